#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove which hardware paths a kernel uses (no GPU needed: cuobjdump
reads the in-tree .so).  tcgen05.mma -> UTCHMMA, TMA -> UTMALDG (.2D / .4D / .MULTICAST / .2CTA), tcgen05.ld -> LDTM,
tcgen05.commit -> UTCBAR, mbarrier -> SYNCS, TMEM alloc -> UTCATOMSWS, multimem.ld_reduce -> LDGMC, .sys-scope
acquire/release traffic -> LDG/STG ... STRONG.SYS.

    python scripts/sass_evidence.py > profiles/sass_evidence.txt
"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sorted(glob.glob(os.path.join(ROOT, "colearn_federated_learning_b200", "ops", "_colearn_C*.so")))[0]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.splitlines()
names = iter(demangle)
pat = re.compile(r"\b(UTCHMMA[.\w]*|UTMALDG[.\w]*|LDTM[.\w]*|UTCBAR[.\w]*|UTCATOMSWS[.\w]*|SYNCS[.\w]*|LDGMC[.\w]*|UCGABAR\w*|"
                 r"(?:LDG|STG|ATOMG|RED)\.[.\w]*STRONG\.SYS[.\w]*|MULTIMEM[.\w]*)")
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    if "Function :" in line:
        cur = next(names)
        cur = re.sub(r"\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"\(CUtensorMap_st.*", "", cur)            # drop the parameter list
        counts[cur] = collections.Counter()
        continue
    m = pat.search(line)
    if m and cur is not None:
        counts[cur][m.group(1)] += 1
print(f"# {os.path.basename(so)} — mnemonic counts per kernel (scripts/sass_evidence.py)")
for k, c in counts.items():
    if not c:
        continue
    print(f"== {k}")
    for mn, n in c.most_common():
        print(f"{n:7d} {mn}")
