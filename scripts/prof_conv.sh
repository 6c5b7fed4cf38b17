#!/bin/sh
# ncu --set full of the ResNet-18 step's own kernels (1 GPU): the conv GEMMs (implicit-GEMM forward / dgrad / wgrad, explicit
# stem + stride-2 + 1x1 convs, split-K), BatchNorm (single-launch reduction, apply, backward), im2col of the stem, pooling,
# split-K reduce.  Skips the first eager step (warm-up), captures 60 matching launches of the second.
#   -> gpurun_out/r2_prof_conv_raw.csv (the report's raw page; the .ncu-rep itself is dropped when it is larger than 20 MB:
#      gpurun brings back at most 64 MiB) + per-launch times of one whole step: r2_convnet_launch_times.csv
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none \
    -k regex:"gemm_tcgen05_kernel|bn_reduce_finalize|bn_apply|bn_bwd|im2col|col2im|maxpool|avgpool|splitk_reduce|softmax_xent" \
    -s 170 -c 60 -f -o gpurun_out/r2_prof_conv python scripts/prof_convnet_only.py 3 > gpurun_out/r2_prof_conv.log 2>&1
echo "ncu conv rc=$?"; tail -n 2 gpurun_out/r2_prof_conv.log
ncu -i gpurun_out/r2_prof_conv.ncu-rep --page raw --csv > gpurun_out/r2_prof_conv_raw.csv 2>/dev/null
[ "$(stat -c %s gpurun_out/r2_prof_conv.ncu-rep 2>/dev/null || echo 0)" -gt 20000000 ] && rm -f gpurun_out/r2_prof_conv.ncu-rep
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 230 -c 260 --csv --log-file gpurun_out/r2_convnet_launch_times.csv \
    python scripts/prof_convnet_only.py 2 > gpurun_out/r2_convnet_launch_times.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2_convnet_launch_times.csv")) if len(r) > 5]
hdr = rows[0] if rows else []
if "Kernel Name" in hdr:
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            agg[r[ki][:70]][0] += 1; agg[r[ki][:70]][1] += float(r[vi].replace(",", ""))
        except ValueError:
            pass
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print("%6.1f us %4d x  %4.1f%%  %s" % (v[1] / 1e3, v[0], 100 * v[1] / tot, k))
    print("total kernel time of the captured launches: %.1f us" % (tot / 1e3))
PY
