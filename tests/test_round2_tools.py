"""The tools the next GPU round relies on must work on first use: scripts/pick_schedule.py on a synthetic gpurun_out/ and the
staged round-2 shell scripts (syntax, stage selection)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench(flags, graph_ms, eager_ms, loss=2.3):
    return json.dumps({"steps_per_fit": 16, "batch": 128, "flags": flags,
                       "native_eager": {"ms_per_fit": eager_ms * 16, "ms_per_step": eager_ms, "launches_per_fit": 9000, "last_loss": loss},
                       "native_graph": {"ms_per_fit": graph_ms * 16, "ms_per_step": graph_ms, "launches_per_fit": 40, "last_loss": loss}})


def test_pick_schedule_digests_a_synthetic_run(tmp_path):
    out = tmp_path / "gpurun_out"
    out.mkdir()
    (out / "r2_unvalidated_tests.log").write_text(
        "=== group: resnet_eval\n1 passed in 20.1s\nrc=0 (resnet_eval)\n"
        "=== group: mn_major_operands or mn_major_b_operand or mn_major_fused\n"
        "FAILED tests/test_zz_round2_gpu.py::test_gemm_mn_major_b_operand[128-64-64-128] - AssertionError\n"
        "1 failed, 5 passed in 31.0s\nrc=1 (mn_major)\n"
        "=== group: implicit_step\nrc=124 (implicit_step)\n")
    (out / "r2_convnet_default.json").write_text(_bench({}, 1.77, 3.9) + "\n")
    (out / "r2_convnet_FUSED_BN_1.json").write_text("some warning line\n" + _bench({"COLEARN_CONV_FUSED_BN": "1"}, 1.61, 3.5) + "\n")
    (out / "r2_convnet_IMPLICIT_2_COLEARN_PDL_1.json").write_text(
        _bench({"COLEARN_CONV_IMPLICIT": "2", "COLEARN_CONV_FUSED_BN": "1", "COLEARN_PDL": "1"}, 1.18, 2.9) + "\n")
    (out / "r2_convnet_SPLITK_2.json").write_text("")                          # crashed run: empty file
    p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "pick_schedule.py"), str(out)], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0, p.stderr
    text = p.stdout
    assert "FAILED tests/test_zz_round2_gpu.py::test_gemm_mn_major_b_operand" in text
    assert "no summary (timeout / crash?)" in text and "unreadable" in text
    order = [ln for ln in text.splitlines() if ln.strip().startswith("r2_convnet_")]
    assert order[0].strip().startswith("r2_convnet_IMPLICIT_2_COLEARN_PDL_1.json") and "1.180" in order[0]
    tail = text[text.index("== fastest schedule"):]
    assert '"IMPLICIT": 2' in tail and '"FUSED_BN": 1' in tail and "COLEARN_PDL=1" in tail


def test_round2_scripts_parse_and_select_stages(tmp_path):
    for name in ("run_round2_first.sh", "run_round2_8gpu.sh"):
        assert subprocess.run(["sh", "-n", os.path.join(ROOT, "scripts", name)], capture_output=True).returncode == 0
    # an unknown stage runs nothing (and needs no GPU): only gpurun_out/ is created
    p = subprocess.run(["sh", os.path.join(ROOT, "scripts", "run_round2_first.sh"), "none"], cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert p.returncode == 0 and p.stdout == "" and os.listdir(tmp_path) == ["gpurun_out"] and os.listdir(tmp_path / "gpurun_out") == []
    text = open(os.path.join(ROOT, "scripts", "run_round2_first.sh")).read()
    for stage in ("tests", "conv", "prof", "gemm"):
        assert f"if want {stage}; then" in text


def test_script_groups_cover_every_gated_gpu_test():
    """scripts/run_round2_first.sh runs the gated GPU tests in per-group pytest processes (-k expressions of the form
    "a or b"): every term must select something and every test of tests/test_zz_round2_gpu.py must be selected."""
    import ast
    import re
    text = open(os.path.join(ROOT, "scripts", "run_round2_first.sh")).read()
    groups = re.findall(r'"([^"]+)"', re.search(r"for grp in (.*?); do", text, flags=re.S).group(1))
    terms = [t.strip() for g in groups for t in g.split(" or ")]
    tree = ast.parse(open(os.path.join(ROOT, "tests", "test_zz_round2_gpu.py")).read())
    tests = [n.name for n in tree.body if isinstance(n, ast.FunctionDef) and n.name.startswith("test_")]
    assert len(groups) >= 14 and len(tests) >= 20
    for t in terms:
        assert any(t in name for name in tests), f"group term {t!r} selects nothing"
    for name in tests:
        assert any(t in name for t in terms), f"{name} is in no group of scripts/run_round2_first.sh"
