"""Digest the round-2 GPU calls (gpurun_out/ is scratch) into the tracked evidence under profiles/:

    python scripts/summarize_round2.py

* profiles/r2_convnet_schedules.json      ResNet-18 step time under every schedule switch (calls 2, 3)
* profiles/r2_cfg5_operand_modes.json     wide-MLP round at N=1 per operand mode / PDL (calls 2, 3, 4)
* profiles/r2_mlp_variants.json           persistent MLP kernel, us per batch-1 step per variant and call
* profiles/r2_ncu_mlp_v5_v6.json (+ raw CSVs)  ncu --set full of the headline kernel, default and packed variant
* profiles/r2_splitk_chaos_control.json   why a 2-step comparison of ResNet schedules measures chaos, not kernels
* copies: first run of the formerly gated tests, GPU suite logs, sanitizer logs, bench lines, experiment logs
"""
import csv
import glob
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def last_json(path):
    try:
        lines = [ln for ln in open(path).read().splitlines() if ln.startswith("{")]
        return json.loads(lines[-1]) if lines else None
    except (OSError, ValueError):
        return None


def dump(name, obj):
    with open(os.path.join(P, name), "w") as f:
        json.dump(obj, f, indent=1)
    print("wrote", name)


def copy(src, dst):
    s = os.path.join(G, src)
    if os.path.exists(s):
        shutil.copyfile(s, os.path.join(P, dst))
        print("copied", dst)


def convnet():
    rows = []
    for pat, call in (("r2_convnet_*.json", 2), ("r2c3_convnet_*.json", 3)):
        for f in sorted(glob.glob(os.path.join(G, pat))):
            d = last_json(f)
            if not d:
                continue
            rows.append({"call": call, "file": os.path.basename(f), "flags": d.get("flags"),
                         "defaults_in_effect": "round-1 schedule (all switches 0)" if call == 2 else "round-2 defaults (STREAMS FUSED_BN SPLITK=1 WGRAD_MN DGRAD_KN IMPLICIT=2)",
                         "graph_ms_per_step": (d.get("native_graph") or {}).get("ms_per_step"),
                         "eager_ms_per_step": (d.get("native_eager") or {}).get("ms_per_step"),
                         "graph_nodes_or_launches_per_fit": (d.get("native_eager") or {}).get("launches_per_fit"),
                         "last_loss": (d.get("native_graph") or {}).get("last_loss"),
                         "cudnn_autocast_ms_per_step": (d.get("torch_cudnn_autocast") or {}).get("ms_per_step")})
    rows.sort(key=lambda r: (r["call"], r["graph_ms_per_step"] or 1e9))
    dump("r2_convnet_schedules.json", {"what": "scripts/bench_convnet.py: 16 SGD steps of batch 128 (32x32 images), best of 3 after 2 warm-ups, CUDA events",
                                      "rows": rows})


def cfg5():
    rows = []
    for f in sorted(glob.glob(os.path.join(G, "r2_bench_cfg5_n1*.json")) + glob.glob(os.path.join(G, "r2c3_bench_cfg5_*.json"))
                    + glob.glob(os.path.join(G, "r2c4_bench_cfg5_*.json"))):
        d = last_json(f)
        if not d:
            rows.append({"file": os.path.basename(f), "error": "no JSON line (the overlapped-reduce run trapped: see profiles/README.md)"})
            continue
        rows.append({"file": os.path.basename(f), "switches": d["config"].get("switches"), "rounds_per_s": d["value"], "ms_per_round": d["ms_per_step"],
                     "e2e_rounds_per_s": (d.get("e2e") or {}).get("value"), "train_path": d["config"].get("train_path"),
                     "roofline": d["config"].get("roofline")})
    dump("r2_cfg5_operand_modes.json", {"what": "bench.py --config cfg5 at N=1 (wide MLP 10-4096x4-2, batch 1024, 8 rounds steps): operand modes and PDL", "rows": rows})


def mlp():
    out = {}
    for f in sorted(glob.glob(os.path.join(G, "r2*_microbench_mlp.json")) + glob.glob(os.path.join(G, "r[A-Z]_microbench_mlp.json"))):
        try:
            d = json.load(open(f))
        except ValueError:
            continue
        res = d.get("results", d) if isinstance(d, dict) else d
        rows = [r for r in res if isinstance(r, dict) and r.get("kernel") == "mlp_local_sgd_persistent"]
        out[os.path.basename(f)] = [{k: r.get(k) for k in ("variant", "net", "batch", "samples", "clients", "us_per_step")} for r in rows]
    dump("r2_mlp_variants.json", {"what": "scripts/microbench.py --only mlp: one client, CUDA events, us per SGD step (batch 1) or per batch", "calls": out})


NCU_KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "smsp__average_warp_latency_per_inst_issued.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def ncu(rep, steps=None):
    path = os.path.join(G, rep)
    if not os.path.exists(path):
        return None
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return None
    with open(os.path.join(P, rep.replace(".ncu-rep", "_ncu_raw.csv")), "w") as f:
        f.write(raw)
    out = []
    hdr = rows[0]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        rec = {"kernel": d.get("Kernel Name"), "grid": d.get("Grid Size"), "block": d.get("Block Size")}
        for k in NCU_KEYS:
            if k in d:
                try:
                    rec[k] = float(d[k].replace(",", ""))
                except ValueError:
                    rec[k] = d[k]
        if steps and "sm__cycles_elapsed.max" in rec:
            rec["cycles_per_sgd_step"] = rec["sm__cycles_elapsed.max"] / steps
            rec["warp_instructions_per_step_and_warp"] = rec.get("smsp__inst_executed.sum", 0) / steps / 4
        out.append(rec)
    return out


def main():
    os.makedirs(P, exist_ok=True)
    convnet()
    cfg5()
    mlp()
    n = {"v5_default_for_ffnn": ncu("r2c3_prof_mlp_v5.ncu-rep", 8192), "v6_packed_ffma2_default_for_mlp64": ncu("r2c3_prof_mlp_v6.ncu-rep", 8192)}
    if any(n.values()):
        dump("r2_ncu_mlp_v5_v6.json", {"what": "ncu --set full --clock-control none of mlp_local_sgd_kernel_v2<MLP 10-64-64-2>, 8192 batch-1 steps, 1 CTA x 128 threads "
                                               "(call 3; the kernels of that call still gathered through a caller-made permutation table)", **n})
    for rep in sorted(glob.glob(os.path.join(G, "r2c*_prof_*.ncu-rep"))):
        name = os.path.basename(rep)
        if "prof_mlp" in name:
            continue
        r = ncu(name)
        if r:
            dump(name.replace(".ncu-rep", "_ncu_summary.json"), {"kernels": r})
    d = last_json(os.path.join(G, "r2c3_debug_splitk.json"))
    if d is None:
        try:
            d = json.load(open(os.path.join(G, "r2c3_debug_splitk.json")))
        except (OSError, ValueError):
            d = None
    if d:
        dump("r2_splitk_chaos_control.json", d)
    copy("r2_unvalidated_tests.log", "r2_call1_first_run_of_the_gated_tests.log")
    for src in sorted(glob.glob(os.path.join(G, "r2c*_pytest_*.log"))):
        copy(os.path.basename(src), os.path.basename(src).replace("r2c", "r2_call"))
    for src in sorted(glob.glob(os.path.join(G, "sanitizer_*.log"))):
        copy(os.path.basename(src), "r2_" + os.path.basename(src))
    for src in sorted(glob.glob(os.path.join(G, "r2c*_bench_*.json")) + glob.glob(os.path.join(G, "r2_8_*.json")) + glob.glob(os.path.join(G, "r2_bench_reference*.json"))):
        if os.path.getsize(src) > 0 and "cfg5_x" not in src:
            copy(os.path.basename(src), os.path.basename(src).replace("r2c", "r2_call"))
    for src in sorted(glob.glob(os.path.join(G, "r2_8_sweep_*.json")) + glob.glob(os.path.join(G, "r2_8_tests.log")) + glob.glob(os.path.join(G, "r2_8_box_cfg3*"))
                      + glob.glob(os.path.join(G, "r2c7_box*.jsonl")) + glob.glob(os.path.join(G, "r2c7_box.log"))):
        copy(os.path.basename(src), os.path.basename(src).replace("r2c", "r2_call"))
    exp = os.path.join(G, "experiment_logs")
    if os.path.isdir(exp):
        dst = os.path.join(P, "r2_experiment_logs_b200")
        shutil.rmtree(dst, ignore_errors=True)
        shutil.copytree(exp, dst, ignore=shutil.ignore_patterns("pi*_*.log", "coordinator_*.log"))
        print("copied experiment logs")


LATE_CALLS = "ABCDEFGHIJKLM"


def late_calls():
    """Calls A .. M (letters): bench lines, test logs, NVLink counters, round-overhead fits, ncu captures of the final kernels."""
    for src in sorted(glob.glob(os.path.join(G, "r[A-Z]_*"))):
        name = os.path.basename(src)
        if name.endswith((".ncu-rep", ".pth", ".err")) or os.path.getsize(src) == 0 or os.path.getsize(src) > 3_000_000:
            continue
        copy(name, "r2_call" + name[1:])
    copy("r2_prof_conv_raw.csv", "r2_callB_prof_conv_ncu_raw.csv")
    copy("r2_convnet_launch_times.csv", "r2_convnet_launch_times.csv")
    # NVLink payload counters (NVML) around the fused collectives
    nv = {"what": "NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX / RX (payload bytes, all links of a GPU) read before and after an untimed loop of "
                  "kernel launches: what the hardware says crossed the links, next to the algorithmic count and the CUDA-event time",
          "twoshot_kernel_world2": {}, "engine_rounds": {}}
    for tag, f in (("nvls", "rC_nvlink_twoshot_nvls.json"), ("p2p", "rC_nvlink_twoshot_p2p.json")):
        try:
            nv["twoshot_kernel_world2"][tag] = [{k: r.get(k) for k in ("P", "bytes", "world", "nvls", "twoshot_ms", "twoshot_busbw_GBps", "nccl_allreduce_ms",
                                                                         "nccl_busbw_GBps", "nvlink")} for r in json.load(open(os.path.join(G, f)))]
        except (OSError, ValueError):
            pass
    for tag, f in (("cfg2_star_n2_before_staged_writeout", "rC_bench_cfg2_n2_nvlink.json"), ("cfg5_twoshot_nvls_n2", "rC_bench_cfg5_n2_nvlink.json"),
                   ("ref_local_star_n2_staged_writeout", "rF_bench_default_n2.json"), ("cfg5_twoshot_p2p_n2", "rE_bench_cfg5_n2_nvlink.json"),
                   ("ref_local_star_n8", "rM_default_n8.json"), ("cfg5_twoshot_nvls_n8", "rM_cfg5_n8.json")):
        d = last_json(os.path.join(G, f))
        if d:
            nv["engine_rounds"][tag] = {"rounds_per_s": d["value"], "nvls": d["config"].get("nvls"), **(d["config"].get("nvlink") or {})}
    dump("r2_nvlink_counters.json", nv)
    # fixed cost of a round: fits of (round time) over (steps per round)
    ov = {"what": "scripts/round_overhead.py: FFNN, SSE, batch 1; least-squares line through (steps per round, ms per round): slope = us per step, "
                  "intercept = fixed cost of a round; K rounds in one run_rounds call / K single-round calls; phases = the engine's event timers",
          "runs": {}}
    for tag, f in (("n1_callD", "rD_round_overhead_ffnn_n1.json"), ("n2_callE_before_stream_alignment", "rE_round_overhead_ffnn_n2.json"),
                   ("n2_callF", "rF_round_overhead_ffnn_n2.json"), ("n8_callM_final_kernel", "rM_round_overhead_ffnn_n8.json")):
        try:
            ov["runs"][tag] = json.load(open(os.path.join(G, f)))
        except (OSError, ValueError):
            pass
    dump("r2_round_overhead.json", ov)
    # ncu --set full of the headline kernel at three points of the latency work
    n = {"callB_before (FFNN v5, BCE / MLP v6)": {"ffnn": ncu("rB_prof_ffnn_default.ncu-rep", 8192), "mlp64": ncu("rB_prof_mlp64_default.ncu-rep", 8192)},
         "callI_after_fast_sigmoid_and_uniform_loss (FFNN v5, SSE / MLP v6)": {"ffnn": ncu("rI_prof_ffnn_sse.ncu-rep", 8192), "mlp64": ncu("rI_prof_mlp64.ncu-rep", 8192)}}
    dump("r2_ncu_mlp_late.json", {"what": "ncu --set full --clock-control none --import-source on, 8 192 batch-1 steps, 1 CTA x 128 threads; per-instruction stall "
                                          "listings: python scripts/ncu_hot_loop.py <rep> (r2_callI_hot_loop_*.txt)", **n})
    for rep, out in (("rI_prof_ffnn_sse.ncu-rep", "r2_callI_hot_loop_ffnn_sse.txt"), ("rI_prof_mlp64.ncu-rep", "r2_callI_hot_loop_mlp64.txt"),
                     ("rB_prof_mlp64_default.ncu-rep", "r2_callB_hot_loop_mlp64.txt")):
        if os.path.exists(os.path.join(G, rep)):
            txt = subprocess.run(["python", os.path.join(ROOT, "scripts", "ncu_hot_loop.py"), os.path.join(G, rep)], capture_output=True, text=True).stdout
            with open(os.path.join(P, out), "w") as f:
                f.write(txt)
            print("wrote", out)


if __name__ == "__main__":
    main()
    late_calls()
