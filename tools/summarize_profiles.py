"""Turns gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum) and gpurun_out/prof_ops.ncu-rep
(ncu --set full) into the tracked summaries under profiles/.   python tools/summarize_profiles.py r01"""
import collections
import csv
import os
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(root, "profiles")
os.makedirs(out_dir, exist_ok=True)


def launches():
    path = os.path.join(root, "gpurun_out", "launches.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    data = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[start + 1:]
            if len(r) > vi and r[mi] == "gpu__time_duration.sum"]
    ups = [i for i, (k, _) in enumerate(data) if "upsample_kernel" in k]      # one per PD-UNet step
    # bench.py --steps 1 --warmup 3 (eager): 3 warm-up steps, then the timed eager step
    a, b = ups[3], ups[4] if len(ups) > 4 else len(data)
    step = data[a:b]
    tot = sum(v for _, v in step)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in step:
        name = re.sub(r"\(.*", "", k)
        name = re.sub(r"^void ", "", name)
        agg[name[:110]][0] += 1
        agg[name[:110]][1] += v
    with open(os.path.join(out_dir, f"{tag}_launches_step.md"), "w") as f:
        f.write(f"# {tag}: one eager PD-UNet inference step (configs[1], batch 16), every launch\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --steps 1 --warmup 3` "
                "(cuDNN autotune off, graph off). Per-launch times are cold-cache and serialised: compare SHARES.\n\n"
                f"launches in the step: {len(step)}; sum of kernel time: {tot / 1e6:.3f} ms\n\n"
                "| share | time (us) | launches | kernel |\n|---:|---:|---:|---|\n")
        for name, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {100 * v / tot:5.1f}% | {v / 1e3:9.1f} | {n} | `{name}` |\n")
        ours = sum(v for k, (n, v) in agg.items() if "pdu::" in k)
        f.write(f"\nlibpdu_b200 kernels: {100 * ours / tot:.1f}% of the step; cuDNN / ATen: {100 - 100 * ours / tot:.1f}%\n")
    print("wrote", f"{tag}_launches_step.md")


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg",
        "smsp__average_warp_latency_issue_stalled_barrier.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def full(rep_name="prof_ops.ncu-rep", suffix=""):
    rep = os.path.join(root, "gpurun_out", rep_name)
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    ki = hdr.index("Kernel Name")
    with open(os.path.join(out_dir, f"{tag}_ncu_full_summary{suffix}.md"), "w") as f:
        f.write(f"# {tag}: ncu --set full, selected counters per captured launch (tools/gpu_final.sh)\n\n"
                "Captured under the profiler (replays, cold caches): use for ratios and stall reasons, not for timing.\n")
        seen = collections.Counter()
        for r in rows[2:]:
            seen[r[ki][:60]] += 1
            if seen[r[ki][:60]] > 2:      # two launches per kernel are enough
                continue
            f.write(f"\n## `{r[ki][:120]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for w, i in idx:
                f.write(f"| {w} | {r[i]} | {units[i]} |\n")
    print("wrote", f"{tag}_ncu_full_summary{suffix}.md")


launches()
full()                                         # forward projector (cell build + strip kernel)
full("prof_adj.ncu-rep", "_adj")
full("prof_fan.ncu-rep", "_fan")
full("prof_filter.ncu-rep", "_filter")
full("prof_nufft.ncu-rep", "_nufft")
