"""Key counters of every kernel in an .ncu-rep as a markdown table (reads the report here, no GPU).
   python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/r02_x.md"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_registers", "occ regs"),
        ("launch__occupancy_limit_shared_mem", "occ smem"),
        ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/smem %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__inst_executed.sum", "warp instr"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank conflicts"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio")]
print(f"`ncu --set full --clock-control none` ({rep.split('/')[-1]}); per launch, cold caches (ncu flushes between kernels).\n")
for r in data:
    if len(r) != len(hdr):
        continue
    print(f"### `{r[ix['Kernel Name']][:110]}`\n")
    print("| counter | value |\n|---|---:|")
    for key, label in cols:
        if key in ix:
            u = units[ix[key]]
            print(f"| {label} (`{key}`) | {r[ix[key]]} {u} |")
    print()
