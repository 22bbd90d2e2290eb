"""Key metrics of an .ncu-rep (one row per captured launch): python profiles/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
KEYS = ["Kernel Name", "launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__average_warp", "smsp__warp_issue_stalled", "smsp__average_warps_issue_stalled"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h, u = r[0], r[1]
for row in r[2:]:
    print("=" * 100)
    items = []
    for a, b, c in zip(h, u, row):
        if any(a.startswith(k) or a == k for k in KEYS) or (("pipe_tensor" in a or "pipe_tc" in a or "tmem" in a) and "pct" in a):
            if "stalled" in a or "average_warp" in a:
                try:
                    if float(c.replace(",", "")) < 0.3: continue
                except ValueError: pass
            items.append((a, b, c))
    for a, b, c in items:
        print(f"{a:110s} {b:12s} {c}")
