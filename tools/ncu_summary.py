"""Headline metrics + stall reasons of one kernel launch in an ncu report, as the text summaries under profiles/ are written.
usage: ncu_summary.py report.ncu-rep "title line" > profiles/<name>_summary.txt"""
import csv, subprocess, sys

rep, title = sys.argv[1:3]
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout.splitlines()))
h, units, r = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__icc_request_hit_rate.pct"]
print(title)
print()
for w in want:
    if w in h:
        i = h.index(w)
        print("%s = %s %s" % (w, r[i], units[i]))
stalls = []
for i, name in enumerate(h):
    if name.startswith("smsp__average_warps_issue_stalled_") and name.endswith("_per_issue_active.ratio") or (name.startswith("smsp__average_warp_latency_issue_stalled_") and name.endswith(".ratio")):
        try:
            stalls.append((float(r[i]), name.split("stalled_")[1].split("_per_issue")[0].replace(".ratio", "")))
        except ValueError:
            pass
stalls.sort(reverse=True)
if stalls:
    print("stall reasons (warps per issue-active cycle): " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:10]))
