"""Print the handful of ncu raw metrics this project watches (FP32-pipe roofline evidence)."""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "sm__icc_request_hit_rate.pct",
        "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_average_branch_targets_threads_uniform.pct", "l1tex__t_sector_hit_rate.pct"]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
for w in want:
    if w in d:
        print(f"{w:75s} {d[w][0]:>16s} {d[w][1]}")
st = [(float(v), h) for h, u, v in zip(hdr, units, vals) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
for f, h in sorted(st, reverse=True)[:10]:
    print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:24s} {f:6.3f} warps/issue")
