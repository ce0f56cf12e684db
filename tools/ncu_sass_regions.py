"""Summarise an ncu report's SASS page by code REGION: consecutive instructions are grouped by the
marker given on the command line (offset ranges), or, without markers, in windows of N instructions.
Prints samples %, executed warp instructions % and active threads per instruction for each window.
usage: ncu_sass_regions.py report.ncu-rep [window=100]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
win = int(sys.argv[2]) if len(sys.argv) > 2 else 100
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hk = next(k for k, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hk]
i_s, i_e, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
body = [r for r in rows[hk + 1:] if len(r) == len(hdr)]
num = lambda x: int(x) if x.strip().lstrip("-").isdigit() else 0
tot_s = sum(num(r[i_s]) for r in body) or 1
tot_e = sum(num(r[i_e]) for r in body) or 1
tot_t = sum(num(r[i_t]) for r in body)
print(f"instructions {len(body)}  samples {tot_s}  warp-instructions executed {tot_e}  threads/instruction {tot_t / tot_e:.2f}")
print("  idx      samples%  exec%   thr/inst  first instruction")
for k in range(0, len(body), win):
    w = body[k:k + win]
    s = sum(num(r[i_s]) for r in w); e = sum(num(r[i_e]) for r in w); t = sum(num(r[i_t]) for r in w)
    print(f"{k:5d}  {100 * s / tot_s:8.2f} {100 * e / tot_e:7.2f} {t / max(e, 1):9.1f}   {w[0][1].strip()[:60]}")
