"""Per SOURCE LINE summary of an ncu report taken with --import-source on: for every file of the kernel, the
lines with the most executed warp instructions -- share of the kernel's instructions, share of its stall
samples, active threads per instruction.  (ncu's combined sass,cuda page has one section per source file.)
usage: python tools/ncu_source_lines.py report.ncu-rep [top=60]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, lines = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and cur and len(r) == len(hdr) and r[0].strip().isdigit() and r[2] == "-":
        num = lambda x: int(x) if x.strip().isdigit() else 0
        lines.append((cur, int(r[0]), r[1].strip(), num(r[hdr.index("# Samples")]), num(r[hdr.index("Instructions Executed")]),
                      num(r[hdr.index("Thread Instructions Executed")])))
tot_s = sum(l[3] for l in lines) or 1
tot_e = sum(l[4] for l in lines) or 1
print(f"source lines {len(lines)}  warp instructions {tot_e}  samples {tot_s}  threads/instruction {sum(l[5] for l in lines) / tot_e:.2f}")
print("  instr%  samp%  thr/inst  file:line  source")
for f, ln, src, s, e, t in sorted(lines, key=lambda l: -l[4])[:top]:
    print(f"  {100 * e / tot_e:6.2f} {100 * s / tot_s:6.2f} {t / max(e, 1):8.1f}  {f}:{ln}  {src[:110]}")
