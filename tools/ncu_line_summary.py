"""Aggregate `ncu --page source --csv --print-source sass,cuda` by CUDA source line."""
import csv
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hk = next(k for k, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hk]
i_line, i_src, i_addr, i_sass = 0, 1, 2, 3
i_s = hdr.index("# Samples")
i_e = hdr.index("Instructions Executed")
i_t = hdr.index("Thread Instructions Executed")
agg = {}
tot = 0
tote = 0
for r in rows[hk + 1:]:
    if len(r) != len(hdr):
        continue
    num = lambda x: int(x) if x.strip().lstrip("-").isdigit() else 0
    s = num(r[i_s]); e = num(r[i_e]); t = num(r[i_t])
    a = agg.setdefault(r[i_line], [0, 0, 0, r[i_src], 0])
    a[0] += s; a[1] += e; a[2] += t; a[4] += 1
    tot += s; tote += e
print("total samples", tot, "instructions executed", tote)
print("line  samples   %samp   %exec  thr/inst  #sass  source")
for line, (s, e, t, src, k) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    print(f"{line:>4s} {s:8d} {100*s/tot:6.2f}% {100*e/tote:6.2f}% {t/max(e,1):6.1f} {k:5d}  {src.strip()[:100]}")
