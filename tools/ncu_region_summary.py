"""Split an `ncu --page source --csv` dump into cull core / rest and summarise stalls."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
num = lambda x: int(x) if x.strip().isdigit() else 0
idx = [k for k, r in enumerate(data) if "FFMA2" in r[ix["Source"]]]
lo, hi = idx[0], idx[-1]
while lo > 0 and "LDS.128" not in data[lo][ix["Source"]]:
    lo -= 1


def summarize(name, rs):
    tot = sum(num(r[ix["# Samples"]]) for r in rs)
    exe = sum(num(r[ix["Instructions Executed"]]) for r in rs)
    thr = sum(num(r[ix["Thread Instructions Executed"]]) for r in rs)
    agg = {s[6:]: sum(num(r[ix[s]]) for r in rs) for s in stalls}
    print(f"{name:12s} sass={len(rs):5d} samples={tot:8d} exec={exe/1e9:6.2f}G thr/inst={thr/max(exe,1):5.1f} ",
          {k: round(v / max(tot, 1), 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]})


print("total samples", sum(num(r[ix["# Samples"]]) for r in data))
summarize("before-cull", data[:lo])
summarize("cull-core", data[lo:hi + 1])
summarize("after-cull", data[hi + 1:])
ops = {}
for r in data[:lo] + data[hi + 1:]:
    src = r[ix["Source"]].split()
    if not src:
        continue
    op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
    o = ops.setdefault(op, [0, 0])
    o[0] += num(r[ix["Instructions Executed"]])
    o[1] += num(r[ix["# Samples"]])
tot_e = sum(v[0] for v in ops.values())
print("non-cull opcodes:", ", ".join(f"{k} {100*v[0]/tot_e:.1f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1][0])[:22]))
