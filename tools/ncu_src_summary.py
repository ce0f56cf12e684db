"""Summarise an `ncu --page source --csv` dump: samples by region and the hottest SASS lines."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
texe = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
print("instructions:", len(data), "samples:", tot, "inst executed:", texe)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
print("stalls:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
# opcode histogram weighted by executed count
ops = {}
for r in data:
    op = r[ix["Source"]].split()[0] if r[ix["Source"]] else "?"
    if op.startswith("@"):
        op = r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    e = int(r[ix["Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    o = ops.setdefault(op, [0, 0])
    o[0] += e
    o[1] += s
print("by opcode (executed%, samples%):")
for op, (e, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"  {op:10s} {100*e/texe:6.2f}% {100*s/tot:6.2f}%")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("hottest lines:")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:n]:
    st = {s[6:]: int(r[ix[s]] or 0) for s in stalls if int(r[ix[s]] or 0)}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"  {r[ix['Address']][-5:]} {int(r[ix['# Samples']]):6d} exe={int(r[ix['Instructions Executed']] or 0):>10d} thr={r[ix['Avg. Threads Executed']][:5]:>5s} {r[ix['Source']][:70]:70s} {top}")
