"""Group ncu source-line samples of render_kernel by functional phase (line ranges are found
from marker comments in rtclj_kernels.cuh, so the script follows the source as it moves)."""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
src = open("raytracing-clj_b200/csrc/rtclj_kernels.cuh").read().splitlines()


def find(pat, start=0):
    for i in range(start, len(src)):
        if pat in src[i]:
            return i + 1
    raise SystemExit("marker not found: " + pat)


marks = [
    ("cull: packed fp32 wrappers", find("typedef unsigned long long f32x2;"), find("// ---------------------------------------------------------------- TMA bulk staging")),
    ("fp64 div/sqrt (noinline)", find("__noinline__ d3 divs("), find("__noinline__ double dsqrt(") + 1),
    ("helpers: philox", find("uint4 philox("), find("double u24(")),
    ("helpers: u24/sym/vec3", find("double u24("), find("unsigned char quantise(")),
    ("helpers: exact_test", find("void exact_test("), find("template <bool kConstTab>")),
    ("staging/init", find("template <bool kConstTab>"), find("// ---- fp32 view of the ray")),
    ("cull setup", find("// ---- fp32 view of the ray"), find("// ---- (A)+(B)")),
    ("cull loop+record", find("// ---- (A)+(B)"), find("// ---- (B) exact closest hit")),
    ("B: survivor walk+prefilter", find("// ---- (B) exact closest hit"), find("// ---- (C) shade")),
    ("C: shade", find("// ---- (C) shade"), find("      if (done) {")),
    ("bookkeeping", find("      if (done) {"), find("// ---- refill:")),
    ("refill/unit decode", find("// ---- refill:"), find("// ---- camera ray:")),
    ("camera ray", find("// ---- camera ray:"), find("// ---- counters:")),
    ("epilogue", find("// ---- counters:"), find("struct FParams")),
]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hk = next(k for k, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hk]
i_s, i_e, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
num = lambda x: int(x) if x.strip().isdigit() else 0
agg = {}
for r in rows[hk + 1:]:
    if len(r) != len(hdr) or not r[0].strip().isdigit():
        continue
    ln = int(r[0])
    name = next((n for n, a, b in marks if a <= ln < b), None)
    if name is None:
        srcline = src[ln - 1] if ln - 1 < len(src) else ""
        name = "fp64 div/sqrt (noinline)" if re.search(r"d3 divs\(|double ddiv\(|double dsqrt\(", srcline) else \
            ("helpers: u24/sym/vec3" if "__forceinline__" in srcline else f"other")
    a = agg.setdefault(name, [0, 0, 0])
    a[0] += num(r[i_s]); a[1] += num(r[i_e]); a[2] += num(r[i_t])
tot = sum(a[0] for a in agg.values()); tote = sum(a[1] for a in agg.values())
print(f"{'group':28s} {'samples%':>8s} {'instr%':>7s} {'thr/inst':>8s} {'cyc/inst*':>9s}")
for n, (s, e, t) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{n:28s} {100*s/tot:8.2f} {100*e/tote:7.2f} {t/max(e,1):8.1f} {(s/tot)/(e/tote) if e else 0:9.2f}")
print("(*relative: share of warp-time over share of instructions)")
