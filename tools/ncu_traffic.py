"""DRAM bytes of one kernel launch from an ncu --set full report -> profiles/r2_traffic.json
(bench.py reports it as roofline.traffic, labelled static).  usage: ncu_traffic.py report.ncu-rep workload-name kernel-name"""
import csv
import json
import subprocess
import sys

rep, workload, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def to_bytes(key):
    v, u = d[key]
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]


print(json.dumps({"workload": workload, "kernel": kernel, "dram_bytes_read": to_bytes("dram__bytes_read.sum"),
                  "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
                  "gpu_time_ms": float(d["gpu__time_duration.sum"][0].replace(",", "")) * {"ms": 1, "us": 1e-3, "s": 1e3, "msecond": 1, "usecond": 1e-3, "second": 1e3}.get(d["gpu__time_duration.sum"][1], 1),
                  "source": "ncu --set full --clock-control none, one launch of the full workload"}, indent=1))
