#!/usr/bin/env python3
"""Per-kernel summary (launches, share of kernel time, average duration, DRAM bytes per launch) of an ncu launch list:
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file X <cmd>
usage: launch_summary.py X.csv "<command>" > profiles/rNN_launch_summary_<workload>.json"""
import collections, csv, json, sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14]
hdr = rows[0]
iN, iM, iU, iV, iI = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "second": 1e3, "s": 1e3,
         "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}
per = collections.defaultdict(lambda: collections.defaultdict(dict))
for r in rows[1:]:
    name = r[iN].replace("void ", "").split("(")[0][:60]
    per[name][r[iI]][r[iM]] = float(r[iV].replace(",", "")) * scale[r[iU]]
tot = sum(m.get("gpu__time_duration.sum", 0.0) for k in per.values() for m in k.values())
out = {"command": sys.argv[2] if len(sys.argv) > 2 else "", "kernels": {}}
for name, launches in sorted(per.items(), key=lambda kv: -sum(m.get("gpu__time_duration.sum", 0.0) for m in kv[1].values())):
    t = [m.get("gpu__time_duration.sum", 0.0) for m in launches.values()]
    out["kernels"][name] = {"launches": len(t), "share_of_kernel_time": sum(t) / tot, "avg_ms": sum(t) / len(t),
                            "dram_read_GB": sum(m.get("dram__bytes_read.sum", 0.0) for m in launches.values()) / len(t),
                            "dram_write_GB": sum(m.get("dram__bytes_write.sum", 0.0) for m in launches.values()) / len(t)}
print(json.dumps(out, indent=1))
