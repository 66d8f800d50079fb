#!/usr/bin/env python
"""per-kernel table from an `ncu --metrics ... --csv` log (tools/gpu_metrics_pass.sh)"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if not hdr or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    n = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "")
    a = agg.setdefault((n, d["ID"]), {})
    v = float(d["Metric Value"].replace(",", ""))
    u = d["Metric Unit"]
    if d["Metric Name"] == "gpu__time_duration.sum":
        v = v / 1e6 if u == "ns" else (v / 1e3 if u == "us" else v)
    if d["Metric Name"].startswith("dram__bytes"):
        v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    a[d["Metric Name"]] = v
per = collections.OrderedDict()
for (n, _), a in agg.items():
    p = per.setdefault(n, dict(n=0, ms=0.0, inst=0.0, issue=0.0, rd=0.0, wr=0.0, warps=0.0, fp64=0.0, regs=0))
    p["n"] += 1
    p["ms"] += a.get("gpu__time_duration.sum", 0)
    p["inst"] += a.get("smsp__inst_executed.sum", 0)
    p["issue"] += a.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0)
    p["rd"] += a.get("dram__bytes_read.sum", 0)
    p["wr"] += a.get("dram__bytes_write.sum", 0)
    p["warps"] += a.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0)
    p["fp64"] += a.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 0)
    p["regs"] = a.get("launch__registers_per_thread", 0)
tot = sum(p["ms"] for p in per.values())
print(f"{'kernel':34s} {'n':>3s} {'ms':>8s} {'ms/l':>7s} {'%':>5s} {'Minst/l':>8s} {'issue%':>6s} {'warps%':>6s} {'fp64%':>6s} {'GB/l':>6s} {'GB/s':>6s} regs")
for n, p in sorted(per.items(), key=lambda x: -x[1]["ms"]):
    k = p["n"]
    gb = (p["rd"] + p["wr"]) / k / 1e9
    print(f"{n[:34]:34s} {k:3d} {p['ms']:8.3f} {p['ms']/k:7.3f} {100*p['ms']/tot:5.1f} {p['inst']/k/1e6:8.1f} {p['issue']/k:6.1f} "
          f"{p['warps']/k:6.1f} {p['fp64']/k:6.1f} {gb:6.3f} {gb/(p['ms']/k/1e3) if p['ms'] else 0:6.0f} {int(p['regs'])}")
print(f"total {tot:.3f} ms, {sum(p['inst'] for p in per.values())/1e6:.0f} M warp instructions, "
      f"{sum(p['rd']+p['wr'] for p in per.values())/1e9:.2f} GB DRAM")
