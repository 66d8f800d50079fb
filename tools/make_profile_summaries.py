#!/usr/bin/env python
"""Turns the raw ncu output of tools/profile_all.sh (gpurun_out/prof) into the tracked summaries under profiles/:
  r1_launches.csv / r1_launch_summary.txt  -- per-launch durations of one time step
  r1_<kernel>.txt                          -- headline metrics, stall reasons, hottest SASS of one launch
  sweep_b_traffic.json                     -- DRAM bytes per launch of k_sweep_b_{sp,node} (bench.py roofline.traffic)
usage: python tools/make_profile_summaries.py [gpurun_out/prof] [round tag, default r2]"""
import collections
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "prof")
dst = os.path.join(ROOT, "profiles")
TAG = sys.argv[2] if len(sys.argv) > 2 else "r2"
ROUND = "round " + TAG[1:]
sys.path.insert(0, ROOT)


def csrc_hash():
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "stress-particle-sph_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(f.encode())
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()


def launch_summary():
    rows = list(csv.reader(open(os.path.join(src, "launches.csv"))))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    seq = []
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1e6 if r[mu] == "ns" else (v / 1e3 if r[mu] == "us" else v)
        seq.append((re.sub(r"\(.*", "", r[kn]).replace("void ", ""), v))
    idx = [i for i, (n, _) in enumerate(seq) if n.startswith("k_domain_bbox")]
    step = seq[idx[0]:idx[1]] if len(idx) > 1 else seq[idx[0]:] if idx else seq
    agg = collections.OrderedDict()
    for n, v in step:
        a = agg.setdefault(n, [0.0, 0])
        a[0] += v
        a[1] += 1
    tot = sum(v for _, v in step)
    out = [ROUND + " -- ncu launch list of ONE time step of the 4 002 483-particle refined Bui column, 1 B200",
           "command: ncu --metrics gpu__time_duration.sum --clock-control none -s 49 -c 60 --csv python tools/run_steps.py --steps 2",
           "(cold-cache, serialised per-launch times: compare SHARES with bench.py's kernels_ms_per_step, not absolutes)", ""]
    for n, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
        out.append(f"{n:38s} {v:8.3f} ms {c:3d} launches {100 * v / tot:5.1f}%")
    out.append(f"{'total':38s} {tot:8.3f} ms")
    open(os.path.join(dst, TAG + "_launch_summary.txt"), "w").write("\n".join(out) + "\n")
    shutil.copy(os.path.join(src, "launches.csv"), os.path.join(dst, TAG + "_launches.csv"))


def kernel_summaries():
    traffic = {}
    for f in sorted(os.listdir(src)):
        if not f.endswith(".ncu-rep"):
            continue
        name = f[:-8]
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(src, f), "24"],
                             capture_output=True, text=True).stdout
        txt += "\ninstruction mix (tools/ncu_opmix.py):\n" + subprocess.run(
            [sys.executable, os.path.join(ROOT, "tools", "ncu_opmix.py"), os.path.join(src, f)], capture_output=True,
            text=True).stdout
        head = (f"{ROUND} -- ncu --set full --clock-control none --import-source on, one launch of {name} in step 2 of the\n"
                f"4 002 483-particle refined Bui column (tools/profile_all.sh); read with tools/ncu_summary.py\n\n")
        open(os.path.join(dst, f"{TAG}_{name}.txt"), "w").write(head + "\n".join(l[:200] for l in txt.splitlines()) + "\n")
        raw = subprocess.run(["ncu", "-i", os.path.join(src, f), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        h, u, v = rows[0], rows[1], rows[2]

        def val(metric):
            i = h.index(metric)
            x = float(v[i].replace(",", ""))
            return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u[i]]
        traffic[name] = {"dram_read": val("dram__bytes_read.sum"), "dram_write": val("dram__bytes_write.sum"),
                         "kernel": v[h.index("Kernel Name")][:60]}
    path = os.path.join(dst, "sweep_b_traffic.json")
    try:  # captures of the same kernel sources taken in another call: keep them
        old = json.load(open(path))
        if old.get("csrc_sha256") == csrc_hash():
            traffic = {**old.get("kernels", {}), **traffic}
    except Exception:
        pass
    b = [traffic[k] for k in ("k_sweep_b_sp", "k_sweep_b_node", "k_artvisc") if k in traffic]
    if len(b) == 3:
        per_launch = sum(t["dram_read"] + t["dram_write"] for t in b)  # one stage
        import bench  # SASS hashes of the captured kernels in the library built from these sources (bench.py compares)
        sass = {m: v for m, v in bench.kernel_sass_hashes().items() if any(f"{len(n)}{n}" in m for n in traffic)}
        json.dump({"dram_bytes_per_launch": per_launch, "csrc_sha256": csrc_hash(), "round": TAG,
                   "kernel_sass_sha256": sass,
                   "sass_tie": "SASS hashes of every instantiation of the captured kernels in the library built from "
                               "csrc_sha256 (the build the capture ran)",
                   "note": "dram__bytes_read.sum + dram__bytes_write.sum of one k_sweep_b_sp + one k_sweep_b_node + one "
                           "k_artvisc launch (= one RK stage of sweep B, the unit bench.py times as 'k_sweep_b'), "
                           "ncu --set full, 4 002 483-particle refined Bui column",
                   "kernels": traffic}, open(path, "w"), indent=1)


os.makedirs(dst, exist_ok=True)
if os.path.exists(os.path.join(src, "launches.csv")):
    launch_summary()
    print(open(os.path.join(dst, TAG + "_launch_summary.txt")).read())
kernel_summaries()
if os.path.exists(os.path.join(src, "profile.log")):
    shutil.copy(os.path.join(src, "profile.log"), os.path.join(dst, TAG + "_event_profile.txt"))
for mode in ("list", "tile"):  # per-kernel metric tables of one step on either path (tools/gpu_metrics_pass.sh)
    f = os.path.join(src, f"metrics_{mode}.csv")
    if os.path.exists(f):
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "metrics_table.py"), f], capture_output=True,
                             text=True).stdout
        which = "id-list path (default)" if mode == "list" else "cell-tile path (SPSPH_TILE=1)"
        head = (f"{ROUND} -- one ncu metrics pass over every kernel of one time step, {which},\n"
                "4 002 483-particle refined Bui column, 1 B200 (tools/gpu_metrics_pass.sh; serialised, cold-cache launches).\n"
                "Minst/l = warp instructions per launch, issue% = issue slots active, warps% = resident warps of 64 per SM,\n"
                "fp64% = fp64 pipe active, GB/l = DRAM bytes read + written per launch, GB/s = that over the launch time.\n\n")
        open(os.path.join(dst, f"{TAG}_kernel_metrics_{mode}.txt"), "w").write(head + txt)
