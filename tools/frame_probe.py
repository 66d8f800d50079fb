#!/usr/bin/env python
"""Times one output frame of the 4 M-particle column two ways (cuda:0): a complete spsph_download of the time-varying
arrays a writer needs, and spsph_download_frame (the same columns packed on the device, one transfer per species).
usage: python tools/frame_probe.py [--ncol 1632]"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stress-particle-sph_b200"))
import spsph  # noqa: E402
from spsph import _abi, decks  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ncol", type=int, default=1632)
a = ap.parse_args()
d = tempfile.mkdtemp()
decks.write_deck(d, decks.refined_bui_spec(ncol=a.ncol))
prob = spsph.load(d, "bui")
p = prob.params
eng = spsph.Engine(prob)
eng.run(1, 0.0, prob.blocks[0]["dt"], 3)
import torch  # pinned host memory, as bench.py's e2e leg uses

keys = ("x", "vel", "stress", "internal_vars", "displ", "disp_10", "rho", "hsml", "bc_or_not")
spec = {n: (dt, sh(p)) for n, _, dt, sh in _abi.STATE_FIELDS}
arrays = {}
for k in keys:
    t = torch.empty(int(np.prod(spec[k][1])) * np.dtype(spec[k][0]).itemsize, dtype=torch.uint8).pin_memory()
    arrays[k] = t.numpy().view(spec[k][0]).reshape(spec[k][1])
node_cols = ["x", "y", "vx", "vy", "sxx", "syy", "sxy", "szz", "epsp", "disp_10", "rho", "hsml", "displ_x", "displ_y", "bc_or_not"]
sp_cols = ["x", "y", "vx", "vy", "sxx", "syy", "sxy", "szz", "epsp", "rho", "hsml"]
nn, ns = p.nnode, p.ntotal - p.nnode
tn = torch.empty(nn * len(node_cols), dtype=torch.float64).pin_memory().numpy().reshape(nn, len(node_cols))
ts = torch.empty(ns * len(sp_cols), dtype=torch.float64).pin_memory().numpy().reshape(ns, len(sp_cols))
eng.download(arrays)  # evaluates the free-surface marks once: neither timing below pays for them twice
eng.download_frame(node_cols, 0, nn, tn)
res = {}
for name, fn in (("download", lambda: eng.download(arrays)),
                 ("frame", lambda: (eng.download_frame(node_cols, 0, nn, tn), eng.download_frame(sp_cols, nn, ns, ts)))):
    best = 1e9
    for _ in range(5):
        eng.sync()
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    res[name] = best
b_full = sum(v.nbytes for v in arrays.values())
b_frame = tn.nbytes + ts.nbytes
assert np.array_equal(tn[:, 0], arrays["x"][:nn, 0]) and np.array_equal(ts[:, 8], arrays["internal_vars"][nn:p.ntotal, 0])
print(f"frame of {p.ntotal} particles: spsph_download {b_full / 1e6:.0f} MB in {res['download'] * 1e3:.2f} ms, "
      f"spsph_download_frame {b_frame / 1e6:.0f} MB in {res['frame'] * 1e3:.2f} ms (both incl. the free-surface pass)")
