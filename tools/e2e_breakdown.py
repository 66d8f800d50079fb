#!/usr/bin/env python
"""Wall-clock split of the e2e leg of bench.py (upload from pinned memory / K steps / download), repeated 3 times."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "stress-particle-sph_b200"))
import bench, spsph
prob = bench.make_problem(1632)
dt = prob.blocks[0]["dt"]
eng = spsph.Engine(prob)
eng.run(1, 0.0, dt, 3)
pinned, keep = bench.pinned_like(dict(prob.arrays))
out_keys = ("x", "vel", "stress", "internal_vars", "displ")
outp, keep2 = bench.pinned_like({k: prob.arrays[k] for k in out_keys})
K = 20
for rep in range(3):
    eng.sync(); t0 = time.perf_counter()
    eng.upload(pinned); t1 = time.perf_counter()
    eng.run(1, 0.0, dt, K); eng.sync(); t2 = time.perf_counter()
    eng.download(outp); t3 = time.perf_counter()
    print(f"rep {rep}: upload {1e3*(t1-t0):.1f} ms, {K} steps {1e3*(t2-t1):.1f} ms, download {1e3*(t3-t2):.1f} ms, total {1e3*(t3-t0):.1f} ms "
          f"-> {prob.params.ntotal*K/(t3-t0):.3e} particle-steps/s; threads {os.cpu_count()} affinity {len(os.sched_getaffinity(0))}")
