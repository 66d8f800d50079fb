#!/usr/bin/env python
"""Compares the dumps written by tools/make_reference_goldens.sh (real gfortran build of the reference) with the
oracle: prints the relative L-inf difference of x, vel, stress, eps_p after steps 1, 10, 100.
usage: python tools/compare_reference_dump.py oracle/_ref/<example dir>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stress-particle-sph_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import spsph  # noqa: E402
from oracle_binding import Oracle  # noqa: E402

d = sys.argv[1]
kind = "bui" if "bui" in d else ("vs" if "vertical" in d else "sl")
prob = spsph.load(d, kind)
orc = Oracle(prob)
dt, t, done = prob.blocks[0]["dt"], 0.0, 0
for step in (1, 10, 100):
    t = orc.run(done + 1, t, dt, step - done)
    done = step
    raw = np.fromfile(os.path.join(d, f"dump.{step:06d}.bin"), dtype=np.uint8)
    nt, n2 = np.frombuffer(raw[:8], np.int32)
    body = np.frombuffer(raw[8:], np.float64)
    x, v, s, e = np.split(body, [2 * nt, 4 * nt, 8 * nt])
    a = orc.download()
    for name, ref, got in (("x", x, a["x"][:nt].ravel()), ("vel", v, a["vel"][:nt].ravel()),
                           ("stress", s, a["stress"][:nt].ravel()), ("eps_p", e, a["internal_vars"][:nt, 0])):
        scale = max(np.abs(ref).max(), 1e-300)
        print(f"step {step:4d} {name:7s} rel L-inf {np.abs(ref - got).max() / scale:.3e}  bitwise equal: {np.array_equal(ref, got)}")
