#!/usr/bin/env python
"""Runs n time steps of a generated deck on cuda:0 (profiling driver for ncu).
usage: python tools/run_steps.py [--kind refined_bui|bui|vs|sl|wide_slope] [--ncol 1632] [--steps 3]"""
import argparse
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stress-particle-sph_b200"))
import spsph  # noqa: E402
from spsph import decks  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="refined_bui")
ap.add_argument("--ncol", type=int, default=1632)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--profile", action="store_true")
ap.add_argument("--warmup", type=int, default=0)
ap.add_argument("--deck", default=None, help="directory to write the deck into, reused when it already holds one")
a = ap.parse_args()
d = a.deck or tempfile.mkdtemp()
reuse = bool(a.deck) and os.path.isdir(d) and len(os.listdir(d)) > 0
os.makedirs(d, exist_ok=True)
if a.kind == "refined_bui":
    spec, var = decks.refined_bui_spec(ncol=a.ncol), "bui"
elif a.kind == "wide_slope":
    spec, var = decks.wide_slope_spec(ncol=a.ncol), "vs"
else:
    spec, var = decks.SHIPPED[a.kind](), a.kind
if not reuse:
    decks.write_deck(d, spec)
prob = spsph.load(d, var)
eng = spsph.Engine(prob)
dt = prob.blocks[0]["dt"]
t0 = eng.run(1, 0.0, dt, a.warmup) if a.warmup else 0.0
if a.profile:
    eng.profile(True)
eng.run(1 + a.warmup, t0, dt, a.steps)
ms, n = eng.last_run()
print(f"{a.kind} ncol={a.ncol}: {prob.params.ntotal} particles, {a.steps} steps, {ms / a.steps:.3f} ms/step, "
      f"{n} launches, {eng.pair_stats()}, tile/list steps {eng.path_counts()}")
if a.profile:
    for k, (t, c) in eng.profile_get().items():
        if c:
            print(f"  {k:20s} {t / a.steps:8.4f} ms/step  {c // a.steps:3d} launches/step  {t / c:8.4f} ms/launch")
