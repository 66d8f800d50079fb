#!/usr/bin/env python
"""Runs reference-executable golden cases on cuda:0 with tolerance 0 and reports which are bit-exact.
usage: python tools/strict_cases.py case [case ...]"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ("stress-particle-sph_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, d))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402

__graft_entry__.build()
import spsph  # noqa: E402
from spsph import decks  # noqa: E402
from ref_cases import golden_path, spec_of  # noqa: E402
from test_reference_pinned_cpu import compare_with_golden  # noqa: E402

for case in sys.argv[1:]:
    g = np.load(golden_path(case))
    variant, spec = spec_of(case)
    d = tempfile.mkdtemp()
    decks.write_deck(d, spec)
    prob = spsph.load(d, variant)
    dt = prob.blocks[0]["dt"]
    eng = spsph.Engine(prob)
    done, t, ok = 0, 0.0, True
    for step in (int(s) for s in g["steps"]):
        t = eng.run(1 + done, t, dt, step - done)
        done = step
        try:
            compare_with_golden(case, g, step, eng.download(), prob.params, "CUDA engine", 0.0)
        except AssertionError as e:
            ok = False
            print(f"{case}: NOT bit-exact at step {step}: {str(e)[:300]}")
            break
    if ok:
        print(f"{case}: BIT-EXACT against the reference executable ({done} steps)")
    eng.close()
