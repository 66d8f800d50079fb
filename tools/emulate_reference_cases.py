#!/usr/bin/env python
"""Runs reference-executable cases (oracle/ref_cases.py) on the HOST-EMULATED CUDA engine (tests/test_step_emulation_cpu.py)
at their full resolution and step count, in lockstep with the oracle (which supplies each step's pair list), and compares
the emulated engine's frames with the goldens of the reference's own executables -- the GPU parity test of
tests/test_gpu_reference.py without a GPU (everything but the neighbour build and CUDA's libm).

    python tools/emulate_reference_cases.py [--standalone] [case ...]     default: every case the emulation supports

--standalone: the emulated engine builds its own neighbour lists (the complete device path; no list is imported from
the oracle, which then only serves as a progress check); otherwise the oracle supplies each step's pair list.
"""
import ctypes as C
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("oracle", "tests", "stress-particle-sph_b200", ""):
    sys.path.insert(0, os.path.join(ROOT, p))
import spsph  # noqa: E402
from oracle_binding import Oracle, lib  # noqa: E402
from ref_cases import CASES, golden_path, spec_of  # noqa: E402
from spsph import decks  # noqa: E402
from test_list_kernels_cpu import growth_rule  # noqa: E402
from test_reference_pinned_cpu import compare_with_golden  # noqa: E402
import test_step_emulation_cpu as T  # noqa: E402


class _Factory:
    def mktemp(self, name):
        import pathlib
        return pathlib.Path(tempfile.mkdtemp(prefix=name))


def main():
    gen = T.emu_engine.__wrapped__(_Factory())  # the fixture's generator: builds the emulated engine library
    E = next(gen)
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    standalone = "--standalone" in sys.argv
    skip = (lambda c: False) if standalone else (
        lambda c: c.endswith("_long") or c.startswith("bui_inside") or c == "bui_full")
    cases = args or [c for c in CASES if not skip(c)]
    L = lib()
    L.oracle_debug_grid.restype = None
    bad = 0
    for case in cases:
        g = np.load(golden_path(case))
        variant, spec = spec_of(case)
        d = tempfile.mkdtemp(prefix="emu_" + case)
        decks.write_deck(d, spec)
        prob = spsph.load(d, variant)
        p = prob.params
        dt = prob.blocks[0]["dt"]
        orc, eng = Oracle(prob), E.Engine(prob)
        cells = np.zeros(p.ntotal2, np.int32)
        mb, npairs = C.c_int64(), C.c_int64()
        t, t0 = 0.0, time.time()
        frames = [int(s) for s in g["steps"]]
        try:
            for step in range(1, frames[-1] + 1):
                if standalone:
                    eng.step(step, t, dt)
                    t = t + dt
                    if step in frames:
                        compare_with_golden(case, g, step, eng.download(), p, "emulated CUDA engine (stand-alone)")
                    continue
                before = orc.download()
                orc.step(step, t, dt)
                L.oracle_debug_grid(C.c_void_p(orc.h), cells.ctypes.data_as(C.c_void_p), C.byref(mb), C.byref(npairs))
                pairs = orc.pairs()
                rule = growth_rule(prob, cells, pairs, mb.value, npairs.value)
                lists, keep = T.build_step(prob, cells, pairs, before["x"], rule, npairs.value, before["hsml"])
                assert E.cuda_lib().spsph_emu_set_lists(eng.h, C.byref(lists)) == 0
                eng.step(step, t, dt)
                t = t + dt
                if step in frames:
                    compare_with_golden(case, g, step, eng.download(), p, "emulated CUDA engine")
            print(f"{case:24s} OK   {frames[-1]} steps, frames {frames}, {time.time() - t0:.0f} s", flush=True)
        except AssertionError as e:
            bad += 1
            print(f"{case:24s} FAIL {str(e)[:300]}", flush=True)
        eng.close()
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
