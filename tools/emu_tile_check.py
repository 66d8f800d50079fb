#!/usr/bin/env python
"""TEST TOOL (CPU, no GPU): builds the SIMT host emulation of csrc/spsph_engine.cu (tests/native/) and runs decks on it
against the oracle, bit for bit, printing which path (cell-tile / id-list) every step took.
usage: python tools/emu_tile_check.py [case ...]   cases: bui bui_dx02 vs sl bui_inside bui_outside refined102 ..."""
import ctypes as C
import os
import sys
import tempfile
import pathlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stress-particle-sph_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from test_step_emulation_cpu import _build_emulated, STATE_KEYS  # noqa: E402
import spsph  # noqa: E402
import spsph.engine as E  # noqa: E402
from spsph import decks  # noqa: E402
from oracle_binding import Oracle  # noqa: E402

CASES = {
    "bui": ("bui", lambda: decks.bui_spec(), 6),
    "bui_dx02": ("bui", lambda: decks.bui_spec(dx=0.2, maxtimestep=1000), 8),
    "vs": ("vs", lambda: decks.vertical_slope_spec(), 6),
    "sl": ("sl", lambda: decks.strain_localisation_spec(), 4),
    "bui_inside": ("bui", lambda: decks.bui_spec(mode="inside", npoints=2), 5),
    "bui_outside": ("bui", lambda: decks.bui_spec(mode="outside"), 5),
    "refined102": ("bui", lambda: decks.refined_bui_spec(ncol=102), 4),
    "wide": ("vs", lambda: decks.wide_slope_spec(ncol=60), 4),
}


def main():
    names = sys.argv[1:] or ["bui_dx02", "vs"]
    d = pathlib.Path(tempfile.mkdtemp(prefix="emu_tile_"))
    so = _build_emulated(d, ("-DSPSPH_EMU_SIMT",))
    E._lib, E._CUDA_SO = None, so
    lib = E.cuda_lib()
    lib.spsph_path_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    bad = 0
    for name in names:
        variant, spec_fn, nsteps = CASES[name]
        dd = tempfile.mkdtemp(prefix="deck_")
        decks.write_deck(dd, spec_fn())
        prob = spsph.load(dd, variant)
        p = prob.params
        dt = prob.blocks[0]["dt"]
        orc, eng = Oracle(prob), E.Engine(prob)
        t = 0.0
        for step in range(1, nsteps + 1):
            orc.step(step, t, dt)
            eng.step(step, t, dt)
            t += dt
            a, b = eng.download(), orc.download()
            tl, ll = C.c_int64(), C.c_int64()
            lib.spsph_path_counts(eng.h, C.byref(tl), C.byref(ll))
            msg = []
            for k in STATE_KEYS:
                x, y = a[k], b[k]
                if k in ("x", "vel", "stress"):
                    x, y = x[:p.ntotal], y[:p.ntotal]
                if not np.array_equal(x, y):
                    msg.append(f"{k}:{int((x != y).sum())}")
            ok = eng.pair_stats() == orc.pair_stats()
            print(f"{name} step {step}: tile/list steps {tl.value}/{ll.value}  stats {'ok' if ok else 'DIFF'} "
                  f"{eng.pair_stats()}  {'BITWISE OK' if not msg else 'DIFF ' + ' '.join(msg)}", flush=True)
            if msg or not ok:
                bad += 1
                break
        eng.close()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
