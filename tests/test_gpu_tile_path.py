"""The cell-tile path (csrc/tile_kernels.cuh, selected with SPSPH_TILE=1) against the oracle and against the reference's
own executables: bit for bit, like the id-list path. Every test also asserts that the steps really ran on the tile
kernels (spsph_path_counts), so a silent fall-back to the list path cannot pass for the tile path.

Covered: the three shipped problems (walls, XSPH, artificial viscosity, velocity-vector re-seating; CSPM, boundary
conditions, damping; von-Mises Perzyna), reversed first step and forward steps, the ordered pair list re-created on
demand, free-surface marks at download (the id lists are materialised for it), inside approach with wall forces,
fixed re-seating, a 600-step run through list growth (growth steps fall back to the list path, the others stay on the
tile path), the refined column at 250 k particles, restart from a download, and reference-executable goldens."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_gpu_parity import _compare, _load, _pairs_equal  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def tile_path(monkeypatch):
    monkeypatch.setenv("SPSPH_TILE", "1")  # read by spsph_create


SETUP_KEYS = ("rho", "mass", "hsml", "itype", "wall_position", "horizontal_or_not", "bc_info")


def _full_state(prob, snap):
    """a download completed with the set-up arrays spsph_download does not return (wall geometry, BC tables)"""
    full = dict(snap)
    for k in SETUP_KEYS:
        full[k] = prob.arrays[k]
    return full


def _run(prob, nsteps, label, check_at=(), pairs_at=(), min_tile_share=0.9):
    import spsph
    from oracle_binding import Oracle
    dt = prob.blocks[0]["dt"]
    eng, orc = spsph.Engine(prob), Oracle(prob)
    t = 0.0
    for it in range(1, nsteps + 1):
        eng.step(it, t, dt)
        orc.step(it, t, dt)
        t = t + dt
        if it in pairs_at:
            _pairs_equal(eng.pairs(), orc.pairs(), f"{label} step {it}")
            assert eng.pair_stats() == orc.pair_stats()
        if it in check_at or it == nsteps:
            _compare(eng.download(), orc.download(), prob.params.ntotal, f"{label} after {it} steps (tile path)")
    tile, lst = eng.path_counts()
    assert tile + lst == nsteps and tile >= min_tile_share * nsteps, f"{label}: only {tile} of {nsteps} steps on the tile path"
    eng.close()
    orc.close()


@pytest.mark.parametrize("kind,nsteps", [("bui", 100), ("vs", 100), ("sl", 60)])
def test_tile_path_shipped_problems(deck_dir, kind, nsteps):
    _run(_load(deck_dir, kind), nsteps, kind, check_at=(1, 2, 3), pairs_at=(1, 2, 3))


@pytest.mark.parametrize("kw", [dict(mode="inside", npoints=2), dict(mode="inside", npoints=1), dict(mode="outside"),
                                dict(mode="inside", npoints=3)], ids=["inside_sp2", "inside_sp1", "outside_fixed", "inside_sp3"])
def test_tile_path_bui_variants(tmp_path, kw):
    import spsph
    from spsph import decks
    decks.write_deck(str(tmp_path), decks.bui_spec(maxtimestep=100, **kw))
    _run(spsph.load(str(tmp_path), "bui"), 40, f"bui {kw}", check_at=(1, 2))


def test_tile_path_long_run_with_list_growth(deck_dir):
    """600 steps of the Bui column: in the steps where the pair list grows the engine takes the id-list path (split
    traversal order), everywhere else the tile path; the two interleave on one state"""
    import spsph
    from oracle_binding import Oracle
    prob = _load(deck_dir, "bui")
    dt = prob.blocks[0]["dt"]
    eng, orc = spsph.Engine(prob), Oracle(prob)
    eng.run(1, 0.0, dt, 600)
    orc.run(1, 0.0, dt, 600)
    _compare(eng.download(), orc.download(), prob.params.ntotal, "bui after 600 steps (tile path)")
    tile, lst = eng.path_counts()
    # (from step ~200 on the column spreads and the pair count passes its previous maximum in every other step)
    assert lst >= 1, "the 600-step run is expected to contain list-growth steps"
    assert tile >= 200 and tile + lst == 600, f"only {tile} of 600 steps on the tile path"


def test_tile_path_refined_column_250k(deck_dir):
    prob = _load(deck_dir, "refined_bui_spec", ncol=408)
    _run(prob, 6, "refined Bui column, 251 k particles", check_at=(1, 2))


def test_tile_path_restart_from_download(deck_dir):
    """download after 5 steps, upload into a fresh engine (first step after an upload walks the pairs in reversed order
    unless the list length is restored), continue: equals the uninterrupted run"""
    import spsph
    prob = _load(deck_dir, "bui")
    dt = prob.blocks[0]["dt"]
    a = spsph.Engine(prob)
    t = a.run(1, 0.0, dt, 5)
    snap, cap = a.download(), a.list_capacity()
    a.run(6, t, dt, 5)
    b = spsph.Engine(prob)
    b.upload(_full_state(prob, snap))
    b.set_list_capacity(cap)
    b.run(6, t, dt, 5)
    _compare(b.download(), a.download(), prob.params.ntotal, "restart on the tile path")
    assert b.path_counts()[0] == 5


@pytest.mark.parametrize("case", ["bui", "vs", "sl", "bui_outside", "bui_inside_sp1", "bui_refined", "vs_wide", "sl_sp2",
                                  "bui_out_domain", "bui_shift5", "bui_outside_sp1", "bui_outside_sp3", "sl_sine_bc"])
def test_tile_path_reproduces_reference_binary(case, tmp_path):
    """the reference's own executables' output (tests/golden/ref_<case>.npz), bit for bit, on the tile path"""
    from test_gpu_reference import run_case_on_engine
    run_case_on_engine(case, tmp_path)


def test_tile_path_falls_back_when_masks_overflow(tmp_path):
    """sml = 1.5: a stencil row holds more candidates than an entry code addresses -> the engine stays on the list path
    (same results), it does not compute something else"""
    import spsph
    from spsph import decks
    spec = decks.bui_spec(maxtimestep=100)
    spec["sml"] = 1.5
    decks.write_deck(str(tmp_path), spec)
    prob = spsph.load(str(tmp_path), "bui")
    _run(prob, 5, "bui sml = 1.5", min_tile_share=0.0)


def test_row_wise_transfers_single_gpu(deck_dir, monkeypatch):
    """spsph_upload_rows / spsph_download_rows: (1) the rows of all particles == the full transfer; (2) a subset of rows
    replaces exactly those particles' time-varying state"""
    monkeypatch.setenv("SPSPH_TILE", "0")
    import spsph
    from spsph.engine import row_arrays
    prob = _load(deck_dir, "bui")
    p = prob.params
    dt = prob.blocks[0]["dt"]
    a = spsph.Engine(prob)
    t = a.run(1, 0.0, dt, 7)
    snap = a.download()
    ids_all = np.arange(p.ntotal2, dtype=np.int32)
    keys = ("x", "vel", "stress", "internal_vars", "f_drucker", "displ", "x_10", "disp_10", "n_int", "bc_int",
            "if_out_domain")
    rows = a.download_rows(ids_all, keys=keys)
    for k in keys:
        x, y = rows[k], snap[k]
        if k in ("vel", "stress"):
            x, y = x[:p.ntotal], y[:p.ntotal]
        assert np.array_equal(x, y), k
    # a fresh engine fed row-wise continues like one fed by spsph_upload
    b, c = spsph.Engine(prob), spsph.Engine(prob)
    b.upload(_full_state(prob, snap))
    c.upload_rows(row_arrays(p, snap, ids_all), ids_all)
    b.run(8, t, dt, 5)
    c.run(8, t, dt, 5)
    _compare(c.download(), b.download(), p.ntotal, "row-wise upload")
    # subset: every third particle takes the snapshot, the others keep the initial state
    sub = ids_all[::3].copy()
    d = spsph.Engine(prob)
    d.upload_rows(row_arrays(p, snap, sub), sub)
    got, init = d.download(), prob.arrays
    mask = np.zeros(p.ntotal2, bool)
    mask[sub] = True
    for k in ("x", "vel", "stress"):
        n = p.ntotal if k != "x" else p.ntotal2
        assert np.array_equal(got[k][:n][mask[:n]], snap[k][:n][mask[:n]]), k
        assert np.array_equal(got[k][:n][~mask[:n]], init[k][:n][~mask[:n]]), k
