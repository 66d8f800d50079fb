"""CUDA path against the reference's own executables (tests/golden/ref_<case>.npz, see
tests/test_reference_pinned_cpu.py): bit-exact positions, velocities, stresses and plastic strain through the
C-ABI on every input set the reference ships."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ref_cases import CASES, DEVICE_TOLERANCE, DEVICE_UNSUPPORTED, DEVICE_UNVERIFIED, golden_path, spec_of  # noqa: E402
from test_reference_pinned_cpu import compare_with_golden  # noqa: E402

pytestmark = pytest.mark.gpu


def run_case_on_engine(case, tmp_path):
    import spsph
    from spsph import decks
    g = np.load(golden_path(case))
    variant, spec = spec_of(case)
    decks.write_deck(str(tmp_path), spec)
    prob = spsph.load(str(tmp_path), variant)
    dt = prob.blocks[0]["dt"]
    if case in DEVICE_UNSUPPORTED:  # the product path refuses what it does not compute (no silent approximation)
        with pytest.raises(Exception, match="not supported"):
            spsph.Engine(prob)
        return
    eng = spsph.Engine(prob)
    done, t = 0, 0.0
    for step in (int(s) for s in g["steps"]):
        t = eng.run(1 + done, t, dt, step - done)
        done = step
        compare_with_golden(case, g, step, eng.download(), prob.params, "CUDA engine", DEVICE_TOLERANCE.get(case, 0.0))
    eng.close()


@pytest.mark.parametrize("case", [c for c in CASES if c not in DEVICE_UNVERIFIED])
def test_engine_reproduces_reference_binary(case, tmp_path):
    run_case_on_engine(case, tmp_path)


def test_driver_frames_match_reference_cadence(tmp_path):
    """host/sph_driver.cpp (the PROGRAM SPH_2018 stand-in) plots at the steps the reference executable plots (its
    output clock is an fp32 accumulation: frames land at steps 301 and 602 for plot_step = 300), prints the reference's
    node and stress-particle rows and lists the same free-surface nodes in surface_points.csv; its frames come from
    spsph_download_frame (device-side packing)"""
    import subprocess
    from spsph import decks
    g = np.load(golden_path("bui_long"))
    variant, spec = spec_of("bui_long")
    for blk in spec["blocks"]:
        blk["plot_step"] = 300
        blk["print_step"] = blk["save_step"] = 10 ** 6
    deck = tmp_path / "deck"
    deck.mkdir()
    decks.write_deck(str(deck), spec)
    drv = os.path.join(ROOT, "stress-particle-sph_b200", "sph_driver")
    subprocess.check_call([drv, str(deck), variant], stdout=subprocess.DEVNULL)
    frames = sorted(int(f.split(".")[-1]) for f in os.listdir(deck) if f.startswith("nodes.csv."))
    assert frames == [0] + [int(s) for s in g["steps"]], frames
    for step in (int(s) for s in g["steps"]):
        rows = [ln.split(",") for ln in open(deck / f"surface_points.csv.{step:06d}").read().splitlines()[1:] if ln.strip()]
        surf = np.array([[float(r[0]), float(r[1])] for r in rows]).reshape(-1, 2)
        assert np.array_equal(surf, g[f"surf{step}"])
        nodes = np.loadtxt(deck / f"nodes.csv.{step:06d}", delimiter=",", skiprows=1, usecols=range(9))
        ref = np.column_stack([g[f"n{step}_x"], g[f"n{step}_vel"], g[f"n{step}_stress"], g[f"n{step}_strain"]])
        assert np.abs(nodes - ref).max() <= 0.5e-8 * 1.0001 + 1e-8 * np.abs(ref).max() * 0, "F16.8 frames differ from the reference"
        sps = np.loadtxt(deck / f"stress_points.csv.{step:06d}", delimiter=",", usecols=range(9))
        ref = np.column_stack([g[f"s{step}_x"], g[f"s{step}_vel"], g[f"s{step}_stress"], g[f"s{step}_strain"]])
        assert np.abs(sps - ref).max() <= 0.5e-8 * 1.0001, "F16.8 stress-particle frames differ from the reference"
