"""The oracle is pinned to the reference itself: tests/golden/ref_<case>.npz hold what the reference's OWN
executables (example_problems/*/sph, built by its authors with gfortran 4.8.5 -O3) computed on every input set the
reference ships (oracle/make_reference_goldens.py ran them here through oracle/gfortran_shim.c). The oracle must
reproduce positions, velocities, stresses and plastic strain of every velocity and stress particle BIT FOR BIT
(north_star asks for 1e-9 relative after 100 steps; the restatement is exact, so the test asks for equality)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from ref_cases import CASES, golden_path, spec_of  # noqa: E402

FIELDS = (("x", "x", None), ("vel", "vel", None), ("stress", "stress", None), ("strain", "internal_vars", 0))


def compare_with_golden(case, g, step, arrays, p, label, rel_tol=0.0):
    """arrays: a download (oracle or engine) after `step` steps; asserts equality with the reference frame
    (rel_tol > 0: relative L-inf per field instead, for paths that call a different libm)"""
    nn = p.nnode
    for tag, sl in (("n", slice(0, nn)), ("s", slice(nn, p.ntotal))):
        for key, mine_key, col in FIELDS:
            ref = g[f"{tag}{step}_{key}"]
            mine = arrays[mine_key][sl] if col is None else arrays[mine_key][sl, col]
            if rel_tol > 0 and np.abs(ref - mine).max() <= rel_tol * max(np.abs(ref).max(), 1e-300):
                continue
            if not np.array_equal(ref, mine):
                d = np.abs(ref - mine)
                k = int(np.argmax(d.reshape(len(ref), -1).max(axis=1)))
                raise AssertionError(
                    f"{label} differs from the reference binary: case {case}, step {step}, "
                    f"{'velocity' if tag == 'n' else 'stress'} particles, field {key}: rel Linf "
                    f"{d.max() / max(np.abs(ref).max(), 1e-300):.3e} (worst particle {k}: ref {ref[k]}, got {mine[k]})")
    # surface_points.csv: positions of the velocity particles that get_nodes_on_free_surface marked (bc_or_not = 2)
    surf = arrays["x"][:nn][arrays["bc_or_not"][:nn] == 2]
    ref = g[f"surf{step}"]
    assert surf.shape == ref.shape and np.array_equal(surf, ref), (
        f"{label}: free-surface nodes differ from the reference binary's surface_points.csv: case {case}, step {step}, "
        f"{len(surf)} vs {len(ref)} nodes")


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_reproduces_reference_binary(case, tmp_path):
    import spsph
    from spsph import decks
    from oracle_binding import Oracle
    if case in ("bui_full", "sl_full") and not os.environ.get("SPSPH_FULL_RUNS"):
        # 16 650 / 2 092 steps of the serial oracle take 5 / 1 minutes: outside the default CPU suite (which runs the
        # complete vertical-slope problem). SPSPH_FULL_RUNS=1 runs them (done for every change of the oracle, result
        # recorded in DESIGN.md); the GPU suite always does (tests/test_zz_gpu_new_paths.py).
        pytest.skip("set SPSPH_FULL_RUNS=1 for the complete Bui and strain-localisation runs on the CPU oracle")
    g = np.load(golden_path(case))
    variant, spec = spec_of(case)
    assert str(g["variant"]) == variant
    decks.write_deck(str(tmp_path), spec)
    prob = spsph.load(str(tmp_path), variant)
    dt = prob.blocks[0]["dt"]
    orc = Oracle(prob)
    done, t = 0, 0.0
    steps = [int(s) for s in g["steps"]]
    for step in steps:
        t = orc.run(1 + done, t, dt, step - done)
        done = step
        compare_with_golden(case, g, step, orc.download(), prob.params, "oracle")


def test_goldens_cover_every_shipped_input_set():
    # the 16 input directories of the reference hold 14 distinct input sets (each top-level input.txt repeats one
    # of its variants: tests/test_oracle_cpu.py::VARIANT_DIRS), plus the two long runs
    shipped = [c for c in CASES if not c.endswith("_long") and c.split("_", 1)[-1] not in (
        "refined", "wide", "gauss", "quintic", "art_stress", "cont_density", "cont_density_sle2", "tresca",
        "mohr_coulomb", "dp_perzyna", "vm_expflow", "vm_powflow", "sigman", "xsph", "sigman_xsph", "full", "out_domain", "sine_bc", "plane_stress", "outside_sp1",
        "outside_sp3", "shift5", "sml15")]
    assert len(shipped) == 14, shipped
    for c in CASES:
        assert os.path.exists(golden_path(c)), c


def test_gfortran_shim_builds_and_regenerates_a_golden(tmp_path):
    """the run-time stand-in builds; where the reference checkout is mounted, re-running its binary reproduces the
    committed vertical-slope golden exactly"""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    assert os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libgfortran.so.3"))
    ref_bin = "/root/reference/example_problems/vertical_slope/sph"
    if not os.path.exists(ref_bin):
        pytest.skip("reference checkout not mounted (GPU box)")
    import make_reference_goldens as m
    before = dict(np.load(golden_path("vs")))
    keep = golden_path("vs") + ".keep"
    os.replace(golden_path("vs"), keep)
    try:
        m.run_case("vs")
        after = dict(np.load(golden_path("vs")))
    finally:
        os.replace(keep, golden_path("vs"))
    assert set(before) == set(after)
    for k in before:
        assert np.array_equal(before[k], after[k]), k
