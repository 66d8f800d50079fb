"""CUDA paths added after round 1's GPU budget was spent, against the reference's own executables
(tests/golden/ref_<case>.npz): Perzyna viscoplasticity with the Tresca, Mohr-Coulomb and Drucker-Prager criteria,
apply_stress_free (ifsigman = 1) with get_nodes_on_free_surface evaluated every step, XSPH together with boundary
conditions. Their per-particle arithmetic is checked on the CPU (tests/test_oracle_cpu.py::
test_device_math_transcription_*); the list traversal of k_fs_normals / k_xsph_marks has its first hardware run here.
Sorted last in the suite on purpose (oracle/ref_cases.py::DEVICE_UNVERIFIED)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ref_cases import DEVICE_UNVERIFIED  # noqa: E402
from test_gpu_reference import run_case_on_engine  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", sorted(DEVICE_UNVERIFIED))
def test_new_path_reproduces_reference_binary(case, tmp_path):
    run_case_on_engine(case, tmp_path)


def test_checkpoint_resume_is_bit_exact(deck_dir, tmp_path):
    """spsph.checkpoint: a Bui run saved after step 330 (the pair list grew at step 301) and resumed on a fresh engine
    continues bit for bit (the oracle-side proof of the concept is tests/test_oracle_cpu.py::
    test_restart_needs_the_list_capacity)"""
    import numpy as np
    import spsph
    from spsph import checkpoint
    prob = spsph.load(deck_dir("bui"), "bui")
    dt = prob.blocks[0]["dt"]
    e1 = spsph.Engine(prob)
    t_mid = e1.run(1, 0.0, dt, 330)
    path = str(tmp_path / "step330.npz")
    checkpoint.save(path, e1, 330, t_mid)
    e1.run(331, t_mid, dt, 60)
    ref = e1.download()
    e2 = spsph.Engine(prob, upload=False)
    it, t = checkpoint.resume(path, e2, prob)
    assert (it, t) == (330, t_mid)
    e2.run(it + 1, t, dt, 60)
    got = e2.download()
    for k in checkpoint.DYNAMIC:
        assert np.array_equal(ref[k], got[k]), k
    e1.close()
    e2.close()


@pytest.mark.parametrize("kind", ["bui", "sl"])
def test_download_frame_equals_download(kind, deck_dir):
    """spsph_download_frame (device-side packing of what OutputRes prints, mat:2919-3060): every column of the packed
    table equals the corresponding array of a full spsph_download bit for bit, for velocity particles, stress
    particles and wall particles, any column order, any row range; the free-surface marks come with it"""
    import numpy as np
    import spsph
    from spsph.engine import FRAME_COLS
    prob = spsph.load(deck_dir(kind), kind)
    p, dt = prob.params, prob.blocks[0]["dt"]
    eng = spsph.Engine(prob)
    eng.run(1, 0.0, dt, 25)
    cols = list(FRAME_COLS)  # all 16 columns
    tab = eng.download_frame(cols)  # the marks are evaluated by the frame call itself (before any spsph_download)
    full = eng.download()
    nn, nt, n2 = p.nnode, p.ntotal, p.ntotal2
    assert tab.shape == (n2, 16)
    want = {
        "x": full["x"][:, 0], "y": full["x"][:, 1], "vx": full["vel"][:, 0], "vy": full["vel"][:, 1],
        "sxx": full["stress"][:, 0], "syy": full["stress"][:, 1], "sxy": full["stress"][:, 2], "szz": full["stress"][:, 3],
        "rho": full["rho"], "hsml": full["hsml"],
    }
    part = {"epsp": full["internal_vars"][:, 0], "f_drucker": full["f_drucker"], "bc_or_not": full["bc_or_not"].astype(np.float64)}
    node = {"disp_10": full["disp_10"], "displ_x": full["displ"][:, 0], "displ_y": full["displ"][:, 1]}
    for k, c in enumerate(cols):
        col = tab[:, k]
        if c in want:
            assert np.array_equal(col.view(np.uint64), np.ascontiguousarray(want[c]).view(np.uint64)), c
        elif c in part:
            assert np.array_equal(col[:nt], part[c]) and not col[nt:].any(), c
        else:
            assert np.array_equal(col[:nn], node[c]) and not col[nn:].any(), c
    # a writer's selection: the ParaView row of a stress particle, a sub-range, reordered columns
    sel = ["y", "x", "szz", "epsp", "vx"]
    first, count = nn + 5, min(300, nt - nn - 5)
    sub = eng.download_frame(sel, first, count)
    assert np.array_equal(sub, tab[first:first + count][:, [FRAME_COLS[c] for c in sel]])
    assert eng.download_frame(["x"], 0, 0).shape == (0, 1)
    for bad in (dict(cols=[99]), dict(cols=["x"], first=n2 - 1, count=2), dict(cols=["x"] * 17)):
        with pytest.raises(RuntimeError):
            eng.download_frame(**bad)
    eng.close()
