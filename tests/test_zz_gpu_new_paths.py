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
