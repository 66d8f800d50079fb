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
