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
from ref_cases import CASES, DEVICE_TOLERANCE, DEVICE_UNSUPPORTED, golden_path, spec_of  # noqa: E402
from test_reference_pinned_cpu import compare_with_golden  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", list(CASES))
def test_engine_reproduces_reference_binary(case, tmp_path):
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
