#!/usr/bin/env python
"""TEST INFRASTRUCTURE. Runs the reference's own Bui executable (oracle/reference_runner.py) on a coarse column with a
plot cadence of 3 steps and stores the GiD files it writes (<name>.post.msh, <name>.post.res: OutputMesh / OutputRes,
3_SPH_material_2018.f90:2707-2744, 2930-3008) gzip-compressed under tests/golden/, for tests/test_gid_writer_cpu.py.
The run-time stand-in (oracle/gfortran_shim.c) prints every REAL of a list-directed record with 17 significant digits."""
import gzip
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stress-particle-sph_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import reference_runner as rr  # noqa: E402
from spsph import decks  # noqa: E402


def spec():
    s = decks.bui_spec(dx=0.2, maxtimestep=7)
    for b in s["blocks"]:
        b["plot_step"] = 3
    return s


if __name__ == "__main__":
    d = tempfile.mkdtemp(prefix="gid_golden_")
    decks.write_deck(d, spec())
    rc, _ = rr.run(d, "bui")
    assert rc == 0, open(os.path.join(d, "stderr.txt")).read()
    for ext in ("post.msh", "post.res"):
        src = os.path.join(d, f"co_soil.{ext}")
        with open(src, "rb") as f, gzip.open(os.path.join(ROOT, "tests", "golden", f"gid_bui_dx02.{ext}.gz"), "wb", 9) as g:
            shutil.copyfileobj(f, g)
        print(ext, os.path.getsize(src), "bytes")
