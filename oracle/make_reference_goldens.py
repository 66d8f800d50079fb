#!/usr/bin/env python
"""TEST INFRASTRUCTURE. Runs the UNMODIFIED reference executables that the reference repository ships
(example_problems/<problem>/sph, built by its authors with gfortran 4.8.5 -O3 from the sources beside them) on
the input sets of spsph/decks.py (regenerations of every input set the reference ships, see
tests/test_oracle_cpu.py::test_all_shipped_input_sets_regenerate) and stores what they print as golden vectors:

    tests/golden/ref_<case>.npz   x, vel, stress, strain of every velocity and stress particle at the plotted
                                  steps, all significant digits (oracle/gfortran_shim.c prints %.17g), plus
                                  the positions the binary lists in surface_points.csv

The image has no Fortran run time; oracle/_ref/libgfortran.so.3 (built from oracle/gfortran_shim.c by
`make -C oracle ref`) supplies the 25 libgfortran entry points the binaries import. The binaries are executed
through the dynamic loader where they lie (read-only, no exec bit): nothing of the reference is copied into this
repository; the scratch directories live under oracle/_ref/run (git-ignored).

    python oracle/make_reference_goldens.py            # all cases -> tests/golden/ref_*.npz
    python oracle/make_reference_goldens.py vs bui     # selected cases

Only usable where /root/reference exists (the development container, not the GPU box): the tests read the
committed .npz files."""
import os
import shutil
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stress-particle-sph_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from spsph import decks  # noqa: E402

REF = os.environ.get("SPSPH_REFERENCE", "/root/reference")
EX = os.path.join(REF, "example_problems")
BIN = {"bui": os.path.join(EX, "soil_failure_bui_et_al_2008", "sph"),
       "vs": os.path.join(EX, "vertical_slope", "sph"),
       "sl": os.path.join(EX, "strain_localisation_in_soil_sample", "sph")}
SHIM_DIR = os.path.join(ROOT, "oracle", "_ref")
LOADER = "/lib64/ld-linux-x86-64.so.2"

import reference_runner  # noqa: E402
from ref_cases import CASES  # noqa: E402


DEFAULT_COLS = ["x-coord", "y-coord", "x-vel", "y-vel", "sxx", "syy", "sxy", "szz", "strain"]


def read_csv(path):
    """-> (column names, rows). The Bui copy writes stress_points.csv without its header line (same columns)."""
    rows, hdr = [], None
    with open(path) as f:
        for k, line in enumerate(f):
            v = [t.strip() for t in line.strip().split(",") if t.strip()]
            if not v:
                continue
            try:
                rows.append([float(t) for t in v])
            except ValueError:
                if k == 0:
                    hdr = v
                else:
                    raise
    if hdr is None:
        hdr = DEFAULT_COLS[:len(rows[0])] if rows else DEFAULT_COLS
    return hdr, (np.array(rows, dtype=np.float64) if rows else np.zeros((0, len(hdr))))


def run_case(name):
    which, spec_fn, nsteps, keep = CASES[name]
    spec = spec_fn(nsteps)
    spec["out"] = [1] * 7 + [0, 0, 0]  # plot every stress / velocity component and the plastic strain
    for blk in spec["blocks"]:
        blk["plot_step"] = 1 if isinstance(keep, tuple) else keep
        blk["print_step"] = 1000000
        blk["save_step"] = 1000000
    run = os.path.join(SHIM_DIR, "run", name)
    shutil.rmtree(run, ignore_errors=True)
    os.makedirs(run)
    decks.write_deck(run, spec)
    t0 = time.time()
    rc, _ = reference_runner.run(run, which)
    if rc != 0:
        raise RuntimeError(f"{name}: the reference binary exited with {rc}: " + open(os.path.join(run, "stderr.txt")).read()[-400:])
    out = {"variant": which, "nsteps": nsteps}
    steps = sorted(int(f.split(".")[-1]) for f in os.listdir(run) if f.startswith("nodes.csv."))
    if isinstance(keep, tuple):
        missing = [s for s in keep if s not in steps]
        if missing:
            raise RuntimeError(f"{name}: the binary wrote no frame for steps {missing}")
        steps = list(keep)
    else:
        steps = [s for s in steps if s > 0]
    out["steps"] = np.array(steps, dtype=np.int32)
    for s in steps:
        for kind, tag in (("nodes", "n"), ("stress_points", "s")):
            hdr, a = read_csv(os.path.join(run, f"{kind}.csv.{s:06d}"))
            col = {h: i for i, h in enumerate(hdr)}
            out[f"{tag}{s}_x"] = a[:, [col["x-coord"], col["y-coord"]]]
            out[f"{tag}{s}_vel"] = a[:, [col["x-vel"], col["y-vel"]]]
            out[f"{tag}{s}_stress"] = a[:, [col["sxx"], col["syy"], col["sxy"], col["szz"]]]
            out[f"{tag}{s}_strain"] = a[:, col["strain"]]
        _, sp = read_csv(os.path.join(run, f"surface_points.csv.{s:06d}"))
        out[f"surf{s}"] = sp.reshape(-1, 2) if sp.size else np.zeros((0, 2))
    dst = os.path.join(ROOT, "tests", "golden", f"ref_{name}.npz")
    np.savez_compressed(dst, **out)
    print(f"{name:22s} {time.time() - t0:6.1f} s  frames at steps {steps}  -> {os.path.relpath(dst, ROOT)} "
          f"({os.path.getsize(dst) / 1024:.0f} KiB)")
    if not os.environ.get('KEEP_RUN'):
        shutil.rmtree(run, ignore_errors=True)


if __name__ == "__main__":
    if not os.path.exists(os.path.join(SHIM_DIR, "libgfortran.so.3")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    for b in BIN.values():
        if not os.path.exists(b):
            raise SystemExit(f"{b} not found: the reference is only mounted in the development container")
    for c in (sys.argv[1:] or list(CASES)):
        run_case(c)
