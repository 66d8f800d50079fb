"""TEST INFRASTRUCTURE. Runs one of the reference's own prebuilt executables (example_problems/*/sph of the
reference repository, gfortran 4.8.5 -O3) on a deck directory, with oracle/_ref/libgfortran.so.3
(oracle/gfortran_shim.c) as its Fortran run-time library.

`make -C oracle ref` places copies of the three executables in oracle/_ref/ (git-ignored build output that travels
to the GPU box like the other built artefacts, so that bench.py can time the real reference there); in the
development container the originals under /root/reference are used where they lie."""
import os
import resource
import subprocess
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
LOADER = "/lib64/ld-linux-x86-64.so.2"
_EX = {"bui": "soil_failure_bui_et_al_2008", "vs": "vertical_slope", "sl": "strain_localisation_in_soil_sample"}


def binary(which):
    """path of the reference executable for a problem family, or None"""
    for p in (os.path.join(REF_DIR, f"sph_{which}"),
              os.path.join(os.environ.get("SPSPH_REFERENCE", "/root/reference"), "example_problems", _EX[which], "sph")):
        if os.path.exists(p):
            return p
    return None


def available(which="bui"):
    return binary(which) is not None and os.path.exists(os.path.join(REF_DIR, "libgfortran.so.3"))


def _unlimit_stack():  # RK4 keeps ~43 doubles per particle in automatic arrays (main:653-680)
    try:
        _, hard = resource.getrlimit(resource.RLIMIT_STACK)
        resource.setrlimit(resource.RLIMIT_STACK, (hard, hard))
    except (ValueError, OSError):
        pass


def run(deck_dir, which, timeout=None):
    """runs the executable in deck_dir (it reads input.txt there and writes its frames there);
    returns (exit status, wall seconds)"""
    t0 = time.perf_counter()
    with open(os.path.join(deck_dir, "stdout.txt"), "w") as so, open(os.path.join(deck_dir, "stderr.txt"), "w") as se:
        rc = subprocess.run([LOADER, "--library-path", REF_DIR, binary(which)], cwd=deck_dir, stdout=so, stderr=se,
                            preexec_fn=_unlimit_stack, timeout=timeout).returncode
    return rc, time.perf_counter() - t0
