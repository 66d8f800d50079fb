"""GiD mesh / result writers of the host driver (host/sph_gid.cpp: OutputMesh mat:2707-2744, OutputRes mat:2930-3008)
against the files the reference's own executable writes (tests/golden/gid_bui_dx02.post.{msh,res}.gz, produced by
oracle/make_gid_golden.py): same records in the same order, integers and keywords identical, every REAL equal as a
double (list-directed formatting -- column widths, 0 vs 0.0 -- is the run-time library's business, not compared)."""
import gzip
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _tokens(text):
    out = []
    for line in text.splitlines():
        if line.strip():
            out.append(line.replace('"', ' " ').split())
    return out


def _same(a, b):
    if a == b:
        return True
    try:
        return float(a) == float(b)
    except ValueError:
        return False


def _compare(mine, ref, what):
    A, B = _tokens(mine), _tokens(ref)
    assert len(A) == len(B), f"{what}: {len(A)} records vs {len(B)} in the reference file"
    for k, (ra, rb) in enumerate(zip(A, B)):
        assert len(ra) == len(rb) and all(_same(x, y) for x, y in zip(ra, rb)), f"{what}, record {k + 1}: {ra} vs {rb}"


def test_gid_files_equal_the_reference_executables(tmp_path):
    import spsph
    from spsph import decks
    from spsph.problem import GidWriter
    from oracle_binding import Oracle
    import make_gid_golden
    d = str(tmp_path / "deck")
    os.makedirs(d)
    decks.write_deck(d, make_gid_golden.spec())
    prob = spsph.load(d, "bui")
    dt = prob.blocks[0]["dt"]
    prefix = str(tmp_path / "co_soil")
    w = GidWriter(d, "bui", prefix)
    w.mesh(prob.arrays["x"])
    orc = Oracle(prob)
    w.frame(orc.download(), 0.0)
    t = 0.0
    for step in range(1, 7):
        orc.step(step, t, dt)
        t = t + dt
        if step % 3 == 0:  # plot_step = 3 (the fp32 cadence counters of the driver agree for this few steps)
            w.frame(orc.download(), t)
    w.close()
    for ext in ("post.msh", "post.res"):
        ref = gzip.open(os.path.join(ROOT, "tests", "golden", f"gid_bui_dx02.{ext}.gz"), "rt").read()
        _compare(open(f"{prefix}.{ext}").read(), ref, ext)
