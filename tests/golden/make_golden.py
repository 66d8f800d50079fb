"""Generates tests/golden/*.npz from the CPU oracle (NOT from the reference: it cannot run here, see
oracle/sph_oracle.cpp header). Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "stress-particle-sph_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import spsph  # noqa: E402
from oracle_binding import Oracle  # noqa: E402
from spsph import decks  # noqa: E402

for kind, nsteps in (("bui", 100), ("vs", 100), ("sl", 100)):
    d = tempfile.mkdtemp()
    decks.write_deck(d, decks.SHIPPED[kind]())
    prob = spsph.load(d, kind)
    orc = Oracle(prob)
    orc.run(1, 0.0, prob.blocks[0]["dt"], nsteps)
    a = orc.download()
    nt = prob.params.ntotal
    out = os.path.join(ROOT, "tests", "golden", f"{kind}_{nsteps}steps.npz")
    np.savez_compressed(out, kind=kind, nsteps=nsteps, x=a["x"][:nt], vel=a["vel"][:nt], stress=a["stress"][:nt],
                        epsp=a["internal_vars"][:, 0], npairs=orc.pair_stats()["npairs"])
    print(out, os.path.getsize(out))
