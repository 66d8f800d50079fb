"""Full-size checks (BASELINE.json configs[3], 4.0 M particles): bitwise parity with the oracle over the first
two steps (step 1 walks the pair list reversed, step 2 forward), size-independent properties after more steps,
and a mid-size (250 k particles) bitwise run."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bitwise(a, b, nt, label):
    for k in ("x", "vel", "stress", "internal_vars", "f_drucker", "displ", "x_10", "disp_10", "n_int", "bc_int"):
        x, y = a[k], b[k]
        if k in ("x", "vel", "stress"):
            x, y = x[:nt], y[:nt]
        assert np.array_equal(x, y), f"{label}: {k} differs in {int((x != y).sum())} entries"


def test_refined_bui_250k_bitwise(deck_dir):
    import spsph
    from oracle_binding import Oracle
    prob = spsph.load(deck_dir("refined_bui_spec", ncol=408), "bui")
    assert prob.params.ntotal == 409 * 205 * 3
    dt = prob.blocks[0]["dt"]
    eng, orc = spsph.Engine(prob), Oracle(prob)
    eng.run(1, 0.0, dt, 6)
    orc.run(1, 0.0, dt, 6)
    assert eng.pair_stats() == orc.pair_stats()
    _bitwise(eng.download(), orc.download(), prob.params.ntotal, "refined Bui 250k, 6 steps")


def test_refined_bui_4m_bitwise_and_properties(deck_dir):
    import spsph
    from oracle_binding import Oracle
    prob = spsph.load(deck_dir("refined_bui_spec", ncol=1632), "bui")
    p = prob.params
    assert p.ntotal == 4002483
    dt = prob.blocks[0]["dt"]
    eng, orc = spsph.Engine(prob), Oracle(prob)
    t = eng.run(1, 0.0, dt, 2)
    orc.run(1, 0.0, dt, 2)
    st = eng.pair_stats()
    assert st == orc.pair_stats()
    _bitwise(eng.download(), orc.download(), p.ntotal, "refined Bui 4M, 2 steps")
    orc.close()
    # ---- size-independent properties after 10 more steps
    eng.run(3, t, dt, 10)
    a = eng.download()
    nt = p.ntotal
    assert all(np.isfinite(a[k]).all() for k in ("x", "vel", "stress", "internal_vars"))
    assert np.array_equal(a["mass"], prob.arrays["mass"]) and np.array_equal(a["rho"], prob.arrays["rho"])
    # Drucker-Prager admissibility after adapt_stress2 (mat:2096-2134): sqrt(J2) <= -3*alpha*sigma_m + k_c >= 0
    tanfi, coh = p.props[12], p.props[13]
    alpha2 = tanfi / np.sqrt(9 + 12 * tanfi ** 2)
    kc = 3 * coh / np.sqrt(9 + 12 * tanfi ** 2)
    s = a["stress"][:nt]
    sm = (s[:, 0] + s[:, 1] + s[:, 3]) / 3
    j2 = s[:, 2] ** 2 + 0.5 * ((s[:, 0] - sm) ** 2 + (s[:, 1] - sm) ** 2 + (s[:, 3] - sm) ** 2)
    yld = -3 * alpha2 * sm + kc
    assert (yld >= -1e-6).all() and (np.sqrt(j2) <= yld * (1 + 1e-9) + 1e-6).all()
    # every pair is seen from both sides: interaction counts sum to twice the pair count
    st = eng.pair_stats()
    assert st["noiac"] == 0 and st["miniac"] >= 1
    # the column has started to settle under gravity: all velocity particles move down, none up
    vy = a["vel"][:p.nnode, 1]
    assert vy.min() < 0 and vy.max() <= 1e-12
    eng.close()


def test_refined_bui_1m_20_steps_bitwise(deck_dir):
    """1.0 M particles x 20 steps against the oracle, bit for bit (about 90 s of CPU time for the serial oracle): the
    large configuration beyond its first two steps"""
    import spsph
    from oracle_binding import Oracle
    prob = spsph.load(deck_dir("refined_bui_spec", ncol=816), "bui")
    assert prob.params.ntotal == 817 * 409 * 3
    dt = prob.blocks[0]["dt"]
    eng, orc = spsph.Engine(prob), Oracle(prob)
    eng.run(1, 0.0, dt, 20)
    orc.run(1, 0.0, dt, 20)
    assert eng.pair_stats() == orc.pair_stats()
    _bitwise(eng.download(), orc.download(), prob.params.ntotal, "refined Bui 1M, 20 steps")
    eng.close()
    orc.close()
