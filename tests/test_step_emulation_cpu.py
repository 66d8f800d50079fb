"""The CUDA engine's time step emulated on the host, against the oracle, without a GPU.

tests/native/make_engine_host.py rewrites csrc/spsph_engine.cu (kernel launches -> emu_launch) and g++ compiles it with
tests/native/cuda_host_emu.h: every kernel thread of the per-particle kernels (k_rk_begin, the interpolation and
gradient sweeps with their fused RK4 / constitutive / boundary-condition epilogues, k_artvisc, k_art_force, k_move,
k_shift, k_free_surface, k_xsph_marks, k_fs_normals, the upload / download conversions) runs serially on the host,
through the engine's own C-ABI and its own step sequence (step_impl). The neighbour build (k_count / k_fill: block
scans, shared-memory queues) is not emulated: each step's sorted arrays and gather lists are built HERE from the
oracle's pair list in the layout the fill pass writes and handed over with spsph_emu_set_lists. The state after the
emulated steps must equal the oracle's bit for bit (CUDA's libm is not involved on the host, so also for the paths
that carry a 1e-9 tolerance on the GPU)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_list_kernels_cpu import _ell, growth_rule  # noqa: E402


class EmuLists(C.Structure):
    _fields_ = [("n_pairs", C.c_int64), ("growth_mode", C.c_int32), ("pad", C.c_int32),
                ("growth_ka", C.c_uint64), ("growth_kb", C.c_uint64),
                ("order", C.c_void_p * 3), ("cell", C.c_void_p * 3), ("pos", C.c_void_p * 3), ("h", C.c_void_p * 3),
                ("pos_of", C.c_void_p), ("tot0", C.c_int64), ("totC", C.c_int64), ("totD", C.c_int64),
                ("idx0", C.c_void_p), ("w0", C.c_void_p), ("gx0", C.c_void_p), ("gy0", C.c_void_p),
                ("idxC", C.c_void_p), ("wC", C.c_void_p), ("gxC", C.c_void_p), ("gyC", C.c_void_p),
                ("xC", C.c_void_p), ("yC", C.c_void_p), ("hC", C.c_void_p),
                ("idxD", C.c_void_p), ("wD", C.c_void_p),
                ("off0", C.c_void_p), ("offC", C.c_void_p), ("offD", C.c_void_p), ("n0", C.c_void_p), ("n1", C.c_void_p),
                ("bc_int", C.c_void_p), ("n_int", C.c_void_p), ("if_out", C.c_void_p)]


def _build_emulated(d, extra=()):
    cpp, so = str(d / "engine_host.cpp"), str(d / "libspsph_emu.so")
    subprocess.run([sys.executable, os.path.join(ROOT, "tests", "native", "make_engine_host.py"),
                    os.path.join(ROOT, "stress-particle-sph_b200", "csrc", "spsph_engine.cu"), cpp], check=True,
                   stdout=subprocess.DEVNULL)
    r = subprocess.run(["g++", "-O1", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-fPIC", "-shared", "-w",
                        "-D__noinline__=", "-fno-gnu-unique", *extra, "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "tests", "native"),
                        "-I" + os.path.join(ROOT, "stress-particle-sph_b200", "csrc"), "-I" + os.path.join(ROOT, "include"),
                        "-o", so, cpp, "-ldl"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return so


@pytest.fixture(scope="module")
def emu_engine(tmp_path_factory):
    """spsph.Engine bound to the host-emulated build of the CUDA engine"""
    d = tmp_path_factory.mktemp("emu_engine")
    cpp, so = str(d / "engine_host.cpp"), str(d / "libspsph_emu.so")
    subprocess.run([sys.executable, os.path.join(ROOT, "tests", "native", "make_engine_host.py"),
                    os.path.join(ROOT, "stress-particle-sph_b200", "csrc", "spsph_engine.cu"), cpp], check=True,
                   stdout=subprocess.DEVNULL)
    r = subprocess.run(["g++", "-O1", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-fPIC", "-shared", "-w",
                        "-D__noinline__=", "-fno-gnu-unique", "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "tests", "native"),
                        "-I" + os.path.join(ROOT, "stress-particle-sph_b200", "csrc"), "-I" + os.path.join(ROOT, "include"),
                        "-o", so, cpp, "-ldl"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    import spsph.engine as E
    saved = (E._lib, E._CUDA_SO)
    E._lib, E._CUDA_SO = None, so
    lib = E.cuda_lib()
    lib.spsph_emu_set_lists.argtypes = [C.c_void_p, C.POINTER(EmuLists)]
    yield E
    E._lib, E._CUDA_SO = saved


def build_step(prob, cells, pairs, x_built, rule, npairs, hsml=None):
    """this step's species-sorted arrays and gather lists in the device layout, from the oracle's grid cells and
    ordered pair list (the payloads of csrc/grid_kernels.cuh::k_fill)"""
    p = prob.params
    nn, ns, nt, n2 = p.nnode, p.nstress, p.ntotal, p.ntotal2
    ids = np.arange(n2)
    species = np.where(ids < nn, 0, np.where(ids < nt, 1, 2))
    cell0 = cells.astype(np.int64) - 1
    hs = prob.arrays["hsml"] if hsml is None else hsml  # sle = 2: the smoothing length the lists were built with
    order, cell, pos, hh = [], [], [], []
    pos_of = np.zeros(n2, np.int32)
    for sp in range(3):
        mine = ids[species == sp]
        inside = mine[cell0[mine] >= 0]
        inside = inside[np.lexsort((inside, cell0[inside]))]
        o = np.concatenate([inside, mine[cell0[mine] < 0]]).astype(np.int32)
        pos_of[o] = np.arange(len(o), dtype=np.int32)
        order.append(o)
        cell.append(np.where(cell0[o] >= 0, cell0[o], -1).astype(np.int32))
        pos.append(np.ascontiguousarray(x_built[o]))
        hh.append(np.ascontiguousarray(hs[o]))
    nnp, nsp = (nn + 31) // 32 * 32, (ns + 31) // 32 * 32
    slot = np.where(ids < nn, pos_of, nnp + pos_of)
    l0 = [[] for _ in range(nnp + nsp)]
    lc = [[] for _ in range(nnp + nsp)]
    ld = [[] for _ in range(nnp + nsp)]
    bc_int = np.zeros(nn, np.int32)
    X = x_built
    for i, j, ty, w, gx, gy in zip(pairs["pair_i"], pairs["pair_j"], pairs["pint_type"], pairs["w"], pairs["dwdx"],
                                   pairs["dwdy"]):
        i, j = int(i) - 1, int(j) - 1
        if ty == 1:
            l0[slot[i]].append((j, w, gx, gy))
            l0[slot[j]].append((i, w, gx, gy))
        elif ty in (6, 9):
            l0[slot[j]].append((i, w, gx, gy))
            if ty == 6:
                bc_int[j] = 1
        elif ty == 3:
            hm = np.float32(0.5 * (hs[i] + hs[j]))
            lc[slot[i]].append((j, w, gx, gy, np.float32(X[i, 0] - X[j, 0]), np.float32(X[i, 1] - X[j, 1]), hm))
            lc[slot[j]].append((i, w, -gx, -gy, np.float32(X[j, 0] - X[i, 0]), np.float32(X[j, 1] - X[i, 1]), hm))
        elif ty == 2:
            ld[slot[i]].append((j, w))
            ld[slot[j]].append((i, w))
    f32 = np.float32
    n0, off0, (idx0, w0, gx0, gy0) = _ell(l0, nnp + nsp, (np.int32, f32, f32, f32))
    nC, offC, (idxC, wC, gxC, gyC, xC, yC, hC) = _ell(lc, nnp + nsp, (np.int32, f32, f32, f32, f32, f32, f32))
    nD, offD, (idxD, wD) = _ell(ld, nnp + nsp, (np.int32, f32))
    n1 = (nC + nD).astype(np.int32)
    n_int = nC[pos_of[:nn]].astype(np.float32)  # node-node interaction count, indexed by particle number

    def total(cnt):
        return int((cnt.reshape(-1, 32).max(axis=1) * 32).sum())
    keep = dict(order=order, cell=cell, pos=pos, h=hh, pos_of=pos_of, idx0=idx0, w0=w0, gx0=gx0, gy0=gy0, idxC=idxC,
                wC=wC, gxC=gxC, gyC=gyC, xC=xC, yC=yC, hC=hC, idxD=idxD, wD=wD, off0=off0, offC=offC, offD=offD, n0=n0,
                n1=n1, bc_int=bc_int, n_int=n_int, if_out=(cells == 0).astype(np.int32))
    E = EmuLists()
    E.n_pairs = int(npairs)
    E.growth_mode, E.growth_ka, E.growth_kb = rule
    for s in range(3):
        E.order[s], E.cell[s], E.pos[s], E.h[s] = (order[s].ctypes.data, cell[s].ctypes.data, pos[s].ctypes.data,
                                                    hh[s].ctypes.data)
    E.tot0, E.totC, E.totD = total(n0), total(nC), total(nD)
    for k in ("pos_of", "idx0", "w0", "gx0", "gy0", "idxC", "wC", "gxC", "gyC", "xC", "yC", "hC", "idxD", "wD", "off0",
              "offC", "offD", "n0", "n1", "bc_int", "n_int", "if_out"):
        setattr(E, k, keep[k].ctypes.data)
    return E, keep


STATE_KEYS = ("x", "vel", "stress", "rho", "hsml", "internal_vars", "f_drucker", "displ", "x_10", "disp_10", "n_int",
              "bc_int", "if_out_domain", "bc_or_not")


def run_lockstep(emu_engine, prob, nsteps, check_at, label):
    from oracle_binding import Oracle, lib
    p = prob.params
    dt = prob.blocks[0]["dt"]
    orc = Oracle(prob)
    eng = emu_engine.Engine(prob)
    L = lib()
    L.oracle_debug_grid.restype = None
    cells = np.zeros(p.ntotal2, np.int32)
    mb, npairs = C.c_int64(), C.c_int64()
    t = 0.0
    for step in range(1, nsteps + 1):
        before = orc.download()
        x_before = before["x"]
        orc.step(step, t, dt)
        L.oracle_debug_grid(C.c_void_p(orc.h), cells.ctypes.data_as(C.c_void_p), C.byref(mb), C.byref(npairs))
        pairs = orc.pairs()
        rule = growth_rule(prob, cells, pairs, mb.value, npairs.value)
        E, keep = build_step(prob, cells, pairs, x_before, rule, npairs.value, before["hsml"])
        assert emu_engine.cuda_lib().spsph_emu_set_lists(eng.h, C.byref(E)) == 0
        eng.step(step, t, dt)
        t = t + dt
        if step in check_at:
            a, b = eng.download(), orc.download()
            nt = p.ntotal
            for k in STATE_KEYS:
                x, y = a[k], b[k]
                if k in ("x", "vel", "stress"):
                    x, y = x[:nt], y[:nt]
                assert np.array_equal(x, y), (f"{label}, step {step}: {k} differs from the oracle in "
                                              f"{int((x != y).sum())} entries")
    final = eng.download()
    eng.close()
    return final


def test_emulated_step_bui(emu_engine, tmp_path):
    import spsph
    from spsph import decks
    decks.write_deck(str(tmp_path), decks.bui_spec(dx=0.2, maxtimestep=1000))
    prob = spsph.load(str(tmp_path), "bui")
    run_lockstep(emu_engine, prob, 12, (1, 2, 3, 12), "Bui column, dx = 0.2")


def _sl(dx=0.05, **kw):
    from spsph import decks
    s = decks.strain_localisation_spec(dx=dx, maxtimestep=1000)
    ncrit = kw.pop("ncrit", 2)
    # yield stress far below the wave stress so that the sample flows within a few steps
    s["props"] = [2, ncrit, 8.e07, 0.25, 1., 2.e3, kw.pop("yield0", 2.e4), -8.e06, kw.pop("frict", 0.), 50.,
                  kw.pop("delta", 1.), kw.pop("nflow", 1)]
    if kw.pop("free_right", False):
        s["segments"] = [g for g in s["segments"] if not (g[0] == 0.5 and g[2] == 0.5)]
    s.update(kw)
    return s


def _vs(dx=1.0, **kw):
    from spsph import decks
    s = decks.vertical_slope_spec(dx=dx, maxtimestep=1000)
    if kw.pop("free_right", False):
        s["segments"] = [g for g in s["segments"] if not (g[0] == 10. and g[2] == 10.)]
    s.update(kw)
    return s


def _bui(dx=0.2, **kw):
    from spsph import decks
    extra = {k: kw.pop(k) for k in list(kw) if k not in ("mode", "npoints")}
    s = decks.bui_spec(dx=dx, maxtimestep=1000, **kw)
    s.update(extra)
    return s


def _sine(s):
    s["bcs"] = list(s["bcs"])
    s["bcs"][4] = (5, 6, 0, 1.0, 0.5, 3000., 0.3, 2.e-4)
    return s


# (label, variant, deck spec, steps): the shipped problems and every option of oracle/ref_cases.py the device
# implements, at a resolution the Python list builder handles in seconds
EMU_CASES = [
    ("bui", "bui", lambda: _bui(), 12),
    ("bui_outside", "bui", lambda: _bui(mode="outside"), 8),
    ("bui_outside_sp1", "bui", lambda: _bui(mode="outside", npoints=1), 8),
    ("bui_outside_sp3", "bui", lambda: _bui(mode="outside", npoints=3), 8),
    ("bui_standard", "bui", lambda: _bui(mode="standard"), 8),
    ("bui_shift5", "bui", lambda: _bui(shift_update=5), 12),
    ("bui_plane_stress", "bui", lambda: _bui(ntype_solid=1), 8),
    ("bui_sml15", "bui", lambda: _bui(sml=1.5), 6),
    ("bui_out_domain", "bui", lambda: _bui(domain=[-10, -10, 4.00001, 41]), 12),
    ("bui_art_stress", "bui", lambda: _bui(art_stress=True), 6),
    ("bui_gauss", "bui", lambda: _bui(skf=2), 6),
    ("bui_quintic", "bui", lambda: _bui(skf=3), 6),
    ("bui_cont_density", "bui", lambda: _bui(cont_density=True), 8),
    ("vs", "vs", lambda: _vs(), 10),
    ("vs_cont_density_sle2", "vs", lambda: _vs(cont_density=True), 8),
    ("sl_cont_density_sle2", "sl", lambda: _sl(cont_density=True), 20),
    ("vs_sp2", "vs", lambda: _vs(npoints=2), 6),
    ("vs_standard", "vs", lambda: _vs(standard=True) if False else dict(_vs(), sp_sph=False), 6),
    ("vs_sigman", "vs", lambda: _vs(free_right=True, ifsigman=1, update_x=True), 12),
    ("sl_von_mises", "sl", lambda: _sl(), 20),
    ("sl_tresca", "sl", lambda: _sl(ncrit=1), 20),
    ("sl_mohr_coulomb", "sl", lambda: _sl(ncrit=3, frict=20., yield0=8.e3), 20),
    ("sl_dp_perzyna", "sl", lambda: _sl(ncrit=4, frict=20.), 20),
    ("sl_vm_expflow", "sl", lambda: _sl(nflow=2, delta=1.5), 20),
    ("sl_vm_powflow", "sl", lambda: _sl(delta=1.5), 20),
    ("sl_sine_bc", "sl", lambda: _sine(_sl()), 20),
    ("sl_sigman", "sl", lambda: _sl(free_right=True, ifsigman=1), 20),
    ("sl_xsph", "sl", lambda: _sl(free_right=True, xsph=True, yield0=5.e3), 20),
    ("sl_sigman_xsph", "sl", lambda: _sl(free_right=True, ifsigman=1, xsph=True, yield0=5.e3), 20),
    # the three decks exactly as shipped (BASELINE configs 0-2), first steps
    ("shipped_bui", "bui", lambda: __import__("spsph").decks.bui_spec(maxtimestep=100), 3),
    ("shipped_vs", "vs", lambda: __import__("spsph").decks.vertical_slope_spec(maxtimestep=100), 4),
    ("shipped_sl", "sl", lambda: __import__("spsph").decks.strain_localisation_spec(maxtimestep=100), 2),
]


# the default suite runs a representative third (the stand-alone and reference-golden tests cover every option again);
# SPSPH_EMU_ALL=1 runs all of them
_DEFAULT = {"bui", "bui_outside_sp3", "bui_plane_stress", "bui_out_domain", "bui_art_stress", "bui_cont_density", "vs",
            "vs_sigman", "sl_tresca", "sl_mohr_coulomb", "sl_vm_expflow", "sl_sigman_xsph", "sl_cont_density_sle2",
            "shipped_bui"}
if not os.environ.get("SPSPH_EMU_ALL"):
    EMU_CASES = [c for c in EMU_CASES if c[0] in _DEFAULT]


@pytest.mark.parametrize("label,variant,spec_fn,nsteps", EMU_CASES, ids=[c[0] for c in EMU_CASES])
def test_emulated_step_matches_oracle(emu_engine, tmp_path, label, variant, spec_fn, nsteps):
    import spsph
    from spsph import decks
    decks.write_deck(str(tmp_path), spec_fn())
    prob = spsph.load(str(tmp_path), variant)
    final = run_lockstep(emu_engine, prob, nsteps, (1, 2, nsteps), label)
    if label.startswith("sl_"):
        assert (final["internal_vars"][prob.params.nnode:prob.params.ntotal, 0] > 0).sum() >= 5, "no plastic flow reached"


def run_standalone(emu_engine, prob, nsteps, check_at, label, pairs_at=()):
    """the emulated engine on its own: the neighbour build runs too (k_cell_id, k_scatter, k_rank, k_count, k_fill /
    k_fill_scan, k_growth_threshold thread by thread; bounding box, prefix scans and slice widths by host loops)"""
    from oracle_binding import Oracle
    p = prob.params
    dt = prob.blocks[0]["dt"]
    orc = Oracle(prob)
    eng = emu_engine.Engine(prob)
    t = 0.0
    for step in range(1, nsteps + 1):
        orc.step(step, t, dt)
        eng.step(step, t, dt)
        t = t + dt
        if step in pairs_at:
            pa, pb = eng.pairs(), orc.pairs()
            assert len(pa["pair_i"]) == len(pb["pair_i"]), f"{label}, step {step}: pair count"
            for f in ("pair_i", "pair_j", "pint_type"):
                assert np.array_equal(pa[f], pb[f]), f"{label}, step {step}: {f} differs"
            for f in ("w", "dwdx", "dwdy"):
                assert np.array_equal(pa[f].view(np.uint32), pb[f].view(np.uint32)), f"{label}, step {step}: {f} differs"
            assert eng.pair_stats() == orc.pair_stats()
        if step in check_at:
            a, b = eng.download(), orc.download()
            for k in STATE_KEYS:
                x, y = a[k], b[k]
                if k in ("x", "vel", "stress"):
                    x, y = x[:p.ntotal], y[:p.ntotal]
                assert np.array_equal(x, y), (f"{label}, step {step}: {k} differs from the oracle in "
                                              f"{int((x != y).sum())} entries")
    eng.close()


def test_emulated_engine_standalone_bui(emu_engine, tmp_path):
    import spsph
    from spsph import decks
    decks.write_deck(str(tmp_path), decks.bui_spec(dx=0.2, maxtimestep=1000))
    prob = spsph.load(str(tmp_path), "bui")
    run_standalone(emu_engine, prob, 6, (1, 2, 6), "Bui column, dx = 0.2, stand-alone", pairs_at=(1, 2, 3))


TINY = [("vs_2x2", "vs", lambda d: d.vertical_slope_spec(dx=10.0, maxtimestep=100), 30),
        ("vs_3x3", "vs", lambda d: d.vertical_slope_spec(dx=5.0, maxtimestep=100), 30),
        ("bui_3x2", "bui", lambda d: d.bui_spec(dx=2.0, maxtimestep=100), 50),
        ("bui_5x3_inside_sp3", "bui", lambda d: d.bui_spec(dx=1.0, maxtimestep=100, mode="inside", npoints=3), 40),
        ("bui_5x3_standard", "bui", lambda d: d.bui_spec(dx=1.0, maxtimestep=100, mode="standard"), 40),
        ("sl_3x5", "sl", lambda d: d.strain_localisation_spec(dx=0.25, maxtimestep=100), 50)]


@pytest.mark.parametrize("label,variant,spec_fn,nsteps", TINY, ids=[c[0] for c in TINY])
def test_emulated_engine_tiny_problems(emu_engine, tmp_path, label, variant, spec_fn, nsteps):
    """ragged edge of the layout: 4 - 45 particles, i.e. one partly filled 32-particle slice per species, list rows
    shorter than one streaming group, cells with a single particle"""
    import spsph
    from spsph import decks
    decks.write_deck(str(tmp_path), spec_fn(decks))
    prob = spsph.load(str(tmp_path), variant)
    run_standalone(emu_engine, prob, nsteps, (1, nsteps), label, pairs_at=(1, 2))


SIMT_CASES = [("bui", "bui", lambda: _bui(), 5),
              ("bui_inside_sp2", "bui", lambda: _bui(mode="inside", npoints=2), 4),
              ("bui_sml15", "bui", lambda: _bui(sml=1.5), 3),
              ("sl_sigman_xsph", "sl", lambda: _sl(free_right=True, ifsigman=1, xsph=True, yield0=5.e3), 4)]


@pytest.fixture(scope="module")
def emu_engine_simt(tmp_path_factory):
    so = _build_emulated(tmp_path_factory.mktemp("emu_simt"), ("-DSPSPH_EMU_SIMT",))
    import spsph.engine as E
    saved = (E._lib, E._CUDA_SO)
    E._lib, E._CUDA_SO = None, so
    E.cuda_lib()
    yield E
    E._lib, E._CUDA_SO = saved


@pytest.mark.parametrize("label,variant,spec_fn,nsteps", SIMT_CASES, ids=[c[0] for c in SIMT_CASES])
def test_emulated_engine_simt_mode(emu_engine_simt, tmp_path, label, variant, spec_fn, nsteps):
    """-DSPSPH_EMU_SIMT: the threads of a block are fibers that switch at every warp / block collective, so the device
    code runs with NO host replacement except the inline PTX (cp.async = memcpy, 256-bit loads / stores): the real
    ell_stream ring (cp.async groups, tail rows, __syncwarp), shuffles, votes, block scans, the bounding-box reduction,
    the slice-width maxima of k_count and k_pair_stats. About 1.5 s per step whatever the problem size (the fixed grids
    of the scans and reductions create ~1.5 M fibers per step)."""
    import spsph
    from spsph import decks
    decks.write_deck(str(tmp_path), spec_fn())
    prob = spsph.load(str(tmp_path), variant)
    run_standalone(emu_engine_simt, prob, nsteps, (1, nsteps), label + ", SIMT emulation", pairs_at=(1, 2))


VARIANT_FLAGS = {"t64": ("-DSPSPH_SWEEP_T=64", "-DSPSPH_MINB=8"),
                 "pipe2": ("-DSPSPH_ELL_PIPE=1", "-DSPSPH_ELL_SUB=2", "-DSPSPH_A_SUB=2"),
                 "pipe4_t32": ("-DSPSPH_ELL_PIPE=1", "-DSPSPH_SWEEP_T=32", "-DSPSPH_MINB=16")}


@pytest.fixture(scope="module", params=list(VARIANT_FLAGS))
def emu_engine_simt_variant(request, tmp_path_factory):
    so = _build_emulated(tmp_path_factory.mktemp("emu_simt_" + request.param),
                         ("-DSPSPH_EMU_SIMT",) + VARIANT_FLAGS[request.param])
    import spsph.engine as E
    saved = (E._lib, E._CUDA_SO)
    E._lib, E._CUDA_SO = None, so
    E.cuda_lib()
    yield E
    E._lib, E._CUDA_SO = saved


@pytest.mark.parametrize("label,variant,spec_fn,nsteps", SIMT_CASES[::3], ids=[c[0] for c in SIMT_CASES[::3]])
def test_emulated_engine_simt_mode_build_variants(emu_engine_simt_variant, tmp_path, label, variant, spec_fn, nsteps):
    """the build-time variants tools/variant_timing.sh times on the GPU give the same bits on the SIMT emulation, where the
    real ell_stream runs: 64- / 32-thread blocks for the pair-sum kernels (block size is a scheduling choice only) and
    the software-pipelined gathers (-DSPSPH_ELL_PIPE=1: partner records of the next part requested before the current
    part is consumed -- same entries, same order)"""
    import spsph
    from spsph import decks
    decks.write_deck(str(tmp_path), spec_fn())
    prob = spsph.load(str(tmp_path), variant)
    run_standalone(emu_engine_simt_variant, prob, nsteps, (1, nsteps), label + ", SIMT emulation, build variant",
                   pairs_at=(1,))


TILE_CASES = [("bui", "bui", lambda: _bui(), 5),
              ("bui_inside_sp2", "bui", lambda: _bui(mode="inside", npoints=2), 3),
              ("sl", "sl", lambda: _sl(), 3)]


@pytest.mark.parametrize("label,variant,spec_fn,nsteps", TILE_CASES, ids=[c[0] for c in TILE_CASES])
def test_emulated_engine_simt_mode_tile_path(emu_engine_simt, tmp_path, monkeypatch, label, variant, spec_fn, nsteps):
    """the cell-tile path (csrc/tile_kernels.cuh, SPSPH_TILE=1) on the SIMT emulation: one-pass build with entry codes,
    shared-memory partner tiles (cp.async = memcpy here), reversed first step and forward steps, walls, XSPH,
    artificial viscosity, CSPM -- state, ordered pair list (re-created on demand from the id lists) and statistics equal
    the oracle's bit for bit, and every step really ran on the tile kernels"""
    import ctypes as C
    import spsph
    from spsph import decks
    monkeypatch.setenv("SPSPH_TILE", "1")
    decks.write_deck(str(tmp_path), spec_fn())
    prob = spsph.load(str(tmp_path), variant)
    counts = []
    real_engine = emu_engine_simt.Engine

    class Counting(real_engine):
        def close(self):
            if getattr(self, "h", None):
                counts.append(self.path_counts())
            super().close()

    monkeypatch.setattr(emu_engine_simt, "Engine", Counting)
    run_standalone(emu_engine_simt, prob, nsteps, (1, nsteps), label + ", SIMT emulation, tile path", pairs_at=(1, 2))
    assert counts and counts[0] == (nsteps, 0), counts
