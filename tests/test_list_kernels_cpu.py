"""Host emulation of the device's list-walking kernels (k_free_surface, k_xsph_marks, k_fs_normals of
csrc/step_kernels.cuh) against the oracle, without a GPU: tests/native/list_kernels_host.cpp compiles the kernels
for the host and runs them thread by thread on gather lists built HERE from the oracle's pair list, in the layout the
device's fill pass writes (species-sorted arrays, warp-sliced ELL, reference orientation of the cross-species
gradients, own-perspective gradients of the node-node list, entries in traversal order)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


class EmuArgs(C.Structure):
    _fields_ = [("nnode", C.c_int32), ("nstress", C.c_int32), ("ndummy", C.c_int32), ("skf", C.c_int32),
                ("growth_mode", C.c_int32), ("pad", C.c_int32), ("pi", C.c_double),
                ("growth_ka", C.c_uint64), ("growth_kb", C.c_uint64),
                ("order", C.c_void_p * 3), ("cell", C.c_void_p * 3), ("pos", C.c_void_p * 3), ("h", C.c_void_p * 3),
                ("pos_of", C.c_void_p),
                ("idx0", C.c_void_p), ("idxC", C.c_void_p), ("idxD", C.c_void_p), ("off0", C.c_void_p),
                ("offC", C.c_void_p), ("offD", C.c_void_p), ("n0", C.c_void_p), ("n1", C.c_void_p),
                ("gx0", C.c_void_p), ("gy0", C.c_void_p), ("gxC", C.c_void_p), ("gyC", C.c_void_p),
                ("x", C.c_void_p), ("mass", C.c_void_p), ("rho", C.c_void_p), ("hsml", C.c_void_p),
                ("bc_or_not", C.c_void_p), ("covered", C.c_void_p), ("fs_normal", C.c_void_p)]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "listk.so")
    r = subprocess.run(["g++", "-O1", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-fPIC", "-shared", "-w",
                        "-D__noinline__=", "-I/usr/local/cuda/include",
                        "-I" + os.path.join(ROOT, "stress-particle-sph_b200", "csrc"), "-I" + os.path.join(ROOT, "include"),
                        "-o", so, os.path.join(ROOT, "tests", "native", "list_kernels_host.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(so)


def _ell(entries, nslots, fields):
    """entries[slot] = list of tuples -> (counts, offsets per slice, one array per field) in the warp-sliced layout"""
    cnt = np.array([len(e) for e in entries], np.int32)
    nsl = nslots // 32
    rows = cnt.reshape(nsl, 32).max(axis=1)
    off = np.zeros(nsl, np.int32)
    off[1:] = np.cumsum(rows[:-1] * 32)
    total = int((rows * 32).sum())
    out = [np.zeros(max(total, 1), dt) for dt in fields]
    for t, ent in enumerate(entries):
        base = off[t // 32] + (t & 31)
        for e, tup in enumerate(ent):
            for a, v in zip(out, tup):
                a[base + 32 * e] = v
    return cnt, off, out


def build(prob, cells, pairs, x_built):
    """the device's per-step structures from the oracle's grid cells and ordered pair list"""
    p = prob.params
    nn, ns, nd, nt, n2 = p.nnode, p.nstress, p.ndummy, p.ntotal, p.ntotal2
    ids = np.arange(n2)
    species = np.where(ids < nn, 0, np.where(ids < nt, 1, 2))
    cell0 = cells.astype(np.int64) - 1  # reference cell id is 1-based, 0 marks an out-of-domain particle
    order, cell, pos, hh = [], [], [], []
    pos_of = np.zeros(n2, np.int32)
    hs = prob.arrays["hsml"]
    for sp in range(3):
        mine = ids[species == sp]
        inside = mine[cell0[mine] >= 0]
        inside = inside[np.lexsort((inside, cell0[inside]))]  # by (cell, particle number)
        o = np.concatenate([inside, mine[cell0[mine] < 0]]).astype(np.int32)
        pos_of[o] = np.arange(len(o), dtype=np.int32)
        order.append(o)
        cell.append(np.where(cell0[o] >= 0, cell0[o], -1).astype(np.int32))
        pos.append(np.ascontiguousarray(x_built[o]))
        hh.append(np.ascontiguousarray(hs[o]))
    nnp, nsp = (nn + 31) // 32 * 32, (ns + 31) // 32 * 32
    slot = np.where(ids < nn, pos_of, nnp + pos_of)  # thread slot of a velocity / stress particle
    l0 = [[] for _ in range(nnp + nsp)]
    l1 = [[] for _ in range(nnp + nsp)]
    for i, j, ty, gx, gy in zip(pairs["pair_i"], pairs["pair_j"], pairs["pint_type"], pairs["dwdx"], pairs["dwdy"]):
        i, j = int(i) - 1, int(j) - 1
        if ty == 1:            # i stress particle, j velocity particle: reference orientation on both sides
            l0[slot[i]].append((j, gx, gy))
            l0[slot[j]].append((i, gx, gy))
        elif ty in (6, 9):     # i wall particle
            l0[slot[j]].append((i, gx, gy))
        elif ty == 3:          # velocity - velocity: own-perspective gradient
            l1[slot[i]].append((j, gx, gy))
            l1[slot[j]].append((i, -gx, -gy))
        elif ty == 2:          # stress - stress: ids only (the gradient is re-evaluated)
            l1[slot[i]].append((j, 0.0, 0.0))
            l1[slot[j]].append((i, 0.0, 0.0))
    f3 = (np.int32, np.float32, np.float32)
    n0, off0, (idx0, gx0, gy0) = _ell(l0, nnp + nsp, f3)
    # list C (velocity particles) and list D (stress particles) have separate storage; offsets are per global slice
    lc = [l1[t] if t < nnp else [] for t in range(nnp + nsp)]
    ld = [l1[t] if t >= nnp else [] for t in range(nnp + nsp)]
    n1 = np.array([len(e) for e in l1], np.int32)
    _, offC, (idxC, gxC, gyC) = _ell(lc, nnp + nsp, f3)
    _, offD, (idxD, _, _) = _ell(ld, nnp + nsp, f3)
    return dict(order=order, cell=cell, pos=pos, h=hh, pos_of=pos_of, idx0=idx0, gx0=gx0, gy0=gy0, off0=off0, n0=n0,
                idxC=idxC, gxC=gxC, gyC=gyC, offC=offC, idxD=idxD, offD=offD, n1=n1)


def growth_rule(prob, cells, pairs, m_before, npairs):
    """(mode, ka, kb) of the device's GrowthRule (grid_kernels.cuh): 1 = every pair is new (first step, reversed),
    0 = no growth (forward), 2 = split after the pair with creation index m_before, which is the LAST pair of the
    traversal order (new pairs first and reversed, then the old ones ascending); keys = (cell, species, id)"""
    if m_before == 0:
        return 1, 0, 0
    if npairs <= m_before:
        return 0, 0, 0
    p = prob.params

    def key(i):  # i: 1-based particle number
        i = int(i) - 1
        sp = 0 if i < p.nnode else (1 if i < p.ntotal else 2)
        return ((int(cells[i]) - 1) << 34) | (sp << 32) | i
    k1, k2 = key(pairs["pair_i"][-1]), key(pairs["pair_j"][-1])
    return 2, min(k1, k2), max(k1, k2)


def _args(prob, b, rule, x, bc, cov, normal, keep):
    p = prob.params
    a = EmuArgs()
    mode, a.growth_ka, a.growth_kb = rule
    a.nnode, a.nstress, a.ndummy, a.skf, a.growth_mode, a.pi = p.nnode, p.nstress, p.ndummy, p.skf, mode, p.pi
    ptr = lambda arr: (keep.append(arr), arr.ctypes.data)[1]  # noqa: E731
    for s in range(3):
        a.order[s], a.cell[s], a.pos[s], a.h[s] = ptr(b["order"][s]), ptr(b["cell"][s]), ptr(b["pos"][s]), ptr(b["h"][s])
    for k in ("pos_of", "idx0", "idxC", "idxD", "off0", "offC", "offD", "n0", "n1", "gx0", "gy0", "gxC", "gyC"):
        setattr(a, k, ptr(b[k]))
    a.x, a.mass, a.rho, a.hsml = ptr(x), ptr(prob.arrays["mass"]), ptr(prob.arrays["rho"]), ptr(prob.arrays["hsml"])
    a.bc_or_not, a.covered, a.fs_normal = ptr(bc), ptr(cov), ptr(normal)
    return a


def _sl_free_right(dx, **kw):
    from spsph import decks
    s = decks.strain_localisation_spec(dx=dx, maxtimestep=1000)
    s["props"] = [2, 2, 8.e07, 0.25, 1., 2.e3, 1.5e5, -8.e06, 0., 50., 1., 1]
    s["segments"] = [g for g in s["segments"] if not (g[0] == 0.5 and g[2] == 0.5)]
    s.update(kw)
    return s


@pytest.mark.parametrize("xsph", [False, True])
def test_free_surface_kernels_match_oracle(emu, tmp_path, xsph):
    """k_xsph_marks + k_free_surface + k_fs_normals on the host == the oracle's XSPH marks, get_nodes_on_free_surface
    classification (every velocity and stress particle) and step-4 normals (every marked velocity particle), for the
    fully reversed first step and for forward steps, on the strain-localisation sample with a free right side"""
    import spsph
    from spsph import decks
    from oracle_binding import Oracle, lib
    decks.write_deck(str(tmp_path), _sl_free_right(0.025, ifsigman=1, xsph=xsph))
    prob = spsph.load(str(tmp_path), "sl")
    p = prob.params
    dt = prob.blocks[0]["dt"]
    orc = Oracle(prob)
    L = lib()
    L.oracle_debug_grid.restype = None
    L.oracle_debug_surface.restype = None
    t, checked, marked_total = 0.0, 0, 0
    for step in range(1, 41):
        before = orc.download()
        orc.step(step, t, dt)
        t = t + dt
        if step not in (1, 2, 3, 10, 25, 40):
            continue
        after = orc.download()
        cells = np.zeros(p.ntotal2, np.int32)
        mb, npairs = C.c_int64(), C.c_int64()
        L.oracle_debug_grid(C.c_void_p(orc.h), cells.ctypes.data_as(C.c_void_p), C.byref(mb), C.byref(npairs))
        pairs = orc.pairs()
        rule = growth_rule(prob, cells, pairs, mb.value, npairs.value)
        normal_o = np.zeros((p.ntotal, 2))
        subset_o = np.zeros(p.ntotal)
        L.oracle_debug_surface(C.c_void_p(orc.h), normal_o.ctypes.data_as(C.c_void_p), subset_o.ctypes.data_as(C.c_void_p))
        b = build(prob, cells, pairs, before["x"])
        keep = []
        bc = before["bc_or_not"].astype(np.int32).copy()
        cov = np.zeros(p.ntotal, np.int32)
        normal_d = np.zeros((p.nnode, 2))
        x_now = np.ascontiguousarray(after["x"])  # inside approach: no re-seating after the classification
        a = _args(prob, b, rule, x_now, bc, cov, normal_d, keep)
        if xsph:
            emu.emu_xsph_marks(C.byref(a))
        emu.emu_free_surface(C.byref(a))
        assert np.array_equal(bc, after["bc_or_not"]), f"step {step}: {int((bc != after['bc_or_not']).sum())} flags differ"
        emu.emu_fs_normals(C.byref(a))
        marked = np.nonzero(after["bc_or_not"][:p.nnode] == 2)[0]
        assert np.array_equal(normal_d[marked], normal_o[marked], equal_nan=True), f"step {step}: normals differ"
        checked += 1
        marked_total += len(marked)
    assert checked >= 4 and marked_total > 50


@pytest.mark.parametrize("mode_kw", [dict(mode="vel_vector"), dict(mode="outside"), dict(mode="inside", npoints=2)])
def test_free_surface_classification_bui(emu, tmp_path, mode_kw):
    """k_free_surface on the host == the oracle's classification of every velocity and stress particle of the Bui
    column (wall partners, stress-stress gradients re-evaluated from the build positions), on the positions the
    reference classifies on: after the position update, before shift_stress_points re-seats the stress particles"""
    import spsph
    from spsph import decks
    from oracle_binding import Oracle, lib
    decks.write_deck(str(tmp_path), decks.bui_spec(maxtimestep=1000, **mode_kw))
    prob = spsph.load(str(tmp_path), "bui")
    p = prob.params
    dt = prob.blocks[0]["dt"]
    orc = Oracle(prob)
    L = lib()
    L.oracle_debug_grid.restype = None
    L.oracle_debug_x_fs.restype = None
    long_run = mode_kw.get("mode") == "vel_vector"  # run on until the pair list has grown twice (split traversal)
    t, checked, grown = 0.0, 0, 0
    cells = np.zeros(p.ntotal2, np.int32)
    mb, npairs = C.c_int64(), C.c_int64()
    for step in range(1, 341 if long_run else 61):
        before = orc.download()
        orc.step(step, t, dt)
        t = t + dt
        L.oracle_debug_grid(C.c_void_p(orc.h), cells.ctypes.data_as(C.c_void_p), C.byref(mb), C.byref(npairs))
        growth = mb.value > 0 and npairs.value > mb.value
        if not (step in (1, 2, 30, 60) or (growth and grown < 2)):
            continue
        grown += growth
        after = orc.download()
        pairs = orc.pairs()
        rule = growth_rule(prob, cells, pairs, mb.value, npairs.value)
        x_fs = np.zeros((p.ntotal2, 2))
        L.oracle_debug_x_fs(C.c_void_p(orc.h), x_fs.ctypes.data_as(C.c_void_p))
        b = build(prob, cells, pairs, before["x"])
        keep = []
        bc = before["bc_or_not"].astype(np.int32).copy()
        cov = np.zeros(p.ntotal, np.int32)
        a = _args(prob, b, rule, x_fs, bc, cov, np.zeros((p.nnode, 2)), keep)
        emu.emu_free_surface(C.byref(a))
        assert np.array_equal(bc, after["bc_or_not"]), (
            f"step {step} (growth rule {rule[0]}): {int((bc != after['bc_or_not']).sum())} flags differ")
        assert (bc == 2).sum() > 50
        checked += 1
    assert checked >= 4 and (grown == 2 or not long_run)
