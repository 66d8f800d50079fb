"""CPU tests (no GPU): the oracle, the host-side reader, the deck writer and the C-ABI surface."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/example_problems"
REF_DIRS = {"bui": "soil_failure_bui_et_al_2008/outside_approach/velocity_vector_update", "vs": "vertical_slope",
            "sl": "strain_localisation_in_soil_sample"}
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.mark.parametrize("kind", ["bui", "vs", "sl"])
@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_regenerated_decks_equal_shipped_inputs(deck_dir, kind):
    import spsph
    a = spsph.load(os.path.join(REF, REF_DIRS[kind]), kind)
    b = spsph.load(deck_dir(kind), kind)
    assert bytes(a.params) == bytes(b.params)
    assert a.blocks == b.blocks
    for k in a.arrays:
        assert np.array_equal(a.arrays[k], b.arrays[k]), k


VARIANT_DIRS = [
    ("bui", "soil_failure_bui_et_al_2008", dict()),
    ("bui", "soil_failure_bui_et_al_2008/outside_approach", dict(mode="outside")),
    ("bui", "soil_failure_bui_et_al_2008/standard_sph", dict(mode="standard")),
    ("bui", "soil_failure_bui_et_al_2008/inside_approach/SP1", dict(mode="inside", npoints=1)),
    ("bui", "soil_failure_bui_et_al_2008/inside_approach/SP2", dict(mode="inside", npoints=2)),
    ("bui", "soil_failure_bui_et_al_2008/inside_approach/SP3", dict(mode="inside", npoints=3)),
    ("vs", "vertical_slope/SP1", dict()), ("vs", "vertical_slope/SP2", dict(npoints=2)),
    ("vs", "vertical_slope/SP3", dict(npoints=3)), ("vs", "vertical_slope/standard", dict(standard=True)),
    ("sl", "strain_localisation_in_soil_sample/SP1", dict()),
    ("sl", "strain_localisation_in_soil_sample/SP2", dict(npoints=2)),
    ("sl", "strain_localisation_in_soil_sample/SP3", dict(npoints=3)),
    ("sl", "strain_localisation_in_soil_sample/standard", dict(standard=True)),
]


@pytest.mark.parametrize("kind,path,kw", VARIANT_DIRS, ids=[p for _, p, _ in VARIANT_DIRS])
@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_all_shipped_input_sets_regenerate(deck_dir, kind, path, kw):
    """all 16 input sets of the reference: the deck writer reproduces particles, parameters and time blocks"""
    import spsph
    d = os.path.join(REF, path)
    if not os.path.exists(os.path.join(d, "input.txt")):
        pytest.skip("no input.txt")
    src = d
    name = {"bui": "co_soil", "vs": "elastic_cut", "sl": "localisation"}[kind]
    if not os.path.exists(os.path.join(d, name + ".dat")):  # top-level Bui input.txt uses the velocity-vector deck
        import shutil
        import tempfile
        src = tempfile.mkdtemp()
        shutil.copy(os.path.join(d, "input.txt"), src)
        base = os.path.join(REF, REF_DIRS[kind])
        for ext in (".dat", ".pts"):
            shutil.copy(os.path.join(base, name + ext), src)
    a = spsph.load(src, kind)
    b = spsph.load(deck_dir(kind, **kw), kind)
    assert bytes(a.params) == bytes(b.params)
    assert [(x["dt"], x["time_end"]) for x in a.blocks] == [(x["dt"], x["time_end"]) for x in b.blocks]
    for k in a.arrays:
        assert np.array_equal(a.arrays[k], b.arrays[k]), k


def test_particle_counts_match_survey(deck_dir):
    """counts derived independently from the input files."""
    import spsph
    want = {"bui": (861, 1722, 348), "vs": (441, 400, 0), "sl": (3321, 3200, 0)}
    for kind, (nn, ns, nd) in want.items():
        p = spsph.load(deck_dir(kind), kind).params
        assert (p.nnode, p.nstress, p.ndummy) == (nn, ns, nd)


def test_pair_counts_match_independent_kdtree(deck_dir):
    """pair census at t = 0 against scipy's KD-tree on the same lattice (strict r < 2*h as main:1353)"""
    import spsph
    from oracle_binding import Oracle
    from scipy.spatial import cKDTree
    for kind in ("bui", "vs", "sl"):
        prob = spsph.load(deck_dir(kind), kind)
        orc = Oracle(prob)
        orc.step(1, 0.0, prob.blocks[0]["dt"])
        pr = orc.pairs()
        x = prob.arrays["x"]
        h = prob.arrays["hsml"][0]
        tree = cKDTree(x)
        cand = tree.query_pairs(2 * h * (1 + 1e-9), output_type="ndarray")
        d = x[cand[:, 0]] - x[cand[:, 1]]
        r = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
        kd = cand[r < 2 * h]
        got = {(min(i, j), max(i, j)) for i, j in zip(pr["pair_i"] - 1, pr["pair_j"] - 1)}
        assert got == {(int(i), int(j)) for i, j in kd}
        # every accepted pair satisfies the reference's criterion, and types follow Pint_Update
        it = prob.arrays["itype"]
        s = it[pr["pair_i"] - 1] + it[pr["pair_j"] - 1]
        tmap = {3: 1, 2: 2, 4: 3, 27: 6, 26: 9, 50: 0}
        assert np.array_equal(pr["pint_type"], np.vectorize(tmap.get)(s))
        t1 = pr["pint_type"] == 1
        assert (it[pr["pair_i"][t1] - 1] == 1).all() and (it[pr["pair_j"][t1] - 1] == 2).all()


def test_oracle_step1_is_reversed_creation_order(deck_dir):
    """the first step walks the list back to front; later steps front to back."""
    import spsph
    from oracle_binding import Oracle
    prob = spsph.load(deck_dir("vs"), "vs")
    dt = prob.blocks[0]["dt"]
    orc = Oracle(prob)
    orc.step(1, 0.0, dt)
    p1 = orc.pairs()
    orc.step(2, dt, dt)
    p2 = orc.pairs()
    # positions do not move in this config (update_x = F): same pair set, opposite order
    assert np.array_equal(p1["pair_i"][::-1], p2["pair_i"]) and np.array_equal(p1["pair_j"][::-1], p2["pair_j"])
    assert np.array_equal(p1["w"][::-1], p2["w"])


def test_oracle_invariants_bui(deck_dir):
    """reference-derived invariants: mass constant, DP stresses inside the yield surface
    after adapt_stress2, Shepard interpolation reproduces constants."""
    import spsph
    from oracle_binding import Oracle
    prob = spsph.load(deck_dir("bui"), "bui")
    p = prob.params
    orc = Oracle(prob)
    orc.run(1, 0.0, prob.blocks[0]["dt"], 50)
    a = orc.download()
    assert np.array_equal(a["mass"], prob.arrays["mass"])
    tanfi, coh = p.props[12], p.props[13]
    alpha2 = tanfi / np.sqrt(9 + 12 * tanfi ** 2)
    kc = 3 * coh / np.sqrt(9 + 12 * tanfi ** 2)
    s = a["stress"][:p.ntotal]
    sm = (s[:, 0] + s[:, 1] + s[:, 3]) / 3
    dev = np.stack([s[:, 0] - sm, s[:, 1] - sm, s[:, 2], s[:, 3] - sm], 1)
    j2 = dev[:, 2] ** 2 + 0.5 * (dev[:, 0] ** 2 + dev[:, 1] ** 2 + dev[:, 3] ** 2)
    yld = -3 * alpha2 * sm + kc
    assert (yld >= -1e-6).all()
    assert (np.sqrt(j2) <= yld * (1 + 1e-9) + 1e-6).all()
    assert np.isfinite(a["x"]).all() and np.isfinite(a["vel"]).all()


def test_oracle_vertical_slope_static_limit(deck_dir):
    """elastic block under ramped gravity with damping settles (|v| -> 0) with sigma_yy ~ -rho*g*depth"""
    import spsph
    from oracle_binding import Oracle
    prob = spsph.load(deck_dir("vs"), "vs")
    p = prob.params
    orc = Oracle(prob)
    orc.run(1, 0.0, prob.blocks[0]["dt"], 2000)
    a = orc.download()
    assert np.abs(a["vel"][:p.nnode]).max() < 1e-3
    sp = slice(p.nnode, p.ntotal)
    y = a["x"][sp, 1]
    xs = a["x"][sp, 0]
    mid = (np.abs(xs - 5.0) < 1.0) & (y < 9.0) & (y > 1.0)
    want = -2000 * 9.81 * (10.0 - y[mid])
    assert np.abs(a["stress"][sp, 1][mid] - want).max() < 0.15 * 2000 * 9.81 * 10


def test_golden_fixtures_match_oracle(deck_dir):
    """tests/golden/*.npz were produced by tests/golden/make_golden.py from this oracle; they pin the oracle
    (and through it the CUDA path) against accidental change. They are NOT reference outputs -- those are
    tests/golden/ref_*.npz, checked by tests/test_reference_pinned_cpu.py."""
    import spsph
    from oracle_binding import Oracle
    files = sorted(f for f in os.listdir(GOLD) if f.endswith("steps.npz")) if os.path.isdir(GOLD) else []
    assert files, "golden fixtures missing"
    for f in files:
        g = np.load(os.path.join(GOLD, f))
        kind, nsteps = str(g["kind"]), int(g["nsteps"])
        prob = spsph.load(deck_dir(kind), kind)
        orc = Oracle(prob)
        orc.run(1, 0.0, prob.blocks[0]["dt"], nsteps)
        a = orc.download()
        nt = prob.params.ntotal
        for k in ("x", "vel", "stress"):
            assert np.array_equal(a[k][:nt], g[k]), (f, k)
        assert np.array_equal(a["internal_vars"][:, 0], g["epsp"])
        assert orc.pair_stats()["npairs"] == int(g["npairs"])


def test_cabi_exports_every_declared_symbol():
    """libspsph_cuda.so loads without a GPU and exports exactly what include/spsph.h declares"""
    import spsph
    hdr = open(os.path.join(ROOT, "include", "spsph.h")).read()
    declared = sorted(set(re.findall(r"\b(spsph_[a-z_]+)\s*\(", hdr)))
    assert declared
    lib = spsph.cuda_lib()
    for name in declared:
        assert hasattr(lib, name), name
    from spsph import engine
    assert sorted(engine.EXPORTS) == declared
    assert b"sm_100a" in lib.spsph_version()


def test_params_layout_matches_header():
    import spsph
    from spsph import _abi
    prob = spsph.load  # noqa: F841
    # struct_bytes is written by the C++ reader (sizeof(spsph_params)); load() raises on mismatch
    import tempfile
    from spsph import decks
    d = tempfile.mkdtemp()
    decks.write_deck(d, decks.vertical_slope_spec())
    p = spsph.load(d, "vs").params
    assert p.struct_bytes == C.sizeof(_abi.Params)


def test_engine_fails_loudly_without_gpu(deck_dir):
    """no CPU fallback: on a box without CUDA the product path must raise, not compute"""
    import spsph
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    prob = spsph.load(deck_dir("vs"), "vs")
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        spsph.Engine(prob)


def test_product_never_touches_oracle():
    pkg = os.path.join(ROOT, "stress-particle-sph_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_binding" not in txt and "liboracle" not in txt and "sph_oracle" not in txt, f


def test_fast_division_identity(tmp_path):
    """div_rn() (grid_kernels.cuh) must equal the IEEE quotient bit for bit: 2e7 random + adversarial cases on the
    host FPU (600 M were run once); the GPU pair-weight parity tests cover the device side."""
    import subprocess
    exe = str(tmp_path / "div_identity")
    src = os.path.join(ROOT, "tests", "native", "div_identity.cpp")
    subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-o", exe, src], check=True)
    r = subprocess.run([exe, "20000000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert "two-step=0" in r.stdout


def test_fortran_shim_matches_header():
    """fortran/spsph_shim.f90 cannot be compiled here (no Fortran compiler); check statically that it declares
    every spsph_params / spsph_state member of include/spsph.h in the same order and binds every entry point."""
    hdr = open(os.path.join(ROOT, "include", "spsph.h")).read()
    f90 = open(os.path.join(ROOT, "stress-particle-sph_b200", "fortran", "spsph_shim.f90")).read().lower()

    def c_members(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(const\s+)?(int32_t|float|double)\s*", "", decl)
            for part in decl.split(","):
                names.append(re.sub(r"\[.*?\]|[\*\s]", "", part))
        return [n.lower() for n in names if n]

    def f_members(typename):
        body = re.search(r"type, bind\(c\) :: %s(.*?)end type" % typename, f90, re.S).group(1)
        names = []
        for line in body.splitlines():
            line = line.split("!")[0]
            if "::" not in line:
                continue
            for part in re.split(r",(?![^()]*\))", line.split("::")[1]):
                names.append(re.sub(r"\(.*\)", "", part).strip())
        return [n for n in names if n]

    assert c_members("spsph_params") == f_members("spsph_params")
    assert c_members("spsph_state") == f_members("spsph_state")
    for name in set(re.findall(r"\b(spsph_[a-z_]+)\s*\(", hdr)):
        assert 'name="%s"' % name in f90, name


def _devmath(tmp_path):
    """the per-particle device functions of csrc/dev_common.cuh compiled for the host (tests/native/dev_math_host.cpp)"""
    import ctypes as C
    import subprocess
    so = str(tmp_path / "devmath.so")
    r = subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-fPIC", "-shared", "-w",
                        "-D__noinline__=", "-I/usr/local/cuda/include",
                        "-I" + os.path.join(ROOT, "stress-particle-sph_b200", "csrc"), "-I" + os.path.join(ROOT, "include"),
                        "-o", so, os.path.join(ROOT, "tests", "native", "dev_math_host.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return C.CDLL(so)


def _ptr(a):
    return a.ctypes.data_as(__import__("ctypes").c_void_p)


@pytest.mark.parametrize("ncrit,frict,nflow,delta", [(1, 0., 1, 1.), (2, 0., 1, 1.), (3, 20., 1, 1.), (4, 20., 1, 1.),
                                                     (3, 35., 2, 1.5), (1, 0., 1, 1.5), (12, 0., 1, 1.)])
def test_device_math_transcription_plastic_terms(ncrit, frict, nflow, delta, tmp_path):
    """plastic_terms of csrc/dev_common.cuh (Perzyna with the four yield criteria of invar09, Drucker-Prager of Bui
    et al.) == the oracle's restatement, bit for bit, on 200 000 random stress states around the yield surface
    (host build of the device source: same libm on both sides)"""
    import ctypes as C
    import spsph
    from spsph import decks
    from oracle_binding import Oracle, lib
    if ncrit == 12:
        spec, variant = decks.bui_spec(dx=0.01), "bui"
    else:
        spec, variant = decks.strain_localisation_spec(dx=0.002), "sl"
        spec["props"] = [2, ncrit, 8.e07, 0.25, 1., 2.e3, 1.5e5, -8.e06, frict, 50., delta, nflow]
    decks.write_deck(str(tmp_path), spec)
    prob = spsph.load(str(tmp_path), variant)
    orc = Oracle(prob)
    n = 200000
    assert prob.params.ntotal >= n
    rng = np.random.default_rng(ncrit * 100 + int(frict))
    scale = 1.5e5 if ncrit != 12 else 2.0e4
    stress = rng.normal(0.0, scale, (n, 4))
    stress[: n // 10] *= 1e-3                       # nearly stress-free
    stress[n // 10: n // 5, 2] = 0.0                # principal axes aligned with x, y
    k = slice(n // 5, n // 4)                       # axisymmetric states: Lode angle at +-30 degrees (sint3 = +-1)
    stress[k, 2] = 0.0
    stress[k, 3] = stress[k, 1]
    stress[n // 4: n // 4 + 100] = 0.0              # exactly zero stress
    grad = rng.normal(0.0, 1.0, (n, 4))
    epsp = np.abs(rng.normal(0.0, 5e-3, n))
    epsp[::3] = 0.0
    props = np.array(list(prob.params.props), dtype=np.float64)
    for time_sph in (0.0, 1.0e-3):
        fd_o = rng.normal(0.0, 1.0e4, n)
        fd_d = fd_o.copy()
        Gs_o, Gs_d = np.zeros((n, 4)), np.zeros((n, 4))
        d_o, d_d = np.zeros(n), np.zeros(n)
        rc = lib().oracle_plastic_terms(C.c_void_p(orc.h), C.c_double(time_sph), C.c_int32(n), _ptr(stress), _ptr(grad),
                                        _ptr(epsp), _ptr(fd_o), _ptr(Gs_o), _ptr(d_o))
        assert rc == 0
        _devmath(tmp_path).devmath_plastic_terms(
            C.c_int(ncrit), C.c_int(prob.params.ntype_eco), C.c_int(prob.params.ntype_solid), C.c_double(time_sph),
            _ptr(props), C.c_int(n), _ptr(stress), _ptr(grad), _ptr(epsp), _ptr(fd_d), _ptr(Gs_d), _ptr(d_d))
        assert np.array_equal(Gs_o, Gs_d, equal_nan=True)
        assert np.array_equal(d_o, d_d, equal_nan=True)
        assert np.array_equal(fd_o, fd_d)
        if time_sph > 0:
            assert (Gs_o != 0).any(axis=1).mean() > 0.2  # the plastic branch is exercised


def test_device_math_transcription_adapt_stress_free_kernel(tmp_path):
    """adapt_stress2, apply_stress_free and the three smoothing kernels of csrc/dev_common.cuh == the oracle's, bit for
    bit, on random inputs (host build of the device source)"""
    import ctypes as C
    import spsph
    from spsph import decks
    from oracle_binding import Oracle, lib
    decks.write_deck(str(tmp_path), decks.bui_spec(dx=0.02))
    prob = spsph.load(str(tmp_path), "bui")
    orc = Oracle(prob)
    dm = _devmath(tmp_path)
    n = min(prob.params.nnode, 20000)
    rng = np.random.default_rng(7)
    props = np.array(list(prob.params.props), dtype=np.float64)
    stress = rng.normal(0.0, 2.0e4, (n, 4))
    stress[: n // 4] += 3.0e4  # tension: apex cut-off branch
    a, b = stress.copy(), stress.copy()
    assert lib().oracle_adapt_stress(C.c_void_p(orc.h), C.c_int32(n), _ptr(a)) == 0
    dm.devmath_adapt_stress(_ptr(props), C.c_int(n), _ptr(b))
    assert np.array_equal(a, b) and not np.array_equal(a, stress)
    th = rng.uniform(0, 2 * np.pi, n)
    normal = np.column_stack([np.cos(th), np.sin(th)])
    a, b = stress.copy(), stress.copy()
    assert lib().oracle_stress_free(C.c_void_p(orc.h), C.c_int32(n), _ptr(a), _ptr(normal)) == 0
    dm.devmath_stress_free(C.c_int(n), _ptr(b), _ptr(normal))
    assert np.array_equal(a, b) and not np.array_equal(a, stress)
    h = np.full(n, 1.2 * 0.02) * rng.uniform(0.9, 1.1, n)
    r = rng.uniform(0.0, 3.2, n) * h
    ang = rng.uniform(0, 2 * np.pi, n)
    dx, dy = r * np.cos(ang), r * np.sin(ang)
    lib().oracle_kernel.restype = None
    for skf in (1, 2, 3):
        orc.p.skf = skf
        o2 = Oracle.__new__(Oracle)
        o2.p = orc.p
        st = prob.state()
        o2.h = lib().oracle_create(C.byref(o2.p), C.byref(st))
        wo, gxo, gyo = np.zeros(n), np.zeros(n), np.zeros(n)
        wd, gxd, gyd = np.zeros(n), np.zeros(n), np.zeros(n)
        lib().oracle_kernel(C.c_void_p(o2.h), C.c_int32(n), _ptr(r), _ptr(dx), _ptr(dy), _ptr(h), _ptr(wo), _ptr(gxo), _ptr(gyo))
        dm.devmath_kernel(C.c_int(skf), C.c_double(prob.params.pi), C.c_int(n), _ptr(r), _ptr(dx), _ptr(dy), _ptr(h),
                          _ptr(wd), _ptr(gxd), _ptr(gyd))
        assert np.array_equal(wo, wd) and np.array_equal(gxo, gxd) and np.array_equal(gyo, gyd), skf
        assert (wo != 0).mean() > 0.5
        o2.close()


def test_restart_needs_the_list_capacity(tmp_path):
    """the concept behind spsph.checkpoint (state arrays + list capacity), checked on the oracle: a Bui run restarted at
    step 330 -- after the pair list has grown (step 301) and before it grows again -- continues bit for bit when
    m_pairs is restored, and does not when it is left at 0 (the first restarted step would walk its list reversed)"""
    import spsph
    from spsph import checkpoint, decks
    from oracle_binding import Oracle
    decks.write_deck(str(tmp_path), decks.bui_spec(maxtimestep=1000))
    prob = spsph.load(str(tmp_path), "bui")
    dt = prob.blocks[0]["dt"]
    a = Oracle(prob)
    t_mid = a.run(1, 0.0, dt, 330)
    mid, cap = a.download(), a.list_capacity()
    assert cap > 0
    a.run(331, t_mid, dt, 60)
    ref = a.download()
    for restore in (True, False):
        p2 = prob.copy()
        for k in checkpoint.DYNAMIC:  # the set-up arrays stay those of the deck reader
            p2.arrays[k] = mid[k]
        b = Oracle(p2)
        if restore:
            b.set_list_capacity(cap)
        b.run(331, t_mid, dt, 60)
        got = b.download()
        same = all(np.array_equal(ref[k], got[k]) for k in ("x", "vel", "stress", "internal_vars", "f_drucker"))
        assert same == restore


def test_product_library_has_no_emulation_code():
    """the host emulation (tests/native/) is test infrastructure: the product library is the nvcc build and carries
    none of the SPSPH_HOST_EMU alternatives"""
    import subprocess
    so = os.path.join(ROOT, "stress-particle-sph_b200", "libspsph_cuda.so")
    syms = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    assert "spsph_step" in syms
    assert "emu" not in syms.lower()


def test_sass_identical_without_emulation_guards():
    """the #ifdef SPSPH_HOST_EMU / SPSPH_EMU_* alternatives inside csrc/ are invisible to nvcc: a build from copies of
    the sources with every such conditional resolved and deleted has the same SASS, function by function, as the
    product library (tools/sass_guard_check.py; the committed record is profiles/r2_sass_guard_check.txt)"""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_guard_check
    text, n = sass_guard_check.strip_guards(
        "a\n#ifdef SPSPH_HOST_EMU\nb\n#else\nc\n#endif\n#if defined(SPSPH_HOST_EMU) && !defined(SPSPH_EMU_SIMT)\nd\n"
        "#endif\n#if !defined(SPSPH_HOST_EMU) || defined(SPSPH_EMU_SIMT)\ne\n#ifdef OTHER\nf\n#endif\n#endif\n")
    assert (text.split(), n) == (["a", "c", "e", "#ifdef", "OTHER", "f", "#endif"], 3)
    ok, report = sass_guard_check.check()
    assert ok, report
