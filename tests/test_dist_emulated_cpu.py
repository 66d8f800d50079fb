"""The multi-GPU layer on the CPU: world_size-2 runs (one process per rank, launched with torch.distributed.run, gloo
for the bootstrap) of the HOST-EMULATED engine -- csrc/spsph_engine.cu and every kernel compiled for the host
(tests/native/cuda_host_emu.h) -- with NCCL replaced by a file-based stand-in (tests/native/fake_nccl.cpp, selected
through SPSPH_NCCL_SO). What runs is the product's own slab code: spsph_dist_init, halo select / pack / unpack,
migration, ghost-list compaction, the all-reduced grid bounds and pair counts, the distributed list-growth search,
slab re-planning. Owned particles of the two ranks, merged, must equal the single-domain oracle bit for bit (the same
assertions tests/test_multi_gpu.py makes on two real GPUs)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NATIVE = os.path.join(ROOT, "tests", "native")
KEYS = ("x", "vel", "stress", "internal_vars", "f_drucker", "displ")


@pytest.fixture(scope="module")
def emu_dist(tmp_path_factory):
    """(emulated engine .so, fake NCCL .so)"""
    d = tmp_path_factory.mktemp("emu_dist")
    cpp, so, nccl = str(d / "engine_host.cpp"), str(d / "libspsph_emu.so"), str(d / "libfake_nccl.so")
    subprocess.run([sys.executable, os.path.join(NATIVE, "make_engine_host.py"),
                    os.path.join(ROOT, "stress-particle-sph_b200", "csrc", "spsph_engine.cu"), cpp], check=True,
                   stdout=subprocess.DEVNULL)
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-fPIC", "-shared", "-w",
                    "-D__noinline__=", "-fno-gnu-unique", "-I/usr/local/cuda/include", "-I" + NATIVE,
                    "-I" + os.path.join(ROOT, "stress-particle-sph_b200", "csrc"), "-I" + os.path.join(ROOT, "include"),
                    "-o", so, cpp, "-ldl"], check=True)
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-I/usr/local/cuda/include", "-o", nccl,
                    os.path.join(NATIVE, "fake_nccl.cpp")], check=True)
    return so, nccl


@pytest.fixture(scope="module")
def emu_dist_simt(emu_dist, tmp_path_factory):
    """the same with the lockstep (SIMT) emulation: warp collectives and block barriers have their real meaning, which
    the cell-tile kernels need (tests/test_step_emulation_cpu.py)"""
    d = tmp_path_factory.mktemp("emu_dist_simt")
    cpp, so = str(d / "engine_host.cpp"), str(d / "libspsph_emu_simt.so")
    subprocess.run([sys.executable, os.path.join(NATIVE, "make_engine_host.py"),
                    os.path.join(ROOT, "stress-particle-sph_b200", "csrc", "spsph_engine.cu"), cpp], check=True,
                   stdout=subprocess.DEVNULL)
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-fPIC", "-shared", "-w",
                    "-DSPSPH_EMU_SIMT", "-D__noinline__=", "-fno-gnu-unique", "-I/usr/local/cuda/include", "-I" + NATIVE,
                    "-I" + os.path.join(ROOT, "stress-particle-sph_b200", "csrc"), "-I" + os.path.join(ROOT, "include"),
                    "-o", so, cpp, "-ldl"], check=True)
    return so, emu_dist[1]


def _run_ranks(emu_dist, tmp_path, kind, steps, port, extra=(), world=2, peel=False, env_extra=None):
    so, nccl = emu_dist
    out = str(tmp_path / "dist")
    env = dict(os.environ, SPSPH_EMU_SO=so, SPSPH_NCCL_SO=nccl, SPSPH_FAKE_NCCL_DIR=str(tmp_path), OMP_NUM_THREADS="1",
               SPSPH_PEEL="1" if peel else "0", **(env_extra or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "dist_worker.py"), "--kind", kind,
           "--steps", str(steps), "--out", out] + list(extra)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return [np.load(os.path.join(out, f"rank{k}.npz")) for k in range(world)]


def _assert_owned_equal_oracle(prob, ranks, steps, label, exact=True, min_share=0.2):
    from spsph import dist
    from oracle_binding import Oracle
    orc = Oracle(prob)
    orc.run(1, 0.0, prob.blocks[0]["dt"], steps)
    ref = orc.download()
    merged = dist.merge_owned([{k: r[k] for k in KEYS} for r in ranks], [r["flags"] for r in ranks], prob.params)
    nt = prob.params.ntotal
    assert int(ranks[0]["npairs"]) == orc.pair_stats()["npairs"], label
    for k in KEYS:
        a, b = merged[k], ref[k]
        if k in ("x", "vel", "stress"):
            a, b = a[:nt], b[:nt]
        assert np.array_equal(a, b), f"{label}: {k} differs, max |diff| {np.abs(a - b).max():.3e}"
    assert all((r["flags"] == 1).sum() > min_share * prob.params.ntotal2 for r in ranks), "a rank owns (almost) nothing"
    orc.close()


# bui, 300 steps: the pair count passes its previous maximum repeatedly from step 208 on -> the distributed search of
# the reference's list-growth traversal rule across the slabs; particles migrate between the slabs
@pytest.mark.parametrize("kind,steps,port", [("vs", 30, 29611), ("sl", 12, 29612), ("bui", 300, 29613)])
def test_two_emulated_slabs_match_oracle(emu_dist, tmp_path, deck_dir, kind, steps, port):
    import spsph
    # halo peeling (SPSPH_PEEL=1, k_peel_counts: per-sweep list lengths with the too-deep ghosts zeroed) rides along in
    # the long Bui run: it must not change a bit of any owned particle
    ranks = _run_ranks(emu_dist, tmp_path, kind, steps, port, peel=(kind == "bui"))
    _assert_owned_equal_oracle(spsph.load(deck_dir(kind), kind), ranks, steps, f"{kind}, 2 emulated slabs")


@pytest.mark.parametrize("case,port", [("bui_cont_density", 29621), ("sl_sigman_xsph", 29622)])
def test_two_emulated_slabs_extended_halo_record(emu_dist, tmp_path, case, port):
    """continuity density and per-step free-surface marks (apply_stress_free, XSPH next to boundary conditions): the
    density, smoothing length, velocity divergence, marks and normals travel in the extended halo record"""
    import spsph
    from spsph import decks
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from ref_cases import spec_of
    steps = 20
    ranks = _run_ranks(emu_dist, tmp_path, "case:" + case, steps, port)
    variant, spec = spec_of(case)
    d = str(tmp_path / "deck")
    os.makedirs(d)
    decks.write_deck(d, spec)
    _assert_owned_equal_oracle(spsph.load(d, variant), ranks, steps, f"{case}, 2 emulated slabs")


def test_two_emulated_slabs_replanned(emu_dist, tmp_path, deck_dir):
    """dynamic re-slabbing (spsph_dist_set_planes): the plane swings by 0.3 halo distances every 40 steps, the
    particles that change owner travel with the next halo exchange"""
    import spsph
    steps = 130
    ranks = _run_ranks(emu_dist, tmp_path, "bui", steps, 29631, extra=["--replan", "40"])
    _assert_owned_equal_oracle(spsph.load(deck_dir("bui"), "bui"), ranks, steps, "bui, re-planned slabs", min_share=0.05)


def test_four_emulated_slabs_interior_ranks(emu_dist, tmp_path):
    """four slabs of a refined Bui column (the BASELINE configs[3] workload at 1/800 of its size): the two interior
    ranks exchange halos with both neighbours, halo peeling on (a third of the ghost list entries skipped in the late
    sweeps of this geometry); distributed pair count and owned particles equal the oracle's"""
    import spsph
    from spsph import decks
    steps, ncol = 8, 204
    ranks = _run_ranks(emu_dist, tmp_path, "refined_bui", steps, 29641, extra=["--ncol", str(ncol)], world=4, peel=True)
    d = str(tmp_path / "deck")
    os.makedirs(d)
    decks.write_deck(d, decks.refined_bui_spec(ncol=ncol))
    _assert_owned_equal_oracle(spsph.load(d, "bui"), ranks, steps, "refined bui, 4 emulated slabs", min_share=0.1)


def test_two_emulated_slabs_gauss_kernel_halo(emu_dist, tmp_path):
    """Gauss kernel (cut-off 3 h on cells of 2 h): the halo distance follows the cut-off. On the
    emulated ranks exp() is the oracle's libm, so the comparison is bitwise here (1e-9 on the GPU)"""
    import spsph
    from spsph import decks
    steps = 20
    ranks = _run_ranks(emu_dist, tmp_path, "vs_gauss", steps, 29651)
    d = str(tmp_path / "deck")
    os.makedirs(d)
    spec = decks.vertical_slope_spec()
    spec["skf"] = 2
    decks.write_deck(d, spec)
    _assert_owned_equal_oracle(spsph.load(d, "vs"), ranks, steps, "vs, Gauss kernel, 2 emulated slabs")


def test_two_emulated_slabs_row_transfers(emu_dist, tmp_path, deck_dir):
    """rank-local transfers in a slab run (what bench.py's end-to-end leg does at N > 1): after spsph_dist_init every
    rank uploads only its slab + halo rows (spsph_upload_rows) and fetches only the rows it owns at the end
    (spsph_download_rows, ownership after migration from spsph_dist_flags); everything else is NaN in what this test
    merges, so a row that was not transferred cannot pass"""
    import spsph
    steps = 60
    ranks = _run_ranks(emu_dist, tmp_path, "bui", steps, 29661, extra=["--rows"])
    _assert_owned_equal_oracle(spsph.load(deck_dir("bui"), "bui"), ranks, steps, "bui, row-wise transfers, 2 emulated slabs")


def test_replanned_slabs_message_limit_regression(emu_dist, tmp_path, monkeypatch):
    """After spsph_dist_set_planes the rank that gains particles first sends a halo band that is short by the moved
    strip, and a full band one step later: up to 50 % more records than the step before, beyond the 12.5 % (+ slack)
    a message may grow by. bench.py --rebalance died of that on the 4 M column ("halo message capacity exceeded");
    the engine now sends full-capacity messages for two exchanges after a change of planes. Small problems hide the
    defect behind the fixed slack of 4096 records, so the test takes the slack away (SPSPH_HALO_SLACK=0)."""
    import spsph
    from spsph import decks
    monkeypatch.setenv("SPSPH_HALO_SLACK", "0")
    steps, ncol = 8, 204
    ranks = _run_ranks(emu_dist, tmp_path, "refined_bui", steps, 29671, extra=["--ncol", str(ncol), "--replan", "4"])
    d = str(tmp_path / "deck")
    os.makedirs(d)
    decks.write_deck(d, decks.refined_bui_spec(ncol=ncol))
    _assert_owned_equal_oracle(spsph.load(d, "bui"), ranks, steps, "refined bui, planes moved every 4 steps", min_share=0.1)


def test_four_emulated_slabs_wide_slope(emu_dist, tmp_path):
    """the weak-scaling workload (BASELINE configs[4]: wide vertical slope, one 10 m x 10 m block per rank, elastic, CSPM,
    boundary conditions with a gravity ramp, inside approach) at 1/550 of its size on four slabs"""
    import spsph
    from spsph import decks
    steps, ncol, world = 15, 60, 4
    ranks = _run_ranks(emu_dist, tmp_path, "wide_slope", steps, 29681, extra=["--ncol", str(ncol)], world=world)
    d = str(tmp_path / "deck")
    os.makedirs(d)
    decks.write_deck(d, decks.wide_slope_spec(ncol=ncol, nslab=world))
    _assert_owned_equal_oracle(spsph.load(d, "vs"), ranks, steps, "wide slope, 4 emulated slabs", min_share=0.1)


def test_two_emulated_slabs_tile_path(emu_dist_simt, tmp_path, deck_dir):
    """the cell-tile path (SPSPH_TILE=1: one-pass build with entry codes, shared-memory partner tiles) on two slabs of
    the SIMT emulation: the ranks agree on the path through the all-reduced pair count and overflow flags, every step
    runs on the tile kernels, owned particles equal the oracle bit for bit"""
    import spsph
    steps = 6
    ranks = _run_ranks(emu_dist_simt, tmp_path, "bui", steps, 29691, env_extra={"SPSPH_TILE": "1"})
    assert all(int(r["tile_steps"]) == steps for r in ranks), [int(r["tile_steps"]) for r in ranks]
    _assert_owned_equal_oracle(spsph.load(deck_dir("bui"), "bui"), ranks, steps, "bui, tile path, 2 SIMT-emulated slabs")
