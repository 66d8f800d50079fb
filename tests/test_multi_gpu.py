"""Multi-GPU parity: the x-slab decomposition (one process per GPU, NCCL halo exchange) must reproduce the
single-domain oracle bit for bit on the owned particles. Needs >= 2 GPUs (gpurun --gpus 2); skipped otherwise."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


# bui, 300 steps: the pair count passes its previous maximum 18 times from step 208 on, which exercises the
# distributed search of the reference's list-growth traversal rule across the slabs
@pytest.mark.parametrize("kind,steps", [("vs", 30), ("sl", 30), ("bui", 300)])
def test_two_slabs_match_oracle(tmp_path, deck_dir, kind, steps):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import spsph
    from spsph import dist
    from oracle_binding import Oracle
    out = str(tmp_path / "dist")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "dist_worker.py"), "--kind", kind,
           "--steps", str(steps), "--out", out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    prob = spsph.load(deck_dir(kind), kind)
    orc = Oracle(prob)
    orc.run(1, 0.0, prob.blocks[0]["dt"], steps)
    ref = orc.download()
    ranks = [np.load(os.path.join(out, f"rank{k}.npz")) for k in range(2)]
    keys = ("x", "vel", "stress", "internal_vars", "f_drucker", "displ")
    merged = dist.merge_owned([{k: r_[k] for k in keys} for r_ in ranks], [r_["flags"] for r_ in ranks], prob.params)
    nt = prob.params.ntotal
    assert int(ranks[0]["npairs"]) == orc.pair_stats()["npairs"]
    for k in keys:
        a, b = merged[k], ref[k]
        if k in ("x", "vel", "stress"):
            a, b = a[:nt], b[:nt]
        assert np.array_equal(a, b), f"{kind}: {k} differs, max |diff| {np.abs(a - b).max():.3e}"
    # both ranks really own a share
    assert all((r_["flags"] == 1).sum() > 0.2 * prob.params.ntotal2 for r_ in ranks)


def test_bench_two_slabs_records_parity_and_weak_scaling(tmp_path):
    """bench.py --gpus 2 (reduced sizes): the line carries a bit-for-bit slab parity check against the single-GPU engine
    and against the reference executable's golden, the weak-scaling sub-record, and an e2e leg that moves rank-local rows"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import json
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "5",
           "--warmup", "3", "--ncol", "408", "--no-cpu", "--weak-ncol", "200", "--weak-steps", "4", "--parity-ncol", "136",
           "--parity-steps", "30"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    pc = line["parity_check"]
    assert pc["ranks"] == 2 and pc["bitwise_equal_single_gpu"] is True and pc["pair_count_equal"] is True, pc
    assert pc["reference_executable_golden"]["bitwise_equal"] is True, pc
    assert line["weak_scaling"]["value"] > 0 and line["e2e"]["value"] > 0
    assert "spsph_upload_rows" in line["config"]["e2e_protocol"]


def test_two_slabs_gauss_kernel_halo(tmp_path):
    """Gauss kernel (cut-off 3 h on cells of 2 h): the halo distance follows the cut-off"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import spsph
    from spsph import decks, dist
    from oracle_binding import Oracle
    out = str(tmp_path / "dist")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29535", os.path.join(ROOT, "tools", "dist_worker.py"), "--kind", "vs_gauss",
           "--steps", "30", "--out", out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    d = str(tmp_path / "deck")
    spec = decks.vertical_slope_spec()
    spec["skf"] = 2
    decks.write_deck(d, spec)
    prob = spsph.load(d, "vs")
    orc = Oracle(prob)
    orc.run(1, 0.0, prob.blocks[0]["dt"], 30)
    ref = orc.download()
    ranks = [np.load(os.path.join(out, f"rank{k}.npz")) for k in range(2)]
    keys = ("x", "vel", "stress", "internal_vars", "f_drucker", "displ")
    merged = dist.merge_owned([{k: r_[k] for k in keys} for r_ in ranks], [r_["flags"] for r_ in ranks], prob.params)
    nt = prob.params.ntotal
    for k in keys:
        a, b = merged[k], ref[k]
        if k in ("x", "vel", "stress"):
            a, b = a[:nt], b[:nt]
        assert np.allclose(a, b, rtol=1e-9, atol=1e-300), f"vs_gauss: {k} differs"


@pytest.mark.parametrize("case", ["bui_cont_density", "vs_cont_density_sle2", "sl_sigman_xsph", "vs_sigman"])
def test_two_slabs_extended_halo_record(tmp_path, case):
    """continuity density (+ smoothing-length update) and per-step free-surface marks (apply_stress_free, XSPH next to
    boundary conditions) on two slabs: the density, smoothing length, velocity divergence, marks and normals travel in
    the extended halo record; owned particles equal the single-domain oracle bit for bit"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import spsph
    from spsph import decks, dist
    from oracle_binding import Oracle
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from ref_cases import spec_of
    steps = 30
    out = str(tmp_path / "dist")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29536", os.path.join(ROOT, "tools", "dist_worker.py"), "--kind", "case:" + case,
           "--steps", str(steps), "--out", out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    variant, spec = spec_of(case)
    d = str(tmp_path / "deck")
    decks.write_deck(d, spec)
    prob = spsph.load(d, variant)
    orc = Oracle(prob)
    orc.run(1, 0.0, prob.blocks[0]["dt"], steps)
    ref = orc.download()
    ranks = [np.load(os.path.join(out, f"rank{k}.npz")) for k in range(2)]
    keys = ("x", "vel", "stress", "internal_vars", "f_drucker", "displ", "rho", "hsml")
    merged = dist.merge_owned([{k: r_[k] for k in keys} for r_ in ranks], [r_["flags"] for r_ in ranks], prob.params)
    nt = prob.params.ntotal
    for k in keys:
        a, b = merged[k][:nt], ref[k][:nt]
        assert np.array_equal(a, b), f"{case}: {k} differs in {int((a != b).sum())} entries, max |diff| {np.abs(a - b).max():.3e}"


def test_two_slabs_dynamic_reslabbing(tmp_path, deck_dir):
    """spsph_dist_set_planes every 40 steps (the plane swings by 0.3 halo distances each time, so thousands of particles
    change owner): the particles that change slab travel with the next halo exchange; owned particles stay bit-identical
    to the single-domain oracle, and every particle keeps exactly one owner"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import spsph
    from spsph import dist
    from oracle_binding import Oracle
    out = str(tmp_path / "dist")
    steps = 130
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29537", os.path.join(ROOT, "tools", "dist_worker.py"), "--kind", "bui",
           "--steps", str(steps), "--out", out, "--replan", "40"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    prob = spsph.load(deck_dir("bui"), "bui")
    orc = Oracle(prob)
    orc.run(1, 0.0, prob.blocks[0]["dt"], steps)
    ref = orc.download()
    ranks = [np.load(os.path.join(out, f"rank{k}.npz")) for k in range(2)]
    keys = ("x", "vel", "stress", "internal_vars", "f_drucker", "displ")
    merged = dist.merge_owned([{k: r_[k] for k in keys} for r_ in ranks], [r_["flags"] for r_ in ranks], prob.params)
    nt = prob.params.ntotal
    for k in keys:
        a, b = merged[k][:nt], ref[k][:nt]
        assert np.array_equal(a, b), f"re-slabbing: {k} differs in {int((a != b).sum())} entries"


def test_two_slabs_tile_path(tmp_path, deck_dir, monkeypatch):
    """the cell-tile path (SPSPH_TILE=1) on two slabs: 120 steps of the Bui column, owned particles bit-identical to the
    single-domain oracle"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import spsph
    from spsph import dist
    from oracle_binding import Oracle
    monkeypatch.setenv("SPSPH_TILE", "1")
    out = str(tmp_path / "dist")
    steps = 120
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29538", os.path.join(ROOT, "tools", "dist_worker.py"), "--kind", "bui",
           "--steps", str(steps), "--out", out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    prob = spsph.load(deck_dir("bui"), "bui")
    orc = Oracle(prob)
    orc.run(1, 0.0, prob.blocks[0]["dt"], steps)
    ref = orc.download()
    ranks = [np.load(os.path.join(out, f"rank{k}.npz")) for k in range(2)]
    assert int(ranks[0]["tile_steps"]) >= steps - 2, "the slab run did not use the tile kernels"
    keys = ("x", "vel", "stress", "internal_vars", "f_drucker", "displ")
    merged = dist.merge_owned([{k: r_[k] for k in keys} for r_ in ranks], [r_["flags"] for r_ in ranks], prob.params)
    nt = prob.params.ntotal
    for k in keys:
        a, b = merged[k][:nt], ref[k][:nt]
        assert np.array_equal(a, b), f"tile path on two slabs: {k} differs in {int((a != b).sum())} entries"
