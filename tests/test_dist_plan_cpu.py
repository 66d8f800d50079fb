"""CPU test of the multi-GPU host logic with a real 2-process group (gloo): slab planning, ownership rule and
halo coverage, checked against the oracle's pair list."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, deck, kind, q):
    sys.path.insert(0, os.path.join(ROOT, "stress-particle-sph_b200"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as td
    import spsph
    from spsph import dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    prob = spsph.load(deck, kind)
    plan = dist.plan_slabs(prob, world)
    flags = dist.initial_flags(prob, plan, rank)
    owned = torch.tensor((flags == dist.OWNED).astype(np.int32))
    total = owned.clone()
    td.all_reduce(total)                      # every particle is owned exactly once
    ok_partition = bool((total == 1).all())
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    td.all_gather(counts, torch.tensor([int(owned.sum())]))
    # halo coverage: every partner of an owned particle is local (first-step pair list from the oracle)
    from oracle_binding import Oracle
    orc = Oracle(prob)
    orc.step(1, 0.0, prob.blocks[0]["dt"])
    pr = orc.pairs()
    i, j = pr["pair_i"] - 1, pr["pair_j"] - 1
    loc = flags != dist.REMOTE
    own = flags == dist.OWNED
    ok_halo = bool(loc[j[own[i]]].all() and loc[i[own[j]]].all())
    q.put((rank, ok_partition, ok_halo, [int(c) for c in counts], plan["halo_cells"], plan["halo_capacity"]))
    td.barrier()
    td.destroy_process_group()


@pytest.mark.parametrize("kind", ["vs", "bui"])
def test_slab_plan_partition_and_halo(deck_dir, kind):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29611 + (0 if kind == "vs" else 1)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, deck_dir(kind), kind, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import spsph
    n2 = spsph.load(deck_dir(kind), kind).params.ntotal2
    for rank, ok_partition, ok_halo, counts, cells, cap in res:
        assert ok_partition and ok_halo
        assert sum(counts) == n2 and min(counts) > 0.3 * n2
        assert cells >= 11 and cap >= 1024


def test_dependent_sweep_count(deck_dir):
    import spsph
    from spsph import dist
    assert dist.dependent_sweeps(spsph.load(deck_dir("bui"), "bui").params) == 11   # shift + 8 + final + XSPH
    assert dist.dependent_sweeps(spsph.load(deck_dir("vs"), "vs").params) == 9


def test_cost_weighted_planes(tmp_path):
    """spsph.dist.rebalance_cost: a rank that pays more per particle (the side wall) hands particles to its neighbours;
    the modelled maximum cost drops, no plane moves further than spsph_dist_set_planes accepts, the order is kept"""
    sys.path.insert(0, os.path.join(ROOT, "stress-particle-sph_b200"))
    import spsph
    from spsph import decks, dist
    d = str(tmp_path / "deck")
    decks.write_deck(d, decks.refined_bui_spec(ncol=408, maxtimestep=5))
    prob = spsph.load(d, "bui")
    plan = dist.plan_slabs(prob, 8)
    planes, H = np.array(plan["planes"]), plan["H"]
    xn = np.sort(prob.arrays["x"][:prob.params.nnode, 0])

    def processed(pl):
        lo = np.where(np.isfinite(pl[:-1]), pl[:-1] - H, -np.inf)
        hi = np.where(np.isfinite(pl[1:]), pl[1:] + H, np.inf)
        return (np.searchsorted(xn, hi) - np.searchsorted(xn, lo)).astype(float)

    n0 = processed(planes)
    per_particle = np.array([1.08, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.96])  # ms per processed particle, relative
    cost0 = per_particle * n0
    new = dist.rebalance_cost(prob, planes, cost0, H)
    cost1 = per_particle * processed(new)
    assert cost1.max() < cost0.max() * 0.985, (cost0, cost1)
    assert cost1[:-1].max() / cost1[:-1].min() < 1.07  # (the last slab takes what is left and may hit the 0.4 H bound)
    assert np.all(np.abs(new[1:-1] - planes[1:-1]) <= 0.4 * H + 1e-12)
    assert np.all(np.diff(new[1:-1]) > H)
    # every rank already costs the same: nothing moves
    same = dist.rebalance_cost(prob, planes, np.full(8, 1.0), H)
    assert np.allclose(same[1:-1], planes[1:-1], atol=0.02 * H)
