#!/usr/bin/env python
"""One rank of a multi-GPU run (launched by torchrun or by tests/test_multi_gpu.py with RANK/WORLD_SIZE set).
Runs `--steps` steps of a generated deck on the x-slab decomposition and writes this rank's download + ownership
flags to `--out`/rank<r>.npz for the parity check against the oracle."""
import argparse
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stress-particle-sph_b200"))
import spsph  # noqa: E402
from spsph import decks, dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="vs")
    ap.add_argument("--ncol", type=int, default=0)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--out", required=True)
    ap.add_argument("--replan", type=int, default=0,
                    help="every N steps: re-slab (spsph_dist_set_planes); the interior planes are pushed back and forth by "
                         "0.3 halo distances and then rebalanced on the owned counts (spsph.dist.rebalance)")
    ap.add_argument("--rows", action="store_true",
                    help="rank-local transfers: after dist_init the rank uploads only its slab + halo rows again "
                         "(spsph_upload_rows) and reads its results back with spsph_download_rows of the rows it owns")
    a = ap.parse_args()
    import torch
    import torch.distributed as td
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    emu = os.environ.get("SPSPH_EMU_SO")
    if emu:  # tests/test_dist_emulated_cpu.py: the host-emulated engine (no GPU), NCCL replaced through SPSPH_NCCL_SO
        import spsph.engine as E
        E._lib, E._CUDA_SO = None, emu
        local = 0
    else:
        torch.cuda.set_device(local)
    td.init_process_group("gloo")
    d = tempfile.mkdtemp()
    if a.kind == "refined_bui":
        spec, var = decks.refined_bui_spec(ncol=a.ncol), "bui"
    elif a.kind == "wide_slope":
        spec, var = decks.wide_slope_spec(ncol=a.ncol, nslab=world), "vs"
    elif a.kind.startswith("case:"):  # a golden case of oracle/ref_cases.py (option combinations no shipped deck uses)
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        from ref_cases import spec_of
        var, spec = spec_of(a.kind[5:])
    elif a.kind == "vs_gauss":
        spec, var = decks.vertical_slope_spec(), "vs"
        spec["skf"] = 2
    else:
        spec, var = decks.SHIPPED[a.kind](), a.kind
    decks.write_deck(d, spec)
    prob = spsph.load(d, var)
    plan = dist.plan_slabs(prob, world)
    uid = [spsph.dist_unique_id() if rank == 0 else None]
    td.broadcast_object_list(uid, src=0)
    eng = spsph.Engine(prob, device=local)
    eng.dist_init(rank, world, uid[0], plan)
    dt = prob.blocks[0]["dt"]
    if a.rows:
        from spsph.engine import row_arrays
        ids_local = dist.local_ids(prob, plan, rank)
        eng.upload_rows(row_arrays(prob.params, prob.arrays, ids_local), ids_local)
    if a.replan > 0:
        planes, t, done, k = np.array(plan["planes"]), 0.0, 0, 0
        while done < a.steps:
            n = min(a.replan, a.steps - done)
            t = eng.run(1 + done, t, dt, n)
            done += n
            if done < a.steps:
                cnt = [None] * world
                td.all_gather_object(cnt, int((eng.dist_flags()[:prob.params.nnode] == 1).sum()))
                new = dist.rebalance(planes, cnt, plan["H"])
                new[1:-1] += (0.3 if k % 2 == 0 else -0.3) * plan["H"] - (new[1:-1] - planes[1:-1])  # forced swing
                planes = new
                eng.set_planes(planes)
                k += 1
    else:
        eng.run(1, 0.0, dt, a.steps)
    ms, launches = eng.last_run()
    arrs = eng.download()
    flags = eng.dist_flags()
    if a.rows:  # the owned rows, fetched row-wise, replace the complete download in what the parity check reads
        from spsph.engine import _row_selector
        keys = ("x", "vel", "stress", "internal_vars", "f_drucker", "displ")
        owned = np.flatnonzero(flags == 1).astype(np.int32)
        rows = eng.download_rows(owned, keys=keys)
        for k in keys:
            full = np.full_like(arrs[k], np.nan)  # anything that is not an owned row must not be looked at
            full[_row_selector(prob.params, owned, k)] = rows[k]
            arrs[k] = full
    os.makedirs(a.out, exist_ok=True)
    np.savez(os.path.join(a.out, f"rank{rank}.npz"), flags=flags, ms=ms, npairs=eng.pair_stats()["npairs"],
             tile_steps=eng.path_counts()[0],
             **{k: arrs[k] for k in ("x", "vel", "stress", "internal_vars", "f_drucker", "displ", "rho", "hsml")})
    td.barrier()
    eng.close()
    td.destroy_process_group()


if __name__ == "__main__":
    main()
