"""Host-side planning of the multi-GPU x-slab decomposition (device side in
csrc/dist_kernels.cuh). Pure numpy so that the partition logic is testable without a GPU.

One process per GPU. Every rank loads the complete problem, creates its Engine, uploads everything and calls
Engine.dist_init(plan): from then on each rank advances its slab plus a wide halo and exchanges ghost
particles / migrants with its two neighbours over NCCL once per time step.
"""
import numpy as np

REMOTE, OWNED, GHOST = 0, 1, 2


def dependent_sweeps(params):
    """number of kernels per time step that read partner values produced earlier in the same step; the
    redundantly computed ghost region loses one cell of validity per such kernel"""
    n = 8 + 1                      # 4 x (stress_point_update, get_derivatives) + the final stress_point_update
    if params.sph_shift:
        n += 1                     # interpolation at the start of the step (main:99-109)
    if params.update_x and params.xsph:
        n += 1                     # XSPH_update reads partner velocities (main:189-239)
    if params.update_x and params.no_bcs > 0 and (params.ifsigman == 1 or params.xsph):
        # get_nodes_on_free_surface runs every step (spsph_create: fs_each_step): it reads the partners' new positions,
        # then the step-4 normals read the partners' fresh marks and covered flags (mat:1116-1411) -- plus one spare
        n += 3
    return n


def plan_slabs(problem, nranks, safety=2.0, balance="processed"):
    """-> dict(planes, halo_cells, halo_capacity, H). A plane is placed midway between two distinct particle
    columns so that no particle sits on it initially. balance = "processed": slabs are sized so that every rank
    PROCESSES the same number of velocity particles (owned + both halos; interior slabs come out narrower than
    the two end slabs); balance = "owned": equal numbers of owned particles."""
    p = problem.params
    x = problem.arrays["x"][:, 0]
    xn = np.sort(x[:p.nnode])
    ux = np.unique(xn)
    halo_cells = dependent_sweeps(p) + 2
    hmax = float(problem.arrays["hsml"].max())
    # a dependent sweep reads partners up to the kernel cut-off away: 2 h for the cubic spline, 3 h for the Gauss and
    # quintic kernels (main:1228-1238; the cell edge stays 2 h) -- the same expression as spsph_dist_init
    scale_k = 2.0 if p.skf == 1 else 3.0
    H = halo_cells * scale_k * hmax

    def snap(v):  # midpoint between the two distinct columns around v
        k = int(np.searchsorted(ux, v))
        k = min(max(k, 1), len(ux) - 1)
        return 0.5 * (ux[k - 1] + ux[k])

    def count(lo, hi):
        return int(np.searchsorted(xn, hi) - np.searchsorted(xn, lo))

    def place(target):  # greedy left-to-right placement for a per-rank processed count `target`
        pl = [-np.inf]
        for r in range(nranks - 1):
            lo = pl[-1] - H if r > 0 else -np.inf
            base = int(np.searchsorted(xn, lo)) if r > 0 else 0
            k = min(base + target, len(xn) - 1)
            pl.append(snap(xn[k] - H))   # processed range of rank r ends at plane + H
            if pl[-1] <= pl[-2] + (0 if r == 0 else H):
                pl[-1] = snap((pl[-2] if r > 0 else xn[0]) + 1.0001 * H)
        pl.append(np.inf)
        return pl

    if nranks > 1 and H > 0.25 * (xn[-1] - xn[0]) / nranks:
        balance = "owned"  # halo comparable to the slab width (tiny problems): equalising the processed count is moot
    if nranks == 1:
        planes = [-np.inf, np.inf]
    elif balance == "owned":
        planes = [-np.inf] + [snap(xn[(r * p.nnode) // nranks]) for r in range(1, nranks)] + [np.inf]
    else:
        lo_t, hi_t = p.nnode // nranks, p.nnode
        for _ in range(40):  # bisection on the common processed count: the last rank takes what is left
            mid = (lo_t + hi_t) // 2
            pl = place(mid)
            last = count(pl[-2] - H, np.inf)
            if last > mid:
                lo_t = mid + 1
            else:
                hi_t = mid
        planes = place(hi_t)
    planes = np.array(planes, dtype=np.float64)
    if nranks > 1 and not np.all(np.diff(planes[1:-1]) > 0) and nranks > 2:
        raise ValueError("slab planning failed: planes are not increasing")
    cap = 1024
    for r in range(1, nranks):
        near = int(((x >= planes[r] - H) & (x < planes[r] + H)).sum())
        cap = max(cap, int(safety * near) + 1024)
    widths = np.diff(planes[1:-1]) if nranks > 2 else np.array([np.inf])
    if nranks > 2 and widths.min() < H:
        raise ValueError(f"slab width {widths.min():.4g} is below the halo distance {H:.4g}: use fewer ranks")
    return dict(planes=planes, halo_cells=halo_cells, halo_capacity=cap, H=H)


def key_x(problem):
    """position that decides ownership (k_dist_init_flags / key_x in dist_kernels.cuh)"""
    p = problem.params
    x = problem.arrays["x"][:, 0].copy()
    if p.sp_sph and not p.inside_approach:
        node_of = (np.arange(p.nstress) // p.npoints)
        x[p.nnode:p.ntotal] = problem.arrays["x"][node_of, 0]
    return x


def initial_flags(problem, plan, rank):
    """numpy restatement of k_dist_init_flags: 0 remote, 1 owned, 2 ghost"""
    lo, hi, H = plan["planes"][rank], plan["planes"][rank + 1], plan["H"]
    xk, xi = key_x(problem), problem.arrays["x"][:, 0]
    f = np.zeros(len(xi), np.int32)
    f[(xi >= lo - H) & (xi < hi + H)] = GHOST
    f[(xk >= lo) & (xk < hi)] = OWNED
    return f


def local_ids(problem, plan, rank, x=None):
    """particle numbers a rank has to hold (owned + ghost), ascending: the row set of spsph_upload_rows"""
    if x is None:
        return np.flatnonzero(initial_flags(problem, plan, rank) != REMOTE).astype(np.int32)
    p = problem.params
    lo, hi, H = plan["planes"][rank], plan["planes"][rank + 1], plan["H"]
    xi = x[:, 0]
    xk = xi.copy()
    if p.sp_sph and not p.inside_approach:
        xk[p.nnode:p.ntotal] = xi[np.arange(p.nstress) // p.npoints]
    keep = ((xi >= lo - H) & (xi < hi + H)) | ((xk >= lo) & (xk < hi))
    return np.flatnonzero(keep).astype(np.int32)


def rebalance(planes, owned_counts, H, gain=0.5):
    """Dynamic re-slabbing: shifts the interior planes towards equal owned velocity-particle counts. planes: current
    planes (nranks + 1), owned_counts: velocity particles every rank owns now (the same list on every rank, e.g. from
    an all-gather of Engine.dist_flags() counts). A plane moves by at most 0.4 H per call (spsph_dist_set_planes
    accepts 0.5 H); the estimate assumes a uniform density inside the two slabs a plane separates."""
    planes = np.array(planes, dtype=np.float64)
    n = np.asarray(owned_counts, dtype=np.float64)
    nr = len(n)
    new = planes.copy()
    target = n.sum() / nr
    excess = 0.0  # particles that should move rightwards across plane r
    for r in range(1, nr):
        excess += n[r - 1] - target
        # width per particle of the slab the plane moves into (the end slabs are unbounded: use their neighbour)
        src = r - 1 if excess > 0 else r
        lo, hi = planes[src], planes[src + 1]
        if not np.isfinite(lo) or not np.isfinite(hi):
            src = min(max(src + (1 if not np.isfinite(lo) else -1), 1), nr - 2) if nr > 2 else src
            lo, hi = planes[src], planes[src + 1]
        if not (np.isfinite(lo) and np.isfinite(hi)) or n[src] <= 0:
            continue
        shift = -gain * excess * (hi - lo) / n[src]
        new[r] = planes[r] + float(np.clip(shift, -0.4 * H, 0.4 * H))
    return new


def rebalance_cost(problem, planes, cost_ms, H, x_nodes=None, gain=0.8, iters=60):
    """Cost-weighted re-slabbing: moves the interior planes so that every rank's MEASURED pair work becomes equal (a rank
    that holds a wall pays more per particle than an interior one). planes: current planes (nranks + 1); cost_ms: what
    every rank spent in its per-particle kernels per step (the same list on every rank, e.g. an all-gather of the sums of
    Engine.profile_get() without the entries that contain waits); x_nodes: current x of the velocity particles (default:
    the problem's initial positions). Model: a rank's cost = its measured cost per processed velocity particle x the
    velocity particles in [lo - H, hi + H]. A plane moves by at most 0.4 H in total (spsph_dist_set_planes accepts 0.5 H)."""
    p = problem.params
    planes0 = np.array(planes, dtype=np.float64)
    nr = len(planes0) - 1
    if nr < 2:
        return planes0
    xn = np.sort(np.asarray(problem.arrays["x"][:p.nnode, 0] if x_nodes is None else x_nodes, dtype=np.float64))

    def processed(pl):
        lo = np.where(np.isfinite(pl[:-1]), pl[:-1] - H, -np.inf)
        hi = np.where(np.isfinite(pl[1:]), pl[1:] + H, np.inf)
        return (np.searchsorted(xn, hi) - np.searchsorted(xn, lo)).astype(np.float64)

    n0 = np.maximum(processed(planes0), 1.0)
    c = np.asarray(cost_ms, dtype=np.float64) / n0  # ms per processed velocity particle, kept fixed
    span = max(xn[-1] - xn[0], 1e-300)
    lam = len(xn) / span                           # velocity particles per unit x (lattice set-ups: uniform in x)
    pl = planes0.copy()
    for _ in range(iters):
        cost = c * processed(pl)
        for r in range(1, nr):
            # the more expensive side gives particles to the cheaper one
            d = gain * (cost[r] - cost[r - 1]) / ((c[r] + c[r - 1]) * lam)
            pl[r] = float(np.clip(pl[r] + d, planes0[r] - 0.4 * H, planes0[r] + 0.4 * H))
    return pl


def merge_owned(per_rank_arrays, per_rank_flags, params):
    """assemble the global state from every rank's download: entry i comes from the rank that owns particle i"""
    out = {k: v.copy() for k, v in per_rank_arrays[0].items()}
    seen = np.zeros(params.ntotal2, bool)
    for arrs, fl in zip(per_rank_arrays, per_rank_flags):
        own = fl == OWNED
        if (seen & own).any():
            raise AssertionError("a particle is owned by two ranks")
        seen |= own
        for k, v in arrs.items():
            n = v.shape[0]
            m = own[:n]
            out[k][m] = v[m]
    if not seen.all():
        raise AssertionError(f"{int((~seen).sum())} particles are owned by no rank")
    return out
