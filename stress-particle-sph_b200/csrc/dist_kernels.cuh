// Multi-GPU x-slab decomposition: ownership, halo selection, pack / unpack of the per-step exchange.
//
// Design (re-thought for NVLink): every rank holds full-size particle arrays indexed by the
// reference's global particle number but only its slab (owned particles) plus a WIDE halo of ghost particles is
// "local"; everything else is flagged remote and never enters the cell grid. Once per step, before the
// neighbour search, each rank sends the full state of its owned particles lying within the halo distance of a
// neighbouring slab (one fixed-capacity NCCL send/recv pair per side, no count round trip); after that the
// whole time step -- neighbour search, 4 RK stages, position update -- runs with NO further communication
// except one 6-double all-reduce for the global grid bounds and one 8-byte all-reduce of the pair count.
// Every sweep reads partners at most one cell (2h) away, so the ghost region computed redundantly degrades by
// one cell per dependent sweep; with a halo of (number of dependent sweeps + 2) cells the owned particles'
// results are bit-identical to the single-GPU run. Particles migrate between slabs as part of the same
// exchange (the receiver re-derives ownership from the position it receives).
#pragma once
#include "step_kernels.cuh"

namespace spsph {

// LF_STALE: a ghost of the previous step waiting to be refreshed by its owner (only between k_halo_select and
// k_list_compact; it is still on the local list, so k_halo_unpack must not append it again)
enum : int { LF_REMOTE = 0, LF_OWNED = 1, LF_GHOST = 2, LF_STALE = 3 };
// doubles per exchanged particle: id, x(2), vel(2), stress(4), eps_p, f_drucker, x_10(2), disp_10, displ(2), out flag, spare
constexpr int HALO_REC = 18;
// extended record (DistGeom::ext != 0): + rho, hsml, div u (continuity density: the density of a stress particle is
// integrated, its smoothing length follows with sle = 2, and density_update reads the last velocity divergence,
// main:706-713, 807-821) + bc_or_not, free-surface normal (get_nodes_on_free_surface every step: apply_stress_free
// and XSPH next to boundary conditions read the marks of the previous step, mat:1756-1839, main:224-230)
constexpr int HALO_REC_EXT = 24;
enum : int { HALO_EXT_DENSITY = 1, HALO_EXT_MARKS = 2 };

struct DistGeom {
  int rank, nranks;
  double lo, hi;  // this rank's slab is lo <= x_key < hi (lo = -inf on rank 0, hi = +inf on the last rank)
  double H;       // halo distance
  int sp_follows_node;  // outside approach: a stress particle is owned by the rank that owns its node
  int cap;        // record capacity of one halo message buffer
  int lim[2];     // records this step's message to the left / right neighbour may hold (<= cap)
  int rec;        // doubles per record: HALO_REC, or HALO_REC_EXT when ext != 0
  int ext;        // HALO_EXT_* bits: what the extended part of a record carries
};

// position that decides ownership: a stress particle of the outside approach follows its velocity particle,
// because shift_stress_points (main:244-368) re-seats it from that particle's thread
__device__ __forceinline__ double key_x(const DevParams &P, const DistGeom &D, const double *__restrict__ x, int i) {
  if (D.sp_follows_node && i >= P.nnode && i < P.ntotal) return x[2 * (size_t)((i - P.nnode) / P.npoints)];
  return x[2 * (size_t)i];
}

// initial flags after upload (all ranks hold identical, complete data)
__global__ void k_dist_init_flags(DevParams P, DistGeom D, const double *__restrict__ x, int *__restrict__ lflag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.ntotal2) return;
  const double xk = key_x(P, D, x, i), xi = x[2 * (size_t)i];
  int f = LF_REMOTE;
  if (xk >= D.lo && xk < D.hi)
    f = LF_OWNED;
  else if (xi >= D.lo - D.H && xi < D.hi + D.H)
    f = LF_GHOST;
  lflag[i] = f;
}

// flags after a row-wise upload (spsph_upload_rows): only the uploaded rows can be local
__global__ void k_dist_flags_rows(DevParams P, DistGeom D, const double *__restrict__ x, const int *__restrict__ ids, int n,
                                  int *__restrict__ lflag) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int i = ids[k];
  const double xk = key_x(P, D, x, i), xi = x[2 * (size_t)i];
  int f = LF_REMOTE;
  if (xk >= D.lo && xk < D.hi)
    f = LF_OWNED;
  else if (xi >= D.lo - D.H && xi < D.hi + D.H)
    f = LF_GHOST;
  lflag[i] = f;
}

// ------------------------------------------------------------------------------------------------------
// The local list: particle numbers with lflag != LF_REMOTE, in no particular order (every consumer is
// order-independent: min/max, per-cell counts, the deterministic in-cell ranking of k_rank). Built once from the
// flags, then maintained per step by k_halo_select (marks stale ghosts) -> k_halo_unpack (appends newcomers) ->
// k_list_compact (drops ghosts that were not refreshed), so no per-step pass scales with the global count.
// ------------------------------------------------------------------------------------------------------
// Order-preserving compaction (keep flags -> exclusive scan -> scatter): the list starts in ascending particle
// number and stays nearly so (newcomers are appended), which keeps the per-particle passes streaming.
// ids_in == nullptr: identity (initial build from the flags of all particles).
__global__ void k_list_flags(const int *__restrict__ ids_in, const int *__restrict__ n_in, int nfull,
                             int *__restrict__ lflag, int *__restrict__ keep) {
  const int n = ids_in ? *n_in : nfull;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int i = ids_in ? ids_in[k] : k;
    const int f = lflag[i];
    const bool kp = (f == LF_OWNED || f == LF_GHOST);
    if (f == LF_STALE) lflag[i] = LF_REMOTE;  // a ghost nobody refreshed: it left our halo
    keep[k] = kp ? 1 : 0;
  }
}
__global__ void k_list_scatter(const int *__restrict__ ids_in, const int *__restrict__ n_in, int nfull,
                               const int *__restrict__ keep, const int *__restrict__ pos, int *__restrict__ ids_out,
                               int *__restrict__ n_out) {
  const int n = ids_in ? *n_in : nfull;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    if (keep[k]) ids_out[pos[k]] = ids_in ? ids_in[k] : k;
    if (k == n - 1) *n_out = pos[k] + keep[k];
  }
  if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) *n_out = 0;
}

// halo selection + migration (see file header). cnt[0]/cnt[1]: number of records for the left/right neighbour.
__global__ void k_halo_select(DevParams P, DistGeom D, const double *__restrict__ x, int *__restrict__ lflag,
                              LocalList LL, int *__restrict__ cnt, int *__restrict__ idsL, int *__restrict__ idsR,
                              int *__restrict__ err) {
  SPSPH_FOR_LOCAL(LL, kk, i) {
    const int f = lflag[i];
    if (f == LF_GHOST) {
      lflag[i] = LF_STALE;  // refreshed by its owner below if it is still inside our halo
      continue;
    }
    if (f != LF_OWNED) continue;
    const double xk = key_x(P, D, x, i), xi = x[2 * (size_t)i];
    bool toL = false, toR = false;
    if (xk < D.lo) {  // migrates to the left neighbour; we keep it as a ghost (its state is current)
      toL = true;
      lflag[i] = LF_GHOST;
    } else if (xk >= D.hi) {
      toR = true;
      lflag[i] = LF_GHOST;
    } else {
      if (D.rank > 0 && xi < D.lo + D.H) toL = true;
      if (D.rank < D.nranks - 1 && xi >= D.hi - D.H) toR = true;
    }
    if (toL) {
      const int k = atomicAdd(&cnt[0], 1);
      if (k < D.lim[0])
        idsL[k] = i;
      else
        *err = 1;
    }
    if (toR) {
      const int k = atomicAdd(&cnt[1], 1);
      if (k < D.lim[1])
        idsR[k] = i;
      else
        *err = 1;
    }
  }
}

struct HaloArrays {
  double *x, *epsp, *fdp, *x_10, *disp_10, *displ;
  int *if_out;
  // extended record
  double *rho, *hsml, *mor, *divu;
  double2 *mrho;
  int *bc_or_not;
  double *fs_normal;
};

// message layout: record 0 = header {count}, records 1..count = particles
__global__ void k_halo_pack(DevParams P, DistGeom D, StatePtrs st, HaloArrays A, const int *__restrict__ cnt_ptr, int cap,
                            const int *__restrict__ ids, double *__restrict__ msg) {
  const int n = min(*cnt_ptr, cap);
  if (blockIdx.x == 0 && threadIdx.x == 0) msg[0] = (double)n;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int i = ids[k];
    double *o = msg + (size_t)D.rec * (k + 1);
    if (D.ext) {
      for (int q = HALO_REC; q < HALO_REC_EXT; ++q) o[q] = 0.0;
      if (D.ext & HALO_EXT_DENSITY) {
        o[18] = A.rho[i];
        o[19] = A.hsml[i];
        if (i >= P.nnode && i < P.ntotal) o[20] = A.divu[i - P.nnode];
      }
      if ((D.ext & HALO_EXT_MARKS) && i < P.ntotal) {
        o[21] = (double)A.bc_or_not[i];
        if (i < P.nnode) {
          o[22] = A.fs_normal[2 * (size_t)i];
          o[23] = A.fs_normal[2 * (size_t)i + 1];
        }
      }
    }
    o[0] = (double)i;
    o[1] = A.x[2 * (size_t)i];
    o[2] = A.x[2 * (size_t)i + 1];
    o[16] = (double)A.if_out[i];
    o[17] = 0.0;
    if (i < P.nnode) {
      const Rec4 r = ldrec(st.NB, i);
      const Stress4 s = ld4(st.NSb, i);
      o[3] = r.a;
      o[4] = r.b;
      o[5] = s.s1;
      o[6] = s.s2;
      o[7] = s.s3;
      o[8] = s.s4;
      o[9] = st.epsp[i];
      o[10] = st.fdp[i];
      o[11] = A.x_10[2 * (size_t)i];
      o[12] = A.x_10[2 * (size_t)i + 1];
      o[13] = A.disp_10[i];
      o[14] = A.displ[2 * (size_t)i];
      o[15] = A.displ[2 * (size_t)i + 1];
    } else if (i < P.ntotal) {
      const int ks = i - P.nnode;
      const Rec4 v = ldrec(st.SVb, ks);
      const Stress4 s = ld4(st.SFb, ks);
      o[3] = v.a;
      o[4] = v.b;
      o[5] = s.s1;
      o[6] = s.s2;
      o[7] = s.s3;
      o[8] = s.s4;
      o[9] = st.epsp[i];
      o[10] = st.fdp[i];
      o[11] = o[12] = o[13] = o[14] = o[15] = 0.0;
    } else {
      for (int q = 3; q < 16; ++q) o[q] = 0.0;
    }
  }
}

__global__ void k_halo_unpack(DevParams P, DistGeom D, StatePtrs st, HaloArrays A, const double *__restrict__ msg,
                              int *__restrict__ lflag, int *__restrict__ list_ids, int *__restrict__ list_n) {
  const int n = (int)msg[0];
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const double *o = msg + (size_t)D.rec * (k + 1);
    const int i = (int)o[0];
    A.x[2 * (size_t)i] = o[1];
    A.x[2 * (size_t)i + 1] = o[2];
    A.if_out[i] = (int)o[16];
    double rho_i = st.rho[i], mor_i = st.mor[i];
    if (D.ext & HALO_EXT_DENSITY) {  // the density (and smoothing length) travel with the particle
      rho_i = o[18];
      mor_i = st.mass[i] / rho_i;
      A.rho[i] = rho_i;
      A.hsml[i] = o[19];
      A.mor[i] = mor_i;
      A.mrho[i] = make_double2(st.mass[i], rho_i);
      if (i >= P.nnode && i < P.ntotal) A.divu[i - P.nnode] = o[20];
    }
    if ((D.ext & HALO_EXT_MARKS) && i < P.ntotal) {
      A.bc_or_not[i] = (int)o[21];
      if (i < P.nnode) {
        A.fs_normal[2 * (size_t)i] = o[22];
        A.fs_normal[2 * (size_t)i + 1] = o[23];
      }
    }
    if (i < P.nnode) {
      strec(st.NB, i, o[3], o[4], st.mass[i], rho_i);
      st4(st.NSb, i, Stress4{o[5], o[6], o[7], o[8]});
      st.epsp[i] = o[9];
      st.fdp[i] = o[10];
      A.x_10[2 * (size_t)i] = o[11];
      A.x_10[2 * (size_t)i + 1] = o[12];
      A.disp_10[i] = o[13];
      A.displ[2 * (size_t)i] = o[14];
      A.displ[2 * (size_t)i + 1] = o[15];
    } else if (i < P.ntotal) {
      const int ks = i - P.nnode;
      strec(st.SVb, ks, o[3], o[4], mor_i, 0.0);
      const Stress4 s{o[5], o[6], o[7], o[8]};
      st4(st.SFb, ks, s);
      const double r = rho_i;
      const double r2 = r * r;
      strec(st.SB, ks, s.s1 / r2, s.s2 / r2, s.s3 / r2, st.mass[i]);
      st.epsp[i] = o[9];
      st.fdp[i] = o[10];
    }
    if (lflag[i] == LF_REMOTE) list_ids[atomicAdd(list_n, 1)] = i;  // newcomer: joins the local list
    lflag[i] = LF_GHOST;  // ownership is settled by k_halo_own once every position has arrived
  }
}

// receiver side of migration: a received particle whose key position lies in our slab becomes owned
__global__ void k_halo_own(DevParams P, DistGeom D, const double *__restrict__ x, const double *__restrict__ msg,
                           int *__restrict__ lflag) {
  const int n = (int)msg[0];
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int i = (int)msg[(size_t)D.rec * (k + 1)];
    const double xk = key_x(P, D, x, i);
    if (xk >= D.lo && xk < D.hi) lflag[i] = LF_OWNED;
  }
}

// local bounding box -> 6 values encoded for a single MAX all-reduce: {-xmin, -ymin, xmax, ymax, hmax, -hmin}
__global__ void k_bbox_final(int nblocks, const double *__restrict__ partial, double *__restrict__ bb6) {
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;  // one warp; min / max are order-independent
  double mn0 = 1.e+10, mn1 = 1.e+10, mx0 = -1.e+10, mx1 = -1.e+10, hmx = 0.0, hmn = 1.e+300;
  for (int b = threadIdx.x; b < nblocks; b += 32) {
    const double *o = partial + 6 * b;
    mn0 = fmin(mn0, o[0]);
    mn1 = fmin(mn1, o[1]);
    mx0 = fmax(mx0, o[2]);
    mx1 = fmax(mx1, o[3]);
    hmx = fmax(hmx, o[4]);
    hmn = fmin(hmn, o[5]);
  }
  mn0 = warp_min(mn0);
  mn1 = warp_min(mn1);
  mx0 = warp_max(mx0);
  mx1 = warp_max(mx1);
  hmx = warp_max(hmx);
  hmn = warp_min(hmn);
  if (threadIdx.x == 0) {
    bb6[0] = -mn0;
    bb6[1] = -mn1;
    bb6[2] = mx0;
    bb6[3] = mx1;
    bb6[4] = hmx;
    bb6[5] = -hmn;
  }
}

// ------------------------------------------------------------------------------------------------------
// List-growth rule across slabs: find the pair whose GLOBAL creation index equals the old list
// capacity M. Creation order = earlier member ascending by (cell id, species, particle number); every rank
// knows, for its owned particles, the number of pairs they open (nfwd_u, local exclusive scan base_u). The
// search narrows row -> cell -> particle with one small collective per level; everything stays on the stream.
// ------------------------------------------------------------------------------------------------------
constexpr int GT_CAP = 1 << 16;  // rows / cells per row the search can handle
constexpr int GT_PCAP = 256;     // pair-opening particles of one cell listed per rank

struct GtSel {
  long long rem;  // remaining 1-based index inside the selected row / cell / particle
  int row, cell, err, pad;
};

__device__ __forceinline__ int unified_of_cell(const SortArrays &S, int c) {
  return S.start[0][c] + S.start[1][c] + S.start[2][c];
}
__device__ __forceinline__ long long gt_base(const SortArrays &S, const GridInfo *G, const int *base_u,
                                             long long local_total, int c) {  // pairs opened in cells < c
  const int U = unified_of_cell(S, c);
  const int nact = unified_of_cell(S, G->ncell);
  return U < nact ? (long long)base_u[U] : local_total;
}

__global__ void k_gt_rows(const GridInfo *__restrict__ G, SortArrays S, const int *__restrict__ base_u,
                          const long long *__restrict__ local_total, long long *__restrict__ rows) {
  const int y = blockIdx.x * blockDim.x + threadIdx.x;
  if (y >= GT_CAP) return;
  const int ndx = G->ndivx[0], ndy = G->ndivx[1];
  long long v = 0;
  if (y < ndy) v = gt_base(S, G, base_u, *local_total, (y + 1) * ndx) - gt_base(S, G, base_u, *local_total, y * ndx);
  rows[y] = v;
}
__global__ void k_gt_pick_row(const GridInfo *__restrict__ G, const long long *__restrict__ rows, long long M,
                              GtSel *__restrict__ sel) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  long long acc = 0;
  sel->err = (G->ndivx[0] > GT_CAP || G->ndivx[1] > GT_CAP) ? 1 : 0;
  sel->row = 0;
  sel->rem = 0;
  for (int y = 0; y < min(G->ndivx[1], GT_CAP); ++y) {
    if (acc + rows[y] >= M) {
      sel->row = y;
      sel->rem = M - acc;
      return;
    }
    acc += rows[y];
  }
  sel->err = 1;
}
__global__ void k_gt_cells(const GridInfo *__restrict__ G, SortArrays S, const int *__restrict__ base_u,
                           const long long *__restrict__ local_total, const GtSel *__restrict__ sel,
                           long long *__restrict__ cells) {
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  if (xx >= GT_CAP) return;
  const int ndx = G->ndivx[0];
  long long v = 0;
  if (xx < ndx) {
    const int c = sel->row * ndx + xx;
    v = gt_base(S, G, base_u, *local_total, c + 1) - gt_base(S, G, base_u, *local_total, c);
  }
  cells[xx] = v;
}
// picks the cell, then lists this rank's pair-opening particles of that cell as (key, count)
__global__ void k_gt_pick_cell(const GridInfo *__restrict__ G, SortArrays S, const int *__restrict__ nfwd_u,
                               const long long *__restrict__ cells, GtSel *__restrict__ sel,
                               long long *__restrict__ mine /* [GT_PCAP][2] */) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int ndx = G->ndivx[0];
  long long acc = 0;
  int found = -1;
  for (int xx = 0; xx < min(ndx, GT_CAP); ++xx) {
    if (acc + cells[xx] >= sel->rem) {
      found = xx;
      break;
    }
    acc += cells[xx];
  }
  for (int i = 0; i < 2 * GT_PCAP; ++i) mine[i] = -1;
  if (found < 0) {
    sel->err = 1;
    return;
  }
  const int c = sel->row * ndx + found;
  sel->cell = c;
  sel->rem = sel->rem - acc;
  int n = 0;
  for (int sp = 0; sp < 3; ++sp)
    for (int k = S.start[sp][c]; k < S.start[sp][c + 1]; ++k) {
      const int nf = nfwd_u[unified_slot(S, c, sp, k)];
      if (nf <= 0) continue;  // ghosts and particles that open no pair
      if (n >= GT_PCAP) {
        sel->err = 1;
        return;
      }
      mine[2 * n] = (long long)make_key(c, sp, S.order[sp][k]);
      mine[2 * n + 1] = nf;
      ++n;
    }
}
// all = gathered lists of every rank; picks the particle and, on the rank that owns it, its rem-th forward partner
__global__ void k_gt_pick_particle(DevParams P, const GridInfo *__restrict__ G, SortArrays S,
                                   const int *__restrict__ pos_of, const long long *__restrict__ all, int nranks,
                                   int myrank, GtSel *__restrict__ sel, long long *__restrict__ out2) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  out2[0] = 0;
  out2[1] = 0;
  out2[2] = 0;
  const int n = nranks * GT_PCAP;
  long long acc = 0, last = -1;
  for (;;) {  // walk the entries in ascending key order (selection without sorting; n is small)
    long long best = -1;
    int bi = -1;
    for (int i = 0; i < n; ++i) {
      const long long kk = all[2 * i];
      if (kk < 0 || kk <= last) continue;
      if (bi < 0 || kk < best) {
        best = kk;
        bi = i;
      }
    }
    if (bi < 0) {
      sel->err = 1;
      return;
    }
    const long long nf = all[2 * bi + 1];
    if (acc + nf >= sel->rem) {
      if (bi / GT_PCAP == myrank) {
        const okey_t ka = (okey_t)best;
        const int c = key_cell(ka), sp = (int)((ka >> 32) & 3), id = (int)(ka & 0xffffffffu);
        out2[0] = best;
        out2[2] = 1;  // found flag (a key may legitimately be 0)
        out2[1] = (long long)mth_forward_partner(P, G, S, c, sp, pos_of[id], (int)(sel->rem - acc));
      }
      return;
    }
    acc += nf;
    last = best;
  }
}
__global__ void k_gt_finish(const long long *__restrict__ out2, const GtSel *__restrict__ sel, GrowthRule *__restrict__ g,
                            int *__restrict__ err) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  g->mode = 2;
  g->ka = (okey_t)out2[0];
  g->kb = (okey_t)out2[1];
  if (sel->err || out2[2] != 1) *err = 1;
}

// ------------------------------------------------------------------------------------------------------
// Halo peeling. The ghost region is computed redundantly and loses one cut-off distance u of validity per dependent
// sweep, so a ghost at depth floor(d / u) (d = distance from this rank's slab) stops mattering once more sweeps have
// run than the halo is cells deep. The pair-sum kernels take the list lengths as an argument: handing them copies in
// which the lengths of the ghosts that are too deep for a sweep are zero skips those ghosts' list walks without any
// change to the kernels. Level l = 1 .. PEEL_LEVELS serves the sweeps 3 l + 1 .. 3 l + 3 of a step (1-based) and keeps
// depth <= halo_cells - (3 l + 1); the sweeps 1 .. 3 use the original arrays.
//   Bit-identical on every particle that is still computed, by induction over the sweeps: a particle computed in
//   sweep j has d < (halo_cells - j + 1) u, its partners are at most u further out, i.e. inside what sweep j - 1 computed.
// ------------------------------------------------------------------------------------------------------
constexpr int PEEL_LEVELS = 3;
__host__ __device__ inline int peel_max_depth(int halo_cells, int level) {
  const int m = halo_cells - (3 * level + 1);
  return m > 0 ? m : 0;  // owned particles (depth 0) are never peeled
}
__global__ void k_peel_counts(SlotMap M, SortArrays S, DistGeom D, int halo_cells, const int *__restrict__ n0,
                              const int *__restrict__ n1, int *__restrict__ p0, int *__restrict__ p1, size_t stride) {
  const double inv_u = (double)halo_cells / D.H;
  const int T = M.nnp + M.nsp;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    int sp, k;
    if (!slot_decode(M, t, sp, k)) continue;  // padding, or a slot past this rank's local particles: never read
    const double x = S.pos[sp][k].x;
    const double d = fmax(fmax(D.lo - x, x - D.hi), 0.0);
    const double q = d * inv_u;
    const int depth = q < 1.0e6 ? (int)q : 1000000;
    const int a = n0[t], b = n1[t];
#pragma unroll
    for (int l = 1; l <= PEEL_LEVELS; ++l) {
      const bool keep = depth <= peel_max_depth(halo_cells, l);
      p0[(size_t)(l - 1) * stride + t] = keep ? a : 0;
      p1[(size_t)(l - 1) * stride + t] = keep ? b : 0;
    }
  }
}

}  // namespace spsph
