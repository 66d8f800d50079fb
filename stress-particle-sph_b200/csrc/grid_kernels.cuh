// Neighbour search: cell grid, counting sort of velocity / stress / dummy particles, per-particle ordered
// neighbour enumeration and the fp32 pair weights (reference: Check_Out_Domain main:1170-1194,
// grid_find_NEW main:1199-1435, kernel main:1440-1538, Pint_Update mat:1574-1634).
//
// Design (not a port of the linked list): the reference's global pair list becomes, per particle, gather
// lists stored as warp-sliced ELL (32 particles per slice, entry e of lane l at off + 32*e + l) so that one
// thread per particle walks its own partners in the reference's traversal order with coalesced loads and no
// atomics in any sum. Restricted to one particle, the reference's creation order is "partners sorted by
// (cell id, particle index)" (SURVEY.md App. B), which is exactly the order a cell-sorted enumeration
// produces; the list re-use rule (new list nodes prepended => visited first, reversed) is applied when the
// entries are written.
#pragma once
#include "dev_common.cuh"

namespace spsph {

constexpr int SLICE = 32;
// The two top bits of a partner id of list 0 may carry the partner's mass/rho class (see step_kernels.cuh, ell_stream)
constexpr int QCLASS_SHIFT = 30;
constexpr int QID_MASK = (1 << QCLASS_SHIFT) - 1;

struct SortArrays {  // per species s in {node, stress, dummy}; species-sorted index k
  const int *start[3];      // [ncell+1] first sorted index of each cell
  const int *order[3];      // [n_s] original 0-based particle id
  const int *ordc[3];       // [n_s] the same id with the particle's mass/rho class in the two top bits (0 without palette)
  const double2 *pos[3];    // [n_s] positions
  const float2 *upos[3];    // [n_s] positions in cell units relative to the grid origin (fp32 prefilter only)
  const double *h[3];       // [n_s] smoothing lengths
  const int *cell[3];       // [n_s] cell id (or -1 for out-of-domain particles parked at the end)
};

// thread-slot space: nodes [0,NNp), stress particles [NNp, NNp+NSp), dummies after; NNp, NSp multiples of 32
struct SlotMap {
  int nn, ns, nd;     // particle counts
  int nnp, nsp, ndp;  // padded to multiples of SLICE
  __host__ __device__ int total() const { return nnp + nsp + ndp; }
};
__device__ __forceinline__ bool slot_decode(const SlotMap &m, int t, int &sp, int &k) {
  if (t < m.nnp) {
    sp = SP_NODE;
    k = t;
    return k < m.nn;
  }
  t -= m.nnp;
  if (t < m.nsp) {
    sp = SP_STRESS;
    k = t;
    return k < m.ns;
  }
  t -= m.nsp;
  sp = SP_DUMMY;
  k = t;
  return k < m.nd;
}

// Particles a kernel has to visit. Single GPU: all of them (ids == nullptr, identity). Multi-GPU: the compact
// list of this rank's local (owned + ghost) particle numbers kept by dist_kernels.cuh, so that no per-step pass
// is proportional to the global particle count.
struct LocalList {
  const int *ids;  // nullptr: identity
  const int *n;    // device count (ids != nullptr)
  int nfull;
};
__device__ __forceinline__ int ll_count(const LocalList &l) { return l.ids ? *l.n : l.nfull; }
__device__ __forceinline__ int ll_id(const LocalList &l, int k) { return l.ids ? l.ids[k] : k; }
#define SPSPH_FOR_LOCAL(LL, K, I)                                                                       \
  for (int K = blockIdx.x * blockDim.x + threadIdx.x, n_ll_ = ll_count(LL), I = 0;                      \
       K < n_ll_ && ((I = ll_id(LL, K)), true); K += gridDim.x * blockDim.x)

// ------------------------------------------------------------------------------------------------------
// Check_Out_Domain + bounding box / max h of the in-domain particles (min/max are order-independent).
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_min(double v) {
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// lflag (multi-GPU only, else nullptr): 0 remote, 1 owned, 2 ghost; only owned particles enter the local
// bounds (the global bounds are the all-reduce of the local ones), remote particles are skipped entirely.
__global__ void k_domain_bbox(DevParams P, const double *__restrict__ x, const double *__restrict__ hsml,
                              int *__restrict__ if_out, const int *__restrict__ lflag, LocalList LL,
                              double *__restrict__ partial /* [gridDim.x][6] */) {
  double xmn = 1.e+10, ymn = 1.e+10, xmx = -1.e+10, ymx = -1.e+10, hmx = 0.0, hmn = 1.e+300;
  SPSPH_FOR_LOCAL(LL, kk, i) {
    const int lf = lflag ? lflag[i] : 1;
    if (lf == 0) continue;
    const double2 p = ld2(x, i);
    int out = if_out[i];
    const double dxx = (p.x - P.xmin_dom[0]) * (p.x - P.xmax_dom[0]);
    const double dyy = (p.y - P.xmin_dom[1]) * (p.y - P.xmax_dom[1]);
    if (dxx > 0.0 || dyy > 0.0) {
      if (!out) if_out[i] = 1;
      out = 1;
    }
    if (!out && lf == 1) {
      xmn = fmin(xmn, p.x);
      xmx = fmax(xmx, p.x);
      ymn = fmin(ymn, p.y);
      ymx = fmax(ymx, p.y);
      hmx = fmax(hmx, hsml[i]);
      hmn = fmin(hmn, hsml[i]);
    }
  }
  __shared__ double sh[6][32];
  xmn = warp_min(xmn);
  ymn = warp_min(ymn);
  xmx = warp_max(xmx);
  ymx = warp_max(ymx);
  hmx = warp_max(hmx);
  hmn = warp_min(hmn);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    sh[0][w] = xmn;
    sh[1][w] = ymn;
    sh[2][w] = xmx;
    sh[3][w] = ymx;
    sh[4][w] = hmx;
    sh[5][w] = hmn;
  }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    xmn = l < nw ? sh[0][l] : 1.e+10;
    ymn = l < nw ? sh[1][l] : 1.e+10;
    xmx = l < nw ? sh[2][l] : -1.e+10;
    ymx = l < nw ? sh[3][l] : -1.e+10;
    hmx = l < nw ? sh[4][l] : 0.0;
    hmn = l < nw ? sh[5][l] : 1.e+300;
    xmn = warp_min(xmn);
    ymn = warp_min(ymn);
    xmx = warp_max(xmx);
    ymx = warp_max(ymx);
    hmx = warp_max(hmx);
    hmn = warp_min(hmn);
    if (l == 0) {
      double *o = partial + 6 * blockIdx.x;
      o[0] = xmn;
      o[1] = ymn;
      o[2] = xmx;
      o[3] = ymx;
      o[4] = hmx;
      o[5] = hmn;
    }
  }
}

// grid_find_NEW Task 1 (main:1245-1258) on one thread. bb6 = {-xmin, -ymin, xmax, ymax, hmax, -hmin} of the
// in-domain particles (already all-reduced over the ranks in a multi-GPU run).
__global__ void k_grid_params(const double *__restrict__ bb6, GridInfo *__restrict__ G, int cell_capacity) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double mn[2] = {-bb6[0], -bb6[1]}, mx[2] = {bb6[2], bb6[3]};
  const double hmx = bb6[4], hmn = -bb6[5];
  for (int d = 0; d < 2; ++d) {
    double xmin = mn[d], xmax = mx[d];
    double deltx = hmx * 2;
    const double length = xmax - xmin;
    const int ndiv = (int)((length / deltx) + 1);
    const double length_new = ndiv * deltx;
    xmin = xmin - (length_new - length) / 2 - (double)0.001f * length;
    xmax = xmax + (length_new - length) / 2 + (double)0.001f * length;
    G->xmin[d] = xmin;
    G->xmax[d] = xmax;
    G->deltx[d] = deltx;
    G->ndivx[d] = ndiv;
    G->rxmin[d] = mn[d];
    G->rxmax[d] = mx[d];
  }
  G->rhmax = hmx;
  G->uniform_h = (hmn == hmx) ? 1 : 0;
  const long long nc = (long long)G->ndivx[0] * (long long)G->ndivx[1];
  G->overflow = (nc > (long long)cell_capacity || nc <= 0) ? 1 : 0;
  G->ncell = G->overflow ? 1 : (int)nc;
  if (G->overflow) {  // keep the remaining kernels in bounds; the host reports the error
    G->ndivx[0] = 1;
    G->ndivx[1] = 1;
  }
}

// zero the per-cell counters (3 species x (ncell+1)) -- size known only on the device
__global__ void k_zero_cells(const GridInfo *__restrict__ G, int *__restrict__ cnt, int cell_stride) {
  const int n = G->ncell + 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    cnt[i] = 0;
    cnt[cell_stride + i] = 0;
    cnt[2 * cell_stride + i] = 0;
  }
}

__device__ __forceinline__ int species_of(const DevParams &P, int i) {
  return i < P.nnode ? SP_NODE : (i < P.ntotal ? SP_STRESS : SP_DUMMY);
}

// grid_find_NEW Task 2 (main:1277-1286): cell id of every in-domain particle + per-cell species counts
__global__ void k_cell_id(DevParams P, const GridInfo *__restrict__ G, const double *__restrict__ x,
                          const int *__restrict__ if_out, LocalList LL, int *__restrict__ which_cell,
                          int *__restrict__ cnt, int cell_stride, int *__restrict__ nout /* [6] */) {
  SPSPH_FOR_LOCAL(LL, kk, i) {
    const int sp = species_of(P, i);
    if (if_out[i] || G->overflow) {  // on a cell-table overflow nothing is binned; the host reports the error
      which_cell[i] = -1 - atomicAdd(&nout[sp], 1);  // parked after the sorted particles, order irrelevant
      continue;
    }
    const double2 p = ld2(x, i);
    int ix = (int)((p.x - G->xmin[0]) / G->deltx[0] + 1);
    int iy = (int)((p.y - G->xmin[1]) / G->deltx[1] + 1);
    if (ix > G->ndivx[0]) ix = G->ndivx[0];
    if (iy > G->ndivx[1]) iy = G->ndivx[1];
    const int c = G->ndivx[0] * (iy - 1) + (ix - 1);  // 0-based cell id, same ordering as the reference's
    which_cell[i] = c;
    atomicAdd(&cnt[sp * cell_stride + c], 1);
  }
}

// ------------------------------------------------------------------------------------------------------
// Device-wide exclusive scan of int32 (three-kernel reduce / scan-of-sums / apply). `n` comes from device
// memory (n_ptr) plus a constant, so no host round trip is needed. blockIdx.y selects one of several
// independent rows (row stride in elements).
// ------------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_BLOCKS = 296;  // 2 per SM

__device__ __forceinline__ int block_excl_scan(int v, int *total) {
  __shared__ int wsum[SCAN_THREADS / 32];
  const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (l >= o) inc += t;
  }
  if (l == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    int s = l < SCAN_THREADS / 32 ? wsum[l] : 0;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, o);
      if (l >= o) s += t;
    }
    if (l < SCAN_THREADS / 32) wsum[l] = s;
  }
  __syncthreads();
  const int base = w > 0 ? wsum[w - 1] : 0;
  if (total) *total = wsum[SCAN_THREADS / 32 - 1];
  __syncthreads();
  return base + inc - v;
}

__device__ __forceinline__ void scan_chunk(int n, int &b, int &e) {
  const int per = (n + SCAN_BLOCKS - 1) / SCAN_BLOCKS;
  const int chunk = ((per + SCAN_THREADS - 1) / SCAN_THREADS) * SCAN_THREADS;
  const long long bb = (long long)blockIdx.x * chunk;
  b = bb < n ? (int)bb : n;
  e = (bb + chunk) < n ? (int)(bb + chunk) : n;
}

__global__ void k_scan_reduce(const int *__restrict__ in, int row_stride, const int *__restrict__ n_ptr, int n_add,
                              int *__restrict__ bsum) {
  const int n = (n_ptr ? *n_ptr : 0) + n_add;
  const int *row = in + (size_t)blockIdx.y * row_stride;
  int b, e;
  scan_chunk(n, b, e);
  int s = 0;
  for (int i = b + threadIdx.x; i < e; i += SCAN_THREADS) s += row[i];
  int tot;
  block_excl_scan(s, &tot);
  if (threadIdx.x == 0) bsum[blockIdx.y * SCAN_BLOCKS + blockIdx.x] = tot;
}
__global__ void k_scan_sums(int *__restrict__ bsum, long long *__restrict__ totals) {
  // one block per row; SCAN_BLOCKS <= 2*SCAN_THREADS
  int *row = bsum + blockIdx.x * SCAN_BLOCKS;
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < SCAN_BLOCKS; base += SCAN_THREADS) {
    const int i = base + threadIdx.x;
    const int v = i < SCAN_BLOCKS ? row[i] : 0;
    int tot;
    const int ex = block_excl_scan(v, &tot);
    const int c = carry;
    if (i < SCAN_BLOCKS) row[i] = c + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0 && totals) totals[blockIdx.x] = carry;
}
__global__ void k_scan_apply(const int *__restrict__ in, int *__restrict__ out, int row_stride,
                             const int *__restrict__ n_ptr, int n_add, const int *__restrict__ bsum) {
  const int n = (n_ptr ? *n_ptr : 0) + n_add;
  const int *row = in + (size_t)blockIdx.y * row_stride;
  int *orow = out + (size_t)blockIdx.y * row_stride;
  int b, e;
  scan_chunk(n, b, e);
  int carry = bsum[blockIdx.y * SCAN_BLOCKS + blockIdx.x];
  for (int base = b; base < e; base += SCAN_THREADS) {
    const int i = base + threadIdx.x;
    const int v = i < e ? row[i] : 0;
    int tot;
    const int ex = block_excl_scan(v, &tot);
    if (i < e) orow[i] = carry + ex;
    carry += tot;
  }
}

// ------------------------------------------------------------------------------------------------------
// Counting sort: scatter into cell segments (arbitrary order inside a cell), then rank inside the cell by
// original index so that the final order is deterministic and equals the reference's list_picell order
// (ascending particle index within a cell, main:1299-1305).
// ------------------------------------------------------------------------------------------------------
__global__ void k_scatter(DevParams P, LocalList LL, const int *__restrict__ which_cell, const int *__restrict__ start,
                          int *__restrict__ fill, int cell_stride, int *__restrict__ tmp /* 3 rows, stride ntotal2 */) {
  SPSPH_FOR_LOCAL(LL, kk, i) {
    const int c = which_cell[i];
    if (c < 0) continue;
    const int sp = species_of(P, i);
    const int slot = start[sp * cell_stride + c] + atomicAdd(&fill[sp * cell_stride + c], 1);
    tmp[(size_t)sp * P.ntotal2 + slot] = i;
  }
}

__global__ void k_rank(DevParams P, LocalList LL, const GridInfo *__restrict__ G, const double *__restrict__ x,
                       const double *__restrict__ hsml, const int *__restrict__ which_cell,
                       const int *__restrict__ start, int cell_stride, const int *__restrict__ tmp,
                       int *__restrict__ order, double2 *__restrict__ spos, double *__restrict__ sh,
                       int *__restrict__ scell, int *__restrict__ pos_of /* [ntotal2] species-sorted index */,
                       float2 *__restrict__ supos, const int *__restrict__ nout, const unsigned char *__restrict__ mcls,
                       int *__restrict__ ordc) {
  // one thread per (local) particle
  SPSPH_FOR_LOCAL(LL, kk, i) {
    const int sp = species_of(P, i);
    const int c = which_cell[i];
    const size_t row = (size_t)sp * P.ntotal2;
    int k;
    if (c < 0) {
      const int nact = start[sp * cell_stride + G->ncell];
      k = nact + (-1 - c);
    } else {
      const int b = start[sp * cell_stride + c], e = start[sp * cell_stride + c + 1];
      int r = 0;
      for (int j = b; j < e; ++j) r += (tmp[row + j] < i) ? 1 : 0;
      k = b + r;
    }
    order[row + k] = i;
    ordc[row + k] = mcls ? (i | ((int)mcls[i] << 30)) : i;  // what k_fill stores as partner id of list 0 (QCLASS_SHIFT)
    const double2 xi = ld2(x, i);
    supos[row + k] =
        make_float2((float)((xi.x - G->xmin[0]) / G->deltx[0]), (float)((xi.y - G->xmin[1]) / G->deltx[1]));
    spos[row + k] = xi;
    sh[row + k] = hsml[i];
    scell[row + k] = c < 0 ? -1 : c;
    pos_of[i] = k;
  }
}

// ------------------------------------------------------------------------------------------------------
// Neighbour enumeration.
// ------------------------------------------------------------------------------------------------------
struct Cand {  // one accepted partner
  int q;       // species-sorted index
  double dxq, dyq, r, mh;  // dx = x_p - x_q (p-perspective); caller flips for the reference orientation
};

// acceptance test of main:1345-1353 (symmetric in the two particles)
__device__ __forceinline__ bool pair_accept(const DevParams &P, double2 pp, double hp, double2 pq, double hq, double &dx,
                                            double &dy, double &r, double &mh) {
  dx = pp.x - pq.x;
  dy = pp.y - pq.y;
  double driac = dx * dx;
  driac = driac + dy * dy;
  mh = (hp + hq) / 2.;
  r = sqrt(driac);
  return r < P.scale_k * mh;
}

// unified (cell, species, index) slot of the reference's creation order
__device__ __forceinline__ int unified_slot(const SortArrays &S, int c, int sp, int k) {
  const int sn = S.start[0][c], ss = S.start[1][c], sd = S.start[2][c];
  int u = sn + ss + sd;
  if (sp >= SP_STRESS) u += S.start[0][c + 1] - sn;
  if (sp >= SP_DUMMY) u += S.start[1][c + 1] - ss;
  return u + (k - S.start[sp][c]);
}

// Growth rule of the reference's list: pairs whose creation index exceeds the old list
// capacity M are visited first and reversed. Particles are compared by the creation-order key
// (cell id, species, particle number) -- globally meaningful, so the same rule serves the multi-GPU slabs:
// pair (ka < kb) is "old" iff (ka, kb) <=lex (g.ka, g.kb), the keys of the pair with creation index M.
typedef unsigned long long okey_t;
__device__ __forceinline__ okey_t make_key(int cell, int sp, int id) {
  return ((okey_t)(unsigned)cell << 34) | ((okey_t)(unsigned)sp << 32) | (okey_t)(unsigned)id;
}
__device__ __forceinline__ int key_cell(okey_t k) { return (int)(k >> 34); }
struct GrowthRule {
  int mode;  // 0: all old (forward), 1: all new (fully reversed: first step), 2: split at (ka, kb)
  int pad;
  okey_t ka, kb;
};
__device__ __forceinline__ bool pair_is_old(const GrowthRule &g, okey_t k1, okey_t k2) {
  if (g.mode == 0) return true;
  if (g.mode == 1) return false;
  const okey_t a = k1 < k2 ? k1 : k2, b = k1 < k2 ? k2 : k1;
  return (a < g.ka) || (a == g.ka && b <= g.kb);
}
__device__ __forceinline__ okey_t sorted_key(const SortArrays &S, int sq, int q) {
  const int *cp = sq == 0 ? S.cell[0] : (sq == 1 ? S.cell[1] : S.cell[2]);
  const int *op = sq == 0 ? S.order[0] : (sq == 1 ? S.order[1] : S.order[2]);
  return make_key(cp[q], sq, op[q]);
}

struct KernelConsts {  // per-thread constants of `kernel` (main:1468-1492) for a fixed smoothing length h
  double h, rh, hh, rhh, factor, f6, m63;
};
__device__ __forceinline__ KernelConsts kernel_consts(const DevParams &P, double h) {
  KernelConsts K;
  K.h = h;
  K.rh = 1.0 / h;
  K.hh = h * h;
  K.rhh = 1.0 / K.hh;
  K.factor = 15.e0 / (7.e0 * P.pi * h * h);
  K.f6 = K.factor * 1.e0 / 6.e0;
  K.m63 = -K.f6 * 3.;
  return K;
}
// same values as sph_kernel(), bit for bit, for mhsml == K.h; GRAD = false skips the gradient
template <bool GRAD>
__device__ __forceinline__ void sph_kernel_fast(const KernelConsts &K, double r, double dx, double dy, double &w,
                                                double &gx, double &gy) {
  const double q = div_rn(r, K.h, K.rh);
  w = 0.;
  gx = 0.;
  gy = 0.;
  if (q >= 0 && q <= 1.e0) {
    w = K.factor * ((double)(2.f / 3.f) - q * q + q * q * q * 0.5);
    if (GRAD) {
      const double t = div_rn(K.factor * (-2. + 1.5 * q), K.hh, K.rhh);
      gx = t * dx;
      gy = t * dy;
    }
  } else if (q > 1.e0 && q <= 2) {
    const double t = 2. - q;
    w = K.f6 * (t * t * t);
    if (GRAD) {
      const double u = div_rn(K.m63 * (t * t), K.h, K.rh);
      const double rr = __drcp_rn(r);
      gx = u * div_rn(dx, r, rr);
      gy = u * div_rn(dy, r, rr);
    }
  }
}

// Acceptance test with a squared-distance prefilter: sqrt() only for candidates within 2e-15 (relative) of
// the cut-off, where the reference's `sqrt(driac) < scale_k*mhsml` decides. Returns the squared distance.
__device__ __forceinline__ bool pair_accept_fast(double scale_k, double2 pp, double hp, double2 pq, double hq,
                                                 double &dx, double &dy, double &driac, double &mh) {
  dx = pp.x - pq.x;
  dy = pp.y - pq.y;
  driac = dx * dx;
  driac = driac + dy * dy;
  mh = (hp + hq) / 2.;
  const double c = scale_k * mh;
  const double c2 = c * c;
  if (driac < c2 * 0.999999999999998) return true;
  if (driac > c2 * 1.000000000000002) return false;
  return sqrt(driac) < c;
}

// fp32 prefilter in cell units: classifies a candidate as sure-accept (1), sure-reject (0) or undecided (2).
// Only valid when all particles share one h (cut-off == scale_k*h) and the grid is small enough for fp32;
// `lo`/`hi` bracket the squared cut-off by the fp32 coordinate error (see prefilter_bounds()).
struct Prefilter {
  int on;
  float lo, hi;
};
__device__ __forceinline__ Prefilter prefilter_bounds(const DevParams &P, const GridInfo *G, double hp) {
  Prefilter f;
  const int nmax = max(G->ndivx[0], G->ndivx[1]);
  // coordinate error <= nmax * 2^-23 per component after the fp32 rounding and the subtraction
  const double eps = 8.0 * (double)nmax * 1.1920929e-07;
  const double c = (double)P.scale_k * hp / G->deltx[0];  // cut-off in cell units (== 1 for skf = 1)
  f.on = (G->uniform_h != 0 && eps < 0.05 * c && G->deltx[0] == G->deltx[1]) ? 1 : 0;
  const double a = (c - eps) > 0 ? (c - eps) : 0.0, b = c + eps;
  f.lo = (float)(a * a * 0.999);
  f.hi = (float)(b * b * 1.001);
  return f;
}
__device__ __forceinline__ int prefilter_test(const Prefilter &f, float2 up, float2 uq) {
  const float du = up.x - uq.x, dv = up.y - uq.y;
  const float d2 = __fmaf_rn(du, du, dv * dv);
  return d2 > f.hi ? 0 : (d2 < f.lo ? 1 : 2);
}

// One row of the 3x3 stencil: the cells (cx-1..cx+1, jy) are consecutive cell ids, so each species' particles
// of the whole row form ONE contiguous range of the species-sorted arrays.
struct RowRange {
  int ca, cb;  // first and last cell id of the row segment
};
__device__ __forceinline__ RowRange row_range(int ndx, int cx, int jy) {
  return RowRange{jy * ndx + max(cx - 1, 0), jy * ndx + min(cx + 1, ndx - 1)};
}

// Count pass: per thread slot the lengths of its gather lists, its forward pair count (creation index) and its
// total interaction count (countiac, main:1355-1356). The accepted partners are also recorded, in list order, in
// a fixed-capacity scratch (CAND_CAP rows per 32-particle slice) so that the fill pass does not have to search
// again; a particle with more partners than that raises `overflow` and the step falls back to k_fill_scan.
constexpr int CAND_CAP = 64;
// Scratch layout: entry c of lane l of a slice at row c, column l (the lanes of a warp that append at the same time
// share a 128-byte row). Measured alternatives on the 4 M column (k_count + k_fill = 1.00 + 1.49 ms with this layout):
// 16-byte groups per lane 1.59 + 1.49 (scattered 4-byte stores); acceptance masks instead of a scratch, one 64-bit
// word per (stencil row, species), walked bit by bit in k_fill: 0.87 + 1.98. Offsets are relative to cand_base(t).
__host__ __device__ __forceinline__ size_t cand_base(int t) { return (size_t)(t / SLICE) * CAND_CAP * SLICE + (size_t)(t & 31); }
__host__ __device__ __forceinline__ size_t cand_off(int c) { return (size_t)c * SLICE; }

#ifndef SPSPH_COUNT_MINB
#define SPSPH_COUNT_MINB 6  // measured on the 4 M column: 6 blocks (78 registers, no spills) 0.97 ms, 8: 1.00, 10: 1.19
#endif
#ifndef SPSPH_COUNT_PAIR
#define SPSPH_COUNT_PAIR 0  // 1: candidate positions loaded as 16-byte pairs (measured slower: 1.31 vs 1.00 ms)
#endif
#ifndef SPSPH_COUNT_LEAN
#define SPSPH_COUNT_LEAN 1  // 0: the general candidate loop only (round-2 mid-point kernel, kept for A/B timing)
#endif
__global__ void __launch_bounds__(128, SPSPH_COUNT_MINB)
k_count(DevParams P, SlotMap M, const GridInfo *__restrict__ G, SortArrays S, int *__restrict__ n0, int *__restrict__ n1,
        int *__restrict__ nfwd_u /* unified order */, int *__restrict__ nall, int *__restrict__ w0 /* slice widths */,
        int *__restrict__ wC, int *__restrict__ wD, const int *__restrict__ lflag, int *__restrict__ cand0,
        int *__restrict__ cand1, int *__restrict__ overflow, const int *__restrict__ nout, int t0, int tn) {
  // slots [t0, t0 + tn): all of them on a single GPU; on a slab one launch per species over the leading slots that
  // can hold local particles (slices that are not visited keep the zero width of the host memset)
  const int t = t0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= t0 + tn) return;
  int sp = 0, k = 0;
  bool live = (t < M.total()) && slot_decode(M, t, sp, k);
  // only the leading slots of a species are occupied this step (in-grid particles, then out-of-domain ones)
  // (single GPU: every particle is sorted, nothing to check)
  if (live && lflag) live = k < (sp == 0 ? S.start[0] : (sp == 1 ? S.start[1] : S.start[2]))[G->ncell] + nout[sp];
  int c0 = 0, c1 = 0, cf = 0, ca = 0;
  if (live) {
    const int *__restrict__ cellp = sp == 0 ? S.cell[0] : (sp == 1 ? S.cell[1] : S.cell[2]);
    const int c = cellp[k];
    if (c >= 0) {
      const double2 *__restrict__ posp = sp == 0 ? S.pos[0] : (sp == 1 ? S.pos[1] : S.pos[2]);
      const double *__restrict__ hpp = sp == 0 ? S.h[0] : (sp == 1 ? S.h[1] : S.h[2]);
      const double2 pp = posp[k];
      const double hp = hpp[k];
      const bool uni = G->uniform_h != 0;
      const double sk = (double)P.scale_k;
      const Prefilter pf = prefilter_bounds(P, G, hp);
      const float2 *__restrict__ uposp = sp == 0 ? S.upos[0] : (sp == 1 ? S.upos[1] : S.upos[2]);
      const float2 up = uposp[k];
      const int ndx = G->ndivx[0], ndy = G->ndivx[1];
      const int cy = c / ndx, cx = c - cy * ndx;
      const size_t cb = cand_base(t);
      const bool owner_of_lists = sp != SP_DUMMY;
      auto take = [&](int sq, int q, int fthr) {  // accepted partner, in list order
        ++ca;
        cf += (q >= fthr) ? 1 : 0;
        if (owner_of_lists) {
          if (sq == sp) {
            if (c1 < CAND_CAP) cand1[cb + cand_off(c1)] = q;
            ++c1;
          } else {  // node<->stress (type 1) and node/stress<->dummy (types 6, 9)
            if (c0 < CAND_CAP) cand0[cb + cand_off(c0)] = sq * P.ntotal2 + q;
            ++c0;
          }
        }
      };
      auto scan = [&](int sq, int b, int e, int fthr) {
        const float2 *__restrict__ uq = sq == 0 ? S.upos[0] : (sq == 1 ? S.upos[1] : S.upos[2]);
        const double2 *__restrict__ pq = sq == 0 ? S.pos[0] : (sq == 1 ? S.pos[1] : S.pos[2]);
        const double *__restrict__ hq = sq == 0 ? S.h[0] : (sq == 1 ? S.h[1] : S.h[2]);
        auto exact = [&](int q) {
          double dx, dy, d2, mh;
          return pair_accept_fast(sk, pp, hp, pq[q], uni ? hp : hq[q], dx, dy, d2, mh);
        };
        int q = b;
        if (pf.on) {
          // four candidates per iteration: independent loads, one combined "anything to do" test
          for (; q + 4 <= e; q += 4) {
            float2 u4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) u4[u] = uq[q + u];
            int cls[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) cls[u] = prefilter_test(pf, up, u4[u]);
            if ((cls[0] | cls[1] | cls[2] | cls[3]) == 0) continue;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (cls[u] == 0 || (sq == sp && q + u == k)) continue;
              if (cls[u] == 2 && !exact(q + u)) continue;
              take(sq, q + u, fthr);
            }
          }
        }
        for (; q < e; ++q) {
          if (sq == sp && q == k) continue;
          int cls = 2;
          if (pf.on) cls = prefilter_test(pf, up, uq[q]);
          if (cls == 0) continue;
          if (cls == 2 && !exact(q)) continue;
          take(sq, q, fthr);
        }
      };
#if SPSPH_COUNT_LEAN
      // Lean scan (the common case: fp32 prefilter valid, list-owning particle). The same candidates in the same
      // order as the general loop below, but the range of one (row | cell, species) is cut beforehand at the
      // particle itself and at the first forward partner, so that the candidate loop carries no per-candidate
      // index tests or counters: 4 candidates per trip, classification into predicates, one rare branch for
      // candidates inside the fp32 uncertainty band, predicated stores. Forward partners are counted by difference.
      if (pf.on && owner_of_lists) {
        const float plo = pf.lo, phi = pf.hi;
        // a group of four goes the careful way when any of its squared distances lies within [plo, phi], tested as
        // |d2 - pmid| <= phalf with the half-width widened beyond the fp32 rounding of pmid and of the difference
        const float pmid = 0.5f * (plo + phi), phalf = 0.5f * (phi - plo) * 1.001f + 4.e-6f * phi;
        for (int jy = max(cy - 1, 0); jy <= min(cy + 1, ndy - 1); ++jy) {
          const RowRange rr = row_range(ndx, cx, jy);
          const bool merged = (S.start[2][rr.cb + 1] - S.start[2][rr.ca]) == 0;  // no wall particles in this row
          const int nsq = merged ? 2 : 3;
          for (int cq0 = rr.ca; cq0 <= rr.cb;) {
            const int cq1 = merged ? rr.cb : cq0;
#pragma unroll 1
            for (int sq = 0; sq < nsq; ++sq) {
              const int *__restrict__ stq = sq == 0 ? S.start[0] : (sq == 1 ? S.start[1] : S.start[2]);
              const int b = stq[cq0], e = stq[cq1 + 1];
              if (b == e) continue;
              const bool same = sq == sp;
              int m;  // first forward partner of the range
              if (jy > cy)
                m = b;
              else if (jy < cy)
                m = e;
              else
                m = min(max(same ? k + 1 : (sq > sp ? stq[c] : stq[c + 1]), b), e);
              const int back_end = (same && jy == cy && k >= b && k < e) ? k : m;  // the particle itself: m == k + 1
              const float2 *__restrict__ uq = sq == 0 ? S.upos[0] : (sq == 1 ? S.upos[1] : S.upos[2]);
              int *__restrict__ cp = (same ? cand1 : cand0) + cb;
              const int tag = same ? 0 : sq * P.ntotal2;  // cross-species partners: index into the unified sorted arrays
              int cc = same ? c1 : c0;
              auto exact = [&](int q) {
                const double2 *__restrict__ pq = sq == 0 ? S.pos[0] : (sq == 1 ? S.pos[1] : S.pos[2]);
                const double *__restrict__ hq = sq == 0 ? S.h[0] : (sq == 1 ? S.h[1] : S.h[2]);
                double dx, dy, d2, mh;
                return pair_accept_fast(sk, pp, hp, pq[q], uni ? hp : hq[q], dx, dy, d2, mh);
              };
              auto one = [&](int q) {  // one candidate, the careful way
                const float2 a = uq[q];
                const float du = up.x - a.x, dv = up.y - a.y;
                const float d2 = __fmaf_rn(du, du, dv * dv);
                if (d2 > phi) return;
                if (!(d2 < plo) && !exact(q)) return;
                if (cc < CAND_CAP) cp[cand_off(cc)] = tag + q;
                ++cc;
              };
              auto seg = [&](int lo, int hi) {
                int q = lo;
#if SPSPH_COUNT_PAIR
                // candidates come in 16-byte pairs: start on a pair boundary
                if (q < hi && (reinterpret_cast<size_t>(uq + q) & 8)) one(q++);
#endif
                for (;;) {
                  // fast groups: four candidates, all of them clear of the uncertainty band
                  for (; q + 4 <= hi && cc + 4 <= CAND_CAP; q += 4) {
#if SPSPH_COUNT_PAIR
                    const float4 a01 = *reinterpret_cast<const float4 *>(uq + q);
                    const float4 a23 = *reinterpret_cast<const float4 *>(uq + q + 2);
                    const float ax[4] = {a01.x, a01.z, a23.x, a23.z}, ay[4] = {a01.y, a01.w, a23.y, a23.w};
#else
                    float ax[4], ay[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                      const float2 a = uq[q + u];
                      ax[u] = a.x;
                      ay[u] = a.y;
                    }
#endif
                    float d2[4], off[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                      const float du = up.x - ax[u], dv = up.y - ay[u];
                      d2[u] = __fmaf_rn(du, du, dv * dv);
                      off[u] = fabsf(d2[u] - pmid);
                    }
                    if (fminf(fminf(off[0], off[1]), fminf(off[2], off[3])) <= phalf) break;  // rare: one by one
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                      if (d2[u] < plo) {
                        cp[cand_off(cc)] = tag + (q + u);
                        ++cc;
                      }
                  }
                  if (q >= hi) break;
                  // up to four candidates one by one: a group that touches the band, the tail of the range, and
                  // every candidate once the scratch is nearly full
                  const int qe = min(q + 4, hi);
#pragma unroll 1
                  for (; q < qe; ++q) one(q);
                }
              };
              seg(b, back_end);
              const int cmid = cc;
              seg(m, e);
              cf += cc - cmid;
              if (same)
                c1 = cc;
              else
                c0 = cc;
            }
            cq0 = cq1 + 1;
          }
        }
        ca = c0 + c1;
      } else
#endif
      for (int jy = max(cy - 1, 0); jy <= min(cy + 1, ndy - 1); ++jy) {
        const RowRange rr = row_range(ndx, cx, jy);
        // forward partners (creation order): later row, or same row from a threshold index on (per species)
        int fthr[3];
#pragma unroll
        for (int sq = 0; sq < 3; ++sq) {
          if (jy > cy)
            fthr[sq] = S.start[sq][rr.ca];
          else if (jy < cy)
            fthr[sq] = S.start[sq][rr.cb + 1];
          else
            fthr[sq] = (sq == sp) ? k + 1 : (sq > sp ? S.start[sq][c] : S.start[sq][c + 1]);
        }
        const int nd_row = S.start[2][rr.cb + 1] - S.start[2][rr.ca];
        if (nd_row == 0) {
          // no wall particles in this row: every list takes partners of a single species, whose order
          // (cell id, particle index) is the order of the merged range
          scan(0, S.start[0][rr.ca], S.start[0][rr.cb + 1], fthr[0]);
          scan(1, S.start[1][rr.ca], S.start[1][rr.cb + 1], fthr[1]);
        } else {
          // wall particles interleave with the other species cell by cell: (cell id, species, index) order
          for (int cq = rr.ca; cq <= rr.cb; ++cq) {
            scan(0, S.start[0][cq], S.start[0][cq + 1], fthr[0]);
            scan(1, S.start[1][cq], S.start[1][cq + 1], fthr[1]);
            scan(2, S.start[2][cq], S.start[2][cq + 1], fthr[2]);
          }
        }
      }
      if (c0 > CAND_CAP || c1 > CAND_CAP) *overflow = 1;
      // creation index / statistics count every pair once: at the owner of its earlier member
      const bool owned = !lflag || lflag[(sp == 0 ? S.order[0] : (sp == 1 ? S.order[1] : S.order[2]))[k]] == 1;
      nfwd_u[unified_slot(S, c, sp, k)] = owned ? cf : 0;
      if (!owned) ca = -1;
    } else if (lflag && lflag[(sp == 0 ? S.order[0] : (sp == 1 ? S.order[1] : S.order[2]))[k]] != 1) {
      ca = -1;
    }
    nall[t] = ca;
  }
  if (t < M.nnp + M.nsp) {  // list-owning slots
    n0[t] = live ? c0 : 0;
    n1[t] = live ? c1 : 0;
    int m0 = c0, m1 = c1;
    for (int o = 16; o > 0; o >>= 1) {
      m0 = max(m0, __shfl_xor_sync(0xffffffffu, m0, o));
      m1 = max(m1, __shfl_xor_sync(0xffffffffu, m1, o));
    }
    if ((threadIdx.x & 31) == 0) {
      const int sl = t / SLICE;
      w0[sl] = m0 * SLICE;
      const bool is_node = t < M.nnp;
      wC[sl] = is_node ? m1 * SLICE : 0;
      wD[sl] = is_node ? 0 : m1 * SLICE;
    }
  }
}

// after the scans: totals -> status block
__global__ void k_status(const GridInfo *__restrict__ G, const long long *__restrict__ totals,
                         const int *__restrict__ start, int cell_stride, const int *__restrict__ nout,
                         const int *__restrict__ cand_overflow, StepStatus *st) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  st->tot0 = totals[0];
  st->totC = totals[1];
  st->totD = totals[2];
  st->n_pairs = totals[3];
  st->ncell = G->ncell;
  st->overflow = G->overflow;
  st->err = 0;
  st->pad[0] = *cand_overflow;
  for (int sp = 0; sp < 3; ++sp) st->nloc[sp] = start[sp * cell_stride + G->ncell] + nout[sp];
}

// key of the want-th (1-based) forward partner, in creation order, of the particle at (cell c, species sp,
// species-sorted index k)
__device__ okey_t mth_forward_partner(const DevParams &P, const GridInfo *G, const SortArrays &S, int c, int sp, int k,
                                      int want) {
  const double2 pp = S.pos[sp][k];
  const double hp = S.h[sp][k];
  const int ndx = G->ndivx[0], ndy = G->ndivx[1];
  const int cy = c / ndx, cx = c - cy * ndx;
  int seen = 0;
  okey_t kb = 0;
  for (int jy = cy; jy <= min(cy + 1, ndy - 1) && seen < want; ++jy)
    for (int jx = max(cx - 1, 0); jx <= min(cx + 1, ndx - 1) && seen < want; ++jx) {
      const int cq = jy * ndx + jx;
      if (cq < c) continue;
      for (int sq = 0; sq < 3 && seen < want; ++sq) {
        const int b = S.start[sq][cq], e = S.start[sq][cq + 1];
        for (int q = b; q < e && seen < want; ++q) {
          const bool fwd = (cq > c) || (sq > sp || (sq == sp && q > k));
          if (!fwd) continue;
          double dx, dy, r, mh;
          if (!pair_accept(P, pp, hp, S.pos[sq][q], S.h[sq][q], dx, dy, r, mh)) continue;
          ++seen;
          kb = make_key(cq, sq, S.order[sq][q]);
        }
      }
    }
  return kb;
}

// Finds the split point of the growth rule: the pair with creation index M (1-based) -> (ka*, kb*).
__global__ void k_growth_threshold(DevParams P, SlotMap Mm, const GridInfo *__restrict__ G, SortArrays S,
                                   const int *__restrict__ base_u /* exclusive scan of nfwd_u */, long long Mold,
                                   GrowthRule *__restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int nact = S.start[0][G->ncell] + S.start[1][G->ncell] + S.start[2][G->ncell];
  // largest unified slot ua with base_u[ua] < Mold  (creation indices of ua are base+1 .. base+nfwd)
  int lo = 0, hi = nact - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if ((long long)base_u[mid] < Mold)
      lo = mid;
    else
      hi = mid - 1;
  }
  const int ua = lo;
  const int want = (int)(Mold - (long long)base_u[ua]);  // the want-th forward partner of ua is the last old pair
  // locate ua's cell: largest c with U(c) <= ua
  int cl = 0, ch = G->ncell - 1;
  while (cl < ch) {
    const int mid = (cl + ch + 1) >> 1;
    const int U = S.start[0][mid] + S.start[1][mid] + S.start[2][mid];
    if (U <= ua)
      cl = mid;
    else
      ch = mid - 1;
  }
  const int c = cl;
  int rem = ua - (S.start[0][c] + S.start[1][c] + S.start[2][c]);
  int sp = 0;
  for (; sp < 3; ++sp) {
    const int n = S.start[sp][c + 1] - S.start[sp][c];
    if (rem < n) break;
    rem -= n;
  }
  const int k = S.start[sp][c] + rem;
  out->mode = 2;
  out->ka = make_key(c, sp, S.order[sp][k]);
  out->kb = mth_forward_partner(P, G, S, c, sp, k, want);
}

struct ListPtrs {
  // list 0: node <- stress/dummy partners and stress <- node/dummy partners (types 1, 6, 9), reference
  //         orientation of the gradient (pair_i - pair_j after Pint_Update)
  int *idx0;
  float *w0, *gx0, *gy0;
  // (mass/rho)_partner * w as fp64 (low / high words), the factor h1/h2 of stress_point_update (main:430-431):
  // streaming it removes the per-entry gather of the partner's mass/rho (the pair-sum kernels are bound by the
  // L1 rate of gathered sectors, not by the streamed bytes)
  int *h0lo, *h0hi;
  // list C: node <- node (type 3), own-perspective gradient; list D: stress <- stress (type 2), weight only
  int *idxC;
  float *wC, *gxC, *gyC;
  // fp32 geometry of artificial_viscosity (main:856-864), frozen during a step: xij, yij, h = 0.5*(h_i+h_j)
  float *xC, *yC, *hC;
  int *idxD;
  float *wD;
  const int *off0, *offC, *offD;  // per slice, exclusive scans of the slice widths
};

// Fill pass: writes every list entry at its traversal position.
// Accepted candidates are first compacted into per-thread shared-memory queues (cheap, divergent scan), then
// the kernel evaluation + stores run as dense loops with (nearly) all lanes active and row-aligned stores.
constexpr int QCAP = 40;
constexpr int FILL_THREADS = 128;

__global__ void __launch_bounds__(FILL_THREADS)
k_fill_scan(DevParams P, SlotMap M, const GridInfo *__restrict__ G, SortArrays S, const int *__restrict__ n0,
       const int *__restrict__ n1, const GrowthRule *__restrict__ growth, ListPtrs L, int *__restrict__ bc_int,
       float *__restrict__ n_int, const double *__restrict__ mor, const unsigned char *__restrict__ mcls, int t0,
       int tn) {
  __shared__ int q0buf[QCAP][FILL_THREADS];  // cross-species partners (species in the top 2 bits)
  __shared__ int q1buf[QCAP][FILL_THREADS];  // same-species partners
  const int tid = threadIdx.x;
  const int t = t0 + blockIdx.x * blockDim.x + tid;
  if (t >= t0 + tn) return;
  int sp, k;
  if (!slot_decode(M, t, sp, k)) return;
  const int *__restrict__ orderp = sp == 0 ? S.order[0] : S.order[1];
  const int *__restrict__ cellp = sp == 0 ? S.cell[0] : S.cell[1];
  const double2 *__restrict__ posp = sp == 0 ? S.pos[0] : S.pos[1];
  const double *__restrict__ hpp = sp == 0 ? S.h[0] : S.h[1];
  const int id = orderp[k];
  const int c = cellp[k];
  const int cnt0 = n0[t], cnt1 = n1[t];
  if (sp == SP_NODE && P.track_nint) n_int[id] = (float)cnt1;  // node-node interaction count (main:870-871 / 221-222)
  if (c < 0) {
    if (sp == SP_NODE) bc_int[id] = 0;
    return;
  }
  const GrowthRule gr = *growth;
  const int lane = t & 31, sl = t / SLICE;
  const size_t o0 = (size_t)L.off0[sl] + lane;
  const size_t o1 = (size_t)(sp == SP_NODE ? L.offC[sl] : L.offD[sl]) + lane;
  const double2 pp = posp[k];
  const double hp = hpp[k];
  const bool uni = G->uniform_h != 0;
  const double sk = (double)P.scale_k;
  const KernelConsts K = kernel_consts(P, hp);
  const Prefilter pf = prefilter_bounds(P, G, hp);
  const float2 up = (sp == 0 ? S.upos[0] : S.upos[1])[k];
  const int ndx = G->ndivx[0], ndy = G->ndivx[1];
  const int cy = c / ndx, cx = c - cy * ndx;
  // number of "old" entries per list (prefix of the ascending order); only the split mode needs a pre-count
  int s0 = (gr.mode == 1) ? 0 : cnt0, s1 = (gr.mode == 1) ? 0 : cnt1;
  if (gr.mode == 2) {
    const okey_t up = make_key(c, sp, id);
    if (up >= gr.ka) {
      s0 = 0;
      s1 = 0;
      // partners at or before ka live in cells <= cell(ka); cheap conservative test on the first stencil cell
      const int cfirst = max(cy - 1, 0) * ndx + max(cx - 1, 0);
      if (cfirst <= key_cell(gr.ka)) {
        for (int jy = max(cy - 1, 0); jy <= min(cy + 1, ndy - 1); ++jy)
          for (int jx = max(cx - 1, 0); jx <= min(cx + 1, ndx - 1); ++jx) {
            const int cq = jy * ndx + jx;
            for (int sq = 0; sq < 3; ++sq) {
              const int b = S.start[sq][cq], e = S.start[sq][cq + 1];
              for (int q = b; q < e; ++q) {
                if (sq == sp && q == k) continue;
                double dx, dy, r, mh;
                if (!pair_accept(P, pp, hp, S.pos[sq][q], S.h[sq][q], dx, dy, r, mh)) continue;
                if (!pair_is_old(gr, up, make_key(cq, sq, S.order[sq][q]))) continue;
                if (sq == sp)
                  ++s1;
                else
                  ++s0;
              }
            }
          }
      }
    }
  }
  int e0 = 0, e1 = 0, has_dummy = 0, nq0 = 0, nq1 = 0;

  auto drain = [&]() {
    // list 0: cross-species partners, reference orientation of the gradient (pair_i - pair_j)
    for (int j = 0; j < nq0; ++j) {
      const int pk = q0buf[j][tid];
      const int sq = (int)((unsigned)pk >> 30), q = pk & 0x3fffffff;
      const double2 pq = S.pos[sq][q];
      const double hq = uni ? hp : S.h[sq][q];
      double dx = pp.x - pq.x, dy = pp.y - pq.y;
      double d2 = dx * dx;
      d2 = d2 + dy * dy;
      const double mh = (hp + hq) / 2.;
      const double r = sqrt(d2);
      // Pint_Update orientation: pair_i = stress particle (type 1) or dummy (types 6, 9)
      const bool p_is_i = (sp == SP_STRESS && sq == SP_NODE);
      if (!p_is_i) {
        dx = -dx;
        dy = -dy;
      }
      double w, gx, gy;
      if (P.skf == 1 && mh == K.h)
        sph_kernel_fast<true>(K, r, dx, dy, w, gx, gy);
      else
        sph_kernel(P, r, dx, dy, mh, w, gx, gy);
      const int pos = (e0 < s0) ? (cnt0 - s0) + e0 : (cnt0 - 1 - e0);
      const size_t a = o0 + (size_t)pos * SLICE;
      const int qid = S.order[sq][q];
      const float wf = (float)w;
      const double h0 = (sq == SP_DUMMY || mcls) ? 0.0 : mor[qid] * (double)wf;
      L.idx0[a] = mcls ? (qid | ((int)mcls[qid] << 30)) : qid;  // mass/rho class of the partner, see ell_stream
      if (L.h0lo) {  // not stored when mass/rho is uniform per species (the sweeps rebuild it from w)
        L.h0lo[a] = __double2loint(h0);
        L.h0hi[a] = __double2hiint(h0);
      }
      L.w0[a] = wf;
      L.gx0[a] = (float)gx;
      L.gy0[a] = (float)gy;
      if (sq == SP_DUMMY) has_dummy = 1;
      ++e0;
    }
    // list C / D: same-species partners, own-perspective gradient (nodes) or weight only (stress particles)
    for (int j = 0; j < nq1; ++j) {
      const int q = q1buf[j][tid];
      const double2 pq = posp[q];
      const double hq = uni ? hp : hpp[q];
      const double dx = pp.x - pq.x, dy = pp.y - pq.y;
      double d2 = dx * dx;
      d2 = d2 + dy * dy;
      const double mh = (hp + hq) / 2.;
      const double r = sqrt(d2);
      const int pos = (e1 < s1) ? (cnt1 - s1) + e1 : (cnt1 - 1 - e1);
      const size_t a = o1 + (size_t)pos * SLICE;
      double w, gx, gy;
      if (sp == SP_NODE) {
        if (P.skf == 1 && mh == K.h)
          sph_kernel_fast<true>(K, r, dx, dy, w, gx, gy);
        else
          sph_kernel(P, r, dx, dy, mh, w, gx, gy);
        L.idxC[a] = orderp[q];
        L.wC[a] = (float)w;
        L.gxC[a] = (float)gx;
        L.gyC[a] = (float)gy;
        L.xC[a] = (float)dx;                 // xij = real(x(1,i) - x(1,j)), main:856
        L.yC[a] = (float)dy;
        if (L.hC) L.hC[a] = (float)(0.5 * (hp + hq));  // main:863 (not stored when h is uniform)
      } else {
        if (P.skf == 1 && mh == K.h)
          sph_kernel_fast<false>(K, r, dx, dy, w, gx, gy);
        else
          sph_kernel(P, r, dx, dy, mh, w, gx, gy);
        L.idxD[a] = orderp[q];
        L.wD[a] = (float)w;
      }
      ++e1;
    }
    nq0 = 0;
    nq1 = 0;
  };

  auto scan = [&](int sq, int b, int e) {
    for (int q = b; q < e; ++q) {
      if (sq == sp && q == k) continue;
      int cls = 2;
      if (pf.on) cls = prefilter_test(pf, up, S.upos[sq][q]);
      if (cls == 0) continue;
      if (cls == 2) {
        double dx, dy, d2, mh;
        if (!pair_accept_fast(sk, pp, hp, S.pos[sq][q], uni ? hp : S.h[sq][q], dx, dy, d2, mh)) continue;
      }
      if (sq == sp) {
        q1buf[nq1++][tid] = q;
      } else {
        q0buf[nq0++][tid] = (sq << 30) | q;
      }
      if (nq0 == QCAP || nq1 == QCAP) drain();
    }
  };

  for (int jy = max(cy - 1, 0); jy <= min(cy + 1, ndy - 1); ++jy) {
    const RowRange rr = row_range(ndx, cx, jy);
    const int nd_row = S.start[2][rr.cb + 1] - S.start[2][rr.ca];
    if (nd_row == 0) {
      // no wall particles in this row: every list takes partners of a single species, whose order
      // (cell id, particle index) is the order of the merged range
      scan(0, S.start[0][rr.ca], S.start[0][rr.cb + 1]);
      scan(1, S.start[1][rr.ca], S.start[1][rr.cb + 1]);
    } else {
      // wall particles interleave with the other species cell by cell: (cell id, species, index) order
      for (int cq = rr.ca; cq <= rr.cb; ++cq)
        for (int sq = 0; sq < 3; ++sq) scan(sq, S.start[sq][cq], S.start[sq][cq + 1]);
    }
  }
  drain();
  if (sp == SP_NODE) bc_int[id] = has_dummy;  // main:506,579
}

// Fill pass, fast path: the accepted partners were recorded by k_count (cand0 / cand1, list order), so every
// thread evaluates the kernel for its entries in a dense loop, two entries in flight at 8 blocks per SM (measured best).
#ifndef SPSPH_FILL_MINB
#define SPSPH_FILL_MINB 7
#endif
#ifndef SPSPH_FILL_U
#define SPSPH_FILL_U 2
#endif
// UNIFORM: cubic spline and one smoothing length for all particles (known to the host after upload): only the
// hoisted-constant kernel evaluation is compiled in, which keeps the general kernels' registers out of the hot path
template <bool UNIFORM>
__global__ void __launch_bounds__(128, SPSPH_FILL_MINB)
k_fill(DevParams P, SlotMap M, const GridInfo *__restrict__ G, SortArrays S, const int *__restrict__ n0,
       const int *__restrict__ n1, const GrowthRule *__restrict__ growth, ListPtrs L, int *__restrict__ bc_int,
       float *__restrict__ n_int, const double *__restrict__ mor, const unsigned char *__restrict__ mcls,
       const int *__restrict__ cand0, const int *__restrict__ cand1, int t0, int tn) {
  const int t = t0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= t0 + tn) return;
  int sp, k;
  if (!slot_decode(M, t, sp, k)) return;
  const int *__restrict__ orderp = sp == 0 ? S.order[0] : S.order[1];
  const int *__restrict__ cellp = sp == 0 ? S.cell[0] : S.cell[1];
  const double2 *__restrict__ posp = sp == 0 ? S.pos[0] : S.pos[1];
  const double *__restrict__ hpp = sp == 0 ? S.h[0] : S.h[1];
  const int id = orderp[k];
  const int c = cellp[k];
  const int cnt0 = n0[t], cnt1 = n1[t];
  if (sp == SP_NODE && P.track_nint) n_int[id] = (float)cnt1;  // node-node interaction count (main:870-871 / 221-222)
  if (c < 0) {
    if (sp == SP_NODE) bc_int[id] = 0;
    return;
  }
  const GrowthRule gr = *growth;
  const int lane = t & 31, sl = t / SLICE;
  // list offsets fit 32 bits (the host refuses lists of 2^31 entries): one IMAD.WIDE per store address
  const int o0 = L.off0[sl] + lane;
  const int o1 = (sp == SP_NODE ? L.offC[sl] : L.offD[sl]) + lane;
  const int n2 = P.ntotal2;  // row stride of the unified sorted arrays: cross-species scratch entries are sq * n2 + q
  const size_t cb = cand_base(t);
  const double2 pp = posp[k];
  const double hp = hpp[k];
  const bool uni = UNIFORM || G->uniform_h != 0;
  const KernelConsts K = kernel_consts(P, hp);
  // number of "old" entries per list (prefix of the ascending order); only the split mode needs a pre-count
  int s0 = (gr.mode == 1) ? 0 : cnt0, s1 = (gr.mode == 1) ? 0 : cnt1;
  if (gr.mode == 2) {
    const okey_t up = make_key(c, sp, id);
    if (up >= gr.ka) {
      s0 = 0;
      s1 = 0;
      for (int e = 0; e < cnt0; ++e) {
        const int pk = cand0[cb + cand_off(e)];
        const int sq = pk >= 2 * n2 ? 2 : (pk >= n2 ? 1 : 0);
        s0 += pair_is_old(gr, up, sorted_key(S, sq, pk - sq * n2)) ? 1 : 0;
      }
      for (int e = 0; e < cnt1; ++e) s1 += pair_is_old(gr, up, sorted_key(S, sp, cand1[cb + cand_off(e)])) ? 1 : 0;
    }
  }
  constexpr int U = SPSPH_FILL_U;
  int has_dummy = 0;
  // list 0: cross-species partners, reference orientation of the gradient (pair_i - pair_j after Pint_Update)
  // the entries of the next trip are requested before this trip's arithmetic (the scratch comes from DRAM);
  // measured: also requesting the next trip's partner records one trip ahead costs registers and is slower
  // (1.34 ms at 5 blocks per SM / 94 registers, 1.46 at 7 / 72 with spills, against 1.27 ms as is)
  int pkn[U];
#pragma unroll
  for (int u = 0; u < U; ++u) pkn[u] = cnt0 > 0 ? cand0[cb + cand_off(min(u, cnt0 - 1))] : 0;
  for (int e0 = 0; e0 < cnt0; e0 += U) {
    int pk[U], idw[U];
    double2 pq[U];
    double hq[U];
#pragma unroll
    for (int u = 0; u < U; ++u) pk[u] = pkn[u];
    if (e0 + U < cnt0) {
#pragma unroll
      for (int u = 0; u < U; ++u) pkn[u] = cand0[cb + cand_off(min(e0 + U + u, cnt0 - 1))];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      pq[u] = S.pos[0][pk[u]];
      idw[u] = S.ordc[0][pk[u]];  // partner id, with its mass/rho class in the top bits when the palette is on
      hq[u] = uni ? hp : S.h[0][pk[u]];
    }
    double w[U], gx[U], gy[U], h0[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      double dx = pp.x - pq[u].x, dy = pp.y - pq[u].y;
      double d2 = dx * dx;
      d2 = d2 + dy * dy;
      const double mh = (hp + hq[u]) / 2.;
      const double r = sqrt(d2);
      // Pint_Update orientation: pair_i = stress particle (type 1) or dummy (types 6, 9)
      const bool p_is_i = (sp == SP_STRESS && pk[u] < n2);
      if (!p_is_i) {
        dx = -dx;
        dy = -dy;
      }
      if (UNIFORM || (P.skf == 1 && mh == K.h))
        sph_kernel_fast<true>(K, r, dx, dy, w[u], gx[u], gy[u]);
      else
        sph_kernel(P, r, dx, dy, mh, w[u], gx[u], gy[u]);
      h0[u] = (pk[u] >= 2 * n2 || mcls) ? 0.0 : mor[idw[u] & QID_MASK] * (double)(float)w[u];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + u;
      if (e >= cnt0) break;
      const int pos = (e < s0) ? (cnt0 - s0) + e : (cnt0 - 1 - e);
      const int a = o0 + pos * SLICE;
      L.idx0[a] = idw[u];  // with the mass/rho class of the partner, see ell_stream
      if (L.h0lo) {  // not stored when mass/rho is uniform per species (the sweeps rebuild it from w)
        L.h0lo[a] = __double2loint(h0[u]);
        L.h0hi[a] = __double2hiint(h0[u]);
      }
      L.w0[a] = (float)w[u];
      L.gx0[a] = (float)gx[u];
      L.gy0[a] = (float)gy[u];
      if (pk[u] >= 2 * n2) has_dummy = 1;
    }
  }
  // list C / D: same-species partners, own-perspective gradient (nodes) or weight only (stress particles)
#pragma unroll
  for (int u = 0; u < U; ++u) pkn[u] = cnt1 > 0 ? cand1[cb + cand_off(min(u, cnt1 - 1))] : 0;
  for (int e0 = 0; e0 < cnt1; e0 += U) {
    int q[U], qid[U];
    double2 pq[U];
    double hq[U];
#pragma unroll
    for (int u = 0; u < U; ++u) q[u] = pkn[u];
    if (e0 + U < cnt1) {
#pragma unroll
      for (int u = 0; u < U; ++u) pkn[u] = cand1[cb + cand_off(min(e0 + U + u, cnt1 - 1))];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      pq[u] = posp[q[u]];
      qid[u] = orderp[q[u]];
      hq[u] = uni ? hp : hpp[q[u]];
    }
    double w[U], gx[U], gy[U], dxs[U], dys[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double dx = pp.x - pq[u].x, dy = pp.y - pq[u].y;
      double d2 = dx * dx;
      d2 = d2 + dy * dy;
      const double mh = (hp + hq[u]) / 2.;
      const double r = sqrt(d2);
      dxs[u] = dx;
      dys[u] = dy;
      if (sp == SP_NODE) {
        if (UNIFORM || (P.skf == 1 && mh == K.h))
          sph_kernel_fast<true>(K, r, dx, dy, w[u], gx[u], gy[u]);
        else
          sph_kernel(P, r, dx, dy, mh, w[u], gx[u], gy[u]);
      } else {
        if (UNIFORM || (P.skf == 1 && mh == K.h))
          sph_kernel_fast<false>(K, r, dx, dy, w[u], gx[u], gy[u]);
        else
          sph_kernel(P, r, dx, dy, mh, w[u], gx[u], gy[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + u;
      if (e >= cnt1) break;
      const int pos = (e < s1) ? (cnt1 - s1) + e : (cnt1 - 1 - e);
      const int a = o1 + pos * SLICE;
      if (sp == SP_NODE) {
        L.idxC[a] = qid[u];
        L.wC[a] = (float)w[u];
        L.gxC[a] = (float)gx[u];
        L.gyC[a] = (float)gy[u];
        L.xC[a] = (float)dxs[u];                 // xij = real(x(1,i) - x(1,j)), main:856
        L.yC[a] = (float)dys[u];
        if (L.hC) L.hC[a] = (float)(0.5 * (hp + hq[u]));   // main:863 (not stored when h is uniform)
      } else {
        L.idxD[a] = qid[u];
        L.wD[a] = (float)w[u];
      }
    }
  }
  if (sp == SP_NODE) bc_int[id] = has_dummy;  // main:506,579
}

// Export of the reference's pair list (creation order + traversal position), for parity tests and tools.
__global__ void k_export_pairs(DevParams P, SlotMap M, const GridInfo *__restrict__ G, SortArrays S,
                               const int *__restrict__ base_u, long long n_pairs, long long Mold, int *__restrict__ pi,
                               int *__restrict__ pj, int *__restrict__ ptype, float *__restrict__ pw,
                               float *__restrict__ pgx, float *__restrict__ pgy) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  int sp, k;
  if (t >= M.total() || !slot_decode(M, t, sp, k)) return;
  const int c = S.cell[sp][k];
  if (c < 0) return;
  const double2 pp = S.pos[sp][k];
  const double hp = S.h[sp][k];
  const int ndx = G->ndivx[0], ndy = G->ndivx[1];
  const int cy = c / ndx, cx = c - cy * ndx;
  long long ci = base_u[unified_slot(S, c, sp, k)];  // 0-based creation index of this particle's first pair
  const int idp = S.order[sp][k];
  const int itp[3] = {2, 1, 25};
  for (int jy = cy; jy <= min(cy + 1, ndy - 1); ++jy)
    for (int jx = max(cx - 1, 0); jx <= min(cx + 1, ndx - 1); ++jx) {
      const int cq = jy * ndx + jx;
      if (cq < c) continue;
      for (int sq = 0; sq < 3; ++sq) {
        const int b = S.start[sq][cq], e = S.start[sq][cq + 1];
        for (int q = b; q < e; ++q) {
          const bool fwd = (cq > c) || (sq > sp || (sq == sp && q > k));
          if (!fwd) continue;
          double dx, dy, r, mh, w, gx, gy;
          if (!pair_accept(P, pp, hp, S.pos[sq][q], S.h[sq][q], dx, dy, r, mh)) continue;
          sph_kernel(P, r, dx, dy, mh, w, gx, gy);
          int i = idp + 1, j = S.order[sq][q] + 1;  // 1-based ids, creation orientation (itotal, jtotal)
          float fgx = (float)gx, fgy = (float)gy;
          const int ii = itp[sp], jj = itp[sq];
          if ((ii == 2 && jj == 1) || (ii == 2 && jj == 25) || (ii == 1 && jj == 25)) {  // Pint_Update swap
            const int tmp = i;
            i = j;
            j = tmp;
            fgx = -fgx;
            fgy = -fgy;
          }
          const int isumm = ii + jj;
          const int ty = isumm == 3 ? 1 : isumm == 2 ? 2 : isumm == 4 ? 3 : isumm == 27 ? 6 : isumm == 26 ? 9 : 0;
          // traversal position
          long long pos = ci;
          if (n_pairs > Mold) {
            const long long over = n_pairs - Mold;
            pos = (ci >= Mold) ? (n_pairs - 1 - ci) : (over + ci);
          }
          pi[pos] = i;
          pj[pos] = j;
          ptype[pos] = ty;
          pw[pos] = (float)w;
          pgx[pos] = fgx;
          pgy[pos] = fgy;
          ++ci;
        }
      }
    }
}

}  // namespace spsph
