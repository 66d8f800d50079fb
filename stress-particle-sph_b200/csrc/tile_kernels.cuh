// Cell-tile path of the time step: neighbour build and pair sums WITHOUT partner ids and without scattered gathers.
//
// Restricted to one particle, the reference's pair order is "partners sorted by (cell id, particle index)"
// (grid_find_NEW main:1322-1396), and a row of the 3x3 cell stencil is ONE contiguous range of the
// cell-sorted particle arrays. So instead of a partner id (round 1: 4 bytes + one scattered 32-byte gather per entry
// and sweep) a list entry is ONE BYTE: stencil row (2 bits) + position of the partner among that row's candidates
// (6 bits); the build derives it from per-row 64-bit acceptance masks. The records the partners expose (velocity of a velocity
// particle, stress of a stress particle, ...) are written in cell-sorted order by the kernel that precedes their
// reader, so a block of consecutive sorted particles stages the three candidate ranges it needs in shared memory
// with contiguous, coalesced loads and every lane then reads its partners from shared memory. What is still
// streamed per entry are the reference's fp32 pair weights (w: 4 bytes in the interpolation sweeps, dwdx/dwdy: 8
// bytes in the gradient sweeps; main:1376-1378), stored per lane in groups of four entries (one 16-byte load).
//
// Traversal order: ascending candidates (the reference's forward order) or, in the first step after an upload, fully
// reversed (list nodes are prepended when new, main:1362-1368). Steps in which the pair list grows later on (split
// order, a few per run) take the round-1 list path of step_kernels.cuh; so do option combinations this path does
// not cover (see tile_eligible() in spsph_engine.cu). Every per-particle sum runs in the reference's order in one
// thread: no atomics in any sum, results bit-identical to the list path.
//
// Reference rows: grid_find_NEW main:1199-1435 + kernel main:1440-1538 + Pint_Update mat:1574-1634 (k_tile_build),
// stress_point_update main:403-482 (k_tile_a_*), get_derivatives main:487-648 + RK4 main:653-802 + plastic_terms
// mat:1884-1954 (k_tile_b_*), artificial_viscosity main:826-904 (k_tile_av), XSPH_update main:189-239 and the
// position update main:140-182 (k_tile_move).
#pragma once
#include <type_traits>

#include "step_kernels.cuh"

namespace spsph {

typedef unsigned long long u64;

struct TileLists {
  // per list-owning thread slot t (velocity particles [0, nnp), stress particles [nnp, nnp + nsp)); stencil row
  // r = 0, 1, 2 (grid rows cy-1, cy, cy+1) at [r * nslots + t]
  int *rowA;     // first cross-species candidate of the row: start[partner species][row*ndx + max(cx-1, 0)]
  int *rowS;     // first same-species candidate of the row
  unsigned *mW;  // acceptance mask over the row's wall particles (read only by particles with wall partners)
  int *n0, *n1;  // entries of list 0 (cross-species + wall, bit 30: has wall partners) and of the same-species list
  // list entries in traversal order, groups of four consecutive entries per lane: entries 4g..4g+3 of slot t at
  // [(G(t/32) + g)*32 + t%32]. code byte: row << 6 | candidate position within the row; 0xC0: wall particle (the
  // wall partners are the set bits of mW in the same traversal order)
  unsigned *code0, *codeS;
  float4 *w0, *gx0, *gy0;  // list 0: w, dwdx, dwdy (fp32, main:1376-1378) in the reference's orientation after Pint_Update
  float4 *gxC, *gyC;       // velocity-velocity list: own-perspective gradient (artificial_viscosity)
  int capN0, capS0;        // rows per 32-slot slice (multiples of 4): list 0 of velocity / stress particles
  int capC, capD;          // ... same-species list of velocity / stress particles
  int nsl_n;               // slices of velocity particles (nnp / 32)
  int nslots;              // nnp + nsp
};
constexpr int TILE_WALL_FLAG = 1 << 30;
constexpr unsigned TILE_CODE_WALL = 0xC0u;

__device__ __forceinline__ size_t ell0_base(const TileLists &L, int t) {
  const int sl = t >> 5;
  const size_t g0 = sl < L.nsl_n ? (size_t)sl * (L.capN0 >> 2)
                                 : (size_t)L.nsl_n * (L.capN0 >> 2) + (size_t)(sl - L.nsl_n) * (L.capS0 >> 2);
  return g0 * 32 + (t & 31);
}
__device__ __forceinline__ size_t ellS_base(const TileLists &L, int t) {
  const int sl = t >> 5;
  const size_t g0 = sl < L.nsl_n ? (size_t)sl * (L.capC >> 2)
                                 : (size_t)L.nsl_n * (L.capC >> 2) + (size_t)(sl - L.nsl_n) * (L.capD >> 2);
  return g0 * 32 + (t & 31);
}
__device__ __forceinline__ int ell0_cap(const TileLists &L, int t) { return (t >> 5) < L.nsl_n ? L.capN0 : L.capS0; }
__device__ __forceinline__ int ellS_cap(const TileLists &L, int t) { return (t >> 5) < L.nsl_n ? L.capC : L.capD; }

// species-sorted copies of per-particle constants (written by k_rank_consts once per step): partners read them from a
// staged tile like the state records
struct SortedConsts {
  const double *mor[2];     // mass/rho
  const double2 *mrho[2];   // {mass, rho}
  const double *rrho[2];    // RN(1/rho)
};

__global__ void k_rank_consts(DevParams P, LocalList LL, const int *__restrict__ pos_of, const double *__restrict__ mass,
                              const double *__restrict__ rho, const double *__restrict__ mor, double *__restrict__ smor,
                              double2 *__restrict__ smrho, double *__restrict__ srrho) {
  SPSPH_FOR_LOCAL(LL, kk, i) {
    if (i >= P.ntotal) continue;
    const size_t a = (size_t)(i < P.nnode ? 0 : 1) * P.ntotal2 + pos_of[i];
    const double r = rho[i];
    smor[a] = mor[i];
    smrho[a] = make_double2(mass[i], r);
    srrho[a] = __drcp_rn(r);
  }
}

// ------------------------------------------------------------------------------------------------------
// Tile geometry of a block: its targets are consecutive species-sorted particles, so the candidates of all of them
// are, per stencil row, one contiguous range of the partner species' sorted arrays.
// ------------------------------------------------------------------------------------------------------
struct TileGeom {
  int base[3], off[3], cnt[3];  // per stencil row: first staged sorted index, offset in the staged array, count
  int staged;                   // 0: the block reads its partners from global memory (row-straddling or oversized tile)
};

// c0, c1: cells of the block's first and last live target; start_q: cell table of the partner species
__device__ __forceinline__ void tile_geom(const GridInfo *__restrict__ G, int c0, int c1, const int *__restrict__ start_q,
                                          int cap, TileGeom &g) {
  g.staged = 0;
  for (int r = 0; r < 3; ++r) g.base[r] = g.off[r] = g.cnt[r] = 0;
  if (c0 < 0 || c1 < 0) return;  // out-of-domain particles are parked behind the sorted ones: no partners
  const int ndx = G->ndivx[0], ndy = G->ndivx[1];
  const int cy0 = c0 / ndx, cy1 = c1 / ndx;
  if (cy0 != cy1) return;  // the block straddles two grid rows
  const int cxa = max(c0 - cy0 * ndx - 1, 0), cxb = min(c1 - cy0 * ndx + 1, ndx - 1);
  int tot = 0;
  for (int r = 0; r < 3; ++r) {
    const int row = cy0 - 1 + r;
    g.off[r] = tot;
    if (row < 0 || row >= ndy) continue;
    const int b = start_q[row * ndx + cxa], e = start_q[row * ndx + cxb + 1];
    g.base[r] = b;
    g.cnt[r] = e - b;
    tot += (e - b + 1) & ~1;  // even offsets: 16-byte alignment of 8-byte arrays
  }
  g.staged = tot <= cap ? 1 : 0;
}

// first candidate (sorted index) of each stencil row of a particle in cell c
__device__ __forceinline__ void lane_rows(const GridInfo *__restrict__ G, int c, const int *__restrict__ start_q, int (&b)[3]) {
  b[0] = b[1] = b[2] = 0;
  if (c < 0) return;
  const int ndx = G->ndivx[0], ndy = G->ndivx[1];
  const int cy = c / ndx, cx = c - cy * ndx;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int row = cy - 1 + r;
    if (row >= 0 && row < ndy) b[r] = start_q[row * ndx + max(cx - 1, 0)];
  }
}
// cooperative copy of the three candidate ranges into shared memory; idx != nullptr: the source is indexed by particle
// number (state between steps) and idx is the sorted order
template <class T>
__device__ __forceinline__ void stage_rows(T *__restrict__ sm, const T *__restrict__ g, const int *__restrict__ idx,
                                           const TileGeom &tg) {
  if (!tg.staged) return;
#pragma unroll
  for (int r = 0; r < 3; ++r)
    for (int i = threadIdx.x; i < tg.cnt[r]; i += blockDim.x) {
      const int j = tg.base[r] + i;
      sm[tg.off[r] + i] = g[idx ? idx[j] : j];
    }
}
// where a lane reads partner records: the staged tile, or global memory (sorted, or by particle number through idx)
template <class T>
struct Src {
  const T *p;
  const int *ix;
  __device__ __forceinline__ T operator()(int j) const { return p[ix ? ix[j] : j]; }
};
template <class T>
__device__ __forceinline__ Src<T> make_src(const T *sm, const T *g, const int *idx, const TileGeom &tg) {
  return tg.staged ? Src<T>{sm, nullptr} : Src<T>{g, idx};
}


// asynchronous copies global -> shared (LDGSTS): the staged tile does not pass through registers, and the copies
// overlap the loads of the block's own records
#ifndef SPSPH_HOST_EMU  // device only (the host emulation of tests/native/ copies synchronously)
__device__ __forceinline__ void cp_async_16(void *smem, const void *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_8(void *smem, const void *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_4(void *smem, const void *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
#else
inline void cp_async_4(void *smem, const void *gmem) { std::memcpy(smem, gmem, 4); }
inline void cp_async_16(void *smem, const void *gmem) { std::memcpy(smem, gmem, 16); }
inline void cp_async_8(void *smem, const void *gmem) { std::memcpy(smem, gmem, 8); }
inline void cp_async_wait_all() {}
#endif
// the three candidate ranges of a SORTED array, element size 8, 16 or 32 bytes
template <class T>
__device__ __forceinline__ void stage_rows_async(T *__restrict__ sm, const T *__restrict__ g, const TileGeom &tg) {
  static_assert(sizeof(T) == 8 || sizeof(T) == 16 || sizeof(T) == 32, "record sizes of the tile arrays");
  if (!tg.staged) return;
#pragma unroll
  for (int r = 0; r < 3; ++r)
    for (int i = threadIdx.x; i < tg.cnt[r]; i += blockDim.x) {
      const char *src = reinterpret_cast<const char *>(g + tg.base[r] + i);
      char *dst = reinterpret_cast<char *>(sm + tg.off[r] + i);
      if (sizeof(T) == 8) {
        cp_async_8(dst, src);
      } else {
        cp_async_16(dst, src);
        if (sizeof(T) == 32) cp_async_16(dst + 16, src + 16);
      }
    }
}

__device__ __forceinline__ float f4c(const float4 &v, int u) { return u == 0 ? v.x : (u == 1 ? v.y : (u == 2 ? v.z : v.w)); }
__device__ __forceinline__ float4 ldcs4(const float4 *p) { return __ldcs(p); }
__device__ __forceinline__ unsigned ldcs1(const unsigned *p) { return __ldcs(p); }
__device__ __forceinline__ int popc_range(u64 m, int lo, int hi) {  // set bits at positions [lo, hi)
  if (hi <= lo) return 0;
  const u64 hm = hi >= 64 ? ~0ull : ((1ull << hi) - 1ull);
  const u64 lm = lo >= 64 ? ~0ull : ((1ull << lo) - 1ull);
  return __popcll(m & hm & ~lm);
}

__device__ __forceinline__ u64 bits_range(int lo, int hi) {  // mask of bit positions [lo, hi)
  if (hi <= lo) return 0ull;
  const u64 hm = hi >= 64 ? ~0ull : ((1ull << hi) - 1ull);
  const u64 lm = lo >= 64 ? ~0ull : ((1ull << lo) - 1ull);
  return hm & ~lm;
}
// wall partners of a particle in traversal order: the set bits of its three wall masks (rare path: particles next
// to a wall); returns the index into the species-sorted wall arrays
struct WalkWall {
  unsigned m;
  int r;
  __device__ __forceinline__ void start(bool rev) {
    m = 0u;
    r = rev ? 3 : -1;
  }
  __device__ __forceinline__ int next(const TileLists &L, int t, bool rev, const int (&jW)[3]) {
    while (m == 0u) {
      r += rev ? -1 : 1;
      m = L.mW[(size_t)r * L.nslots + t];
    }
    int b;
    if (!rev) {
      b = __ffs((int)m) - 1;
      m &= m - 1;
    } else {
      b = 31 - __clz((int)m);
      m ^= 1u << b;
    }
    return (r == 0 ? jW[0] : (r == 1 ? jW[1] : jW[2])) + b;
  }
};

// ------------------------------------------------------------------------------------------------------
// Tile geometry of every block of the pair-sum kernels, computed once per step: a block then starts with ONE
// uniform load instead of a chain of dependent ones (cell of its first / last target -> cell table -> ranges).
//   kind 0: stress-particle targets, velocity-particle partners (sweeps A and B, stress side)
//   kind 1: velocity-particle targets, stress-particle partners (sweeps A and B, velocity side)
//   kind 2: velocity-particle targets and partners (XSPH)
//   kind 3: stress-particle targets and partners (XSPH)
//   kind 4: velocity-particle targets and partners, half-size blocks (artificial viscosity)
// ------------------------------------------------------------------------------------------------------
constexpr int TS_T = 128;  // targets per block: stress-particle side, artificial viscosity, position update
constexpr int TN_T = 64;   // velocity-particle side of sweeps A and B (a tile of stress particles is twice as large)
struct GeomTables {
  TileGeom *g[5];
};
__device__ __forceinline__ void geom_kind(int kind, int &tsp, int &qsp, int &T) {
  tsp = (kind == 0 || kind == 3) ? SP_STRESS : SP_NODE;
  qsp = (kind == 0 || kind == 2 || kind == 4) ? SP_NODE : SP_STRESS;
  T = (kind == 1 || kind == 4) ? TN_T : TS_T;
}
__global__ void k_tile_geoms(const GridInfo *__restrict__ G, SortArrays S, const int *__restrict__ nout, GeomTables GT,
                             int cap0, int cap1, int cap2, int cap3, int cap4) {
  const int kind = blockIdx.y;
  int tsp, qsp, T;
  geom_kind(kind, tsp, qsp, T);
  const int nlive = S.start[tsp][G->ncell] + nout[tsp];
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b * T >= nlive) return;
  const int cap = kind == 0 ? cap0 : (kind == 1 ? cap1 : (kind == 2 ? cap2 : (kind == 3 ? cap3 : cap4)));
  TileGeom g;
  tile_geom(G, S.cell[tsp][b * T], S.cell[tsp][min(b * T + T, nlive) - 1], S.start[qsp], cap, g);
  GT.g[kind][b] = g;
}

// ------------------------------------------------------------------------------------------------------
// Neighbour build, one pass: acceptance masks -> entry codes, list lengths, pair count, fp32 weights; also the two
// per-step sums that depend on geometry only -- cspm_norm of stress_point_update (main:430-433, 446-462) and the CSPM
// matrix of get_derivatives (main:539-550, 596-605).
// ------------------------------------------------------------------------------------------------------
constexpr int TB_T = 64;  // targets per block
constexpr int TB_CAP0 = 320, TB_CAP1 = 640, TB_CAP2 = 256;  // staged candidates per species
constexpr int TB_CODES = 64;  // entries per list the code scratch holds

struct TileStatus {  // read back by the host after the build
  long long n_pairs;
  int flags;  // bit 0: a stencil row / a list has more entries than the masks and codes hold; bit 1: a list exceeds its slice capacity
  int max_n0n, max_n0s, max_n1n, max_n1s;  // longest lists (size the slice capacities)
  int ncell, overflow;
  int nloc[3];
};

template <int SP>
__global__ void __launch_bounds__(TB_T)
k_tile_build(DevParams P, const GridInfo *__restrict__ G, SortArrays S, TileLists L, SortedConsts C, int nnp, int rev,
             int want_c, int want_s, const int *__restrict__ lflag, const int *__restrict__ nout, int *__restrict__ nall,
             int *__restrict__ bc_int, float *__restrict__ n_int, double *__restrict__ norm, double *__restrict__ AE,
             u64 *__restrict__ acc_pairs,
             int *__restrict__ flags /* [0] mask overflow, [1..4] longest lists, [5] slice overflow */) {
  constexpr bool LISTS = SP != SP_DUMMY;
  constexpr int SQA = SP == SP_NODE ? SP_STRESS : SP_NODE;  // cross-species partner of a list-owning target
  __shared__ float2 su0[LISTS ? TB_CAP0 : 1], su1[LISTS ? TB_CAP1 : 1], su2[LISTS ? TB_CAP2 : 1];
  __shared__ double2 sx0[LISTS ? TB_CAP0 : 1], sx1[LISTS ? TB_CAP1 : 1], sx2[LISTS ? TB_CAP2 : 1];
  // entry codes of the block's targets in creation order, one byte per entry: [entry][thread]
  __shared__ unsigned char scode0[LISTS ? TB_CODES * TB_T : 4], scodeS[LISTS ? TB_CODES * TB_T : 4];
  __shared__ TileGeom tgs[3];
  const int nlive = S.start[SP][G->ncell] + nout[SP];  // sorted particles of this species, incl. out-of-domain ones
  const int kb0 = blockIdx.x * TB_T;
  if (kb0 >= nlive) return;
  const int kb1 = min(kb0 + TB_T, nlive) - 1;
  if (threadIdx.x < 3) {
    const int caps[3] = {TB_CAP0, TB_CAP1, TB_CAP2};
    if (LISTS)
      tile_geom(G, S.cell[SP][kb0], S.cell[SP][kb1], S.start[threadIdx.x], caps[threadIdx.x], tgs[threadIdx.x]);
    else
      tgs[threadIdx.x].staged = 0;
  }
  const int k0 = kb0 + threadIdx.x;
  const bool live = k0 < nlive;
  const int k = live ? k0 : kb0;
  const int t = (SP == SP_NODE ? 0 : nnp) + k0;  // list slot (LISTS only)
  const int id = S.order[SP][k];
  const int c = live ? S.cell[SP][k] : -1;
  const double2 pp = S.pos[SP][k];
  const double hp = S.h[SP][k];
  const float2 up = S.upos[SP][k];
  __syncthreads();
  if (LISTS) {
    stage_rows_async(su0, S.upos[0], tgs[0]);
    stage_rows_async(sx0, S.pos[0], tgs[0]);
    stage_rows_async(su1, S.upos[1], tgs[1]);
    stage_rows_async(sx1, S.pos[1], tgs[1]);
    stage_rows_async(su2, S.upos[2], tgs[2]);
    stage_rows_async(sx2, S.pos[2], tgs[2]);
    cp_async_wait_all();
    __syncthreads();
  }
  const Prefilter pf = prefilter_bounds(P, G, hp);
  const double sk = (double)P.scale_k;
  // where the candidates are read: the staged tile or the global sorted arrays (generic pointers, set once)
  const float2 *U[3];
  const double2 *X[3];
#pragma unroll
  for (int sq = 0; sq < 3; ++sq) {
    const bool stg = LISTS && tgs[sq].staged;
    U[sq] = stg ? (sq == 0 ? su0 : (sq == 1 ? su1 : su2)) : S.upos[sq];
    X[sq] = stg ? (sq == 0 ? sx0 : (sq == 1 ? sx1 : sx2)) : S.pos[sq];
  }
  const bool rv = rev != 0;
  // same-species entries: velocity-velocity for artificial viscosity (want_c: with gradients) and XSPH, stress-stress
  // for XSPH only
  const bool want1 = LISTS && (SP == SP_NODE ? (want_c != 0 || want_s != 0) : want_s != 0);
  int ovf = 0;
  int cnt0 = 0, cnt1 = 0;  // entries of list 0 (cross-species + wall) and of the same-species list
  int cf = 0;              // forward partners (pairs this particle opens in creation order, main:1322-1341)
  unsigned mW[3] = {0u, 0u, 0u};
  int jbs[3][3], gbs[3][3];  // [species][row]: tile index / global sorted index of candidate 0
  // One candidate range. Accepted candidates get their code appended (creation order).
  //   fwd_from: candidates from this position on are forward partners; iself: position of the particle itself or -1
  auto scan = [&](int sq, int jb, int i0, int i1, unsigned rowtag, int fwd_from, int iself, unsigned char *sc, int &n,
                  bool store, unsigned *wmask) {
    const float2 *__restrict__ Uq = U[sq] + jb;
    for (int i = i0; i < i1; ++i) {
      const float2 uq = Uq[i];
      const float du = up.x - uq.x, dv = up.y - uq.y;
      const float d2 = __fmaf_rn(du, du, dv * dv);
      bool acc = pf.on && d2 < pf.lo;
      if (!acc && (!pf.on || d2 <= pf.hi)) {  // undecided by the fp32 prefilter: the reference's own test
        double dx, dy, dd, mh;
        acc = pair_accept_fast(sk, pp, hp, X[sq][jb + i], hp, dx, dy, dd, mh);
      }
      if (i == iself) acc = false;
      if (acc) {
        if (store && n < TB_CODES) sc[n * TB_T + threadIdx.x] = (unsigned char)(wmask ? TILE_CODE_WALL : (rowtag | (unsigned)i));
        if (wmask) *wmask |= 1u << i;
        ++n;
        cf += (i >= fwd_from) ? 1 : 0;
      }
    }
  };
  if (c >= 0) {
    const int ndx = G->ndivx[0], ndy = G->ndivx[1];
    const int cy = c / ndx, cx = c - cy * ndx;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int sq = 0; sq < 3; ++sq) jbs[sq][r] = gbs[sq][r] = 0;
      const int row = cy - 1 + r;
      if (row < 0 || row >= ndy) continue;
      const int ca_ = row * ndx + max(cx - 1, 0), cb_ = row * ndx + min(cx + 1, ndx - 1);
      int b[3], len[3], fw[3];
#pragma unroll
      for (int sq = 0; sq < 3; ++sq) {
        b[sq] = S.start[sq][ca_];
        len[sq] = S.start[sq][cb_ + 1] - b[sq];
        gbs[sq][r] = b[sq];
        jbs[sq][r] = (LISTS && tgs[sq].staged) ? b[sq] - tgs[sq].base[r] + tgs[sq].off[r] : b[sq];
        // creation order: later row, or own row from a threshold index on (own cell: species, then index)
        fw[sq] = r == 2 ? 0 : (r == 0 ? (1 << 30) : ((sq == SP ? k + 1 : (sq > SP ? S.start[sq][c] : S.start[sq][c + 1])) - b[sq]));
        if (len[sq] > ((sq == SP_DUMMY && LISTS) ? 32 : 64)) {
          ovf |= 1;
          len[sq] = 0;
        }
      }
      const int iself = r == 1 ? k - b[SP] : -1;
      if (!LISTS) {  // wall particle: counts only
        int n = 0;
        scan(0, jbs[0][r], 0, len[0], 0u, fw[0], SP == 0 ? iself : -1, nullptr, n, false, nullptr);
        scan(1, jbs[1][r], 0, len[1], 0u, fw[1], SP == 1 ? iself : -1, nullptr, n, false, nullptr);
        scan(2, jbs[2][r], 0, len[2], 0u, fw[2], SP == 2 ? iself : -1, nullptr, n, false, nullptr);
        cnt0 += n;
        continue;
      }
      const unsigned rowtag = (unsigned)r << 6;
      if (len[SP_DUMMY] == 0) {
        scan(SQA, jbs[SQA][r], 0, len[SQA], rowtag, fw[SQA], -1, scode0, cnt0, true, nullptr);
      } else {
        // wall particles in this row: (cell, species, index) order -- cell by cell, the cross-species partners of the
        // cell and then its wall particles
        for (int cq = ca_; cq <= cb_; ++cq) {
          scan(SQA, jbs[SQA][r], S.start[SQA][cq] - b[SQA], min(S.start[SQA][cq + 1] - b[SQA], len[SQA]), rowtag, fw[SQA],
               -1, scode0, cnt0, true, nullptr);
          scan(SP_DUMMY, jbs[SP_DUMMY][r], S.start[SP_DUMMY][cq] - b[SP_DUMMY],
               min(S.start[SP_DUMMY][cq + 1] - b[SP_DUMMY], len[SP_DUMMY]), rowtag, fw[SP_DUMMY], -1, scode0, cnt0, true,
               &mW[r]);
        }
      }
      scan(SP, jbs[SP][r], 0, len[SP], rowtag, fw[SP], iself, scodeS, cnt1, want1, nullptr);
    }
  } else {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int sq = 0; sq < 3; ++sq) jbs[sq][r] = gbs[sq][r] = 0;
  }
  // statistics: every pair is counted once, at the owner of its earlier member
  const bool owned = !lflag || lflag[id] == 1;
  {
    u64 f = (live && c >= 0 && owned) ? (u64)cf : 0ull;
    for (int o = 16; o > 0; o >>= 1) f += __shfl_xor_sync(0xffffffffu, f, o);
    if ((threadIdx.x & 31) == 0 && f) atomicAdd(acc_pairs, f);
    if (live) nall[(SP == SP_NODE ? 0 : (SP == SP_STRESS ? nnp : L.nslots)) + k0] = owned ? (c >= 0 ? cnt0 + cnt1 : 0) : -1;
  }
  if (!LISTS) {
    if (__any_sync(0xffffffffu, ovf != 0) && (threadIdx.x & 31) == 0) atomicOr(&flags[0], 1);
    return;
  }
  const bool wallp = (mW[0] | mW[1] | mW[2]) != 0u;
  if (cnt0 > TB_CODES || (want1 && cnt1 > TB_CODES)) ovf |= 1;
  const int cap0 = ell0_cap(L, t), cap1 = ellS_cap(L, t);
  if (cnt0 > cap0 || (want1 && cnt1 > cap1)) ovf |= 2;
  if (live) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const size_t a = (size_t)r * L.nslots + t;
      L.rowA[a] = gbs[SQA][r];
      L.rowS[a] = gbs[SP][r];
      if (wallp) L.mW[a] = mW[r];
    }
    L.n0[t] = cnt0 | (wallp ? TILE_WALL_FLAG : 0);
    L.n1[t] = cnt1;
    if (SP == SP_NODE) {
      bc_int[id] = wallp ? 1 : 0;                  // main:506,579
      if (P.track_nint) n_int[id] = (float)cnt1;   // velocity-velocity interaction count (main:870-871 / 221-222)
    }
  }
  {
    int m0 = live ? cnt0 : 0, m1 = live ? cnt1 : 0;
    m0 = warp_max_i(m0);
    m1 = warp_max_i(m1);
    const int o = warp_max_i(ovf);
    if ((threadIdx.x & 31) == 0) {
      if (o & 1) atomicOr(&flags[0], 1);
      if (o & 2) atomicOr(&flags[5], 1);
      atomicMax(&flags[SP == SP_NODE ? 1 : 2], m0);
      atomicMax(&flags[SP == SP_NODE ? 3 : 4], m1);
    }
  }
  const bool ok_w = live && ovf == 0;
  // the code word of entries 4g..4g+3 in TRAVERSAL order: creation order, or its mirror image in the first step
  auto code_word = [&](const unsigned char *sc, int g, int cnt) {
    unsigned cw = 0u;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = g * 4 + u;
      if (e < cnt) cw |= (unsigned)sc[(rv ? cnt - 1 - e : e) * TB_T + threadIdx.x] << (8 * u);
    }
    return cw;
  };
  // ---- weights, in traversal order; entries are evaluated in groups of four (independent division / sqrt chains) ----
  const KernelConsts K = kernel_consts(P, hp);
  const double2 *__restrict__ mrq = C.mrho[SQA];
  double nrm = 0.0, ae1 = 0.0, ae2 = 0.0, ae3 = 0.0, ae4 = 0.0;
  {
    const int n0w = ok_w ? cnt0 : 0;
    WalkWall ww;
    ww.start(rv);
    const int jW[3] = {jbs[SP_DUMMY][0], jbs[SP_DUMMY][1], jbs[SP_DUMMY][2]};
    const size_t base = ell0_base(L, t);
    for (int g = 0; g * 4 < n0w; ++g) {
      const unsigned cw = code_word(scode0, g, n0w);
      float wv[4] = {0.f, 0.f, 0.f, 0.f}, gxv[4] = {0.f, 0.f, 0.f, 0.f}, gyv[4] = {0.f, 0.f, 0.f, 0.f};
      double2 pq[4];
      int jg[4];
      bool isw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        isw[u] = false;
        jg[u] = 0;
        pq[u] = pp;
        if (g * 4 + u < n0w) {
          const unsigned cd = (cw >> (8 * u)) & 0xffu;
          const int r = (int)(cd >> 6), b = (int)(cd & 63u);
          if (r == 3) {
            isw[u] = true;
            pq[u] = X[SP_DUMMY][ww.next(L, t, rv, jW)];
          } else {
            pq[u] = X[SQA][(r == 0 ? jbs[SQA][0] : (r == 1 ? jbs[SQA][1] : jbs[SQA][2])) + b];
            jg[u] = (r == 0 ? gbs[SQA][0] : (r == 1 ? gbs[SQA][1] : gbs[SQA][2])) + b;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (g * 4 + u >= n0w) continue;
        double dx = pp.x - pq[u].x, dy = pp.y - pq[u].y;
        double d2 = dx * dx;
        d2 = d2 + dy * dy;
        const double r = sqrt(d2);
        // Pint_Update orientation (mat:1574-1634): pair_i = stress particle (type 1) or wall particle (types 6, 9)
        const bool p_is_i = (SP == SP_STRESS && !isw[u]);
        if (!p_is_i) {
          dx = -dx;
          dy = -dy;
        }
        double w, gx, gy;
        sph_kernel_fast<true>(K, r, dx, dy, w, gx, gy);
        wv[u] = (float)w;
        gxv[u] = (float)gx;
        gyv[u] = (float)gy;
        if (!isw[u]) {
          const double2 mr = mrq[jg[u]];
          const double rr = __drcp_rn(mr.y);
          nrm = nrm + div_rn((double)wv[u] * mr.x, mr.y, rr);  // cspm_norm, main:433
          if (P.cspm) {  // CSPM matrix, main:539-550
            const double gxd = (double)gxv[u], gyd = (double)gyv[u];
            const double h1 = SP == SP_STRESS ? div_rn(gxd * mr.x, mr.y, rr) : div_rn(-gxd * mr.x, mr.y, rr);
            const double h2 = SP == SP_STRESS ? div_rn(gyd * mr.x, mr.y, rr) : div_rn(-gyd * mr.x, mr.y, rr);
            ae1 = ae1 + (pq[u].x - pp.x) * h1;
            ae2 = ae2 + (pq[u].y - pp.y) * h1;
            ae3 = ae3 + (pq[u].x - pp.x) * h2;
            ae4 = ae4 + (pq[u].y - pp.y) * h2;
          }
        }
      }
      const size_t a = base + (size_t)g * 32;
      L.code0[a] = cw;
      L.w0[a] = make_float4(wv[0], wv[1], wv[2], wv[3]);
      L.gx0[a] = make_float4(gxv[0], gxv[1], gxv[2], gxv[3]);
      L.gy0[a] = make_float4(gyv[0], gyv[1], gyv[2], gyv[3]);
    }
  }
  if (want1) {  // same-species list: codes; velocity particles also the gradient from their own perspective
    const int n1w = ok_w ? cnt1 : 0;
    const size_t base = ellS_base(L, t);
    for (int g = 0; g * 4 < n1w; ++g) {
      const unsigned cw = code_word(scodeS, g, n1w);
      if (SP == SP_NODE && want_c) {
        float gxv[4] = {0.f, 0.f, 0.f, 0.f}, gyv[4] = {0.f, 0.f, 0.f, 0.f};
        double2 pq[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          pq[u] = pp;
          if (g * 4 + u < n1w) {
            const unsigned cd = (cw >> (8 * u)) & 0xffu;
            const int r = (int)(cd >> 6), b = (int)(cd & 63u);
            pq[u] = X[SP][(r == 0 ? jbs[SP][0] : (r == 1 ? jbs[SP][1] : jbs[SP][2])) + b];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (g * 4 + u >= n1w) continue;
          const double dx = pp.x - pq[u].x, dy = pp.y - pq[u].y;
          double d2 = dx * dx;
          d2 = d2 + dy * dy;
          const double r = sqrt(d2);
          double w, gx, gy;
          sph_kernel_fast<true>(K, r, dx, dy, w, gx, gy);
          gxv[u] = (float)gx;
          gyv[u] = (float)gy;
        }
        const size_t a = base + (size_t)g * 32;
        L.gxC[a] = make_float4(gxv[0], gxv[1], gxv[2], gxv[3]);
        L.gyC[a] = make_float4(gyv[0], gyv[1], gyv[2], gyv[3]);
      }
      L.codeS[base + (size_t)g * 32] = cw;
    }
  }
  if (!live) return;
  norm[id] = nrm;
  if (P.cspm) {  // inversion of the CSPM matrix, main:596-605
    double ae5 = ae1 * ae4 - ae2 * ae3;
    if (fabs(ae5) < P.ae_thr) {
      ae5 = 1;
      ae1 = 1;
      ae2 = 0;
      ae3 = 0;
      ae4 = 1;
    } else {
      ae5 = 1 / ae5;
    }
    double *AEp = AE + 5 * (size_t)id;
    AEp[0] = ae1;
    AEp[1] = ae2;
    AEp[2] = ae3;
    AEp[3] = ae4;
    AEp[4] = ae5;
  }
}

__global__ void k_tile_status(const GridInfo *__restrict__ G, const int *__restrict__ start, int cell_stride,
                              const int *__restrict__ nout, const u64 *__restrict__ acc_pairs,
                              const int *__restrict__ flags, TileStatus *st) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  st->n_pairs = (long long)*acc_pairs;
  st->flags = (flags[0] ? 1 : 0) | (flags[5] ? 2 : 0);
  st->max_n0n = flags[1];
  st->max_n0s = flags[2];
  st->max_n1n = flags[3];
  st->max_n1s = flags[4];
  st->ncell = G->ncell;
  st->overflow = G->overflow;
  for (int sp = 0; sp < 3; ++sp) st->nloc[sp] = start[sp * cell_stride + G->ncell] + nout[sp];
}

// sorted partner-visible records of one step (written in species-sorted order by the kernel preceding their reader)
struct TileRecs {
  double2 *NAs;  // [nn] velocity of a velocity particle: input of sweep A (stress side)
  Rec4 *SAs;     // [ns] stress of a stress particle: input of sweep A (velocity side)
  Rec4 *NBs;     // [nn] {vx, vy, m, rho}: input of sweep B (stress side), artificial viscosity, XSPH
  Rec4 *SBs;     // [ns] {s1/rho^2, s2/rho^2, s3/rho^2, m}: input of sweep B (velocity side)
  double2 *SVs;  // [ns] velocity of a stress particle after the final interpolation: XSPH
};

// groups of four entries staged per lane: cross-species list of a stress / velocity particle, velocity-velocity list,
// same-species list in the position update (typical lengths in the Bui layout: 18, 36, 17, 35 entries)
#ifndef SPSPH_LG_SP0
#define SPSPH_LG_SP0 7
#endif
#ifndef SPSPH_LG_N0
#define SPSPH_LG_N0 0
#endif
#ifndef SPSPH_LG_NC
#define SPSPH_LG_NC 0
#endif
#ifndef SPSPH_LG_SS
#define SPSPH_LG_SS 11
#endif
constexpr int LG_SP0 = SPSPH_LG_SP0, LG_N0 = SPSPH_LG_N0, LG_NC = SPSPH_LG_NC, LG_SS = SPSPH_LG_SS;  // 0: not staged
#ifndef SPSPH_TILE_WARPS
#define SPSPH_TILE_WARPS 16  // resident warps per SM requested from ptxas
#endif
#define TILE_MINB(T_) (SPSPH_TILE_WARPS * 32 / (T_))

// per-lane state of a list walk: its length, the tile index of candidate 0 of each stencil row, code / weight cursor
struct LaneList {
  int cnt;
  bool wallp;
  int j0, j1, j2;
  size_t base;
};
// cross-species list (list 0) of slot t
__device__ __forceinline__ LaneList lane_list0(const TileLists &L, int t, bool live, const TileGeom &tg) {
  LaneList q;
  const int n0r = live ? L.n0[t] : 0;
  q.cnt = n0r & ~TILE_WALL_FLAG;
  q.wallp = (n0r & TILE_WALL_FLAG) != 0;
  q.j0 = q.j1 = q.j2 = 0;
  q.base = 0;
  if (q.cnt == 0) return q;
  q.j0 = L.rowA[t];
  q.j1 = L.rowA[(size_t)L.nslots + t];
  q.j2 = L.rowA[2 * (size_t)L.nslots + t];
  if (tg.staged) {
    q.j0 += tg.off[0] - tg.base[0];
    q.j1 += tg.off[1] - tg.base[1];
    q.j2 += tg.off[2] - tg.base[2];
  }
  q.base = ell0_base(L, t);
  return q;
}
// same-species list of slot t
__device__ __forceinline__ LaneList lane_listS(const TileLists &L, int t, bool live, const TileGeom &tg) {
  LaneList q;
  q.cnt = live ? L.n1[t] : 0;
  q.wallp = false;
  q.j0 = q.j1 = q.j2 = 0;
  q.base = 0;
  if (q.cnt == 0) return q;
  q.j0 = L.rowS[t];
  q.j1 = L.rowS[(size_t)L.nslots + t];
  q.j2 = L.rowS[2 * (size_t)L.nslots + t];
  if (tg.staged) {
    q.j0 += tg.off[0] - tg.base[0];
    q.j1 += tg.off[1] - tg.base[1];
    q.j2 += tg.off[2] - tg.base[2];
  }
  q.base = ellS_base(L, t);
  return q;
}
// tile index of the partner of an entry code
__device__ __forceinline__ int code_index(const LaneList &q, unsigned cd) {
  const unsigned r = cd >> 6;
  return (r == 0 ? q.j0 : (r == 1 ? q.j1 : q.j2)) + (int)(cd & 63u);
}

// A block's list entries staged in shared memory: every lane copies ITS OWN groups (codes + NA float4 weight arrays)
// with cp.async right at the start of the block, together with the partner tile, so that all global loads of the
// pair sum are in flight at once and the walk itself runs out of shared memory. LG groups are staged; a lane with a
// longer list reads the rest from global memory.
template <int NA, int T, int LG>
struct ListSmem {
  unsigned c[LG > 0 ? LG : 1][LG > 0 ? T : 1];
  float4 x[(NA > 0 && LG > 0) ? LG : 1][(NA > 0 && LG > 0) ? T : 1];
  float4 y[(NA > 1 && LG > 0) ? LG : 1][(NA > 1 && LG > 0) ? T : 1];
};
template <int NA, int T, int LG>
__device__ __forceinline__ void list_stage(ListSmem<NA, T, LG> &sm, const unsigned *__restrict__ code,
                                           const float4 *__restrict__ a0, const float4 *__restrict__ a1,
                                           const LaneList &q) {
  if (LG == 0) return;
  const int ng = min((q.cnt + 3) >> 2, LG);
  for (int g = 0; g < ng; ++g) {
    const size_t a = q.base + (size_t)g * 32;
    cp_async_4(&sm.c[g][threadIdx.x], code + a);
    if (NA > 0) cp_async_16(&sm.x[g][threadIdx.x], a0 + a);
    if (NA > 1) cp_async_16(&sm.y[g][threadIdx.x], a1 + a);
  }
}
// body(nowall, code, w0, w1) is called for every entry of this lane in traversal order. Full groups of a lane without
// wall partners run as straight-line code (nowall = std::true_type: no branch per entry), so the four entries'
// load / conversion / division chains interleave; ragged tails and lanes next to a wall take the guarded form.
// LG = 0: nothing is staged, the groups are read from global memory two groups ahead.
template <int NA, int T, int LG, class Body>
__device__ __forceinline__ void walk_list(const ListSmem<NA, T, LG> &sm, const unsigned *__restrict__ code,
                                          const float4 *__restrict__ a0, const float4 *__restrict__ a1,
                                          const LaneList &q, Body body) {
  const int ng = (q.cnt + 3) >> 2;
  if (ng == 0) return;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  unsigned c1 = 0u, c2 = 0u;
  float4 x1 = z4, x2 = z4, y1 = z4, y2 = z4;
  auto gload = [&](int g, unsigned &c, float4 &x, float4 &y) {
    const size_t a = q.base + (size_t)g * 32;
    c = ldcs1(code + a);
    if (NA > 0) x = ldcs4(a0 + a);
    if (NA > 1) y = ldcs4(a1 + a);
  };
  if (LG == 0) {
    gload(0, c1, x1, y1);
    if (ng > 1) gload(1, c2, x2, y2);
  }
  for (int g = 0; g < ng; ++g) {
    unsigned c;
    float4 x = z4, y = z4;
    if (LG == 0) {
      c = c1;
      x = x1;
      y = y1;
      c1 = c2;
      x1 = x2;
      y1 = y2;
      if (g + 2 < ng) gload(g + 2, c2, x2, y2);
    } else if (g < LG) {
      c = sm.c[g][threadIdx.x];
      if (NA > 0) x = sm.x[g][threadIdx.x];
      if (NA > 1) y = sm.y[g][threadIdx.x];
    } else {
      gload(g, c, x, y);
    }
    const int nv = q.cnt - g * 4;
    if (nv >= 4 && !q.wallp) {
#pragma unroll
      for (int u = 0; u < 4; ++u) body(std::true_type{}, (c >> (8 * u)) & 0xffu, f4c(x, u), f4c(y, u));
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (u < nv) body(std::false_type{}, (c >> (8 * u)) & 0xffu, f4c(x, u), f4c(y, u));
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// Sweep A, stress-particle side (stress_point_update main:403-482: velocity of a stress particle from its velocity
// particles; + adapt_stress2 / BCs that follow it). FROMB: input is the state between steps (SPH_shift, main:99-109).
// ------------------------------------------------------------------------------------------------------
constexpr int TA_SP_CAP = 384;
template <bool FROMB>
__global__ void __launch_bounds__(TS_T, TILE_MINB(TS_T))
k_tile_a_sp(DevParams P, SlotMap M, SortArrays S, TileLists L, SortedConsts C, TileRecs R, StatePtrs st,
            const TileGeom *__restrict__ geoms, int do_adapt, int do_bc, int final_sweep) {
  __shared__ double2 sv[TA_SP_CAP];
  __shared__ double smo[TA_SP_CAP];
  __shared__ ListSmem<1, TS_T, LG_SP0> sl;
  const int kb0 = blockIdx.x * TS_T;
  if (kb0 >= M.ns) return;
  const TileGeom tg = geoms[blockIdx.x];
  const int k0 = kb0 + threadIdx.x;
  const bool live = k0 < M.ns;
  const int k = live ? k0 : kb0;
  const int t = M.nnp + k0;
  const int id = S.order[1][k];
  const int ks = id - P.nnode;
  const LaneList q = lane_list0(L, t, live, tg);
  list_stage(sl, L.code0, L.w0, nullptr, q);
  if (tg.staged) {
    if (FROMB) {
      for (int r = 0; r < 3; ++r)
        for (int i = threadIdx.x; i < tg.cnt[r]; i += blockDim.x) {
          const Rec4 p = ldrec(st.NBr, S.order[0][tg.base[r] + i]);
          sv[tg.off[r] + i] = make_double2(p.a, p.b);
        }
    } else {
      stage_rows_async(sv, (const double2 *)R.NAs, tg);
    }
    stage_rows_async(smo, C.mor[0], tg);
  }
  double2 v;
  Stress4 s;
  if (FROMB) {
    const Rec4 r = ldrec(st.SVbr, ks);
    v = make_double2(r.a, r.b);
    s = ld4(st.SFbr, ks);
  } else {
    v = ld2(st.SVa, ks);
    const Rec4 r = R.SAs[k];
    s = Stress4{r.a, r.b, r.c, r.d};
  }
  const double nrm = st.norm[id];
  const double own_mor = st.mor[id], own_rho = st.rho[id], own_m = st.mass[id];
  cp_async_wait_all();
  __syncthreads();
  const bool staged = tg.staged != 0;
  const double2 *__restrict__ vsrc = staged ? sv : (const double2 *)R.NAs;
  const double *__restrict__ msrc = staged ? smo : C.mor[0];
  double vtx = 0.0, vty = 0.0;
  walk_list(sl, L.code0, L.w0, nullptr, q, [&](auto nowall, unsigned cd, float w, float) {
    if (!decltype(nowall)::value && cd >= TILE_CODE_WALL) return;  // wall partners (type 9) take no part
    const int j = code_index(q, cd);
    double2 vq;
    if (FROMB && !staged) {
      const Rec4 p = ldrec(st.NBr, S.order[0][j]);
      vq = make_double2(p.a, p.b);
    } else {
      vq = vsrc[j];
    }
    const double h2 = msrc[j] * (double)w;  // (mass(i)/rho(i))*w, main:431
    vtx = vtx + vq.x * h2;
    vty = vty + vq.y * h2;
  });
  if (!live) return;
  if (nrm != 0) {
    const double rn = __drcp_rn(nrm);
    v.x = div_rn(vtx, nrm, rn);
    v.y = div_rn(vty, nrm, rn);
  }
  if (do_adapt) adapt_stress(P, s);
  if (do_bc) apply_bcs(P, st.bc_or_not, st.bc_info, st.bc_int, st.fs_normal, id, v, s);
  strec(st.SVb, ks, v.x, v.y, own_mor, 0.0);
  st4(st.SFb, ks, s);
  const double r2 = own_rho * own_rho, rr2 = __drcp_rn(r2);
  R.SBs[k] = Rec4{div_rn(s.s1, r2, rr2), div_rn(s.s2, r2, rr2), div_rn(s.s3, r2, rr2), own_m};
  if (final_sweep) R.SVs[k] = v;
}

// ------------------------------------------------------------------------------------------------------
// Sweep A, velocity-particle side (stress and plastic strain of a velocity particle from its stress particles).
// ------------------------------------------------------------------------------------------------------
constexpr int TA_N_CAP = 608;
template <bool FROMB, bool EPSP>
__global__ void __launch_bounds__(TN_T, TILE_MINB(TN_T))
k_tile_a_node(DevParams P, SlotMap M, SortArrays S, TileLists L, SortedConsts C, TileRecs R, StatePtrs st,
              const TileGeom *__restrict__ geoms, int do_adapt, int do_bc, int final_sweep) {
  __shared__ Rec4 ss[TA_N_CAP];
  __shared__ double smo[TA_N_CAP];
  __shared__ double sep[EPSP ? TA_N_CAP : 1];
  __shared__ ListSmem<1, TN_T, LG_N0> sl;
  const int kb0 = blockIdx.x * TN_T;
  if (kb0 >= M.nn) return;
  const TileGeom tg = geoms[blockIdx.x];
  const int k0 = kb0 + threadIdx.x;
  const bool live = k0 < M.nn;
  const int k = live ? k0 : kb0;
  const int t = k0;
  const int id = S.order[0][k];
  const LaneList q = lane_list0(L, t, live, tg);
  list_stage(sl, L.code0, L.w0, nullptr, q);
  if (tg.staged) {
    if (FROMB) {
      for (int r = 0; r < 3; ++r)
        for (int i = threadIdx.x; i < tg.cnt[r]; i += blockDim.x)
          ss[tg.off[r] + i] = ld256(st.SFbr + 4 * (size_t)(S.order[1][tg.base[r] + i] - P.nnode));
    } else {
      stage_rows_async(ss, (const Rec4 *)R.SAs, tg);
    }
    stage_rows_async(smo, C.mor[1], tg);
    if (EPSP) stage_rows(sep, (const double *)st.epsp, S.order[1], tg);
  }
  double2 v;
  Stress4 s;
  if (FROMB) {
    const Rec4 r = ldrec(st.NBr, id);
    v = make_double2(r.a, r.b);
    s = ld4(st.NSbr, id);
  } else {
    v = R.NAs[k];
    s = ld4(st.NSa, id);
  }
  const double nrm = st.norm[id];
  const double2 mr = C.mrho[0][k];
  cp_async_wait_all();
  __syncthreads();
  const bool staged = tg.staged != 0;
  const Rec4 *__restrict__ ssrc = staged ? ss : (const Rec4 *)R.SAs;
  const double *__restrict__ msrc = staged ? smo : C.mor[1];
  double t1 = 0.0, t2 = 0.0, t3 = 0.0, t4 = 0.0, te = 0.0;
  walk_list(sl, L.code0, L.w0, nullptr, q, [&](auto nowall, unsigned cd, float w, float) {
    if (!decltype(nowall)::value && cd >= TILE_CODE_WALL) return;  // wall partners (type 6) take no part
    const int j = code_index(q, cd);
    Rec4 p;
    double ep = 0.0;
    if (staged) {
      p = ssrc[j];
      if (EPSP) ep = sep[j];
    } else {
      const int qid = (FROMB || EPSP) ? S.order[1][j] : 0;
      p = FROMB ? ld256(st.SFbr + 4 * (size_t)(qid - P.nnode)) : ssrc[j];
      if (EPSP) ep = st.epsp[qid];
    }
    const double h1 = msrc[j] * (double)w;  // (mass(j)/rho(j))*w, main:430
    t1 = t1 + p.a * h1;
    t2 = t2 + p.b * h1;
    t3 = t3 + p.c * h1;
    t4 = t4 + p.d * h1;
    if (EPSP) te = te + ep * h1;
  });
  if (!live) return;
  if (nrm != 0) {
    const double rn = __drcp_rn(nrm);
    s.s1 = div_rn(t1, nrm, rn);
    s.s2 = div_rn(t2, nrm, rn);
    s.s3 = div_rn(t3, nrm, rn);
    s.s4 = div_rn(t4, nrm, rn);
    if (EPSP) st.epsp[id] = div_rn(te, nrm, rn);
  } else {
    v.x = 0;
    v.y = 0;
  }
  if (do_adapt) adapt_stress(P, s);
  if (do_bc) apply_bcs(P, st.bc_or_not, st.bc_info, st.bc_int, st.fs_normal, id, v, s);
  R.NBs[k] = Rec4{v.x, v.y, mr.x, mr.y};
  if (final_sweep || FROMB) strec(st.NB, id, v.x, v.y, mr.x, mr.y);  // state between steps / input of k_rk_begin
  st4(st.NSb, id, s);
}

// ------------------------------------------------------------------------------------------------------
// Sweep B, stress-particle side: velocity gradient (main:519-525, wall term main:552-575), CSPM correction, div1,
// plastic_terms, Jaumann terms, RK4 stage accumulation and next-stage predictor (or the final update).
// ------------------------------------------------------------------------------------------------------
constexpr int TB_SP_CAP = 384;
__global__ void __launch_bounds__(TS_T, TILE_MINB(TS_T))
k_tile_b_sp(DevParams P, SlotMap M, const GridInfo *__restrict__ G, SortArrays S, TileLists L, SortedConsts C,
            TileRecs R, StatePtrs st, const TileGeom *__restrict__ geoms, int rev, double f1next, double f2, int last) {
  __shared__ Rec4 sn_[TB_SP_CAP];
  __shared__ double srr[TB_SP_CAP];
  __shared__ ListSmem<2, TS_T, LG_SP0> sl;
  const int kb0 = blockIdx.x * TS_T;
  if (kb0 >= M.ns) return;
  const TileGeom tg = geoms[blockIdx.x];
  const int k0 = kb0 + threadIdx.x;
  const bool live = k0 < M.ns;
  const int k = live ? k0 : kb0;
  const int t = M.nnp + k0;
  const int id = S.order[1][k];
  const int ks = id - P.nnode;
  const LaneList q = lane_list0(L, t, live, tg);
  list_stage(sl, L.code0, L.gx0, L.gy0, q);
  stage_rows_async(sn_, (const Rec4 *)R.NBs, tg);
  stage_rows_async(srr, C.rrho[0], tg);
  const Rec4 selfv = ldrec(st.SVb, ks);
  const double2 vp = make_double2(selfv.a, selfv.b);
  const Stress4 sp_ = ld4(st.SFb, ks);
  int jW[3] = {0, 0, 0};
  if (q.wallp) lane_rows(G, S.cell[1][k], S.start[2], jW);
  WalkWall ww;
  ww.start(rev != 0);
  // operands of the stage epilogue: requested now, consumed after the pair sum
  double ep_ = st.epsp[id], fd_ = st.fdp[id];
  const double rke0 = st.RKe[ks];
  Stress4 rk = ld4(st.RKs, ks);
  const Stress4 s0 = ld4(st.stress0, ks);
  cp_async_wait_all();
  __syncthreads();
  const bool staged = tg.staged != 0;
  const Rec4 *__restrict__ nsrc = staged ? sn_ : (const Rec4 *)R.NBs;
  const double *__restrict__ rsrc = staged ? srr : C.rrho[0];
  double g11 = 0.0, g12 = 0.0, g21 = 0.0, g22 = 0.0;  // grad1_tmp(d,k): d velocity component, k direction
  walk_list(sl, L.code0, L.gx0, L.gy0, q, [&](auto nowall, unsigned cd, float gxf, float gyf) {
    const double gx = (double)gxf, gy = (double)gyf;
    if (decltype(nowall)::value || cd < TILE_CODE_WALL) {  // type 1: velocity particle {vx, vy, m, rho}
      const int j = code_index(q, cd);
      const Rec4 p = nsrc[j];
      const double rr = rsrc[j];
      const double h1 = div_rn(gx * p.c, p.d, rr);  // dwdx*mass(i)/rho(i), main:514
      const double h2 = div_rn(gy * p.c, p.d, rr);
      const double dvx = p.a - vp.x, dvy = p.b - vp.y;
      g11 = g11 + dvx * h1;
      g12 = g12 + dvx * h2;
      g21 = g21 + dvy * h1;
      g22 = g22 + dvy * h2;
    } else {  // type 9: wall particle (no-slip mirror velocity), main:552-575
      const int pq = S.order[2][ww.next(L, t, rev != 0, jW)];
      const double2 xp = ld2(st.x, id);
      const double beta_max = 1.5, vel_wall = 0.0;
      const double wall = (double)st.wallpos[pq];
      const double2 xq = ld2(st.x, pq);
      double da, db;
      if (st.horiz[pq] == 1.f) {
        da = fabs(xp.y - wall);
        db = fabs(xq.y - wall);
      } else {
        da = fabs(xp.x - wall);
        db = fabs(xq.x - wall);
      }
      const double bq = 1 + (db / da);
      const double beta = (bq < beta_max) ? bq : beta_max;
      const double dvx = vp.x * (1 - beta) + beta * vel_wall;
      const double dvy = vp.y * (1 - beta) + beta * vel_wall;
      const double mq = st.mass[pq], rq = st.rho[pq];
      const double h1 = gx * mq / rq;
      const double h2 = gy * mq / rq;
      g11 = g11 + (vp.x - dvx) * h1;
      g12 = g12 + (vp.x - dvx) * h2;
      g21 = g21 + (vp.y - dvy) * h1;
      g22 = g22 + (vp.y - dvy) * h2;
    }
  });
  if (!live) return;
  if (P.cspm) {
    const double *AEp = st.AE + 5 * (size_t)id;
    const double ae1 = AEp[0], ae2 = AEp[1], ae3 = AEp[2], ae4 = AEp[3], ae5 = AEp[4];
    // main:619-622: the second statement sees the already-corrected first column
    g11 = ae5 * (ae1 * g11 + ae2 * g12);
    g12 = ae5 * (ae3 * g11 + ae4 * g12);
    g21 = ae5 * (ae1 * g21 + ae2 * g22);
    g22 = ae5 * (ae3 * g21 + ae4 * g22);
  }
  // div1, main:633-636
  const double d1 = -(P.D11 * g11 + P.D12 * g22);
  const double d2 = -(P.D12 * g11 + P.D22 * g22);
  const double d3 = -(P.D33 * g21 + P.D33 * g12);
  const double d4 = -(P.D41 * g11 + P.D42 * g22);
  // plastic_terms, mat:1884-1954
  double Gs[4] = {0.0, 0.0, 0.0, 0.0}, der1 = 0.0;
  plastic_terms(P, sp_, g11, g12, g21, g22, &ep_, &fd_, Gs, der1);
  if (P.ncrit == 12) st.fdp[id] = fd_;  // f_drucker is read and rewritten by drucker_prager only
  const double rke = rke0 + der1 * f2;
  // Jaumann terms, main:751-757
  double sp1 = 0.0, sp2 = 0.0, sp3 = 0.0, sp4 = 0.0;
  if (P.update_x) {
    const double o1 = 0.5 * (g12 - g21), o2 = -0.5 * (g12 - g21);
    sp1 = 2 * o1 * sp_.s3;
    sp2 = 2 * o2 * sp_.s3;
    sp3 = o2 * sp_.s1 + o1 * sp_.s2;
  }
  const double r1 = -d1 + sp1 + Gs[0];
  const double r2 = -d2 + sp2 + Gs[1];
  const double r3 = -d3 + sp3 + Gs[2];
  const double r4 = -d4 + sp4 + Gs[3];
  rk.s1 = rk.s1 + f2 * r1;
  rk.s2 = rk.s2 + f2 * r2;
  rk.s3 = rk.s3 + f2 * r3;
  rk.s4 = rk.s4 + f2 * r4;
  Stress4 sn;
  if (!last) {
    st4(st.RKs, ks, rk);
    st.RKe[ks] = rke;
    sn.s1 = s0.s1 + f1next * (P.dt) * r1;
    sn.s2 = s0.s2 + f1next * (P.dt) * r2;
    sn.s3 = s0.s3 + f1next * (P.dt) * r3;
    sn.s4 = s0.s4 + f1next * (P.dt) * r4;
  } else {
    sn.s1 = s0.s1 + (P.dt / 6) * rk.s1;
    sn.s2 = s0.s2 + (P.dt / 6) * rk.s2;
    sn.s3 = s0.s3 + (P.dt / 6) * rk.s3;
    sn.s4 = s0.s4 + (P.dt / 6) * rk.s4;
    // update_strain, mat:1864-1880 with Ddev_strn = RK_dev_strain/6 (main:799)
    st.epsp[id] = ep_ + P.dt * (rke / 6);
  }
  if (P.adapt) adapt_stress(P, sn);
  double2 vn = vp;
  apply_bcs(P, st.bc_or_not, st.bc_info, st.bc_int, st.fs_normal, id, vn, sn);
  R.SAs[k] = Rec4{sn.s1, sn.s2, sn.s3, sn.s4};
  st2(st.SVa, ks, vn);
}

// ------------------------------------------------------------------------------------------------------
// Sweep B, velocity-particle side: (1/rho) grad sigma (main:529-536, wall term main:577-588), CSPM correction, div2,
// gravity / damping, artificial viscosity of the stage, RK4 accumulation and predictor.
// ------------------------------------------------------------------------------------------------------
constexpr int TB_N_CAP = 544;
__global__ void __launch_bounds__(TN_T, TILE_MINB(TN_T))
k_tile_b_node(DevParams P, SlotMap M, const GridInfo *__restrict__ G, SortArrays S, TileLists L, TileRecs R,
              StatePtrs st, const TileGeom *__restrict__ geoms, int rev, double f1next, double f2, int last,
              int extra_forces) {
  __shared__ Rec4 ssb[TB_N_CAP];
  __shared__ ListSmem<2, TN_T, LG_N0> sl;
  const int kb0 = blockIdx.x * TN_T;
  if (kb0 >= M.nn) return;
  const TileGeom tg = geoms[blockIdx.x];
  const int k0 = kb0 + threadIdx.x;
  const bool live = k0 < M.nn;
  const int k = live ? k0 : kb0;
  const int t = k0;
  const int id = S.order[0][k];
  const LaneList q = lane_list0(L, t, live, tg);
  list_stage(sl, L.code0, L.gx0, L.gy0, q);
  stage_rows_async(ssb, (const Rec4 *)R.SBs, tg);
  const Rec4 self = R.NBs[k];  // {vx, vy, m, rho}
  const double2 vp = make_double2(self.a, self.b);
  const double rp = self.d;
  const Stress4 sp_ = ld4(st.NSb, id);
  int jW[3] = {0, 0, 0};
  if (q.wallp) lane_rows(G, S.cell[0][k], S.start[2], jW);
  WalkWall ww;
  ww.start(rev != 0);
  const double r2p = rp * rp, rr2p = __drcp_rn(r2p);
  const double so1 = div_rn(sp_.s1, r2p, rr2p), so2 = div_rn(sp_.s2, r2p, rr2p),
               so3 = div_rn(sp_.s3, r2p, rr2p);  // stress(1:3,i)/rho(i)**2
  // operands of the stage epilogue: requested now, consumed after the pair sum
  double2 avp = make_double2(0.0, 0.0), fb = make_double2(0.0, 0.0), af = make_double2(0.0, 0.0);
  if (P.alpha > 0 || P.beta > 0) avp = ld2(st.av, id);  // artificial viscosity of this stage (k_tile_av)
  if (extra_forces) {
    fb = ld2(st.fbound, id);  // f_bound (main:764): zero unless boundary_forces ran
    af = ld2(st.aforce, id);  // art_force: zero unless art_stress = T
  }
  double2 rk = ld2(st.RKv, id);
  const double2 v0 = ld2(st.vel0, id);
  cp_async_wait_all();
  __syncthreads();
  const Rec4 *__restrict__ ssrc = tg.staged ? ssb : (const Rec4 *)R.SBs;
  double a11 = 0.0, a12 = 0.0, a21 = 0.0, a22 = 0.0, a31 = 0.0, a32 = 0.0;  // grad2_tmp(s,k)
  walk_list(sl, L.code0, L.gx0, L.gy0, q, [&](auto nowall, unsigned cd, float gxf, float gyf) {
    const double gx = (double)gxf, gy = (double)gyf;
    double q1, q2, q3, mq;
    if (decltype(nowall)::value || cd < TILE_CODE_WALL) {  // type 1: stress particle {s1/rho^2, s2/rho^2, s3/rho^2, m}
      const Rec4 p = ssrc[code_index(q, cd)];
      q1 = p.a;
      q2 = p.b;
      q3 = p.c;
      mq = p.d;
    } else {  // type 6: the wall particle takes the velocity particle's stress (main:580)
      const int pq = S.order[2][ww.next(L, t, rev != 0, jW)];
      const double rq = st.rho[pq];
      mq = st.mass[pq];
      q1 = sp_.s1 / (rq * rq);
      q2 = sp_.s2 / (rq * rq);
      q3 = sp_.s3 / (rq * rq);
    }
    const double c1 = so1 + q1, c2 = so2 + q2, c3 = so3 + q3;
    a11 = a11 - mq * (gx * c1);
    a12 = a12 - mq * (gy * c1);
    a21 = a21 - mq * (gx * c2);
    a22 = a22 - mq * (gy * c2);
    a31 = a31 - mq * (gx * c3);
    a32 = a32 - mq * (gy * c3);
  });
  if (!live) return;
  if (P.cspm) {
    const double *AEp = st.AE + 5 * (size_t)id;
    const double ae1 = AEp[0], ae2 = AEp[1], ae3 = AEp[2], ae4 = AEp[3], ae5 = AEp[4];
    // main:623-626: only stress components 1..ndimn are corrected
    a11 = ae5 * (ae1 * a11 + ae2 * a12);
    a12 = ae5 * (ae3 * a11 + ae4 * a12);
    a21 = ae5 * (ae1 * a21 + ae2 * a22);
    a22 = ae5 * (ae3 * a21 + ae4 * a22);
  }
  const double dv1 = -(a11 + a32);  // div2, main:641-642
  const double dv2 = -(a31 + a22);
  // gravity_force, mat:2809-2871
  const double sg1 = P.grav[0] - P.damping * vp.x;
  const double sg2 = P.grav[1] - P.damping * vp.y;
  // artificial viscosity of this stage is zero when alpha = beta = 0 (art_visc stays 0, main:688)
  const double r1 = -dv1 + sg1 + avp.x + fb.x + af.x;
  const double r2 = -dv2 + sg2 + avp.y + fb.y + af.y;
  rk.x = rk.x + f2 * r1;
  rk.y = rk.y + f2 * r2;
  double2 vn;
  if (!last) {
    st2(st.RKv, id, rk);
    vn.x = v0.x + f1next * (P.dt) * r1;
    vn.y = v0.y + f1next * (P.dt) * r2;
  } else {
    vn.x = v0.x + (P.dt / 6) * rk.x;
    vn.y = v0.y + (P.dt / 6) * rk.y;
  }
  Stress4 sn = sp_;
  if (P.adapt) adapt_stress(P, sn);
  apply_bcs(P, st.bc_or_not, st.bc_info, st.bc_int, st.fs_normal, id, vn, sn);
  R.NAs[k] = vn;
  st4(st.NSa, id, sn);
}

// ------------------------------------------------------------------------------------------------------
// artificial_viscosity, main:826-904 (fp32 locals and accumulators in list order) over the velocity-velocity list;
// xij, yij are re-derived from the staged positions (the list path stored their fp32 roundings).
// ------------------------------------------------------------------------------------------------------
constexpr int TAV_CAP = 288;
__global__ void __launch_bounds__(TN_T, TILE_MINB(TN_T))
k_tile_av(DevParams P, SlotMap M, SortArrays S, TileLists L, TileRecs R, StatePtrs st,
          const TileGeom *__restrict__ geoms, float h_u) {
  __shared__ Rec4 sn_[TAV_CAP];
  __shared__ double2 sx[TAV_CAP];
  __shared__ ListSmem<2, TN_T, LG_NC> sl;
  const int kb0 = blockIdx.x * TN_T;
  if (kb0 >= M.nn) return;
  const TileGeom tg = geoms[blockIdx.x];
  const int k0 = kb0 + threadIdx.x;
  const bool live = k0 < M.nn;
  const int k = live ? k0 : kb0;
  const int t = k0;
  const int id = S.order[0][k];
  const LaneList q = lane_listS(L, t, live, tg);
  list_stage(sl, L.codeS, L.gxC, L.gyC, q);
  stage_rows_async(sn_, (const Rec4 *)R.NBs, tg);
  stage_rows_async(sx, S.pos[0], tg);
  const Rec4 self = R.NBs[k];
  const double2 vp = make_double2(self.a, self.b);
  const double rp = self.d;
  const double2 pp = S.pos[0][k];
  cp_async_wait_all();
  __syncthreads();
  const Rec4 *__restrict__ nsrc = tg.staged ? sn_ : (const Rec4 *)R.NBs;
  const double2 *__restrict__ xsrc = tg.staged ? sx : S.pos[0];
  float acc1 = 0.f, acc2 = 0.f;
  walk_list(sl, L.codeS, L.gxC, L.gyC, q, [&](auto, unsigned cd, float gxf, float gyf) {
    const int j = code_index(q, cd);
    const Rec4 p = nsrc[j];
    const double2 pq = xsrc[j];
    const float xij = (float)(pp.x - pq.x), yij = (float)(pp.y - pq.y);  // main:856-857
    const float h = h_u;
    const float rho2 = (float)(0.5 * (rp + p.d));
    const float cs = 600.f;
    float div_u = (float)((double)xij * (vp.x - p.a));
    div_u = (float)((double)div_u + (double)yij * (vp.y - p.b));
    const float sq = sqrtf(xij * xij + yij * yij);
    const float theta = (h * div_u) / (sq * sq + 0.01f * (h * h));
    const double rho2d = (double)rho2;
    const double num = -P.alpha * (double)cs * (double)theta + P.beta * (double)(theta * theta);
    const float vv = (float)div_rn(num, rho2d, __drcp_rn(rho2d));
    const float visc = (div_u < 0) ? vv : 0.f;
    acc1 = (float)((double)acc1 + (double)(visc * gxf) * p.c);  // ordered fp32 accumulation
    acc2 = (float)((double)acc2 + (double)(visc * gyf) * p.c);
  });
  if (!live) return;
  st2(st.av, id, make_double2((double)(-acc1), (double)(-acc2)));  // art_visc = -art_visc_temp, main:901
}

// ------------------------------------------------------------------------------------------------------
// Position update, main:140-182: XSPH_update (main:189-239; w re-evaluated from the staged positions instead of
// streamed: the same-species weights have no other reader) or the fp32 mid-velocity rule; displ.
// One launch, block-uniform role: velocity particles first, then stress particles.
// ------------------------------------------------------------------------------------------------------
constexpr int TMV_CAP = 640;
__global__ void __launch_bounds__(TS_T, TILE_MINB(TS_T))
k_tile_move(DevParams P, SlotMap M, SortArrays S, TileLists L, SortedConsts C, TileRecs R, StatePtrs st,
            const TileGeom *__restrict__ geoms_n, const TileGeom *__restrict__ geoms_s, int nb_node,
            double *__restrict__ x, const double *__restrict__ x00, double *__restrict__ displ) {
  const bool is_node = (int)blockIdx.x < nb_node;  // block-uniform
  const int sp = is_node ? SP_NODE : SP_STRESS;
  __shared__ double2 sv[TMV_CAP], sx[TMV_CAP];
  __shared__ double smo[TMV_CAP];
  __shared__ ListSmem<0, TS_T, LG_SS> sl;
  const int bq = is_node ? blockIdx.x : blockIdx.x - nb_node;
  const int kb0 = bq * TS_T;
  const int nlive = is_node ? M.nn : M.ns;
  if (kb0 >= nlive) return;
  const bool xs = P.update_x && P.xsph;
  TileGeom tg;
  tg.staged = 0;
  if (xs) tg = (is_node ? geoms_n : geoms_s)[bq];
  const int k0 = kb0 + threadIdx.x;
  const bool live = k0 < nlive;
  const int k = live ? k0 : kb0;
  const int t = is_node ? k0 : M.nnp + k0;
  const int id = S.order[sp][k];
  LaneList q;
  q.cnt = 0;
  if (xs) {
    q = lane_listS(L, t, live, tg);
    list_stage(sl, L.codeS, nullptr, nullptr, q);
    if (tg.staged) {
      if (is_node) {
        for (int r = 0; r < 3; ++r)
          for (int i = threadIdx.x; i < tg.cnt[r]; i += blockDim.x) {
            const Rec4 p = R.NBs[tg.base[r] + i];
            sv[tg.off[r] + i] = make_double2(p.a, p.b);
          }
      } else {
        stage_rows_async(sv, (const double2 *)R.SVs, tg);
      }
      stage_rows_async(sx, S.pos[sp], tg);
      stage_rows_async(smo, C.mor[sp], tg);
    }
  }
  double2 vp;
  if (is_node) {
    const Rec4 r = R.NBs[k];
    vp = make_double2(r.a, r.b);
  } else {
    vp = R.SVs[k];
  }
  const double2 pp = S.pos[sp][k];
  const double2 xp = ld2(x, id);
  cp_async_wait_all();
  __syncthreads();
  double sx_ = 0.0, sy_ = 0.0;
  if (xs) {
    const bool staged = tg.staged != 0;
    const double2 *__restrict__ xsrc = staged ? sx : S.pos[sp];
    const double *__restrict__ msrc = staged ? smo : C.mor[sp];
    const KernelConsts K = kernel_consts(P, S.h[sp][k]);
    walk_list(sl, L.codeS, nullptr, nullptr, q, [&](auto, unsigned cd, float, float) {
      const int j = code_index(q, cd);
      double2 vq;
      if (staged) {
        vq = sv[j];
      } else if (is_node) {
        const Rec4 p = ldrec(R.NBs, j);
        vq = make_double2(p.a, p.b);
      } else {
        vq = R.SVs[j];
      }
      const double2 pq = xsrc[j];
      const double dx = pp.x - pq.x, dy = pp.y - pq.y;
      double d2 = dx * dx;
      d2 = d2 + dy * dy;
      const double r = sqrt(d2);
      double w, gx, gy;
      sph_kernel_fast<false>(K, r, dx, dy, w, gx, gy);
      const double wd = (double)(float)w;  // pairs%w is fp32 (main:1376)
      const double mo = msrc[j];
      sx_ = sx_ + mo * (vq.x - vp.x) * wd;
      sy_ = sy_ + mo * (vq.y - vp.y) * wd;
    });
  }
  if (!live) return;
  if (P.update_x) {
    double2 xn;
    if (P.xsph) {
      const double eps = 0.5;
      xn.x = xp.x + P.dt * (vp.x + eps * sx_);
      xn.y = xp.y + P.dt * (vp.y + eps * sy_);
    } else {
      const double2 v0 = ld2(st.vx0, id);
      const float hx = (float)(0.5 * (v0.x + vp.x));  // real :: vel_half, main:89,145
      const float hy = (float)(0.5 * (v0.y + vp.y));
      xn.x = xp.x + (double)hx * P.dt;
      xn.y = xp.y + (double)hy * P.dt;
    }
    st2(x, id, xn);
    if (is_node) {
      const double2 x0 = ld2(x00, id);
      st2(displ, id, make_double2(xn.x - x0.x, xn.y - x0.y));  // main:171
    }
  } else if (is_node) {
    const double2 v0 = ld2(st.vx0, id);
    double2 d = ld2(displ, id);
    d.x = d.x + 0.5 * (v0.x + vp.x) * P.dt;  // main:180
    d.y = d.y + 0.5 * (v0.y + vp.y) * P.dt;
    st2(displ, id, d);
  }
}

}  // namespace spsph
