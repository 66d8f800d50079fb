// Pair-sum and per-particle kernels of one time step (reference: stress_point_update main:403-482,
// get_derivatives main:487-648, RK4 main:653-802, artificial_viscosity main:826-904, plastic_terms
// mat:1884-1954, adapt_stress2 mat:2087-2161, Normal_BCs mat:1665-1770, gravity_force mat:2809-2871,
// XSPH_update main:189-239, shift_stress_points main:244-368, position update main:140-182).
//
// One thread per velocity / stress particle, walking its gather list front to back: every per-particle sum
// is accumulated in the reference's traversal order (fp32 sums bit-exact, fp64 sums too), no atomics.
//
// Data layout. The pair-sum kernels are bound by the L1 rate of *gathered sectors* (about one 32-byte sector
// per clock per SM for scattered lanes; ncu: l1tex__data_pipe_lsu_wavefronts at 86 % in round-1 profiles), not by
// HBM. The state therefore lives in two buffer sets with *different record formats*, each written by the kernel
// that precedes its reader, such that every neighbour gather is ONE aligned sector (or half of one):
//   format A (input of sweep A = stress_point_update):
//       NA[node] = {vx, vy} (16 B)               SA[sp] = {s1, s2, s3, s4} (32 B)
//       NSa[node] = own stress (carried)         SVa[sp] = own velocity (carried)
//       the factor (m/rho)_partner * w of main:430-431 is streamed with the list (ListPtrs::h0lo/h0hi)
//   format B (input of sweep B = get_derivatives; also the state between time steps):
//       NB[node] = {vx, vy, m, rho}              SB[sp] = {s1/rho^2, s2/rho^2, s3/rho^2, m}
//       NSb[node] = own stress                   SFb[sp] = own stress, SVb[sp] = {vx, vy, m/rho, -}
// Sweep A maps A -> B, sweep B (with the fused RK4 stage epilogue and next-stage predictor) maps B -> A; a
// kernel only gathers from the format it does not write, so there are no read/write races.
// List rows are streamed through shared memory with cp.async (ell_stream), four entries in flight per thread.
#pragma once
#include "spsph.h"  // SPSPH_COL_* column codes of k_pack_frame
#include "grid_kernels.cuh"

namespace spsph {

struct __align__(32) Rec4 {
  double a, b, c, d;
};

struct StatePtrs {
  // constant per step / persistent (original particle order)
  const double *x;  // (2, ntotal2)
  const double *mass, *rho, *hsml, *mor;
  const double2 *mrho;  // {mass, rho}: one gather where a sweep needs both (the cspm_norm pass)
  const float *wallpos, *horiz;
  const int *bc_or_not, *bc_info;
  const int *bc_int;        // (nnode) velocity particle next to a wall particle (k_fill, main:506,579)
  const double *fs_normal;  // (2, nnode) free-surface normals of apply_stress_free (k_fs_normals; ifsigman = 1 only)
  // format A
  double *NA;   // (2, nnode) velocity
  Rec4 *SA;     // [nstress] stress
  double *NSa;  // (4, nnode)
  double *SVa;  // (2, nstress)
  // format B (== state between steps)
  Rec4 *NB;     // [nnode] {vx, vy, m, rho}
  Rec4 *SB;     // [nstress] {s1/rho^2, s2/rho^2, s3/rho^2, m}
  double *NSb;  // (4, nnode)
  double *SFb;  // (4, nstress)
  Rec4 *SVb;    // [nstress] {vx, vy, m/rho, -}
  double *av;   // (2, nnode) artificial viscosity acceleration of the current stage (k_artvisc -> k_sweep_b_node)
  double *fbound;  // (2, nnode) boundary_forces of the current step (zero unless the inside approach has walls)
  double *aforce;  // (2, nnode) artificial_force of the current stage (zero unless art_stress = T)
  int has_fbound, has_aforce;  // 0: the array holds zeros (the feature is off): sweep B does not read it
  Rec4 *RN;        // [nnode] artificial-stress terms R(1:3) of a node (k_art_force_prep -> k_art_force)
  // continuity density (cont_density = T): the density of the stress particles is integrated by RK4, that of the
  // velocity particles interpolated by every stress_point_update; with sle = 2 the smoothing length follows
  double *rho_w, *hsml_w, *mor_w;  // writable views of rho, hsml, mor
  double2 *mrho_w;                 // writable view of mrho
  double *rho_new;                 // (nnode) interpolated density of a sweep A, committed by k_commit_node_rho
  double *rho0, *hsml0, *RKrho, *RKh;  // (nstress) RK4 work arrays, main:686-689
  double *divu;                    // (nstress) grad_u(1,1) + grad_u(2,2) of the last get_derivatives (persists across steps)
  // read side of a format-B -> format-B sweep (the SPH_shift interpolation): the other B buffer set
  const Rec4 *NBr, *SVbr;
  const double *NSbr, *SFbr;
  double *epsp;     // (ntotal) Internal_Vars(1,:)
  double *fdp;      // (ntotal) f_drucker
  double *norm;     // (ntotal) cspm_norm of stress_point_update (frozen within a step)
  double *AE;       // (5, ntotal) inverted CSPM matrix of get_derivatives (frozen within a step)
  double *vel0;     // (2, nnode)
  double *stress0;  // (4, nstress)
  double *vx0;      // (2, ntotal)
  double *RKv;      // (2, nnode)
  double *RKs;      // (4, nstress)
  double *RKe;      // (nstress)
};

// 32-byte records move with one 256-bit access (LDG.E.ENL2.256 / STG.E.ENL2.256 on sm_100a): a gathered record
// costs one L1 sector access instead of two 128-bit ones -- the L1 sector rate bounds the pair-sum kernels.
#ifndef SPSPH_HOST_EMU  // device only (the host emulation of tests/native/ has its own version)
__device__ __forceinline__ Rec4 ld256(const void *p) {
  Rec4 r;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p));
  return r;
}
__device__ __forceinline__ void st256(void *p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
#else
inline Rec4 ld256(const void *p) { return *static_cast<const Rec4 *>(p); }
inline void st256(void *p, double a, double b, double c, double d) { *static_cast<Rec4 *>(p) = Rec4{a, b, c, d}; }
#endif  // SPSPH_HOST_EMU
__device__ __forceinline__ Rec4 ldrec(const Rec4 *p, int i) { return ld256(p + i); }
__device__ __forceinline__ void strec(Rec4 *p, int i, double a, double b, double c, double d) {
  st256(p + i, a, b, c, d);
}
__device__ __forceinline__ int ldcs_i(const int *p) { return __ldcs(p); }
__device__ __forceinline__ float ldcs_f(const float *p) { return __ldcs(p); }

constexpr int UNR = 4;  // list entries in flight per thread

// ------------------------------------------------------------------------------------------------------
// Streaming of a warp's ELL slice through shared memory.
// The slice of one warp is contiguous in every list array (row e = 32 consecutive 4-byte entries, rows back
// to back), so a group of 4 rows of one array is 512 contiguous bytes = one 16-byte cp.async per lane. Each
// warp owns a private ring of ELL_NG groups per array; HBM latency is hidden by the ring depth, independent
// of register count, and the neighbour gathers for group g+1 are issued before group g is consumed.
// ------------------------------------------------------------------------------------------------------
#ifndef SPSPH_MINB
#define SPSPH_MINB 4      // min resident blocks per SM requested from ptxas for the sweep kernels
#endif
#ifndef SPSPH_SWEEP_T
#define SPSPH_SWEEP_T 128  // threads per block of the five pair-sum kernels (tools/variant_timing.sh times 64 x 8 blocks)
#endif
constexpr int SWEEP_T = SPSPH_SWEEP_T;
static_assert(SWEEP_T % 32 == 0 && SWEEP_T >= 32 && SWEEP_T <= 256, "whole warps; each warp owns its ring in shared memory");
#ifndef SPSPH_ELL_GROUP
#define SPSPH_ELL_GROUP 4
#endif
constexpr int ELL_GROUP = SPSPH_ELL_GROUP;  // rows per cp.async group
#ifndef SPSPH_ELL_SUB
#define SPSPH_ELL_SUB 4
#endif
#ifndef SPSPH_ELL_PIPE
#define SPSPH_ELL_PIPE 0  // 1: software-pipelined gathers in ell_stream (variant, see there)
#endif
constexpr int ELL_SUB = SPSPH_ELL_SUB;  // entries gathered + consumed together (in flight per thread)
#ifndef SPSPH_ELL_NG
#define SPSPH_ELL_NG 4
#endif
constexpr int ELL_NG = SPSPH_ELL_NG;     // groups in the ring (16 rows = 2 KB per array per warp in flight)

#ifndef SPSPH_HOST_EMU  // device only (the host emulation of tests/native/ has its own version)
// The list rows are read exactly once per sweep: L2 evict-first, so that the 0.7-1.3 GB streamed per sweep do
// not push the gathered particle records (43-171 MB, re-read ~20-40 times each) out of the 126 MB L2.
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, unsigned long long pol) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

#else
inline unsigned long long l2_policy_evict_first() { return 0; }
inline unsigned long long l2_policy_evict_last() { return 0; }
inline void cp_async16(void *smem, const void *gmem, unsigned long long) { std::memcpy(smem, gmem, 16); }
inline void cp_async_commit() {}
template <int N>
inline void cp_async_wait() {}
#endif  // SPSPH_HOST_EMU
// NARR arrays of 4-byte entries are streamed (array 0 holds the partner ids). `rows` is the slice width
// (warp-uniform), `cnt` the calling lane's own list length. gather(q) -> R fetches the partner record (q < 0:
// past the end); compute(q[4], pay[NARR-1][4], rec[4], nvalid) consumes one group of entries in list order
// (pay: raw 32-bit payloads of arrays 1..NARR-1; entries u >= nvalid are past the end of this lane's list).
// The two top bits of a partner id of list 0 may carry the partner's mass/rho class (QCLASS_SHIFT, written by the fill
// pass when the host found at most four distinct values per species); gather() always receives the plain id, and so
// does compute() unless RAWQ asks for the stored word (sweep A, which turns the class into the factor (m/rho)*w).
#if !defined(SPSPH_HOST_EMU) || defined(SPSPH_EMU_SIMT)  // device, and the lockstep (SIMT) host emulation
template <int NARR, int NG, class R, int GR = ELL_GROUP, int SUB = ELL_SUB, bool RAWQ = false, class GatherF,
          class ComputeF>
__device__ __forceinline__ void ell_stream(const int *const *arr, size_t slice_off, int rows, int cnt, int *smw,
                                           GatherF gather, ComputeF compute) {
  static_assert(GR % 4 == 0 && GR % SUB == 0, "group = whole 16-byte cp.async rows, consumed in SUB-entry parts");
  const int lane = threadIdx.x & 31;
  const int ng = (rows + GR - 1) / GR;
  if (ng == 0) return;
  const unsigned long long pol = l2_policy_evict_first();
  auto issue = [&](int g) {
    if (g < ng) {
#pragma unroll
      for (int a = 0; a < NARR; ++a)
#pragma unroll
        for (int c = 0; c < GR / 4; ++c)
          cp_async16(smw + ((g % NG) * NARR + a) * (GR * 32) + c * 128 + lane * 4,
                     arr[a] + slice_off + (size_t)g * (GR * 32) + c * 128 + lane * 4, pol);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int g = 0; g < NG; ++g) issue(g);
#if SPSPH_ELL_PIPE
  // Variant (tools/variant_timing.sh): the partner records of part s+1 are gathered BEFORE part s is consumed, so the
  // gather latency overlaps the pair arithmetic inside one warp as well (the sweeps run at 16 warps per SM and wait on
  // these gathers ~45 % of the time). Same entries, same order, same arithmetic: bit-identical results.
  static_assert(NG >= 2, "the next group has to be resident while the current one is consumed");
  constexpr int NS = GR / SUB;  // parts per group
  const int nsub = ng * NS;
  int qn[SUB];
  R nxt[SUB];
  auto fetch = [&](int sidx) {
    const int g = sidx / NS, hh = sidx % NS;
    const int *sl = smw + ((g % NG) * NARR) * (GR * 32);
#pragma unroll
    for (int u = 0; u < SUB; ++u) {
      const int raw = sl[(hh * SUB + u) * 32 + lane];
      qn[u] = RAWQ ? raw : (raw & QID_MASK);
      nxt[u] = gather((g * GR + hh * SUB + u) < cnt ? (raw & QID_MASK) : -1);
    }
  };
  cp_async_wait<NG - 1>();
  __syncwarp();
  fetch(0);
  for (int sidx = 0; sidx < nsub; ++sidx) {
    const int g = sidx / NS, hh = sidx % NS;
    int q[SUB], pay[NARR > 1 ? NARR - 1 : 1][SUB];
    R cur[SUB];
#pragma unroll
    for (int u = 0; u < SUB; ++u) {
      q[u] = qn[u];
      cur[u] = nxt[u];
    }
    if (sidx + 1 < nsub) {
      if (hh == NS - 1) {  // the next part opens group g+1: NG+g groups are committed, g+2 of them must have landed
        cp_async_wait<NG - 2>();
        __syncwarp();
      }
      fetch(sidx + 1);
    }
    const int *sl = smw + ((g % NG) * NARR) * (GR * 32);
#pragma unroll
    for (int a = 1; a < NARR; ++a)
#pragma unroll
      for (int u = 0; u < SUB; ++u) pay[a - 1][u] = sl[a * (GR * 32) + (hh * SUB + u) * 32 + lane];
    compute(q, pay, cur, cnt - g * GR - hh * SUB);
    if (hh == NS - 1) {
      __syncwarp();
      issue(g + NG);
    }
  }
  cp_async_wait<0>();
  return;
#endif
  for (int g = 0; g < ng; ++g) {
    cp_async_wait<NG - 1>();
    __syncwarp();
    const int *sl = smw + ((g % NG) * NARR) * (GR * 32);
#pragma unroll
    for (int hh = 0; hh < GR / SUB; ++hh) {
      int q[SUB], pay[NARR > 1 ? NARR - 1 : 1][SUB];
      R cur[SUB];
#pragma unroll
      for (int u = 0; u < SUB; ++u) {
        const int raw = sl[(hh * SUB + u) * 32 + lane];
        q[u] = RAWQ ? raw : (raw & QID_MASK);
        cur[u] = gather((g * GR + hh * SUB + u) < cnt ? (raw & QID_MASK) : -1);
      }
#pragma unroll
      for (int a = 1; a < NARR; ++a)
#pragma unroll
        for (int u = 0; u < SUB; ++u) pay[a - 1][u] = sl[a * (GR * 32) + (hh * SUB + u) * 32 + lane];
      compute(q, pay, cur, cnt - g * GR - hh * SUB);
    }
    __syncwarp();
    issue(g + NG);
  }
  cp_async_wait<0>();
}
#else
// host emulation: the same entries in the same SUB-sized groups, read straight from the list arrays
template <int NARR, int NG, class R, int GR = ELL_GROUP, int SUB = ELL_SUB, bool RAWQ = false, class GatherF,
          class ComputeF>
inline void ell_stream(const int *const *arr, size_t slice_off, int rows, int cnt, int *, GatherF gather,
                       ComputeF compute) {
  const int lane = threadIdx.x & 31;
  const int ng = (rows + GR - 1) / GR;
  for (int g = 0; g < ng; ++g)
    for (int hh = 0; hh < GR / SUB; ++hh) {
      int q[SUB], pay[NARR > 1 ? NARR - 1 : 1][SUB];
      R cur[SUB];
      for (int u = 0; u < SUB; ++u) {
        const int row = g * GR + hh * SUB + u;
        const int raw = row < cnt ? arr[0][slice_off + (size_t)row * 32 + lane] : 0;
        q[u] = RAWQ ? raw : (raw & QID_MASK);
        cur[u] = gather(row < cnt ? (raw & QID_MASK) : -1);
        for (int a = 1; a < NARR; ++a) pay[a - 1][u] = row < cnt ? arr[a][slice_off + (size_t)row * 32 + lane] : 0;
      }
      compute(q, pay, cur, cnt - g * GR - hh * SUB);
    }
}
#endif  // SPSPH_HOST_EMU
__device__ __forceinline__ int warp_max_i(int v) {
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#define ELL_SMEM(NARR, NG) ((NG) * (NARR) * ELL_GROUP * 32)  // ints per warp
#define ELL_SMEM_G(NARR, NG, GR) ((NG) * (NARR) * (GR) * 32)
#ifndef SPSPH_A_GR
#define SPSPH_A_GR 4
#endif
#ifndef SPSPH_A_NG
#define SPSPH_A_NG 4
#endif
#ifndef SPSPH_A_SUB
#define SPSPH_A_SUB SPSPH_A_GR
#endif
constexpr int A_SUB = SPSPH_A_SUB;  // sweep A: entries gathered + consumed together
constexpr int A_GR = SPSPH_A_GR, A_NG = SPSPH_A_NG;  // sweep A: rows per group (all in flight per thread), ring depth

// Per-kernel build switches of the five pair-sum kernels: entries gathered + consumed together (registers) and the
// resident blocks per SM requested from ptxas. Defaults = the values measured best on the 4 M-particle column
// (per kernel, against 4 entries / 4 blocks: k_sweep_b_sp 2 / 6 -0.08 ms per step, k_artvisc 2 / 6 -0.09, k_sweep_a_sp
// 2 / 8 -0.03; the velocity-particle sides are best left at 4 / 4).
#ifndef SPSPH_ASP_SUB
#define SPSPH_ASP_SUB 2
#endif
#ifndef SPSPH_AN_SUB
#define SPSPH_AN_SUB SPSPH_A_SUB
#endif
#ifndef SPSPH_BSP_SUB
#define SPSPH_BSP_SUB 2
#endif
#ifndef SPSPH_BN_SUB
#define SPSPH_BN_SUB SPSPH_ELL_SUB
#endif
#ifndef SPSPH_AV_SUB
#define SPSPH_AV_SUB 2
#endif
#ifndef SPSPH_ASP_MINB
#define SPSPH_ASP_MINB 8
#endif
#ifndef SPSPH_AN_MINB
#define SPSPH_AN_MINB SPSPH_MINB
#endif
#ifndef SPSPH_BSP_MINB
#define SPSPH_BSP_MINB 6
#endif
#ifndef SPSPH_BN_MINB
#define SPSPH_BN_MINB SPSPH_MINB
#endif
#ifndef SPSPH_AV_MINB
#define SPSPH_AV_MINB 6
#endif
constexpr int ASP_SUB = SPSPH_ASP_SUB, AN_SUB = SPSPH_AN_SUB, BSP_SUB = SPSPH_BSP_SUB, BN_SUB = SPSPH_BN_SUB,
              AV_SUB = SPSPH_AV_SUB;

// state format conversions at the boundary of the time loop ---------------------------------------------
// pack: reference-layout vel/stress (upload) -> format B
__global__ void k_pack_state(DevParams P, const double *__restrict__ vel, const double *__restrict__ stress,
                             StatePtrs st) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= P.ntotal) return;
  const double2 v = ld2(vel, id);
  const Stress4 s = ld4(stress, id);
  if (id < P.nnode) {
    strec(st.NB, id, v.x, v.y, st.mass[id], st.rho[id]);
    st4(st.NSb, id, s);
  } else {
    const int ks = id - P.nnode;
    const double r = st.rho[id];
    const double r2 = r * r;
    strec(st.SB, ks, s.s1 / r2, s.s2 / r2, s.s3 / r2, st.mass[id]);
    st4(st.SFb, ks, s);
    strec(st.SVb, ks, v.x, v.y, st.mor[id], 0.0);
  }
}
// unpack: format B -> reference-layout vel (2,ntotal) and stress (4,ntotal) for download
__global__ void k_unpack_state(DevParams P, StatePtrs st, double *__restrict__ vel, double *__restrict__ stress) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= P.ntotal) return;
  if (id < P.nnode) {
    const Rec4 r = ldrec(st.NB, id);
    st2(vel, id, make_double2(r.a, r.b));
    st4(stress, id, ld4(st.NSb, id));
  } else {
    const int ks = id - P.nnode;
    const Rec4 r = ldrec(st.SVb, ks);
    st2(vel, id, make_double2(r.a, r.b));
    st4(stress, id, ld4(st.SFb, ks));
  }
}

// Output frame packed on the device (OutputRes, mat:2919-3060: the ParaView rows "x, y, [vel], [stress], [strain],
// [disp_10], [density], [sml]"; GiD results mat:2930-3008): one thread per (particle, column) element of a
// row-major (count, ncols) table, so a frame leaves the device in ONE copy holding only the columns the writers print.
// Velocity-particle-only columns read 0 for stress particles, wall particles carry positions / rho / h only.
struct FrameCols {
  int n;
  unsigned long long codes;  // 4 bits per column (SPSPH_COL_COUNT <= 16, SPSPH_FRAME_MAX_COLS = 16)
};
__global__ void k_pack_frame(DevParams P, StatePtrs st, FrameCols C, const double *__restrict__ displ,
                             const double *__restrict__ disp_10, int first, long long nelem, double *__restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nelem) return;
  const int id = first + (int)(e / C.n);
  const int c = (int)(C.codes >> (4 * (int)(e % C.n))) & 15;
  const bool node = id < P.nnode, part = id < P.ntotal;
  double v = 0.0;
  switch (c) {
    case SPSPH_COL_X: v = st.x[2 * (size_t)id]; break;
    case SPSPH_COL_Y: v = st.x[2 * (size_t)id + 1]; break;
    case SPSPH_COL_VX:
    case SPSPH_COL_VY:
      if (part) {
        const Rec4 r = node ? ldrec(st.NB, id) : ldrec(st.SVb, id - P.nnode);
        v = c == SPSPH_COL_VX ? r.a : r.b;
      }
      break;
    case SPSPH_COL_SXX:
    case SPSPH_COL_SYY:
    case SPSPH_COL_SXY:
    case SPSPH_COL_SZZ:
      if (part) v = (node ? st.NSb + 4 * (size_t)id : st.SFb + 4 * (size_t)(id - P.nnode))[c - SPSPH_COL_SXX];
      break;
    case SPSPH_COL_EPSP: if (part) v = st.epsp[id]; break;
    case SPSPH_COL_FDRUCKER: if (part) v = st.fdp[id]; break;
    case SPSPH_COL_DISP10: if (node) v = disp_10[id]; break;
    case SPSPH_COL_DISPLX: if (node) v = displ[2 * (size_t)id]; break;
    case SPSPH_COL_DISPLY: if (node) v = displ[2 * (size_t)id + 1]; break;
    case SPSPH_COL_RHO: v = st.rho[id]; break;
    case SPSPH_COL_HSML: v = st.hsml[id]; break;
    case SPSPH_COL_BC_OR_NOT: if (part) v = (double)st.bc_or_not[id]; break;
    default: break;
  }
  out[e] = v;
}

// ------------------------------------------------------------------------------------------------------
// RK4 prologue (main:681-690 + first predictor main:700-701 with f1rk = 0 + adapt_stress2/BCs main:715-716):
// format B (state) -> format A (stage-1 input); saves vel0/stress0/vx0 and zeroes the RK accumulators.
// Stress-particle velocities and node stresses start from zero (main:690).
// ------------------------------------------------------------------------------------------------------
// pos_of != nullptr (cell-tile path): the partner-visible stage-1 records go to the species-sorted arrays NAs / SAs.
__global__ void k_rk_begin(DevParams P, StatePtrs st, LocalList LL, const int *__restrict__ pos_of,
                           double2 *__restrict__ NAs, Rec4 *__restrict__ SAs) {
  SPSPH_FOR_LOCAL(LL, kk, id) {
  if (id >= P.ntotal) continue;  // wall particle
  double2 vn;
  Stress4 sn;
  if (id < P.nnode) {
    const Rec4 r = ldrec(st.NB, id);
    const double2 v = make_double2(r.a, r.b);
    st2(st.vx0, id, v);
    st2(st.vel0, id, v);
    if (pos_of) st2(st.RKv, id, make_double2(0.0, 0.0));  // id-list path: stage 1 of sweep B starts from zero itself
    vn.x = v.x + 0. * (P.dt) * 0.0;  // vel0 + f1rk(1)*dt*RHS_2 with RHS_2 = 0
    vn.y = v.y + 0. * (P.dt) * 0.0;
    sn = Stress4{0.0, 0.0, 0.0, 0.0};
    if (P.adapt) adapt_stress(P, sn);
    apply_bcs(P, st.bc_or_not, st.bc_info, st.bc_int, st.fs_normal, id, vn, sn);
    if (pos_of)
      NAs[pos_of[id]] = vn;
    else
      st2(st.NA, id, vn);
    st4(st.NSa, id, sn);
  } else {
    const int ks = id - P.nnode;
    const Rec4 r = ldrec(st.SVb, ks);
    const double2 v = make_double2(r.a, r.b);
    const Stress4 s = ld4(st.SFb, ks);
    st2(st.vx0, id, v);
    st4(st.stress0, ks, s);
    if (pos_of) {
      st4(st.RKs, ks, Stress4{0.0, 0.0, 0.0, 0.0});
      st.RKe[ks] = 0.0;
    }
    vn = make_double2(0.0, 0.0);
    sn.s1 = s.s1 + 0. * (P.dt) * 0.0;
    sn.s2 = s.s2 + 0. * (P.dt) * 0.0;
    sn.s3 = s.s3 + 0. * (P.dt) * 0.0;
    sn.s4 = s.s4 + 0. * (P.dt) * 0.0;
    if (P.adapt) adapt_stress(P, sn);
    apply_bcs(P, st.bc_or_not, st.bc_info, st.bc_int, st.fs_normal, id, vn, sn);
    if (pos_of)
      strec(SAs, pos_of[id], sn.s1, sn.s2, sn.s3, sn.s4);
    else
      strec(st.SA, ks, sn.s1, sn.s2, sn.s3, sn.s4);
    st2(st.SVa, ks, vn);
    if (P.cont_density) {  // main:686-689 and the stage-1 block main:706-713 (f1rk = 0, f2rk = 1)
      const double r0 = st.rho[id], h0 = st.hsml[id], m = st.mass[id];
      st.rho0[ks] = r0;
      st.hsml0[ks] = h0;
      const double rhs = -r0 * st.divu[ks];  // density_update with the grad_u of the previous step's last sweep
      const double rn = r0 + 0. * (P.dt) * rhs;
      st.rho_w[id] = rn;
      st.mrho_w[id] = make_double2(m, rn);
      st.mor_w[id] = m / rn;
      st.RKrho[ks] = 0.0 + 1. * rhs;
      if (P.sle == 2) {
        const double rh = -(h0 / (rn * 2)) * rhs;
        st.hsml_w[id] = h0 + 0. * (P.dt) * rh;
        st.RKh[ks] = 0.0 + 1. * rh;
      }
    }
  }
  }
}

__device__ __forceinline__ double h0_of(int lo, int hi) { return __hiloint2double(hi, lo); }

// ------------------------------------------------------------------------------------------------------
// Sweep A (stress_point_update + the adapt_stress2 / BCs that follow it).
//   FROMB = false: input in format A (inside RK4 and the final interpolation of the step)
//   FROMB = true : input in format B (the SPH_shift interpolation at the start of a step, main:99-109)
//   FIRST: first sweep A of the step -> computes and stores cspm_norm (needs w and the partner's m, rho)
// Streams {partner id, (m/rho)_partner*w [, w]} and gathers ONE 16-byte (node velocity) or 32-byte (stress) record.
// ------------------------------------------------------------------------------------------------------
// UMOR: mass/rho takes at most four distinct values within a species (lattice set-ups: interior, edge and corner
// particles; the host checks it at upload). The fill pass then stores the partner's class in the two top bits of its
// id instead of the 8-byte product (m/rho)*w, and the sweep rebuilds the product from the class and the streamed
// weight -- bit-identical -- streaming 8 bytes per entry instead of 12.
struct MorPalette {
  double v0, v1, v2, v3;
};
__device__ __forceinline__ double palette_value(const MorPalette &p, unsigned c) {
  return c == 0 ? p.v0 : (c == 1 ? p.v1 : (c == 2 ? p.v2 : p.v3));
}
template <bool FIRST, bool FROMB, bool UMOR>
__global__ void __launch_bounds__(SWEEP_T, SPSPH_ASP_MINB)
k_sweep_a_sp(DevParams P, SlotMap M, const int *__restrict__ order_s, ListPtrs L, const int *__restrict__ n0,
             StatePtrs st, int do_adapt, int do_bc, MorPalette mor_u) {
  const int k0 = blockIdx.x * blockDim.x + threadIdx.x;  // species-sorted index of the stress particle
  if ((k0 & ~31) >= M.ns) return;                         // whole warp past the end
  const bool live = k0 < M.ns;
  const int k = live ? k0 : 0;
  const int t = M.nnp + k0;
  const int id = order_s[k];
  const int ks = id - P.nnode;
  const int cnt = live ? n0[t] : 0;
  const int wrows = warp_max_i(cnt);
  // the particle's own velocity survives only where cspm_norm == 0 (no velocity-particle partner): read it there
  double2 v;
  Stress4 s;
  if (FROMB) {
    s = ld4(st.SFbr, ks);
  } else {
    const Rec4 r = ldrec(st.SA, ks);
    s = Stress4{r.a, r.b, r.c, r.d};
  }
  double vtx = 0.0, vty = 0.0, nrm = 0.0;
  {
    constexpr int NARR = UMOR ? 2 : (FIRST ? 4 : 3);
    __shared__ __align__(16) int smem[(SWEEP_T / 32) * ELL_SMEM_G(NARR, A_NG, A_GR)];
    const int *arrs[4] = {L.idx0, UMOR ? reinterpret_cast<const int *>(L.w0) : L.h0lo, L.h0hi,
                          reinterpret_cast<const int *>(L.w0)};
    const double *__restrict__ NAv = st.NA;
    const Rec4 *__restrict__ NBv = st.NBr;
    // FIRST also needs the partner's mass and density (cspm_norm, main:433): they ride in the same 32-byte record
    // (format B) or come with one extra 16-byte gather (format A)
    struct RecN {
      double2 v, mr;
    };
    ell_stream<NARR, A_NG, RecN, A_GR, ASP_SUB, UMOR>(
        arrs, (size_t)L.off0[t / SLICE], wrows, cnt, smem + (threadIdx.x >> 5) * ELL_SMEM_G(NARR, A_NG, A_GR),
        [&](int q) {
          const int qq = (q < 0 || q >= P.nnode) ? 0 : q;
          RecN o;
          if (FROMB) {
            const Rec4 r = ldrec(NBv, qq);
            o.v = make_double2(r.a, r.b);
            o.mr = make_double2(r.c, r.d);
          } else {
            o.v = ld2(NAv, qq);
            o.mr = FIRST ? st.mrho[qq] : make_double2(0.0, 0.0);
          }
          return o;
        },
        [&](const int(&q)[ASP_SUB], const int(&pay)[NARR - 1][ASP_SUB], const RecN(&r)[ASP_SUB], int nvalid) {
          if (UMOR && !FIRST) {
            // full group without a wall partner (the common case): straight-line code, no per-entry predicates
            bool plain = nvalid >= ASP_SUB;
#pragma unroll
            for (int u = 0; u < ASP_SUB; ++u) plain = plain && ((q[u] & QID_MASK) < P.nnode);
            if (plain) {
#pragma unroll
              for (int u = 0; u < ASP_SUB; ++u) {
                const double h2 = palette_value(mor_u, (unsigned)q[u] >> QCLASS_SHIFT) * (double)__int_as_float(pay[0][u]);
                vtx = vtx + r[u].v.x * h2;
                vty = vty + r[u].v.y * h2;
              }
              return;
            }
          }
#pragma unroll
          for (int u = 0; u < ASP_SUB; ++u) {
            const int qid = UMOR ? (q[u] & QID_MASK) : q[u];
            const bool ok = (u < nvalid) && (qid < P.nnode);  // dummy partners (type 9) take no part
            double h2;  // (mass(i)/rho(i))*w, main:431
            if constexpr (UMOR)
              h2 = palette_value(mor_u, (unsigned)q[u] >> QCLASS_SHIFT) * (double)__int_as_float(pay[0][u]);
            else
              h2 = h0_of(pay[0][u], pay[1][u]);
            if (FIRST && P.cont_density) {  // the density moves: the streamed product of the fill pass is stale
              const double rq = ok ? r[u].mr.y : 1.0;
              h2 = div_rn(r[u].mr.x, rq, __drcp_rn(rq)) * (double)__int_as_float(pay[NARR - 2][u]);
            }
            const double tx = vtx + r[u].v.x * h2, ty = vty + r[u].v.y * h2;
            vtx = ok ? tx : vtx;
            vty = ok ? ty : vty;
            if (FIRST) {
              const double wd = (double)__int_as_float(pay[NARR - 2][u]);
              const double rq = ok ? r[u].mr.y : 1.0;
              const double nn = nrm + div_rn(wd * r[u].mr.x, rq, __drcp_rn(rq));
              nrm = ok ? nn : nrm;
            }
          }
        });
  }
  if (!live) return;
  if (FIRST)
    st.norm[id] = nrm;
  else
    nrm = st.norm[id];
  if (nrm != 0) {  // one reciprocal, exactly rounded quotients (div_rn == IEEE division)
    const double rn = __drcp_rn(nrm);
    v.x = div_rn(vtx, nrm, rn);
    v.y = div_rn(vty, nrm, rn);
  } else if (FROMB) {
    const Rec4 r = ldrec(st.SVbr, ks);
    v = make_double2(r.a, r.b);
  } else {
    v = ld2(st.SVa, ks);
  }
  if (do_adapt) adapt_stress(P, s);
  if (do_bc) apply_bcs(P, st.bc_or_not, st.bc_info, st.bc_int, st.fs_normal, id, v, s);
  strec(st.SVb, ks, v.x, v.y, st.mor[id], 0.0);
  st4(st.SFb, ks, s);
  const double rr = st.rho[id];
  const double r2 = rr * rr, rr2 = __drcp_rn(r2);
  strec(st.SB, ks, div_rn(s.s1, r2, rr2), div_rn(s.s2, r2, rr2), div_rn(s.s3, r2, rr2), st.mass[id]);
}

template <bool FIRST, bool FROMB, bool EPSP, bool UMOR>
__global__ void __launch_bounds__(SWEEP_T, SPSPH_AN_MINB)
k_sweep_a_node(DevParams P, SlotMap M, const int *__restrict__ order_n, ListPtrs L, const int *__restrict__ n0,
               StatePtrs st, int do_adapt, int do_bc, MorPalette mor_u) {
  const int k0 = blockIdx.x * blockDim.x + threadIdx.x;
  if ((k0 & ~31) >= M.nn) return;  // whole warp past the end
  const bool live = k0 < M.nn;
  const int k = live ? k0 : 0;
  const int t = k0;
  const int id = order_n[k];
  const int cnt = live ? n0[t] : 0;
  const int wrows = warp_max_i(cnt);
  // the particle's own stress survives only where cspm_norm == 0 (no stress-particle partner): read it there
  double2 v;
  Stress4 s;
  if (FROMB) {
    const Rec4 r = ldrec(st.NBr, id);
    v = make_double2(r.a, r.b);
  } else {
    v = ld2(st.NA, id);
  }
  double t1 = 0.0, t2 = 0.0, t3 = 0.0, t4 = 0.0, te = 0.0, nrm = 0.0, trho = 0.0;
  {
    struct RecS {
      Rec4 s;
      double ep;
      double2 mr;
    };
    constexpr int NARR = UMOR ? 2 : (FIRST ? 4 : 3);
    __shared__ __align__(16) int smem[(SWEEP_T / 32) * ELL_SMEM_G(NARR, A_NG, A_GR)];
    const int *arrs[4] = {L.idx0, UMOR ? reinterpret_cast<const int *>(L.w0) : L.h0lo, L.h0hi,
                          reinterpret_cast<const int *>(L.w0)};
    ell_stream<NARR, A_NG, RecS, A_GR, AN_SUB, UMOR>(
        arrs, (size_t)L.off0[t / SLICE], wrows, cnt, smem + (threadIdx.x >> 5) * ELL_SMEM_G(NARR, A_NG, A_GR),
        [&](int q) {
          const int qs = (q < 0 || q >= P.ntotal) ? 0 : q - P.nnode;
          RecS r;
          r.s = FROMB ? ld256(st.SFbr + 4 * (size_t)qs) : ldrec(st.SA, qs);
          r.ep = EPSP ? st.epsp[qs + P.nnode] : 0.0;
          r.mr = FIRST ? st.mrho[qs + P.nnode] : make_double2(0.0, 0.0);
          return r;
        },
        [&](const int(&q)[AN_SUB], const int(&pay)[NARR - 1][AN_SUB], const RecS(&r)[AN_SUB], int nvalid) {
          if (UMOR && !FIRST) {
            // full group without a wall partner (the common case): straight-line code, no per-entry predicates
            bool plain = nvalid >= AN_SUB;
#pragma unroll
            for (int u = 0; u < AN_SUB; ++u) plain = plain && ((q[u] & QID_MASK) < P.ntotal);
            if (plain) {
#pragma unroll
              for (int u = 0; u < AN_SUB; ++u) {
                const double h1 = palette_value(mor_u, (unsigned)q[u] >> QCLASS_SHIFT) * (double)__int_as_float(pay[0][u]);
                t1 = t1 + r[u].s.a * h1;
                t2 = t2 + r[u].s.b * h1;
                t3 = t3 + r[u].s.c * h1;
                t4 = t4 + r[u].s.d * h1;
                if (EPSP) te = te + r[u].ep * h1;
              }
              return;
            }
          }
#pragma unroll
          for (int u = 0; u < AN_SUB; ++u) {
            const int qid = UMOR ? (q[u] & QID_MASK) : q[u];
            const bool ok = (u < nvalid) && (qid < P.ntotal);  // dummy partners (type 6) take no part
            double h1;  // (mass(j)/rho(j))*w, main:430
            if constexpr (UMOR)
              h1 = palette_value(mor_u, (unsigned)q[u] >> QCLASS_SHIFT) * (double)__int_as_float(pay[0][u]);
            else
              h1 = h0_of(pay[0][u], pay[1][u]);
            if (FIRST && P.cont_density) {  // the density moves: recompute the factor, and interpolate rho (main:437)
              const double rq = ok ? r[u].mr.y : 1.0;
              h1 = div_rn(r[u].mr.x, rq, __drcp_rn(rq)) * (double)__int_as_float(pay[NARR - 2][u]);
              const double nr = trho + r[u].mr.y * h1;
              trho = ok ? nr : trho;
            }
            const double n1 = t1 + r[u].s.a * h1, n2 = t2 + r[u].s.b * h1, n3 = t3 + r[u].s.c * h1,
                         n4 = t4 + r[u].s.d * h1;
            t1 = ok ? n1 : t1;
            t2 = ok ? n2 : t2;
            t3 = ok ? n3 : t3;
            t4 = ok ? n4 : t4;
            if (EPSP) {
              const double ne = te + r[u].ep * h1;
              te = ok ? ne : te;
            }
            if (FIRST) {
              const double wd = (double)__int_as_float(pay[NARR - 2][u]);
              const double rq = ok ? r[u].mr.y : 1.0;
              const double nn = nrm + div_rn(wd * r[u].mr.x, rq, __drcp_rn(rq));
              nrm = ok ? nn : nrm;
            }
          }
        });
  }
  if (!live) return;
  if (FIRST)
    st.norm[id] = nrm;
  else
    nrm = st.norm[id];
  if (nrm != 0) {
    const double rn = __drcp_rn(nrm);
    s.s1 = div_rn(t1, nrm, rn);
    s.s2 = div_rn(t2, nrm, rn);
    s.s3 = div_rn(t3, nrm, rn);
    s.s4 = div_rn(t4, nrm, rn);
    if (EPSP) st.epsp[id] = div_rn(te, nrm, rn);
  } else {
    s = FROMB ? ld4(st.NSbr, id) : ld4(st.NSa, id);
    v.x = 0;
    v.y = 0;
  }
  if (do_adapt) adapt_stress(P, s);
  if (do_bc) apply_bcs(P, st.bc_or_not, st.bc_info, st.bc_int, st.fs_normal, id, v, s);
  double rnode = st.rho[id];
  if (FIRST && P.cont_density) {  // rho(1:nnode) = rho_temp/cspm_norm, main:464-465 (unconditional)
    rnode = trho / nrm;
    st.rho_new[id] = rnode;  // committed after the sweep: the stress-particle side still reads the old value
  }
  strec(st.NB, id, v.x, v.y, st.mass[id], rnode);
  st4(st.NSb, id, s);
}

// cont_density: the interpolated density of the velocity particles becomes current once both sides of sweep A are done
__global__ void k_commit_node_rho(DevParams P, SlotMap M, const int *__restrict__ order_n, StatePtrs st) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= M.nn) return;
  const int id = order_n[k];
  const double r = st.rho_new[id], m = st.mass[id];
  st.rho_w[id] = r;
  st.mrho_w[id] = make_double2(m, r);
  st.mor_w[id] = m / r;
}

// ------------------------------------------------------------------------------------------------------
// Sweep B (get_derivatives + plastic_terms + gravity/damping + Jaumann terms + RK4 stage accumulation +
// next-stage predictor, or the final RK4 update when `last`): format B -> format A.
// ------------------------------------------------------------------------------------------------------
template <bool FIRST>
__global__ void __launch_bounds__(SWEEP_T, SPSPH_BSP_MINB)
k_sweep_b_sp(DevParams P, SlotMap M, const int *__restrict__ order_s, ListPtrs L, const int *__restrict__ n0,
             StatePtrs st, double f1next, double f2, int stage_flags, double f2next) {
  const int last = stage_flags & 1;      // last RK4 stage: final update instead of the next predictor
  const bool stage1 = stage_flags & 2;   // first RK4 stage: the accumulators start from zero (not read, main:686)
  const int k0 = blockIdx.x * blockDim.x + threadIdx.x;
  if ((k0 & ~31) >= M.ns) return;  // whole warp past the end
  const bool live = k0 < M.ns;
  const int k = live ? k0 : 0;
  const int t = M.nnp + k0;
  const int id = order_s[k];
  const int ks = id - P.nnode;
  const int cnt = live ? n0[t] : 0;
  const int wrows = warp_max_i(cnt);
  const Rec4 selfv = ldrec(st.SVb, ks);
  const double2 vp = make_double2(selfv.a, selfv.b);
  const Stress4 sp_ = ld4(st.SFb, ks);
  // own position: the CSPM matrix pass needs it for every entry, the wall-partner terms only next to a wall (there
  // it is read where it is used instead of by every stress particle in every stage)
  double2 xp = make_double2(0.0, 0.0);
  if (FIRST && P.cspm) xp = ld2(st.x, id);
  double ae1 = 0.0, ae2 = 0.0, ae3 = 0.0, ae4 = 0.0, ae5 = 1.0;
  double g11 = 0.0, g12 = 0.0, g21 = 0.0, g22 = 0.0;  // grad1_tmp(d,k): d velocity component, k direction
    auto entry_slow = [&](int q, int a1, int a2, const Rec4 &r) {
      const double gx = (double)__int_as_float(a1), gy = (double)__int_as_float(a2);
      if (q < P.nnode) {  // type 1: q is the node; r = {vx, vy, m, rho}
        const double h1 = gx * r.c / r.d;
        const double h2 = gy * r.c / r.d;
        g11 = g11 + (r.a - vp.x) * h1;
        g12 = g12 + (r.a - vp.x) * h2;
        g21 = g21 + (r.b - vp.y) * h1;
        g22 = g22 + (r.b - vp.y) * h2;
        if (FIRST && P.cspm) {
          const double2 xq = ld2(st.x, q);
          ae1 = ae1 + (xq.x - xp.x) * h1;
          ae2 = ae2 + (xq.y - xp.y) * h1;
          ae3 = ae3 + (xq.x - xp.x) * h2;
          ae4 = ae4 + (xq.y - xp.y) * h2;
        }
      } else {  // type 9: q is a dummy wall particle (no-slip mirror velocity), main:552-575
        const double beta_max = 1.5, vel_wall = 0.0;
        const double wall = (double)st.wallpos[q];
        const double2 xq = ld2(st.x, q);
        const double2 xo = (FIRST && P.cspm) ? xp : ld2(st.x, id);
        double da, db;
        if (st.horiz[q] == 1.f) {
          da = fabs(xo.y - wall);
          db = fabs(xq.y - wall);
        } else {
          da = fabs(xo.x - wall);
          db = fabs(xq.x - wall);
        }
        const double bq = 1 + (db / da);
        const double beta = (bq < beta_max) ? bq : beta_max;
        const double dvx = vp.x * (1 - beta) + beta * vel_wall;
        const double dvy = vp.y * (1 - beta) + beta * vel_wall;
        const double mq = st.mass[q], rq = st.rho[q];
        const double h1 = gx * mq / rq;
        const double h2 = gy * mq / rq;
        g11 = g11 + (vp.x - dvx) * h1;
        g12 = g12 + (vp.x - dvx) * h2;
        g21 = g21 + (vp.y - dvy) * h1;
        g22 = g22 + (vp.y - dvy) * h2;
      }
    };
  {
    __shared__ __align__(16) int smem[(SWEEP_T / 32) * ELL_SMEM(3, ELL_NG)];
    const int *arrs[3] = {L.idx0, reinterpret_cast<const int *>(L.gx0), reinterpret_cast<const int *>(L.gy0)};
    ell_stream<3, ELL_NG, Rec4, ELL_GROUP, BSP_SUB>(
        arrs, (size_t)L.off0[t / SLICE], wrows, cnt, smem + (threadIdx.x >> 5) * ELL_SMEM(3, ELL_NG),
        [&](int q) { return ldrec(st.NB, (q < 0 || q >= P.nnode) ? 0 : q); },
        [&](const int(&q)[BSP_SUB], const int(&pay)[2][BSP_SUB], const Rec4(&r)[BSP_SUB], int nvalid) {
          bool special = false;  // wall partner in this group (type 9), or the once-per-step CSPM matrix pass
#pragma unroll
          for (int u = 0; u < BSP_SUB; ++u) special |= (u < nvalid) && (q[u] >= P.nnode);
          if ((FIRST && P.cspm) || __any_sync(0xffffffffu, special)) {
#pragma unroll
            for (int u = 0; u < BSP_SUB; ++u)
              if (u < nvalid) entry_slow(q[u], pay[0][u], pay[1][u], r[u]);
            return;
          }
          // branch-free path: the four entries' division chains are independent and interleave
          double h1[BSP_SUB], h2[BSP_SUB];
#pragma unroll
          for (int u = 0; u < BSP_SUB; ++u) {
            const double gx = (double)__int_as_float(pay[0][u]), gy = (double)__int_as_float(pay[1][u]);
            const double rr = __drcp_rn(r[u].d);
            h1[u] = div_rn(gx * r[u].c, r[u].d, rr);  // dwdx*mass(i)/rho(i), main:514
            h2[u] = div_rn(gy * r[u].c, r[u].d, rr);
          }
          if (nvalid >= BSP_SUB) {  // full group: no per-entry predicates
#pragma unroll
            for (int u = 0; u < BSP_SUB; ++u) {
              const double dvx = r[u].a - vp.x, dvy = r[u].b - vp.y;
              g11 = g11 + dvx * h1[u];
              g12 = g12 + dvx * h2[u];
              g21 = g21 + dvy * h1[u];
              g22 = g22 + dvy * h2[u];
            }
            return;
          }
#pragma unroll
          for (int u = 0; u < BSP_SUB; ++u) {
            const bool ok = u < nvalid;
            const double dvx = r[u].a - vp.x, dvy = r[u].b - vp.y;
            const double n11 = g11 + dvx * h1[u], n12 = g12 + dvx * h2[u], n21 = g21 + dvy * h1[u],
                         n22 = g22 + dvy * h2[u];
            g11 = ok ? n11 : g11;
            g12 = ok ? n12 : g12;
            g21 = ok ? n21 : g21;
            g22 = ok ? n22 : g22;
          }
        });
  }
  if (!live) return;
  if (P.cspm) {
    double *AEp = st.AE + 5 * (size_t)id;
    if (FIRST) {
      ae5 = ae1 * ae4 - ae2 * ae3;
      if (fabs(ae5) < P.ae_thr) {
        ae5 = 1;
        ae1 = 1;
        ae2 = 0;
        ae3 = 0;
        ae4 = 1;
      } else {
        ae5 = 1 / ae5;
      }
      AEp[0] = ae1;
      AEp[1] = ae2;
      AEp[2] = ae3;
      AEp[3] = ae4;
      AEp[4] = ae5;
    } else {
      ae1 = AEp[0];
      ae2 = AEp[1];
      ae3 = AEp[2];
      ae4 = AEp[3];
      ae5 = AEp[4];
    }
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
  plastic_terms(P, sp_, g11, g12, g21, g22, st.epsp + id, st.fdp + id, Gs, der1);
  const double rke = (stage1 ? 0.0 : st.RKe[ks]) + der1 * f2;
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
  Stress4 rk = stage1 ? Stress4{0.0, 0.0, 0.0, 0.0} : ld4(st.RKs, ks);
  rk.s1 = rk.s1 + f2 * r1;
  rk.s2 = rk.s2 + f2 * r2;
  rk.s3 = rk.s3 + f2 * r3;
  rk.s4 = rk.s4 + f2 * r4;
  const Stress4 s0 = ld4(st.stress0, ks);
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
    st.epsp[id] = st.epsp[id] + P.dt * (rke / 6);
  }
  if (P.adapt) adapt_stress(P, sn);
  double2 vn = vp;
  apply_bcs(P, st.bc_or_not, st.bc_info, st.bc_int, st.fs_normal, id, vn, sn);
  strec(st.SA, ks, sn.s1, sn.s2, sn.s3, sn.s4);
  st2(st.SVa, ks, vn);
  if (P.cont_density) {
    // the next stage's density block (main:706-713: density_update reads the grad_u just computed) or the final
    // update (main:792-797); RK_rho takes the next stage's RHS at once, it does not change during that stage
    const double div = g11 + g22;
    st.divu[ks] = div;
    const double m = st.mass[id];
    double rn;
    if (!last) {
      const double rhs = -st.rho[id] * div;
      rn = st.rho0[ks] + f1next * (P.dt) * rhs;
      st.RKrho[ks] = st.RKrho[ks] + f2next * rhs;
      if (P.sle == 2) {
        const double h0 = st.hsml0[ks];
        const double rh = -(h0 / (rn * 2)) * rhs;
        st.hsml_w[id] = h0 + f1next * (P.dt) * rh;
        st.RKh[ks] = st.RKh[ks] + f2next * rh;
      }
    } else {
      rn = st.rho0[ks] + (P.dt / 6.) * st.RKrho[ks];
      if (P.sle == 2) st.hsml_w[id] = st.hsml0[ks] + (P.dt / 6.) * st.RKh[ks];
    }
    st.rho_w[id] = rn;
    st.mrho_w[id] = make_double2(m, rn);
    st.mor_w[id] = m / rn;
  }
}


// artificial_viscosity, main:826-904 (fp32 locals and accumulators, list order): one thread per node over its
// node-node list; xij, yij, h were rounded to fp32 when the list was built, so ONE 32-byte gather per entry.
// UH: one smoothing length for every particle (host-checked at upload): h = 0.5*(h_i + h_j) is that constant and
// is not streamed (24 instead of 28 bytes per entry).
template <bool UH>
__global__ void __launch_bounds__(SWEEP_T, SPSPH_AV_MINB)
k_artvisc(DevParams P, SlotMap M, const int *__restrict__ order_n, ListPtrs L, const int *__restrict__ n1, StatePtrs st,
          float h_u) {
  const int k0 = blockIdx.x * blockDim.x + threadIdx.x;
  if ((k0 & ~31) >= M.nn) return;  // whole warp past the end
  const bool live = k0 < M.nn;
  const int k = live ? k0 : 0;
  const int t = k0;
  const int id = order_n[k];
  const int cntc = live ? n1[t] : 0;
  const int wrowsC = warp_max_i(cntc);
  const Rec4 self = ldrec(st.NB, id);  // {vx, vy, m, rho}
  const double2 vp = make_double2(self.a, self.b);
  const double rp = self.d;
  float acc1 = 0.f, acc2 = 0.f;
  constexpr int NG = 3;
  constexpr int NARR = UH ? 5 : 6;
  __shared__ __align__(16) int smem[(SWEEP_T / 32) * ELL_SMEM(NARR, NG)];
  const int *arrs[6] = {L.idxC, reinterpret_cast<const int *>(L.gxC), reinterpret_cast<const int *>(L.gyC),
                        reinterpret_cast<const int *>(L.xC), reinterpret_cast<const int *>(L.yC),
                        reinterpret_cast<const int *>(L.hC)};
  ell_stream<NARR, NG, Rec4, ELL_GROUP, AV_SUB>(
      arrs, (size_t)L.offC[t / SLICE], wrowsC, cntc, smem + (threadIdx.x >> 5) * ELL_SMEM(NARR, NG),
      [&](int q) { return ldrec(st.NB, q < 0 ? 0 : q); },
      [&](const int(&)[AV_SUB], const int(&pay)[NARR - 1][AV_SUB], const Rec4(&r)[AV_SUB], int nvalid) {
        float visc[AV_SUB];
#pragma unroll
        for (int u = 0; u < AV_SUB; ++u) {  // independent per entry: interleaves
          const float xij = __int_as_float(pay[2][u]), yij = __int_as_float(pay[3][u]);
          float h;
          if constexpr (UH)
            h = h_u;
          else
            h = __int_as_float(pay[4][u]);
          const float rho2 = (float)(0.5 * (rp + r[u].d));
          const float cs = 600.f;
          float div_u = (float)((double)xij * (vp.x - r[u].a));
          div_u = (float)((double)div_u + (double)yij * (vp.y - r[u].b));
          const float sq = sqrtf(xij * xij + yij * yij);
          const float theta = (h * div_u) / (sq * sq + 0.01f * (h * h));
          // evaluated for every entry (no divergent branch; the four division chains interleave), kept if approaching
          const double rho2d = (double)rho2;
          const double num = -P.alpha * (double)cs * (double)theta + P.beta * (double)(theta * theta);
          const float vv = (float)div_rn(num, rho2d, __drcp_rn(rho2d));
          visc[u] = (div_u < 0) ? vv : 0.f;
        }
#pragma unroll
        for (int u = 0; u < AV_SUB; ++u) {  // ordered fp32 accumulation
          const float gxf = __int_as_float(pay[0][u]), gyf = __int_as_float(pay[1][u]);
          const float a1 = (float)((double)acc1 + (double)(visc[u] * gxf) * r[u].c);
          const float a2 = (float)((double)acc2 + (double)(visc[u] * gyf) * r[u].c);
          acc1 = (u < nvalid) ? a1 : acc1;
          acc2 = (u < nvalid) ? a2 : acc2;
        }
      });
  if (!live) return;
  st2(st.av, id, make_double2((double)(-acc1), (double)(-acc2)));  // art_visc = -art_visc_temp, main:901
}

template <bool FIRST>
__global__ void __launch_bounds__(SWEEP_T, SPSPH_BN_MINB)
k_sweep_b_node(DevParams P, SlotMap M, const int *__restrict__ order_n, ListPtrs L, const int *__restrict__ n0,
               StatePtrs st, double f1next, double f2, int stage_flags) {
  const int last = stage_flags & 1;     // see k_sweep_b_sp
  const bool stage1 = stage_flags & 2;
  const int k0 = blockIdx.x * blockDim.x + threadIdx.x;
  if ((k0 & ~31) >= M.nn) return;  // whole warp past the end
  const bool live = k0 < M.nn;
  const int k = live ? k0 : 0;
  const int t = k0;
  const int id = order_n[k];
  const int cnt = live ? n0[t] : 0;
  const int wrows = warp_max_i(cnt);
  const int sl = t / SLICE;
  const Rec4 self = ldrec(st.NB, id);  // {vx, vy, m, rho}
  const double2 vp = make_double2(self.a, self.b);
  const double rp = self.d;
  const Stress4 sp_ = ld4(st.NSb, id);
  const double r2p = rp * rp, rr2p = __drcp_rn(r2p);
  const double so1 = div_rn(sp_.s1, r2p, rr2p), so2 = div_rn(sp_.s2, r2p, rr2p),
               so3 = div_rn(sp_.s3, r2p, rr2p);  // stress(1:3,i)/rho(i)**2
  double2 xp = make_double2(0.0, 0.0);
  if (FIRST && P.cspm) xp = ld2(st.x, id);
  double ae1 = 0.0, ae2 = 0.0, ae3 = 0.0, ae4 = 0.0, ae5 = 1.0;
  double a11 = 0.0, a12 = 0.0, a21 = 0.0, a22 = 0.0, a31 = 0.0, a32 = 0.0;  // grad2_tmp(s,k)
    auto entry_slow = [&](int q, int a1, int a2, const Rec4 &r) {
      const double gx = (double)__int_as_float(a1), gy = (double)__int_as_float(a2);
      double q1, q2, q3, mq;
      if (q < P.ntotal) {  // type 1: q is the stress particle; r = {s1/rho^2, s2/rho^2, s3/rho^2, m}
        q1 = r.a;
        q2 = r.b;
        q3 = r.c;
        mq = r.d;
        if (FIRST && P.cspm) {
          const double rq = st.rho[q];
          const double h1b = -gx * mq / rq;
          const double h2b = -gy * mq / rq;
          const double2 xq = ld2(st.x, q);
          ae1 = ae1 + (xq.x - xp.x) * h1b;
          ae2 = ae2 + (xq.y - xp.y) * h1b;
          ae3 = ae3 + (xq.x - xp.x) * h2b;
          ae4 = ae4 + (xq.y - xp.y) * h2b;
        }
      } else {  // type 6: dummy takes the node's stress (main:580)
        const double rq = st.rho[q];
        mq = st.mass[q];
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
    };
  {
    __shared__ __align__(16) int smem[(SWEEP_T / 32) * ELL_SMEM(3, ELL_NG)];
    const int *arrs[3] = {L.idx0, reinterpret_cast<const int *>(L.gx0), reinterpret_cast<const int *>(L.gy0)};
    ell_stream<3, ELL_NG, Rec4, ELL_GROUP, BN_SUB>(
        arrs, (size_t)L.off0[sl], wrows, cnt, smem + (threadIdx.x >> 5) * ELL_SMEM(3, ELL_NG),
        [&](int q) { return ldrec(st.SB, (q < 0 || q >= P.ntotal) ? 0 : q - P.nnode); },
        [&](const int(&q)[BN_SUB], const int(&pay)[2][BN_SUB], const Rec4(&r)[BN_SUB], int nvalid) {
          bool special = false;  // wall partner in this group (type 6), or the once-per-step CSPM matrix pass
#pragma unroll
          for (int u = 0; u < BN_SUB; ++u) special |= (u < nvalid) && (q[u] >= P.ntotal);
          if ((FIRST && P.cspm) || __any_sync(0xffffffffu, special)) {
#pragma unroll
            for (int u = 0; u < BN_SUB; ++u)
              if (u < nvalid) entry_slow(q[u], pay[0][u], pay[1][u], r[u]);
            return;
          }
          if (nvalid >= BN_SUB) {  // full group: no per-entry predicates
#pragma unroll
            for (int u = 0; u < BN_SUB; ++u) {
              const double gx = (double)__int_as_float(pay[0][u]), gy = (double)__int_as_float(pay[1][u]);
              const double c1 = so1 + r[u].a, c2 = so2 + r[u].b, c3 = so3 + r[u].c, mq = r[u].d;
              a11 = a11 - mq * (gx * c1);
              a12 = a12 - mq * (gy * c1);
              a21 = a21 - mq * (gx * c2);
              a22 = a22 - mq * (gy * c2);
              a31 = a31 - mq * (gx * c3);
              a32 = a32 - mq * (gy * c3);
            }
            return;
          }
#pragma unroll
          for (int u = 0; u < BN_SUB; ++u) {
            const bool ok = u < nvalid;
            const double gx = (double)__int_as_float(pay[0][u]), gy = (double)__int_as_float(pay[1][u]);
            const double c1 = so1 + r[u].a, c2 = so2 + r[u].b, c3 = so3 + r[u].c, mq = r[u].d;
            const double n11 = a11 - mq * (gx * c1), n12 = a12 - mq * (gy * c1);
            const double n21 = a21 - mq * (gx * c2), n22 = a22 - mq * (gy * c2);
            const double n31 = a31 - mq * (gx * c3), n32 = a32 - mq * (gy * c3);
            a11 = ok ? n11 : a11;
            a12 = ok ? n12 : a12;
            a21 = ok ? n21 : a21;
            a22 = ok ? n22 : a22;
            a31 = ok ? n31 : a31;
            a32 = ok ? n32 : a32;
          }
        });
  }
  if (!live) return;
  if (P.cspm) {
    double *AEp = st.AE + 5 * (size_t)id;
    if (FIRST) {
      ae5 = ae1 * ae4 - ae2 * ae3;
      if (fabs(ae5) < P.ae_thr) {
        ae5 = 1;
        ae1 = 1;
        ae2 = 0;
        ae3 = 0;
        ae4 = 1;
      } else {
        ae5 = 1 / ae5;
      }
      AEp[0] = ae1;
      AEp[1] = ae2;
      AEp[2] = ae3;
      AEp[3] = ae4;
      AEp[4] = ae5;
    } else {
      ae1 = AEp[0];
      ae2 = AEp[1];
      ae3 = AEp[2];
      ae4 = AEp[3];
      ae5 = AEp[4];
    }
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
  // artificial viscosity of this stage (k_artvisc), zero when alpha = beta = 0 (art_visc stays 0, main:688)
  double av1 = 0.0, av2 = 0.0;
  if (P.alpha > 0 || P.beta > 0) {
    const double2 a = ld2(st.av, id);
    av1 = a.x;
    av2 = a.y;
  }
  // f_bound (main:764): zero unless boundary_forces ran; art_force: zero unless art_stress = T
  const double2 fb = st.has_fbound ? ld2(st.fbound, id) : make_double2(0.0, 0.0);
  const double2 af = st.has_aforce ? ld2(st.aforce, id) : make_double2(0.0, 0.0);
  const double r1 = -dv1 + sg1 + av1 + fb.x + af.x;
  const double r2 = -dv2 + sg2 + av2 + fb.y + af.y;
  double2 rk = stage1 ? make_double2(0.0, 0.0) : ld2(st.RKv, id);
  rk.x = rk.x + f2 * r1;
  rk.y = rk.y + f2 * r2;
  const double2 v0 = ld2(st.vel0, id);
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
  st2(st.NA, id, vn);
  st4(st.NSa, id, sn);
}


// ------------------------------------------------------------------------------------------------------
// Standard SPH mode (SP_SPH = F): stress_point_update degenerates to copies between a velocity particle and the
// stress particle that shadows it (main:472-480); no adapt_stress2 / BCs follow it (main:720-723,132-135).
// ------------------------------------------------------------------------------------------------------
__global__ void k_sweep_a_std(DevParams P, StatePtrs st, LocalList LL) {
  SPSPH_FOR_LOCAL(LL, kk, id) {
  if (id >= P.ntotal) continue;
  if (id < P.nnode) {  // stress(:,i) = stress(:,nnode+i); Internal_Vars(1,i) = Internal_Vars(1,nnode+i)
    const double2 v = ld2(st.NA, id);
    const Rec4 s = ldrec(st.SA, id);
    double rn = st.rho[id];
    if (P.cont_density) {  // rho(1:nnode) = rho(nnode+1:ntotal), main:478
      rn = st.rho[P.nnode + id];
      st.rho_w[id] = rn;
      st.mrho_w[id] = make_double2(st.mass[id], rn);
      st.mor_w[id] = st.mass[id] / rn;
    }
    strec(st.NB, id, v.x, v.y, st.mass[id], rn);
    st4(st.NSb, id, Stress4{s.a, s.b, s.c, s.d});
    st.epsp[id] = st.epsp[P.nnode + id];
  } else {             // vel(:,nnode+i) = vel(:,i)
    const int ks = id - P.nnode;
    const double2 v = ld2(st.NA, ks);
    const Rec4 s = ldrec(st.SA, ks);
    strec(st.SVb, ks, v.x, v.y, st.mor[id], 0.0);
    st4(st.SFb, ks, Stress4{s.a, s.b, s.c, s.d});
    const double rr = st.rho[id];
    const double r2 = rr * rr;
    strec(st.SB, ks, s.a / r2, s.b / r2, s.c / r2, st.mass[id]);
  }
  }
}
// x(:,nnode+1:ntotal) = x(:,1:nnode), main:166
__global__ void k_sp_follow(DevParams P, double *__restrict__ x, LocalList LL) {
  SPSPH_FOR_LOCAL(LL, kk, i) {
    if (i < P.nnode) st2(x, P.nnode + i, ld2(x, i));
  }
}

// ------------------------------------------------------------------------------------------------------
// boundary_forces, main:1039-1165 (branch test == 2): repulsion of a velocity particle by the innermost layer of
// wall particles; fp32 locals, fp64 accumulation in ascending wall-particle number. Only wall particles closer
// than 0.75*dx contribute a non-zero term, so the brute-force double loop becomes a 3x3-cell query.
// ------------------------------------------------------------------------------------------------------
__global__ void k_bound_force(DevParams P, SlotMap M, const GridInfo *__restrict__ G, SortArrays S, int ndummy2,
                              double *__restrict__ fbound) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= M.nn) return;
  const int id = S.order[0][k];
  const int c = S.cell[0][k];
  double f1 = 0.0, f2acc = 0.0;
  if (c >= 0) {
    const double2 pp = S.pos[0][k];
    const double hp = S.h[0][k];
    const int ndx = G->ndivx[0], ndy = G->ndivx[1];
    const int cy = c / ndx, cx = c - cy * ndx;
    constexpr int MAXC = 16;
    int cid[MAXC];
    double tx[MAXC], ty[MAXC];
    int n = 0;
    const float d0 = (float)(P.dx / 2.);
    const float cs = 20.f;
    for (int jy = max(cy - 1, 0); jy <= min(cy + 1, ndy - 1); ++jy)
      for (int jx = max(cx - 1, 0); jx <= min(cx + 1, ndx - 1); ++jx) {
        const int cq = jy * ndx + jx;
        for (int q = S.start[2][cq]; q < S.start[2][cq + 1]; ++q) {
          const int j = S.order[2][q];
          if (j >= P.ntotal + ndummy2) continue;  // only the innermost wall layer (main:1119-1121)
          const double2 pq = S.pos[2][q];
          const float r0 = (float)(pp.x - pq.x), r1 = (float)(pp.y - pq.y);
          const float r2 = sqrtf(r0 * r0 + r1 * r1);
          if (!(r2 > 0 && r2 < 1.5f * d0)) continue;  // f2 = 0: the term is (+-)0 and leaves the sum unchanged
          const float f2 = 1 - (r2 / (1.5f * d0));
          const float h = (float)(0.5 * (hp + S.h[2][q]));
          const float rb = r2 / (0.75f * h);
          float f;
          if (0 < rb && rb <= 2.f / 3.f)
            f = 2.f / 3.f;
          else if (2.f / 3.f < rb && rb <= 1)
            f = 2 * rb - 1.5f * (rb * rb);
          else if (1 < rb && rb < 2)
            f = 0.5f * ((2 - rb) * (2 - rb));
          else
            f = 0.f;
          const float pre = (0.01f * (cs * cs)) * f2 * f;
          if (n < MAXC) {
            cid[n] = j;
            tx[n] = (double)(pre * (r0 / (r2 * r2)));
            ty[n] = (double)(pre * (r1 / (r2 * r2)));
            ++n;
          }
        }
      }
    // ascending wall-particle number = the reference's loop order (j outer)
    for (int a = 1; a < n; ++a) {
      const int ci = cid[a];
      const double ax = tx[a], ay = ty[a];
      int b = a - 1;
      while (b >= 0 && cid[b] > ci) {
        cid[b + 1] = cid[b];
        tx[b + 1] = tx[b];
        ty[b + 1] = ty[b];
        --b;
      }
      cid[b + 1] = ci;
      tx[b + 1] = ax;
      ty[b + 1] = ay;
    }
    for (int a = 0; a < n; ++a) {
      f1 = f1 + tx[a];
      f2acc = f2acc + ty[a];
    }
  }
  st2(fbound, id, make_double2(f1, f2acc));
}

// ------------------------------------------------------------------------------------------------------
// Position update, main:140-182: XSPH_update (main:189-239) or the fp32 mid-velocity rule; displ.
// Velocities come from format B (the state at the end of the step); one 32-byte gather per list entry.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_move(DevParams P, SlotMap M, SortArrays So, ListPtrs L, const int *__restrict__ n1, StatePtrs st,
       double *__restrict__ x, const double *__restrict__ x00, double *__restrict__ displ) {
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (t0 >= M.nnp + M.nsp) return;
  const bool is_node = t0 < M.nnp;  // warp-uniform (nnp is a multiple of 32)
  const int k0 = is_node ? t0 : t0 - M.nnp;
  const int nlive = is_node ? M.nn : M.ns;
  if ((k0 & ~31) >= nlive) return;  // whole warp past the end
  const bool live = k0 < nlive;
  const int k = live ? k0 : 0;
  const int id = is_node ? So.order[0][k] : So.order[1][k];
  double2 vp;
  {
    const Rec4 r = is_node ? ldrec(st.NB, id) : ldrec(st.SVb, id - P.nnode);
    vp = make_double2(r.a, r.b);
  }
  double sx = 0.0, sy = 0.0;
  if (P.update_x && P.xsph) {
    const int cnt = live ? n1[t0] : 0;
    const int wrows = warp_max_i(cnt);
    const int sl = t0 / SLICE;
    __shared__ __align__(16) int smem[4 * ELL_SMEM(2, ELL_NG)];
    int *smw = smem + (threadIdx.x >> 5) * ELL_SMEM(2, ELL_NG);
    auto body = [&](const int(&)[ELL_SUB], const int(&pay)[1][ELL_SUB], const Rec4(&r)[ELL_SUB], int nvalid,
                    bool node) {
#pragma unroll
      for (int u = 0; u < ELL_SUB; ++u) {
        const double wd = (double)__int_as_float(pay[0][u]);
        const double mr = node ? (r[u].c / r[u].d) : r[u].c;  // mass(j)/rho(j)
        const double nx = sx + mr * (r[u].a - vp.x) * wd, ny = sy + mr * (r[u].b - vp.y) * wd;
        sx = (u < nvalid) ? nx : sx;
        sy = (u < nvalid) ? ny : sy;
      }
    };
    if (is_node) {
      const int *arrs[2] = {L.idxC, reinterpret_cast<const int *>(L.wC)};
      ell_stream<2, ELL_NG, Rec4>(
          arrs, (size_t)L.offC[sl], wrows, cnt, smw, [&](int q) { return ldrec(st.NB, q < 0 ? 0 : q); },
          [&](const int(&q)[ELL_SUB], const int(&pay)[1][ELL_SUB], const Rec4(&r)[ELL_SUB], int nvalid) {
            body(q, pay, r, nvalid, true);
          });
    } else {
      const int *arrs[2] = {L.idxD, reinterpret_cast<const int *>(L.wD)};
      ell_stream<2, ELL_NG, Rec4>(
          arrs, (size_t)L.offD[sl], wrows, cnt, smw, [&](int q) { return ldrec(st.SVb, q < 0 ? 0 : q - P.nnode); },
          [&](const int(&q)[ELL_SUB], const int(&pay)[1][ELL_SUB], const Rec4(&r)[ELL_SUB], int nvalid) {
            body(q, pay, r, nvalid, false);
          });
    }
  }
  if (!live) return;
  const double2 xp = ld2(x, id);
  if (P.update_x) {
    double2 xn;
    if (P.xsph) {
      const double eps = 0.5;
      xn.x = xp.x + P.dt * (vp.x + eps * sx);
      xn.y = xp.y + P.dt * (vp.y + eps * sy);
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

// shift_stress_points, main:244-368 (outside approach): one thread per node re-seats its stress particles.
__global__ void k_shift(DevParams P, const Rec4 *__restrict__ NB, double *__restrict__ x, double *__restrict__ x_10,
                        double *__restrict__ disp_10, const int *__restrict__ bc_int, const float *__restrict__ n_int,
                        LocalList LL) {
  SPSPH_FOR_LOCAL(LL, kk, i) {
  if (i >= P.nnode) continue;
  const double2 xi = ld2(x, i);
  const Rec4 rr = ldrec(NB, i);
  const double2 v = make_double2(rr.a, rr.b);
  const int k = P.nnode + i * P.npoints;  // 0-based id of this node's first stress particle
  const double dx = P.dx;
  if (P.itimestep % P.shift_update == 0) {
    if (P.vel_vector) {
      const double2 xo = ld2(x_10, i);
      const double ddx = xi.x - xo.x, ddy = xi.y - xo.y;
      double d10 = sqrt(ddx * ddx + ddy * ddy);
      d10 = d10 / dx;
      disp_10[i] = d10;
      st2(x_10, i, xi);
      const double vn = sqrt(v.x * v.x + v.y * v.y);
      const float cos_theta = (float)(v.x / vn), sin_theta = (float)(v.y / vn);
      float r1 = (float)((dx / 2) * (double)cos_theta), r2 = (float)((dx / 2) * (double)sin_theta);
      if (r1 > 0 && (double)r1 < dx / 5) r1 = (float)((double)r1 + dx / 3.);
      if (r2 > 0 && (double)r2 < dx / 5) r2 = (float)((double)r2 + dx / 3.);
      if (r1 < 0 && (double)r1 > -dx / 5) r1 = (float)((double)r1 - dx / 3.);
      if (r2 < 0 && (double)r2 > -dx / 5) r2 = (float)((double)r2 - dx / 3.);
      if (d10 > P.disp_tol) {
        st2(x, k, make_double2(xi.x + (double)r1, xi.y + (double)r2));
        st2(x, k + 1, make_double2(xi.x - (double)r1, xi.y - (double)r2));
      }
    } else {
      const double r_x = P.r_x, r_y = P.r_y;
      if (P.npoints == 1) {
        st2(x, k, make_double2(xi.x + r_x, xi.y + r_y));
      } else if (P.npoints == 2) {
        st2(x, k, make_double2(xi.x + r_x, xi.y + r_y));
        st2(x, k + 1, make_double2(xi.x - r_x, xi.y - r_x));  // sic, main:309
      } else if (P.npoints == 3) {
        st2(x, k, make_double2(xi.x, xi.y + r_y));
        st2(x, k + 1, make_double2(xi.x - r_x, xi.y - r_y));
        st2(x, k + 2, make_double2(xi.x + r_x, xi.y - r_y));
      }
    }
  }
  const float abs_vel = (float)sqrt(v.x * v.x + v.y * v.y);
  bool collapse = false;
  if (bc_int[i] == 1 && abs_vel > 0.4f) collapse = true;
  if (n_int[i] < 2) collapse = true;
  if (collapse)
    for (int q = 0; q < P.npoints && q < 3; ++q) st2(x, k + q, xi);
  }
}

// ------------------------------------------------------------------------------------------------------
// artificial_force, main:908-1016 (art_stress = T; Monaghan's artificial stress on node-node pairs, all fp64).
// The principal-axis terms R(1:3) of a node depend on that node alone, so they are evaluated once per node and
// stage (k_art_force_prep) and gathered per pair (k_art_force, ordered sums over the node-node list).
// atan / sin / cos / pow are CUDA's where the reference calls glibc's: results agree to ~1e-15 relative per term,
// not bit for bit (tests: 1e-9 relative after 40 steps against the reference executable).
// ------------------------------------------------------------------------------------------------------
__global__ void k_art_force_prep(DevParams P, SlotMap M, const int *__restrict__ order_n, StatePtrs st) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= M.nn) return;
  const int id = order_n[k];
  const Stress4 s = ld4(st.NSb, id);
  const double rp = st.rho[id];
  const double eps = (double)0.1f;
  const double s12 = s.s1 - s.s2;
  const double theta = (s12 >= (double)1e-08f) ? 0.5 * atan(2 * s.s3 / s12) : 0.0;
  const double c = cos(theta), sn = sin(theta);
  const double sg1 = (c * c) * s.s1 + 2 * c * sn * s.s3 + (sn * sn) * s.s2;
  const double sg2 = (sn * sn) * s.s1 - 2 * c * sn * s.s3 + (c * c) * s.s2;
  const double R21 = sg1 > 0 ? -eps * (sg1 / (rp * rp)) : 0.0;
  const double R22 = sg2 > 0 ? -eps * (sg2 / (rp * rp)) : 0.0;
  strec(st.RN, id, R21 * (c * c) + R22 * (sn * sn), R21 * ((c * c) + (sn * sn)), (R21 - R22) * (c * sn), 0.0);
}
__global__ void __launch_bounds__(128)
k_art_force(DevParams P, SlotMap M, const int *__restrict__ order_n, ListPtrs L, const int *__restrict__ n1,
            StatePtrs st, double w2) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M.nn) return;
  const int id = order_n[t];
  const int cnt = n1[t];
  const size_t o1 = (size_t)L.offC[t / SLICE] + (t & 31);
  const Rec4 Rp = ldrec(st.RN, id);
  const double nexp = (double)2.55f;
  double t11 = 0.0, t21 = 0.0, t31 = 0.0, t12 = 0.0, t22 = 0.0, t32 = 0.0;  // art_force_temp(istre, d)
  for (int e = 0; e < cnt; ++e) {
    const size_t a = o1 + (size_t)e * SLICE;
    const int q = L.idxC[a];
    const Rec4 Rq = ldrec(st.RN, q);
    const double mq = st.mass[q];
    const double fn = pow((double)L.wC[a] / w2, nexp);
    const double gx = (double)L.gxC[a], gy = (double)L.gyC[a];  // this node's perspective: +dwdx as pair_i, -dwdx as pair_j
    const double s1 = Rp.a + Rq.a, s2 = Rp.b + Rq.b, s3 = Rp.c + Rq.c;
    t11 = t11 + mq * (gx * fn * s1);
    t21 = t21 + mq * (gx * fn * s2);
    t31 = t31 + mq * (gx * fn * s3);
    t12 = t12 + mq * (gy * fn * s1);
    t22 = t22 + mq * (gy * fn * s2);
    t32 = t32 + mq * (gy * fn * s3);
  }
  st2(st.aforce, id, make_double2(t11 + t32, t31 + t22));
}

// ------------------------------------------------------------------------------------------------------
// get_nodes_on_free_surface, mat:1116-1411, steps 1-3 and the bc_or_not rewrite (mat:1333-1349): which particles
// lie on the free surface. Runs on demand (spsph_download) with the pair lists of the last step and the positions
// after it, exactly what the reference holds when it writes surface_points.csv. One thread per particle.
//   Step 1 sums fp32 accumulators over pair types 1, 2, 3 in the reference's traversal order; the two gather
//   lists of a particle (cross-species / same-species) are each in that order and are merged on the fly by the
//   creation-order keys (new pairs first and descending, then old pairs ascending: SURVEY App. B).
//   Step 3 is an OR over pairs, so its order is irrelevant.
// Stress-stress entries carry no stored gradient; it is re-evaluated from the positions the list was built with
// (S.pos), the same arithmetic as k_fill. Step 4 (the refined normal) only feeds apply_stress_free (ifsigman = 1,
// not supported) and is not evaluated.
// Deviation: the reference's `x**0.5` calls libm powf/pow, which are not correctly rounded (powf differs from sqrtf
// in 6e-4 of all arguments, by one ulp); the device uses the correctly rounded square root. A classification can
// differ only when a comparison of step 3 is decided by that last bit.
// ------------------------------------------------------------------------------------------------------
__global__ void k_free_surface(DevParams P, SlotMap M, SortArrays S, const int *__restrict__ pos_of, ListPtrs L,
                               const int *__restrict__ n0, const int *__restrict__ n1,
                               const GrowthRule *__restrict__ growth, const double *__restrict__ x,
                               const double *__restrict__ mass, const double *__restrict__ rho,
                               const double *__restrict__ hsml, int *__restrict__ bc_or_not,
                               int *__restrict__ covered_out /* f_int = -1 marks for k_fs_normals, or null */) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M.nnp + M.nsp) return;
  int sp, k;
  if (!slot_decode(M, t, sp, k)) return;
  const int id = S.order[sp][k];
  const int c = S.cell[sp][k];
  const int cnt0 = c < 0 ? 0 : n0[t], cnt1 = c < 0 ? 0 : n1[t];
  const int lane = t & 31, sl = t / SLICE;
  const size_t o0 = (size_t)L.off0[sl] + lane;
  const size_t o1 = (size_t)(sp == SP_NODE ? L.offC[sl] : L.offD[sl]) + lane;
  const GrowthRule gr = *growth;
  const okey_t kp = make_key(c < 0 ? 0 : c, sp, id);
  const double2 xp = ld2(x, id);
  const double hp = hsml[id];
  const double2 pp_old = S.pos[sp][k];
  const KernelConsts K = kernel_consts(P, S.h[sp][k]);
  auto species = [&](int q) { return q < P.nnode ? SP_NODE : (q < P.ntotal ? SP_STRESS : SP_DUMMY); };
  auto key_of = [&](int q) {
    const int sq = species(q);
    return make_key(S.cell[sq][pos_of[q]], sq, q);
  };
  // traversal precedence of an entry: new pairs first (descending key), then old pairs (ascending key)
  auto before = [&](okey_t ka, okey_t kb) {
    const bool oa = pair_is_old(gr, kp, ka), ob = pair_is_old(gr, kp, kb);
    if (oa != ob) return !oa;
    return oa ? ka < kb : ka > kb;
  };
  float A1 = 0.f, A2 = 0.f, A3 = 0.f, A4 = 0.f, f1 = 0.f, f2 = 0.f;
  auto step1 = [&](int q, float gx, float gy) {  // gx, gy: gradient from this particle's perspective
    const double mq = mass[q], rq = rho[q];
    const float h1 = (float)(mq * (double)gx / rq), h2 = (float)(mq * (double)gy / rq);
    const double2 xq = ld2(x, q);
    A1 = (float)((double)A1 + (xq.x - xp.x) * (double)h1);
    f1 = f1 + h1;
    A2 = (float)((double)A2 + (xq.y - xp.y) * (double)h1);
    A3 = (float)((double)A3 + (xq.x - xp.x) * (double)h2);
    A4 = (float)((double)A4 + (xq.y - xp.y) * (double)h2);
    f2 = f2 + h2;
  };
  auto entry1 = [&](int e, int &q, float &gx, float &gy) {  // same-species entry e
    const size_t a = o1 + (size_t)e * SLICE;
    if (sp == SP_NODE) {
      q = L.idxC[a];
      gx = L.gxC[a];
      gy = L.gyC[a];
    } else {
      q = L.idxD[a];
      const double2 pq = S.pos[SP_STRESS][pos_of[q]];
      const double hq = S.h[SP_STRESS][pos_of[q]];
      const double dx = pp_old.x - pq.x, dy = pp_old.y - pq.y;
      double d2 = dx * dx;
      d2 = d2 + dy * dy;
      const double mh = (K.h + hq) / 2.;
      const double r = sqrt(d2);
      double w, gxd = 0.0, gyd = 0.0;
      if (P.skf == 1 && mh == K.h)
        sph_kernel_fast<true>(K, r, dx, dy, w, gxd, gyd);
      else
        sph_kernel(P, r, dx, dy, mh, w, gxd, gyd);
      gx = (float)gxd;
      gy = (float)gyd;
    }
  };
  // Step 1: merged traversal of the type-1 entries of list 0 and all entries of list C / D
  {
    int e0 = 0, e1 = 0;
    auto skip_walls = [&]() {
      while (e0 < cnt0 && (L.idx0[o0 + (size_t)e0 * SLICE] & QID_MASK) >= P.ntotal) ++e0;
    };
    skip_walls();
    while (e0 < cnt0 || e1 < cnt1) {
      bool take0;
      if (e0 >= cnt0)
        take0 = false;
      else if (e1 >= cnt1)
        take0 = true;
      else {
        const int q0 = L.idx0[o0 + (size_t)e0 * SLICE] & QID_MASK;
        const int q1 = sp == SP_NODE ? L.idxC[o1 + (size_t)e1 * SLICE] : L.idxD[o1 + (size_t)e1 * SLICE];
        take0 = before(key_of(q0), key_of(q1));
      }
      if (take0) {
        const size_t a = o0 + (size_t)e0 * SLICE;
        const float gx = L.gx0[a], gy = L.gy0[a];  // reference orientation: pair_i = stress particle
        if (sp == SP_STRESS)
          step1(L.idx0[a] & QID_MASK, gx, gy);
        else
          step1(L.idx0[a] & QID_MASK, -gx, -gy);
        ++e0;
        skip_walls();
      } else {
        int q;
        float gx, gy;
        entry1(e1, q, gx, gy);
        step1(q, gx, gy);
        ++e1;
      }
    }
  }
  // Step 2: first approximation of the normal, scan point and tangent
  if (fabsf(A1) <= 1.e-8f) A1 = 0.f;
  if (fabsf(A2) <= 1.e-8f) A2 = 0.f;
  if (fabsf(A3) <= 1.e-8f) A3 = 0.f;
  if (fabsf(A4) <= 1.e-8f) A4 = 0.f;
  const float v1 = -(A1 * f1 + A2 * f2), v2 = -(A3 * f1 + A4 * f2);
  const float v3 = __fsqrt_rn(v1 * v1 + v2 * v2);
  const double nx = (double)(v1 / v3), ny = (double)(v2 / v3);
  const float tt1 = (float)(xp.x + hp * nx), tt2 = (float)(xp.y + hp * ny);
  const float tau1 = (float)(-ny), tau2 = (float)nx;
  const float limit = (float)((double)1.41421354f * hp);
  // Step 3: is any same-species or wall partner inside the scan region?
  bool covered = false;
  auto scan = [&](int q) {
    const double2 xq = ld2(x, q);
    const double dx = xq.x - xp.x, dy = xq.y - xp.y;
    const float dist = (float)sqrt(dx * dx + dy * dy);
    const float xt1 = (float)(xq.x - (double)tt1), xt2 = (float)(xq.y - (double)tt2);
    const float xt_norm = __fsqrt_rn(xt1 * xt1 + xt2 * xt2);
    const float prod_scal = (float)(fabs(nx * (double)xt1 + ny * (double)xt2) + (double)fabsf(tau1 * xt1 + tau2 * xt2));
    if (dist >= limit && (double)xt_norm < hp)
      covered = true;
    else if (dist < limit && (double)prod_scal < hp)
      covered = true;
  };
  for (int e = 0; e < cnt1 && !covered; ++e) scan(sp == SP_NODE ? L.idxC[o1 + (size_t)e * SLICE] : L.idxD[o1 + (size_t)e * SLICE]);
  for (int e = 0; e < cnt0 && !covered; ++e) {
    const int q = L.idx0[o0 + (size_t)e * SLICE] & QID_MASK;
    if (q >= P.ntotal) scan(q);
  }
  if (bc_or_not[id] != 1) bc_or_not[id] = covered ? 0 : 2;
  if (covered_out) covered_out[id] = covered ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------------
// XSPH_update's side effect on the BC flags, main:224-230: a velocity particle that is not on the free surface
// (bc_or_not /= 2) and has a velocity-particle partner that is (bc_or_not == 2) is marked 3 -- a particle with
// boundary conditions (1) thereby loses them when get_nodes_on_free_surface rewrites the flags at the end of the
// step. The marks never create or destroy a 2, so the result does not depend on the pair order. Launched only for
// XSPH together with boundary conditions, right before the per-step k_free_surface.
// ------------------------------------------------------------------------------------------------------
__global__ void k_xsph_marks(DevParams P, SlotMap M, SortArrays S, ListPtrs L, const int *__restrict__ n1,
                             int *bc_or_not) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M.nn) return;
  const int id = S.order[SP_NODE][t];
  if (S.cell[SP_NODE][t] < 0) return;  // out of the domain: no pairs
  if (bc_or_not[id] == 2) return;
  const int cnt1 = n1[t];
  const size_t o1 = (size_t)L.offC[t / SLICE] + (t & 31);
  for (int e = 0; e < cnt1; ++e) {
    const int q = L.idxC[o1 + (size_t)e * SLICE];
    if (bc_or_not[q] == 2) {
      bc_or_not[id] = 3;
      return;
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// get_nodes_on_free_surface step 4, mat:1340-1411 (ifsigman = 1 only): the normal of every marked velocity particle
// from the chords to its marked partners (pair types 1 and 3; fp32 coordinates, fp64 sums in traversal order),
// normalised and oriented against grad(f_int) (fp32 sums in traversal order; f_int = -1 on the particles that step 3
// found covered). Runs after k_free_surface of the same step: bc_or_not == 2 <=> subset == 1. A particle without a
// marked partner gets 0/0 = NaN, which apply_stress_free skips (mat:1791).
// Deviation: x**0.5 -> correctly rounded sqrt, as in k_free_surface.
// ------------------------------------------------------------------------------------------------------
__global__ void k_fs_normals(DevParams P, SlotMap M, SortArrays S, const int *__restrict__ pos_of, ListPtrs L,
                             const int *__restrict__ n0, const int *__restrict__ n1,
                             const GrowthRule *__restrict__ growth, const double *__restrict__ x,
                             const double *__restrict__ mass, const double *__restrict__ rho,
                             const int *__restrict__ bc_or_not, const int *__restrict__ covered,
                             double *__restrict__ fs_normal) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M.nn) return;
  const int id = S.order[SP_NODE][t];
  const int c = S.cell[SP_NODE][t];
  if (bc_or_not[id] != 2) {
    st2(fs_normal, id, make_double2(0.0, 0.0));
    return;
  }
  const int cnt0 = c < 0 ? 0 : n0[t], cnt1 = c < 0 ? 0 : n1[t];  // out of the domain: no pairs -> 0/0
  const int lane = t & 31, sl = t / SLICE;
  const size_t o0 = (size_t)L.off0[sl] + lane;
  const size_t o1 = (size_t)L.offC[sl] + lane;
  const GrowthRule gr = *growth;
  const okey_t kp = make_key(c < 0 ? 0 : c, SP_NODE, id);
  const double2 xp = ld2(x, id);
  const float xpf = (float)xp.x, ypf = (float)xp.y;
  const float fp = covered[id] ? -1.f : 0.f;
  auto species = [&](int q) { return q < P.nnode ? SP_NODE : (q < P.ntotal ? SP_STRESS : SP_DUMMY); };
  auto key_of = [&](int q) {
    const int sq = species(q);
    return make_key(S.cell[sq][pos_of[q]], sq, q);
  };
  auto before = [&](okey_t ka, okey_t kb) {  // traversal precedence, as in k_free_surface
    const bool oa = pair_is_old(gr, kp, ka), ob = pair_is_old(gr, kp, kb);
    if (oa != ob) return !oa;
    return oa ? ka < kb : ka > kb;
  };
  double nx = 0.0, ny = 0.0;
  float neighbour = 0.f, gf1 = 0.f, gf2 = 0.f;
  // q: partner; gx, gy: kernel gradient from this particle's perspective; first: this particle is pair_i
  auto visit = [&](int q, float gx, float gy, bool first) {
    const double2 xq = ld2(x, q);
    if (bc_or_not[q] == 2) {
      const float xqf = (float)xq.x, yqf = (float)xq.y;
      const float x_vect = first ? (xqf - xpf) : (xpf - xqf);
      const float y_vect = first ? (yqf - ypf) : (ypf - yqf);
      nx = nx - (double)y_vect;
      ny = ny + (double)x_vect;
      neighbour = neighbour + 1.f;
    }
    const double mq = mass[q], rq = rho[q];
    const float h1 = (float)(mq * (double)gx / rq), h2 = (float)(mq * (double)gy / rq);
    const float fq = covered[q] ? -1.f : 0.f;
    gf1 = gf1 + (fq - fp) * h1;
    gf2 = gf2 + (fq - fp) * h2;
  };
  int e0 = 0, e1 = 0;
  auto skip_walls = [&]() {
    while (e0 < cnt0 && (L.idx0[o0 + (size_t)e0 * SLICE] & QID_MASK) >= P.ntotal) ++e0;
  };
  skip_walls();
  while (e0 < cnt0 || e1 < cnt1) {
    bool take0;
    if (e0 >= cnt0)
      take0 = false;
    else if (e1 >= cnt1)
      take0 = true;
    else
      take0 = before(key_of(L.idx0[o0 + (size_t)e0 * SLICE] & QID_MASK), key_of(L.idxC[o1 + (size_t)e1 * SLICE]));
    if (take0) {
      const size_t a = o0 + (size_t)e0 * SLICE;
      // type 1: pair_i is the stress particle, the stored gradient is in its orientation
      visit(L.idx0[a] & QID_MASK, -L.gx0[a], -L.gy0[a], false);
      ++e0;
      skip_walls();
    } else {
      const size_t a = o1 + (size_t)e1 * SLICE;
      const int q = L.idxC[a];
      visit(q, L.gxC[a], L.gyC[a], kp < key_of(q));  // type 3: pair_i is the creator, the smaller (cell, id) key
      ++e1;
    }
  }
  nx = nx / (double)neighbour;
  ny = ny / (double)neighbour;
  const float norm_vect = (float)sqrt(nx * nx + ny * ny);
  nx = nx / (double)norm_vect;
  ny = ny / (double)norm_vect;
  const float p_scal = (float)((double)gf1 * nx + (double)gf2 * ny);
  if (p_scal < 0) {
    nx = -nx;
    ny = -ny;
  }
  st2(fs_normal, id, make_double2(nx, ny));
}

__global__ void k_pair_stats(SlotMap M, const int *__restrict__ nall, int *__restrict__ out /* max,min,zero */) {
  int mx = 0, mn = 1000, nz = 0;
  const int n = M.total();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int sp, k;
    if (!slot_decode(M, i, sp, k)) continue;  // padding slot
    const int c = nall[i];
    if (c < 0) continue;  // not owned by this rank (multi-GPU)
    mx = max(mx, c);
    mn = min(mn, c);
    nz += (c == 0);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    nz += __shfl_xor_sync(0xffffffffu, nz, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&out[0], mx);
    atomicMin(&out[1], mn);
    atomicAdd(&out[2], nz);
  }
}

}  // namespace spsph
