// Pair-sum and per-particle kernels of one time step (reference: stress_point_update main:403-482,
// get_derivatives main:487-648, RK4 main:653-802, artificial_viscosity main:826-904, plastic_terms
// mat:1884-1954, adapt_stress2 mat:2087-2161, Normal_BCs mat:1665-1770, gravity_force mat:2809-2871,
// XSPH_update main:189-239, shift_stress_points main:244-368, position update main:140-182).
//
// One thread per velocity / stress particle, walking its gather list front to back: every per-particle sum
// is accumulated in the reference's traversal order (fp32 sums bit-exact, fp64 sums too), no atomics.
//
// Data layout. The state lives in two buffer sets with *different record formats*, each written by the
// kernel that precedes its reader so that a neighbour gather is one aligned 32-byte sector:
//   format A (input of sweep A = stress_point_update):
//       NA[node] = {vx, vy, m/rho, -}            SA[sp] = {s1, s2, s3, s4, m/rho, eps_p, -, -}
//       NSa[node] = own stress (carried)         SVa[sp] = own velocity (carried)
//   format B (input of sweep B = get_derivatives; also the state between time steps):
//       NB[node] = {vx, vy, m, rho}              SB[sp] = {s1/rho^2, s2/rho^2, s3/rho^2, m}
//       NSb[node] = own stress                   SFb[sp] = own stress, SVb[sp] = own velocity
// Sweep A maps A -> B, sweep B (with the fused RK4 stage epilogue and next-stage predictor) maps B -> A; a
// kernel only gathers from the format it does not write, so there are no read/write races.
// List entries are streamed with ld.global.cs (read once per sweep), four entries in flight per thread.
#pragma once
#include "grid_kernels.cuh"

namespace spsph {

struct __align__(32) Rec4 {
  double a, b, c, d;
};
struct __align__(32) Rec8 {
  double s1, s2, s3, s4, mor, epsp, p0, p1;
};

struct StatePtrs {
  // constant per step / persistent (original particle order)
  const double *x;  // (2, ntotal2)
  const double *mass, *rho, *hsml, *mor;
  const float *wallpos, *horiz;
  const int *bc_or_not, *bc_info;
  // format A
  Rec4 *NA;     // [nnode]
  Rec8 *SA;     // [nstress]
  double *NSa;  // (4, nnode)
  double *SVa;  // (2, nstress)
  // format B (== state between steps)
  Rec4 *NB;     // [nnode]
  Rec4 *SB;     // [nstress]
  double *NSb;  // (4, nnode)
  double *SFb;  // (4, nstress)
  double *SVb;  // (2, nstress)
  // read side of a format-B -> format-B sweep (the SPH_shift interpolation): the other B buffer set
  const Rec4 *NBr;
  const double *NSbr, *SFbr, *SVbr;
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
__device__ __forceinline__ Rec4 ld256(const void *p) {
  Rec4 r;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p));
  return r;
}
__device__ __forceinline__ void st256(void *p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
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
#ifndef SPSPH_PREFETCH
#define SPSPH_PREFETCH 0  // 1: gathers of group g+1 are issued before group g is consumed (register double buffer)
#endif
#ifndef SPSPH_MINB
#define SPSPH_MINB 4      // min resident blocks per SM requested from ptxas for the sweep kernels
#endif
constexpr int ELL_GROUP = 4;  // rows per cp.async group
constexpr int ELL_NG = 4;     // groups in the ring (16 rows = 2 KB per array per warp in flight)

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// NARR arrays are streamed (array 0 holds the partner ids). `rows` is the slice width (warp-uniform), `cnt` the
// calling lane's own list length. gather(q) -> R fetches the partner record (q < 0: past the end);
// compute(q[4], a1[4], a2[4], rec[4], nvalid) consumes one group of entries in list order (a1/a2: raw 32-bit
// payloads of arrays 1 and 2; entries u >= nvalid are past the end of this lane's list and must be ignored).
template <int NARR, class R, class GatherF, class ComputeF>
__device__ __forceinline__ void ell_stream(const int *const *arr, size_t slice_off, int rows, int cnt, int *smw,
                                           GatherF gather, ComputeF compute) {
  const int lane = threadIdx.x & 31;
  const int ng = (rows + ELL_GROUP - 1) / ELL_GROUP;
  if (ng == 0) return;
  auto issue = [&](int g) {
    if (g < ng) {
#pragma unroll
      for (int a = 0; a < NARR; ++a)
        cp_async16(smw + ((g % ELL_NG) * NARR + a) * (ELL_GROUP * 32) + lane * 4,
                   arr[a] + slice_off + (size_t)g * (ELL_GROUP * 32) + lane * 4);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int g = 0; g < ELL_NG; ++g) issue(g);
  cp_async_wait<ELL_NG - 1>();
  __syncwarp();
  R cur[ELL_GROUP];
#if SPSPH_PREFETCH
#pragma unroll
  for (int u = 0; u < ELL_GROUP; ++u) {
    const int q = smw[u * 32 + lane];
    cur[u] = gather(u < cnt ? q : -1);
  }
#endif
  for (int g = 0; g < ng; ++g) {
#if SPSPH_PREFETCH
    R nxt[ELL_GROUP];
    if (g + 1 < ng) {
      cp_async_wait<ELL_NG - 2>();
      __syncwarp();
      const int *sn = smw + (((g + 1) % ELL_NG) * NARR) * (ELL_GROUP * 32);
#pragma unroll
      for (int u = 0; u < ELL_GROUP; ++u) {
        const int q = sn[u * 32 + lane];
        nxt[u] = gather(((g + 1) * ELL_GROUP + u) < cnt ? q : -1);
      }
    }
    const int *sl = smw + ((g % ELL_NG) * NARR) * (ELL_GROUP * 32);
#else
    if (g > 0) {
      cp_async_wait<ELL_NG - 1>();
      __syncwarp();
    }
    const int *sl = smw + ((g % ELL_NG) * NARR) * (ELL_GROUP * 32);
#pragma unroll
    for (int u = 0; u < ELL_GROUP; ++u) {
      const int qq = sl[u * 32 + lane];
      cur[u] = gather((g * ELL_GROUP + u) < cnt ? qq : -1);
    }
#endif
    int q[ELL_GROUP], a1[ELL_GROUP], a2[ELL_GROUP];
#pragma unroll
    for (int u = 0; u < ELL_GROUP; ++u) {
      q[u] = sl[u * 32 + lane];
      a1[u] = NARR > 1 ? sl[(ELL_GROUP * 32) + u * 32 + lane] : 0;
      a2[u] = NARR > 2 ? sl[2 * (ELL_GROUP * 32) + u * 32 + lane] : 0;
    }
    const int nvalid = cnt - g * ELL_GROUP;  // may be <= 0 or > ELL_GROUP
    compute(q, a1, a2, cur, nvalid);
    __syncwarp();
    issue(g + ELL_NG);
#if SPSPH_PREFETCH
#pragma unroll
    for (int u = 0; u < ELL_GROUP; ++u) cur[u] = nxt[u];
#endif
  }
  cp_async_wait<0>();
}
__device__ __forceinline__ int warp_max_i(int v) {
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#define ELL_SMEM(NARR) (ELL_NG * (NARR) * ELL_GROUP * 32)  // ints per warp

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
    st2(st.SVb, ks, v);
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
    st2(vel, id, ld2(st.SVb, ks));
    st4(stress, id, ld4(st.SFb, ks));
  }
}

// ------------------------------------------------------------------------------------------------------
// RK4 prologue (main:681-690 + first predictor main:700-701 with f1rk = 0 + adapt_stress2/BCs main:715-716):
// format B (state) -> format A (stage-1 input); saves vel0/stress0/vx0 and zeroes the RK accumulators.
// Stress-particle velocities and node stresses start from zero (main:690).
// ------------------------------------------------------------------------------------------------------
__global__ void k_rk_begin(DevParams P, StatePtrs st, const int *__restrict__ lflag) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= P.ntotal) return;
  if (lflag && lflag[id] == 0) return;  // remote particle (multi-GPU)
  double2 vn;
  Stress4 sn;
  if (id < P.nnode) {
    const Rec4 r = ldrec(st.NB, id);
    const double2 v = make_double2(r.a, r.b);
    st2(st.vx0, id, v);
    st2(st.vel0, id, v);
    st2(st.RKv, id, make_double2(0.0, 0.0));
    vn.x = v.x + 0. * (P.dt) * 0.0;  // vel0 + f1rk(1)*dt*RHS_2 with RHS_2 = 0
    vn.y = v.y + 0. * (P.dt) * 0.0;
    sn = Stress4{0.0, 0.0, 0.0, 0.0};
    if (P.adapt) adapt_stress(P, sn);
    apply_bcs(P, st.bc_or_not, st.bc_info, id, vn, sn);
    strec(st.NA, id, vn.x, vn.y, st.mor[id], 0.0);
    st4(st.NSa, id, sn);
  } else {
    const int ks = id - P.nnode;
    const double2 v = ld2(st.SVb, ks);
    const Stress4 s = ld4(st.SFb, ks);
    st2(st.vx0, id, v);
    st4(st.stress0, ks, s);
    st4(st.RKs, ks, Stress4{0.0, 0.0, 0.0, 0.0});
    st.RKe[ks] = 0.0;
    vn = make_double2(0.0, 0.0);
    sn.s1 = s.s1 + 0. * (P.dt) * 0.0;
    sn.s2 = s.s2 + 0. * (P.dt) * 0.0;
    sn.s3 = s.s3 + 0. * (P.dt) * 0.0;
    sn.s4 = s.s4 + 0. * (P.dt) * 0.0;
    if (P.adapt) adapt_stress(P, sn);
    apply_bcs(P, st.bc_or_not, st.bc_info, id, vn, sn);
    Rec8 *o = st.SA + ks;
    reinterpret_cast<double2 *>(o)[0] = make_double2(sn.s1, sn.s2);
    reinterpret_cast<double2 *>(o)[1] = make_double2(sn.s3, sn.s4);
    reinterpret_cast<double2 *>(o)[2] = make_double2(st.mor[id], st.epsp[id]);
    st2(st.SVa, ks, vn);
  }
}

// ------------------------------------------------------------------------------------------------------
// Sweep A (stress_point_update + the adapt_stress2 / BCs that follow it).
//   FROMB = false: input in format A (inside RK4 and the final interpolation of the step)
//   FROMB = true : input in format B (the SPH_shift interpolation at the start of a step, main:99-109)
//   FIRST: first sweep A of the step -> computes and stores cspm_norm
// ------------------------------------------------------------------------------------------------------
template <bool FIRST, bool FROMB>
__global__ void __launch_bounds__(128, SPSPH_MINB)
k_sweep_a_sp(DevParams P, SlotMap M, const int *__restrict__ order_s, ListPtrs L, const int *__restrict__ n0,
             StatePtrs st, int do_adapt, int do_bc) {
  const int k0 = blockIdx.x * blockDim.x + threadIdx.x;  // species-sorted index of the stress particle
  if ((k0 & ~31) >= M.ns) return;                         // whole warp past the end
  const bool live = k0 < M.ns;
  const int k = live ? k0 : 0;
  const int t = M.nnp + k0;
  const int id = order_s[k];
  const int ks = id - P.nnode;
  const int cnt = live ? n0[t] : 0;
  const int wrows = warp_max_i(cnt);
  double2 v;
  Stress4 s;
  if (FROMB) {
    v = ld2(st.SVbr, ks);
    s = ld4(st.SFbr, ks);
  } else {
    v = ld2(st.SVa, ks);
    const Rec8 *o = st.SA + ks;
    const double2 a = reinterpret_cast<const double2 *>(o)[0], b = reinterpret_cast<const double2 *>(o)[1];
    s = Stress4{a.x, a.y, b.x, b.y};
  }
  double vtx = 0.0, vty = 0.0, nrm = 0.0;
  {
    __shared__ __align__(16) int smem[4 * ELL_SMEM(2)];
    const int *arrs[2] = {L.idx0, reinterpret_cast<const int *>(L.w0)};
    const Rec4 *__restrict__ NR = FROMB ? st.NBr : (const Rec4 *)st.NA;
    ell_stream<2, Rec4>(
        arrs, (size_t)L.off0[t / SLICE], wrows, cnt, smem + (threadIdx.x >> 5) * ELL_SMEM(2),
        [&](int q) { return ldrec(NR, (q < 0 || q >= P.nnode) ? 0 : q); },
        [&](const int(&q)[ELL_GROUP], const int(&a1)[ELL_GROUP], const int(&)[ELL_GROUP], const Rec4(&r)[ELL_GROUP],
            int nvalid) {
#pragma unroll
          for (int u = 0; u < ELL_GROUP; ++u) {
            const bool ok = (u < nvalid) && (q[u] < P.nnode);  // dummy partners (type 9) take no part
            const double wd = (double)__int_as_float(a1[u]);
            const double mor = FROMB ? (r[u].c / r[u].d) : r[u].c;  // mass/rho
            const double h2 = mor * wd;
            const double tx = vtx + r[u].a * h2, ty = vty + r[u].b * h2;
            vtx = ok ? tx : vtx;
            vty = ok ? ty : vty;
            if (FIRST) {
              if (ok) {
                const double mq = FROMB ? r[u].c : st.mass[q[u]];
                const double rq = FROMB ? r[u].d : st.rho[q[u]];
                nrm = nrm + (wd * mq) / rq;
              }
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
    v.x = vtx / nrm;
    v.y = vty / nrm;
  }
  if (do_adapt) adapt_stress(P, s);
  if (do_bc) apply_bcs(P, st.bc_or_not, st.bc_info, id, v, s);
  st2(st.SVb, ks, v);
  st4(st.SFb, ks, s);
  const double rr = st.rho[id];
  const double r2 = rr * rr;
  strec(st.SB, ks, s.s1 / r2, s.s2 / r2, s.s3 / r2, st.mass[id]);
}

template <bool FIRST, bool FROMB, bool EPSP>
__global__ void __launch_bounds__(128, SPSPH_MINB)
k_sweep_a_node(DevParams P, SlotMap M, const int *__restrict__ order_n, ListPtrs L, const int *__restrict__ n0,
               StatePtrs st, int do_adapt, int do_bc) {
  const int k0 = blockIdx.x * blockDim.x + threadIdx.x;
  if ((k0 & ~31) >= M.nn) return;  // whole warp past the end
  const bool live = k0 < M.nn;
  const int k = live ? k0 : 0;
  const int t = k0;
  const int id = order_n[k];
  const int cnt = live ? n0[t] : 0;
  const int wrows = warp_max_i(cnt);
  double2 v;
  Stress4 s;
  if (FROMB) {
    const Rec4 r = ldrec(st.NBr, id);
    v = make_double2(r.a, r.b);
    s = ld4(st.NSbr, id);
  } else {
    const Rec4 r = ldrec(st.NA, id);
    v = make_double2(r.a, r.b);
    s = ld4(st.NSa, id);
  }
  double t1 = 0.0, t2 = 0.0, t3 = 0.0, t4 = 0.0, te = 0.0, nrm = 0.0;
  {
    struct RecS {
      Stress4 s;
      double mor, ep;
    };
    __shared__ __align__(16) int smem[4 * ELL_SMEM(2)];
    const int *arrs[2] = {L.idx0, reinterpret_cast<const int *>(L.w0)};
    ell_stream<2, RecS>(
        arrs, (size_t)L.off0[t / SLICE], wrows, cnt, smem + (threadIdx.x >> 5) * ELL_SMEM(2),
        [&](int q) {
          const int qs = (q < 0 || q >= P.ntotal) ? 0 : q - P.nnode;
          RecS r;
          if (FROMB) {
            r.s = ld4(st.SFbr, qs);
            r.mor = st.mor[qs + P.nnode];
            r.ep = EPSP ? st.epsp[qs + P.nnode] : 0.0;
          } else {
            const Rec8 *o = st.SA + qs;
            const Rec4 a = ld256(o);
            const double2 c = reinterpret_cast<const double2 *>(o)[2];
            r.s = Stress4{a.a, a.b, a.c, a.d};
            r.mor = c.x;
            r.ep = c.y;
          }
          return r;
        },
        [&](const int(&q)[ELL_GROUP], const int(&a1)[ELL_GROUP], const int(&)[ELL_GROUP], const RecS(&r)[ELL_GROUP],
            int nvalid) {
#pragma unroll
          for (int u = 0; u < ELL_GROUP; ++u) {
            const bool ok = (u < nvalid) && (q[u] < P.ntotal);  // dummy partners (type 6) take no part
            const double wd = (double)__int_as_float(a1[u]);
            const double h1 = r[u].mor * wd;
            const double n1 = t1 + r[u].s.s1 * h1, n2 = t2 + r[u].s.s2 * h1, n3 = t3 + r[u].s.s3 * h1,
                         n4 = t4 + r[u].s.s4 * h1;
            t1 = ok ? n1 : t1;
            t2 = ok ? n2 : t2;
            t3 = ok ? n3 : t3;
            t4 = ok ? n4 : t4;
            if (EPSP) {
              const double ne = te + r[u].ep * h1;
              te = ok ? ne : te;
            }
            if (FIRST) {
              if (ok) nrm = nrm + (wd * st.mass[q[u]]) / st.rho[q[u]];
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
    s.s1 = t1 / nrm;
    s.s2 = t2 / nrm;
    s.s3 = t3 / nrm;
    s.s4 = t4 / nrm;
    if (EPSP) st.epsp[id] = te / nrm;
  } else {
    v.x = 0;
    v.y = 0;
  }
  if (do_adapt) adapt_stress(P, s);
  if (do_bc) apply_bcs(P, st.bc_or_not, st.bc_info, id, v, s);
  strec(st.NB, id, v.x, v.y, st.mass[id], st.rho[id]);
  st4(st.NSb, id, s);
}

// ------------------------------------------------------------------------------------------------------
// Sweep B (get_derivatives + plastic_terms + gravity/damping + artificial viscosity + Jaumann terms + RK4
// stage accumulation + next-stage predictor, or the final RK4 update when `last`): format B -> format A.
// ------------------------------------------------------------------------------------------------------
template <bool FIRST>
__global__ void __launch_bounds__(128, SPSPH_MINB)
k_sweep_b_sp(DevParams P, SlotMap M, const int *__restrict__ order_s, ListPtrs L, const int *__restrict__ n0,
             StatePtrs st, double f1next, double f2, int last) {
  const int k0 = blockIdx.x * blockDim.x + threadIdx.x;
  if ((k0 & ~31) >= M.ns) return;  // whole warp past the end
  const bool live = k0 < M.ns;
  const int k = live ? k0 : 0;
  const int t = M.nnp + k0;
  const int id = order_s[k];
  const int ks = id - P.nnode;
  const int cnt = live ? n0[t] : 0;
  const int wrows = warp_max_i(cnt);
  const double2 vp = ld2(st.SVb, ks);
  const Stress4 sp_ = ld4(st.SFb, ks);
  double2 xp = make_double2(0.0, 0.0);
  if ((FIRST && P.cspm) || P.ndummy > 0) xp = ld2(st.x, id);
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
        double da, db;
        if (st.horiz[q] == 1.f) {
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
    __shared__ __align__(16) int smem[4 * ELL_SMEM(3)];
    const int *arrs[3] = {L.idx0, reinterpret_cast<const int *>(L.gx0), reinterpret_cast<const int *>(L.gy0)};
    ell_stream<3, Rec4>(
        arrs, (size_t)L.off0[t / SLICE], wrows, cnt, smem + (threadIdx.x >> 5) * ELL_SMEM(3),
        [&](int q) { return ldrec(st.NB, (q < 0 || q >= P.nnode) ? 0 : q); },
        [&](const int(&q)[ELL_GROUP], const int(&a1)[ELL_GROUP], const int(&a2)[ELL_GROUP], const Rec4(&r)[ELL_GROUP],
            int nvalid) {
          bool special = false;  // wall partner in this group (type 9), or the once-per-step CSPM matrix pass
#pragma unroll
          for (int u = 0; u < ELL_GROUP; ++u) special |= (u < nvalid) && (q[u] >= P.nnode);
          if ((FIRST && P.cspm) || __any_sync(0xffffffffu, special)) {
#pragma unroll
            for (int u = 0; u < ELL_GROUP; ++u)
              if (u < nvalid) entry_slow(q[u], a1[u], a2[u], r[u]);
            return;
          }
          // branch-free path: the four entries' division chains are independent and interleave
          double h1[ELL_GROUP], h2[ELL_GROUP];
#pragma unroll
          for (int u = 0; u < ELL_GROUP; ++u) {
            const double gx = (double)__int_as_float(a1[u]), gy = (double)__int_as_float(a2[u]);
            const double rr = __drcp_rn(r[u].d);
            h1[u] = div_rn(gx * r[u].c, r[u].d, rr);  // dwdx*mass(i)/rho(i), main:514
            h2[u] = div_rn(gy * r[u].c, r[u].d, rr);
          }
#pragma unroll
          for (int u = 0; u < ELL_GROUP; ++u) {
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
  if (P.ntype_eco > 1) {
    Stress4 s2 = sp_;
    if (P.ntype_solid == 1) s2.s4 = P.props[3] * (s2.s1 + s2.s2);
    double vivel[4] = {0.0, 0.0, 0.0, 0.0};
    if (P.ncrit <= 5) {
      von_mises_perzyna(P, s2, st.epsp[id], Gs, vivel);
    } else if (P.ncrit == 12) {
      double G2[4];
      double fd = st.fdp[id];
      drucker_prager(P, s2, g11, g12, g21, g22, fd, G2, vivel);
      st.fdp[id] = fd;
      Gs[0] = -G2[0];
      Gs[1] = -G2[1];
      Gs[2] = -G2[2];
      Gs[3] = -G2[3];
    }
    if (P.ntype_solid == 0)
      der1 = vivel[0];
    else
      der1 = sqrt((2.0 * (vivel[0] * vivel[0] + vivel[1] * vivel[1] + vivel[3] * vivel[3]) + vivel[2] * vivel[2]) / 3.0);
  }
  const double rke = st.RKe[ks] + der1 * f2;
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
  Stress4 rk = ld4(st.RKs, ks);
  rk.s1 = rk.s1 + f2 * r1;
  rk.s2 = rk.s2 + f2 * r2;
  rk.s3 = rk.s3 + f2 * r3;
  rk.s4 = rk.s4 + f2 * r4;
  const Stress4 s0 = ld4(st.stress0, ks);
  Stress4 sn;
  double ep = st.epsp[id];
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
    ep = ep + P.dt * (rke / 6);
    st.epsp[id] = ep;
  }
  if (P.adapt) adapt_stress(P, sn);
  double2 vn = vp;
  apply_bcs(P, st.bc_or_not, st.bc_info, id, vn, sn);
  Rec8 *o = st.SA + ks;
  reinterpret_cast<double2 *>(o)[0] = make_double2(sn.s1, sn.s2);
  reinterpret_cast<double2 *>(o)[1] = make_double2(sn.s3, sn.s4);
  reinterpret_cast<double2 *>(o)[2] = make_double2(st.mor[id], ep);
  st2(st.SVa, ks, vn);
}

template <bool FIRST>
__global__ void __launch_bounds__(128, SPSPH_MINB)
k_sweep_b_node(DevParams P, SlotMap M, const int *__restrict__ order_n, ListPtrs L, const int *__restrict__ n0,
               const int *__restrict__ n1, StatePtrs st, double f1next, double f2, int last) {
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
  const double r2p = rp * rp;
  const double so1 = sp_.s1 / r2p, so2 = sp_.s2 / r2p, so3 = sp_.s3 / r2p;  // stress(1:3,i)/rho(i)**2
  const double2 xp = ld2(st.x, id);
  double ae1 = 0.0, ae2 = 0.0, ae3 = 0.0, ae4 = 0.0, ae5 = 1.0;
  double a11 = 0.0, a12 = 0.0, a21 = 0.0, a22 = 0.0, a31 = 0.0, a32 = 0.0;  // grad2_tmp(s,k)
  __shared__ __align__(16) int smem[4 * ELL_SMEM(3)];
  int *smw = smem + (threadIdx.x >> 5) * ELL_SMEM(3);
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
    const int *arrs[3] = {L.idx0, reinterpret_cast<const int *>(L.gx0), reinterpret_cast<const int *>(L.gy0)};
    ell_stream<3, Rec4>(
        arrs, (size_t)L.off0[sl], wrows, cnt, smw,
        [&](int q) { return ldrec(st.SB, (q < 0 || q >= P.ntotal) ? 0 : q - P.nnode); },
        [&](const int(&q)[ELL_GROUP], const int(&a1)[ELL_GROUP], const int(&a2)[ELL_GROUP], const Rec4(&r)[ELL_GROUP],
            int nvalid) {
          bool special = false;  // wall partner in this group (type 6), or the once-per-step CSPM matrix pass
#pragma unroll
          for (int u = 0; u < ELL_GROUP; ++u) special |= (u < nvalid) && (q[u] >= P.ntotal);
          if ((FIRST && P.cspm) || __any_sync(0xffffffffu, special)) {
#pragma unroll
            for (int u = 0; u < ELL_GROUP; ++u)
              if (u < nvalid) entry_slow(q[u], a1[u], a2[u], r[u]);
            return;
          }
#pragma unroll
          for (int u = 0; u < ELL_GROUP; ++u) {
            const bool ok = u < nvalid;
            const double gx = (double)__int_as_float(a1[u]), gy = (double)__int_as_float(a2[u]);
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
  // artificial_viscosity, main:826-904 (fp32 locals and accumulators, list order)
  double av1 = 0.0, av2 = 0.0;
  if (P.alpha > 0 || P.beta > 0) {
    const int cntc = live ? n1[t] : 0;
    const int wrowsC = warp_max_i(cntc);
    const double hp = st.hsml[id];
    float acc1 = 0.f, acc2 = 0.f;
    struct RecV {
      Rec4 n;
      double2 x;
      double h;
    };
    const int *arrs[3] = {L.idxC, reinterpret_cast<const int *>(L.gxC), reinterpret_cast<const int *>(L.gyC)};
    __syncwarp();
    ell_stream<3, RecV>(
        arrs, (size_t)L.offC[sl], wrowsC, cntc, smw,
        [&](int q) {
          const int qq = q < 0 ? 0 : q;
          RecV r;
          r.n = ldrec(st.NB, qq);
          r.x = ld2(st.x, qq);
          r.h = st.hsml[qq];
          return r;
        },
        [&](const int(&)[ELL_GROUP], const int(&a1v)[ELL_GROUP], const int(&a2v)[ELL_GROUP], const RecV(&rv)[ELL_GROUP],
            int nvalid) {
#pragma unroll
          for (int u = 0; u < ELL_GROUP; ++u) {
            if (u >= nvalid) continue;
            const int a1 = a1v[u], a2 = a2v[u];
            const RecV &r = rv[u];
            const float gxf = __int_as_float(a1), gyf = __int_as_float(a2);
            const float xij = (float)(xp.x - r.x.x);
            const float yij = (float)(xp.y - r.x.y);
            const float h = (float)(0.5 * (hp + r.h));
            const float rho2 = (float)(0.5 * (rp + r.n.d));
            const float cs = 600.f;
            float div_u = (float)((double)xij * (vp.x - r.n.a));
            div_u = (float)((double)div_u + (double)yij * (vp.y - r.n.b));
            const float sq = sqrtf(xij * xij + yij * yij);
            const float theta = (h * div_u) / (sq * sq + 0.01f * (h * h));
            float visc = 0.f;
            if (div_u < 0)
              visc = (float)((-P.alpha * (double)cs * (double)theta + P.beta * (double)(theta * theta)) / (double)rho2);
            acc1 = (float)((double)acc1 + (double)(visc * gxf) * r.n.c);
            acc2 = (float)((double)acc2 + (double)(visc * gyf) * r.n.c);
          }
        });
    av1 = (double)(-acc1);
    av2 = (double)(-acc2);
  }
  if (!live) return;
  const double r1 = -dv1 + sg1 + av1 + 0.0 + 0.0;  // + f_bound + art_force (both zero here, main:763-764)
  const double r2 = -dv2 + sg2 + av2 + 0.0 + 0.0;
  double2 rk = ld2(st.RKv, id);
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
  apply_bcs(P, st.bc_or_not, st.bc_info, id, vn, sn);
  strec(st.NA, id, vn.x, vn.y, st.mor[id], 0.0);
  st4(st.NSa, id, sn);
}

// ------------------------------------------------------------------------------------------------------
// Position update, main:140-182: XSPH_update (main:189-239) or the fp32 mid-velocity rule; displ.
// Velocities come from format B (the state at the end of the step).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_move(DevParams P, SlotMap M, SortArrays So, ListPtrs L, const int *__restrict__ n1, StatePtrs st,
       double *__restrict__ x, const double *__restrict__ x00, double *__restrict__ displ) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M.nnp + M.nsp) return;
  const bool is_node = t < M.nnp;
  const int k = is_node ? t : t - M.nnp;
  if (k >= (is_node ? M.nn : M.ns)) return;
  const int id = is_node ? So.order[0][k] : So.order[1][k];
  double2 vp;
  if (is_node) {
    const Rec4 r = ldrec(st.NB, id);
    vp = make_double2(r.a, r.b);
  } else {
    vp = ld2(st.SVb, id - P.nnode);
  }
  const double2 xp = ld2(x, id);
  if (P.update_x) {
    double2 xn;
    if (P.xsph) {
      const int cnt = n1[t];
      const int lane = t & 31, sl = t / SLICE;
      double sx = 0.0, sy = 0.0;
      if (is_node) {
        const size_t oc = (size_t)L.offC[sl] + lane;
        for (int e0 = 0; e0 < cnt; e0 += UNR) {
          int q[UNR];
          float w[UNR];
          Rec4 r[UNR];
          double mr[UNR];
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            const size_t a = oc + (size_t)(e0 + u) * SLICE;
            q[u] = ldcs_i(L.idxC + a);
            w[u] = ldcs_f(L.wC + a);
          }
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            if (e0 + u >= cnt) q[u] = -1;
            r[u] = ldrec(st.NB, q[u] < 0 ? 0 : q[u]);
            mr[u] = st.mor[q[u] < 0 ? 0 : q[u]];
          }
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            if (q[u] < 0) continue;
            const double wd = (double)w[u];
            sx = sx + mr[u] * (r[u].a - vp.x) * wd;
            sy = sy + mr[u] * (r[u].b - vp.y) * wd;
          }
        }
      } else {
        const size_t od = (size_t)L.offD[sl] + lane;
        for (int e0 = 0; e0 < cnt; e0 += UNR) {
          int q[UNR];
          float w[UNR];
          double2 vq[UNR];
          double mr[UNR];
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            const size_t a = od + (size_t)(e0 + u) * SLICE;
            q[u] = ldcs_i(L.idxD + a);
            w[u] = ldcs_f(L.wD + a);
          }
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            if (e0 + u >= cnt) q[u] = -1;
            const int qq = q[u] < 0 ? P.nnode : q[u];
            vq[u] = ld2(st.SVb, qq - P.nnode);
            mr[u] = st.mor[qq];
          }
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            if (q[u] < 0) continue;
            const double wd = (double)w[u];
            sx = sx + mr[u] * (vq[u].x - vp.x) * wd;
            sy = sy + mr[u] * (vq[u].y - vp.y) * wd;
          }
        }
      }
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
                        const int *__restrict__ lflag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.nnode) return;
  if (lflag && lflag[i] == 0) return;  // remote particle (multi-GPU)
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

// pair statistics of grid_find_NEW, main:1405-1422
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
