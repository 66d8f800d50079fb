// Device-side common definitions for the Stress-Particle SPH step engine (sm_100a).
//
// Arithmetic contract: this translation unit is compiled with -fmad=false, so every a*b+c below is two
// IEEE roundings exactly as written (the reference is built without FMA contraction);
// fp64/fp32 division and sqrt are nvcc's IEEE-compliant defaults (-prec-div/-prec-sqrt true, -ftz false).
// Explicit __fma_rn() calls are used only inside exactly-rounded division helpers.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace spsph {

enum : int { SP_NODE = 0, SP_STRESS = 1, SP_DUMMY = 2 };

// Scalars the kernels need; passed by value as a kernel parameter (one simulation == one copy).
struct DevParams {
  int nnode, nstress, ntotal, ntotal2, ndummy, npoints;
  int skf, scale_k, cspm, update_x, xsph, ncrit, ntype_eco, ntype_solid;
  int cont_density, sle;  // continuity density on the stress particles (main:706-713); sle = 2: smoothing length follows
  int no_bcs, bc_nloop;  // bc_nloop: Normal_BCs loop bound (ntotal or nnode), mat:1683
  int ifsigman;          // 1: BCs also calls apply_stress_free (mat:1633-1637)
  int sp_sph, inside_approach, vel_vector, shift_update;
  int adapt;  // ncrit == 12
  int track_nint;  // n_int is only refreshed by artificial_viscosity / XSPH_update / isolated_nodes
  double pi, D11, D12, D22, D33, D41, D42;
  double alpha, beta, damping, dx, r_x, r_y, disp_tol, ae_thr;
  double props[20];
  double xmin_dom[2], xmax_dom[2];
  // Drucker-Prager constants of adapt_stress2 (fp64, mat:2096-2100), computed once on the host with the
  // same expressions (sqrt and division are correctly rounded on both sides)
  double dp_alpha2, dp_kc;
  // sin(props(1,9)*0.017453292) of invar09 / yieldf09 (Mohr-Coulomb, Drucker-Prager Perzyna), host libm
  double snphi;
  // per step (host-computed: time curves, exp/sin of Normal_BCs and gravity factor involve libm)
  int itimestep;
  double time_sph, dt;
  double grav[2];      // factg*ft_grav*cgrav(:), mat:2688-2711
  double bcval[16];    // bc_value per BC id at t_actual, mat:1696-1722
  int bcvar[16];       // bc_list(2, id)
};

// Cell grid of one step (grid_find_NEW Task 1, main:1245-1258); lives in device memory, written by
// k_grid_params, read by every neighbour kernel.
struct GridInfo {
  double xmin[2], xmax[2], deltx[2];
  int ndivx[2];
  int ncell;
  int overflow;  // 1 if ncell exceeds the allocated cell capacity
  int uniform_h;  // all in-domain particles share one smoothing length (enables hoisted kernel constants)
  // raw reductions
  double rxmin[2], rxmax[2], rhmax;
  int nactive[3];  // in-domain particles per species
};

struct StepStatus {  // copied to pinned host memory after the count pass
  long long n_pairs;
  long long tot0, totC, totD;  // list storage needed (entries incl. slice padding)
  int ncell, overflow, err;
  int nloc[3];  // particles per species that need processing: in the grid + out-of-domain (remote ones excluded)
  int pad[2];
};

__device__ __forceinline__ double2 ld2(const double *p, int i) { return reinterpret_cast<const double2 *>(p)[i]; }
__device__ __forceinline__ void st2(double *p, int i, double2 v) { reinterpret_cast<double2 *>(p)[i] = v; }

struct Stress4 {
  double s1, s2, s3, s4;
};
// (4, n) arrays are 32-byte aligned per particle: one 256-bit access each (LDG/STG.E.ENL2.256 on sm_100a)
#ifndef SPSPH_HOST_EMU  // device only (the host emulation of tests/native/ has its own version)
__device__ __forceinline__ Stress4 ld4(const double *p, int i) {
  Stress4 s;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(s.s1), "=d"(s.s2), "=d"(s.s3), "=d"(s.s4)
               : "l"(p + 4 * (size_t)i));
  return s;
}
__device__ __forceinline__ void st4(double *p, int i, const Stress4 &s) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p + 4 * (size_t)i), "d"(s.s1), "d"(s.s2), "d"(s.s3),
               "d"(s.s4)
               : "memory");
}
#else
inline Stress4 ld4(const double *p, int i) { return Stress4{p[4 * (size_t)i], p[4 * (size_t)i + 1], p[4 * (size_t)i + 2], p[4 * (size_t)i + 3]}; }
inline void st4(double *p, int i, const Stress4 &s) {
  p[4 * (size_t)i] = s.s1;
  p[4 * (size_t)i + 1] = s.s2;
  p[4 * (size_t)i + 2] = s.s3;
  p[4 * (size_t)i + 3] = s.s4;
}
#endif  // SPSPH_HOST_EMU

// ---- smoothing kernel, main:1440-1538 (ndimn = 2): cubic spline (skf = 1), Gauss (2), quintic (3); fp64
// evaluation, caller rounds to fp32. dx,dy = x(pair_i) - x(pair_j); r = sqrt(dx*dx + dy*dy) as computed by the
// neighbour search. Integer powers follow libgcc's __powidf2, which is what gfortran calls for `**5` / `**4`:
// x**5 = x * ((x*x)*(x*x)), x**4 = (x*x)*(x*x). The Gauss kernel's exp() is CUDA's (<= 1 ulp) where the reference
// calls glibc's: the fp32-rounded weights can differ in the last bit with probability ~1e-9 per pair.
__device__ __forceinline__ double powi5(double a) {
  const double a2 = a * a;
  return a * (a2 * a2);
}
__device__ __forceinline__ double powi4(double a) {
  const double a2 = a * a;
  return a2 * a2;
}
__device__ __forceinline__ void sph_kernel(const DevParams &P, double r, double dx, double dy, double h, double &w,
                                           double &gx, double &gy) {
  const double q = r / h;
  w = 0.;
  gx = 0.;
  gy = 0.;
  if (P.skf == 1) {
    const double factor = 15.e0 / (7.e0 * P.pi * h * h);
    if (q >= 0 && q <= 1.e0) {
      w = factor * ((double)(2.f / 3.f) - q * q + q * q * q / 2.);
      const double t = factor * (-2. + 1.5 * q) / (h * h);
      gx = t * dx;
      gy = t * dy;
    } else if (q > 1.e0 && q <= 2) {
      const double t = 2. - q;
      w = factor * 1.e0 / 6.e0 * (t * t * t);
      const double u = -factor * 1.e0 / 6.e0 * 3. * (t * t) / h;
      gx = u * (dx / r);
      gy = u * (dy / r);
    }
  } else if (P.skf == 2) {  // main:1494-1504
    const double factor = 1.e0 / ((h * h) * P.pi);
    if (q >= 0 && q <= 3) {
      w = factor * exp(-q * q);
      gx = w * (-2. * dx / h / h);
      gy = w * (-2. * dy / h / h);
    }
  } else if (P.skf == 3) {  // main:1506-1535
    const double factor = 7.e0 / (478.e0 * P.pi * h * h);
    if (q >= 0 && q <= 1) {
      w = factor * (powi5(3 - q) - 6 * powi5(2 - q) + 15 * powi5(1 - q));
      const double t = (-120 + 120 * q - 50 * (q * q)) / (h * h);
      gx = factor * (t * dx);
      gy = factor * (t * dy);
    } else if (q > 1 && q <= 2) {
      w = factor * (powi5(3 - q) - 6 * powi5(2 - q));
      const double u = factor * (-5 * powi4(3 - q) + 30 * powi4(2 - q)) / h;
      gx = u * (dx / r);
      gy = u * (dy / r);
    } else if (q > 2 && q <= 3) {
      w = factor * powi5(3 - q);
      const double u = factor * (-5 * powi4(3 - q)) / h;
      gx = u * (dx / r);
      gy = u * (dy / r);
    }
  }
}

// ---- exactly rounded division by a value whose correctly rounded reciprocal is known (Markstein): the
// result equals IEEE a/b bit for bit (checked exhaustively against the hardware divide in
// tests/test_oracle_cpu.py::test_fast_division_identity); used to hoist the per-pair divisions by h and r.
__device__ __forceinline__ double div_rn(double a, double b, double rb) {
  double q = a * rb;
  double r = __fma_rn(-b, q, a);
  q = __fma_rn(r, rb, q);
  r = __fma_rn(-b, q, a);
  return __fma_rn(r, rb, q);
}

// ---- adapt_stress2 body for one particle, mat:2087-2161 (fp64); the divisions by 3 and by sqrt(J2) go through div_rn
// (bit-identical to IEEE division, a fraction of its instructions) ----
__device__ __forceinline__ void adapt_stress(const DevParams &P, Stress4 &s) {
  const double alpha2 = P.dp_alpha2, kc = P.dp_kc;
  const double r3 = 1.0 / 3.0;  // RN(1/3), folded at compile time
  double smean = div_rn(s.s1 + s.s2 + s.s4, 3.0, r3);
  double d1 = s.s1 - smean, d2 = s.s2 - smean, d3 = s.s3, d4 = s.s4 - smean;
  double varj2 = d3 * d3 + 0.5 * (d1 * d1 + d2 * d2 + d4 * d4);
  double yield = -alpha2 * 3 * smean + kc;
  if (yield < 0) {
    const double sh = kc / (3 * alpha2);
    s.s1 = s.s1 - smean + sh;
    s.s2 = s.s2 - smean + sh;
    s.s4 = s.s4 - smean + sh;
    smean = div_rn(s.s1 + s.s2 + s.s4, 3.0, r3);
    d1 = s.s1 - smean;
    d2 = s.s2 - smean;
    d3 = s.s3;
    d4 = s.s4 - smean;
    varj2 = d3 * d3 + 0.5 * (d1 * d1 + d2 * d2 + d4 * d4);
    yield = -alpha2 * 3 * smean + kc;
  }
  const double sq = sqrt(varj2);
  if (yield < sq) {
    double rn = div_rn(-3 * alpha2 * smean + kc, sq, __drcp_rn(sq));
    if (sq <= (double)10e-06f) rn = 0;
    s.s1 = rn * d1 + smean;
    s.s2 = rn * d2 + smean;
    s.s4 = rn * d4 + smean;
    s.s3 = rn * d3;
  }
}

// ---- Normal_BCs for one particle, mat:1665-1770 ----
__device__ __forceinline__ void normal_bcs(const DevParams &P, const int *__restrict__ bc_or_not,
                                           const int *__restrict__ bc_info, int ip, double2 &v, Stress4 &s) {
  if (P.no_bcs <= 0 || ip >= P.bc_nloop) return;
  if (bc_or_not[ip] != 1) return;
  const int *bi = bc_info + 8 * (size_t)ip;
  if (bi[1] == 0) return;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const int t = bi[2 + i];
    if (t == 0) continue;
    const double val = P.bcval[t - 1];
    const int var = P.bcvar[t - 1];
    if (var == 5)
      v.x = val;
    else if (var == 6)
      v.y = val;
    else if (var == 1)
      s.s1 = val;
    else if (var == 3)
      s.s3 = val;
    else if (var == 2)
      s.s2 = val;
  }
}

// ---- apply_stress_free body for one marked velocity particle, vertical_slope copy mat:1795-1826: only the
// stress component tangential to the free surface (unit normal nx, ny) survives ----
__device__ __forceinline__ void stress_free(Stress4 &s, double nx, double ny) {
  const double costh = nx, sinth = ny;
  const double s2 = sinth * sinth, c2 = costh * costh, sc = sinth * costh;
  const double sigmatt = s2 * s.s1 - 2 * sc * s.s3 + c2 * s.s2;
  s.s1 = s2 * sigmatt;
  s.s2 = c2 * sigmatt;
  s.s3 = -sc * sigmatt;
  s.s4 = c2 * sigmatt;
}

// ---- BCs for one particle, mat:1622-1640: Normal_BCs, then (ifsigman = 1) apply_stress_free on the velocity
// particles that get_nodes_on_free_surface marked at the end of the previous step (bc_or_not = 2) and that are not
// next to a wall (bc_int /= 1); fs_normal: their step-4 normals (k_fs_normals), NaN when no marked neighbour ----
__device__ __forceinline__ void apply_bcs(const DevParams &P, const int *__restrict__ bc_or_not,
                                          const int *__restrict__ bc_info, const int *__restrict__ bc_int,
                                          const double *__restrict__ fs_normal, int ip, double2 &v, Stress4 &s) {
  if (P.no_bcs <= 0) return;
  normal_bcs(P, bc_or_not, bc_info, ip, v, s);
  if (P.ifsigman == 1 && ip < P.nnode) {
    if (bc_or_not[ip] == 2 && bc_int[ip] != 1) {
      const double2 n = ld2(fs_normal, ip);
      if (n.x == n.x && n.y == n.y) stress_free(s, n.x, n.y);  // isnan(normal) -> cycle, mat:1791
    }
  }
}

// ---- drucker_prager, mat:1958-2083: fp32 locals; returns G (= -Gs) and vivel ----
__device__ __forceinline__ void drucker_prager(const DevParams &P, const Stress4 &st, double g11, double g12, double g21,
                                               double g22, double &f_drucker, double G[4], double vivel[4]) {
  const double f0 = f_drucker;
  const float tanfi = (float)P.props[12], coh = (float)P.props[13];
  const float young = (float)P.props[2], poiss = (float)P.props[3];
  const float smean = (float)((st.s1 + st.s2 + st.s4) / 3.0);
  const float d1 = (float)(st.s1 - (double)smean);
  const float d2 = (float)(st.s2 - (double)smean);
  const float d3 = (float)st.s3;
  const float d4 = (float)(st.s4 - (double)smean);
  const float varj2 = d3 * d3 + 0.5f * (d1 * d1 + d2 * d2 + d4 * d4);
  const float vari1 = 3 * smean;
  const float eps11 = (float)g11;
  const float eps12 = (float)(0.5 * (g12 + g21));
  const float eps22 = (float)g22;
  const float emean = eps11 + eps22;
  const float alpha2 = tanfi / (sqrtf(9 + 12 * (tanfi * tanfi)));
  const float kc = (3 * coh) / (sqrtf(9 + 12 * (tanfi * tanfi)));
  const float yield = -alpha2 * vari1 + kc;
  const float sq = sqrtf(varj2);
  const float f1 = sq - yield;
  f_drucker = (double)f1;
  const double df = (double)f1 - f0;
  const float G_mod = young / (2.f * (1.f + poiss));
  const float K_mod = young / (3.f * (1.f - 2.f * poiss));
  const float s_eps = d1 * eps11 + 2 * d3 * eps12 + d2 * eps22;
  if (P.time_sph > 0 && f1 >= 0 && df >= 0 && sq >= 10e-06f) {
    const float gq = G_mod / sq;
    const float lambda_1 = 3 * alpha2 * K_mod * emean;
    const float lambda_2 = gq * s_eps;
    const float G2 = (lambda_1 + lambda_2) / G_mod;
    G[0] = (double)((gq * d1) * G2);
    G[1] = (double)((gq * d2) * G2);
    G[2] = (double)((gq * d3) * G2);
    G[3] = (double)((gq * d4) * G2);
    const float is = 1.f / sq;
    const float c6 = 1.f / 6.f;
    vivel[0] = (double)((c6 * is) * (2 * d1 - d2 - d4)) * (double)G2;
    vivel[1] = (double)((c6 * is) * (2 * d2 - d1 - d4)) * (double)G2;
    vivel[2] = (double)(is * d3) * (double)G2;
    vivel[3] = (double)((c6 * is) * (2 * d4 - d1 - d2)) * (double)G2;
  } else {
    G[0] = G[1] = G[2] = G[3] = 0.0;
    vivel[0] = vivel[1] = vivel[2] = vivel[3] = 0.0;
  }
}

// ---- Gs = -De * vivel with the plane-strain De of Get_Dmatx (strain_localisation copy :2545-2576) ----
__device__ __forceinline__ void perzyna_gs(const DevParams &P, const double vivel[4], double Gs[4]) {
  const double young = P.props[2], poiss = P.props[3];
  const double cst = young * (1.0 - poiss) / ((1.0 + poiss) * (1.0 - 2.0 * poiss));
  const double off = cst * poiss / (1.0 - poiss);
  const double d33 = (1.0 - 2.0 * poiss) * cst / (2.0 * (1.0 - poiss));
  // Gs(i) = Gs(i) - Dmatx(i,j)*vivel(j), j = 1..4 in order, zero entries included (mat:1933-1937)
  const double D[4][4] = {{cst, off, 0.0, off}, {off, cst, 0.0, off}, {0.0, 0.0, d33, 0.0}, {off, off, 0.0, cst}};
  for (int a = 0; a < 4; ++a) {
    double g = 0.0;
    for (int b = 0; b < 4; ++b) g = g - D[a][b] * vivel[b];
    Gs[a] = g;
  }
}

// ---- Get_Vivel (von Mises branch, ncrit = 2) + Get_Dmatx: strain_localisation copy :2169-2576, fp64 ----
// Returns Gs = -De * vivel and vivel. The trigonometric terms of invar09/yieldf09 only feed cons1..3 of
// the other criteria and are not evaluated.
__device__ __forceinline__ void von_mises_perzyna(const DevParams &P, const Stress4 &st, double evpstn, double Gs[4],
                                                  double vivel[4]) {
  const double root3 = (double)1.7320507764816284f;  // sqrt(3.00) in default REAL
  const double smean = (st.s1 + st.s2 + st.s4) / 3.0;
  double devia[4];
  devia[0] = st.s1 - smean;
  devia[1] = st.s2 - smean;
  devia[2] = st.s3;
  devia[3] = st.s4 - smean;
  const double varj2 = devia[2] * devia[2] + 0.5 * (devia[0] * devia[0] + devia[1] * devia[1] + devia[3] * devia[3]);
  const double steff = sqrt(varj2);
  const double yield = root3 * steff;
  const double fdatm0 = P.props[6], hards = P.props[7];
  double fdatm = fdatm0 + hards * evpstn;
  double fact = fabs(fdatm) / fabs(fdatm0);
  if (fact < (double)0.1f) fact = (double)0.1f;
  fdatm = fdatm0 * fact;
  vivel[0] = vivel[1] = vivel[2] = vivel[3] = 0.0;
  if (yield > fdatm) {
    double veca2[4] = {0, 0, 0, 0}, veca3[4];
    if (steff > 0) {
      for (int s = 0; s < 4; ++s) veca2[s] = devia[s] / (2.0 * steff);
      veca2[2] = devia[2] / steff;
    }
    veca3[0] = devia[1] * devia[3] + varj2 / 3.0;
    veca3[1] = devia[0] * devia[3] + varj2 / 3.0;
    veca3[2] = -2.0 * devia[2] * devia[3];
    veca3[3] = devia[0] * devia[1] - devia[2] * devia[2] + varj2 / 3.0;
    const double veca1[4] = {1.0, 1.0, 0.0, 1.0};
    const double cons1 = 0.0, cons2 = root3, cons3 = 0.0;
    double avect[4];
    for (int s = 0; s < 4; ++s) avect[s] = cons1 * veca1[s] + cons2 * veca2[s] + cons3 * veca3[s];
    const double allow = (double)0.01f;
    const double gamma = P.props[9], delta = P.props[10], nflow = P.props[11];
    const double fcurr = yield - fdatm;
    const double fnorm = fcurr / fdatm;
    if (fnorm >= allow) {
      double cmult;
      if (nflow != 1)
        cmult = gamma * (exp(delta * fnorm) - 1.0);
      else
        cmult = gamma * ((delta == 1.0) ? fnorm : pow(fnorm, delta));
      for (int s = 0; s < 4; ++s) vivel[s] = cmult * avect[s];
    }
  }
  perzyna_gs(P, vivel, Gs);
}

// ---- Get_Vivel for the other yield criteria of invar09 / yieldf09 (strain_localisation copy :2277-2296,
// 2423-2461): ncrit = 1 Tresca, 3 Mohr-Coulomb, 4 Drucker-Prager (Perzyna, linear hardening). Cam Clay (5) is
// refused at spsph_create. The Lode-angle terms use CUDA's asin / sin / cos / tan (<= 2 ulp) where the reference
// calls glibc's: agreement to the north star's 1e-9, not bit for bit (ncrit = 4 does not use them and is exact).
// Kept out of line and fed by value so that the sweep kernel's register allocation and its vivel[] registers are
// those of the von Mises / Drucker-Prager builds.
__device__ __noinline__ double4 perzyna_vivel_general(int ncrit, double snphi, double fdatm0, double hards, double gamma,
                                                      double delta, double nflow, double s1, double s2, double s3,
                                                      double s4, double evpstn) {
  double4 out = make_double4(0.0, 0.0, 0.0, 0.0);
  const double root3 = (double)1.7320507764816284f;  // sqrt(3.00) in default REAL
  const double smean = (s1 + s2 + s4) / 3.0;
  double devia[5];
  devia[1] = s1 - smean;
  devia[2] = s2 - smean;
  devia[3] = s3;
  devia[4] = s4 - smean;
  const double varj2 = devia[3] * devia[3] + 0.5 * (devia[1] * devia[1] + devia[2] * devia[2] + devia[4] * devia[4]);
  const double varj3 = devia[4] * (devia[4] * devia[4] - varj2);
  const double steff = sqrt(varj2);
  double sint3;
  if (steff != 0.0) {
    sint3 = -3.0 * root3 * varj3 / (2.0 * varj2 * steff);
    if (sint3 > 1.0) sint3 = 1.0;
  } else {
    sint3 = 0.0;
  }
  if (sint3 < -1.0) sint3 = -1.0;
  if (sint3 > 1.0) sint3 = 1.0;
  const double theta = asin(sint3) / 3.0;
  double yield = 0.0;
  if (ncrit == 1) {
    yield = 2.0 * cos(theta) * steff;
  } else if (ncrit == 2) {
    yield = root3 * steff;
  } else if (ncrit == 3) {
    yield = smean * snphi + steff * (cos(theta) - sin(theta) * snphi / root3);
  } else {  // ncrit == 4
    yield = 6.0 * smean * snphi / (root3 * (3.0 - snphi)) + steff;
  }
  double fdatm = fdatm0 + hards * evpstn;
  double fact = fabs(fdatm) / fabs(fdatm0);
  if (fact < (double)0.1f) fact = (double)0.1f;
  fdatm = fdatm0 * fact;
  if (!(yield > fdatm)) return out;
  // yieldf09
  double veca2[4] = {0, 0, 0, 0}, veca3[4];
  if (steff > 0) {
    for (int s = 0; s < 4; ++s) veca2[s] = devia[s + 1] / (2.0 * steff);
    veca2[2] = devia[3] / steff;
  }
  veca3[0] = devia[2] * devia[4] + varj2 / 3.0;
  veca3[1] = devia[1] * devia[4] + varj2 / 3.0;
  veca3[2] = -2.0 * devia[3] * devia[4];
  veca3[3] = devia[1] * devia[2] - devia[3] * devia[3] + varj2 / 3.0;
  const double veca1[4] = {1.0, 1.0, 0.0, 1.0};
  double cons1 = 0.0, cons2 = 0.0, cons3 = 0.0;
  if (ncrit == 1 || ncrit == 3) {
    const double tanth = tan(theta), tant3 = tan(3.0 * theta), sinth = sin(theta), costh = cos(theta),
                 cost3 = cos(3.0 * theta);
    const double abthe = fabs(theta * 57.29577951308);
    if (ncrit == 1) {
      cons1 = 0.0;
      if (abthe >= 29.0) {
        cons2 = root3;
        cons3 = 0.0;
      } else {
        cons2 = 2.0 * (costh + sinth * tant3);
        cons3 = root3 * sinth / (varj2 * cost3);
      }
    } else {
      cons1 = snphi / 3.0;
      if (abthe >= 29.0) {
        cons3 = 0.0;
        double plumi = 1.0;
        if (theta > 0.0) plumi = -1.0;
        cons2 = 0.5 * (root3 + plumi * cons1 * root3);
      } else {
        cons2 = costh * ((1.0 + tanth * tant3) + cons1 * (tant3 - tanth) * root3);
        cons3 = (root3 * sinth + 3.0 * cons1 * costh) / (2.0 * varj2 * cost3);
      }
    }
  } else if (ncrit == 2) {
    cons2 = root3;
  } else {  // ncrit == 4
    cons1 = 2.0 * snphi / (root3 * (3.0 - snphi));
    cons2 = 1.0;
  }
  double avect[4];
  for (int s = 0; s < 4; ++s) avect[s] = cons1 * veca1[s] + cons2 * veca2[s] + cons3 * veca3[s];
  // flowvp09
  const double allow = (double)0.01f;
  const double fcurr = yield - fdatm;
  const double fnorm = fcurr / fdatm;
  if (fnorm >= allow) {
    double cmult;
    if (nflow != 1)
      cmult = gamma * (exp(delta * fnorm) - 1.0);
    else
      cmult = gamma * ((delta == 1.0) ? fnorm : pow(fnorm, delta));
    out.x = cmult * avect[0];
    out.y = cmult * avect[1];
    out.z = cmult * avect[2];
    out.w = cmult * avect[3];
  }
  return out;
}

__device__ __forceinline__ void perzyna_other(const DevParams &P, const Stress4 &st, double evpstn, double Gs[4],
                                              double vivel[4]) {
  const double4 v = perzyna_vivel_general(P.ncrit, P.snphi, P.props[6], P.props[7], P.props[9], P.props[10], P.props[11],
                                          st.s1, st.s2, st.s3, st.s4, evpstn);
  vivel[0] = v.x;
  vivel[1] = v.y;
  vivel[2] = v.z;
  vivel[3] = v.w;
  perzyna_gs(P, vivel, Gs);
}

// ---- plastic_terms + Get_derivative_intvars for one stress particle, mat:1884-1954, 2697-2760 ----
// sp: its stress; g11..g22: grad_u(1,1) grad_u(1,2) grad_u(2,1) grad_u(2,2); epsp: its accumulated deviatoric
// viscoplastic strain (read for ncrit <= 5); fdp: its f_drucker (read and written for ncrit == 12).
__device__ __forceinline__ void plastic_terms(const DevParams &P, const Stress4 &sp, double g11, double g12, double g21,
                                              double g22, const double *epsp, double *fdp, double Gs[4], double &der1) {
  if (P.ntype_eco > 1) {
    Stress4 s2 = sp;
    if (P.ntype_solid == 1) s2.s4 = P.props[3] * (s2.s1 + s2.s2);
    double vivel[4] = {0.0, 0.0, 0.0, 0.0};
    if (P.ncrit == 2) {
      von_mises_perzyna(P, s2, *epsp, Gs, vivel);
    } else if (P.ncrit <= 5) {
      perzyna_other(P, s2, *epsp, Gs, vivel);
    } else if (P.ncrit == 12) {
      double G2[4];
      double fd = *fdp;
      drucker_prager(P, s2, g11, g12, g21, g22, fd, G2, vivel);
      *fdp = fd;
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
}

}  // namespace spsph
