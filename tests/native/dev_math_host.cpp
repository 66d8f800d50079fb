// TEST INFRASTRUCTURE. The per-particle device functions of stress-particle-sph_b200/csrc/dev_common.cuh are plain
// C++ arithmetic (no intrinsics); this file compiles them for the HOST (g++ -ffp-contract=off, the CUDA build uses
// -fmad=false) so that tests/test_oracle_cpu.py can check their transcription against the oracle's restatement of
// the same reference routines on random inputs, bit for bit, without a GPU. What it cannot see is the difference
// between CUDA's and glibc's asin/sin/cos/tan/exp/pow (the GPU tests carry a 1e-9 tolerance for those paths).
//   g++ -O2 -ffp-contract=off -std=c++17 -fPIC -shared -D__noinline__= -I/usr/local/cuda/include -I<csrc> ...
#include <cmath>
#include <cstring>

// the two intrinsics of div_rn (exactly rounded division by a value with a known reciprocal) on the host
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double __drcp_rn(double x) { return 1.0 / x; }

#include "dev_common.cuh"

using namespace spsph;

static DevParams make_params(int ncrit, int ntype_eco, int ntype_solid, double time_sph, const double *props20) {
  DevParams P;
  std::memset(&P, 0, sizeof(P));
  P.ncrit = ncrit;
  P.ntype_eco = ntype_eco;
  P.ntype_solid = ntype_solid;
  P.time_sph = time_sph;
  for (int k = 0; k < 20; ++k) P.props[k] = props20[k];
  P.snphi = std::sin(props20[8] * (double)0.017453292f);  // as spsph_create does
  const double tanfi = props20[12], coh = props20[13];     //
  P.dp_alpha2 = tanfi / (std::sqrt(9 + 12 * (tanfi * tanfi)));
  P.dp_kc = (3 * coh) / (std::sqrt(9 + 12 * (tanfi * tanfi)));
  return P;
}

extern "C" {

// plastic_terms + Get_derivative_intvars of n stress particles: stress (4,n), grad_u (4,n) as g11 g12 g21 g22,
// epsp (n), f_drucker (n, in/out) -> Gs (4,n), der1 (n)
void devmath_plastic_terms(int ncrit, int ntype_eco, int ntype_solid, double time_sph, const double *props20, int n,
                           const double *stress, const double *grad, const double *epsp, double *f_drucker, double *Gs,
                           double *der1) {
  const DevParams P = make_params(ncrit, ntype_eco, ntype_solid, time_sph, props20);
  for (int i = 0; i < n; ++i) {
    Stress4 s = {stress[4 * i], stress[4 * i + 1], stress[4 * i + 2], stress[4 * i + 3]};
    double g[4] = {0, 0, 0, 0}, d = 0.0;
    plastic_terms(P, s, grad[4 * i], grad[4 * i + 1], grad[4 * i + 2], grad[4 * i + 3], epsp + i, f_drucker + i, g, d);
    for (int k = 0; k < 4; ++k) Gs[4 * i + k] = g[k];
    der1[i] = d;
  }
}

// adapt_stress2 of n particles, stress (4,n) in place
void devmath_adapt_stress(const double *props20, int n, double *stress) {
  const DevParams P = make_params(12, 2, 2, 1.0, props20);
  for (int i = 0; i < n; ++i) {
    Stress4 s = {stress[4 * i], stress[4 * i + 1], stress[4 * i + 2], stress[4 * i + 3]};
    adapt_stress(P, s);
    stress[4 * i] = s.s1;
    stress[4 * i + 1] = s.s2;
    stress[4 * i + 2] = s.s3;
    stress[4 * i + 3] = s.s4;
  }
}

// apply_stress_free of n velocity particles (all marked, none next to a wall): stress (4,n) in place, normal (2,n)
void devmath_stress_free(int n, double *stress, const double *normal) {
  for (int i = 0; i < n; ++i) {
    Stress4 s = {stress[4 * i], stress[4 * i + 1], stress[4 * i + 2], stress[4 * i + 3]};
    stress_free(s, normal[2 * i], normal[2 * i + 1]);
    stress[4 * i] = s.s1;
    stress[4 * i + 1] = s.s2;
    stress[4 * i + 2] = s.s3;
    stress[4 * i + 3] = s.s4;
  }
}

// smoothing kernel of n pairs: r, dx, dy, h -> w, gx, gy (fp64, before the fp32 rounding of the pair record)
void devmath_kernel(int skf, double pi, int n, const double *r, const double *dx, const double *dy, const double *h,
                    double *w, double *gx, double *gy) {
  DevParams P;
  std::memset(&P, 0, sizeof(P));
  P.skf = skf;
  P.pi = pi;
  for (int i = 0; i < n; ++i) sph_kernel(P, r[i], dx[i], dy[i], h[i], w[i], gx[i], gy[i]);
}
}
