// Brute-force check of the identity behind div_rn() in grid_kernels.cuh: with y = RN(1/b), two Markstein
// corrections of q = a*y give exactly RN(a/b) (also after one correction). usage: div_identity [ncases]
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <random>
#include <cstring>
#include <cstdlib>
static inline double div_fast(double a, double b, double y) {
  double q = a * y;
  double r = std::fma(-b, q, a);
  q = std::fma(r, y, q);
  r = std::fma(-b, q, a);
  return std::fma(r, y, q);
}
static inline double div_fast1(double a, double b, double y) {
  double q = a * y;
  double r = std::fma(-b, q, a);
  return std::fma(r, y, q);
}
int main(int argc, char **argv) {
  const uint64_t N = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 20000000ull;
  std::mt19937_64 g(12345);
  uint64_t bad2 = 0, bad1 = 0, n = 0;
  auto rnd = [&](int emin, int emax) {
    uint64_t m = g() & ((1ull << 52) - 1);
    int e = emin + (int)(g() % (uint64_t)(emax - emin + 1));
    uint64_t bits = ((uint64_t)(e + 1023) << 52) | m;
    double d; std::memcpy(&d, &bits, 8);
    if (g() & 1) d = -d;
    return d;
  };
  for (uint64_t it = 0; it < N; ++it) {
    double b = std::fabs(rnd(-40, 40));
    double a = rnd(-60, 60);
    // adversarial: make a/b close to a representable value or midpoint sometimes
    if ((it & 7) == 0) { double k = rnd(-5, 5); a = k * b; a = std::nextafter(a, (it & 8) ? 1e300 : -1e300); }
    if ((it & 7) == 1) { uint64_t bits; std::memcpy(&bits, &b, 8); bits |= ((1ull<<52)-1) & ~(g() & 0xff); std::memcpy(&b, &bits, 8); }
    double y = 1.0 / b;
    double ref = a / b;
    if (div_fast(a, b, y) != ref) ++bad2;
    if (div_fast1(a, b, y) != ref) ++bad1;
    ++n;
  }
  printf("n=%llu mismatches: two-step=%llu one-step=%llu\n", (unsigned long long)n, (unsigned long long)bad2, (unsigned long long)bad1);
  return bad2 ? 1 : 0;
}
