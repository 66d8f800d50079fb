// GiD mesh / result writers of the host driver (see sph_gid.hpp). Reference: OutputMesh
// 3_SPH_material_2018.f90:2707-2744, element table :517-535, OutputRes :2930-3008.
#include "sph_gid.hpp"

#include <cmath>
#include <cstdio>
#include <stdexcept>
#include <vector>

namespace spsph {

namespace {
struct File {
  FILE *f;
  File(const std::string &path, const char *mode) : f(std::fopen(path.c_str(), mode)) {
    if (!f) throw std::runtime_error("cannot open " + path);
  }
  ~File() { std::fclose(f); }
};
}  // namespace

void gid_write_mesh(const Problem &P, const double *x, const std::string &path_prefix) {
  const spsph_params &p = P.p;
  {
    File m(path_prefix + ".post.msh", "w");
    std::fprintf(m.f, " MESH    dimension 3 ElemType Quadrilateral  Nnode 4 \n");
    std::fprintf(m.f, "  coordinates\n");
    for (int i = 0; i < p.nnode; ++i)  // velocity particles only (mat:2720-2722)
      std::fprintf(m.f, " %d %.17g %.17g 0\n", i + 1, x[2 * (size_t)i], x[2 * (size_t)i + 1]);
    std::fprintf(m.f, "  End coordinates\n");
    std::fprintf(m.f, "  Elements\n");
    // quadrilaterals of the velocity-particle lattice, numbered column by column (mat:517-535): node k of column i
    // with its left neighbour k - ndivy, that one's lower neighbour and its own lower neighbour
    const int ndivx = P.ndivx, ndivy = P.ndivy;
    int i = 2, k = ndivy + 2, j = 1;
    while (i <= ndivx) {
      while (k <= i * ndivy) {
        std::fprintf(m.f, " %d %d %d %d %d  1\n", j, k, k - ndivy, k - ndivy - 1, k - 1);
        ++k;
        ++j;
      }
      ++i;
      ++k;
    }
    if (j - 1 != P.nelem) throw std::runtime_error("gid_write_mesh: element count differs from nelem");
    std::fprintf(m.f, "  End elements\n");
  }
  File r(path_prefix + ".post.res", "w");
  std::fprintf(r.f, " GiD Post Results File 1.0\n");
  std::fprintf(r.f, " GaussPoints \"Group1\" ElemType  Quadrilateral\n");
  std::fprintf(r.f, " Number Of Gauss Points: 1\n");
  std::fprintf(r.f, " Natural Coordinates: Internal\n");
  std::fprintf(r.f, " End GaussPoints\n");
}

void gid_append_results(const Problem &P, const spsph_state &s, double time_sph, const std::string &path_prefix) {
  const spsph_params &p = P.p;
  File r(path_prefix + ".post.res", "a");
  FILE *f = r.f;
  const int nn = p.nnode;
  // displacement (always written: results deformation), mat:2937-2954
  std::fprintf(f, " Result \"disp\" \"disp\"  %.17g  Vector OnNodes \"where\" \n  Values \n", time_sph);
  for (int i = 0; i < nn; ++i) {
    double dx, dy;
    if (p.update_x) {
      dx = s.x[2 * (size_t)i] - s.x00[2 * (size_t)i];
      dy = s.x[2 * (size_t)i + 1] - s.x00[2 * (size_t)i + 1];
    } else {
      dx = s.displ[2 * (size_t)i];
      dy = s.displ[2 * (size_t)i + 1];
    }
    std::fprintf(f, " %d %.17g %.17g 0.\n", i + 1, dx, dy);
  }
  std::fprintf(f, "  End Values \n");
  // velocity vector and its modulus, mat:2959-2967
  std::fprintf(f, " Result \"vel\" \"veloc\"  %.17g  Vector OnNodes \"where\" \n  Values \n", time_sph);
  for (int i = 0; i < nn; ++i) {
    const double ux = s.vel[2 * (size_t)i], uy = s.vel[2 * (size_t)i + 1];
    std::fprintf(f, " %d %.17g %.17g %.17g\n", i + 1, ux, uy, std::sqrt(ux * ux + uy * uy));
  }
  std::fprintf(f, "  End Values \n");
  const char *names[3][2] = {{"sigmaxx", "sxx"}, {"sigmayy", "syy"}, {"sigmaxy", "sxy"}};
  for (int c = 0; c < 3; ++c) {  // mat:2972-3001
    if (P.stress_out[c] != 1) continue;
    std::fprintf(f, " Result \"%s\" \"%s\"  %.17g  Vector OnNodes \"where\" \n  Values \n", names[c][0], names[c][1], time_sph);
    for (int i = 0; i < nn; ++i) std::fprintf(f, " %d %.17g  0.   0.   \n", i + 1, s.stress[4 * (size_t)i + c]);
    std::fprintf(f, "  End Values \n");
  }
  if (P.strain_out == 1) {  // mat:3003-3010
    std::fprintf(f, " Result \"plastic strain\" \"esp\"  %.17g  Vector OnNodes \"where\" \n  Values  \n", time_sph);
    for (int i = 0; i < nn; ++i)
      std::fprintf(f, " %d %.17g  0.  0.   \n", i + 1, s.internal_vars[(size_t)SPSPH_NINT_VARS * i]);
    std::fprintf(f, "  End Values \n");
  }
}

}  // namespace spsph
