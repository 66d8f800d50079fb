// sph_driver -- C++ stand-in for PROGRAM SPH_2018 (code/1_SPH_2018.f90:95-229): reads the three input files,
// uploads the particles, runs the time blocks through the C-ABI (spsph_step == time_integration) and writes the
// ParaView CSV frames of OutputRes (3_SPH_material_2018.f90:2919-3060, Bui copy) at the reference's cadence.
// It exists because no Fortran compiler is available here; with gfortran the unmodified Fortran driver calls the
// same C-ABI through fortran/spsph_shim.f90.
//
// usage: sph_driver <deck directory> <variant: code|bui|vs|sl> [--max-steps N] [--out DIR] [--device D]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "sph_gid.hpp"
#include "sph_problem.hpp"

static int variant_of(const char *s) {
  if (!std::strcmp(s, "code")) return SPSPH_VARIANT_CODE;
  if (!std::strcmp(s, "bui")) return SPSPH_VARIANT_BUI;
  if (!std::strcmp(s, "vs")) return SPSPH_VARIANT_VS;
  if (!std::strcmp(s, "sl")) return SPSPH_VARIANT_SL;
  return -1;
}

// OutputRes: the GiD frame (sph_gid.cpp), then the ParaView part: nodes.csv.NNNNNN, stress_points.csv.NNNNNN (column
// selection by the *_out flags)
static void output_res(spsph::Problem &P, int itimestep_sph, double time_sph, const std::string &out) {
  const spsph_params &p = P.p;
  spsph::gid_append_results(P, P.view(), time_sph, out + "/" + P.name);
  char tag[16];
  std::snprintf(tag, sizeof tag, "%06d", itimestep_sph);
  auto row = [&](FILE *f, int i, bool node) {
    std::fprintf(f, "%16.8f,%16.8f,", P.x[2 * (size_t)i], P.x[2 * (size_t)i + 1]);
    if (P.vel_out[0] == 1) std::fprintf(f, "%16.8f,", P.vel[2 * (size_t)i]);
    if (P.vel_out[1] == 1) std::fprintf(f, "%16.8f,", P.vel[2 * (size_t)i + 1]);
    for (int s = 0; s < 4; ++s)
      if (P.stress_out[s] == 1) std::fprintf(f, "%16.8f,", P.stress[4 * (size_t)i + s]);
    if (P.strain_out == 1) std::fprintf(f, "%16.8f,", P.internal_vars[(size_t)SPSPH_NINT_VARS * i]);
    if (node && P.disp_out == 1) std::fprintf(f, "%16.8f,", P.disp_10[i]);
    if (P.rho_out == 1) std::fprintf(f, "%16.8f,", P.rho[i]);
    if (P.sml_out == 1) std::fprintf(f, "%16.8f,", P.hsml[i]);
    std::fprintf(f, "  \n");
  };
  FILE *f = std::fopen((out + "/nodes.csv." + tag).c_str(), "w");
  if (!f) return;
  std::fprintf(f, "x-coord,y-coord,");
  if (P.vel_out[0] == 1) std::fprintf(f, "x-vel,");
  if (P.vel_out[1] == 1) std::fprintf(f, "y-vel,");
  const char *sn[4] = {"sxx,", "syy,", "sxy,", "szz,"};
  for (int s = 0; s < 4; ++s)
    if (P.stress_out[s] == 1) std::fprintf(f, "%s", sn[s]);
  if (P.strain_out == 1) std::fprintf(f, "strain,");
  if (P.disp_out == 1) std::fprintf(f, "disp_10,");
  if (P.rho_out == 1) std::fprintf(f, "density,");
  if (P.sml_out == 1) std::fprintf(f, "sml,");
  std::fprintf(f, "  \n");
  for (int i = 0; i < p.nnode; ++i) row(f, i, true);
  std::fclose(f);
  f = std::fopen((out + "/stress_points.csv." + tag).c_str(), "w");
  if (!f) return;
  for (int i = p.nnode; i < p.ntotal; ++i) row(f, i, false);
  std::fclose(f);
  // nodes on the free surface (get_nodes_on_free_surface marks them bc_or_not = 2), mat:3061-3078
  f = std::fopen((out + "/surface_points.csv." + tag).c_str(), "w");
  if (!f) return;
  std::fprintf(f, " x-coord,y-coord,\n");
  for (int i = 0; i < p.nnode; ++i)
    if (P.bc_or_not[i] == 2) std::fprintf(f, " %.17g , %.17g\n", P.x[2 * (size_t)i], P.x[2 * (size_t)i + 1]);
  std::fclose(f);
}

// What the writers above print, fetched as ONE table packed on the device (spsph_download_frame) instead of a
// spsph_download of every array: 15 columns per velocity particle, 11 per stress particle, against the 40 doubles per
// particle of the full state. The rows are spread over the Problem's arrays, which the writers read.
static int download_frame(spsph_handle *h, spsph::Problem &P) {
  const spsph_params &p = P.p;
  static const int32_t node_cols[] = {SPSPH_COL_X,    SPSPH_COL_Y,      SPSPH_COL_VX,     SPSPH_COL_VY,       SPSPH_COL_SXX,
                                      SPSPH_COL_SYY,  SPSPH_COL_SXY,    SPSPH_COL_SZZ,    SPSPH_COL_EPSP,     SPSPH_COL_DISP10,
                                      SPSPH_COL_RHO,  SPSPH_COL_HSML,   SPSPH_COL_DISPLX, SPSPH_COL_DISPLY,   SPSPH_COL_BC_OR_NOT};
  static const int32_t sp_cols[] = {SPSPH_COL_X,   SPSPH_COL_Y,   SPSPH_COL_VX,   SPSPH_COL_VY,  SPSPH_COL_SXX, SPSPH_COL_SYY,
                                    SPSPH_COL_SXY, SPSPH_COL_SZZ, SPSPH_COL_EPSP, SPSPH_COL_RHO, SPSPH_COL_HSML};
  const int wn = sizeof node_cols / sizeof *node_cols, ws = sizeof sp_cols / sizeof *sp_cols;
  const int nn = p.nnode, ns = p.ntotal - p.nnode;
  static std::vector<double> tab;
  tab.resize(std::max((size_t)nn * wn, (size_t)ns * ws) + 1);
  auto common = [&](const double *r, size_t i) {
    P.x[2 * i] = r[0];
    P.x[2 * i + 1] = r[1];
    P.vel[2 * i] = r[2];
    P.vel[2 * i + 1] = r[3];
    for (int c = 0; c < 4; ++c) P.stress[4 * i + c] = r[4 + c];
    P.internal_vars[(size_t)SPSPH_NINT_VARS * i] = r[8];
  };
  if (spsph_download_frame(h, node_cols, wn, 0, nn, tab.data())) return 1;
  for (size_t i = 0; i < (size_t)nn; ++i) {
    const double *r = &tab[i * wn];
    common(r, i);
    P.disp_10[i] = r[9];
    P.rho[i] = r[10];
    P.hsml[i] = r[11];
    P.displ[2 * i] = r[12];
    P.displ[2 * i + 1] = r[13];
    P.bc_or_not[i] = (int)r[14];
  }
  if (spsph_download_frame(h, sp_cols, ws, nn, ns, tab.data())) return 1;
  for (size_t k = 0; k < (size_t)ns; ++k) {
    const double *r = &tab[k * ws];
    common(r, nn + k);
    P.rho[nn + k] = r[9];
    P.hsml[nn + k] = r[10];
  }
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s <deck dir> <code|bui|vs|sl> [--max-steps N] [--out DIR] [--device D]\n", argv[0]);
    return 2;
  }
  const int variant = variant_of(argv[2]);
  long max_steps = -1;
  int device = 0;
  std::string out = argv[1];
  for (int a = 3; a + 1 < argc; a += 2) {
    if (!std::strcmp(argv[a], "--max-steps")) max_steps = std::atol(argv[a + 1]);
    if (!std::strcmp(argv[a], "--out")) out = argv[a + 1];
    if (!std::strcmp(argv[a], "--device")) device = std::atoi(argv[a + 1]);
  }
  if (variant < 0) {
    std::fprintf(stderr, "unknown variant %s\n", argv[2]);
    return 2;
  }
  spsph::Problem P;
  try {
    P = spsph::load_problem(argv[1], variant);  // Init_sph -> problem_input_data
  } catch (const std::exception &e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  spsph_handle *h = nullptr;
  auto die = [&](const char *what) {
    std::fprintf(stderr, "%s: %s\n", what, spsph_last_error(h));  // the reference prints and STOPs
    spsph_destroy(h);
    return 1;
  };
  if (spsph_create(&h, &P.p, device)) return die("spsph_create");
  spsph_state st = P.view();
  if (spsph_upload(h, &st)) return die("spsph_upload");

  spsph::gid_write_mesh(P, P.x.data(), out + "/" + P.name);  // Init_sph -> OutputMesh, main:42-44
  // main loop, 1_SPH_2018.f90:134-199 (fp32 output-cadence counters as in 6_SPH_time_vars_2018.f90:40-45)
  double time_sph = 0.0;
  int itimestep_sph = 0;
  long total = 0;
  for (const spsph::TimeBlock &b : P.blocks) {
    const double dt = b.dt;
    output_res(P, itimestep_sph, time_sph, out);  // initial frame, 1_SPH_2018.f90:156
    const float time_print = (float)(b.print_step * dt), time_plot = (float)(b.plot_step * dt);
    float t_print_reset = 0.f, t_plot_reset = 0.f;
    double time = time_sph;
    for (int it = 1; it <= b.maxtimestep; ++it) {
      t_print_reset = (float)((double)t_print_reset + dt);
      t_plot_reset = (float)((double)t_plot_reset + dt);
      itimestep_sph += 1;
      if (spsph_step(h, itimestep_sph, time_sph, dt)) return die("spsph_step");
      time_sph = time_sph + dt;
      time = time + dt;
      ++total;
      if (time > b.time_end) break;  // 1_SPH_2018.f90:82: leaves the block before any output
      const bool stop = (max_steps >= 0 && total >= max_steps);  // --max-steps (not in the reference) ends with a frame
      if (t_plot_reset >= time_plot || stop) {
        if (download_frame(h, P)) return die("spsph_download_frame");
        output_res(P, itimestep_sph, time_sph, out);
        t_plot_reset = 0.f;
      }
      if (t_print_reset >= time_print) {  // Out_print_sph, main:51-73
        long long npairs = 0;
        int mx = 0, mn = 0, no = 0;
        spsph_pair_stats(h, (int64_t *)&npairs, &mx, &mn, &no);
        std::printf(" time step_sph is %d time_sph = %.9g dt_sph = %g pairs = %lld max/min interactions %d/%d\n",
                    itimestep_sph, time_sph, dt, npairs, mx, mn);
        t_print_reset = 0.f;
      }
      if (stop) break;
    }
    if (max_steps >= 0 && total >= max_steps) break;
  }
  spsph_destroy(h);
  return 0;
}
