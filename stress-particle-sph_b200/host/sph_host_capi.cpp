// C entry points of the host-side reader (libspsph_host.so), used by the Python binding and the tests.
#include <cstring>
#include <string>

#include "sph_gid.hpp"
#include "sph_problem.hpp"

extern "C" {

void *spsph_problem_load(const char *dir, int variant, char *err, int errlen) {
  try {
    return new spsph::Problem(spsph::load_problem(dir, variant));
  } catch (const std::exception &e) {
    if (err && errlen > 0) {
      std::strncpy(err, e.what(), (size_t)errlen - 1);
      err[errlen - 1] = 0;
    }
    return nullptr;
  }
}

const spsph_params *spsph_problem_params(void *h) { return &((spsph::Problem *)h)->p; }

int spsph_problem_state(void *h, spsph_state *out) {
  *out = ((spsph::Problem *)h)->view();
  return 0;
}

int spsph_problem_nblocks(void *h) { return (int)((spsph::Problem *)h)->blocks.size(); }

int spsph_problem_block(void *h, int k, double *dt, double *time_end, int *maxtimestep, int *print_step, int *save_step,
                        int *plot_step) {
  auto *P = (spsph::Problem *)h;
  if (k < 0 || k >= (int)P->blocks.size()) return 1;
  const auto &b = P->blocks[k];
  *dt = b.dt;
  *time_end = b.time_end;
  *maxtimestep = b.maxtimestep;
  *print_step = b.print_step;
  *save_step = b.save_step;
  *plot_step = b.plot_step;
  return 0;
}

const char *spsph_problem_name(void *h) { return ((spsph::Problem *)h)->name.c_str(); }

// GiD writers of the host driver (sph_gid.hpp): mesh + result header from the positions x, one result frame from a state
int spsph_problem_gid_mesh(void *h, const double *x, const char *path_prefix) {
  try {
    spsph::gid_write_mesh(*(spsph::Problem *)h, x, path_prefix);
    return 0;
  } catch (const std::exception &) {
    return 1;
  }
}
int spsph_problem_gid_results(void *h, const spsph_state *s, double time_sph, const char *path_prefix) {
  try {
    spsph::gid_append_results(*(spsph::Problem *)h, *s, time_sph, path_prefix);
    return 0;
  } catch (const std::exception &) {
    return 1;
  }
}

void spsph_problem_free(void *h) { delete (spsph::Problem *)h; }

}  // extern "C"
