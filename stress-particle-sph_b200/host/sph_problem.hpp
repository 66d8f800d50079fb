// Host-side problem set-up: reader for input.txt / <name>.dat / <name>.pts and particle generation.
//
// This is the C++ stand-in for the part of the reference's Fortran driver that stays on the host
// (problem_input_data, Read_2DMesh, set_up_dummy_nodes, Setup_Global_Arrays, Get_BCs_on_node,
// Initial_conditions: example_problems/soil_failure_bui_et_al_2008/3_SPH_material_2018.f90:37-1131,
// 1433-1569). It exists in C++ only because no Fortran compiler is available in this image; the
// Fortran driver can keep doing this work itself and hand the same arrays to spsph_upload().
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "spsph.h"

namespace spsph {

struct TimeBlock {  // one "dt, time_end, maxtimestep / print, save, plot" block of input.txt (1_SPH_2018.f90:141-152)
  double dt = 0, time_end = 0;
  int maxtimestep = 0, print_step = 0, save_step = 0, plot_step = 0;
};

struct Problem {
  spsph_params p{};
  std::string name, dir;
  // particle arrays, reference layout (column-major, 0-based storage of the 1-based Fortran arrays)
  std::vector<double> x, vel, stress, rho, mass, hsml, internal_vars, f_drucker, x00, displ, x_10, disp_10;
  std::vector<int32_t> itype, if_out_domain, bc_or_not, bc_info, bc_int;
  std::vector<float> wall_position, horizontal_or_not, n_int;
  std::vector<TimeBlock> blocks;
  int stress_out[4] = {0, 0, 0, 0}, vel_out[2] = {0, 0}, strain_out = 0, rho_out = 0, sml_out = 0, disp_out = 0;
  int ndivx = 0, ndivy = 0, nelem = 0;
  std::vector<std::string> chk;  // echo of what was parsed (the reference's .chk file, cosmetic spacing aside)

  spsph_state view();  // pointers into the vectors above
};

// Throws std::runtime_error on malformed input.
Problem load_problem(const std::string &dir, int variant);

}  // namespace spsph
