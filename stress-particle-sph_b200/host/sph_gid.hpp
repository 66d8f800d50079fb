// GiD post-processing files of the reference driver: <name>.post.msh written once by OutputMesh
// (3_SPH_material_2018.f90:2707-2744: node coordinates, the quadrilateral elements of the velocity-particle lattice,
// and the header of the result file) and one result block per plotted frame appended to <name>.post.res by OutputRes
// (:2930-3008: displacement and velocity vectors, the stress components and the plastic strain selected by
// stress_out / strain_out, all on the velocity particles). Host-only: the writers read the arrays a spsph_download
// filled. The reference writes list-directed (free-format) records; tokens are separated by blanks here and every
// REAL carries 17 significant digits, so that a frame read back equals the downloaded values bit for bit.
#pragma once
#include <string>

#include "sph_problem.hpp"

namespace spsph {

// OutputMesh: writes <path_prefix>.post.msh and starts <path_prefix>.post.res. x: (2, ntotal2) positions to list.
void gid_write_mesh(const Problem &P, const double *x, const std::string &path_prefix);
// OutputRes, GiD part: appends one frame to <path_prefix>.post.res
void gid_append_results(const Problem &P, const spsph_state &s, double time_sph, const std::string &path_prefix);

}  // namespace spsph
