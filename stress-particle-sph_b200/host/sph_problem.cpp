// Host-side reader + particle set-up. See sph_problem.hpp for scope.
// Line references "mat:" are to example_problems/soil_failure_bui_et_al_2008/3_SPH_material_2018.f90
// (the Bui copy; the other copies differ only where noted).
//
// Arithmetic notes: the reference mixes default REAL (fp32) and REAL(DP); every
// fp32 rounding that feeds particle data is reproduced here (wall geometry, pi, 4./3. literals).
// Build with -ffp-contract=off so no FMA contraction changes a rounding.
#include "sph_problem.hpp"

#include <cmath>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace spsph {

namespace {

// Fortran list-directed (free-format) record reader: values separated by blanks or commas, a READ
// always starts on a fresh record, continues over records until its list is satisfied and then
// discards the rest of the record.
class ListReader {
 public:
  explicit ListReader(const std::string &path) : path_(path), in_(path) {
    if (!in_) throw std::runtime_error("cannot open " + path);
  }
  // read(unit,*) text : first token of the next non-blank record
  std::string text() {
    std::vector<std::string> t = tokens(1);
    return t[0];
  }
  std::vector<std::string> tokens(size_t n) {
    std::vector<std::string> out;
    while (out.size() < n) {
      std::string line;
      if (!std::getline(in_, line)) throw std::runtime_error("unexpected end of file in " + path_);
      for (char &c : line)
        if (c == ',' || c == '\t' || c == '\r') c = ' ';
      std::istringstream ss(line);
      std::string tok;
      while (out.size() < n && (ss >> tok)) out.push_back(tok);
    }
    return out;
  }
  std::vector<double> reals(size_t n) {
    std::vector<double> v;
    for (auto &t : tokens(n)) v.push_back(to_double(t));
    return v;
  }
  std::vector<float> reals32(size_t n) {  // items read into default REAL variables
    std::vector<float> v;
    for (auto &t : tokens(n)) v.push_back(to_float(t));
    return v;
  }
  std::vector<int> ints(size_t n) {
    std::vector<int> v;
    for (auto &t : tokens(n)) v.push_back((int)to_double(t));
    return v;
  }
  std::vector<int> logicals(size_t n) {
    std::vector<int> v;
    for (auto &t : tokens(n)) {
      size_t k = (t[0] == '.') ? 1 : 0;
      char c = k < t.size() ? t[k] : '?';
      if (c == 'T' || c == 't')
        v.push_back(1);
      else if (c == 'F' || c == 'f')
        v.push_back(0);
      else
        throw std::runtime_error("bad logical '" + t + "' in " + path_);
    }
    return v;
  }

 private:
  static std::string fix(std::string t) {
    for (char &c : t)
      if (c == 'd' || c == 'D') c = 'e';
    return t;
  }
  double to_double(const std::string &t) {
    std::string f = fix(t);
    char *end = nullptr;
    double v = std::strtod(f.c_str(), &end);
    if (end == f.c_str()) throw std::runtime_error("bad number '" + t + "' in " + path_);
    return v;
  }
  float to_float(const std::string &t) {
    std::string f = fix(t);
    char *end = nullptr;
    float v = std::strtof(f.c_str(), &end);
    if (end == f.c_str()) throw std::runtime_error("bad number '" + t + "' in " + path_);
    return v;
  }
  std::string path_;
  std::ifstream in_;
};

std::string join(const std::vector<double> &v) {
  std::ostringstream ss;
  ss.precision(17);
  for (double d : v) ss << ' ' << d;
  return ss.str();
}

}  // namespace

spsph_state Problem::view() {
  spsph_state s{};
  s.x = x.data();
  s.vel = vel.data();
  s.stress = stress.data();
  s.rho = rho.data();
  s.mass = mass.data();
  s.hsml = hsml.data();
  s.itype = itype.data();
  s.internal_vars = internal_vars.data();
  s.f_drucker = f_drucker.data();
  s.x00 = x00.data();
  s.displ = displ.data();
  s.x_10 = x_10.data();
  s.disp_10 = disp_10.data();
  s.wall_position = wall_position.data();
  s.horizontal_or_not = horizontal_or_not.data();
  s.n_int = n_int.data();
  s.bc_int = bc_int.data();
  s.if_out_domain = if_out_domain.data();
  s.bc_or_not = bc_or_not.data();
  s.bc_info = bc_info.data();
  return s;
}

Problem load_problem(const std::string &dir, int variant) {
  Problem P;
  spsph_params &p = P.p;
  p.struct_bytes = (int32_t)sizeof(spsph_params);
  p.variant = variant;
  P.dir = dir;
  auto chk = [&](const std::string &s) { P.chk.push_back(s); };

  ListReader inp(dir + "/input.txt");  // unit 997, 1_SPH_2018.f90:120
  inp.text();                          // mat:64
  P.name = inp.text();                 // mat:65 problem_name
  ListReader dat(dir + "/" + P.name + ".dat");
  ListReader pts(dir + "/" + P.name + ".pts");

  // ---- problem_input_data, mat:77-160
  int nline = dat.ints(1)[0];
  chk(std::to_string(nline));
  for (int i = 0; i < nline; ++i) chk(dat.text());
  chk(dat.text());
  p.ndimn = dat.ints(1)[0];
  chk(std::to_string(p.ndimn));
  if (p.ndimn != 2) throw std::runtime_error("only ndimn = 2 is supported (all shipped inputs)");
  chk(dat.text());
  int nmats;
  if (variant == SPSPH_VARIANT_BUI) {  // mat:99 (Bui copy reads three integers)
    auto v = dat.ints(3);
    p.ntype_solid = v[0];
    p.nstre = v[1];
    nmats = v[2];
  } else {  // code/3_SPH_material_2018.f90:101
    auto v = dat.ints(2);
    p.nstre = v[0];
    nmats = v[1];
    // ntype_solid is never assigned in the other copies: module variable in .bss == 0.
    // The SL and code/ copies carry no ntype_solid branches, which is the plane-strain path (== 2).
    p.ntype_solid = (variant == SPSPH_VARIANT_VS) ? 0 : 2;
  }
  if (p.nstre != 4) throw std::runtime_error("only nstre = 4 (plane strain) is supported");
  if (nmats != 1) throw std::runtime_error("only nmats = 1 is supported");
  chk(dat.text());
  for (int im = 0; im < nmats; ++im) {  // mat:150-160
    auto v = dat.reals(13);
    for (int k = 0; k < 12; ++k) p.props[k] = v[k + 1];
    int ncrit = (int)p.props[1];
    if (ncrit == 5 || ncrit == 12) {
      chk(dat.text());
      auto e = dat.reals(SPSPH_NPROP - 12);
      for (int k = 12; k < SPSPH_NPROP; ++k) p.props[k] = e[k - 12];
    }
  }
  p.ntype_eco = (int)p.props[0];
  p.ncrit = (int)p.props[1];
  p.pi = (double)(4 * std::atan(1.0f));  // mat:163, single-precision atan

  // ---- setup_particles -> Read_2DMesh, mat:344-751
  chk(dat.text());
  dat.ints(1);  // icunk_s
  inp.text();   // mat:383
  {
    auto l = inp.logicals(3);  // SP_SPH, art_stress, particle_shift
    p.sp_sph = l[0];
    p.art_stress = l[1];
  }
  double rx_factor = 0, ry_factor = 0;  // uninitialised in the reference when not read (App. C-5)
  if (p.sp_sph) {
    inp.text();
    p.inside_approach = inp.logicals(1)[0];
    inp.text();
    p.npoints = inp.ints(1)[0];
    if (!p.inside_approach) {
      inp.text();
      auto t = inp.tokens(6);  // SPH_shift, vel_vector, shift_update, rx_factor, ry_factor, disp_tol
      p.sph_shift = (t[0][0] == 'T' || t[0][0] == 't' || (t[0].size() > 1 && (t[0][1] == 'T' || t[0][1] == 't')));
      p.vel_vector = (t[1][0] == 'T' || t[1][0] == 't' || (t[1].size() > 1 && (t[1][1] == 'T' || t[1][1] == 't')));
      p.shift_update = std::atoi(t[2].c_str());
      rx_factor = std::strtod(t[3].c_str(), nullptr);
      ry_factor = std::strtod(t[4].c_str(), nullptr);
      p.disp_tol = std::strtod(t[5].c_str(), nullptr);
    }
  }
  if (p.vel_vector) p.npoints = 2;  // mat:396
  if (!p.sp_sph) p.npoints = 1;     // mat:397

  chk(pts.text());
  int geom_type = pts.ints(1)[0];
  (void)geom_type;
  chk(pts.text());
  auto xr = pts.reals(5);  // x1 x2 x3 x4 dx
  chk(join(xr));
  chk(pts.text());
  auto yr = pts.reals(5);  // y1 y2 y3 y4 dy
  chk(join(yr));
  chk(pts.text());
  p.dummy_nodes = pts.logicals(1)[0];
  const double x1 = xr[0], x4 = xr[3], y1 = yr[0], y4 = yr[3];
  const double dx = xr[4], dy = yr[4];
  p.dx = dx;
  p.dy = dy;
  const double lx = x4 - x1, ly = y4 - y1;
  const int ndivx = (int)(lx / dx + 1);  // mat:434 (integer truncation of the fp64 value)
  const int ndivy = (int)(ly / dy + 1);
  P.ndivx = ndivx;
  P.ndivy = ndivy;
  const int nnode = ndivx * ndivy;
  P.nelem = (ndivx - 1) * (ndivy - 1);
  int nstress;
  if (p.inside_approach)
    nstress = p.npoints * (ndivx - 1) * (ndivy - 1);  // mat:443-453
  else
    nstress = p.npoints * nnode;
  const int ntotal = nnode + nstress;
  p.nnode = nnode;
  p.nstress = nstress;
  p.ntotal = ntotal;

  inp.text();
  p.sml = inp.reals(1)[0];  // mat:466-467
  const double sml = p.sml;
  const double rho_mat = p.props[5];

  // ---- set_up_dummy_nodes, mat:755-897 (all wall geometry in default REAL = fp32)
  std::vector<double> x_dummy;
  std::vector<float> hor_dummy, wallpos_dummy;
  int ndummy = 0;
  if (p.dummy_nodes) {
    const int nrow = 3;
    const float dx2 = (float)dx, dy2 = (float)dy;
    chk(pts.text());
    const int n_walls = pts.ints(1)[0];
    std::vector<int> wall_id(n_walls), ndiv_wall(n_walls);
    std::vector<float> wall_position2(n_walls), wx1(n_walls), wx2(n_walls);
    int ndummy2 = 0;
    for (int i = 0; i < n_walls; ++i) {
      pts.text();
      pts.text();
      auto t = pts.tokens(4);
      wall_id[i] = std::atoi(t[0].c_str());
      wall_position2[i] = std::strtof(t[1].c_str(), nullptr);
      wx1[i] = std::strtof(t[2].c_str(), nullptr);
      wx2[i] = std::strtof(t[3].c_str(), nullptr);
      const float l_wall = wx2[i] - wx1[i];
      ndiv_wall[i] = (int)(l_wall / dx2 + 1);  // mat:800, fp32 arithmetic then truncation
      ndummy2 += ndiv_wall[i];
    }
    ndummy = nrow * ndummy2;
    p.ndummy2 = ndummy2;
    x_dummy.assign(2 * (size_t)ndummy, 0.0);
    hor_dummy.assign(ndummy, 0.f);
    wallpos_dummy.assign(ndummy, 0.f);
    int k = 0;
    // innermost layer of every wall first (mat:821-837) ...
    for (int i = 0; i < n_walls; ++i)
      for (int j = 1; j <= ndiv_wall[i]; ++j) {
        if (wall_id[i] == 1 || wall_id[i] == 3) {
          x_dummy[2 * k] = wall_position2[i];
          x_dummy[2 * k + 1] = wx1[i] + (float)(j - 1) * dy2;
          hor_dummy[k] = (wall_id[i] == 1) ? 2.f : 22.f;
        } else if (wall_id[i] == 2) {
          x_dummy[2 * k] = wx1[i] + (float)(j - 1) * dx2;
          x_dummy[2 * k + 1] = wall_position2[i];
          hor_dummy[k] = 1.f;
        }
        wallpos_dummy[k] = wall_position2[i];
        ++k;
      }
    // ... then the outer layers, wall-major (mat:841-859)
    for (int i = 0; i < n_walls; ++i)
      for (int m = 2; m <= nrow; ++m)
        for (int j = 1; j <= ndiv_wall[i]; ++j) {
          if (wall_id[i] == 1 || wall_id[i] == 3) {
            x_dummy[2 * k] = wall_position2[i] - (float)(m - 1) * dx2;
            x_dummy[2 * k + 1] = wx1[i] + (float)(j - 1) * dy2;
            hor_dummy[k] = (wall_id[i] == 1) ? 2.f : 22.f;
          } else if (wall_id[i] == 2) {
            x_dummy[2 * k] = wx1[i] + (float)(j - 1) * dx2;
            x_dummy[2 * k + 1] = wall_position2[i] - (float)(m - 1) * dy2;
            hor_dummy[k] = 1.f;
          }
          wallpos_dummy[k] = wall_position2[i];
          ++k;
        }
    if (k != ndummy) throw std::runtime_error("dummy particle count mismatch");
  }
  p.ndummy = ndummy;
  const int ntotal2 = ntotal + ndummy;
  p.ntotal2 = ntotal2;

  // ---- nodes, mat:517-533
  std::vector<double> x_1(2 * (size_t)nnode), mass_1(nnode);
  std::vector<int> no_int_node(nnode);
  {
    int k = 0;
    for (int i = 1; i <= ndivx; ++i)
      for (int j = 1; j <= ndivy; ++j) {
        const double xx = x1 + (i - 1) * dx, yy = y1 + (j - 1) * dy;
        x_1[2 * k] = xx;
        x_1[2 * k + 1] = yy;
        no_int_node[k] = (xx == x1 || xx == x4 || yy == y4) ? 4 : 8;
        if ((xx == x1 && (yy == y1 || yy == y4)) || (xx == x4 && (yy == y1 || yy == y4))) no_int_node[k] = 2;
        ++k;
      }
  }
  // ---- stress particles, mat:560-676
  std::vector<double> x_s(2 * (size_t)nstress);
  auto X1 = [&](int d, int i) -> double { return x_1[2 * (size_t)(i - 1) + (d - 1)]; };  // 1-based
  if (p.inside_approach) {
    if (p.npoints == 1) {
      int k = 0;
      for (int i = 2; i <= ndivx; ++i)
        for (int j = 2; j <= ndivy; ++j) {
          x_s[2 * k] = x1 - dx / 2. + (i - 1) * dx;
          x_s[2 * k + 1] = y1 - dx / 2. + (j - 1) * dy;
          ++k;
        }
    } else {
      int k = 0;
      for (int j = 1; j <= ndivx - 1; ++j)
        for (int i = 1; i <= ndivy - 1; ++i) {
          const int o = (j - 1) * ndivy;
          if (p.npoints == 2) {  // mat:578-582
            x_s[2 * k] = (X1(1, i + o) + X1(1, i + 1 + o) + X1(1, i + ndivy + o)) / 3.;
            x_s[2 * k + 1] = (X1(2, i) + X1(2, i + 1) + X1(2, i + ndivy)) / 3.;
            x_s[2 * (k + 1)] = (X1(1, i + 1 + o) + X1(1, i + ndivy + o) + X1(1, i + ndivy + o)) / 3.;
            x_s[2 * (k + 1) + 1] = (X1(2, i + 1) + X1(2, i + ndivy) + X1(2, i + ndivy + 1)) / 3.;
            k += 2;
          } else if (p.npoints == 3) {  // mat:592-600
            x_s[2 * k] = (X1(1, i + o) + X1(1, i + 1 + o) + X1(1, i + ndivy + o)) / 3.;
            x_s[2 * k + 1] = (X1(2, i) + X1(2, i + 1) + X1(2, i + ndivy)) / 3.;
            x_s[2 * (k + 1)] = (X1(1, i + o) + X1(1, i + ndivy + o) + X1(1, i + 1 + ndivy + o)) / 3.;
            x_s[2 * (k + 1) + 1] = (X1(2, i) + X1(2, i + ndivy) + X1(2, i + ndivy + 1)) / 3.;
            x_s[2 * (k + 2)] =
                (X1(1, i + 1 + o) + X1(1, i + 1 + ndivy + o) + (X1(1, i + ndivy + o) + X1(1, i + o)) / 2.) / 3.;
            x_s[2 * (k + 2) + 1] = (X1(2, i + 1) + X1(2, i + 1 + ndivy) + (X1(2, i) + X1(2, i + 1)) / 2.) / 3.;
            k += 3;
          } else if (p.npoints == 4) {  // mat:610-618
            x_s[2 * k] = (X1(1, i + o) + X1(1, i + o) + X1(1, i + 1 + ndivy + o)) / 3.;
            x_s[2 * k + 1] = (X1(2, i) + X1(2, i + 1) + X1(2, i + ndivy)) / 3.;
            x_s[2 * (k + 1)] = (X1(1, i + 1 + o) + X1(1, i + ndivy + o) + X1(1, i + ndivy + 1 + o)) / 3.;
            x_s[2 * (k + 1) + 1] = (X1(2, i + 1) + X1(2, i + ndivy) + X1(2, i + ndivy + 1)) / 3.;
            x_s[2 * (k + 2)] = (X1(1, i + o) + X1(1, i + 1 + o) + X1(1, i + 1 + ndivy + o)) / 3.;
            x_s[2 * (k + 2) + 1] = (X1(2, i) + X1(2, i + 1) + X1(2, i + 1 + ndivy)) / 3.;
            x_s[2 * (k + 3)] = (X1(1, i + o) + X1(1, i + ndivy + o) + X1(1, i + 1 + ndivy + o)) / 3.;
            x_s[2 * (k + 3) + 1] = (X1(2, i) + X1(2, i + ndivy) + X1(2, i + ndivy + 1)) / 3.;
            k += 4;
          } else {
            throw std::runtime_error("inside approach supports npoints 1..4");
          }
        }
    }
  } else {
    p.r_x = dx * rx_factor;  // mat:628
    p.r_y = dx * ry_factor;
    const double r_x = p.r_x, r_y = p.r_y;
    int k = 0;
    for (int i = 1; i <= nnode; ++i) {
      const double xi = X1(1, i), yi = X1(2, i);
      auto put = [&](int kk, double a, double b) {
        x_s[2 * (size_t)kk] = a;
        x_s[2 * (size_t)kk + 1] = b;
      };
      if (p.npoints == 1) {
        put(k, xi + r_x, yi + r_y);
      } else if (p.npoints == 2) {
        put(k, xi + r_x, yi + r_y);
        put(k + 1, xi - r_x, yi - r_y);
      } else if (p.npoints == 3) {
        put(k, xi, yi + r_y);
        put(k + 1, xi - r_x, yi - r_y);
        put(k + 2, xi + r_x, yi - r_y);
      } else if (p.npoints == 4) {
        put(k, xi - r_x, yi - r_y);
        put(k + 1, xi - r_x, yi + r_y);
        put(k + 2, xi + r_x, yi + r_y);
        put(k + 3, xi + r_x, yi - r_y);
      } else {
        throw std::runtime_error("outside approach supports npoints 1..4");
      }
      k += p.npoints;
    }
  }
  // ---- density, mass, smoothing length, mat:682-729
  const double area = (rho_mat * dx * dy) / 8;
  for (int i = 0; i < nnode; ++i) mass_1[i] = no_int_node[i] * area;
  std::vector<double> mass_s(nstress);
  if (p.inside_approach) {
    for (int i = 0; i < nstress; ++i) mass_s[i] = (8 * area) / p.npoints;
  } else {
    int k = 0;
    for (int i = 0; i < nnode; ++i) {
      for (int j = 0; j < p.npoints; ++j) mass_s[k + j] = mass_1[i] / p.npoints;
      k += p.npoints;
    }
  }
  // ---- elastic constants, mat:738-749 (4./3. and 2./3. are fp32 literals)
  {
    const double young = p.props[2], poiss = p.props[3];
    const double K_mod = young / (3 * (1 - 2 * poiss));
    const double G_mod = young / (2 * (1 + poiss));
    p.D11 = (double)(4.f / 3.f) * G_mod + K_mod;
    p.D22 = p.D11;
    p.D12 = -(double)(2.f / 3.f) * G_mod + K_mod;
    p.D33 = G_mod;
    p.D41 = p.D12;
    p.D42 = p.D12;
  }

  // ---- Setup_Global_Arrays, mat:902-1045
  const size_t n2 = (size_t)ntotal2;
  P.x.assign(2 * n2, 0.0);
  P.vel.assign(2 * n2, 0.0);
  P.stress.assign(4 * n2, 0.0);
  P.rho.assign(n2, 0.0);
  P.mass.assign(n2, 0.0);
  P.hsml.assign(n2, 0.0);
  P.itype.assign(n2, 0);
  P.if_out_domain.assign(n2, 0);
  P.wall_position.assign(n2, 0.f);
  P.horizontal_or_not.assign(n2, 0.f);
  P.f_drucker.assign(ntotal, 0.0);
  P.internal_vars.assign((size_t)SPSPH_NINT_VARS * ntotal, 0.0);
  P.displ.assign(2 * (size_t)nnode, 0.0);
  P.disp_10.assign(nnode, 0.0);
  P.n_int.assign(nnode, 0.f);
  P.bc_int.assign(nnode, 0);
  const double h0 = sml * dx;
  for (int i = 0; i < nnode; ++i) {
    P.x[2 * (size_t)i] = x_1[2 * (size_t)i];
    P.x[2 * (size_t)i + 1] = x_1[2 * (size_t)i + 1];
    P.itype[i] = 2;
    P.rho[i] = rho_mat;
    P.mass[i] = mass_1[i];
    P.hsml[i] = h0;
  }
  for (int i = 0; i < nstress; ++i) {
    const size_t g = (size_t)nnode + i;
    P.x[2 * g] = x_s[2 * (size_t)i];
    P.x[2 * g + 1] = x_s[2 * (size_t)i + 1];
    P.itype[g] = 1;
    P.rho[g] = rho_mat;
    P.mass[g] = mass_s[i];
    P.hsml[g] = h0;
  }
  if (p.dummy_nodes) {
    const float dx2 = (float)dx;
    const double mass_dummy = (double)(dx2 * dx2) * rho_mat;  // mat:867
    for (int i = 0; i < ndummy; ++i) {
      const size_t g = (size_t)ntotal + i;
      P.x[2 * g] = x_dummy[2 * (size_t)i];
      P.x[2 * g + 1] = x_dummy[2 * (size_t)i + 1];
      P.itype[g] = 25;
      P.rho[g] = rho_mat;
      P.mass[g] = mass_dummy;
      P.hsml[g] = h0;
      P.wall_position[g] = wallpos_dummy[i];
      P.horizontal_or_not[g] = hor_dummy[i];
    }
  }
  P.x00 = P.x;
  P.x_10.assign(P.x.begin(), P.x.begin() + 2 * (size_t)nnode);  // mat:1032

  // ---- boundary conditions, mat:182-231
  chk(dat.text());
  {
    auto v = dat.ints(2);
    p.no_bcs = v[0];
    p.ifsigman = v[1];
  }
  if (p.no_bcs > SPSPH_MAX_BCS) throw std::runtime_error("too many BCs");
  if (p.no_bcs > 0) {
    const double twopi = (double)(std::acos(0.0f) * 4.0f / 360.f);  // mat:191 (fp32 expression)
    chk(dat.text());
    for (int i = 0; i < p.no_bcs; ++i) {
      auto v = dat.reals(8);
      for (int j = 0; j < 8; ++j) p.bc_list[i][j] = v[j];
      p.bc_list[i][6] = p.bc_list[i][6] * twopi;
    }
  }
  chk(dat.text());
  const int no_segments_bc = dat.ints(1)[0];
  std::vector<std::vector<double>> seg;
  if (no_segments_bc > 0) {
    chk(dat.text());
    for (int i = 0; i < no_segments_bc; ++i) seg.push_back(dat.reals(5));
  }
  chk(dat.text());
  const int no_nodes_bc = dat.ints(1)[0];
  std::vector<std::vector<double>> nodal;
  if (no_nodes_bc > 0) {
    chk(dat.text());
    for (int i = 0; i < no_nodes_bc; ++i) nodal.push_back(dat.reals(2));
  }
  // Get_BCs_on_node, mat:1049-1131
  P.bc_info.assign(8 * (size_t)ntotal, 0);
  P.bc_or_not.assign(ntotal, 0);
  for (int ip = 0; ip < ntotal; ++ip) P.bc_info[8 * (size_t)ip] = ip + 1;
  for (int is = 0; is < no_segments_bc; ++is) {
    const double xx1 = seg[is][0], yy1 = seg[is][1], xx2 = seg[is][2], yy2 = seg[is][3];
    const double xmax = std::fmax(xx1, xx2), xmin = std::fmin(xx1, xx2);
    const double ymax = std::fmax(yy1, yy2), ymin = std::fmin(yy1, yy2);
    const int no_bc = (int)seg[is][4];
    for (int ip = 0; ip < ntotal; ++ip) {
      const double xx = P.x[2 * (size_t)ip], yy = P.x[2 * (size_t)ip + 1];
      const double z = (yy1 - yy2) * (xx - xx1) + (xx2 - xx1) * (yy - yy1);
      if (z == 0. && xx >= xmin && xx <= xmax && yy >= ymin && yy <= ymax) {
        int32_t *bi = &P.bc_info[8 * (size_t)ip];
        bi[1] += 1;
        const int k = bi[1];
        if (k + 2 > 8) throw std::runtime_error("more than 6 BCs on one particle");
        bi[k + 1] = no_bc;  // bc_info(k+2,ipoin)
        if (P.bc_or_not[ip] == 0) P.bc_or_not[ip] = 1;
      }
    }
  }
  for (int in = 0; in < no_nodes_bc; ++in) {  // mat:1110-1128 (without the interactive conflict prompt)
    const int ic_node = (int)nodal[in][0], no_bc = (int)nodal[in][1];
    const int bc_var = (int)p.bc_list[no_bc - 1][1];
    int32_t *bi = &P.bc_info[8 * (size_t)(ic_node - 1)];
    bi[1] += 1;
    const int k = bc_var + 2;
    if (k < 1 || k > 8) throw std::runtime_error("nodal BC variable out of range");
    if (bi[k - 1] != 0 && bi[k - 1] != no_bc)
      throw std::runtime_error("conflicting nodal BCs (the reference prompts interactively here)");
    bi[k - 1] = no_bc;
    if (P.bc_or_not[ic_node - 1] == 0) P.bc_or_not[ic_node - 1] = 1;
  }

  // ---- time curves, mat:234-264
  chk(dat.text());
  const int ic_tcurve = dat.ints(1)[0];
  if (ic_tcurve == 1) {
    chk(dat.text());
    auto v = dat.ints(2);
    p.ntcurves = v[0];
    if (p.ntcurves > SPSPH_MAX_TCURVES) throw std::runtime_error("too many time curves");
    for (int i = 0; i < p.ntcurves; ++i) {
      chk(dat.text());
      p.nptstcurves[i] = dat.ints(1)[0];
      if (p.nptstcurves[i] > SPSPH_MAX_TCURVE_PTS) throw std::runtime_error("time curve too long");
      chk(dat.text());
      auto t = dat.reals(p.nptstcurves[i]);
      auto f = dat.reals32(p.nptstcurves[i]);
      for (int j = 0; j < p.nptstcurves[i]; ++j) {
        p.ttcurves[i][j] = t[j];
        p.ftcurves[i][j] = f[j];
      }
    }
  }
  // ---- domain limits, mat:288-291
  chk(dat.text());
  {
    auto v = dat.reals(4);
    p.xmin_domain[0] = v[0];
    p.xmin_domain[1] = v[1];
    p.xmax_domain[0] = v[2];
    p.xmax_domain[1] = v[3];
  }
  // ---- Initial_conditions, mat:1433-1569
  chk(dat.text());
  const int icunkno = dat.ints(1)[0];
  if (icunkno == 2) {
    for (int k = 0; k < p.nstre + p.ndimn; ++k) {  // Get_Init2D: every shipped ICtype is 0 -> zero fields
      chk(dat.text());
      dat.ints(1);
    }
  } else if (icunkno == 30) {  // Get_Init_sigma0, constant state
    chk(dat.text());
    const int type_sigma0 = dat.ints(1)[0];
    if (type_sigma0 == 1) {
      chk(dat.text());
      auto s0 = dat.reals(p.nstre);
      chk(dat.text());
      auto v0 = dat.reals(SPSPH_NINT_VARS);
      for (int ip = 0; ip < ntotal; ++ip) {
        for (int a = 0; a < 4; ++a) P.stress[4 * (size_t)ip + a] = s0[a];
        for (int a = 0; a < SPSPH_NINT_VARS; ++a) P.internal_vars[(size_t)SPSPH_NINT_VARS * ip + a] = v0[a];
      }
    }
  }
  // NB Setup_Global_Arrays ends with stress = 0; vel = 0 (mat:1036) *before* Initial_conditions runs.

  // ---- control parameters, mat:302-336
  chk(dat.text());
  {
    auto t = dat.tokens(7);  // pa_sph nnps sle skf CSPM update_x XSPH
    p.sle = std::atoi(t[2].c_str());
    p.skf = std::atoi(t[3].c_str());
    auto lg = [](const std::string &s) {
      size_t k = (s[0] == '.') ? 1 : 0;
      return (s[k] == 'T' || s[k] == 't') ? 1 : 0;
    };
    p.cspm = lg(t[4]);
    p.update_x = lg(t[5]);
    p.xsph = lg(t[6]);
  }
  chk(dat.text());
  {
    auto l = dat.logicals(2);
    p.cont_density = l[1];
  }
  dat.text();
  p.damping = dat.reals(1)[0];
  dat.text();
  {
    auto v = dat.reals(2);
    p.alpha = v[0];
    p.beta = v[1];
  }
  chk(dat.text());
  p.ic_grav = dat.ints(1)[0];
  if (p.ic_grav == 1) {
    chk(dat.text());
    auto v = dat.reals(4);
    p.cgrav[0] = v[0];
    p.cgrav[1] = v[1];
    p.tcurve_grav = (int)v[2];
    p.ft_grav = v[3];
  }
  dat.text();
  dat.text();
  {
    auto v = dat.ints(10);
    for (int k = 0; k < 4; ++k) P.stress_out[k] = v[k];
    P.vel_out[0] = v[4];
    P.vel_out[1] = v[5];
    P.strain_out = v[6];
    P.rho_out = v[7];
    P.sml_out = v[8];
    P.disp_out = v[9];
  }

  // per-copy switches
  p.bc_loop_ntotal = (variant == SPSPH_VARIANT_BUI || variant == SPSPH_VARIANT_VS) ? 1 : 0;
  p.ae_threshold = (variant == SPSPH_VARIANT_BUI) ? 1e-07f : 1e-03f;

  // ---- time blocks of the main program, 1_SPH_2018.f90:141-152
  for (;;) {
    TimeBlock b;
    try {
      inp.text();
      auto v = inp.reals(3);
      b.dt = v[0];
      b.time_end = v[1];
      b.maxtimestep = (int)v[2];
    } catch (const std::exception &) {
      break;
    }
    if (b.dt <= 0) break;
    inp.text();
    auto s = inp.ints(3);
    b.print_step = s[0];
    b.save_step = s[1];
    b.plot_step = s[2];
    P.blocks.push_back(b);
  }
  return P;
}

}  // namespace spsph
