// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product; nothing under stress-particle-sph_b200/
// may include, link or call this file.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, and only as the checker / CPU baseline.
//
// PARITY PINNED against the reference itself. The reference ships no golden vectors or tests and cannot be
// COMPILED here (no Fortran compiler in this image), but its repository ships the authors' own executables
// (example_problems/*/sph, gfortran 4.8.5 -O3, built from the sources beside them) that only lack
// libgfortran.so.3. oracle/gfortran_shim.c supplies that run-time library, oracle/make_reference_goldens.py runs
// the unmodified binaries on every input set the reference ships (100 steps of the three problems, 60 / 40 steps
// of the 11 variants, 602 steps through list growth, 1501 steps into the wall forces) and stores what they print,
// all digits, in tests/golden/ref_*.npz. This file reproduces all of it BIT FOR BIT -- positions, velocities,
// stresses, plastic strain of every velocity and stress particle (tests/test_reference_pinned_cpu.py).
// It is a serial C++ restatement of the reference's time step that follows the Fortran statement
// by statement -- same loop order, same expression order, same fp32/fp64 mix, same
// pair creation and traversal order (App. B), zero-initialised reading of the undefined values (App. C).
//
// Reference files followed ("main:" = code/2_SPH_main_2018.f90, identical in all copies except two
// lines; "mat:" = example_problems/soil_failure_bui_et_al_2008/3_SPH_material_2018.f90 unless a copy is
// named; per-copy differences are runtime switches in spsph_params):
//   time_integration main:78-184        XSPH_update main:189-239     shift_stress_points main:244-368
//   isolated_nodes main:373-398         stress_point_update main:403-482
//   get_derivatives main:487-648        RK4 main:653-802             density_update main:807-821
//   artificial_viscosity main:826-904   get_spin_rate_tensor main:1021-1034
//   Check_Out_Domain main:1170-1194     grid_find_NEW main:1199-1435 kernel main:1440-1538
//   Pint_Update mat:1574-1634           BCs/Normal_BCs mat:1641-1770 update_strain mat:1864-1880
//   apply_stress_free mat:1775-1858
//   plastic_terms mat:1884-1954         drucker_prager mat:1958-2083 adapt_stress2 mat:2087-2161
//   Get_Vivel/invar09/yieldf09/flowvp09/Get_Dmatx (strain_localisation copy) :2169-2576
//   Get_derivative_intvars mat:2697-2760  gravity_force mat:2809-2871
//
// Build: g++ -O2 -ffp-contract=off (no -march, no -ffast-math): IEEE double/float, no FMA contraction,
// which is what `gfortran -O3` emits for x86-64 by default.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "spsph.h"

namespace {

struct Pair {  // TYPE pairs, code/5_SPH_global_vars_2018.f90:14-23
  int32_t pair_i, pair_j, pint_type;
  float w, dwdx, dwdy;
};

struct Oracle {
  spsph_params p;
  int nnode, ntotal, ntotal2;
  // state (1-based Fortran arrays stored 0-based, column-major)
  std::vector<double> x, vel, stress, rho, mass, hsml, internal_vars, f_drucker, x0, x00, vx0, displ, x_10, disp_10;
  std::vector<double> grad_u, art_visc, Ddev_strn, f_bound;
  std::vector<double> subset, normal;  // get_nodes_on_free_surface (module arrays, DP): subset(ntotal), normal(2,ntotal)
  std::vector<int32_t> itype, if_out, bc_or_not, bc_info, bc_int, countiac;
  std::vector<float> wall_position, horizontal_or_not, n_int;
  // pair list: `created` in creation order; traversal order through order_of()
  std::vector<Pair> created;
  std::vector<double> x_at_fs;     // positions get_nodes_on_free_surface saw in the last step (before the re-seating)
  std::vector<int32_t> last_cell;  // which_cell of the last grid_find (1-based reference cell id, 0: out of the domain)
  int64_t m_pairs = 0;       // list capacity = max pair count of all previous steps (main:1210)
  int64_t m_before = 0;      // m_pairs at the start of the current step
  double time_sph = 0, dt_sph = 0;
  int itimestep_sph = 0;
  int maxiac = 0, miniac = 0, noiac = 0;
  std::string err;

  // ---- accessors (1-based) -------------------------------------------------------------------
  double &X(int d, int i) { return x[2 * (size_t)(i - 1) + (d - 1)]; }
  double &V(int d, int i) { return vel[2 * (size_t)(i - 1) + (d - 1)]; }
  double &S(int s, int i) { return stress[4 * (size_t)(i - 1) + (s - 1)]; }
  double &IV1(int i) { return internal_vars[(size_t)SPSPH_NINT_VARS * (i - 1)]; }
  double &GU(int a, int b, int i) { return grad_u[4 * (size_t)(i - 1) + 2 * (b - 1) + (a - 1)]; }  // grad_u(a,b,i)

  // traversal position t (0-based) -> creation index (0-based); SURVEY App. B, main:1361-1383
  inline int64_t creation_index(int64_t t) const {
    const int64_t n = (int64_t)created.size();
    const int64_t M = m_before;
    if (n <= M) return t;
    const int64_t over = n - M;  // pairs beyond the old capacity are prepended: visited first, reversed
    return (t < over) ? (n - 1 - t) : (t - over);
  }

  // ---- kernel, main:1440-1538 (skf = 1, ndimn = 2) ---------------------------------------------
  void kernel(double r, const double dx2[2], double h, double &w, double dwdx[2]) const {
    const double pi = p.pi;
    const double q = r / h;
    w = 0.;
    dwdx[0] = dwdx[1] = 0.;
    if (p.skf == 1) {
      const double factor = 15.e0 / (7.e0 * pi * h * h);
      if (q >= 0 && q <= 1.e0) {
        w = factor * ((double)(2.f / 3.f) - q * q + q * q * q / 2.);
        for (int d = 0; d < 2; ++d) dwdx[d] = factor * (-2. + (double)(3.f / 2.f) * q) / (h * h) * dx2[d];
      } else if (q > 1.e0 && q <= 2) {
        const double t = 2. - q;
        w = factor * 1.e0 / 6.e0 * (t * t * t);
        for (int d = 0; d < 2; ++d) dwdx[d] = -factor * 1.e0 / 6.e0 * 3. * (t * t) / h * (dx2[d] / r);
      }
    } else if (p.skf == 2) {  // Gauss, main:1494-1504
      const double factor = 1.e0 / (std::pow(h, 2) * std::pow(pi, 2 / 2.));
      if (q >= 0 && q <= 3) {
        w = factor * std::exp(-q * q);
        for (int d = 0; d < 2; ++d) dwdx[d] = w * (-2. * dx2[d] / h / h);
      }
    } else if (p.skf == 3) {  // quintic, main:1506-1535
      const double factor = 7.e0 / (478.e0 * pi * h * h);
      // gfortran calls libgcc's __powidf2 for x**5 / x**4: square-and-multiply, x**5 = x*((x*x)*(x*x))
      auto p5 = [](double a) { const double a2 = a * a; return a * (a2 * a2); };
      auto p4 = [](double a) { const double a2 = a * a; return a2 * a2; };
      if (q >= 0 && q <= 1) {
        w = factor * (p5(3 - q) - 6 * p5(2 - q) + 15 * p5(1 - q));
        for (int d = 0; d < 2; ++d) dwdx[d] = factor * ((-120 + 120 * q - 50 * (q * q)) / (h * h) * dx2[d]);
      } else if (q > 1 && q <= 2) {
        w = factor * (p5(3 - q) - 6 * p5(2 - q));
        for (int d = 0; d < 2; ++d) dwdx[d] = factor * (-5 * p4(3 - q) + 30 * p4(2 - q)) / h * (dx2[d] / r);
      } else if (q > 2 && q <= 3) {
        w = factor * p5(3 - q);
        for (int d = 0; d < 2; ++d) dwdx[d] = factor * (-5 * p4(3 - q)) / h * (dx2[d] / r);
      }
    }
  }

  // ---- Check_Out_Domain, main:1170-1194 --------------------------------------------------------
  void check_out_domain() {
    for (int ip = 1; ip <= ntotal2; ++ip)
      for (int d = 1; d <= 2; ++d) {
        const double dxx = (X(d, ip) - p.xmin_domain[d - 1]) * (X(d, ip) - p.xmax_domain[d - 1]);
        if (dxx > 0.0) if_out[ip - 1] = 1;
      }
  }

  // ---- grid_find_NEW, main:1199-1435 -----------------------------------------------------------
  void grid_find() {
    const int scale_k = (p.skf == 1) ? 2 : 3;
    double xmin[2] = {1.e+10, 1.e+10}, xmax[2] = {-1.e+10, -1.e+10}, deltx[2] = {0, 0};
    int ndivx[2] = {1, 1};
    std::fill(countiac.begin(), countiac.end(), 0);
    for (int d = 0; d < 2; ++d) {
      for (int i = 1; i <= ntotal2; ++i) {
        if (if_out[i - 1] == 1) continue;
        if (X(d + 1, i) < xmin[d]) xmin[d] = X(d + 1, i);
        if (X(d + 1, i) > xmax[d]) xmax[d] = X(d + 1, i);
        if (deltx[d] < hsml[i - 1]) deltx[d] = hsml[i - 1];
      }
      deltx[d] = deltx[d] * 2;
      const double length = xmax[d] - xmin[d];
      ndivx[d] = (int)((length / deltx[d]) + 1);
      const double length_new = ndivx[d] * deltx[d];
      xmin[d] = xmin[d] - (length_new - length) / 2 - (double)0.001f * length;
      xmax[d] = xmax[d] + (length_new - length) / 2 + (double)0.001f * length;
    }
    const int64_t ndivt = (int64_t)ndivx[0] * ndivx[1];
    std::vector<int32_t> which_cell(ntotal2, 0), cnt(ndivt + 1, 0), start(ndivt + 2, 0), fill(ndivt + 1, 0),
        list_picell(ntotal2, 0);
    for (int i = 1; i <= ntotal2; ++i) {
      if (if_out[i - 1] == 1) continue;
      int idv[2];
      for (int d = 0; d < 2; ++d) {
        idv[d] = (int)((X(d + 1, i) - xmin[d]) / deltx[d] + 1);
        if (idv[d] > ndivx[d]) idv[d] = ndivx[d];
      }
      const int idivt = ndivx[0] * (idv[1] - 1) + idv[0];
      which_cell[i - 1] = idivt;
      cnt[idivt] += 1;
    }
    int ipost = 1;
    for (int64_t c = 1; c <= ndivt; ++c) {
      start[c] = ipost;
      ipost += cnt[c];
    }
    for (int i = 1; i <= ntotal2; ++i) {
      if (if_out[i - 1] == 1) continue;
      const int c = which_cell[i - 1];
      list_picell[start[c] + fill[c] - 1] = i;
      fill[c] += 1;
    }
    m_before = m_pairs;
    created.clear();
    last_cell = which_cell;
    for (int idy = 1; idy <= ndivx[1]; ++idy)
      for (int idx = 1; idx <= ndivx[0]; ++idx) {
        const int idivt = ndivx[0] * (idy - 1) + idx;
        const int npicsi = cnt[idivt];
        for (int i = 1; i <= npicsi; ++i) {
          const int itotal = list_picell[start[idivt] + i - 2];
          const int idx0 = std::max(1, idx - 1), idx1 = std::min(ndivx[0], idx + 1);
          const int idy1 = std::min(ndivx[1], idy + 1);
          for (int jdy = idy; jdy <= idy1; ++jdy)
            for (int jdx = idx0; jdx <= idx1; ++jdx) {
              const int jdivt = ndivx[0] * (jdy - 1) + jdx;
              if (jdivt < idivt) continue;
              const int jj = (idivt == jdivt) ? i + 1 : 1;
              const int npicsj = cnt[jdivt];
              for (int j = jj; j <= npicsj; ++j) {
                const int jtotal = list_picell[start[jdivt] + j - 2];
                double dxiac[2], tdwdx[2], w;
                dxiac[0] = X(1, itotal) - X(1, jtotal);
                double driac = dxiac[0] * dxiac[0];
                dxiac[1] = X(2, itotal) - X(2, jtotal);
                driac = driac + dxiac[1] * dxiac[1];
                const double mhsml = (hsml[itotal - 1] + hsml[jtotal - 1]) / 2.;
                const double r = std::sqrt(driac);
                if (r < scale_k * mhsml) {
                  countiac[itotal - 1] += 1;
                  countiac[jtotal - 1] += 1;
                  kernel(r, dxiac, mhsml, w, tdwdx);
                  Pair pr;
                  pr.pair_i = itotal;
                  pr.pair_j = jtotal;
                  pr.pint_type = 0;
                  pr.w = (float)w;
                  pr.dwdx = (float)tdwdx[0];
                  pr.dwdy = (float)tdwdx[1];
                  created.push_back(pr);
                }
              }
            }
        }
      }
    if ((int64_t)created.size() > m_pairs) m_pairs = (int64_t)created.size();
    // statistics, main:1405-1422
    maxiac = 0;
    miniac = 1000;
    noiac = 0;
    for (int i = 0; i < ntotal2; ++i) {
      if (countiac[i] > maxiac) maxiac = countiac[i];
      if (countiac[i] < miniac) miniac = countiac[i];
      if (countiac[i] == 0) noiac += 1;
    }
  }

  // ---- Pint_Update, mat:1574-1634 --------------------------------------------------------------
  void pint_update() {
    for (Pair &c : created) {  // order-independent
      const int i = c.pair_i, j = c.pair_j;
      const int ii = std::abs(itype[i - 1]), jj = std::abs(itype[j - 1]);
      const int isumm = itype[i - 1] + itype[j - 1];
      if ((ii == 2 && jj == 1) || (ii == 2 && jj == 25) || (ii == 1 && jj == 25)) {
        c.pair_i = j;
        c.pair_j = i;
        c.dwdx = -c.dwdx;
        c.dwdy = -c.dwdy;
      }
      if (isumm == 3)
        c.pint_type = 1;  // velocity particle -- stress particle
      else if (isumm == 2)
        c.pint_type = 2;  // stress -- stress
      else if (isumm == 4)
        c.pint_type = 3;  // node -- node
      else if (isumm == 27)
        c.pint_type = 6;  // node -- dummy
      else if (isumm == 26)
        c.pint_type = 9;  // stress particle -- dummy
    }
  }

  // ---- time curve evaluation shared by Normal_BCs (mat:1707-1722) and gravity_force (mat:2688-2704)
  double tcurve(int it_curves, double t) const {
    double tt0 = 0, tt1 = 0;
    int ipts = 1;
    if (it_curves < 1 || it_curves > p.ntcurves) return 0.0;
    const int npts = p.nptstcurves[it_curves - 1];
    for (ipts = 1; ipts <= npts - 1; ++ipts) {
      tt0 = p.ttcurves[it_curves - 1][ipts - 1];
      tt1 = p.ttcurves[it_curves - 1][ipts];
      if (t >= tt0 && t <= tt1) break;
      tt0 = -1000.;
    }
    if (tt0 >= 0.0) {
      const double xi = (t - tt0) / (tt1 - tt0);
      return (1. - xi) * (double)p.ftcurves[it_curves - 1][ipts - 1] + xi * (double)p.ftcurves[it_curves - 1][ipts];
    }
    return 0.0;
  }

  // ---- BCs -> Normal_BCs, mat:1641-1770 --------------------------------------------------------
  void bcs() {
    if (p.no_bcs <= 0) return;
    const double ic_time = 0.0;  // uninitialised in the reference: zero reading
    const double t_actual = time_sph + ic_time * dt_sph;
    const int nloop = p.bc_loop_ntotal ? ntotal : nnode;
    for (int ip = 1; ip <= nloop; ++ip) {
      const int32_t *bi = &bc_info[8 * (size_t)(ip - 1)];
      const int nber_BC = bi[1];
      if (bc_or_not[ip - 1] != 1) continue;
      if (nber_BC == 0) continue;
      for (int i = 1; i <= 5; ++i) {
        const int bc_type = bi[i + 1];  // bc_info(i+2,ipoin)
        if (bc_type == 0) continue;
        const double *bl = p.bc_list[bc_type - 1];
        const int it_curves = (int)bl[2];
        double bc_value = 0.0;
        if (it_curves == 0) {
          const double a0 = bl[4], a1 = bl[3], w = bl[5], phi = bl[6], tt = bl[7];
          const double fact = 1.0 - std::exp(-t_actual / tt);
          const double argum = w * t_actual - phi;
          bc_value = (a0 + a1 * std::sin(argum)) * fact;
        } else if (it_curves > 0) {
          const double a1 = bl[3];
          bc_value = tcurve(it_curves, t_actual);
          bc_value = bc_value * a1;
        }
        const double bc_var = bl[1];
        if (bc_var == 5)
          V(1, ip) = bc_value;
        else if (bc_var == 6)
          V(2, ip) = bc_value;
        else if (bc_var == 1)
          S(1, ip) = bc_value;
        else if (bc_var == 3)
          S(3, ip) = bc_value;
        else if (bc_var == 2)
          S(2, ip) = bc_value;
      }
    }
    if (p.ifsigman == 1) apply_stress_free();
  }

  // ---- apply_stress_free, vertical_slope copy mat:1756-1839 (identical in all copies) -----------------
  // Velocity particles that get_nodes_on_free_surface marked (bc_or_not == 2) at the end of the previous step and
  // that are not next to a wall (bc_int /= 1) keep only the stress component tangential to the surface:
  // sigma_tt = n_y^2 sxx - 2 n_x n_y sxy + n_x^2 syy, rotated back; szz <- n_x^2 sigma_tt.
  void apply_stress_free() {
    for (int ip = 1; ip <= nnode; ++ip) {
      if (!(bc_or_not[ip - 1] == 2 && bc_int[ip - 1] != 1)) continue;
      const double costh = normal[2 * (size_t)(ip - 1)], sinth = normal[2 * (size_t)(ip - 1) + 1];
      if (std::isnan(costh) || std::isnan(sinth)) continue;
      const double s2 = sinth * sinth, c2 = costh * costh, sc = sinth * costh;
      const double sxx0 = S(1, ip), syy0 = S(2, ip), sxy0 = S(3, ip);
      const double sigmatt = s2 * sxx0 - 2 * sc * sxy0 + c2 * syy0;
      S(1, ip) = s2 * sigmatt;
      S(2, ip) = c2 * sigmatt;
      S(3, ip) = -sc * sigmatt;
      S(4, ip) = c2 * sigmatt;
    }
  }

  // ---- adapt_stress2, mat:2087-2161 (fp64) -----------------------------------------------------
  void adapt_stress2() {
    const double tanfi = p.props[12], coh = p.props[13];
    const double alpha2 = tanfi / (std::sqrt(9 + 12 * (tanfi * tanfi)));
    const double kc = (3 * coh) / (std::sqrt(9 + 12 * (tanfi * tanfi)));
    for (int i = 1; i <= ntotal; ++i) {
      double smean = (S(1, i) + S(2, i) + S(4, i)) / 3.0;
      double devia[5];
      devia[1] = S(1, i) - smean;
      devia[2] = S(2, i) - smean;
      devia[3] = S(3, i);
      devia[4] = S(4, i) - smean;
      double varj2 = devia[3] * devia[3] + 0.5 * (devia[1] * devia[1] + devia[2] * devia[2] + devia[4] * devia[4]);
      double yield = -alpha2 * 3 * smean + kc;
      if (yield < 0) {
        S(1, i) = S(1, i) - smean + kc / (3 * alpha2);
        S(2, i) = S(2, i) - smean + kc / (3 * alpha2);
        S(4, i) = S(4, i) - smean + kc / (3 * alpha2);
        smean = (S(1, i) + S(2, i) + S(4, i)) / 3.0;
        devia[1] = S(1, i) - smean;
        devia[2] = S(2, i) - smean;
        devia[3] = S(3, i);
        devia[4] = S(4, i) - smean;
        varj2 = devia[3] * devia[3] + 0.5 * (devia[1] * devia[1] + devia[2] * devia[2] + devia[4] * devia[4]);
        yield = -alpha2 * 3 * smean + kc;
      }
      if (yield < std::sqrt(varj2)) {
        double rn = (-3 * alpha2 * smean + kc) / (std::sqrt(varj2));
        if (std::sqrt(varj2) <= (double)10e-06f) rn = 0;
        S(1, i) = rn * devia[1] + smean;
        S(2, i) = rn * devia[2] + smean;
        S(4, i) = rn * devia[4] + smean;
        S(3, i) = rn * devia[3];
      }
    }
  }

  // ---- stress_point_update, main:403-482 -------------------------------------------------------
  std::vector<double> vel_temp, stress_temp, iv_temp, rho_temp, cspm_norm;
  void stress_point_update() {
    if (p.sp_sph) {
      vel_temp.assign(2 * (size_t)ntotal, 0.0);
      stress_temp.assign(4 * (size_t)ntotal, 0.0);
      iv_temp.assign(ntotal, 0.0);
      rho_temp.assign(ntotal, 0.0);
      cspm_norm.assign(ntotal, 0.0);
      const int64_t n = (int64_t)created.size();
      for (int64_t t = 0; t < n; ++t) {
        const Pair &c = created[creation_index(t)];
        if (c.pint_type != 1) continue;
        const int j = c.pair_i, i = c.pair_j;  // j stress particle, i node
        const double w = (double)c.w;
        const double h1 = (mass[j - 1] / rho[j - 1]) * w;
        const double h2 = (mass[i - 1] / rho[i - 1]) * w;
        vel_temp[2 * (size_t)(j - 1)] += V(1, i) * h2;
        vel_temp[2 * (size_t)(j - 1) + 1] += V(2, i) * h2;
        for (int s = 1; s <= 4; ++s) stress_temp[4 * (size_t)(i - 1) + s - 1] += S(s, j) * h1;
        iv_temp[i - 1] += IV1(j) * h1;
        if (p.cont_density) rho_temp[i - 1] += rho[j - 1] * h1;
        cspm_norm[j - 1] += (w * mass[i - 1]) / rho[i - 1];
        cspm_norm[i - 1] += (w * mass[j - 1]) / rho[j - 1];
      }
      for (int i = nnode + 1; i <= ntotal; ++i) {
        if (cspm_norm[i - 1] == 0) continue;
        V(1, i) = vel_temp[2 * (size_t)(i - 1)] / cspm_norm[i - 1];
        V(2, i) = vel_temp[2 * (size_t)(i - 1) + 1] / cspm_norm[i - 1];
      }
      for (int i = 1; i <= nnode; ++i) {
        if (cspm_norm[i - 1] != 0) {
          for (int s = 1; s <= 4; ++s) S(s, i) = stress_temp[4 * (size_t)(i - 1) + s - 1] / cspm_norm[i - 1];
          IV1(i) = iv_temp[i - 1] / cspm_norm[i - 1];
        } else {
          V(1, i) = 0;
          V(2, i) = 0;
        }
      }
      if (p.cont_density)
        for (int i = 1; i <= nnode; ++i) rho[i - 1] = rho_temp[i - 1] / cspm_norm[i - 1];
    } else {  // standard SPH: nodes = stress particles, main:472-480
      for (int i = 1; i <= nnode; ++i) {
        V(1, nnode + i) = V(1, i);
        V(2, nnode + i) = V(2, i);
        for (int s = 1; s <= 4; ++s) S(s, i) = S(s, nnode + i);
        IV1(i) = IV1(nnode + i);
        if (p.cont_density) rho[i - 1] = rho[nnode + i - 1];
      }
    }
  }

  // ---- get_derivatives(vel, divf1, stress, divf2), main:487-648 -----------------------------------
  std::vector<double> grad1, grad2, AE, divf1, divf2;
  void get_derivatives() {
    grad1.assign(4 * (size_t)ntotal, 0.0);  // grad1_tmp(a,b,i) at 4*(i-1)+2*(b-1)+(a-1)
    grad2.assign(6 * (size_t)ntotal, 0.0);  // grad2_tmp(s,b,i) at 6*(i-1)+3*(b-1)+(s-1)
    AE.assign(5 * (size_t)ntotal, 0.0);
    divf1.assign(4 * (size_t)ntotal, 0.0);
    divf2.assign(2 * (size_t)ntotal, 0.0);
    std::fill(bc_int.begin(), bc_int.end(), 0);
    auto G1 = [&](int a, int b, int i) -> double & { return grad1[4 * (size_t)(i - 1) + 2 * (b - 1) + (a - 1)]; };
    auto G2 = [&](int s, int b, int i) -> double & { return grad2[6 * (size_t)(i - 1) + 3 * (b - 1) + (s - 1)]; };
    auto A = [&](int k, int i) -> double & { return AE[5 * (size_t)(i - 1) + (k - 1)]; };
    const int64_t n = (int64_t)created.size();
    for (int64_t t = 0; t < n; ++t) {
      const Pair &c = created[creation_index(t)];
      const int j = c.pair_i, i = c.pair_j;
      const double dwdx = (double)c.dwdx, dwdy = (double)c.dwdy;
      if (c.pint_type == 1) {  // j stress particle, i node
        const double h1 = dwdx * mass[i - 1] / rho[i - 1];
        const double h2 = dwdy * mass[i - 1] / rho[i - 1];
        for (int d = 1; d <= 2; ++d) {
          G1(d, 1, j) = G1(d, 1, j) + (V(d, i) - V(d, j)) * h1;
          G1(d, 2, j) = G1(d, 2, j) + (V(d, i) - V(d, j)) * h2;
        }
        for (int s = 1; s <= 3; ++s) {
          const double g1 = dwdx * (S(s, i) / (rho[i - 1] * rho[i - 1]) + S(s, j) / (rho[j - 1] * rho[j - 1]));
          G2(s, 1, i) = G2(s, 1, i) - mass[j - 1] * g1;
          const double g2 = dwdy * (S(s, i) / (rho[i - 1] * rho[i - 1]) + S(s, j) / (rho[j - 1] * rho[j - 1]));
          G2(s, 2, i) = G2(s, 2, i) - mass[j - 1] * g2;
        }
        if (p.cspm) {
          const double h1b = -dwdx * mass[j - 1] / rho[j - 1];
          const double h2b = -dwdy * mass[j - 1] / rho[j - 1];
          A(1, j) = A(1, j) + (X(1, i) - X(1, j)) * h1;
          A(2, j) = A(2, j) + (X(2, i) - X(2, j)) * h1;
          A(3, j) = A(3, j) + (X(1, i) - X(1, j)) * h2;
          A(4, j) = A(4, j) + (X(2, i) - X(2, j)) * h2;
          A(1, i) = A(1, i) + (X(1, j) - X(1, i)) * h1b;
          A(2, i) = A(2, i) + (X(2, j) - X(2, i)) * h1b;
          A(3, i) = A(3, i) + (X(1, j) - X(1, i)) * h2b;
          A(4, i) = A(4, i) + (X(2, j) - X(2, i)) * h2b;
        }
      } else if (c.pint_type == 9) {  // i stress particle, j dummy
        const double beta_max = 1.5, vel_wall = 0.0;
        const double wall = (double)wall_position[j - 1];
        double da, db;
        if (horizontal_or_not[j - 1] == 1) {
          da = std::fabs(X(2, i) - wall);
          db = std::fabs(X(2, j) - wall);
        } else {
          da = std::fabs(X(1, i) - wall);
          db = std::fabs(X(1, j) - wall);
        }
        const double bq = 1 + (db / da);
        const double beta = (bq < beta_max) ? bq : beta_max;  // MIN(beta_max, 1+db/da); NaN -> beta_max
        const double dv[2] = {V(1, i) * (1 - beta) + beta * vel_wall, V(2, i) * (1 - beta) + beta * vel_wall};
        const double h1 = dwdx * mass[j - 1] / rho[j - 1];
        const double h2 = dwdy * mass[j - 1] / rho[j - 1];
        for (int d = 1; d <= 2; ++d) {
          G1(d, 1, i) = G1(d, 1, i) + (V(d, i) - dv[d - 1]) * h1;
          G1(d, 2, i) = G1(d, 2, i) + (V(d, i) - dv[d - 1]) * h2;
        }
      } else if (c.pint_type == 6) {  // i node, j dummy
        bc_int[i - 1] = 1;
        for (int s = 1; s <= 4; ++s) S(s, j) = S(s, i);
        for (int s = 1; s <= 3; ++s) {
          const double h1 = dwdx * (S(s, i) / (rho[i - 1] * rho[i - 1]) + S(s, j) / (rho[j - 1] * rho[j - 1]));
          G2(s, 1, i) = G2(s, 1, i) - mass[j - 1] * h1;
          const double h2 = dwdy * (S(s, i) / (rho[i - 1] * rho[i - 1]) + S(s, j) / (rho[j - 1] * rho[j - 1]));
          G2(s, 2, i) = G2(s, 2, i) - mass[j - 1] * h2;
        }
      }
    }
    if (p.cspm) {
      const double thr = (double)p.ae_threshold;
      for (int i = 1; i <= ntotal; ++i) {
        A(5, i) = A(1, i) * A(4, i) - A(2, i) * A(3, i);
        if (std::fabs(A(5, i)) < thr) {
          A(5, i) = 1;
          A(1, i) = 1;
          A(2, i) = 0;
          A(3, i) = 0;
          A(4, i) = 1;
        } else {
          A(5, i) = 1 / A(5, i);
        }
      }
      // main:617-628: array statements, so the second of each pair sees the already-updated first column;
      // only components 1..ndimn of grad2_tmp are corrected.
      for (int a = 1; a <= 2; ++a) {
        for (int i = nnode + 1; i <= ntotal; ++i) G1(a, 1, i) = A(5, i) * (A(1, i) * G1(a, 1, i) + A(2, i) * G1(a, 2, i));
        for (int i = nnode + 1; i <= ntotal; ++i) G1(a, 2, i) = A(5, i) * (A(3, i) * G1(a, 1, i) + A(4, i) * G1(a, 2, i));
        for (int i = 1; i <= nnode; ++i) G2(a, 1, i) = A(5, i) * (A(1, i) * G2(a, 1, i) + A(2, i) * G2(a, 2, i));
        for (int i = 1; i <= nnode; ++i) G2(a, 2, i) = A(5, i) * (A(3, i) * G2(a, 1, i) + A(4, i) * G2(a, 2, i));
      }
    }
    for (int i = nnode + 1; i <= ntotal; ++i) {
      double *d1 = &divf1[4 * (size_t)(i - 1)];
      d1[0] = -(p.D11 * G1(1, 1, i) + p.D12 * G1(2, 2, i));
      d1[1] = -(p.D12 * G1(1, 1, i) + p.D22 * G1(2, 2, i));
      d1[2] = -(p.D33 * G1(2, 1, i) + p.D33 * G1(1, 2, i));
      d1[3] = -(p.D41 * G1(1, 1, i) + p.D42 * G1(2, 2, i));
    }
    for (int i = 1; i <= nnode; ++i) {
      divf2[2 * (size_t)(i - 1)] = -(G2(1, 1, i) + G2(3, 2, i));
      divf2[2 * (size_t)(i - 1) + 1] = -(G2(3, 1, i) + G2(2, 2, i));
    }
    grad_u = grad1;
  }

  // ---- drucker_prager, mat:1958-2083 (fp32 locals) ---------------------------------------------
  void drucker_prager(int ie, const double stress2[4], double vivel[4], double G[4]) {
    const double f0 = f_drucker[ie - 1];
    const float tanfi = (float)p.props[12], coh = (float)p.props[13];
    const float young = (float)p.props[2], poiss = (float)p.props[3];
    const float smean = (float)((stress2[0] + stress2[1] + stress2[3]) / 3.0);
    float devia[5];
    devia[1] = (float)(stress2[0] - (double)smean);
    devia[2] = (float)(stress2[1] - (double)smean);
    devia[3] = (float)stress2[2];
    devia[4] = (float)(stress2[3] - (double)smean);
    const float varj2 = devia[3] * devia[3] + 0.5f * (devia[1] * devia[1] + devia[2] * devia[2] + devia[4] * devia[4]);
    const float vari1 = 3 * smean;
    const float eps11 = (float)GU(1, 1, ie);
    const float eps12 = (float)(0.5 * (GU(1, 2, ie) + GU(2, 1, ie)));
    const float eps22 = (float)GU(2, 2, ie);
    const float emean = eps11 + eps22;
    const float alpha2 = tanfi / (std::sqrt(9 + 12 * (tanfi * tanfi)));
    const float kc = (3 * coh) / (std::sqrt(9 + 12 * (tanfi * tanfi)));
    const float yield = -alpha2 * vari1 + kc;
    const float sq = std::sqrt(varj2);
    f_drucker[ie - 1] = (double)(sq - yield);
    const float f1 = (float)f_drucker[ie - 1];
    const double df = (double)f1 - f0;
    const float G_mod = young / (2.f * (1.f + poiss));
    const float K_mod = young / (3.f * (1.f - 2.f * poiss));
    const float s_eps = devia[1] * eps11 + 2 * devia[3] * eps12 + devia[2] * eps22;
    if (time_sph > 0 && f1 >= 0 && df >= 0 && sq >= 10e-06f) {
      float G1[5];
      for (int i = 1; i <= 4; ++i) G1[i] = (G_mod / sq) * devia[i];
      const float lambda_1 = 3 * alpha2 * K_mod * emean;
      const float lambda_2 = (G_mod / sq) * s_eps;
      const float lambda_3 = G_mod;
      const float G2 = (lambda_1 + lambda_2) / lambda_3;
      for (int i = 1; i <= 4; ++i) G[i - 1] = (double)(G1[i] * G2);
      vivel[0] = (double)((1.f / 6.f) * (1.f / sq) * (2 * devia[1] - devia[2] - devia[4]));
      vivel[1] = (double)((1.f / 6.f) * (1.f / sq) * (2 * devia[2] - devia[1] - devia[4]));
      vivel[2] = (double)((1.f / sq) * devia[3]);
      vivel[3] = (double)((1.f / 6.f) * (1 / sq) * (2 * devia[4] - devia[1] - devia[2]));
      for (int i = 0; i < 4; ++i) vivel[i] = vivel[i] * (double)G2;
    } else {
      for (int i = 0; i < 4; ++i) {
        G[i] = 0.0;
        vivel[i] = 0.0;
      }
    }
  }

  // ---- Get_Vivel -> invar09 / yieldf09 / flowvp09 (strain_localisation copy :2169-2540), fp64 -----
  // Yield criteria 1..4 are restated; ncrit = 5 (Cam Clay) is out of scope (no shipped input uses it).
  bool get_vivel(int ie, const double stress2[4], double vivel[4]) {
    const int ncrit = p.ncrit;
    const double root3 = (double)std::sqrt(3.00f);
    const double smean = (stress2[0] + stress2[1] + stress2[3]) / 3.0;
    double devia[5];
    devia[1] = stress2[0] - smean;
    devia[2] = stress2[1] - smean;
    devia[3] = stress2[2];
    devia[4] = stress2[3] - smean;
    const double varj2 = devia[3] * devia[3] + 0.5 * (devia[1] * devia[1] + devia[2] * devia[2] + devia[4] * devia[4]);
    const double varj3 = devia[4] * (devia[4] * devia[4] - varj2);
    const double steff = std::sqrt(varj2);
    double sint3;
    if (steff != 0.0) {
      sint3 = -3.0 * root3 * varj3 / (2.0 * varj2 * steff);
      if (sint3 > 1.0) sint3 = 1.0;
    } else {
      sint3 = 0.0;
    }
    if (sint3 < -1.0) sint3 = -1.0;
    if (sint3 > 1.0) sint3 = 1.0;
    const double theta = std::asin(sint3) / 3.0;
    double yield = 0;
    if (ncrit == 1) {
      yield = 2.0 * std::cos(theta) * steff;
    } else if (ncrit == 2) {
      yield = root3 * steff;
    } else if (ncrit == 3) {
      const double phira = p.props[8] * (double)0.017453292f;
      const double snphi = std::sin(phira);
      yield = smean * snphi + steff * (std::cos(theta) - std::sin(theta) * snphi / root3);
    } else if (ncrit == 4) {
      const double phira = p.props[8] * (double)0.017453292f;
      const double snphi = std::sin(phira);
      yield = 6.0 * smean * snphi / (root3 * (3.0 - snphi)) + steff;
    } else {
      err = "ncrit = 5 (Cam Clay) is not supported";
      return false;
    }
    const double evpstn = IV1(ie);
    const double fdatm0 = p.props[6], hards = p.props[7];
    double fdatm;
    if (fdatm0 > (double)0.001f) {
      fdatm = fdatm0 + hards * evpstn;
      double fact = std::fabs(fdatm) / std::fabs(fdatm0);
      if (fact < (double)0.1f) fact = (double)0.1f;
      fdatm = fdatm0 * fact;
    } else {
      err = "initial yield surface size too small (reference STOPs, mat:2322-2332)";
      return false;
    }
    for (int i = 0; i < 4; ++i) vivel[i] = 0.0;
    if (!(yield > fdatm)) return true;
    // yieldf09
    const double tanth = std::tan(theta), tant3 = std::tan(3.0 * theta), sinth = std::sin(theta), costh = std::cos(theta),
                 cost3 = std::cos(3.0 * theta);
    const double veca1[4] = {1.0, 1.0, 0.0, 1.0};
    double veca2[4] = {0, 0, 0, 0}, veca3[4];
    if (steff > 0) {
      for (int s = 0; s < 4; ++s) veca2[s] = devia[s + 1] / (2.0 * steff);
      veca2[2] = devia[3] / steff;
    }
    veca3[0] = devia[2] * devia[4] + varj2 / 3.0;
    veca3[1] = devia[1] * devia[4] + varj2 / 3.0;
    veca3[2] = -2.0 * devia[3] * devia[4];
    veca3[3] = devia[1] * devia[2] - devia[3] * devia[3] + varj2 / 3.0;
    double cons1 = 0, cons2 = 0, cons3 = 0;
    const double frict = p.props[8];
    if (ncrit == 1) {
      cons1 = 0.0;
      const double abthe = std::fabs(theta * 57.29577951308);
      if (abthe >= 29.0) {
        cons2 = root3;
        cons3 = 0.0;
      } else {
        cons2 = 2.0 * (costh + sinth * tant3);
        cons3 = root3 * sinth / (varj2 * cost3);
      }
    } else if (ncrit == 2) {
      cons1 = 0.0;
      cons2 = root3;
      cons3 = 0.0;
    } else if (ncrit == 3) {
      cons1 = std::sin(frict * (double)0.017453292f) / 3.0;
      const double abthe = std::fabs(theta * 57.29577951308);
      if (abthe >= 29.0) {
        cons3 = 0.0;
        double plumi = 1.0;
        if (theta > 0.0) plumi = -1.0;
        cons2 = 0.5 * (root3 + plumi * cons1 * root3);
      } else {
        cons2 = costh * ((1.0 + tanth * tant3) + cons1 * (tant3 - tanth) * root3);
        cons3 = (root3 * sinth + 3.0 * cons1 * costh) / (2.0 * varj2 * cost3);
      }
    } else if (ncrit == 4) {
      const double snphi = std::sin(frict * (double)0.017453292f);
      cons1 = 2.0 * snphi / (root3 * (3.0 - snphi));
      cons2 = 1.0;
      cons3 = 0.0;
    }
    double avect[4];
    for (int s = 0; s < 4; ++s) avect[s] = cons1 * veca1[s] + cons2 * veca2[s] + cons3 * veca3[s];
    // flowvp09
    const double allow = (double)0.01f;
    const double gamma = p.props[9], delta = p.props[10], nflow = p.props[11];
    const double fcurr = yield - fdatm;
    const double fnorm = fcurr / fdatm;
    if (fnorm >= allow) {
      double cmult;
      if (nflow != 1)
        cmult = gamma * (std::exp(delta * fnorm) - 1.0);
      else
        cmult = gamma * ((delta == 1.0) ? fnorm : std::pow(fnorm, delta));  // fnorm**1.0 == fnorm exactly
      for (int s = 0; s < 4; ++s) vivel[s] = cmult * avect[s];
    }
    return true;
  }

  // ---- plastic_terms, mat:1884-1954 ------------------------------------------------------------
  bool plastic_terms(int ie, const double fi[4], double Gs[4], double &der_intvars1) {
    for (int s = 0; s < 4; ++s) Gs[s] = 0.0;
    der_intvars1 = 0.0;
    if (p.ntype_eco <= 1) return true;
    double stress2[4] = {fi[0], fi[1], fi[2], fi[3]}, vivel[4] = {0, 0, 0, 0};
    if (p.ntype_solid == 1) stress2[3] = p.props[3] * (stress2[0] + stress2[1]);  // plane stress (Bui/VS copies only)
    if (p.ncrit <= 5) {
      if (!get_vivel(ie, stress2, vivel)) return false;
      // Get_Dmatx, plane strain (SL copy :2545-2576; ntype_solid == 2 branch of the Bui/VS copies)
      const double young = p.props[2], poiss = p.props[3];
      double D[4][4] = {{0}};
      const double cst = young * (1.0 - poiss) / ((1.0 + poiss) * (1.0 - 2.0 * poiss));
      D[0][0] = cst;
      D[1][1] = cst;
      D[0][1] = cst * poiss / (1.0 - poiss);
      D[1][0] = cst * poiss / (1.0 - poiss);
      D[2][2] = (1.0 - 2.0 * poiss) * cst / (2.0 * (1.0 - poiss));
      D[0][3] = cst * poiss / (1.0 - poiss);
      D[1][3] = cst * poiss / (1.0 - poiss);
      D[3][0] = cst * poiss / (1.0 - poiss);
      D[3][1] = cst * poiss / (1.0 - poiss);
      D[3][3] = cst;
      for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) Gs[a] = Gs[a] - D[a][b] * vivel[b];
    } else if (p.ncrit == 12) {
      double G2[4];
      drucker_prager(ie, stress2, vivel, G2);
      for (int s = 0; s < 4; ++s) Gs[s] = -G2[s];
    }
    // Get_derivative_intvars, mat:2697-2760
    if (p.ntype_solid == 0) {
      der_intvars1 = vivel[0];
    } else {
      der_intvars1 =
          std::sqrt((2.0 * (vivel[0] * vivel[0] + vivel[1] * vivel[1] + vivel[3] * vivel[3]) + vivel[2] * vivel[2]) / 3.0);
    }
    return true;
  }

  // ---- artificial_viscosity, main:826-904 (fp32 locals and accumulators) ---------------------------
  void artificial_viscosity() {
    std::fill(n_int.begin(), n_int.end(), 0.f);
    std::vector<float> av(2 * (size_t)nnode, 0.f);
    float visc = 0.f;  // keeps its previous value if div_u is NaN (neither branch of main:881-885 fires)
    const int64_t n = (int64_t)created.size();
    for (int64_t t = 0; t < n; ++t) {
      const Pair &c = created[creation_index(t)];
      if (c.pint_type != 3) continue;
      const int i = c.pair_i, j = c.pair_j;
      const float xij = (float)(X(1, i) - X(1, j));
      const float yij = (float)(X(2, i) - X(2, j));
      const float h = (float)(0.5 * (hsml[i - 1] + hsml[j - 1]));
      const float rho2 = (float)(0.5 * (rho[i - 1] + rho[j - 1]));
      const float cs = 600;
      n_int[i - 1] = n_int[i - 1] + 1;
      n_int[j - 1] = n_int[j - 1] + 1;
      float div_u = (float)((double)xij * (V(1, i) - V(1, j)));
      div_u = (float)((double)div_u + (double)yij * (V(2, i) - V(2, j)));
      const float sq = std::sqrt(xij * xij + yij * yij);
      const float theta = (h * div_u) / (sq * sq + 0.01f * (h * h));
      if (div_u < 0)
        visc = (float)((-p.alpha * (double)cs * (double)theta + p.beta * (double)(theta * theta)) / (double)rho2);
      else if (div_u >= 0)
        visc = 0;
      av[2 * (size_t)(i - 1)] = (float)((double)av[2 * (size_t)(i - 1)] + (double)(visc * c.dwdx) * mass[j - 1]);
      av[2 * (size_t)(j - 1)] = (float)((double)av[2 * (size_t)(j - 1)] - (double)(visc * c.dwdx) * mass[i - 1]);
      av[2 * (size_t)(i - 1) + 1] = (float)((double)av[2 * (size_t)(i - 1) + 1] + (double)(visc * c.dwdy) * mass[j - 1]);
      av[2 * (size_t)(j - 1) + 1] = (float)((double)av[2 * (size_t)(j - 1) + 1] - (double)(visc * c.dwdy) * mass[i - 1]);
    }
    for (size_t k = 0; k < av.size(); ++k) art_visc[k] = (double)(-av[k]);
  }

  // ---- boundary_forces, main:1039-1165 (branch test == 2; default REAL locals) ---------------------
  // Repulsion of the velocity particles by the innermost layer of wall particles, brute force over
  // (wall particle j outer, node i inner); every local is fp32, f_bound accumulates in fp64.
  void boundary_forces() {
    std::fill(f_bound.begin(), f_bound.end(), 0.0);
    const float c = 20;
    const int start = ntotal + 1, finish = ntotal + p.ndummy2;
    for (int j = start; j <= finish; ++j)
      for (int i = 1; i <= nnode; ++i) {
        float r[2];
        r[0] = (float)(X(1, i) - X(1, j));
        r[1] = (float)(X(2, i) - X(2, j));
        const float r2 = std::sqrt(r[0] * r[0] + r[1] * r[1]);
        const float d0 = (float)(p.dx / 2.);
        float f2;
        const float h = (float)(0.5 * (hsml[i - 1] + hsml[j - 1]));
        const float r_bound = r2 / (0.75f * h);
        if (r2 > 0 && r2 < 1.5f * d0)
          f2 = 1 - (r2 / (1.5f * d0));
        else
          f2 = 0.0f;
        float f;
        if (0 < r_bound && r_bound <= 2.f / 3.f)
          f = 2.f / 3.f;
        else if (2.f / 3.f < r_bound && r_bound <= 1)
          f = 2 * r_bound - 1.5f * (r_bound * r_bound);
        else if (1 < r_bound && r_bound < 2)
          f = 0.5f * ((2 - r_bound) * (2 - r_bound));
        else
          f = 0.f;
        const float pre = (0.01f * (c * c)) * f2 * f;
        for (int d = 0; d < 2; ++d)
          f_bound[2 * (size_t)(i - 1) + d] = f_bound[2 * (size_t)(i - 1) + d] + (double)(pre * (r[d] / (r2 * r2)));
      }
  }

  // ---- RK4, main:653-802 -----------------------------------------------------------------------
  bool rk4() {
    const double f1rk[4] = {0., 0.5, 0.5, 1.0}, f2rk[4] = {1., 2., 2., 1.0};
    const size_t nt = (size_t)ntotal;
    std::vector<double> vel0(vel.begin(), vel.begin() + 2 * nt), stress0(stress.begin(), stress.begin() + 4 * nt);
    std::vector<double> RK_stress(4 * nt, 0.0), RK_vel(2 * nt, 0.0), RK_dev_strain(nt, 0.0);
    std::vector<double> RHS_1(4 * nt, 0.0), RHS_2(2 * nt, 0.0), source_stress(4 * nt, 0.0), spin(4 * nt, 0.0),
        omega(2 * nt, 0.0);
    std::vector<double> source_grav(2 * (size_t)nnode, 0.0), art_force(2 * (size_t)nnode, 0.0);
    // continuity density (main:686-689): rho0, hsml0 of the stress particles; RK_rho, RK_hsml, RHS_rho, RHS_hsml(ntotal)
    const size_t ns = nt - (size_t)nnode;
    std::vector<double> rho0(rho.begin() + nnode, rho.begin() + ntotal), hsml0(hsml.begin() + nnode, hsml.begin() + ntotal);
    std::vector<double> RK_rho(nt, 0.0), RK_hsml(nt, 0.0), RHS_rho(nt, 0.0), RHS_hsml(nt, 0.0);
    std::fill(art_visc.begin(), art_visc.end(), 0.0);
    std::fill(vel.begin(), vel.end(), 0.0);        // main:690: all entries incl. dummies
    std::fill(stress.begin(), stress.end(), 0.0);  //
    const int ncrit = p.ncrit;
    const double dt = dt_sph;
    for (int st = 0; st < 4; ++st) {
      for (int i = 1; i <= nnode; ++i)
        for (int d = 1; d <= 2; ++d)
          V(d, i) = vel0[2 * (size_t)(i - 1) + d - 1] + f1rk[st] * (dt)*RHS_2[2 * (size_t)(i - 1) + d - 1];
      for (int i = nnode + 1; i <= ntotal; ++i)
        for (int s = 1; s <= 4; ++s)
          S(s, i) = stress0[4 * (size_t)(i - 1) + s - 1] + f1rk[st] * (dt)*RHS_1[4 * (size_t)(i - 1) + s - 1];
      if (p.cont_density) {  // main:706-713; density_update main:807-821 reads the grad_u of the PREVIOUS sweep
        for (size_t k = 0; k < ns; ++k) {
          const int i = nnode + 1 + (int)k;
          RHS_rho[i - 1] = -rho[i - 1] * (GU(1, 1, i) + GU(2, 2, i));
        }
        for (size_t k = 0; k < ns; ++k) rho[nnode + k] = rho0[k] + f1rk[st] * (dt)*RHS_rho[nnode + k];
        if (p.sle == 2) {
          for (size_t k = 0; k < ns; ++k) RHS_hsml[nnode + k] = -(hsml0[k] / (rho[nnode + k] * 2)) * RHS_rho[nnode + k];
          for (size_t k = 0; k < ns; ++k) hsml[nnode + k] = hsml0[k] + f1rk[st] * (dt)*RHS_hsml[nnode + k];
        }
      }
      if (ncrit == 12) adapt_stress2();
      if (p.no_bcs > 0) bcs();
      stress_point_update();
      if (p.sp_sph) {
        if (ncrit == 12) adapt_stress2();
        if (p.no_bcs > 0) bcs();
      }
      get_derivatives();
      for (int j = nnode + 1; j <= ntotal; ++j) {
        double G_local[4], der1;
        if (!plastic_terms(j, &stress[4 * (size_t)(j - 1)], G_local, der1)) return false;
        for (int s = 0; s < 4; ++s) source_stress[4 * (size_t)(j - 1) + s] = G_local[s];
        RK_dev_strain[j - 1] = RK_dev_strain[j - 1] + der1 * f2rk[st];
      }
      // gravity_force, mat:2809-2871
      {
        double g[2] = {0.0, 0.0};
        if (p.ic_grav != 0) {
          double factg = tcurve(p.tcurve_grav, time_sph);
          factg = factg * p.ft_grav;
          if (p.ic_grav == 1) {
            g[0] = factg * p.cgrav[0];
            g[1] = factg * p.cgrav[1];
          }
        }
        for (int i = 1; i <= nnode; ++i)
          for (int d = 1; d <= 2; ++d) source_grav[2 * (size_t)(i - 1) + d - 1] = g[d - 1] - p.damping * V(d, i);
      }
      if (p.inside_approach) boundary_forces();  // Bui copy main:742; elsewhere .and. dummy_nodes (ndummy2 = 0 then)
      if (p.alpha > 0 || p.beta > 0) artificial_viscosity();
      if (p.art_stress) artificial_force(art_force);  // main:746
      if (p.update_x) {  // main:751-757, get_spin_rate_tensor main:1021-1034
        for (int i = 1; i <= ntotal; ++i) {
          omega[2 * (size_t)(i - 1)] = 0.5 * (GU(1, 2, i) - GU(2, 1, i));
          omega[2 * (size_t)(i - 1) + 1] = -0.5 * (GU(1, 2, i) - GU(2, 1, i));
        }
        for (int i = nnode + 1; i <= ntotal; ++i) {
          const double o1 = omega[2 * (size_t)(i - 1)], o2 = omega[2 * (size_t)(i - 1) + 1];
          spin[4 * (size_t)(i - 1)] = 2 * o1 * S(3, i);
          spin[4 * (size_t)(i - 1) + 1] = 2 * o2 * S(3, i);
          spin[4 * (size_t)(i - 1) + 2] = o2 * S(1, i) + o1 * S(2, i);
        }
      }
      for (int i = nnode + 1; i <= ntotal; ++i)
        for (int s = 0; s < 4; ++s) {
          const size_t k = 4 * (size_t)(i - 1) + s;
          RHS_1[k] = -divf1[k] + spin[k] + source_stress[k];
        }
      for (int i = 1; i <= nnode; ++i)
        for (int d = 0; d < 2; ++d) {
          const size_t k = 2 * (size_t)(i - 1) + d;
          // f_bound is zero unless boundary_forces runs (App. C-2); art_force is zero unless art_stress = T
          RHS_2[k] = -divf2[k] + source_grav[k] + art_visc[k] + f_bound[k] + art_force[k];
        }
      for (int i = 1; i <= nnode; ++i)
        for (int d = 0; d < 2; ++d) {
          const size_t k = 2 * (size_t)(i - 1) + d;
          RK_vel[k] = RK_vel[k] + f2rk[st] * RHS_2[k];
        }
      for (int i = nnode + 1; i <= ntotal; ++i)
        for (int s = 0; s < 4; ++s) {
          const size_t k = 4 * (size_t)(i - 1) + s;
          RK_stress[k] = RK_stress[k] + f2rk[st] * RHS_1[k];
        }
      if (p.cont_density) {  // main:772-777
        for (size_t k = 0; k < nt; ++k) RK_rho[k] = RK_rho[k] + f2rk[st] * RHS_rho[k];
        if (p.sle == 2)
          for (size_t k = 0; k < nt; ++k) RK_hsml[k] = RK_hsml[k] + f2rk[st] * RHS_hsml[k];
      }
    }
    for (int i = 1; i <= nnode; ++i)
      for (int d = 1; d <= 2; ++d) V(d, i) = vel0[2 * (size_t)(i - 1) + d - 1] + (dt / 6) * RK_vel[2 * (size_t)(i - 1) + d - 1];
    for (int i = nnode + 1; i <= ntotal; ++i)
      for (int s = 1; s <= 4; ++s)
        S(s, i) = stress0[4 * (size_t)(i - 1) + s - 1] + (dt / 6) * RK_stress[4 * (size_t)(i - 1) + s - 1];
    if (ncrit == 12) adapt_stress2();
    if (p.no_bcs > 0) bcs();
    if (p.cont_density) {  // main:792-797
      for (size_t k = 0; k < ns; ++k) rho[nnode + k] = rho0[k] + (dt / 6.) * RK_rho[nnode + k];
      if (p.sle == 2)
        for (size_t k = 0; k < ns; ++k) hsml[nnode + k] = hsml0[k] + (dt / 6.) * RK_hsml[nnode + k];
    }
    for (size_t i = 0; i < nt; ++i) Ddev_strn[i] = RK_dev_strain[i] / 6;
    return true;
  }

  // ---- artificial_force, main:908-1016 (Monaghan 2000 artificial stress on node-node pairs; all DP) ----------
  void artificial_force(std::vector<double> &art_force) {
    const size_t nn = (size_t)nnode;
    std::vector<double> sigma2(2 * nn, 0.0), R2(2 * nn, 0.0), R(3 * nn, 0.0), tmp(6 * nn, 0.0);  // tmp(istre,d,i)
    const double eps = (double)0.1f, nexp = (double)2.55f;  // default-real literals assigned to DP variables
    double w2 = 0.0, gradw2[2] = {0.0, 0.0};
    const double hsml0 = (double)1.2f * p.dx;
    const double dx2[2] = {p.dx, p.dy};
    kernel(p.dx, dx2, hsml0, w2, gradw2);
    const int64_t n = (int64_t)created.size();
    for (int64_t t = 0; t < n; ++t) {
      const Pair &c = created[creation_index(t)];
      if (c.pint_type != 3) continue;
      const int i = c.pair_i, j = c.pair_j;
      const size_t a = (size_t)(i - 1), b = (size_t)(j - 1);
      const double s12i = S(1, i) - S(2, i), s12j = S(1, j) - S(2, j);
      const double theta = (s12i >= (double)1e-08f) ? 0.5 * std::atan(2 * S(3, i) / s12i) : 0.0;
      const double theta2 = (s12j >= (double)1e-08f) ? 0.5 * std::atan(2 * S(3, j) / s12j) : 0.0;
      const double ci = std::cos(theta), si = std::sin(theta), cj = std::cos(theta2), sj = std::sin(theta2);
      sigma2[2 * a] = (ci * ci) * S(1, i) + 2 * ci * si * S(3, i) + (si * si) * S(2, i);
      sigma2[2 * b] = (cj * cj) * S(1, j) + 2 * cj * sj * S(3, j) + (sj * sj) * S(2, j);
      sigma2[2 * a + 1] = (si * si) * S(1, i) - 2 * ci * si * S(3, i) + (ci * ci) * S(2, i);
      sigma2[2 * b + 1] = (sj * sj) * S(1, j) - 2 * cj * sj * S(3, j) + (cj * cj) * S(2, j);
      const double f = ((double)c.w) / w2;
      for (int k = 0; k < 2; ++k) {
        R2[2 * a + k] = sigma2[2 * a + k] > 0 ? -eps * (sigma2[2 * a + k] / (rho[a] * rho[a])) : 0.0;
        R2[2 * b + k] = sigma2[2 * b + k] > 0 ? -eps * (sigma2[2 * b + k] / (rho[b] * rho[b])) : 0.0;
      }
      R[3 * a] = R2[2 * a] * (ci * ci) + R2[2 * a + 1] * (si * si);
      R[3 * a + 1] = R2[2 * a] * ((ci * ci) + (si * si));
      R[3 * a + 2] = (R2[2 * a] - R2[2 * a + 1]) * (ci * si);
      R[3 * b] = R2[2 * b] * (cj * cj) + R2[2 * b + 1] * (sj * sj);
      R[3 * b + 1] = R2[2 * b] * ((cj * cj) + (sj * sj));
      R[3 * b + 2] = (R2[2 * b] - R2[2 * b + 1]) * (cj * sj);
      const double fn = std::pow(f, nexp);
      for (int is = 0; is < 3; ++is) {
        const double h1 = (double)c.dwdx * fn * (R[3 * a + is] + R[3 * b + is]);
        tmp[6 * a + is] = tmp[6 * a + is] + mass[b] * h1;
        tmp[6 * b + is] = tmp[6 * b + is] - mass[a] * h1;
        const double h2 = (double)c.dwdy * fn * (R[3 * a + is] + R[3 * b + is]);
        tmp[6 * a + 3 + is] = tmp[6 * a + 3 + is] + mass[b] * h2;
        tmp[6 * b + 3 + is] = tmp[6 * b + 3 + is] - mass[a] * h2;
      }
    }
    for (size_t a = 0; a < nn; ++a) {
      art_force[2 * a] = tmp[6 * a + 0] + tmp[6 * a + 3 + 2];      // (1,1) + (3,2)
      art_force[2 * a + 1] = tmp[6 * a + 2] + tmp[6 * a + 3 + 1];  // (3,1) + (2,2)
    }
  }

  // ---- get_nodes_on_free_surface, mat:1116-1411 (every local is default REAL = fp32; subset, normal, x, mass,
  // rho, hsml are DP). Uses this step's pair list with the already updated positions (App. C-7). Rewrites
  // bc_or_not: 1 stays, free-surface particles become 2, all others 0.
  void get_nodes_on_free_surface() {
    const size_t nt = (size_t)ntotal;
    std::vector<float> A(4 * nt, 0.f), ff(2 * nt, 0.f), tt(2 * nt, 0.f), tau(2 * nt, 0.f), f_int(nt, 0.f),
        grad_f(2 * nt, 0.f), neighbour(nt, 0.f);
    subset.assign(nt, 0.0);
    normal.assign(2 * nt, 0.0);
    const int64_t n = (int64_t)created.size();
    auto add = [](float acc, double term) { return (float)((double)acc + term); };  // real = real + dp
    // Step 1: renormalisation matrix and sum of the kernel gradients (pair types 1, 2, 3)
    for (int64_t t = 0; t < n; ++t) {
      const Pair &c = created[creation_index(t)];
      if (c.pint_type < 1 || c.pint_type > 3) continue;
      const int i = c.pair_i, j = c.pair_j;
      const size_t a = (size_t)(i - 1), b = (size_t)(j - 1);
      const float hf1i = (float)(mass[b] * (double)c.dwdx / rho[b]);
      const float hf1j = (float)(-mass[a] * (double)c.dwdx / rho[a]);
      A[4 * a + 0] = add(A[4 * a + 0], (X(1, j) - X(1, i)) * (double)hf1i);
      A[4 * b + 0] = add(A[4 * b + 0], (X(1, i) - X(1, j)) * (double)hf1j);
      ff[2 * a + 0] = ff[2 * a + 0] + hf1i;
      ff[2 * b + 0] = ff[2 * b + 0] + hf1j;
      const float hf2i = (float)(mass[b] * (double)c.dwdy / rho[b]);
      const float hf2j = (float)(-mass[a] * (double)c.dwdy / rho[a]);
      A[4 * a + 1] = add(A[4 * a + 1], (X(2, j) - X(2, i)) * (double)hf1i);
      A[4 * a + 2] = add(A[4 * a + 2], (X(1, j) - X(1, i)) * (double)hf2i);
      A[4 * a + 3] = add(A[4 * a + 3], (X(2, j) - X(2, i)) * (double)hf2i);
      A[4 * b + 1] = add(A[4 * b + 1], (X(2, i) - X(2, j)) * (double)hf1j);
      A[4 * b + 2] = add(A[4 * b + 2], (X(1, i) - X(1, j)) * (double)hf2j);
      A[4 * b + 3] = add(A[4 * b + 3], (X(2, i) - X(2, j)) * (double)hf2j);
      ff[2 * a + 1] = ff[2 * a + 1] + hf2i;
      ff[2 * b + 1] = ff[2 * b + 1] + hf2j;
    }
    // Step 2: first approximation of the normal; Step 3a: scan point and tangent
    for (int i = 1; i <= ntotal; ++i) {
      const size_t a = (size_t)(i - 1);
      for (int k = 0; k < 4; ++k)
        if (std::fabs(A[4 * a + k]) <= 1.e-8f) A[4 * a + k] = 0.f;
      const float v1 = -(A[4 * a + 0] * ff[2 * a] + A[4 * a + 1] * ff[2 * a + 1]);
      const float v2 = -(A[4 * a + 2] * ff[2 * a] + A[4 * a + 3] * ff[2 * a + 1]);
      const float v3 = powf(v1 * v1 + v2 * v2, 0.5f);
      normal[2 * a] = (double)(v1 / v3);
      normal[2 * a + 1] = (double)(v2 / v3);
    }
    for (int i = 1; i <= ntotal; ++i) {
      const size_t a = (size_t)(i - 1);
      tt[2 * a] = (float)(X(1, i) + hsml[a] * normal[2 * a]);
      tt[2 * a + 1] = (float)(X(2, i) + hsml[a] * normal[2 * a + 1]);
      tau[2 * a] = (float)(-normal[2 * a + 1]);
      tau[2 * a + 1] = (float)(normal[2 * a]);
    }
    // Step 3b: is there a particle inside the scan region? `me` looks at `other`
    const float sqrt2 = 1.41421354f;  // 2**0.5 folded by the compiler to the fp32 constant
    auto scan = [&](int me, int other) {
      const size_t a = (size_t)(me - 1);
      if (subset[a] != 0) return;
      const double dx = X(1, other) - X(1, me), dy = X(2, other) - X(2, me);
      const float dist = (float)std::pow(dx * dx + dy * dy, (double)0.5f);
      const float xt1 = (float)(X(1, other) - (double)tt[2 * a]);
      const float xt2 = (float)(X(2, other) - (double)tt[2 * a + 1]);
      const float xt_norm = powf(xt1 * xt1 + xt2 * xt2, 0.5f);
      const float prod_scal = (float)(std::fabs(normal[2 * a] * (double)xt1 + normal[2 * a + 1] * (double)xt2) +
                                      (double)std::fabs(tau[2 * a] * xt1 + tau[2 * a + 1] * xt2));
      const float limit = (float)((double)sqrt2 * hsml[a]);
      if (dist >= limit && (double)xt_norm < hsml[a]) {
        subset[a] = 2;
        f_int[a] = -1;
      } else if (dist < limit && (double)prod_scal < hsml[a]) {
        subset[a] = 2;
        f_int[a] = -1;
      }
    };
    for (int64_t t = 0; t < n; ++t) {
      const Pair &c = created[creation_index(t)];
      if (c.pint_type == 2 || c.pint_type == 3) {
        scan(c.pair_i, c.pair_j);
        scan(c.pair_j, c.pair_i);
      }
      if (c.pint_type == 6 || c.pint_type == 9) scan(c.pair_j, c.pair_i);  // pair_i is the wall particle
    }
    for (size_t a = 0; a < nt; ++a)
      if (subset[a] == 0) subset[a] = 1;
    for (size_t a = 0; a < nt; ++a)
      if (bc_or_not[a] != 1) bc_or_not[a] = 0;
    for (size_t a = 0; a < nt; ++a) {
      if (subset[a] == 2 && bc_or_not[a] != 1)
        bc_or_not[a] = 0;
      else if (subset[a] == 1 && bc_or_not[a] != 1)
        bc_or_not[a] = 2;
      else if (subset[a] == 1 && bc_or_not[a] == 1)
        subset[a] = 2;
    }
    // Step 4: better normal from the neighbouring surface particles, oriented by the gradient of f_int
    std::fill(normal.begin(), normal.end(), 0.0);
    for (int64_t t = 0; t < n; ++t) {
      const Pair &c = created[creation_index(t)];
      if (c.pint_type < 1 || c.pint_type > 3) continue;
      const int i = c.pair_i, j = c.pair_j;
      const size_t a = (size_t)(i - 1), b = (size_t)(j - 1);
      if (subset[a] == 1 && subset[b] == 1) {
        const float x1 = (float)X(1, i), y1 = (float)X(2, i), x2 = (float)X(1, j), y2 = (float)X(2, j);
        const float x_vect = x2 - x1, y_vect = y2 - y1;
        normal[2 * a] = normal[2 * a] - (double)y_vect;
        normal[2 * a + 1] = normal[2 * a + 1] + (double)x_vect;
        normal[2 * b] = normal[2 * b] - (double)y_vect;
        normal[2 * b + 1] = normal[2 * b + 1] + (double)x_vect;
        neighbour[a] = neighbour[a] + 1;
        neighbour[b] = neighbour[b] + 1;
      }
      const float hf1i = (float)(mass[b] * (double)c.dwdx / rho[b]);
      const float hf1j = (float)(-mass[a] * (double)c.dwdx / rho[a]);
      const float hf2i = (float)(mass[b] * (double)c.dwdy / rho[b]);
      const float hf2j = (float)(-mass[a] * (double)c.dwdy / rho[a]);
      grad_f[2 * a] = grad_f[2 * a] + (f_int[b] - f_int[a]) * hf1i;
      grad_f[2 * b] = grad_f[2 * b] + (f_int[a] - f_int[b]) * hf1j;
      grad_f[2 * a + 1] = grad_f[2 * a + 1] + (f_int[b] - f_int[a]) * hf2i;
      grad_f[2 * b + 1] = grad_f[2 * b + 1] + (f_int[a] - f_int[b]) * hf2j;
    }
    for (size_t a = 0; a < nt; ++a) {
      if (subset[a] != 1) continue;
      normal[2 * a] = normal[2 * a] / (double)neighbour[a];
      normal[2 * a + 1] = normal[2 * a + 1] / (double)neighbour[a];
      const float norm_vect = (float)std::pow(normal[2 * a] * normal[2 * a] + normal[2 * a + 1] * normal[2 * a + 1], (double)0.5f);
      normal[2 * a] = normal[2 * a] / (double)norm_vect;
      normal[2 * a + 1] = normal[2 * a + 1] / (double)norm_vect;
      const float p_scal = (float)((double)grad_f[2 * a] * normal[2 * a] + (double)grad_f[2 * a + 1] * normal[2 * a + 1]);
      if (p_scal < 0) {
        normal[2 * a] = -normal[2 * a];
        normal[2 * a + 1] = -normal[2 * a + 1];
      }
    }
  }

  // ---- XSPH_update, main:189-239 ---------------------------------------------------------------
  void xsph_update() {
    std::vector<double> vel_sum(2 * (size_t)ntotal, 0.0);
    const double eps = 0.5;
    std::fill(n_int.begin(), n_int.end(), 0.f);
    const int64_t n = (int64_t)created.size();
    for (int64_t t = 0; t < n; ++t) {
      const Pair &c = created[creation_index(t)];
      const int i = c.pair_i, j = c.pair_j;
      const double w = (double)c.w;
      if (c.pint_type == 2) {
        for (int d = 1; d <= 2; ++d) {
          vel_sum[2 * (size_t)(i - 1) + d - 1] += (mass[j - 1] / rho[j - 1]) * (V(d, j) - V(d, i)) * w;
          vel_sum[2 * (size_t)(j - 1) + d - 1] += (mass[i - 1] / rho[i - 1]) * (V(d, i) - V(d, j)) * w;
        }
      } else if (c.pint_type == 3) {
        for (int d = 1; d <= 2; ++d) {
          vel_sum[2 * (size_t)(j - 1) + d - 1] += (mass[i - 1] / rho[i - 1]) * (V(d, i) - V(d, j)) * w;
          vel_sum[2 * (size_t)(i - 1) + d - 1] += (mass[j - 1] / rho[j - 1]) * (V(d, j) - V(d, i)) * w;
        }
        n_int[i - 1] = n_int[i - 1] + 1;
        n_int[j - 1] = n_int[j - 1] + 1;
        // main:224-230: neighbours of a free-surface node are marked 3 until get_nodes_on_free_surface rewrites
        // bc_or_not at the end of this step (a marked particle with bc_or_not == 1 thereby loses its BCs)
        if (bc_or_not[i - 1] == 2 && bc_or_not[j - 1] != 2) bc_or_not[j - 1] = 3;
        if (bc_or_not[j - 1] == 2 && bc_or_not[i - 1] != 2) bc_or_not[i - 1] = 3;
      }
    }
    for (int i = 1; i <= ntotal; ++i)
      for (int d = 1; d <= 2; ++d) X(d, i) = X(d, i) + dt_sph * (V(d, i) + eps * vel_sum[2 * (size_t)(i - 1) + d - 1]);
  }

  // ---- isolated_nodes, main:373-398 ------------------------------------------------------------
  void isolated_nodes() {
    std::fill(n_int.begin(), n_int.end(), 0.f);
    for (const Pair &c : created)
      if (c.pint_type == 3) {
        n_int[c.pair_i - 1] += 1;
        n_int[c.pair_j - 1] += 1;
      }
  }

  // ---- shift_stress_points, main:244-368 -------------------------------------------------------
  void shift_stress_points() {
    const double dx = p.dx;
    const int npoints = p.npoints;
    if (itimestep_sph % p.shift_update == 0) {
      if (p.vel_vector) {
        int k = nnode + 1;
        for (int i = 1; i <= nnode; ++i) {
          const double ddx = X(1, i) - x_10[2 * (size_t)(i - 1)], ddy = X(2, i) - x_10[2 * (size_t)(i - 1) + 1];
          double d10 = std::sqrt(ddx * ddx + ddy * ddy);
          d10 = d10 / dx;
          disp_10[i - 1] = d10;
          x_10[2 * (size_t)(i - 1)] = X(1, i);
          x_10[2 * (size_t)(i - 1) + 1] = X(2, i);
          const double vn = std::sqrt(V(1, i) * V(1, i) + V(2, i) * V(2, i));
          const float cos_theta = (float)(V(1, i) / vn), sin_theta = (float)(V(2, i) / vn);
          float r1 = (float)((dx / 2) * (double)cos_theta), r2 = (float)((dx / 2) * (double)sin_theta);
          if (r1 > 0 && (double)r1 < dx / 5) r1 = (float)((double)r1 + dx / 3.);
          if (r2 > 0 && (double)r2 < dx / 5) r2 = (float)((double)r2 + dx / 3.);
          if (r1 < 0 && (double)r1 > -dx / 5) r1 = (float)((double)r1 - dx / 3.);
          if (r2 < 0 && (double)r2 > -dx / 5) r2 = (float)((double)r2 - dx / 3.);
          if (d10 > p.disp_tol) {
            X(1, k) = X(1, i) + (double)r1;
            X(2, k) = X(2, i) + (double)r2;
            X(1, k + 1) = X(1, i) - (double)r1;
            X(2, k + 1) = X(2, i) - (double)r2;
          }
          k = k + 2;
        }
      } else {
        const double r_x = p.r_x, r_y = p.r_y;
        if (npoints == 1) {
          for (int i = 1; i <= nnode; ++i) {
            X(1, nnode + i) = X(1, i) + r_x;
            X(2, nnode + i) = X(2, i) + r_y;
          }
        } else if (npoints == 2) {
          int k = nnode + 1;
          for (int i = 1; i <= nnode; ++i) {
            X(1, k) = X(1, i) + r_x;
            X(2, k) = X(2, i) + r_y;
            X(1, k + 1) = X(1, i) - r_x;
            X(2, k + 1) = X(2, i) - r_x;  // sic: r_x (main:309)
            k = k + 2;
          }
        } else if (npoints == 3) {
          int k = nnode + 1;
          for (int i = 1; i <= nnode; ++i) {
            X(1, k) = X(1, i);
            X(2, k) = X(2, i) + r_y;
            X(1, k + 1) = X(1, i) - r_x;
            X(2, k + 1) = X(2, i) - r_y;
            X(1, k + 2) = X(1, i) + r_x;
            X(2, k + 2) = X(2, i) - r_y;
            k = k + 3;
          }
        }
      }
    }
    if ((p.alpha == 0 && p.beta == 0) && (!p.xsph)) isolated_nodes();
    int k = nnode + 1;
    for (int i = 1; i <= nnode; ++i) {
      const float abs_vel = (float)std::sqrt(V(1, i) * V(1, i) + V(2, i) * V(2, i));
      bool collapse = false;
      if (bc_int[i - 1] == 1 && abs_vel > 0.4f) collapse = true;
      if (n_int[i - 1] < 2) collapse = true;
      if (collapse)
        for (int q = 0; q < npoints && q < 3; ++q) {
          X(1, k + q) = X(1, i);
          X(2, k + q) = X(2, i);
        }
      k = k + npoints;
    }
  }

  // ---- time_integration, main:78-184 -----------------------------------------------------------
  bool step(int itimestep, double time, double dt) {
    itimestep_sph = itimestep;
    time_sph = time;
    dt_sph = dt;
    const int ncrit = p.ncrit;
    check_out_domain();
    grid_find();
    pint_update();
    if (p.sph_shift && itimestep_sph > 1 && ((itimestep_sph - 1) % p.shift_update == 0)) {
      stress_point_update();
      if (ncrit == 12) adapt_stress2();
    }
    vx0.assign(vel.begin(), vel.begin() + 2 * (size_t)ntotal);
    x0.assign(x.begin(), x.begin() + 2 * (size_t)ntotal);
    if (!rk4()) return false;
    for (int ip = nnode + 1; ip <= ntotal; ++ip) IV1(ip) = IV1(ip) + dt * Ddev_strn[ip - 1];  // update_strain
    stress_point_update();
    if (p.sp_sph) {
      if (ncrit == 12) adapt_stress2();
      if (p.no_bcs > 0) bcs();
    }
    if (p.update_x) {
      if (p.xsph) {
        xsph_update();
      } else {
        for (size_t k = 0; k < 2 * (size_t)ntotal; ++k) {
          const float vel_half = (float)(0.5 * (vx0[k] + vel[k]));  // real :: vel_half, main:89,145
          x[k] = x0[k] + (double)vel_half * dt_sph;
        }
      }
      x_at_fs = x;
      get_nodes_on_free_surface();  // main:152-154 (ndimn == 2 always here)
      if (p.sp_sph && !p.inside_approach) shift_stress_points();
      if (!p.sp_sph)
        for (int i = 1; i <= nnode; ++i) {
          X(1, nnode + i) = X(1, i);
          X(2, nnode + i) = X(2, i);
        }
      for (size_t k = 0; k < 2 * (size_t)nnode; ++k) displ[k] = x[k] - x00[k];
    } else {
      for (size_t k = 0; k < 2 * (size_t)nnode; ++k) displ[k] = displ[k] + 0.5 * (vx0[k] + vel[k]) * dt_sph;
    }
    return true;
  }
};

template <class T>
void copy_in(std::vector<T> &dst, const T *src, size_t n) {
  dst.assign(n, T());
  if (src) std::memcpy(dst.data(), src, n * sizeof(T));
}
template <class T>
void copy_out(T *dst, const std::vector<T> &src, size_t n) {
  if (dst) std::memcpy(dst, src.data(), n * sizeof(T));
}

}  // namespace

extern "C" {

void *oracle_create(const spsph_params *p, const spsph_state *s) {
  if (!p || p->struct_bytes != (int32_t)sizeof(spsph_params)) return nullptr;
  Oracle *o = new Oracle();
  o->p = *p;
  o->nnode = p->nnode;
  o->ntotal = p->ntotal;
  o->ntotal2 = p->ntotal2;
  const size_t n2 = (size_t)p->ntotal2, nt = (size_t)p->ntotal, nn = (size_t)p->nnode;
  copy_in(o->x, s->x, 2 * n2);
  copy_in(o->vel, s->vel, 2 * n2);
  copy_in(o->stress, s->stress, 4 * n2);
  copy_in(o->rho, s->rho, n2);
  copy_in(o->mass, s->mass, n2);
  copy_in(o->hsml, s->hsml, n2);
  copy_in(o->itype, s->itype, n2);
  copy_in(o->internal_vars, s->internal_vars, (size_t)SPSPH_NINT_VARS * nt);
  copy_in(o->f_drucker, s->f_drucker, nt);
  copy_in(o->x00, s->x00, 2 * n2);
  copy_in(o->displ, s->displ, 2 * nn);
  copy_in(o->x_10, s->x_10, 2 * nn);
  copy_in(o->disp_10, s->disp_10, nn);
  copy_in(o->wall_position, s->wall_position, n2);
  copy_in(o->horizontal_or_not, s->horizontal_or_not, n2);
  copy_in(o->n_int, s->n_int, nn);
  copy_in(o->bc_int, s->bc_int, nn);
  copy_in(o->if_out, s->if_out_domain, n2);
  copy_in(o->bc_or_not, s->bc_or_not, nt);
  copy_in(o->bc_info, s->bc_info, 8 * nt);
  o->grad_u.assign(4 * nt, 0.0);
  o->normal.assign(2 * nt, 0.0);  // module array, first written by get_nodes_on_free_surface
  o->art_visc.assign(2 * nn, 0.0);
  o->f_bound.assign(2 * nn, 0.0);
  o->Ddev_strn.assign(nt, 0.0);
  o->countiac.assign(n2, 0);
  return o;
}

int oracle_step(void *h, int32_t itimestep_sph, double time_sph, double dt_sph) {
  Oracle *o = (Oracle *)h;
  return o->step(itimestep_sph, time_sph, dt_sph) ? 0 : 1;
}

int oracle_run(void *h, int32_t first_itimestep, double time_sph, double dt_sph, int32_t nsteps, double *time_out) {
  Oracle *o = (Oracle *)h;
  for (int k = 0; k < nsteps; ++k) {
    if (!o->step(first_itimestep + k, time_sph, dt_sph)) return 1;
    time_sph = time_sph + dt_sph;  // 1_SPH_2018.f90:174
  }
  if (time_out) *time_out = time_sph;
  return 0;
}

int oracle_download(void *h, const spsph_state *s) {
  Oracle *o = (Oracle *)h;
  const size_t n2 = (size_t)o->ntotal2, nt = (size_t)o->ntotal, nn = (size_t)o->nnode;
  copy_out(s->x, o->x, 2 * n2);
  copy_out(s->vel, o->vel, 2 * n2);
  copy_out(s->stress, o->stress, 4 * n2);
  copy_out(s->rho, o->rho, n2);
  copy_out(s->mass, o->mass, n2);
  copy_out(s->hsml, o->hsml, n2);
  copy_out(s->itype, o->itype, n2);
  copy_out(s->internal_vars, o->internal_vars, (size_t)SPSPH_NINT_VARS * nt);
  copy_out(s->f_drucker, o->f_drucker, nt);
  copy_out(s->x00, o->x00, 2 * n2);
  copy_out(s->displ, o->displ, 2 * nn);
  copy_out(s->x_10, o->x_10, 2 * nn);
  copy_out(s->disp_10, o->disp_10, nn);
  copy_out(s->n_int, o->n_int, nn);
  copy_out(s->bc_int, o->bc_int, nn);
  copy_out(s->if_out_domain, o->if_out, n2);
  copy_out(s->bc_or_not, o->bc_or_not, nt);
  return 0;
}

int oracle_pair_stats(void *h, int64_t *npairs, int32_t *maxiac, int32_t *miniac, int32_t *noiac) {
  Oracle *o = (Oracle *)h;
  if (npairs) *npairs = (int64_t)o->created.size();
  if (maxiac) *maxiac = o->maxiac;
  if (miniac) *miniac = o->miniac;
  if (noiac) *noiac = o->noiac;
  return 0;
}

// pair list of the last step in traversal order, after Pint_Update
int oracle_pairs(void *h, int64_t *npairs, int32_t *pair_i, int32_t *pair_j, int32_t *pint_type, float *w, float *dwdx,
                 float *dwdy) {
  Oracle *o = (Oracle *)h;
  const int64_t n = (int64_t)o->created.size();
  if (npairs) *npairs = n;
  if (!pair_i) return 0;
  for (int64_t t = 0; t < n; ++t) {
    const Pair &c = o->created[o->creation_index(t)];
    pair_i[t] = c.pair_i;
    pair_j[t] = c.pair_j;
    pint_type[t] = c.pint_type;
    w[t] = c.w;
    dwdx[t] = c.dwdx;
    dwdy[t] = c.dwdy;
  }
  return 0;
}

// ---- hooks for tests/test_oracle_cpu.py::test_device_math_transcription: single routines on caller-given inputs ----
int oracle_plastic_terms(void *h, double time_sph, int32_t n, const double *stress, const double *grad, const double *epsp,
                         double *f_drucker, double *Gs, double *der1) {
  Oracle *o = (Oracle *)h;
  if (n > o->ntotal) return 1;
  o->time_sph = time_sph;
  for (int i = 1; i <= n; ++i) {
    o->GU(1, 1, i) = grad[4 * (i - 1)];
    o->GU(1, 2, i) = grad[4 * (i - 1) + 1];
    o->GU(2, 1, i) = grad[4 * (i - 1) + 2];
    o->GU(2, 2, i) = grad[4 * (i - 1) + 3];
    o->IV1(i) = epsp[i - 1];
    o->f_drucker[i - 1] = f_drucker[i - 1];
    if (!o->plastic_terms(i, stress + 4 * (i - 1), Gs + 4 * (i - 1), der1[i - 1])) return 1;
    f_drucker[i - 1] = o->f_drucker[i - 1];
  }
  return 0;
}

int oracle_adapt_stress(void *h, int32_t n, double *stress) {  // adapt_stress2 on the first n particles
  Oracle *o = (Oracle *)h;
  if (n > o->ntotal) return 1;
  const std::vector<double> keep = o->stress;
  std::copy(stress, stress + 4 * (size_t)n, o->stress.begin());
  o->adapt_stress2();
  std::copy(o->stress.begin(), o->stress.begin() + 4 * (size_t)n, stress);
  o->stress = keep;
  return 0;
}

int oracle_stress_free(void *h, int32_t n, double *stress, const double *normal) {  // apply_stress_free, first n nodes
  Oracle *o = (Oracle *)h;
  if (n > o->nnode) return 1;
  const std::vector<double> keep = o->stress;
  const std::vector<int32_t> kb = o->bc_or_not, ki = o->bc_int;
  o->normal.assign(2 * (size_t)o->ntotal, 0.0);
  std::copy(stress, stress + 4 * (size_t)n, o->stress.begin());
  std::copy(normal, normal + 2 * (size_t)n, o->normal.begin());
  std::fill(o->bc_or_not.begin(), o->bc_or_not.end(), 0);
  std::fill(o->bc_or_not.begin(), o->bc_or_not.begin() + n, 2);
  std::fill(o->bc_int.begin(), o->bc_int.end(), 0);
  o->apply_stress_free();
  std::copy(o->stress.begin(), o->stress.begin() + 4 * (size_t)n, stress);
  o->stress = keep;
  o->bc_or_not = kb;
  o->bc_int = ki;
  return 0;
}

void oracle_kernel(void *h, int32_t n, const double *r, const double *dx, const double *dy, const double *hh, double *w,
                   double *gx, double *gy) {
  Oracle *o = (Oracle *)h;
  for (int i = 0; i < n; ++i) {
    const double d[2] = {dx[i], dy[i]};
    double g[2];
    o->kernel(r[i], d, hh[i], w[i], g);
    gx[i] = g[0];
    gy[i] = g[1];
  }
}

// internals of the last step for tests/test_list_kernels_cpu.py (host emulation of the device's list kernels)
void oracle_debug_grid(void *h, int32_t *cell, int64_t *m_before, int64_t *npairs) {
  Oracle *o = (Oracle *)h;
  std::copy(o->last_cell.begin(), o->last_cell.end(), cell);
  *m_before = o->m_before;
  *npairs = (int64_t)o->created.size();
}
void oracle_debug_x_fs(void *h, double *x) {
  Oracle *o = (Oracle *)h;
  std::copy(o->x_at_fs.begin(), o->x_at_fs.end(), x);
}
void oracle_debug_surface(void *h, double *normal, double *subset) {
  Oracle *o = (Oracle *)h;
  std::copy(o->normal.begin(), o->normal.end(), normal);
  std::copy(o->subset.begin(), o->subset.end(), subset);
}

// list-capacity history (m_pairs), the one piece of a checkpoint that is not in spsph_state
int64_t oracle_get_list_capacity(void *h) { return ((Oracle *)h)->m_pairs; }
void oracle_set_list_capacity(void *h, int64_t m) { ((Oracle *)h)->m_pairs = m; }

const char *oracle_last_error(void *h) { return ((Oracle *)h)->err.c_str(); }

void oracle_destroy(void *h) { delete (Oracle *)h; }

}  // extern "C"
