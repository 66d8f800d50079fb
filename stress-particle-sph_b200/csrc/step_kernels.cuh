// Pair-sum and per-particle kernels of one time step (reference: stress_point_update main:403-482,
// get_derivatives main:487-648, RK4 main:653-802, artificial_viscosity main:826-904, plastic_terms
// mat:1884-1954, adapt_stress2 mat:2087-2161, Normal_BCs mat:1665-1770, gravity_force mat:2809-2871,
// XSPH_update main:189-239, shift_stress_points main:244-368, position update main:140-182).
//
// One thread per velocity / stress particle, walking its gather list front to back: every per-particle sum
// is accumulated in the reference's traversal order (fp32 sums bit-exact, fp64 sums too), no atomics.
//
// Buffering: a sweep A (interpolation) maps state buffer `src` to `dst` completely, the following sweep B
// (gradients + constitutive update + RK stage epilogue + next-stage predictor) maps `dst` back to `src`.
// Each kernel only reads neighbour values from the buffer it does not write, so there are no races.
#pragma once
#include "grid_kernels.cuh"

namespace spsph {

struct StatePtrs {
  // constant per step / persistent (original particle order)
  const double *x;         // (2, ntotal2)
  const double *mass, *rho, *hsml, *mor;  // mor = mass/rho
  const float *wallpos, *horiz;
  const int *bc_or_not, *bc_info;
  // ping-pong buffers
  double *V[2];    // (2, ntotal)
  double *S[2];    // (4, ntotal)
  double *sor;     // (3, ntotal) stress(1:3)/rho**2 of the buffer written by the last sweep A
  double *epsp;    // (ntotal) Internal_Vars(1,:)
  double *fdp;     // (ntotal) f_drucker
  double *norm;    // (ntotal) cspm_norm of stress_point_update (frozen within a step)
  double *AE;      // (5, ntotal) inverted CSPM matrix of get_derivatives (frozen within a step)
  double *vel0;    // (2, nnode)
  double *stress0; // (4, nstress)
  double *vx0;     // (2, ntotal)
  double *RKv;     // (2, nnode)
  double *RKs;     // (4, nstress)
  double *RKe;     // (nstress)
};

// ------------------------------------------------------------------------------------------------------
// RK4 prologue (main:681-690 + first predictor main:700-701 with f1rk = 0 + adapt_stress2/BCs main:715-716):
// saves vel0/stress0/vx0, zeroes the accumulators, and builds the stage-1 input buffer where stress-particle
// velocities, node stresses (and all dummy values) start from zero (main:690).
// ------------------------------------------------------------------------------------------------------
__global__ void k_rk_begin(DevParams P, StatePtrs st, int cur, int dst) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= P.ntotal) return;
  const double2 v = ld2(st.V[cur], id);
  const Stress4 s = ld4(st.S[cur], id);
  st2(st.vx0, id, v);
  double2 vn;
  Stress4 sn;
  if (id < P.nnode) {
    st2(st.vel0, id, v);
    st2(st.RKv, id, make_double2(0.0, 0.0));
    // RHS_2 = 0: vel0 + f1rk(1)*dt*RHS_2
    vn.x = v.x + 0. * (P.dt) * 0.0;
    vn.y = v.y + 0. * (P.dt) * 0.0;
    sn = Stress4{0.0, 0.0, 0.0, 0.0};
  } else {
    const int ks = id - P.nnode;
    st4(st.stress0, ks, s);
    st4(st.RKs, ks, Stress4{0.0, 0.0, 0.0, 0.0});
    st.RKe[ks] = 0.0;
    vn = make_double2(0.0, 0.0);
    sn.s1 = s.s1 + 0. * (P.dt) * 0.0;
    sn.s2 = s.s2 + 0. * (P.dt) * 0.0;
    sn.s3 = s.s3 + 0. * (P.dt) * 0.0;
    sn.s4 = s.s4 + 0. * (P.dt) * 0.0;
  }
  if (P.adapt) adapt_stress(P, sn);
  apply_bcs(P, st.bc_or_not, st.bc_info, id, vn, sn);
  st2(st.V[dst], id, vn);
  st4(st.S[dst], id, sn);
}

// ------------------------------------------------------------------------------------------------------
// Sweep A: stress_point_update (+ the adapt_stress2 / BCs that follow it).
// ------------------------------------------------------------------------------------------------------
template <bool FIRST>
__global__ void __launch_bounds__(128)
k_sweep_a(DevParams P, SlotMap M, SortArrays So, ListPtrs L, const int *__restrict__ n0, StatePtrs st, int src, int dst,
          int do_adapt, int do_bc, int want_epsp) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M.nnp + M.nsp) return;
  int sp, k;
  if (!slot_decode(M, t, sp, k)) return;
  const int id = So.order[sp][k];
  const int cnt = n0[t];
  const size_t o0 = (size_t)L.off0[t / SLICE] + (t & 31);
  const double *__restrict__ Vs = st.V[src];
  const double *__restrict__ Ss = st.S[src];
  double2 v = ld2(Vs, id);
  Stress4 s = ld4(Ss, id);
  double nrm = 0.0;
  if (sp == SP_STRESS) {
    double vtx = 0.0, vty = 0.0;
    for (int e = 0; e < cnt; ++e) {
      const size_t a = o0 + (size_t)e * SLICE;
      const int q = L.idx0[a];
      if (q >= P.ntotal) continue;  // dummy partner (type 9): no part in the interpolation
      const double w = (double)L.w0[a];
      const double h2 = st.mor[q] * w;
      const double2 vq = ld2(Vs, q);
      vtx = vtx + vq.x * h2;
      vty = vty + vq.y * h2;
      if (FIRST) nrm = nrm + (w * st.mass[q]) / st.rho[q];
    }
    if (FIRST)
      st.norm[id] = nrm;
    else
      nrm = st.norm[id];
    if (nrm != 0) {
      v.x = vtx / nrm;
      v.y = vty / nrm;
    }
  } else {
    double t1 = 0.0, t2 = 0.0, t3 = 0.0, t4 = 0.0, te = 0.0;
    for (int e = 0; e < cnt; ++e) {
      const size_t a = o0 + (size_t)e * SLICE;
      const int q = L.idx0[a];
      if (q >= P.ntotal) continue;  // dummy partner (type 6)
      const double w = (double)L.w0[a];
      const double h1 = st.mor[q] * w;
      const Stress4 sq = ld4(Ss, q);
      t1 = t1 + sq.s1 * h1;
      t2 = t2 + sq.s2 * h1;
      t3 = t3 + sq.s3 * h1;
      t4 = t4 + sq.s4 * h1;
      if (want_epsp) te = te + st.epsp[q] * h1;
      if (FIRST) nrm = nrm + (w * st.mass[q]) / st.rho[q];
    }
    if (FIRST)
      st.norm[id] = nrm;
    else
      nrm = st.norm[id];
    if (nrm != 0) {
      s.s1 = t1 / nrm;
      s.s2 = t2 / nrm;
      s.s3 = t3 / nrm;
      s.s4 = t4 / nrm;
      if (want_epsp) st.epsp[id] = te / nrm;
    } else {
      v.x = 0;
      v.y = 0;
    }
  }
  if (do_adapt) adapt_stress(P, s);
  if (do_bc) apply_bcs(P, st.bc_or_not, st.bc_info, id, v, s);
  st2(st.V[dst], id, v);
  st4(st.S[dst], id, s);
  const double r = st.rho[id];
  const double r2 = r * r;
  st.sor[3 * (size_t)id] = s.s1 / r2;
  st.sor[3 * (size_t)id + 1] = s.s2 / r2;
  st.sor[3 * (size_t)id + 2] = s.s3 / r2;
}

// ------------------------------------------------------------------------------------------------------
// Sweep B: get_derivatives + plastic_terms + gravity/damping + artificial viscosity + Jaumann terms +
// RK4 stage accumulation + next-stage predictor (or the final RK4 update when `last`).
//   reads buffer b (written by sweep A), writes buffer a.
// ------------------------------------------------------------------------------------------------------
template <bool FIRST>
__global__ void __launch_bounds__(128)
k_sweep_b(DevParams P, SlotMap M, SortArrays So, ListPtrs L, const int *__restrict__ n0, const int *__restrict__ n1,
          StatePtrs st, int b, int a_, double f1next, double f2, int last) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M.nnp + M.nsp) return;
  int sp, k;
  if (!slot_decode(M, t, sp, k)) return;
  const int id = So.order[sp][k];
  const int cnt = n0[t];
  const int lane = t & 31, sl = t / SLICE;
  const size_t o0 = (size_t)L.off0[sl] + lane;
  const double *__restrict__ Vb = st.V[b];
  const double *__restrict__ Sb = st.S[b];
  const double2 vp = ld2(Vb, id);
  const Stress4 sp_ = ld4(Sb, id);
  const double2 xp = ld2(st.x, id);
  double ae1 = 0.0, ae2 = 0.0, ae3 = 0.0, ae4 = 0.0, ae5 = 1.0;

  if (sp == SP_STRESS) {
    const int ks = id - P.nnode;
    double g11 = 0.0, g12 = 0.0, g21 = 0.0, g22 = 0.0;  // grad1_tmp(d,k): d velocity component, k direction
    for (int e = 0; e < cnt; ++e) {
      const size_t a = o0 + (size_t)e * SLICE;
      const int q = L.idx0[a];
      const double gx = (double)L.gx0[a], gy = (double)L.gy0[a];
      if (q < P.ntotal) {  // type 1: q is the node
        const double mq = st.mass[q], rq = st.rho[q];
        const double h1 = gx * mq / rq;
        const double h2 = gy * mq / rq;
        const double2 vq = ld2(Vb, q);
        g11 = g11 + (vq.x - vp.x) * h1;
        g12 = g12 + (vq.x - vp.x) * h2;
        g21 = g21 + (vq.y - vp.y) * h1;
        g22 = g22 + (vq.y - vp.y) * h2;
        if (FIRST && P.cspm) {
          const double2 xq = ld2(st.x, q);
          ae1 = ae1 + (xq.x - xp.x) * h1;
          ae2 = ae2 + (xq.y - xp.y) * h1;
          ae3 = ae3 + (xq.x - xp.x) * h2;
          ae4 = ae4 + (xq.y - xp.y) * h2;
        }
      } else {  // type 9: q is a dummy wall particle (no-slip mirror velocity), main:552-575
        const double beta_max = 1.5, vel_wall = 0.0;
        const double wall = (double)st.wallpos[q];
        const double2 xq = ld2(st.x, q);
        double da, db;
        if (st.horiz[q] == 1.f) {
          da = fabs(xp.y - wall);
          db = fabs(xq.y - wall);
        } else {
          da = fabs(xp.x - wall);
          db = fabs(xq.x - wall);
        }
        const double bq = 1 + (db / da);
        const double beta = (bq < beta_max) ? bq : beta_max;
        const double dvx = vp.x * (1 - beta) + beta * vel_wall;
        const double dvy = vp.y * (1 - beta) + beta * vel_wall;
        const double mq = st.mass[q], rq = st.rho[q];
        const double h1 = gx * mq / rq;
        const double h2 = gy * mq / rq;
        g11 = g11 + (vp.x - dvx) * h1;
        g12 = g12 + (vp.x - dvx) * h2;
        g21 = g21 + (vp.y - dvy) * h1;
        g22 = g22 + (vp.y - dvy) * h2;
      }
    }
    if (P.cspm) {
      double *AEp = st.AE + 5 * (size_t)id;
      if (FIRST) {
        ae5 = ae1 * ae4 - ae2 * ae3;
        if (fabs(ae5) < P.ae_thr) {
          ae5 = 1;
          ae1 = 1;
          ae2 = 0;
          ae3 = 0;
          ae4 = 1;
        } else {
          ae5 = 1 / ae5;
        }
        AEp[0] = ae1;
        AEp[1] = ae2;
        AEp[2] = ae3;
        AEp[3] = ae4;
        AEp[4] = ae5;
      } else {
        ae1 = AEp[0];
        ae2 = AEp[1];
        ae3 = AEp[2];
        ae4 = AEp[3];
        ae5 = AEp[4];
      }
      // main:619-622: the second statement sees the already-corrected first column
      g11 = ae5 * (ae1 * g11 + ae2 * g12);
      g12 = ae5 * (ae3 * g11 + ae4 * g12);
      g21 = ae5 * (ae1 * g21 + ae2 * g22);
      g22 = ae5 * (ae3 * g21 + ae4 * g22);
    }
    // div1, main:633-636
    const double d1 = -(P.D11 * g11 + P.D12 * g22);
    const double d2 = -(P.D12 * g11 + P.D22 * g22);
    const double d3 = -(P.D33 * g21 + P.D33 * g12);
    const double d4 = -(P.D41 * g11 + P.D42 * g22);
    // plastic_terms, mat:1884-1954
    double Gs[4] = {0.0, 0.0, 0.0, 0.0}, der1 = 0.0;
    if (P.ntype_eco > 1) {
      Stress4 s2 = sp_;
      if (P.ntype_solid == 1) s2.s4 = P.props[3] * (s2.s1 + s2.s2);
      double vivel[4] = {0.0, 0.0, 0.0, 0.0};
      if (P.ncrit <= 5) {
        von_mises_perzyna(P, s2, st.epsp[id], Gs, vivel);
      } else if (P.ncrit == 12) {
        double G2[4];
        double fd = st.fdp[id];
        drucker_prager(P, s2, g11, g12, g21, g22, fd, G2, vivel);
        st.fdp[id] = fd;
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
    double rke = st.RKe[ks] + der1 * f2;
    // Jaumann terms, main:751-757
    double sp1 = 0.0, sp2 = 0.0, sp3 = 0.0, sp4 = 0.0;
    if (P.update_x) {
      const double o1 = 0.5 * (g12 - g21), o2 = -0.5 * (g12 - g21);
      sp1 = 2 * o1 * sp_.s3;
      sp2 = 2 * o2 * sp_.s3;
      sp3 = o2 * sp_.s1 + o1 * sp_.s2;
    }
    const double r1 = -d1 + sp1 + Gs[0];
    const double r2 = -d2 + sp2 + Gs[1];
    const double r3 = -d3 + sp3 + Gs[2];
    const double r4 = -d4 + sp4 + Gs[3];
    Stress4 rk = ld4(st.RKs, ks);
    rk.s1 = rk.s1 + f2 * r1;
    rk.s2 = rk.s2 + f2 * r2;
    rk.s3 = rk.s3 + f2 * r3;
    rk.s4 = rk.s4 + f2 * r4;
    const Stress4 s0 = ld4(st.stress0, ks);
    Stress4 sn;
    if (!last) {
      st4(st.RKs, ks, rk);
      st.RKe[ks] = rke;
      sn.s1 = s0.s1 + f1next * (P.dt) * r1;
      sn.s2 = s0.s2 + f1next * (P.dt) * r2;
      sn.s3 = s0.s3 + f1next * (P.dt) * r3;
      sn.s4 = s0.s4 + f1next * (P.dt) * r4;
    } else {
      sn.s1 = s0.s1 + (P.dt / 6) * rk.s1;
      sn.s2 = s0.s2 + (P.dt / 6) * rk.s2;
      sn.s3 = s0.s3 + (P.dt / 6) * rk.s3;
      sn.s4 = s0.s4 + (P.dt / 6) * rk.s4;
      // update_strain, mat:1864-1880 with Ddev_strn = RK_dev_strain/6 (main:799)
      st.epsp[id] = st.epsp[id] + P.dt * (rke / 6);
    }
    if (P.adapt) adapt_stress(P, sn);
    double2 vn = vp;
    apply_bcs(P, st.bc_or_not, st.bc_info, id, vn, sn);
    st2(st.V[a_], id, vn);
    st4(st.S[a_], id, sn);
  } else {
    // ---------------- node ----------------
    const double *sorp = st.sor + 3 * (size_t)id;
    const double so1 = sorp[0], so2 = sorp[1], so3 = sorp[2];
    const double rp = st.rho[id];
    double a11 = 0.0, a12 = 0.0, a21 = 0.0, a22 = 0.0, a31 = 0.0, a32 = 0.0;  // grad2_tmp(s,k)
    for (int e = 0; e < cnt; ++e) {
      const size_t a = o0 + (size_t)e * SLICE;
      const int q = L.idx0[a];
      const double gx = (double)L.gx0[a], gy = (double)L.gy0[a];
      const double mq = st.mass[q];
      double q1, q2, q3;
      if (q < P.ntotal) {  // type 1: q is the stress particle
        const double *sorq = st.sor + 3 * (size_t)q;
        q1 = sorq[0];
        q2 = sorq[1];
        q3 = sorq[2];
        if (FIRST && P.cspm) {
          const double rq = st.rho[q];
          const double h1b = -gx * mq / rq;
          const double h2b = -gy * mq / rq;
          const double2 xq = ld2(st.x, q);
          ae1 = ae1 + (xq.x - xp.x) * h1b;
          ae2 = ae2 + (xq.y - xp.y) * h1b;
          ae3 = ae3 + (xq.x - xp.x) * h2b;
          ae4 = ae4 + (xq.y - xp.y) * h2b;
        }
      } else {  // type 6: dummy takes the node's stress (main:580)
        const double rq = st.rho[q];
        q1 = sp_.s1 / (rq * rq);
        q2 = sp_.s2 / (rq * rq);
        q3 = sp_.s3 / (rq * rq);
      }
      const double c1 = so1 + q1, c2 = so2 + q2, c3 = so3 + q3;
      a11 = a11 - mq * (gx * c1);
      a12 = a12 - mq * (gy * c1);
      a21 = a21 - mq * (gx * c2);
      a22 = a22 - mq * (gy * c2);
      a31 = a31 - mq * (gx * c3);
      a32 = a32 - mq * (gy * c3);
    }
    if (P.cspm) {
      double *AEp = st.AE + 5 * (size_t)id;
      if (FIRST) {
        ae5 = ae1 * ae4 - ae2 * ae3;
        if (fabs(ae5) < P.ae_thr) {
          ae5 = 1;
          ae1 = 1;
          ae2 = 0;
          ae3 = 0;
          ae4 = 1;
        } else {
          ae5 = 1 / ae5;
        }
        AEp[0] = ae1;
        AEp[1] = ae2;
        AEp[2] = ae3;
        AEp[3] = ae4;
        AEp[4] = ae5;
      } else {
        ae1 = AEp[0];
        ae2 = AEp[1];
        ae3 = AEp[2];
        ae4 = AEp[3];
        ae5 = AEp[4];
      }
      // main:623-626: only stress components 1..ndimn are corrected
      a11 = ae5 * (ae1 * a11 + ae2 * a12);
      a12 = ae5 * (ae3 * a11 + ae4 * a12);
      a21 = ae5 * (ae1 * a21 + ae2 * a22);
      a22 = ae5 * (ae3 * a21 + ae4 * a22);
    }
    const double dv1 = -(a11 + a32);  // div2, main:641-642
    const double dv2 = -(a31 + a22);
    // gravity_force, mat:2809-2871
    const double sg1 = P.grav[0] - P.damping * vp.x;
    const double sg2 = P.grav[1] - P.damping * vp.y;
    // artificial_viscosity, main:826-904 (fp32 locals and accumulators, list order)
    double av1 = 0.0, av2 = 0.0;
    if (P.alpha > 0 || P.beta > 0) {
      const int cntc = n1[t];
      const size_t oc = (size_t)L.offC[sl] + lane;
      const double hp = st.hsml[id];
      float acc1 = 0.f, acc2 = 0.f;
      for (int e = 0; e < cntc; ++e) {
        const size_t a = oc + (size_t)e * SLICE;
        const int q = L.idxC[a];
        const float gxf = L.gxC[a], gyf = L.gyC[a];
        const double2 xq = ld2(st.x, q);
        const double2 vq = ld2(Vb, q);
        const float xij = (float)(xp.x - xq.x);
        const float yij = (float)(xp.y - xq.y);
        const float h = (float)(0.5 * (hp + st.hsml[q]));
        const float rho2 = (float)(0.5 * (rp + st.rho[q]));
        const float cs = 600.f;
        float div_u = (float)((double)xij * (vp.x - vq.x));
        div_u = (float)((double)div_u + (double)yij * (vp.y - vq.y));
        const float sq = sqrtf(xij * xij + yij * yij);
        const float theta = (h * div_u) / (sq * sq + 0.01f * (h * h));
        float visc = 0.f;
        if (div_u < 0)
          visc = (float)((-P.alpha * (double)cs * (double)theta + P.beta * (double)(theta * theta)) / (double)rho2);
        const double mq = st.mass[q];
        acc1 = (float)((double)acc1 + (double)(visc * gxf) * mq);
        acc2 = (float)((double)acc2 + (double)(visc * gyf) * mq);
      }
      av1 = (double)(-acc1);
      av2 = (double)(-acc2);
    }
    const double r1 = -dv1 + sg1 + av1 + 0.0 + 0.0;  // + f_bound + art_force (both zero here, main:763-764)
    const double r2 = -dv2 + sg2 + av2 + 0.0 + 0.0;
    double2 rk = ld2(st.RKv, id);
    rk.x = rk.x + f2 * r1;
    rk.y = rk.y + f2 * r2;
    const double2 v0 = ld2(st.vel0, id);
    double2 vn;
    if (!last) {
      st2(st.RKv, id, rk);
      vn.x = v0.x + f1next * (P.dt) * r1;
      vn.y = v0.y + f1next * (P.dt) * r2;
    } else {
      vn.x = v0.x + (P.dt / 6) * rk.x;
      vn.y = v0.y + (P.dt / 6) * rk.y;
    }
    Stress4 sn = sp_;
    if (P.adapt) adapt_stress(P, sn);
    apply_bcs(P, st.bc_or_not, st.bc_info, id, vn, sn);
    st2(st.V[a_], id, vn);
    st4(st.S[a_], id, sn);
  }
}

// ------------------------------------------------------------------------------------------------------
// Position update, main:140-182: XSPH_update (main:189-239) or the fp32 mid-velocity rule; displ.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_move(DevParams P, SlotMap M, SortArrays So, ListPtrs L, const int *__restrict__ n1, StatePtrs st, int cur,
       double *__restrict__ x, const double *__restrict__ x00, double *__restrict__ displ) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M.nnp + M.nsp) return;
  int sp, k;
  if (!slot_decode(M, t, sp, k)) return;
  const int id = So.order[sp][k];
  const double *__restrict__ Vc = st.V[cur];
  const double2 vp = ld2(Vc, id);
  const double2 xp = ld2(x, id);
  if (P.update_x) {
    double2 xn;
    if (P.xsph) {
      const int cnt = n1[t];
      const int lane = t & 31, sl = t / SLICE;
      double sx = 0.0, sy = 0.0;
      if (sp == SP_NODE) {
        const size_t oc = (size_t)L.offC[sl] + lane;
        for (int e = 0; e < cnt; ++e) {
          const size_t a = oc + (size_t)e * SLICE;
          const int q = L.idxC[a];
          const double w = (double)L.wC[a];
          const double2 vq = ld2(Vc, q);
          const double mr = st.mor[q];
          sx = sx + mr * (vq.x - vp.x) * w;
          sy = sy + mr * (vq.y - vp.y) * w;
        }
      } else {
        const size_t od = (size_t)L.offD[sl] + lane;
        for (int e = 0; e < cnt; ++e) {
          const size_t a = od + (size_t)e * SLICE;
          const int q = L.idxD[a];
          const double w = (double)L.wD[a];
          const double2 vq = ld2(Vc, q);
          const double mr = st.mor[q];
          sx = sx + mr * (vq.x - vp.x) * w;
          sy = sy + mr * (vq.y - vp.y) * w;
        }
      }
      const double eps = 0.5;
      xn.x = xp.x + P.dt * (vp.x + eps * sx);
      xn.y = xp.y + P.dt * (vp.y + eps * sy);
    } else {
      const double2 v0 = ld2(st.vx0, id);
      const float hx = (float)(0.5 * (v0.x + vp.x));  // real :: vel_half, main:89,145
      const float hy = (float)(0.5 * (v0.y + vp.y));
      xn.x = xp.x + (double)hx * P.dt;
      xn.y = xp.y + (double)hy * P.dt;
    }
    st2(x, id, xn);
    if (sp == SP_NODE) {
      const double2 x0 = ld2(x00, id);
      st2(displ, id, make_double2(xn.x - x0.x, xn.y - x0.y));  // main:171
    }
  } else if (sp == SP_NODE) {
    const double2 v0 = ld2(st.vx0, id);
    double2 d = ld2(displ, id);
    d.x = d.x + 0.5 * (v0.x + vp.x) * P.dt;  // main:180
    d.y = d.y + 0.5 * (v0.y + vp.y) * P.dt;
    st2(displ, id, d);
  }
}

// shift_stress_points, main:244-368 (outside approach): one thread per node re-seats its stress particles.
__global__ void k_shift(DevParams P, const double *__restrict__ Vc, double *__restrict__ x, double *__restrict__ x_10,
                        double *__restrict__ disp_10, const int *__restrict__ bc_int, const float *__restrict__ n_int) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.nnode) return;
  const double2 xi = ld2(x, i);
  const double2 v = ld2(Vc, i);
  const int k = P.nnode + i * P.npoints;  // 0-based id of this node's first stress particle
  const double dx = P.dx;
  if (P.itimestep % P.shift_update == 0) {
    if (P.vel_vector) {
      const double2 xo = ld2(x_10, i);
      const double ddx = xi.x - xo.x, ddy = xi.y - xo.y;
      double d10 = sqrt(ddx * ddx + ddy * ddy);
      d10 = d10 / dx;
      disp_10[i] = d10;
      st2(x_10, i, xi);
      const double vn = sqrt(v.x * v.x + v.y * v.y);
      const float cos_theta = (float)(v.x / vn), sin_theta = (float)(v.y / vn);
      float r1 = (float)((dx / 2) * (double)cos_theta), r2 = (float)((dx / 2) * (double)sin_theta);
      if (r1 > 0 && (double)r1 < dx / 5) r1 = (float)((double)r1 + dx / 3.);
      if (r2 > 0 && (double)r2 < dx / 5) r2 = (float)((double)r2 + dx / 3.);
      if (r1 < 0 && (double)r1 > -dx / 5) r1 = (float)((double)r1 - dx / 3.);
      if (r2 < 0 && (double)r2 > -dx / 5) r2 = (float)((double)r2 - dx / 3.);
      if (d10 > P.disp_tol) {
        st2(x, k, make_double2(xi.x + (double)r1, xi.y + (double)r2));
        st2(x, k + 1, make_double2(xi.x - (double)r1, xi.y - (double)r2));
      }
    } else {
      const double r_x = P.r_x, r_y = P.r_y;
      if (P.npoints == 1) {
        st2(x, k, make_double2(xi.x + r_x, xi.y + r_y));
      } else if (P.npoints == 2) {
        st2(x, k, make_double2(xi.x + r_x, xi.y + r_y));
        st2(x, k + 1, make_double2(xi.x - r_x, xi.y - r_x));  // sic, main:309
      } else if (P.npoints == 3) {
        st2(x, k, make_double2(xi.x, xi.y + r_y));
        st2(x, k + 1, make_double2(xi.x - r_x, xi.y - r_y));
        st2(x, k + 2, make_double2(xi.x + r_x, xi.y - r_y));
      }
    }
  }
  const float abs_vel = (float)sqrt(v.x * v.x + v.y * v.y);
  bool collapse = false;
  if (bc_int[i] == 1 && abs_vel > 0.4f) collapse = true;
  if (n_int[i] < 2) collapse = true;
  if (collapse)
    for (int q = 0; q < P.npoints && q < 3; ++q) st2(x, k + q, xi);
}

// pair statistics of grid_find_NEW, main:1405-1422
__global__ void k_pair_stats(SlotMap M, const int *__restrict__ nall, int *__restrict__ out /* max,min,zero */) {
  int mx = 0, mn = 1000, nz = 0;
  const int n = M.total();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int sp, k;
    if (!slot_decode(M, i, sp, k)) continue;  // padding slot
    const int c = nall[i];
    mx = max(mx, c);
    mn = min(mn, c);
    nz += (c == 0);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    nz += __shfl_xor_sync(0xffffffffu, nz, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&out[0], mx);
    atomicMin(&out[1], mn);
    atomicAdd(&out[2], nz);
  }
}

}  // namespace spsph
