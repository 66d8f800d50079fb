// TEST INFRASTRUCTURE. Host emulation of the device kernels that walk the per-particle gather lists outside the
// cp.async / warp-shuffle machinery -- k_free_surface, k_xsph_marks, k_fs_normals (csrc/step_kernels.cuh). They are
// one-thread-per-particle scalar code; this file compiles them for the host (SPSPH_HOST_EMU hides the rest of the
// kernel headers, blockIdx / threadIdx become thread-local variables) and runs them thread by thread on list
// structures that tests/test_list_kernels_cpu.py builds from the ORACLE's pair list in the layout the fill pass
// writes (warp-sliced ELL, species-sorted arrays, creation keys). What it proves: the traversal logic and the
// arithmetic of these kernels against the oracle without a GPU. What it cannot: that k_count / k_fill produce these
// lists (the GPU parity tests do that).
//   g++ -O1 -ffp-contract=off -std=c++17 -fPIC -shared -w -D__noinline__= -I/usr/local/cuda/include -I<csrc> -I<include>
#include "cuda_host_emu.h"

#include "step_kernels.cuh"

using namespace spsph;

extern "C" struct EmuArgs {
  // scalars
  int32_t nnode, nstress, ndummy, skf, growth_mode, pad;
  double pi;
  uint64_t growth_ka, growth_kb;  // mode 2: creation keys (creator, partner) of the last old pair
  // species-sorted arrays (velocity, stress, wall particles)
  const int32_t *order[3];
  const int32_t *cell[3];
  const double *pos[3];  // (2, n_s) positions the lists were built with
  const double *h[3];
  const int32_t *pos_of;  // [ntotal2]
  // lists
  const int32_t *idx0, *idxC, *idxD, *off0, *offC, *offD, *n0, *n1;
  const float *gx0, *gy0, *gxC, *gyC;
  // particle arrays, original numbering
  const double *x, *mass, *rho, *hsml;
  int32_t *bc_or_not;  // in/out
  int32_t *covered;    // out (k_free_surface), in (k_fs_normals)
  double *fs_normal;   // out (k_fs_normals)
};

static void fill(const EmuArgs &a, DevParams &P, SlotMap &M, SortArrays &S, ListPtrs &L, GrowthRule &g) {
  std::memset(&P, 0, sizeof(P));
  P.nnode = a.nnode;
  P.nstress = a.nstress;
  P.ntotal = a.nnode + a.nstress;
  P.ntotal2 = P.ntotal + a.ndummy;
  P.ndummy = a.ndummy;
  P.skf = a.skf;
  P.scale_k = a.skf == 1 ? 2 : 3;
  P.pi = a.pi;
  M.nn = a.nnode;
  M.ns = a.nstress;
  M.nd = a.ndummy;
  M.nnp = (a.nnode + 31) / 32 * 32;
  M.nsp = (a.nstress + 31) / 32 * 32;
  M.ndp = (a.ndummy + 31) / 32 * 32;
  for (int s = 0; s < 3; ++s) {
    S.start[s] = nullptr;
    S.order[s] = a.order[s];
    S.pos[s] = reinterpret_cast<const double2 *>(a.pos[s]);
    S.upos[s] = nullptr;
    S.h[s] = a.h[s];
    S.cell[s] = a.cell[s];
  }
  std::memset(&L, 0, sizeof(L));
  L.idx0 = const_cast<int *>(a.idx0);
  L.gx0 = const_cast<float *>(a.gx0);
  L.gy0 = const_cast<float *>(a.gy0);
  L.idxC = const_cast<int *>(a.idxC);
  L.gxC = const_cast<float *>(a.gxC);
  L.gyC = const_cast<float *>(a.gyC);
  L.idxD = const_cast<int *>(a.idxD);
  L.off0 = a.off0;
  L.offC = a.offC;
  L.offD = a.offD;
  std::memset(&g, 0, sizeof(g));
  g.mode = a.growth_mode;
  g.ka = a.growth_ka;
  g.kb = a.growth_kb;
}

template <class F>
static void launch(int nthreads, F f) {
  emu_blockDim = {128, 1, 1};
  emu_gridDim = {(unsigned)((nthreads + 127) / 128), 1, 1};
  for (int t = 0; t < nthreads; ++t) {
    emu_blockIdx = {(unsigned)(t / 128), 0, 0};
    emu_threadIdx = {(unsigned)(t % 128), 0, 0};
    f();
  }
}

extern "C" {

void emu_xsph_marks(const EmuArgs *a) {
  DevParams P;
  SlotMap M;
  SortArrays S;
  ListPtrs L;
  GrowthRule g;
  fill(*a, P, M, S, L, g);
  launch(M.nn, [&] { k_xsph_marks(P, M, S, L, a->n1, a->bc_or_not); });
}

void emu_free_surface(const EmuArgs *a) {
  DevParams P;
  SlotMap M;
  SortArrays S;
  ListPtrs L;
  GrowthRule g;
  fill(*a, P, M, S, L, g);
  launch(M.nnp + M.nsp, [&] {
    k_free_surface(P, M, S, a->pos_of, L, a->n0, a->n1, &g, a->x, a->mass, a->rho, a->hsml, a->bc_or_not, a->covered);
  });
}

void emu_fs_normals(const EmuArgs *a) {
  DevParams P;
  SlotMap M;
  SortArrays S;
  ListPtrs L;
  GrowthRule g;
  fill(*a, P, M, S, L, g);
  launch(M.nn, [&] {
    k_fs_normals(P, M, S, a->pos_of, L, a->n0, a->n1, &g, a->x, a->mass, a->rho, a->bc_or_not, a->covered, a->fs_normal);
  });
}
}
