// libspsph_cuda.so -- C-ABI (include/spsph.h) of the B200-native Stress-Particle SPH time-step engine.
//
// Host orchestration of one time_integration (2_SPH_main_2018.f90:78-184) as a fixed sequence of kernel
// launches on one stream; the only host round trip per step is the read-back of the pair / list totals
// after the count pass (needed to size the neighbour lists and to apply the reference's list-growth rule).
// There is no CPU fallback: every entry point fails if CUDA is unavailable.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <chrono>
#include <string>
#ifndef SPSPH_HOST_EMU
#include <thread>
#endif
#include <type_traits>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>

#include "spsph.h"
#include "dist_kernels.cuh"
#include "tile_kernels.cuh"

using namespace spsph;

#define NCCL_TRY(call)                                                                            \
  do {                                                                                            \
    ncclResult_t r_ = (call);                                                                     \
    if (r_ != ncclSuccess) {                                                                      \
      h->err = std::string(#call) + ": " + (h->p_ncclGetErrorString ? h->p_ncclGetErrorString(r_) : "nccl error"); \
      return 1;                                                                                   \
    }                                                                                             \
  } while (0)

#define CUDA_TRY(call)                                                                            \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) {                                                                      \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                \
      return 1;                                                                                   \
    }                                                                                             \
  } while (0)

#if defined(SPSPH_HOST_EMU) && !defined(SPSPH_EMU_SIMT)
#define SPSPH_EMU_SERIAL 1  // threads run one after the other: cooperative kernels are replaced by host loops
#endif
#ifdef SPSPH_HOST_EMU
// what build_neighbours produces on the device, handed over by the host-emulation harness instead (arrays of the
// caller, read during the next spsph_step): species-sorted arrays, warp-sliced ELL lists, growth rule
struct spsph_emu_lists {
  int64_t n_pairs;
  int32_t growth_mode, pad;
  uint64_t growth_ka, growth_kb;
  const int32_t *order[3], *cell[3];
  const double *pos[3], *h[3];
  const int32_t *pos_of;
  int64_t tot0, totC, totD;
  const int32_t *idx0;
  const float *w0, *gx0, *gy0;
  const int32_t *idxC;
  const float *wC, *gxC, *gyC, *xC, *yC, *hC;
  const int32_t *idxD;
  const float *wD;
  const int32_t *off0, *offC, *offD, *n0, *n1;
  const int32_t *bc_int;
  const float *n_int;
  const int32_t *if_out;  // [ntotal2] Check_Out_Domain flags after this step's check (k_domain_bbox sets them on the device)
};
#endif

struct spsph_handle {
  spsph_params hp{};
  DevParams P{};
  SlotMap M{};
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // node-side sweeps run beside the stress-particle-side ones (independent data)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool dual = true;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  std::vector<void *> allocs;
  bool uploaded = false;
  bool umor = false;  // mass/rho takes <= 4 values within the velocity particles and within the stress particles
  MorPalette pal_node{}, pal_sp{};
  unsigned char *mcls = nullptr;  // [ntotal2] palette class of every particle
  std::vector<double> up_mor;        // host scratch of spsph_upload (mass/rho per particle)
  std::vector<unsigned char> up_cls;  // host scratch of spsph_upload (palette classes)
  bool uniform_h = false;         // one smoothing length for every particle, constant in time
  float h_uniform = 0.f;          // (float)(0.5*(h + h)) of artificial_viscosity, main:863
  bool uniform_cubic = false;  // skf = 1 and one smoothing length for every particle (set at upload)
  bool have_lists = false;  // the pair lists of a completed step are on the device (free-surface detection)

  // particle state (original order)
  double *x = nullptr, *x00 = nullptr, *rho = nullptr, *mass = nullptr, *hsml = nullptr, *mor = nullptr;
  double2 *mrho = nullptr;  // {mass, rho} per particle
  Rec4 *NB[2] = {nullptr, nullptr}, *SB[2] = {nullptr, nullptr}, *SVb[2] = {nullptr, nullptr}, *SA = nullptr;
  double *NA = nullptr, *NSa = nullptr, *SVa = nullptr, *av = nullptr, *fbound = nullptr, *aforce = nullptr;
  Rec4 *RN = nullptr;
  double *rho_new = nullptr, *rho0 = nullptr, *hsml0 = nullptr, *RKrho = nullptr, *RKh = nullptr, *divu = nullptr;  // cont_density
  double art_w2 = 0.0;  // kernel(dx, (dx, dy), 1.2 dx): the reference spacing weight of artificial_force (main:926)
  double *NSb[2] = {nullptr, nullptr}, *SFb[2] = {nullptr, nullptr};
  double *stage_vel = nullptr, *stage_stress = nullptr;  // reference-layout staging for upload / download
  double *epsp = nullptr, *fdp = nullptr, *norm = nullptr, *AE = nullptr;
  double *vel0 = nullptr, *stress0 = nullptr, *vx0 = nullptr, *RKv = nullptr, *RKs = nullptr, *RKe = nullptr;
  double *displ = nullptr, *x_10 = nullptr, *disp_10 = nullptr;
  float *wallpos = nullptr, *horiz = nullptr, *n_int = nullptr;
  int *bc_int = nullptr, *bc_or_not = nullptr, *bc_info = nullptr, *if_out = nullptr;
  // get_nodes_on_free_surface every step (instead of on demand at download): needed when its marks feed back into
  // the time step -- apply_stress_free (ifsigman = 1) or XSPH stripping boundary conditions (main:224-230)
  bool fs_each_step = false;
  int *fs_cov = nullptr;        // (ntotal) covered in step 3 (f_int = -1)
  double *fs_normal = nullptr;  // (2, nnode) step-4 normals
  // outside approach: positions between the position update and shift_stress_points, where the reference runs
  // get_nodes_on_free_surface (main:152-160); the on-demand evaluation at download classifies on this snapshot
  double *x_fs = nullptr;       // (2, ntotal2), single GPU only
  bool x_fs_valid = false;
#ifdef SPSPH_HOST_EMU
  const struct spsph_emu_lists *emu_lists = nullptr;  // this step's lists, handed over by the test harness
#endif
  int cur = 0;
  double *ivars = nullptr;  // device mirror of Internal_Vars(10, ntotal): rows 2..10 never change on the hot path
  std::vector<int32_t> h_itype;

  // grid / sort
  GridInfo *G = nullptr;
  double *bbox_partial = nullptr;
  int bbox_blocks = 0;
  int cell_capacity = 0, cell_stride = 0;
  int *cell_cnt = nullptr, *cell_start = nullptr, *cell_fill = nullptr;
  int *ordc = nullptr;  // order with the mass/rho class in the top bits (k_rank; what k_fill stores in list 0)
  int *which_cell = nullptr, *tmp_ids = nullptr, *order = nullptr, *scell = nullptr, *pos_of = nullptr, *nout = nullptr;
  double2 *spos = nullptr;
  float2 *supos = nullptr;
  double *sh = nullptr;
  int *scan_bsum = nullptr;
  long long *scan_totals = nullptr;  // [0..2] list storage, [3] pairs, [4..6] cells
  // lists
  int *n0 = nullptr, *n1 = nullptr, *nall = nullptr, *nfwd_u = nullptr, *base_u = nullptr;
  int *cand0 = nullptr, *cand1 = nullptr, *cand_overflow = nullptr;  // accepted partners recorded by k_count
  bool force_fill_scan = false;  // SPSPH_FORCE_FILL_SCAN=1: always take the overflow path (tests)
  int *wslice = nullptr, *oslice = nullptr;  // 3 rows x nslices
  int nslices = 0;
  GrowthRule *growth = nullptr;
  long long cap0 = 0, capC = 0, capD = 0;
  ListPtrs L{};
  StepStatus *status_d = nullptr, *status_h = nullptr;
  int *stats_d = nullptr;

  // ---- cell-tile path (tile_kernels.cuh): acceptance masks + fp32 weights instead of partner-id lists ----
  bool tile_cfg = false;    // the option combination is covered by the tile kernels (decided at upload)
  bool tile_off = false;    // a stencil row outgrew the 64-bit masks: list path until the next upload
  bool tile_env = false;    // SPSPH_TILE=1 selects the cell-tile path where the options allow it (default: id lists)
  bool tile_last = false;   // the last step ran on the tile path (the id lists are materialised on demand)
  int tile_last_mode = 0;   // its traversal order: 0 forward, 1 reversed
  TileLists TL{};
  TileRecs TR{};
  double *smor = nullptr, *srrho = nullptr;  // species-sorted per-step copies of mass/rho and RN(1/rho)
  double2 *smrho = nullptr;                  // ... and {mass, rho}
  long long tile_w0_cap = 0, tile_c_cap = 0; // allocated float4 groups
  u64 *tile_acc = nullptr;                   // device: forward pair count of the step
  int *tile_flags = nullptr;                 // device: [0] mask overflow, [1..4] longest lists, [5] slice overflow
  GeomTables GT{};                           // per-block tile geometry of the pair-sum kernels, rebuilt every step
  long long tile_s_cap = 0;
  TileStatus *tstat_d = nullptr, *tstat_h = nullptr;
  long long tile_steps = 0, list_steps = 0;  // which path the steps took (spsph_path_counts)

  // row-wise transfers (spsph_upload_rows / spsph_download_rows): device copy of the row ids + one staging buffer
  int *rows_ids = nullptr;
  char *rows_buf = nullptr;
  size_t rows_cap = 0;
  // halo peeling (dist_kernels.cuh, k_peel_counts): per-level copies of the list-length arrays n0 / n1
  int *peel0 = nullptr, *peel1 = nullptr;
  size_t peel_stride = 0;
  int halo_cells = 0;
  bool peel = false;
  double *frame_buf = nullptr;  // spsph_download_frame: the packed (count, ncols) table
  size_t frame_cap = 0;

  long long m_pairs = 0;       // max pair count of all previous steps (main:1210)
  long long last_n_pairs = 0;  // of the last step
  long long last_m_before = 0;
  float last_ms = 0.f;
  long long last_launches = 0, launches = 0;

  // multi-GPU x-slab decomposition (dist_kernels.cuh); NCCL is resolved with dlopen so that a single-GPU user
  // needs no NCCL and a torch process re-uses the libnccl it has already loaded
  bool dist = false;
  DistGeom D{};
  int *lflag = nullptr;       // [ntotal2] 0 remote, 1 owned, 2 ghost (nullptr-equivalent when !dist)
  int *halo_cnt = nullptr;    // [2] + err flag [1]
  int *list_ids[2] = {nullptr, nullptr};  // local list (double-buffered for the per-step compaction)
  int *list_n = nullptr;                  // [2] device counts
  int *list_keep = nullptr, *list_pos = nullptr;  // [ntotal2] scratch of the order-preserving compaction
  int list_cur = 0;
  bool nloc_valid = false;                // nloc[] holds the previous step's counts (bounds the k_count grid)
  // record counts of the previous step's halo messages (sent / received, left / right): they size this step's
  // NCCL transfers; both ends of a link derive the same size from the same number (the message header)
  int halo_prev_send[2] = {0, 0}, halo_prev_recv[2] = {0, 0};
  bool halo_prev_valid = false;
  int halo_full_msgs = 0;  // exchanges still to run with full-capacity messages after a change of the slab planes
  int halo_slack = 4096;   // records a halo message may grow by from one step to the next, on top of 12.5 % (SPSPH_HALO_SLACK)
  double *dist_h = nullptr;               // pinned: [0..1] received header counts, [2..3] halo_cnt + error flags (as int)
  int *halo_ids[2] = {nullptr, nullptr};
  double *halo_send[2] = {nullptr, nullptr}, *halo_recv[2] = {nullptr, nullptr};
  double *bb6 = nullptr;
  long long *gt_buf = nullptr, *gt_mine = nullptr, *gt_all = nullptr, *gt_out2 = nullptr;  // growth-rule search
  GtSel *gt_sel = nullptr;
  void *nccl_lib = nullptr;
  ncclComm_t comm = nullptr;
  ncclResult_t (*p_ncclCommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*p_ncclCommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*p_ncclSend)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*p_ncclRecv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*p_ncclAllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                                  cudaStream_t) = nullptr;
  ncclResult_t (*p_ncclAllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*p_ncclGroupStart)() = nullptr;
  ncclResult_t (*p_ncclGroupEnd)() = nullptr;
  const char *(*p_ncclGetErrorString)(ncclResult_t) = nullptr;
  int nloc[3] = {0, 0, 0};  // particles per species to process in the current step

  // optional per-kernel timing (CUDA events on the engine stream between launches)
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<int> ev_kid;
  size_t ev_used = 0;
  double prof_ms[32] = {0};
  long long prof_n[32] = {0};
};

enum KernelId {
  KID_BBOX = 0, KID_GRID, KID_ZERO, KID_CELLID, KID_SCAN, KID_SCATTER, KID_RANK, KID_COUNT, KID_STATUS, KID_THRESH,
  KID_FILL, KID_RKBEGIN, KID_SWEEPA, KID_SWEEPB, KID_MOVE, KID_SHIFT, KID_HALO, KID_TBUILD, KID_N
};
static const char *kKernelNames[KID_N] = {"k_domain_bbox", "k_grid_params", "k_zero_cells", "k_cell_id", "k_scan_*",
                                          "k_scatter", "k_rank", "k_count", "k_status", "k_growth_threshold", "k_fill",
                                          "k_rk_begin", "k_sweep_a", "k_sweep_b", "k_move", "k_shift",
                                          "halo_exchange", "k_tile_build"};

namespace {

__global__ void k_upload_derive(int n2, int nt, const double *__restrict__ mass, const double *__restrict__ rho,
                                double *__restrict__ mor, double2 *__restrict__ mrho, const double *__restrict__ ivars,
                                double *__restrict__ epsp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n2) {
    mor[i] = mass[i] / rho[i];  // same IEEE division as the reference's mass(j)/rho(j)
    mrho[i] = make_double2(mass[i], rho[i]);  // one 16-byte gather where a sweep needs both
  }
  if (i < nt) epsp[i] = ivars[(size_t)SPSPH_NINT_VARS * i];
}
__global__ void k_download_ivars(int nt, const double *__restrict__ epsp, double *__restrict__ ivars) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nt) ivars[(size_t)SPSPH_NINT_VARS * i] = epsp[i];
}


// ---- row-wise transfers: rows k = 0..n-1 of a compact array <-> particle ids[k] of the device arrays ----
template <class T>
__global__ void k_rows_scatter(int n, int width, const int *__restrict__ ids, int id_base, const T *__restrict__ src,
                               T *__restrict__ dst) {
  const long long m = (long long)n * width;
  for (long long a = blockIdx.x * (long long)blockDim.x + threadIdx.x; a < m; a += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(a / width), c = (int)(a - (long long)k * width);
    dst[(size_t)(ids[k] - id_base) * width + c] = src[a];
  }
}
template <class T>
__global__ void k_rows_gather(int n, int width, const int *__restrict__ ids, int id_base, const T *__restrict__ src,
                              T *__restrict__ dst) {
  const long long m = (long long)n * width;
  for (long long a = blockIdx.x * (long long)blockDim.x + threadIdx.x; a < m; a += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(a / width), c = (int)(a - (long long)k * width);
    dst[a] = src[(size_t)(ids[k] - id_base) * width + c];
  }
}
// compact reference-layout vel (2, n) / stress (4, n) rows -> state between steps (format B), and back
__global__ void k_rows_pack(DevParams P, int n, const int *__restrict__ ids, const double *__restrict__ vel,
                            const double *__restrict__ stress, StatePtrs st) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int id = ids[k];
  if (id >= P.ntotal) return;
  const double2 v = ld2(vel, k);
  const Stress4 s = ld4(stress, k);
  if (id < P.nnode) {
    strec(st.NB, id, v.x, v.y, st.mass[id], st.rho[id]);
    st4(st.NSb, id, s);
  } else {
    const int ks = id - P.nnode;
    const double r = st.rho[id];
    const double r2 = r * r;
    strec(st.SB, ks, s.s1 / r2, s.s2 / r2, s.s3 / r2, st.mass[id]);
    st4(st.SFb, ks, s);
    strec(st.SVb, ks, v.x, v.y, st.mor[id], 0.0);
  }
}
__global__ void k_rows_unpack(DevParams P, int n, const int *__restrict__ ids, StatePtrs st, double *__restrict__ vel,
                              double *__restrict__ stress) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int id = ids[k];
  double2 v = make_double2(0.0, 0.0);
  Stress4 s{0.0, 0.0, 0.0, 0.0};  // wall particles carry no velocity / stress on the device
  if (id < P.nnode) {
    const Rec4 r = ldrec(st.NB, id);
    v = make_double2(r.a, r.b);
    s = ld4(st.NSb, id);
  } else if (id < P.ntotal) {
    const Rec4 r = ldrec(st.SVb, id - P.nnode);
    v = make_double2(r.a, r.b);
    s = ld4(st.SFb, id - P.nnode);
  }
  if (vel) st2(vel, k, v);
  if (stress) st4(stress, k, s);
}
__global__ void k_rows_epsp(int n, const int *__restrict__ ids, int ntotal, const double *__restrict__ ivars,
                            double *__restrict__ epsp) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n && ids[k] < ntotal) epsp[ids[k]] = ivars[(size_t)SPSPH_NINT_VARS * ids[k]];
}

template <class T>
int dalloc(spsph_handle *h, T **p, size_t n) {
  void *q = nullptr;
  if (n == 0) n = 1;
  CUDA_TRY(cudaMalloc(&q, n * sizeof(T)));
  h->allocs.push_back(q);
  *p = (T *)q;
  return 0;
}

int round_up(int a, int b) { return ((a + b - 1) / b) * b; }

// bookkeeping after every kernel launch: launch counter (+ an event when profiling)
void mark(spsph_handle *h, int kid, int n = 1) {
  h->launches += n;
  if (!h->profiling) return;
  if (h->ev_used == h->ev_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    h->ev_pool.push_back(e);
    h->ev_kid.push_back(0);
  }
  h->ev_kid[h->ev_used] = kid;
  cudaEventRecord(h->ev_pool[h->ev_used], h->stream);
  ++h->ev_used;
}
// fold the recorded events into per-kernel totals (interval since the previous event on the stream)
void prof_collect(spsph_handle *h) {
  if (!h->profiling || h->ev_used < 2) {
    h->ev_used = 0;
    return;
  }
  cudaEventSynchronize(h->ev_pool[h->ev_used - 1]);
  for (size_t i = 1; i < h->ev_used; ++i) {
    if (h->ev_kid[i] < 0) continue;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev_pool[i - 1], h->ev_pool[i]);
    h->prof_ms[h->ev_kid[i]] += ms;
    h->prof_n[h->ev_kid[i]] += 1;
  }
  h->ev_used = 0;
}

SortArrays sort_arrays(spsph_handle *h) {
  SortArrays S;
  const size_t n2 = (size_t)h->P.ntotal2;
  for (int s = 0; s < 3; ++s) {
    S.start[s] = h->cell_start + (size_t)s * h->cell_stride;
    S.order[s] = h->order + s * n2;
    S.ordc[s] = h->ordc + s * n2;
    S.pos[s] = h->spos + s * n2;
    S.upos[s] = h->supos + s * n2;
    S.h[s] = h->sh + s * n2;
    S.cell[s] = h->scell + s * n2;
  }
  return S;
}

// `wb`: format-B buffer set that is written / current; the other set is the read side of a B -> B sweep
StatePtrs state_ptrs(spsph_handle *h, int wb) {
  StatePtrs s;
  s.x = h->x;
  s.mass = h->mass;
  s.rho = h->rho;
  s.hsml = h->hsml;
  s.mor = h->mor;
  s.mrho = h->mrho;
  s.wallpos = h->wallpos;
  s.horiz = h->horiz;
  s.bc_or_not = h->bc_or_not;
  s.bc_info = h->bc_info;
  s.bc_int = h->bc_int;
  s.fs_normal = h->fs_normal;
  s.NA = h->NA;
  s.SA = h->SA;
  s.NSa = h->NSa;
  s.SVa = h->SVa;
  s.NB = h->NB[wb];
  s.SB = h->SB[wb];
  s.NSb = h->NSb[wb];
  s.SFb = h->SFb[wb];
  s.SVb = h->SVb[wb];
  s.NBr = h->NB[1 - wb];
  s.NSbr = h->NSb[1 - wb];
  s.SFbr = h->SFb[1 - wb];
  s.SVbr = h->SVb[1 - wb];
  s.av = h->av;
  s.fbound = h->fbound;
  s.aforce = h->aforce;
  s.has_fbound = (h->hp.inside_approach && h->hp.ndummy2 > 0) ? 1 : 0;
  s.has_aforce = h->hp.art_stress ? 1 : 0;
  s.RN = h->RN;
  s.rho_w = h->rho;
  s.hsml_w = h->hsml;
  s.mor_w = h->mor;
  s.mrho_w = h->mrho;
  s.rho_new = h->rho_new;
  s.rho0 = h->rho0;
  s.hsml0 = h->hsml0;
  s.RKrho = h->RKrho;
  s.RKh = h->RKh;
  s.divu = h->divu;
  s.epsp = h->epsp;
  s.fdp = h->fdp;
  s.norm = h->norm;
  s.AE = h->AE;
  s.vel0 = h->vel0;
  s.stress0 = h->stress0;
  s.vx0 = h->vx0;
  s.RKv = h->RKv;
  s.RKs = h->RKs;
  s.RKe = h->RKe;
  return s;
}

// piecewise-linear time curve, shared by Normal_BCs (mat:1707-1722) and gravity_force (mat:2688-2704)
double tcurve(const spsph_params &p, int it_curves, double t) {
  if (it_curves < 1 || it_curves > p.ntcurves) return 0.0;
  double tt0 = 0, tt1 = 0;
  int ipts = 1;
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

// per-step scalars: gravity factor and the value of every prescribed BC at t_actual
void step_scalars(spsph_handle *h, int itimestep, double time_sph, double dt) {
  const spsph_params &p = h->hp;
  DevParams &P = h->P;
  P.itimestep = itimestep;
  P.time_sph = time_sph;
  P.dt = dt;
  P.grav[0] = P.grav[1] = 0.0;
  if (p.ic_grav != 0) {
    double factg = tcurve(p, p.tcurve_grav, time_sph);
    factg = factg * p.ft_grav;
    if (p.ic_grav == 1) {
      P.grav[0] = factg * p.cgrav[0];
      P.grav[1] = factg * p.cgrav[1];
    }
  }
  const double ic_time = 0.0;  // uninitialised in the reference: zero reading
  const double t_actual = time_sph + ic_time * dt;
  for (int b = 0; b < p.no_bcs; ++b) {
    const double *bl = p.bc_list[b];
    const int it_curves = (int)bl[2];
    double v = 0.0;
    if (it_curves == 0) {
      const double a0 = bl[4], a1 = bl[3], w = bl[5], phi = bl[6], tt = bl[7];
      const double fact = 1.0 - std::exp(-t_actual / tt);
      const double argum = w * t_actual - phi;
      v = (a0 + a1 * std::sin(argum)) * fact;
    } else if (it_curves > 0) {
      v = tcurve(p, it_curves, t_actual);
      v = v * bl[3];
    }
    P.bcval[b] = v;
    P.bcvar[b] = (int)bl[1];
  }
}

// (re)allocation of one 4-byte list array: on failure the old buffer is gone, the pointer is null and the caller
// zeroes the capacity, so that nothing can write through a stale pointer or free it twice
template <class T>
static cudaError_t realloc4(T **p, long long n, bool wanted = true) {
  cudaFree(*p);
  *p = nullptr;
  if (!wanted) return cudaSuccess;
  const cudaError_t e = cudaMalloc((void **)p, (size_t)n * 4);
  // the sweeps copy whole groups of list rows, including slots past the end of a lane's list that nobody writes:
  // zero them once so that those (unused) reads are reads of initialised memory (compute-sanitizer initcheck)
  return e != cudaSuccess ? e : cudaMemset(*p, 0, (size_t)n * 4);
}
int ensure_lists(spsph_handle *h, long long t0, long long tC, long long tD) {
  // 1024 entries of slack: the sweeps stream whole groups of rows and may read (never use) up to 7 rows past a slice
  auto need_cap = [](long long need, long long cap) { return need + 1024 > cap ? need + need / 8 + 2048 : 0ll; };
  if (const long long c = need_cap(t0, h->cap0)) {
    h->cap0 = 0;
    // (m/rho)_partner * w is only stored when it is not a per-species constant times w
    CUDA_TRY(realloc4(&h->L.h0lo, c, !h->umor));
    CUDA_TRY(realloc4(&h->L.h0hi, c, !h->umor));
    CUDA_TRY(realloc4(&h->L.idx0, c));
    CUDA_TRY(realloc4(&h->L.w0, c));
    CUDA_TRY(realloc4(&h->L.gx0, c));
    CUDA_TRY(realloc4(&h->L.gy0, c));
    h->cap0 = c;
  }
  if (const long long c = need_cap(tC, h->capC)) {
    h->capC = 0;
    CUDA_TRY(realloc4(&h->L.xC, c));
    CUDA_TRY(realloc4(&h->L.yC, c));
    CUDA_TRY(realloc4(&h->L.hC, c, !h->uniform_h));  // else a constant
    CUDA_TRY(realloc4(&h->L.idxC, c));
    CUDA_TRY(realloc4(&h->L.wC, c));
    CUDA_TRY(realloc4(&h->L.gxC, c));
    CUDA_TRY(realloc4(&h->L.gyC, c));
    h->capC = c;
  }
  if (const long long c = need_cap(tD, h->capD)) {
    h->capD = 0;
    CUDA_TRY(realloc4(&h->L.idxD, c));
    CUDA_TRY(realloc4(&h->L.wD, c));
    h->capD = c;
  }
  return 0;
}

// exclusive scan of `rows` rows of int32 (stride elements apart) over n = *n_ptr + n_add elements each
void launch_scan(spsph_handle *h, const int *in, int *out, int rows, int stride, const int *n_ptr, int n_add,
                 long long *totals, int kid) {
#ifndef SPSPH_EMU_SERIAL
  dim3 g(SCAN_BLOCKS, rows);
  k_scan_reduce<<<g, SCAN_THREADS, 0, h->stream>>>(in, stride, n_ptr, n_add, h->scan_bsum);
  k_scan_sums<<<rows, SCAN_THREADS, 0, h->stream>>>(h->scan_bsum, totals);
  k_scan_apply<<<g, SCAN_THREADS, 0, h->stream>>>(in, out, stride, n_ptr, n_add, h->scan_bsum);
  mark(h, kid, 3);
#else
  const int n = (n_ptr ? *n_ptr : 0) + n_add;
  for (int r = 0; r < rows; ++r) {
    long long acc = 0;
    for (int i = 0; i < n; ++i) {
      const int v = in[(size_t)r * stride + i];
      out[(size_t)r * stride + i] = (int)acc;
      acc += v;
    }
    if (totals) totals[r] = acc;
  }
  mark(h, kid, 3);
#endif
}

// particles a per-particle kernel visits, and the grid for it
static LocalList local_list(const spsph_handle *h, int nfull) {
  if (!h->dist) return LocalList{nullptr, nullptr, nfull};
  return LocalList{h->list_ids[h->list_cur], h->list_n + h->list_cur, nfull};
}
static int list_grid(const spsph_handle *h, int nfull, int tb) { return h->dist ? 148 * 8 : (nfull + tb - 1) / tb; }
void launch_scan(spsph_handle *h, const int *in, int *out, int rows, int stride, const int *n_ptr, int n_add,
                 long long *totals, int kid);
// order-preserving compaction of the local list (ids_in == nullptr: build from the flags of all particles)
static int compact_local_list(spsph_handle *h, const int *ids_in, const int *n_in, int *ids_out, int *n_out) {
  const int n2 = h->P.ntotal2;
  cudaStream_t s = h->stream;
  const int g = ids_in ? 148 * 8 : (n2 + 255) / 256;
  k_list_flags<<<g, 256, 0, s>>>(ids_in, n_in, n2, h->lflag, h->list_keep);
  launch_scan(h, h->list_keep, h->list_pos, 1, n2, ids_in ? n_in : nullptr, ids_in ? 0 : n2, h->scan_totals + 7,
              KID_HALO);
  k_list_scatter<<<g, 256, 0, s>>>(ids_in, n_in, n2, h->list_keep, h->list_pos, ids_out, n_out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
static int rebuild_local_list(spsph_handle *h) {
  h->list_cur = 0;
  h->nloc_valid = false;
  h->halo_prev_valid = false;
  return compact_local_list(h, nullptr, nullptr, h->list_ids[0], h->list_n);
}

// per-step halo exchange + migration (dist_kernels.cuh); everything is enqueued on the engine stream
// Size of a halo message: the full buffer until the previous step's record count is known, then that count plus
// 12.5 % + 4096 records (counts drift by a few particles per step; an overflow raises the error flag).
static int halo_limit(const spsph_handle *h, int prev_count) {
  if (!h->halo_prev_valid) return h->D.cap;
  const long long l = (long long)prev_count + prev_count / 8 + h->halo_slack;
  return l < h->D.cap ? (int)l : h->D.cap;
}
int halo_exchange(spsph_handle *h) {
  const DevParams &P = h->P;
  h->D.lim[0] = halo_limit(h, h->halo_prev_send[0]);
  h->D.lim[1] = halo_limit(h, h->halo_prev_send[1]);
  const DistGeom &D = h->D;
  cudaStream_t s = h->stream;
  const int n2 = P.ntotal2;
  const StatePtrs st = state_ptrs(h, h->cur);
  const HaloArrays A{h->x,   h->epsp, h->fdp,  h->x_10, h->disp_10,   h->displ,    h->if_out,
                     h->rho, h->hsml, h->mor,  h->divu, h->mrho,      h->bc_or_not, h->fs_normal};
  if (h->profiling) mark(h, -1, 0);
  CUDA_TRY(cudaMemsetAsync(h->halo_cnt, 0, 2 * sizeof(int), s));
  const LocalList LL = local_list(h, n2);
  k_halo_select<<<list_grid(h, n2, 256), 256, 0, s>>>(P, D, h->x, h->lflag, LL, h->halo_cnt, h->halo_ids[0],
                                                       h->halo_ids[1], h->halo_cnt + 2);
  for (int side = 0; side < 2; ++side)
    k_halo_pack<<<148, 256, 0, s>>>(P, D, st, A, h->halo_cnt + side, D.lim[side], h->halo_ids[side], h->halo_send[side]);
  const int left = D.rank - 1, right = D.rank + 1;
  const size_t msg_s[2] = {(size_t)D.rec * (D.lim[0] + 1), (size_t)D.rec * (D.lim[1] + 1)};
  const size_t msg_r[2] = {(size_t)D.rec * (halo_limit(h, h->halo_prev_recv[0]) + 1),
                           (size_t)D.rec * (halo_limit(h, h->halo_prev_recv[1]) + 1)};
  NCCL_TRY(h->p_ncclGroupStart());
  if (left >= 0) {
    NCCL_TRY(h->p_ncclSend(h->halo_send[0], msg_s[0], ncclDouble, left, h->comm, s));
    NCCL_TRY(h->p_ncclRecv(h->halo_recv[0], msg_r[0], ncclDouble, left, h->comm, s));
  }
  if (right < D.nranks) {
    NCCL_TRY(h->p_ncclSend(h->halo_send[1], msg_s[1], ncclDouble, right, h->comm, s));
    NCCL_TRY(h->p_ncclRecv(h->halo_recv[1], msg_r[1], ncclDouble, right, h->comm, s));
  }
  NCCL_TRY(h->p_ncclGroupEnd());
  for (int side = 0; side < 2; ++side) {
    const int peer = side == 0 ? left : right;
    if (peer < 0 || peer >= D.nranks) continue;
    k_halo_unpack<<<148, 256, 0, s>>>(P, D, st, A, h->halo_recv[side], h->lflag, h->list_ids[h->list_cur],
                                      h->list_n + h->list_cur);
  }
  for (int side = 0; side < 2; ++side) {
    const int peer = side == 0 ? left : right;
    if (peer < 0 || peer >= D.nranks) continue;
    k_halo_own<<<148, 256, 0, s>>>(P, D, h->x, h->halo_recv[side], h->lflag);
  }
  // drop the ghosts nobody refreshed (they left our halo) from the local list
  if (compact_local_list(h, h->list_ids[h->list_cur], h->list_n + h->list_cur, h->list_ids[1 - h->list_cur],
                         h->list_n + (1 - h->list_cur)))
    return 1;
  h->list_cur = 1 - h->list_cur;
  mark(h, KID_HALO, 9);
  return 0;
}

// neighbour search up to and including the list fill; leaves the pair totals in h->status_h
#ifdef SPSPH_HOST_EMU
// host emulation, lockstep mode: the test harness hands over this step's sorted arrays and gather lists (built from
// the oracle's pair list in the layout k_count / k_fill write), see tests/test_step_emulation_cpu.py
int emu_import_lists(spsph_handle *h) {
  const spsph_emu_lists *E = h->emu_lists;
  h->emu_lists = nullptr;
  const DevParams &P = h->P;
  const size_t n2 = (size_t)P.ntotal2;
  const int cnt[3] = {P.nnode, P.nstress, P.ndummy};
  for (int sp = 0; sp < 3; ++sp) {
    std::memcpy(h->order + sp * n2, E->order[sp], cnt[sp] * sizeof(int));
    for (int k = 0; k < cnt[sp]; ++k) {
      const int i = E->order[sp][k];
      h->ordc[sp * n2 + k] = h->umor ? (i | ((int)h->mcls[i] << 30)) : i;
    }
    std::memcpy(h->scell + sp * n2, E->cell[sp], cnt[sp] * sizeof(int));
    std::memcpy(h->spos + sp * n2, E->pos[sp], cnt[sp] * sizeof(double2));
    std::memcpy(h->sh + sp * n2, E->h[sp], cnt[sp] * sizeof(double));
  }
  std::memcpy(h->pos_of, E->pos_of, n2 * sizeof(int));
  for (int k = 0; k < 3; ++k) h->nloc[k] = cnt[k];
  h->nloc_valid = true;
  if (ensure_lists(h, E->tot0, E->totC, E->totD)) return 1;
  const int T = h->M.nnp + h->M.nsp;
  std::memcpy(h->n0, E->n0, T * sizeof(int));
  std::memcpy(h->n1, E->n1, T * sizeof(int));
  std::memcpy(h->oslice, E->off0, h->nslices * sizeof(int));
  std::memcpy(h->oslice + h->nslices, E->offC, h->nslices * sizeof(int));
  std::memcpy(h->oslice + 2 * h->nslices, E->offD, h->nslices * sizeof(int));
  h->L.off0 = h->oslice;
  h->L.offC = h->oslice + h->nslices;
  h->L.offD = h->oslice + 2 * h->nslices;
  for (long long a = 0; a < E->tot0; ++a) {
    const int q = E->idx0[a];
    h->L.idx0[a] = h->umor ? (q | ((int)h->mcls[q] << 30)) : q;
    h->L.w0[a] = E->w0[a];
    h->L.gx0[a] = E->gx0[a];
    h->L.gy0[a] = E->gy0[a];
    if (h->L.h0lo) {  // (m/rho)_partner * w, zero for wall partners (k_fill)
      const double h0 = q >= P.ntotal ? 0.0 : h->mor[q] * (double)E->w0[a];
      h->L.h0lo[a] = __double2loint(h0);
      h->L.h0hi[a] = __double2hiint(h0);
    }
  }
  for (long long a = 0; a < E->totC; ++a) {
    h->L.idxC[a] = E->idxC[a];
    h->L.wC[a] = E->wC[a];
    h->L.gxC[a] = E->gxC[a];
    h->L.gyC[a] = E->gyC[a];
    h->L.xC[a] = E->xC[a];
    h->L.yC[a] = E->yC[a];
    if (h->L.hC) h->L.hC[a] = E->hC[a];
  }
  for (long long a = 0; a < E->totD; ++a) {
    h->L.idxD[a] = E->idxD[a];
    h->L.wD[a] = E->wD[a];
  }
  std::memcpy(h->bc_int, E->bc_int, P.nnode * sizeof(int));
  std::memcpy(h->if_out, E->if_out, n2 * sizeof(int));
  if (P.track_nint) std::memcpy(h->n_int, E->n_int, P.nnode * sizeof(float));
  h->last_m_before = h->m_pairs;
  h->last_n_pairs = E->n_pairs;
  GrowthRule gr{E->growth_mode, 0, E->growth_ka, E->growth_kb};
  *h->growth = gr;
  if (E->n_pairs > h->m_pairs) h->m_pairs = E->n_pairs;
  return 0;
}

// host emulation, stand-alone mode: the cooperative pieces of the neighbour build done by plain host loops (the
// per-particle kernels k_cell_id, k_scatter, k_rank, k_count, k_fill run thread by thread like all the others)
void emu_bbox(spsph_handle *h) {  // k_domain_bbox + k_bbox_final: Check_Out_Domain, bounds and max h of the in-domain particles
  const DevParams &P = h->P;
  double xmn = 1.e+10, ymn = 1.e+10, xmx = -1.e+10, ymx = -1.e+10, hmx = 0.0, hmn = 1.e+300;
  for (int i = 0; i < P.ntotal2; ++i) {
    const int lf = h->dist ? h->lflag[i] : 1;  // slab run: remote particles are skipped, only owned ones enter the bounds
    if (lf == 0) continue;
    const double px = h->x[2 * (size_t)i], py = h->x[2 * (size_t)i + 1];
    const double dxx = (px - P.xmin_dom[0]) * (px - P.xmax_dom[0]);
    const double dyy = (py - P.xmin_dom[1]) * (py - P.xmax_dom[1]);
    if (dxx > 0.0 || dyy > 0.0) h->if_out[i] = 1;
    if (h->if_out[i] || lf != 1) continue;
    xmn = std::fmin(xmn, px);
    xmx = std::fmax(xmx, px);
    ymn = std::fmin(ymn, py);
    ymx = std::fmax(ymx, py);
    hmx = std::fmax(hmx, h->hsml[i]);
    hmn = std::fmin(hmn, h->hsml[i]);
  }
  const double bb[6] = {-xmn, -ymn, xmx, ymx, hmx, -hmn};  // the layout k_bbox_final writes (max-reducible)
  std::memcpy(h->bb6, bb, sizeof(bb));
}
void emu_slice_widths(spsph_handle *h) {  // the warp-wide maxima at the end of k_count
  for (int sl = 0; sl < h->nslices; ++sl) {
    int m0 = 0, m1 = 0;
    for (int l = 0; l < SLICE; ++l) {
      m0 = std::max(m0, h->n0[sl * SLICE + l]);
      m1 = std::max(m1, h->n1[sl * SLICE + l]);
    }
    const bool is_node = sl * SLICE < h->M.nnp;
    h->wslice[sl] = m0 * SLICE;
    h->wslice[h->nslices + sl] = is_node ? m1 * SLICE : 0;
    h->wslice[2 * h->nslices + sl] = is_node ? 0 : m1 * SLICE;
  }
}
#endif

// cell grid + counting sort of the local particles (grid_find_NEW Tasks 1-2, main:1245-1305); shared by both paths
int sort_particles(spsph_handle *h) {
  const DevParams &P = h->P;
  const int n2 = P.ntotal2;
  const int TB = 256;
  cudaStream_t s = h->stream;
  if (h->profiling) {  // anchor event so that the first kernel's interval is well defined
    mark(h, -1, 0);
  }
  const int *lflag = h->dist ? h->lflag : nullptr;
  const LocalList LL = local_list(h, n2);
  const int GL = list_grid(h, n2, TB);
#ifndef SPSPH_EMU_SERIAL
  k_domain_bbox<<<h->bbox_blocks, TB, 0, s>>>(P, h->x, h->hsml, h->if_out, lflag, LL, h->bbox_partial);
  mark(h, KID_BBOX);
  k_bbox_final<<<1, 32, 0, s>>>(h->bbox_blocks, h->bbox_partial, h->bb6);
#else
  emu_bbox(h);
#endif
  if (h->dist)  // global grid bounds: the reference's cell grid (hence its pair order) is a global property
    NCCL_TRY(h->p_ncclAllReduce(h->bb6, h->bb6, 6, ncclDouble, ncclMax, h->comm, s));
  k_grid_params<<<1, 32, 0, s>>>(h->bb6, h->G, h->cell_capacity);
  mark(h, KID_GRID, 2);
  k_zero_cells<<<296, TB, 0, s>>>(h->G, h->cell_cnt, h->cell_stride);
  k_zero_cells<<<296, TB, 0, s>>>(h->G, h->cell_fill, h->cell_stride);
  CUDA_TRY(cudaMemsetAsync(h->nout, 0, 6 * sizeof(int), s));
  mark(h, KID_ZERO, 2);
  k_cell_id<<<GL, TB, 0, s>>>(P, h->G, h->x, h->if_out, LL, h->which_cell, h->cell_cnt, h->cell_stride, h->nout);
  mark(h, KID_CELLID);
  launch_scan(h, h->cell_cnt, h->cell_start, 3, h->cell_stride, &h->G->ncell, 1, h->scan_totals + 4, KID_SCAN);
  k_scatter<<<GL, TB, 0, s>>>(P, LL, h->which_cell, h->cell_start, h->cell_fill, h->cell_stride, h->tmp_ids);
  mark(h, KID_SCATTER);
  k_rank<<<GL, TB, 0, s>>>(P, LL, h->G, h->x, h->hsml, h->which_cell, h->cell_start, h->cell_stride, h->tmp_ids,
                           h->order, h->spos, h->sh, h->scell, h->pos_of, h->supos, h->nout,
                           h->umor ? h->mcls : nullptr, h->ordc);
  if (h->tile_cfg)  // species-sorted copies of the per-particle constants the tile kernels stage
    k_rank_consts<<<GL, TB, 0, s>>>(P, LL, h->pos_of, h->mass, h->rho, h->mor, h->smor, h->smrho, h->srrho);
  mark(h, KID_RANK, h->tile_cfg ? 2 : 1);
  return 0;
}

// slots of one species that can hold local particles: all of them, or (slab) last step's local count plus what two
// halo messages can add
static void slot_bounds(const spsph_handle *h, int (&bound)[3]) {
  bound[0] = h->M.nnp;
  bound[1] = h->M.nsp;
  bound[2] = h->M.ndp;
  if (h->dist && h->nloc_valid)
    for (int k = 0; k < 3; ++k) {
      const long long b = (long long)h->nloc[k] + halo_limit(h, h->halo_prev_recv[0]) +
                          halo_limit(h, h->halo_prev_recv[1]) + 32;
      if (b < bound[k]) bound[k] = (int)b;
    }
}

// slab runs: the read-back that follows the neighbour build also brings the halo counts (they size the next step's
// messages) and the error flags of the exchange
static int dist_status_enqueue(spsph_handle *h) {
  cudaStream_t s = h->stream;
  h->dist_h[0] = h->dist_h[1] = 0.0;
  if (h->D.rank > 0) CUDA_TRY(cudaMemcpyAsync(h->dist_h, h->halo_recv[0], sizeof(double), cudaMemcpyDeviceToHost, s));
  if (h->D.rank < h->D.nranks - 1)
    CUDA_TRY(cudaMemcpyAsync(h->dist_h + 1, h->halo_recv[1], sizeof(double), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(h->dist_h + 2, h->halo_cnt, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
  return 0;
}
static int dist_status_apply(spsph_handle *h) {
  const int *hc = reinterpret_cast<const int *>(h->dist_h + 2);  // halo_cnt[0..1], error flags [2..3]
  if (hc[2]) {
    h->err = "multi-GPU: halo message capacity exceeded (too many particles near a slab boundary)";
    return 1;
  }
  if (hc[3]) {
    h->err = "multi-GPU: the distributed list-growth search failed in the previous step (grid too large?)";
    return 1;
  }
  for (int side = 0; side < 2; ++side) {
    h->halo_prev_send[side] = hc[side];
    h->halo_prev_recv[side] = (int)h->dist_h[side];
  }
  // After spsph_dist_set_planes the counts of two exchanges are no guide for the next one: the first carries the
  // particles that change slab on top of the halo band, and the rank that gains them sends a band that is short by
  // the moved strip -- its band of the following step is larger by up to 50 %, beyond the 12.5 % the limit allows.
  if (h->halo_full_msgs > 0) --h->halo_full_msgs;
  h->halo_prev_valid = h->halo_full_msgs == 0;
  return 0;
}

// Neighbour lists of the list path: count pass, list sizing (one host round trip), growth rule, fill pass.
// forced_mode >= 0: materialise the id lists of a step that ran on the tile path (free-surface detection at
// download, spsph_pairs) with that traversal order; the step bookkeeping is left alone.
int build_lists(spsph_handle *h, int forced_mode = -1) {
#ifdef SPSPH_HOST_EMU
  if (h->emu_lists) return emu_import_lists(h);
#endif
  const DevParams &P = h->P;
  const int n2 = P.ntotal2;
  cudaStream_t s = h->stream;
  const int *lflag = h->dist ? h->lflag : nullptr;
  CUDA_TRY(cudaMemsetAsync(h->nfwd_u, 0, (size_t)n2 * sizeof(int), s));
  CUDA_TRY(cudaMemsetAsync(h->wslice, 0, 3 * (size_t)h->nslices * sizeof(int), s));
  CUDA_TRY(cudaMemsetAsync(h->cand_overflow, 0, sizeof(int), s));
  const SortArrays S = sort_arrays(h);
  int bound[3];
  slot_bounds(h, bound);
  if (!h->dist) {  // single GPU: every slot is live, one launch over all of them
    const int tn = h->M.total();
    k_count<<<(tn + 127) / 128, 128, 0, s>>>(P, h->M, h->G, S, h->n0, h->n1, h->nfwd_u, h->nall, h->wslice,
                                             h->wslice + h->nslices, h->wslice + 2 * h->nslices, lflag, h->cand0,
                                             h->cand1, h->cand_overflow, h->nout, 0, tn);
  } else {  // slab: one launch per species over the leading slots that can hold local particles
    const int t0s[3] = {0, h->M.nnp, h->M.nnp + h->M.nsp};
    for (int k = 0; k < 3; ++k) {
      const int tn = (bound[k] + 31) & ~31;
      if (tn > 0)
        k_count<<<(tn + 127) / 128, 128, 0, s>>>(P, h->M, h->G, S, h->n0, h->n1, h->nfwd_u, h->nall, h->wslice,
                                                 h->wslice + h->nslices, h->wslice + 2 * h->nslices, lflag,
                                                 h->cand0, h->cand1, h->cand_overflow, h->nout, t0s[k], tn);
    }
  }
  mark(h, KID_COUNT, h->dist ? 3 : 1);
#ifdef SPSPH_EMU_SERIAL
  emu_slice_widths(h);
#endif
  launch_scan(h, h->wslice, h->oslice, 3, h->nslices, nullptr, h->nslices, h->scan_totals, KID_SCAN);
  if (h->dist)  // unified slots of local particles are the leading ones: scan only those
    launch_scan(h, h->nfwd_u, h->base_u, 1, n2, h->list_n + h->list_cur, 0, h->scan_totals + 3, KID_SCAN);
  else
    launch_scan(h, h->nfwd_u, h->base_u, 1, n2, nullptr, n2, h->scan_totals + 3, KID_SCAN);
  if (h->dist && forced_mode < 0) {  // every pair is counted once, at the owner of its earlier member
    CUDA_TRY(cudaMemcpyAsync(h->scan_totals + 6, h->scan_totals + 3, sizeof(long long), cudaMemcpyDeviceToDevice, s));
    NCCL_TRY(h->p_ncclAllReduce(h->scan_totals + 3, h->scan_totals + 3, 1, ncclInt64, ncclSum, h->comm, s));
  }
  k_status<<<1, 32, 0, s>>>(h->G, h->scan_totals, h->cell_start, h->cell_stride, h->nout, h->cand_overflow,
                            h->status_d);
  mark(h, KID_STATUS);
  CUDA_TRY(cudaMemcpyAsync(h->status_h, h->status_d, sizeof(StepStatus), cudaMemcpyDeviceToHost, s));
  if (h->dist && forced_mode < 0 && dist_status_enqueue(h)) return 1;
  CUDA_TRY(cudaStreamSynchronize(s));
  const StepStatus st = *h->status_h;
  if (h->dist && forced_mode < 0 && dist_status_apply(h)) return 1;
  if (st.overflow) {
    h->err = "cell grid larger than the capacity derived from Xmin_Domain/Xmax_Domain";
    return 1;
  }
  if (st.tot0 >= (1ll << 31) || st.totC >= (1ll << 31) || st.totD >= (1ll << 31)) {
    h->err = "neighbour list exceeds 2^31 entries on one device";
    return 1;
  }
  for (int k = 0; k < 3; ++k) h->nloc[k] = st.nloc[k];
  h->nloc_valid = true;
  if (ensure_lists(h, st.tot0, st.totC, st.totD)) return 1;
  h->L.off0 = h->oslice;
  h->L.offC = h->oslice + h->nslices;
  h->L.offD = h->oslice + 2 * h->nslices;
  // list-growth rule
  GrowthRule gr{0, 0, 0};
  if (forced_mode >= 0) {
    gr.mode = forced_mode;
  } else {
    h->last_m_before = h->m_pairs;
    h->last_n_pairs = st.n_pairs;
    if (st.n_pairs > h->m_pairs) gr.mode = (h->m_pairs == 0) ? 1 : 2;
  }
  CUDA_TRY(cudaMemcpyAsync(h->growth, &gr, sizeof(gr), cudaMemcpyHostToDevice, s));
  if (h->profiling) mark(h, -1, 0);  // do not charge the host round trip to the next kernel
  if (gr.mode == 2 && h->dist) {
    // distributed search for the pair with global creation index m_pairs (dist_kernels.cuh); no host sync
    const long long *ltot = h->scan_totals + 6;  // local pair total saved before the all-reduce
    k_gt_rows<<<GT_CAP / 256, 256, 0, s>>>(h->G, S, h->base_u, ltot, h->gt_buf);
    NCCL_TRY(h->p_ncclAllReduce(h->gt_buf, h->gt_buf, GT_CAP, ncclInt64, ncclSum, h->comm, s));
    k_gt_pick_row<<<1, 32, 0, s>>>(h->G, h->gt_buf, h->m_pairs, h->gt_sel);
    k_gt_cells<<<GT_CAP / 256, 256, 0, s>>>(h->G, S, h->base_u, ltot, h->gt_sel, h->gt_buf);
    NCCL_TRY(h->p_ncclAllReduce(h->gt_buf, h->gt_buf, GT_CAP, ncclInt64, ncclSum, h->comm, s));
    k_gt_pick_cell<<<1, 32, 0, s>>>(h->G, S, h->nfwd_u, h->gt_buf, h->gt_sel, h->gt_mine);
    NCCL_TRY(h->p_ncclAllGather(h->gt_mine, h->gt_all, 2 * GT_PCAP, ncclInt64, h->comm, s));
    k_gt_pick_particle<<<1, 32, 0, s>>>(P, h->G, S, h->pos_of, h->gt_all, h->D.nranks, h->D.rank, h->gt_sel,
                                        h->gt_out2);
    NCCL_TRY(h->p_ncclAllReduce(h->gt_out2, h->gt_out2, 3, ncclInt64, ncclSum, h->comm, s));
    k_gt_finish<<<1, 32, 0, s>>>(h->gt_out2, h->gt_sel, h->growth, h->halo_cnt + 3);
    mark(h, KID_THRESH, 6);
  } else if (gr.mode == 2) {
    k_growth_threshold<<<1, 32, 0, s>>>(P, h->M, h->G, S, h->base_u, h->m_pairs, h->growth);
    mark(h, KID_THRESH);
  }
  if (forced_mode < 0 && st.n_pairs > h->m_pairs) h->m_pairs = st.n_pairs;
  SlotMap ML = h->M;  // only the slots of this rank's local particles hold data
  ML.nn = h->nloc[0];
  ML.ns = h->nloc[1];
  ML.nd = h->nloc[2];
  // single GPU: one launch over all list-owning slots; slab: one launch per species over its local slots
  const int nfl = h->dist ? 2 : 1;
  const int ft0[2] = {0, ML.nnp}, ftn[2] = {h->dist ? ((ML.nn + 31) & ~31) : ML.nnp + ML.nsp, (ML.ns + 31) & ~31};
  for (int k = 0; k < nfl; ++k) {
    if (ftn[k] == 0) continue;
    if (st.pad[0] || h->force_fill_scan)  // a particle has more partners than the candidate scratch holds: search again while filling
      k_fill_scan<<<(ftn[k] + 127) / 128, 128, 0, s>>>(P, ML, h->G, S, h->n0, h->n1, h->growth, h->L, h->bc_int,
                                                       h->n_int, h->mor, h->umor ? h->mcls : nullptr, ft0[k], ftn[k]);
    else if (h->uniform_cubic)
      k_fill<true><<<(ftn[k] + 127) / 128, 128, 0, s>>>(P, ML, h->G, S, h->n0, h->n1, h->growth, h->L, h->bc_int,
                                                        h->n_int, h->mor, h->umor ? h->mcls : nullptr, h->cand0, h->cand1, ft0[k],
                                                        ftn[k]);
    else
      k_fill<false><<<(ftn[k] + 127) / 128, 128, 0, s>>>(P, ML, h->G, S, h->n0, h->n1, h->growth, h->L, h->bc_int,
                                                         h->n_int, h->mor, h->umor ? h->mcls : nullptr, h->cand0, h->cand1, ft0[k],
                                                        ftn[k]);
  }
  mark(h, KID_FILL, h->dist ? 2 : 1);
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// Cell-tile path (tile_kernels.cuh)
// ------------------------------------------------------------------------------------------------------
// Which inputs the tile kernels cover; everything else runs on the list path of round 1 (same results).
static bool tile_config_ok(const spsph_handle *h) {
  const spsph_params &p = h->hp;
  return p.sp_sph && !p.cont_density && !p.art_stress && !h->fs_each_step && h->uniform_cubic &&
         (long long)h->M.nnp + h->M.nsp + h->M.ndp < (1ll << 30);
}

static int tile_alloc_weights(spsph_handle *h) {
  TileLists &T = h->TL;
  const long long g0 = ((long long)T.nsl_n * (T.capN0 >> 2) + (long long)(h->M.nsp / SLICE) * (T.capS0 >> 2)) * 32 + 64;
  const long long gc = (long long)T.nsl_n * (T.capC >> 2) * 32 + 64;
  const long long gs = ((long long)T.nsl_n * (T.capC >> 2) + (long long)(h->M.nsp / SLICE) * (T.capD >> 2)) * 32 + 64;
  if (g0 > h->tile_w0_cap) {
    cudaFree(T.w0);
    cudaFree(T.gx0);
    cudaFree(T.gy0);
    cudaFree(T.code0);
    T.w0 = T.gx0 = T.gy0 = nullptr;
    T.code0 = nullptr;
    h->tile_w0_cap = 0;
    CUDA_TRY(cudaMalloc((void **)&T.code0, (size_t)g0 * sizeof(unsigned)));
    CUDA_TRY(cudaMalloc((void **)&T.w0, (size_t)g0 * sizeof(float4)));
    CUDA_TRY(cudaMalloc((void **)&T.gx0, (size_t)g0 * sizeof(float4)));
    CUDA_TRY(cudaMalloc((void **)&T.gy0, (size_t)g0 * sizeof(float4)));
    h->tile_w0_cap = g0;
  }
  if (gc > h->tile_c_cap) {
    cudaFree(T.gxC);
    cudaFree(T.gyC);
    T.gxC = T.gyC = nullptr;
    h->tile_c_cap = 0;
    CUDA_TRY(cudaMalloc((void **)&T.gxC, (size_t)gc * sizeof(float4)));
    CUDA_TRY(cudaMalloc((void **)&T.gyC, (size_t)gc * sizeof(float4)));
    h->tile_c_cap = gc;
  }
  if (gs > h->tile_s_cap) {
    cudaFree(T.codeS);
    T.codeS = nullptr;
    h->tile_s_cap = 0;
    CUDA_TRY(cudaMalloc((void **)&T.codeS, (size_t)gs * sizeof(unsigned)));
    h->tile_s_cap = gs;
  }
  return 0;
}

// one-time allocations of the tile path (sizes depend on the particle counts only)
static int tile_setup(spsph_handle *h) {
  if (h->TL.rowA) return 0;
  const size_t nsl = (size_t)h->M.nnp + h->M.nsp, n2 = (size_t)h->P.ntotal2;
  TileLists &T = h->TL;
  T.nslots = (int)nsl;
  T.nsl_n = h->M.nnp / SLICE;
  T.capN0 = 48;
  T.capS0 = 28;
  T.capC = 28;
  T.capD = 44;
  T.n0 = h->n0;
  T.n1 = h->n1;
  int rc = dalloc(h, &T.rowA, 3 * nsl) | dalloc(h, &T.rowS, 3 * nsl) | dalloc(h, &T.mW, 3 * nsl);
  rc |= dalloc(h, &h->GT.g[0], (size_t)h->M.nsp / TS_T + 2) | dalloc(h, &h->GT.g[1], (size_t)h->M.nnp / TN_T + 2);
  rc |= dalloc(h, &h->GT.g[2], (size_t)h->M.nnp / TS_T + 2) | dalloc(h, &h->GT.g[3], (size_t)h->M.nsp / TS_T + 2);
  rc |= dalloc(h, &h->GT.g[4], (size_t)h->M.nnp / TN_T + 2);
  rc |= dalloc(h, &h->smor, 2 * n2) | dalloc(h, &h->srrho, 2 * n2) | dalloc(h, &h->smrho, 2 * n2);
  rc |= dalloc(h, &h->TR.NAs, (size_t)h->M.nnp) | dalloc(h, &h->TR.NBs, (size_t)h->M.nnp);
  rc |= dalloc(h, &h->TR.SAs, (size_t)h->M.nsp) | dalloc(h, &h->TR.SBs, (size_t)h->M.nsp) | dalloc(h, &h->TR.SVs, (size_t)h->M.nsp);
  rc |= dalloc(h, &h->tile_acc, 2) | dalloc(h, &h->tile_flags, 8) | dalloc(h, &h->tstat_d, 1);
  if (rc) return 1;
  CUDA_TRY(cudaMallocHost((void **)&h->tstat_h, sizeof(TileStatus)));
  return tile_alloc_weights(h);
}

static SortedConsts sorted_consts(const spsph_handle *h) {
  SortedConsts C;
  const size_t n2 = (size_t)h->P.ntotal2;
  for (int sp = 0; sp < 2; ++sp) {
    C.mor[sp] = h->smor + sp * n2;
    C.mrho[sp] = h->smrho + sp * n2;
    C.rrho[sp] = h->srrho + sp * n2;
  }
  return C;
}

// One-pass neighbour build of the tile path. *use = false: this step has to take the list path (the pair list grew
// after the first step -> split traversal order, or a stencil row outgrew the masks).
int tile_build(spsph_handle *h, bool *use) {
  *use = false;
  const DevParams &P = h->P;
  cudaStream_t s = h->stream;
  const int *lflag = h->dist ? h->lflag : nullptr;
  const SortArrays S = sort_arrays(h);
  const SortedConsts C = sorted_consts(h);
  const int rev = h->m_pairs == 0 ? 1 : 0;  // first step after an upload: every list node is new (main:1362-1368)
  const int want_c = (P.alpha > 0 || P.beta > 0) ? 1 : 0;  // velocity-velocity gradients: artificial viscosity
  const int want_s = (P.update_x && P.xsph) ? 1 : 0;       // same-species entries: XSPH
  TileStatus st{};
  for (int attempt = 0; attempt < 2; ++attempt) {
    CUDA_TRY(cudaMemsetAsync(h->tile_acc, 0, 2 * sizeof(u64), s));
    CUDA_TRY(cudaMemsetAsync(h->tile_flags, 0, 8 * sizeof(int), s));
    int bound[3];
    slot_bounds(h, bound);
    if (bound[0] > 0)
      k_tile_build<SP_NODE><<<(bound[0] + TB_T - 1) / TB_T, TB_T, 0, s>>>(
          P, h->G, S, h->TL, C, h->M.nnp, rev, want_c, want_s, lflag, h->nout, h->nall, h->bc_int, h->n_int, h->norm, h->AE,
          h->tile_acc, h->tile_flags);
    if (bound[1] > 0)
      k_tile_build<SP_STRESS><<<(bound[1] + TB_T - 1) / TB_T, TB_T, 0, s>>>(
          P, h->G, S, h->TL, C, h->M.nnp, rev, want_c, want_s, lflag, h->nout, h->nall, h->bc_int, h->n_int, h->norm, h->AE,
          h->tile_acc, h->tile_flags);
    if (bound[2] > 0)
      k_tile_build<SP_DUMMY><<<(bound[2] + TB_T - 1) / TB_T, TB_T, 0, s>>>(
          P, h->G, S, h->TL, C, h->M.nnp, rev, want_c, want_s, lflag, h->nout, h->nall, h->bc_int, h->n_int, h->norm, h->AE,
          h->tile_acc, h->tile_flags);
    mark(h, KID_TBUILD, 3);
    if (h->dist) {  // all ranks must take the same path: global pair count and overflow flags
      NCCL_TRY(h->p_ncclAllReduce(h->tile_acc, h->tile_acc, 1, ncclInt64, ncclSum, h->comm, s));
      NCCL_TRY(h->p_ncclAllReduce(h->tile_flags, h->tile_flags, 6, ncclInt32, ncclMax, h->comm, s));
    }
    k_tile_status<<<1, 32, 0, s>>>(h->G, h->cell_start, h->cell_stride, h->nout, h->tile_acc, h->tile_flags, h->tstat_d);
    mark(h, KID_STATUS);
    CUDA_TRY(cudaMemcpyAsync(h->tstat_h, h->tstat_d, sizeof(TileStatus), cudaMemcpyDeviceToHost, s));
    if (h->dist && dist_status_enqueue(h)) return 1;
    CUDA_TRY(cudaStreamSynchronize(s));
    if (h->profiling) mark(h, -1, 0);  // do not charge the host round trip to the next kernel
    st = *h->tstat_h;
    if (h->dist && dist_status_apply(h)) return 1;
    if (st.overflow) {
      h->err = "cell grid larger than the capacity derived from Xmin_Domain/Xmax_Domain";
      return 1;
    }
    for (int k = 0; k < 3; ++k) h->nloc[k] = st.nloc[k];
    h->nloc_valid = true;
    if (st.flags & 1) {  // more candidates in a stencil row than a mask holds: this problem stays on the list path
      h->tile_off = true;
      return 0;
    }
    if (!(st.flags & 2)) break;
    if (attempt == 1) return 0;
    // a list is longer than its slice: enlarge the slices and build again
    auto fit = [](int need, int cap) { return need > cap ? ((need + need / 4 + 4 + 3) & ~3) : cap; };
    h->TL.capN0 = fit(st.max_n0n, h->TL.capN0);
    h->TL.capS0 = fit(st.max_n0s, h->TL.capS0);
    if (want_c || want_s) h->TL.capC = fit(st.max_n1n, h->TL.capC);
    if (want_s) h->TL.capD = fit(st.max_n1s, h->TL.capD);
    if (tile_alloc_weights(h)) return 1;
  }
  {  // tile geometry of every block of the pair-sum kernels
    const int nb = std::max((h->M.nsp + TS_T - 1) / TS_T, (h->M.nnp + TN_T - 1) / TN_T);
    k_tile_geoms<<<dim3((nb + 127) / 128, 5), 128, 0, s>>>(h->G, S, h->nout, h->GT, TB_SP_CAP, TB_N_CAP, TMV_CAP, TMV_CAP,
                                                           TAV_CAP);
    mark(h, KID_TBUILD);
  }
  // list-growth rule: forward order unless the list grew; a list that grows after the first step
  // is walked in split order, which only the list path implements
  if (st.n_pairs > h->m_pairs && h->m_pairs > 0) return 0;
  h->last_m_before = h->m_pairs;
  h->last_n_pairs = st.n_pairs;
  h->tile_last_mode = rev;  // (no pairs at all in a first step: nothing is walked in either order)
  if (st.n_pairs > h->m_pairs) h->m_pairs = st.n_pairs;
  *use = true;
  return 0;
}

int tile_step(spsph_handle *h, int itimestep) {
  const DevParams &P = h->P;
  const spsph_params &p = h->hp;
  cudaStream_t s = h->stream;
  const SortArrays S = sort_arrays(h);
  const SortedConsts C = sorted_consts(h);
  const TileLists &L = h->TL;
  const TileRecs &R = h->TR;
  SlotMap M = h->M;  // launch extents: only the particles this rank has to process (same slot layout)
  M.nn = h->nloc[0];
  M.ns = h->nloc[1];
  const int rev = h->tile_last_mode;
  const int GN = (M.nn + 127) / 128;
  const int GSs = (M.ns + TS_T - 1) / TS_T, GNn = (M.nn + TN_T - 1) / TN_T, GNs = (M.nn + TS_T - 1) / TS_T;
  const int adapt = P.adapt, bc = p.no_bcs > 0 ? 1 : 0;
  cudaStream_t s2 = h->dual ? h->stream2 : s;
  auto fork = [&]() {
    if (s2 != s) {
      cudaEventRecord(h->ev_fork, s);
      cudaStreamWaitEvent(s2, h->ev_fork, 0);
    }
  };
  auto join = [&]() {
    if (s2 != s) {
      cudaEventRecord(h->ev_join, s2);
      cudaStreamWaitEvent(s, h->ev_join, 0);
    }
  };
  // SPH_shift block, main:99-109: state between steps -> the other format-B buffer set
  if (p.sph_shift && itimestep > 1 && ((itimestep - 1) % p.shift_update == 0)) {
    const StatePtrs sw = state_ptrs(h, 1 - h->cur);
    fork();
    if (GSs) k_tile_a_sp<true><<<GSs, TS_T, 0, s>>>(P, M, S, L, C, R, sw, h->GT.g[0], adapt, 0, 0);
    if (GNn) k_tile_a_node<true, true><<<GNn, TN_T, 0, s2>>>(P, M, S, L, C, R, sw, h->GT.g[1], adapt, 0, 0);
    join();
    mark(h, KID_SWEEPA, 2);
    h->cur = 1 - h->cur;
  }
  const StatePtrs st = state_ptrs(h, h->cur);
  const int extra = (p.inside_approach && p.ndummy2 > 0) ? 1 : 0;
  if (extra) {  // boundary_forces (main:742): x is frozen during the 4 stages
    k_bound_force<<<GN, 128, 0, s>>>(P, M, h->G, S, p.ndummy2, h->fbound);
    mark(h, KID_MOVE);
  }
  // RK4, main:653-802
  k_rk_begin<<<list_grid(h, P.ntotal, 256), 256, 0, s>>>(P, st, local_list(h, P.ntotal), h->pos_of, R.NAs, R.SAs);
  mark(h, KID_RKBEGIN);
  const double f1rk[4] = {0., 0.5, 0.5, 1.0}, f2rk[4] = {1., 2., 2., 1.0};
  const bool artv = (P.alpha > 0 || P.beta > 0);
  for (int stg = 0; stg < 4; ++stg) {
    fork();
    if (GSs) k_tile_a_sp<false><<<GSs, TS_T, 0, s>>>(P, M, S, L, C, R, st, h->GT.g[0], adapt, bc, 0);
    if (GNn) k_tile_a_node<false, false><<<GNn, TN_T, 0, s2>>>(P, M, S, L, C, R, st, h->GT.g[1], adapt, bc, 0);
    join();
    mark(h, KID_SWEEPA, 2);
    const int last = (stg == 3);
    const double f1n = last ? 0.0 : f1rk[stg + 1];
    fork();
    if (GSs) k_tile_b_sp<<<GSs, TS_T, 0, s>>>(P, M, h->G, S, L, C, R, st, h->GT.g[0], rev, f1n, f2rk[stg], last);
    if (artv && GNn) k_tile_av<<<GNn, TN_T, 0, s2>>>(P, M, S, L, R, st, h->GT.g[4], h->h_uniform);
    if (GNn) k_tile_b_node<<<GNn, TN_T, 0, s2>>>(P, M, h->G, S, L, R, st, h->GT.g[1], rev, f1n, f2rk[stg], last, extra);
    join();
    mark(h, KID_SWEEPB, artv ? 3 : 2);
  }
  // final stress_point_update + adapt_stress2 + BCs, main:130-135
  fork();
  if (GSs) k_tile_a_sp<false><<<GSs, TS_T, 0, s>>>(P, M, S, L, C, R, st, h->GT.g[0], adapt, bc, 1);
  if (GNn) k_tile_a_node<false, true><<<GNn, TN_T, 0, s2>>>(P, M, S, L, C, R, st, h->GT.g[1], adapt, bc, 1);
  join();
  mark(h, KID_SWEEPA, 2);
  // positions, main:140-182
  if (GNs + GSs)
    k_tile_move<<<GNs + GSs, TS_T, 0, s>>>(P, M, S, L, C, R, st, h->GT.g[2], h->GT.g[3], GNs, h->x, h->x00, h->displ);
  mark(h, KID_MOVE);
  if (h->x_fs && !h->dist) {
    CUDA_TRY(cudaMemcpyAsync(h->x_fs, h->x, 2 * (size_t)P.ntotal2 * sizeof(double), cudaMemcpyDeviceToDevice, s));
    h->x_fs_valid = true;
  }
  if (p.update_x && p.sp_sph && !p.inside_approach) {
    k_shift<<<list_grid(h, P.nnode, 256), 256, 0, s>>>(P, st.NB, h->x, h->x_10, h->disp_10, h->bc_int, h->n_int,
                                                       local_list(h, P.nnode));
    mark(h, KID_SHIFT);
  }
  CUDA_TRY(cudaGetLastError());
  h->have_lists = false;
  h->tile_last = true;
  ++h->tile_steps;
  if (h->profiling) prof_collect(h);
  return 0;
}

// the id lists of the last step, built on demand after a tile step (free-surface detection at download, spsph_pairs)
static int materialize_lists(spsph_handle *h) {
  if (h->have_lists || !h->tile_last) return 0;
  const long long keep_launches = h->launches;
  const bool prof = h->profiling;
  h->profiling = false;
  const int rc = build_lists(h, h->tile_last_mode);
  h->profiling = prof;
  h->launches = keep_launches;
  if (rc) return 1;
  h->have_lists = true;
  return 0;
}

// sweep A launches: the UMOR variants (uniform mass/rho per species, see k_sweep_a_sp) are chosen at run time
#define SPSPH_LAUNCH_A_SP(FIRST, FROMB, GRID, STREAM, ST, DOBC)                                                       \
  do {                                                                                                               \
    if (h->umor)                                                                                                     \
      k_sweep_a_sp<FIRST, FROMB, true><<<GRID, SWEEP_T, 0, STREAM>>>(P, M, ord_s, h->L, cnt0, ST, adapt, DOBC, h->pal_node); \
    else                                                                                                             \
      k_sweep_a_sp<FIRST, FROMB, false><<<GRID, SWEEP_T, 0, STREAM>>>(P, M, ord_s, h->L, cnt0, ST, adapt, DOBC, h->pal_node); \
  } while (0)
#define SPSPH_LAUNCH_A_NODE(FIRST, FROMB, EPSP, GRID, STREAM, ST, DOBC)                                                      \
  do {                                                                                                                      \
    if (h->umor)                                                                                                            \
      k_sweep_a_node<FIRST, FROMB, EPSP, true><<<GRID, SWEEP_T, 0, STREAM>>>(P, M, ord_n, h->L, cnt0, ST, adapt, DOBC, h->pal_sp); \
    else                                                                                                                    \
      k_sweep_a_node<FIRST, FROMB, EPSP, false><<<GRID, SWEEP_T, 0, STREAM>>>(P, M, ord_n, h->L, cnt0, ST, adapt, DOBC, h->pal_sp); \
  } while (0)

int step_impl(spsph_handle *h, int itimestep, double time_sph, double dt) {
  if (!h->uploaded) {
    h->err = "spsph_step called before spsph_upload";
    return 1;
  }
  step_scalars(h, itimestep, time_sph, dt);
  if (h->dist && halo_exchange(h)) return 1;
  bool sorted = false;
#ifdef SPSPH_HOST_EMU
  if (h->emu_lists) sorted = true;  // lockstep emulation: the harness hands the sorted arrays and lists over
#endif
  if (!sorted) {
    if (sort_particles(h)) return 1;
    if (h->tile_cfg && h->tile_env && !h->tile_off) {
      bool use_tile = false;
      if (tile_build(h, &use_tile)) return 1;
      if (use_tile) return tile_step(h, itimestep);
    }
  }
  if (build_lists(h)) return 1;
  h->tile_last = false;
  ++h->list_steps;
  const DevParams &P = h->P;
  const spsph_params &p = h->hp;
  cudaStream_t s = h->stream;
  const SortArrays S = sort_arrays(h);
  SlotMap M = h->M;  // launch extents: only the particles this rank has to process (same slot layout)
  M.nn = h->nloc[0];
  M.ns = h->nloc[1];
  const int *lflag = h->dist ? h->lflag : nullptr;
  const int GN = (M.nn + 127) / 128, GB = (M.nnp + M.nsp + 127) / 128;
  const int GNW = (M.nn + SWEEP_T - 1) / SWEEP_T, GSW = (M.ns + SWEEP_T - 1) / SWEEP_T;  // pair-sum kernels
  const int adapt = P.adapt, bc = p.no_bcs > 0 ? 1 : 0;
  const int *ord_n = S.order[0], *ord_s = S.order[1];
  bool first_a = true;
  // list lengths the pair-sum kernels walk: the arrays of the neighbour build, or (slab runs) the peeled copy that
  // serves the current dependent sweep -- ghosts too deep to matter any more have length zero there (k_peel_counts)
  const int *cnt0 = h->n0, *cnt1 = h->n1;
  int sweep_no = 0;
  auto next_sweep = [&]() {
    ++sweep_no;
    const int level = h->peel ? std::min(PEEL_LEVELS, (sweep_no - 1) / 3) : 0;
    cnt0 = level ? h->peel0 + (size_t)(level - 1) * h->peel_stride : h->n0;
    cnt1 = level ? h->peel1 + (size_t)(level - 1) * h->peel_stride : h->n1;
  };
  if (h->peel) {
    k_peel_counts<<<148 * 4, 256, 0, s>>>(M, S, h->D, h->halo_cells, h->n0, h->n1, h->peel0, h->peel1, h->peel_stride);
    mark(h, KID_HALO);
  }
  // SPH_shift block, main:99-109: format B -> the other format-B buffer set
  if (p.sph_shift && itimestep > 1 && ((itimestep - 1) % p.shift_update == 0)) {
    next_sweep();
    const StatePtrs sw = state_ptrs(h, 1 - h->cur);
    SPSPH_LAUNCH_A_SP(true, true, GSW, s, sw, 0);
    SPSPH_LAUNCH_A_NODE(true, true, true, GNW, s, sw, 0);
    if (p.cont_density) k_commit_node_rho<<<GN, 128, 0, s>>>(P, M, ord_n, sw);
    mark(h, KID_SWEEPA, 2);
    h->cur = 1 - h->cur;
    first_a = false;
  }
  const StatePtrs st = state_ptrs(h, h->cur);
  const bool std_sph = !p.sp_sph;
  if (p.inside_approach && p.ndummy2 > 0) {  // boundary_forces (main:742): x is frozen during the 4 stages
    k_bound_force<<<GN, 128, 0, s>>>(P, M, h->G, S, p.ndummy2, h->fbound);
    mark(h, KID_MOVE);
  }
  // RK4, main:653-802
  k_rk_begin<<<list_grid(h, P.ntotal, 256), 256, 0, s>>>(P, st, local_list(h, P.ntotal), nullptr, nullptr, nullptr);
  mark(h, KID_RKBEGIN);
  const double f1rk[4] = {0., 0.5, 0.5, 1.0}, f2rk[4] = {1., 2., 2., 1.0};
  // The node-side and the stress-particle-side kernel of a sweep touch disjoint outputs and only read the
  // other side's previous-format records, so they run side by side on two streams (s2 forks from / joins s).
  // Continuity density (cont_density = T): the densities move inside the step, so every sweep A recomputes
  // cspm_norm and the (m/rho) w factors (FIRST variants), every sweep B the CSPM matrix, and the kernels run in the
  // reference's read-before-write order on one stream.
  const bool cd = p.cont_density != 0;
  cudaStream_t s2 = (h->dual && !cd) ? h->stream2 : s;
  auto fork = [&]() {
    if (s2 != s) {
      cudaEventRecord(h->ev_fork, s);
      cudaStreamWaitEvent(s2, h->ev_fork, 0);
    }
  };
  auto join = [&]() {
    if (s2 != s) {
      cudaEventRecord(h->ev_join, s2);
      cudaStreamWaitEvent(s, h->ev_join, 0);
    }
  };
  for (int stg = 0; stg < 4; ++stg) {
    next_sweep();
    if (std_sph) {
      k_sweep_a_std<<<list_grid(h, P.ntotal, 256), 256, 0, s>>>(P, st, local_list(h, P.ntotal));
    } else {
      fork();
      if (first_a || cd) {
        SPSPH_LAUNCH_A_SP(true, false, GSW, s, st, bc);
        SPSPH_LAUNCH_A_NODE(true, false, false, GNW, s2, st, bc);
        if (cd) k_commit_node_rho<<<GN, 128, 0, s>>>(P, M, ord_n, st);
      } else {
        SPSPH_LAUNCH_A_SP(false, false, GSW, s, st, bc);
        SPSPH_LAUNCH_A_NODE(false, false, false, GNW, s2, st, bc);
      }
      join();
    }
    mark(h, KID_SWEEPA, 2);
    first_a = false;
    const int last = (stg == 3);
    const int sflags = (last ? 1 : 0) | (stg == 0 ? 2 : 0);  // k_sweep_b_*: last stage / accumulators start from zero
    const double f1n = last ? 0.0 : f1rk[stg + 1];
    const bool artv = (P.alpha > 0 || P.beta > 0);
    next_sweep();
    fork();
    if (artv) {
      if (h->uniform_h)
        k_artvisc<true><<<GNW, SWEEP_T, 0, s2>>>(P, M, ord_n, h->L, cnt1, st, h->h_uniform);
      else
        k_artvisc<false><<<GNW, SWEEP_T, 0, s2>>>(P, M, ord_n, h->L, cnt1, st, 0.f);
    }
    if (p.art_stress) {  // main:746
      k_art_force_prep<<<GN, 128, 0, s2>>>(P, M, ord_n, st);
      k_art_force<<<GN, 128, 0, s2>>>(P, M, ord_n, h->L, cnt1, st, h->art_w2);
    }
    const double f2n = last ? 0.0 : f2rk[stg + 1];
    if (cd) {  // the node side reads the stress particles' density before their side integrates it
      k_sweep_b_node<true><<<GNW, SWEEP_T, 0, s>>>(P, M, ord_n, h->L, cnt0, st, f1n, f2rk[stg], sflags);
      k_sweep_b_sp<true><<<GSW, SWEEP_T, 0, s>>>(P, M, ord_s, h->L, cnt0, st, f1n, f2rk[stg], sflags, f2n);
    } else if (stg == 0) {
      k_sweep_b_sp<true><<<GSW, SWEEP_T, 0, s>>>(P, M, ord_s, h->L, cnt0, st, f1n, f2rk[stg], sflags, f2n);
      k_sweep_b_node<true><<<GNW, SWEEP_T, 0, s2>>>(P, M, ord_n, h->L, cnt0, st, f1n, f2rk[stg], sflags);
    } else {
      k_sweep_b_sp<false><<<GSW, SWEEP_T, 0, s>>>(P, M, ord_s, h->L, cnt0, st, f1n, f2rk[stg], sflags, f2n);
      k_sweep_b_node<false><<<GNW, SWEEP_T, 0, s2>>>(P, M, ord_n, h->L, cnt0, st, f1n, f2rk[stg], sflags);
    }
    join();
    mark(h, KID_SWEEPB, artv ? 3 : 2);
  }
  // final stress_point_update + adapt_stress2 + BCs, main:130-135
  next_sweep();
  if (std_sph) {
    k_sweep_a_std<<<list_grid(h, P.ntotal, 256), 256, 0, s>>>(P, st, local_list(h, P.ntotal));
  } else {
    fork();
    if (cd) {
      SPSPH_LAUNCH_A_SP(true, false, GSW, s, st, bc);
      SPSPH_LAUNCH_A_NODE(true, false, true, GNW, s, st, bc);
      k_commit_node_rho<<<GN, 128, 0, s>>>(P, M, ord_n, st);
    } else {
      SPSPH_LAUNCH_A_SP(false, false, GSW, s, st, bc);
      SPSPH_LAUNCH_A_NODE(false, false, true, GNW, s2, st, bc);
    }
    join();
  }
  mark(h, KID_SWEEPA, 2);
  // positions, main:140-182
  next_sweep();
  k_move<<<GB, 128, 0, s>>>(P, M, S, h->L, cnt1, st, h->x, h->x00, h->displ);
  mark(h, KID_MOVE);
  if (h->fs_each_step) {
    // get_nodes_on_free_surface, main:152-154: after the position update, before the stress particles are re-seated;
    // its marks (bc_or_not = 2) and normals feed the BCs of the next step
    if (p.xsph) k_xsph_marks<<<GN, 128, 0, s>>>(P, M, S, h->L, h->n1, h->bc_or_not);
    k_free_surface<<<GB, 128, 0, s>>>(P, M, S, h->pos_of, h->L, h->n0, h->n1, h->growth, h->x, h->mass, h->rho, h->hsml,
                                      h->bc_or_not, h->fs_cov);
    if (P.ifsigman)
      k_fs_normals<<<GN, 128, 0, s>>>(P, M, S, h->pos_of, h->L, h->n0, h->n1, h->growth, h->x, h->mass, h->rho,
                                      h->bc_or_not, h->fs_cov, h->fs_normal);
    mark(h, KID_MOVE, P.ifsigman ? (p.xsph ? 3 : 2) : (p.xsph ? 2 : 1));
  }
  if (p.update_x && std_sph) {  // main:166
    k_sp_follow<<<list_grid(h, P.nnode, 256), 256, 0, s>>>(P, h->x, local_list(h, P.nnode));
    mark(h, KID_SHIFT);
  }
  if (h->x_fs && !h->dist) {
    CUDA_TRY(cudaMemcpyAsync(h->x_fs, h->x, 2 * (size_t)P.ntotal2 * sizeof(double), cudaMemcpyDeviceToDevice, s));
    h->x_fs_valid = true;
  }
  if (p.update_x && p.sp_sph && !p.inside_approach) {
    k_shift<<<list_grid(h, P.nnode, 256), 256, 0, s>>>(P, st.NB, h->x, h->x_10, h->disp_10, h->bc_int, h->n_int,
                                                       local_list(h, P.nnode));
    mark(h, KID_SHIFT);
  }
  CUDA_TRY(cudaGetLastError());
  h->have_lists = true;
  if (h->profiling) prof_collect(h);
  return 0;
}

}  // namespace


// host-side passes over the particles of an upload: a few threads, contiguous chunks in index order
static int host_threads(size_t n) {
#ifdef SPSPH_HOST_EMU
  (void)n;
  return 1;
#else
  if (n < (1u << 18)) return 1;
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 1;
  if (const char *e = getenv("SPSPH_HOST_THREADS")) hw = (unsigned)std::max(1, atoi(e));
  return (int)std::min<unsigned>(hw, 8u);
#endif
}
template <class F>
static void parallel_chunks(size_t n, int nthr, F fn) {
  const size_t per = (n + nthr - 1) / nthr;
  auto run = [&](int c) {
    const size_t i0 = std::min(n, (size_t)c * per), i1 = std::min(n, i0 + per);
    fn(c, i0, i1);
  };
#ifndef SPSPH_HOST_EMU
  if (nthr > 1) {
    std::vector<std::thread> th;
    for (int c = 1; c < nthr; ++c) th.emplace_back(run, c);
    run(0);
    for (auto &t : th) t.join();
    return;
  }
#endif
  for (int c = 0; c < nthr; ++c) run(c);
}

extern "C" {

const char *spsph_version(void) { return "spsph-b200 0.1 (sm_100a)"; }

const char *spsph_last_error(spsph_handle *h) { return h ? h->err.c_str() : "null handle"; }

// w of kernel(r, ., h) on the host (main:1468-1536), for the constant w2 = W(dx, 1.2 dx) of artificial_force (main:921-926)
static double host_kernel_w(int skf, double pi, double r, double h) {
  const double q = r / h;
  auto p5 = [](double a) { const double a2 = a * a; return a * (a2 * a2); };
  if (skf == 1) {
    const double factor = 15.e0 / (7.e0 * pi * h * h);
    if (q >= 0 && q <= 1.e0) return factor * ((double)(2.f / 3.f) - q * q + q * q * q / 2.);
    if (q > 1.e0 && q <= 2) return factor * 1.e0 / 6.e0 * ((2. - q) * (2. - q) * (2. - q));
  } else if (skf == 2) {
    const double factor = 1.e0 / (std::pow(h, 2) * std::pow(pi, 2 / 2.));
    if (q >= 0 && q <= 3) return factor * std::exp(-q * q);
  } else if (skf == 3) {
    const double factor = 7.e0 / (478.e0 * pi * h * h);
    if (q >= 0 && q <= 1) return factor * (p5(3 - q) - 6 * p5(2 - q) + 15 * p5(1 - q));
    if (q > 1 && q <= 2) return factor * (p5(3 - q) - 6 * p5(2 - q));
    if (q > 2 && q <= 3) return factor * p5(3 - q);
  }
  return 0.0;
}

int spsph_create(spsph_handle **out, const spsph_params *p, int device) {
  if (!out || !p) return 1;
  *out = nullptr;
  spsph_handle *h = new spsph_handle();
  *out = h;  // returned even on failure so that spsph_last_error can be read; caller destroys it
  if (p->struct_bytes != (int32_t)sizeof(spsph_params)) {
    h->err = "spsph_params layout mismatch (struct_bytes)";
    return 1;
  }
  h->hp = *p;
  {
    const char *e = std::getenv("SPSPH_FORCE_FILL_SCAN");
    h->force_fill_scan = e && e[0] == '1';
  }
  // ---- scope checks: everything the reference would run for these inputs must exist on the device ----
  auto fail = [&](const char *m) {
    h->err = m;
    return 1;
  };
  if (p->ndimn != 2 || p->nstre != 4) return fail("only ndimn = 2, nstre = 4 (plane strain) is supported");
  if (p->skf < 1 || p->skf > 3) return fail("skf must be 1 (cubic spline), 2 (Gauss) or 3 (quintic)");
  if (p->ntype_eco > 1 && !((p->ncrit >= 1 && p->ncrit <= 4) || p->ncrit == 12))
    return fail("this yield criterion is not supported: ncrit must be 1 (Tresca), 2 (von Mises), 3 (Mohr-Coulomb), "
                "4 (Drucker-Prager, Perzyna) or 12 (Drucker-Prager, Bui et al.); 5 (Cam Clay) is not built");
  if (p->ntype_eco > 1 && p->ncrit <= 4 && !(p->props[6] > (double)0.001f))
    return fail("initial yield surface size too small (the reference STOPs, mat:2322-2332)");
  if (p->sph_shift && p->shift_update <= 0) return fail("shift_update must be positive");
  if (p->no_bcs > 16) return fail("too many BCs");
  if (p->nnode + p->nstress != p->ntotal || p->ntotal + p->ndummy != p->ntotal2) return fail("inconsistent counts");
  // indices into the species-sorted arrays (3 rows of ntotal2) and partner ids with two class bits are 32-bit
  if ((long long)p->ntotal2 * 3 >= (1ll << 31)) return fail("more than 7.1e8 particles on one device");
  if (!p->inside_approach && p->nstress != p->nnode * p->npoints) return fail("nstress != npoints*nnode");
  if (!p->sp_sph && (p->nstress != p->nnode || p->sph_shift)) return fail("standard SPH needs nstress == nnode");
  if (p->ndummy2 < 0 || p->ndummy2 > p->ndummy) return fail("ndummy2 out of range");
  if (p->no_bcs < 0) return fail("no_bcs must not be negative");
  if (p->ntcurves < 0 || p->ntcurves > SPSPH_MAX_TCURVES) return fail("ntcurves exceeds SPSPH_MAX_TCURVES");
  for (int k = 0; k < p->ntcurves; ++k)
    if (p->nptstcurves[k] < 0 || p->nptstcurves[k] > SPSPH_MAX_TCURVE_PTS)
      return fail("a time curve has more points than SPSPH_MAX_TCURVE_PTS");
  if (p->npoints < 1 || p->npoints > 3) return fail("npoints must be 1, 2 or 3");
  if (p->vel_vector && p->sp_sph && !p->inside_approach && p->npoints != 2)
    return fail("vel_vector re-seating needs two stress particles per velocity particle (main:284-293)");

  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail("no CUDA device available (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail("bad CUDA device index");
  h->device = device;
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  if (const char *e = getenv("SPSPH_DUAL_STREAM")) h->dual = atoi(e) != 0;
  if (const char *e = getenv("SPSPH_TILE")) h->tile_env = atoi(e) != 0;
  CUDA_TRY(cudaEventCreate(&h->ev0));
  CUDA_TRY(cudaEventCreate(&h->ev1));

  DevParams &P = h->P;
  P.nnode = p->nnode;
  P.nstress = p->nstress;
  P.ntotal = p->ntotal;
  P.ntotal2 = p->ntotal2;
  P.ndummy = p->ndummy;
  P.npoints = p->npoints;
  P.skf = p->skf;
  P.cont_density = p->cont_density ? 1 : 0;
  P.sle = p->sle;
  P.scale_k = (p->skf == 1) ? 2 : 3;
  P.cspm = p->cspm;
  P.update_x = p->update_x;
  P.xsph = p->xsph;
  P.ncrit = p->ncrit;
  P.ntype_eco = p->ntype_eco;
  P.ntype_solid = p->ntype_solid;
  P.no_bcs = p->no_bcs;
  h->fs_each_step = p->update_x && p->no_bcs > 0 && (p->ifsigman == 1 || p->xsph);
  // without update_x no particle is ever marked (main:140-154), apply_stress_free has nothing to act on
  P.ifsigman = (h->fs_each_step && p->ifsigman == 1) ? 1 : 0;
  P.bc_nloop = p->bc_loop_ntotal ? p->ntotal : p->nnode;
  P.sp_sph = p->sp_sph;
  P.inside_approach = p->inside_approach;
  P.vel_vector = p->vel_vector;
  P.shift_update = p->shift_update > 0 ? p->shift_update : 1;
  P.adapt = (p->ncrit == 12) ? 1 : 0;
  // main:843 (artificial_viscosity), main:203 (XSPH_update), main:347/384 (isolated_nodes via shift_stress_points)
  P.track_nint = ((p->alpha > 0 || p->beta > 0) || (p->update_x && (p->xsph || (p->sp_sph && !p->inside_approach)))) ? 1 : 0;
  P.pi = p->pi;
  P.D11 = p->D11;
  P.D12 = p->D12;
  P.D22 = p->D22;
  P.D33 = p->D33;
  P.D41 = p->D41;
  P.D42 = p->D42;
  P.alpha = p->alpha;
  P.beta = p->beta;
  P.damping = p->damping;
  P.dx = p->dx;
  h->art_w2 = host_kernel_w(p->skf, p->pi, p->dx, (double)1.2f * p->dx);
  P.r_x = p->r_x;
  P.r_y = p->r_y;
  P.disp_tol = p->disp_tol;
  P.ae_thr = (double)p->ae_threshold;
  for (int k = 0; k < 20; ++k) P.props[k] = p->props[k];
  for (int d = 0; d < 2; ++d) {
    P.xmin_dom[d] = p->xmin_domain[d];
    P.xmax_dom[d] = p->xmax_domain[d];
  }
  P.snphi = std::sin(p->props[8] * (double)0.017453292f);  // invar09 :2287-2288, yieldf09 :2441, 2456
  {  // adapt_stress2 constants, mat:2096-2100 (fp64)
    const double tanfi = p->props[12], coh = p->props[13];
    P.dp_alpha2 = tanfi / (std::sqrt(9 + 12 * (tanfi * tanfi)));
    P.dp_kc = (3 * coh) / (std::sqrt(9 + 12 * (tanfi * tanfi)));
  }

  SlotMap &M = h->M;
  M.nn = p->nnode;
  M.ns = p->nstress;
  M.nd = p->ndummy;
  M.nnp = round_up(M.nn, SLICE);
  M.nsp = round_up(M.ns, SLICE);
  M.ndp = round_up(M.nd, SLICE);
  h->nslices = (M.nnp + M.nsp) / SLICE;

  const size_t n2 = (size_t)p->ntotal2, nt = (size_t)p->ntotal, nn = (size_t)p->nnode, ns = (size_t)p->nstress;
  int rc = 0;
  rc |= dalloc(h, &h->x, 2 * n2) | dalloc(h, &h->x00, 2 * n2) | dalloc(h, &h->rho, n2) | dalloc(h, &h->mass, n2);
  rc |= dalloc(h, &h->hsml, n2) | dalloc(h, &h->mor, n2) | dalloc(h, &h->mrho, n2) | dalloc(h, &h->mcls, n2);
  rc |= dalloc(h, &h->NA, 2 * nn) | dalloc(h, &h->SA, ns) | dalloc(h, &h->NSa, 4 * nn) | dalloc(h, &h->SVa, 2 * ns);
  rc |= dalloc(h, &h->av, 2 * nn) | dalloc(h, &h->fbound, 2 * nn) | dalloc(h, &h->aforce, 2 * nn) | dalloc(h, &h->RN, nn);
  if (p->cont_density)
    rc |= dalloc(h, &h->rho_new, nn) | dalloc(h, &h->rho0, nt - nn) | dalloc(h, &h->hsml0, nt - nn) |
          dalloc(h, &h->RKrho, nt - nn) | dalloc(h, &h->RKh, nt - nn) | dalloc(h, &h->divu, nt - nn);
  for (int b = 0; b < 2; ++b) {
    rc |= dalloc(h, &h->NB[b], nn) | dalloc(h, &h->SB[b], ns) | dalloc(h, &h->NSb[b], 4 * nn);
    rc |= dalloc(h, &h->SFb[b], 4 * ns) | dalloc(h, &h->SVb[b], ns);
  }
  rc |= dalloc(h, &h->stage_vel, 2 * nt) | dalloc(h, &h->stage_stress, 4 * nt);
  rc |= dalloc(h, &h->epsp, nt) | dalloc(h, &h->fdp, nt) | dalloc(h, &h->norm, nt);
  rc |= dalloc(h, &h->ivars, (size_t)SPSPH_NINT_VARS * nt);
  rc |= dalloc(h, &h->AE, 5 * nt) | dalloc(h, &h->vel0, 2 * nn) | dalloc(h, &h->stress0, 4 * ns);
  rc |= dalloc(h, &h->vx0, 2 * nt) | dalloc(h, &h->RKv, 2 * nn) | dalloc(h, &h->RKs, 4 * ns) | dalloc(h, &h->RKe, ns);
  rc |= dalloc(h, &h->displ, 2 * nn) | dalloc(h, &h->x_10, 2 * nn) | dalloc(h, &h->disp_10, nn);
  rc |= dalloc(h, &h->wallpos, n2) | dalloc(h, &h->horiz, n2) | dalloc(h, &h->n_int, nn);
  rc |= dalloc(h, &h->bc_int, nn) | dalloc(h, &h->bc_or_not, nt) | dalloc(h, &h->bc_info, 8 * nt);
  if (h->fs_each_step) rc |= dalloc(h, &h->fs_cov, nt) | dalloc(h, &h->fs_normal, 2 * nn);
  if (p->update_x && p->sp_sph && !p->inside_approach && !h->fs_each_step) rc |= dalloc(h, &h->x_fs, 2 * n2);
  rc |= dalloc(h, &h->if_out, n2);
  rc |= dalloc(h, &h->G, 1);
  h->bbox_blocks = 148 * 8;
  rc |= dalloc(h, &h->bbox_partial, 6 * (size_t)h->bbox_blocks);
  rc |= dalloc(h, &h->which_cell, n2) | dalloc(h, &h->tmp_ids, 3 * n2) | dalloc(h, &h->order, 3 * n2) | dalloc(h, &h->ordc, 3 * n2);
  rc |= dalloc(h, &h->scell, 3 * n2) | dalloc(h, &h->pos_of, n2) | dalloc(h, &h->nout, 8) | dalloc(h, &h->bb6, 8);
  rc |= dalloc(h, &h->spos, 3 * n2) | dalloc(h, &h->sh, 3 * n2) | dalloc(h, &h->supos, 3 * n2);
  rc |= dalloc(h, &h->scan_bsum, 4 * (size_t)SCAN_BLOCKS) | dalloc(h, &h->scan_totals, 8);
  const size_t T = (size_t)M.total();
  rc |= dalloc(h, &h->n0, T) | dalloc(h, &h->n1, T) | dalloc(h, &h->nall, T);
  rc |= dalloc(h, &h->nfwd_u, n2) | dalloc(h, &h->base_u, n2) | dalloc(h, &h->cand_overflow, 4);
  rc |= dalloc(h, &h->cand0, (size_t)h->nslices * CAND_CAP * SLICE) | dalloc(h, &h->cand1, (size_t)h->nslices * CAND_CAP * SLICE);
  rc |= dalloc(h, &h->wslice, 3 * (size_t)h->nslices) | dalloc(h, &h->oslice, 3 * (size_t)h->nslices);
  rc |= dalloc(h, &h->growth, 1) | dalloc(h, &h->status_d, 1) | dalloc(h, &h->stats_d, 4);
  if (rc) return 1;
  CUDA_TRY(cudaMallocHost((void **)&h->status_h, sizeof(StepStatus)));
  CUDA_TRY(cudaMallocHost((void **)&h->dist_h, 4 * sizeof(double)));
  CUDA_TRY(cudaMemset(h->nall, 0, T * sizeof(int)));
  CUDA_TRY(cudaMemset(h->AE, 0, 5 * nt * sizeof(double)));
  CUDA_TRY(cudaMemset(h->norm, 0, nt * sizeof(double)));
  CUDA_TRY(cudaMemset(h->fbound, 0, 2 * nn * sizeof(double)));
  CUDA_TRY(cudaMemset(h->aforce, 0, 2 * nn * sizeof(double)));
  if (h->fs_each_step) {
    CUDA_TRY(cudaMemset(h->fs_cov, 0, nt * sizeof(int)));
    CUDA_TRY(cudaMemset(h->fs_normal, 0, 2 * nn * sizeof(double)));
  }
  return 0;
}

int spsph_upload(spsph_handle *h, const spsph_state *s) {
  if (!h || !s) return 1;
  const spsph_params &p = h->hp;
  const size_t n2 = (size_t)p.ntotal2, nt = (size_t)p.ntotal, nn = (size_t)p.nnode;
  if (!s->x || !s->vel || !s->stress || !s->rho || !s->mass || !s->hsml || !s->itype) {
    h->err = "spsph_upload: x, vel, stress, rho, mass, hsml and itype are required";
    return 1;
  }
  CUDA_TRY(cudaSetDevice(h->device));
  const bool up_timing = getenv("SPSPH_UPLOAD_TIMING") != nullptr;
  const auto up_t0 = std::chrono::steady_clock::now();
  auto up_lap = [&](const char *what) {
    if (up_timing)
      fprintf(stderr, "spsph_upload: %-28s at %7.2f ms\n", what,
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - up_t0).count());
  };
  // Every plain copy is queued first: the DMA engine works through them while the host analyses the input below
  // (on an error return the device state is incomplete: the run has to be uploaded again before it can step)
  h->uploaded = false;
  cudaStream_t st = h->stream;
  auto up = [&](void *d, const void *src, size_t bytes) {
    if (!src) return cudaMemsetAsync(d, 0, bytes, st);
    return cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, st);
  };
  CUDA_TRY(up(h->x, s->x, 2 * n2 * 8));
  CUDA_TRY(up(h->hsml, s->hsml, n2 * 8));
  CUDA_TRY(up(h->if_out, s->if_out_domain, n2 * 4));
  CUDA_TRY(up(h->rho, s->rho, n2 * 8));
  CUDA_TRY(up(h->mass, s->mass, n2 * 8));
  CUDA_TRY(up(h->stage_vel, s->vel, 2 * nt * 8));
  CUDA_TRY(up(h->stage_stress, s->stress, 4 * nt * 8));
  CUDA_TRY(up(h->ivars, s->internal_vars, (size_t)SPSPH_NINT_VARS * nt * 8));
  k_upload_derive<<<((int)n2 + 255) / 256, 256, 0, st>>>((int)n2, (int)nt, h->mass, h->rho, h->mor, h->mrho, h->ivars,
                                                          h->epsp);
  h->cur = 0;
  k_pack_state<<<((int)nt + 255) / 256, 256, 0, st>>>(h->P, h->stage_vel, h->stage_stress, state_ptrs(h, 0));
  CUDA_TRY(up(h->x00, s->x00 ? s->x00 : s->x, 2 * n2 * 8));
  CUDA_TRY(up(h->fdp, s->f_drucker, nt * 8));
  CUDA_TRY(up(h->displ, s->displ, 2 * nn * 8));
  CUDA_TRY(up(h->x_10, s->x_10 ? s->x_10 : s->x, 2 * nn * 8));
  CUDA_TRY(up(h->disp_10, s->disp_10, nn * 8));
  CUDA_TRY(up(h->wallpos, s->wall_position, n2 * 4));
  CUDA_TRY(up(h->horiz, s->horizontal_or_not, n2 * 4));
  CUDA_TRY(up(h->n_int, s->n_int, nn * 4));
  CUDA_TRY(up(h->bc_int, s->bc_int, nn * 4));
  CUDA_TRY(up(h->bc_or_not, s->bc_or_not, nt * 4));
  CUDA_TRY(up(h->bc_info, s->bc_info, 8 * nt * 4));
  up_lap("copies queued");
  // Host-side analysis of the input in ONE pass over the particles, on a few threads: the particle order contract
  // (nodes, stress particles, dummies: mat:961-1026), the range of the smoothing lengths, and the mass/rho palette per
  // species (sweep A rebuilds (m/rho)*w from the partner's class and the streamed weight). The chunks are merged in
  // order, so the classes are numbered by first appearance whatever the thread count.
  struct Chunk {
    bool bad_itype = false, pal_over = false;
    double hmax = 0.0, hmin = 1.e300;
    double pal[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    int npal[2] = {0, 0};
  };
  const bool want_pal = !h->hp.cont_density && nn > 0 && nt > nn && n2 < (size_t)QID_MASK;  // two id bits carry the class
  const int nthr = host_threads(n2);
  std::vector<Chunk> chunks(nthr);
  std::vector<double> &mor_h = h->up_mor;  // the reference's mass(j)/rho(j), kept for the class pass
  if (want_pal && mor_h.size() < nt) mor_h.resize(nt);  // host scratch of the handle: page-faulted once, not per upload
  const int32_t *it_h = s->itype;
  const double *hs_h = s->hsml, *ms_h = s->mass, *rh_h = s->rho;
  double *mor_p = want_pal ? mor_h.data() : nullptr;
  parallel_chunks(n2, nthr, [&, it_h, hs_h, ms_h, rh_h, mor_p](int c, size_t i0, size_t i1) {
    Chunk ck;  // thread-local: neighbouring entries of `chunks` share cache lines
    double lm = 0.0, lr = 0.0, lv = 0.0;  // last quotient: lattice set-ups repeat the same mass and density
    bool have = false;
    double hmx = 0.0, hmn = 1.e300;
    for (size_t i = i0; i < i1; ++i) {
      const int want = i < nn ? 2 : (i < nt ? 1 : 25);
      if (it_h[i] != want) ck.bad_itype = true;
      const double hh = hs_h[i];
      hmx = hh > hmx ? hh : hmx;
      hmn = hh < hmn ? hh : hmn;
      if (mor_p && i < nt) {
        const double m = ms_h[i], r = rh_h[i];
        if (!have || m != lm || r != lr) {
          lm = m;
          lr = r;
          lv = m / r;
          have = true;
        }
        mor_p[i] = lv;
        if (!ck.pal_over) {
          const int sp = i < nn ? 0 : 1;
          int k = 0;
          while (k < ck.npal[sp] && ck.pal[sp][k] != lv) ++k;
          if (k == ck.npal[sp]) {
            if (k == 4)
              ck.pal_over = true;
            else
              ck.pal[sp][ck.npal[sp]++] = lv;
          }
        }
      }
    }
    ck.hmax = hmx;
    ck.hmin = hmn;
    chunks[c] = ck;
  });
  up_lap("analysis pass done");
  double hmax = 0.0, hmin = 1.e300;
  bool u = want_pal;
  double pal[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  int npal[2] = {0, 0};
  for (const Chunk &ck : chunks) {
    if (ck.bad_itype) {
      h->err = "spsph_upload: itype does not follow the order nodes(2), stress particles(1), dummies(25)";
      return 1;
    }
    hmax = std::fmax(hmax, ck.hmax);
    hmin = std::fmin(hmin, ck.hmin);
    if (ck.pal_over) u = false;
    for (int sp = 0; u && sp < 2; ++sp)
      for (int k = 0; u && k < ck.npal[sp]; ++k) {
        int j = 0;
        while (j < npal[sp] && pal[sp][j] != ck.pal[sp][k]) ++j;
        if (j == npal[sp]) {
          if (j == 4)
            u = false;
          else
            pal[sp][npal[sp]++] = ck.pal[sp][k];
        }
      }
  }
  if (!u) std::memset(pal, 0, sizeof(pal));
  if (const char *e = getenv("SPSPH_NO_UMOR")) u = u && atoi(e) == 0;
  h->h_itype.assign(s->itype, s->itype + n2);
  up_lap("itype kept");
  // hsml only changes on the device with cont_density and sle = 2 (main:709-712)
  const bool uh = (hmin == hmax && !(h->hp.cont_density && h->hp.sle == 2));
  if (uh != h->uniform_h) h->capC = 0;  // list C was sized for the other mode
  h->uniform_h = uh;
  h->h_uniform = (float)(0.5 * (hmax + hmax));
  h->uniform_cubic = (h->hp.skf == 1 && uh);
  if (h->hp.cont_density) CUDA_TRY(cudaMemsetAsync(h->divu, 0, (nt - nn) * sizeof(double), st));  // grad_u = 0, mat:930
  if (u != h->umor) h->cap0 = 0;  // the list-0 arrays were sized for the other mode: start over
  h->umor = u;
  h->pal_node = MorPalette{pal[0][0], pal[0][1], pal[0][2], pal[0][3]};
  h->pal_sp = MorPalette{pal[1][0], pal[1][1], pal[1][2], pal[1][3]};
  std::vector<unsigned char> &cls = h->up_cls;  // host scratch of the handle
  if (u) {
    if (cls.size() < n2) cls.resize(n2);
    std::memset(cls.data() + nt, 0, n2 - nt);  // wall particles: class 0
    parallel_chunks(nt, nthr, [&](int, size_t i0, size_t i1) {
      for (size_t i = i0; i < i1; ++i) {
        const double *pl = pal[i < nn ? 0 : 1];
        const double v = mor_h[i];
        cls[i] = (unsigned char)(v == pl[0] ? 0 : (v == pl[1] ? 1 : (v == pl[2] ? 2 : 3)));
      }
    });
    CUDA_TRY(cudaMemcpyAsync(h->mcls, cls.data(), n2, cudaMemcpyHostToDevice, st));
  }
  up_lap("classes queued");
  CUDA_TRY(cudaStreamSynchronize(st));  // the caller may reuse its arrays as soon as upload returns
  up_lap("stream drained");
  // cell-table capacity: the in-domain bounding box can never exceed the control domain (main:1187-1192)
  if (!h->cell_cnt) {
    if (!(hmax > 0.0)) {
      h->err = "spsph_upload: hsml must be positive";
      return 1;
    }
    double nc = 1.0;
    for (int d = 0; d < 2; ++d) {
      const double len = p.xmax_domain[d] - p.xmin_domain[d];
      nc *= std::floor(len / (2 * hmax)) + 2.0;
    }
    const double cap_max = 400e6;
    if (!(nc > 0) || nc > cap_max) nc = cap_max;
    h->cell_capacity = (int)nc;
    h->cell_stride = h->cell_capacity + 8;
    if (dalloc(h, &h->cell_cnt, 3 * (size_t)h->cell_stride)) return 1;
    if (dalloc(h, &h->cell_start, 3 * (size_t)h->cell_stride)) return 1;
    if (dalloc(h, &h->cell_fill, 3 * (size_t)h->cell_stride)) return 1;
  }
  h->m_pairs = 0;
  h->have_lists = false;
  h->tile_last = false;
  h->tile_off = false;
  h->tile_cfg = h->tile_env && tile_config_ok(h);
  if (h->tile_cfg && tile_setup(h)) return 1;
  h->x_fs_valid = false;
  h->uploaded = true;
#ifndef SPSPH_HOST_EMU
  if (h->dist) {  // a fresh upload holds complete data on every rank: re-derive owned / ghost / remote
    k_dist_init_flags<<<((int)n2 + 255) / 256, 256, 0, st>>>(h->P, h->D, h->x, h->lflag);
    if (rebuild_local_list(h)) return 1;
    CUDA_TRY(cudaStreamSynchronize(st));
  }
#endif
  return 0;
}

int spsph_step(spsph_handle *h, int32_t itimestep_sph, double time_sph, double dt_sph) {
  if (!h) return 1;
  CUDA_TRY(cudaSetDevice(h->device));
  return step_impl(h, itimestep_sph, time_sph, dt_sph);
}

int spsph_run(spsph_handle *h, int32_t first_itimestep, double time_sph, double dt_sph, int32_t nsteps,
              double *time_sph_out) {
  if (!h) return 1;
  CUDA_TRY(cudaSetDevice(h->device));
  const long long l0 = h->launches;
  CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
  for (int k = 0; k < nsteps; ++k) {
    if (step_impl(h, first_itimestep + k, time_sph, dt_sph)) return 1;
    time_sph = time_sph + dt_sph;  // 1_SPH_2018.f90:174
  }
  CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
  CUDA_TRY(cudaEventSynchronize(h->ev1));
  CUDA_TRY(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  h->last_launches = h->launches - l0;
  if (time_sph_out) *time_sph_out = time_sph;
  return 0;
}

int spsph_last_run_ms(spsph_handle *h, float *ms, int64_t *kernel_launches) {
  if (!h) return 1;
  if (ms) *ms = h->last_ms;
  if (kernel_launches) *kernel_launches = h->last_launches;
  return 0;
}

int spsph_profile(spsph_handle *h, int enable) {
  if (!h) return 1;
  h->profiling = enable != 0;
  h->ev_used = 0;
  for (int k = 0; k < 32; ++k) {
    h->prof_ms[k] = 0.0;
    h->prof_n[k] = 0;
  }
  return 0;
}

int spsph_profile_get(spsph_handle *h, int kid, const char **name, double *total_ms, int64_t *launches) {
  if (!h || kid < 0 || kid >= KID_N) return 1;
  if (name) *name = kKernelNames[kid];
  if (total_ms) *total_ms = h->prof_ms[kid];
  if (launches) *launches = h->prof_n[kid];
  return 0;
}

int spsph_path_counts(spsph_handle *h, int64_t *tile_steps, int64_t *list_steps) {
  if (!h) return 1;
  if (tile_steps) *tile_steps = (int64_t)h->tile_steps;
  if (list_steps) *list_steps = (int64_t)h->list_steps;
  return 0;
}

int spsph_get_list_capacity(spsph_handle *h, int64_t *m_pairs) {
  if (!h || !m_pairs) return 1;
  *m_pairs = (int64_t)h->m_pairs;
  return 0;
}

int spsph_set_list_capacity(spsph_handle *h, int64_t m_pairs) {
  if (!h) return 1;
  if (!h->uploaded || m_pairs < 0) {
    h->err = "spsph_set_list_capacity: call it after spsph_upload, with a non-negative length";
    return 1;
  }
  h->m_pairs = (long long)m_pairs;
  return 0;
}

#ifdef SPSPH_HOST_EMU
int spsph_emu_set_lists(spsph_handle *h, const spsph_emu_lists *lists) {
  if (!h || !lists) return 1;
  h->emu_lists = lists;
  return 0;
}
#endif

// ---- row-wise transfers of the time-varying state ----
static int rows_prepare(spsph_handle *h, const int32_t *ids, int32_t n, int *n_node, int *n_part) {
  const spsph_params &p = h->hp;
  if (!h->uploaded) {
    h->err = "row-wise transfers need a complete spsph_upload first (it carries the set-up arrays)";
    return 1;
  }
  if (n < 0 || (n > 0 && !ids)) {
    h->err = "row-wise transfer: bad id list";
    return 1;
  }
  int nn = 0, nt = 0;
  for (int k = 0; k < n; ++k) {
    if (ids[k] < 0 || ids[k] >= p.ntotal2 || (k > 0 && ids[k] <= ids[k - 1])) {
      h->err = "row-wise transfer: ids must be ascending particle numbers in [0, ntotal2)";
      return 1;
    }
    nn += ids[k] < p.nnode;
    nt += ids[k] < p.ntotal;
  }
  *n_node = nn;
  *n_part = nt;
  const size_t need = (size_t)(n > 0 ? n : 1) * 256;  // every array of a row set at once (< 200 bytes per row)
  if (need > h->rows_cap) {
    cudaFree(h->rows_buf);
  cudaFree(h->frame_buf);
    cudaFree(h->rows_ids);
    h->rows_buf = nullptr;
    h->rows_ids = nullptr;
    h->rows_cap = 0;
    CUDA_TRY(cudaMalloc((void **)&h->rows_buf, need + need / 8));
    CUDA_TRY(cudaMalloc((void **)&h->rows_ids, ((size_t)n + n / 8 + 16) * sizeof(int)));
    h->rows_cap = need + need / 8;
  }
  CUDA_TRY(cudaMemcpyAsync(h->rows_ids, ids, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  return 0;
}

int spsph_upload_rows(spsph_handle *h, const spsph_state *s, const int32_t *ids, int32_t n) {
  if (!h || !s) return 1;
  CUDA_TRY(cudaSetDevice(h->device));
  int nn = 0, nt = 0;
  if (rows_prepare(h, ids, n, &nn, &nt)) return 1;
  const spsph_params &p = h->hp;
  if (!s->x || !s->vel || !s->stress) {
    h->err = "spsph_upload_rows: x, vel and stress rows are required";
    return 1;
  }
  cudaStream_t st = h->stream;
  char *buf = h->rows_buf;
  const int G = 148 * 8;
  size_t off = 0;
  auto put = [&](auto *dst, const auto *src, int rows, int width, int id_base) -> int {
    using T = std::remove_cv_t<std::remove_pointer_t<decltype(src)>>;
    if (!src || rows == 0) return 0;
    const size_t bytes = (size_t)rows * width * sizeof(T);
    T *d = reinterpret_cast<T *>(buf + off);
    off += (bytes + 255) & ~(size_t)255;
    CUDA_TRY(cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, st));
    k_rows_scatter<T><<<G, 256, 0, st>>>(rows, width, h->rows_ids, id_base, d, dst);
    return 0;
  };
  if (put(h->x, (const double *)s->x, n, 2, 0)) return 1;
  {  // vel / stress: compact rows -> format B
    double *dv = reinterpret_cast<double *>(buf + off);
    off += ((size_t)n * 2 * 8 + 255) & ~(size_t)255;
    double *ds = reinterpret_cast<double *>(buf + off);
    off += ((size_t)n * 4 * 8 + 255) & ~(size_t)255;
    CUDA_TRY(cudaMemcpyAsync(dv, s->vel, (size_t)n * 2 * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(ds, s->stress, (size_t)n * 4 * 8, cudaMemcpyHostToDevice, st));
    h->cur = h->cur;  // rows go into the current format-B buffer set
    if (n > 0) k_rows_pack<<<(n + 255) / 256, 256, 0, st>>>(h->P, n, h->rows_ids, dv, ds, state_ptrs(h, h->cur));
  }
  if (s->internal_vars && nt > 0) {
    if (put(h->ivars, (const double *)s->internal_vars, nt, SPSPH_NINT_VARS, 0)) return 1;
    k_rows_epsp<<<(nt + 255) / 256, 256, 0, st>>>(nt, h->rows_ids, p.ntotal, h->ivars, h->epsp);
  }
  if (put(h->fdp, (const double *)s->f_drucker, nt, 1, 0)) return 1;
  if (put(h->displ, (const double *)s->displ, nn, 2, 0)) return 1;
  if (put(h->x_10, (const double *)s->x_10, nn, 2, 0)) return 1;
  if (put(h->disp_10, (const double *)s->disp_10, nn, 1, 0)) return 1;
  if (put(h->n_int, (const float *)s->n_int, nn, 1, 0)) return 1;
  if (put(h->bc_int, (const int *)s->bc_int, nn, 1, 0)) return 1;
  if (put(h->if_out, (const int *)s->if_out_domain, n, 1, 0)) return 1;
  if (put(h->bc_or_not, (const int *)s->bc_or_not, nt, 1, 0)) return 1;
  CUDA_TRY(cudaGetLastError());
  h->m_pairs = 0;
  h->have_lists = false;
  h->tile_last = false;
  h->x_fs_valid = false;
#ifndef SPSPH_HOST_EMU
  if (h->dist) {  // the uploaded rows are this rank's local particles; everything else is remote
    CUDA_TRY(cudaMemsetAsync(h->lflag, 0, (size_t)p.ntotal2 * sizeof(int), st));
    if (n > 0) k_dist_flags_rows<<<(n + 255) / 256, 256, 0, st>>>(h->P, h->D, h->x, h->rows_ids, n, h->lflag);
    if (rebuild_local_list(h)) return 1;
  }
#endif
  CUDA_TRY(cudaStreamSynchronize(st));  // the caller may reuse its arrays as soon as the call returns
  return 0;
}

int spsph_download_rows(spsph_handle *h, const spsph_state *s, const int32_t *ids, int32_t n) {
  if (!h || !s) return 1;
  CUDA_TRY(cudaSetDevice(h->device));
  int nn = 0, nt = 0;
  if (rows_prepare(h, ids, n, &nn, &nt)) return 1;
  cudaStream_t st = h->stream;
  char *buf = h->rows_buf;
  const int G = 148 * 8;
  size_t off = 0;
  auto get = [&](auto *dst, const auto *src, int rows, int width) -> int {
    using T = std::remove_cv_t<std::remove_pointer_t<decltype(src)>>;
    if (!dst || rows == 0) return 0;
    const size_t bytes = (size_t)rows * width * sizeof(T);
    T *d = reinterpret_cast<T *>(buf + off);
    off += (bytes + 255) & ~(size_t)255;
    k_rows_gather<T><<<G, 256, 0, st>>>(rows, width, h->rows_ids, 0, src, d);
    CUDA_TRY(cudaMemcpyAsync(dst, d, bytes, cudaMemcpyDeviceToHost, st));
    return 0;
  };
  if (get(s->x, (const double *)h->x, n, 2)) return 1;
  if ((s->vel || s->stress) && n > 0) {
    double *dv = reinterpret_cast<double *>(buf + off);
    off += ((size_t)n * 2 * 8 + 255) & ~(size_t)255;
    double *ds = reinterpret_cast<double *>(buf + off);
    off += ((size_t)n * 4 * 8 + 255) & ~(size_t)255;
    k_rows_unpack<<<(n + 255) / 256, 256, 0, st>>>(h->P, n, h->rows_ids, state_ptrs(h, h->cur), s->vel ? dv : nullptr,
                                                   s->stress ? ds : nullptr);
    if (s->vel) CUDA_TRY(cudaMemcpyAsync(s->vel, dv, (size_t)n * 2 * 8, cudaMemcpyDeviceToHost, st));
    if (s->stress) CUDA_TRY(cudaMemcpyAsync(s->stress, ds, (size_t)n * 4 * 8, cudaMemcpyDeviceToHost, st));
  }
  if (s->internal_vars && nt > 0) {
    k_download_ivars<<<(h->hp.ntotal + 255) / 256, 256, 0, st>>>(h->hp.ntotal, h->epsp, h->ivars);
    if (get(s->internal_vars, (const double *)h->ivars, nt, SPSPH_NINT_VARS)) return 1;
  }
  if (get(s->f_drucker, (const double *)h->fdp, nt, 1)) return 1;
  if (get(s->displ, (const double *)h->displ, nn, 2)) return 1;
  if (get(s->x_10, (const double *)h->x_10, nn, 2)) return 1;
  if (get(s->disp_10, (const double *)h->disp_10, nn, 1)) return 1;
  if (get(s->n_int, (const float *)h->n_int, nn, 1)) return 1;
  if (get(s->bc_int, (const int *)h->bc_int, nn, 1)) return 1;
  if (get(s->if_out_domain, (const int *)h->if_out, n, 1)) return 1;
  if (get(s->bc_or_not, (const int *)h->bc_or_not, nt, 1)) return 1;  // (as stored: no free-surface pass here)
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

int spsph_sync(spsph_handle *h) {
  if (!h) return 1;
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

// get_nodes_on_free_surface (main:152-154 runs it at the end of every step; only bc_or_not leaves it): evaluated when
// somebody asks for the marks, with the last step's pair lists, unless the configuration needs it every step anyway
static int free_surface_on_demand(spsph_handle *h) {
  if (!h->hp.update_x || h->fs_each_step) return 0;
  if (materialize_lists(h)) return 1;
  if (!h->have_lists) return 0;
  SlotMap ML = h->M;
  ML.nn = h->nloc[0];
  ML.ns = h->nloc[1];
  ML.nd = h->nloc[2];
  const int T = ML.nnp + ML.nsp;
  k_free_surface<<<(T + 127) / 128, 128, 0, h->stream>>>(h->P, ML, sort_arrays(h), h->pos_of, h->L, h->n0, h->n1,
                                                         h->growth, h->x_fs_valid ? h->x_fs : h->x, h->mass, h->rho,
                                                         h->hsml, h->bc_or_not, nullptr);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int spsph_download_frame(spsph_handle *h, const int32_t *cols, int32_t ncols, int32_t first, int32_t count,
                         double *out) {
  if (!h) return 1;
  if (!h->uploaded) {
    h->err = "spsph_download_frame: nothing uploaded yet";
    return 1;
  }
  if (!cols || ncols < 1 || ncols > SPSPH_FRAME_MAX_COLS) {
    h->err = "spsph_download_frame: 1 .. SPSPH_FRAME_MAX_COLS column codes are required";
    return 1;
  }
  if (first < 0 || count < 0 || (long long)first + count > h->hp.ntotal2 || (count > 0 && !out)) {
    h->err = "spsph_download_frame: particle range outside [0, ntotal2) or no output table";
    return 1;
  }
  static_assert(SPSPH_COL_COUNT <= 16 && SPSPH_FRAME_MAX_COLS <= 16, "FrameCols packs 4 bits per column");
  FrameCols C;
  C.n = ncols;
  C.codes = 0;
  bool marks = false;
  for (int k = 0; k < ncols; ++k) {
    if (cols[k] < 0 || cols[k] >= SPSPH_COL_COUNT) {
      h->err = "spsph_download_frame: unknown column code";
      return 1;
    }
    C.codes |= (unsigned long long)cols[k] << (4 * k);
    marks |= cols[k] == SPSPH_COL_BC_OR_NOT;
  }
  CUDA_TRY(cudaSetDevice(h->device));
  if (count == 0) return 0;
  if (marks && free_surface_on_demand(h)) return 1;
  const long long nelem = (long long)count * ncols;
  const size_t bytes = (size_t)nelem * sizeof(double);
  if (bytes > h->frame_cap) {
    cudaFree(h->frame_buf);
    h->frame_buf = nullptr;
    h->frame_cap = 0;
    CUDA_TRY(cudaMalloc((void **)&h->frame_buf, bytes + bytes / 8));
    h->frame_cap = bytes + bytes / 8;
  }
  cudaStream_t st = h->stream;
  k_pack_frame<<<(unsigned)((nelem + 255) / 256), 256, 0, st>>>(h->P, state_ptrs(h, h->cur), C, h->displ, h->disp_10,
                                                                first, nelem, h->frame_buf);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(out, h->frame_buf, bytes, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

int spsph_download(spsph_handle *h, const spsph_state *s) {
  if (!h || !s) return 1;
  const spsph_params &p = h->hp;
  const size_t n2 = (size_t)p.ntotal2, nt = (size_t)p.ntotal, nn = (size_t)p.nnode;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  auto down = [&](void *dst, const void *src, size_t bytes) {
    if (!dst) return cudaSuccess;
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
  };
  CUDA_TRY(down(s->x, h->x, 2 * n2 * 8));
  if (s->vel || s->stress)
    k_unpack_state<<<((int)nt + 255) / 256, 256, 0, st>>>(h->P, state_ptrs(h, h->cur), h->stage_vel, h->stage_stress);
  if (s->vel) {  // dummy particles carry no velocity/stress on the device (zero, as after main:690)
    std::memset(s->vel + 2 * nt, 0, 2 * (n2 - nt) * 8);
    CUDA_TRY(down(s->vel, h->stage_vel, 2 * nt * 8));
  }
  if (s->stress) {
    std::memset(s->stress + 4 * nt, 0, 4 * (n2 - nt) * 8);
    CUDA_TRY(down(s->stress, h->stage_stress, 4 * nt * 8));
  }
  CUDA_TRY(down(s->rho, h->rho, n2 * 8));
  CUDA_TRY(down(s->mass, h->mass, n2 * 8));
  CUDA_TRY(down(s->hsml, h->hsml, n2 * 8));
  if (s->internal_vars) {
    k_download_ivars<<<((int)nt + 255) / 256, 256, 0, st>>>((int)nt, h->epsp, h->ivars);
    CUDA_TRY(down(s->internal_vars, h->ivars, (size_t)SPSPH_NINT_VARS * nt * 8));
  }
  if (s->bc_or_not && free_surface_on_demand(h)) return 1;
  CUDA_TRY(down(s->f_drucker, h->fdp, nt * 8));
  CUDA_TRY(down(s->x00, h->x00, 2 * n2 * 8));
  CUDA_TRY(down(s->displ, h->displ, 2 * nn * 8));
  CUDA_TRY(down(s->x_10, h->x_10, 2 * nn * 8));
  CUDA_TRY(down(s->disp_10, h->disp_10, nn * 8));
  CUDA_TRY(down(s->n_int, h->n_int, nn * 4));
  CUDA_TRY(down(s->bc_int, h->bc_int, nn * 4));
  CUDA_TRY(down(s->if_out_domain, h->if_out, n2 * 4));
  CUDA_TRY(down(s->bc_or_not, h->bc_or_not, nt * 4));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (s->itype) std::memcpy(s->itype, h->h_itype.data(), n2 * sizeof(int32_t));
  return 0;
}

int spsph_pair_stats(spsph_handle *h, int64_t *npairs, int32_t *maxiac, int32_t *miniac, int32_t *noiac) {
  if (!h) return 1;
  CUDA_TRY(cudaSetDevice(h->device));
  const int init[4] = {0, 1000, 0, 0};
  CUDA_TRY(cudaMemcpyAsync(h->stats_d, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
  SlotMap M = h->M;  // only the slots of this rank's local particles were written by the last k_count
  M.nn = h->nloc[0];
  M.ns = h->nloc[1];
  M.nd = h->nloc[2];
#ifndef SPSPH_EMU_SERIAL
  k_pair_stats<<<148, 256, 0, h->stream>>>(M, h->nall, h->stats_d);
#else
  k_pair_stats<<<1, 1, 0, h->stream>>>(M, h->nall, h->stats_d);  // one thread sees everything: no warp reduction needed
#endif
  int out[4];
  CUDA_TRY(cudaMemcpyAsync(out, h->stats_d, sizeof(out), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (npairs) *npairs = h->last_n_pairs;
  if (maxiac) *maxiac = out[0];
  if (miniac) *miniac = out[1];
  if (noiac) *noiac = out[2];
  return 0;
}

int spsph_pairs(spsph_handle *h, int64_t *npairs, int32_t *pair_i, int32_t *pair_j, int32_t *pint_type, float *w,
                float *dwdx, float *dwdy) {
  if (!h) return 1;
  if (h->dist) {  // the creation indices of a slab run are rank-local: there is no global ordered list to export
    h->err = "spsph_pairs is not available in a multi-GPU run (use spsph_pair_stats for the global pair count)";
    return 1;
  }
  const long long n = h->last_n_pairs;
  if (npairs) *npairs = n;
  if (!pair_i) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  if (materialize_lists(h)) return 1;  // creation indices (base_u) come from the count pass of the list path
  void *d[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  const size_t b = (size_t)(n > 0 ? n : 1) * 4;
  auto body = [&]() -> int {
    for (int k = 0; k < 6; ++k) CUDA_TRY(cudaMalloc(&d[k], b));
    const int T = h->M.total();
    SlotMap ML = h->M;  // only the slots of this rank's local particles hold data
    ML.nn = h->nloc[0];
    ML.ns = h->nloc[1];
    ML.nd = h->nloc[2];
    // NB: valid until the next spsph_step (positions in the sorted arrays are those of the last search)
    k_export_pairs<<<(T + 127) / 128, 128, 0, h->stream>>>(h->P, ML, h->G, sort_arrays(h), h->base_u, n, h->last_m_before,
                                                           (int *)d[0], (int *)d[1], (int *)d[2], (float *)d[3],
                                                           (float *)d[4], (float *)d[5]);
    void *host[6] = {pair_i, pair_j, pint_type, w, dwdx, dwdy};
    for (int k = 0; k < 6; ++k)
      CUDA_TRY(cudaMemcpyAsync(host[k], d[k], (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
  };
  const int rc = body();
  for (int k = 0; k < 6; ++k) cudaFree(d[k]);
  return rc;
}

// the NCCL library: libnccl.so.2 from the loader's search path (torch's bundled copy when the caller imported torch),
// or the file SPSPH_NCCL_SO names (another NCCL build; the CPU tests put a file-based stand-in there)
static const char *nccl_library() {
  const char *e = std::getenv("SPSPH_NCCL_SO");
  return (e && *e) ? e : "libnccl.so.2";
}

int spsph_dist_unique_id(char *id128) {
  void *lib = dlopen(nccl_library(), RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return 1;
  auto f = (ncclResult_t(*)(ncclUniqueId *))dlsym(lib, "ncclGetUniqueId");
  if (!f) return 1;
  ncclUniqueId u;
  if (f(&u) != ncclSuccess) return 1;
  std::memcpy(id128, u.internal, 128);
  return 0;
}

int spsph_dist_init(spsph_handle *h, int32_t rank, int32_t nranks, const char *id128, const double *planes,
                    int32_t halo_cells, int32_t halo_capacity) {
  if (!h || !planes || nranks < 1 || rank < 0 || rank >= nranks) return 1;
  if (!h->uploaded) {
    h->err = "spsph_dist_init must follow spsph_upload (every rank uploads the complete problem)";
    return 1;
  }
  CUDA_TRY(cudaSetDevice(h->device));
  const spsph_params &p = h->hp;
  h->nccl_lib = dlopen(nccl_library(), RTLD_NOW | RTLD_GLOBAL);
  if (!h->nccl_lib) {
    h->err = std::string("cannot load ") + nccl_library() + ": " + dlerror();
    return 1;
  }
#define NCCL_SYM(name)                                                            \
  h->p_##name = (decltype(h->p_##name))dlsym(h->nccl_lib, #name);                \
  if (!h->p_##name) {                                                             \
    h->err = "libnccl.so.2 lacks " #name;                                         \
    return 1;                                                                     \
  }
  NCCL_SYM(ncclCommInitRank)
  NCCL_SYM(ncclCommDestroy)
  NCCL_SYM(ncclSend)
  NCCL_SYM(ncclRecv)
  NCCL_SYM(ncclAllReduce)
  NCCL_SYM(ncclAllGather)
  NCCL_SYM(ncclGroupStart)
  NCCL_SYM(ncclGroupEnd)
  NCCL_SYM(ncclGetErrorString)
#undef NCCL_SYM
  ncclUniqueId u;
  std::memcpy(u.internal, id128, 128);
  NCCL_TRY(h->p_ncclCommInitRank(&h->comm, nranks, u, rank));
  // halo distance: every dependent sweep of a step reads partners at most one cell (2*max h) away
  double hmax = 0.0;
  {
    std::vector<double> hs((size_t)p.ntotal2);
    CUDA_TRY(cudaMemcpy(hs.data(), h->hsml, hs.size() * 8, cudaMemcpyDeviceToHost));
    for (double v : hs) hmax = std::fmax(hmax, v);
  }
  DistGeom &D = h->D;
  D.rank = rank;
  D.nranks = nranks;
  D.lo = planes[rank];
  D.hi = planes[rank + 1];
  // every dependent sweep reads partners up to the kernel cut-off away: scale_k * h (2 h cubic spline, 3 h Gauss / quintic)
  D.H = (double)halo_cells * (double)h->P.scale_k * hmax;
  D.sp_follows_node = (!p.inside_approach) ? 1 : 0;  // outside approach and standard SPH
  D.cap = halo_capacity;
  // continuity density and per-step free-surface marks travel in an extended halo record
  D.ext = (p.cont_density ? HALO_EXT_DENSITY : 0) | (h->fs_each_step ? HALO_EXT_MARKS : 0);
  D.rec = D.ext ? HALO_REC_EXT : HALO_REC;
  if (nranks > 1 && (D.hi - D.lo) < D.H && rank > 0 && rank < nranks - 1) {
    h->err = "multi-GPU: slab thinner than the halo distance";
    return 1;
  }
  const size_t n2 = (size_t)p.ntotal2;
  const size_t msg = (size_t)D.rec * ((size_t)D.cap + 1);
  if (dalloc(h, &h->lflag, n2) || dalloc(h, &h->halo_cnt, 4)) return 1;
  if (dalloc(h, &h->list_ids[0], n2) || dalloc(h, &h->list_ids[1], n2) || dalloc(h, &h->list_n, 2) ||
      dalloc(h, &h->list_keep, n2) || dalloc(h, &h->list_pos, n2))
    return 1;
  if (dalloc(h, &h->gt_buf, (size_t)GT_CAP) || dalloc(h, &h->gt_mine, 2 * (size_t)GT_PCAP) ||
      dalloc(h, &h->gt_all, 2 * (size_t)GT_PCAP * nranks) || dalloc(h, &h->gt_out2, 4) || dalloc(h, &h->gt_sel, 1))
    return 1;
  for (int side = 0; side < 2; ++side)
    if (dalloc(h, &h->halo_ids[side], (size_t)D.cap) || dalloc(h, &h->halo_send[side], msg) ||
        dalloc(h, &h->halo_recv[side], msg))
      return 1;
  CUDA_TRY(cudaMemset(h->halo_cnt, 0, 4 * sizeof(int)));
  h->halo_cells = halo_cells;
  if (const char *e = std::getenv("SPSPH_HALO_SLACK")) h->halo_slack = std::max(0, std::atoi(e));
  {  // halo peeling (k_peel_counts): opt-in with SPSPH_PEEL=1 -- bit-identical (tests/test_dist_emulated_cpu.py, and the
     // parity check of bench.py --gpus 2 on hardware) but not faster where it could be measured: 1.79 against 1.76 ms per
     // step on two slabs of a 1 M-particle column (170-cell slabs, 13-cell halos)
    const char *e = std::getenv("SPSPH_PEEL");
    h->peel = nranks > 1 && halo_cells > 4 && D.H > 0.0 && e && std::atoi(e) != 0;
    if (h->peel) {
      h->peel_stride = (size_t)h->M.nnp + h->M.nsp;
      if (dalloc(h, &h->peel0, PEEL_LEVELS * h->peel_stride) || dalloc(h, &h->peel1, PEEL_LEVELS * h->peel_stride)) return 1;
    }
  }
  h->dist = true;
  k_dist_init_flags<<<((int)n2 + 255) / 256, 256, 0, h->stream>>>(h->P, h->D, h->x, h->lflag);
  if (rebuild_local_list(h)) return 1;
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

int spsph_dist_set_planes(spsph_handle *h, const double *planes) {
  if (!h || !planes) return 1;
  if (!h->dist) {
    h->err = "spsph_dist_set_planes: not a multi-GPU run";
    return 1;
  }
  DistGeom &D = h->D;
  const double lo = planes[D.rank], hi = planes[D.rank + 1];
  // A particle whose slab changes must already sit in the new owner's halo (it travels with the next exchange, as
  // a migrant does): one call may move a plane by less than the halo distance, keeping a margin of two cells.
  const double lim = 0.5 * D.H;
  if ((D.rank > 0 && std::fabs(lo - D.lo) > lim) || (D.rank < D.nranks - 1 && std::fabs(hi - D.hi) > lim)) {
    h->err = "spsph_dist_set_planes: a slab plane may move by at most half the halo distance per call";
    return 1;
  }
  if (!(lo < hi) || (D.nranks > 2 && D.rank > 0 && D.rank < D.nranks - 1 && (hi - lo) < D.H)) {
    h->err = "spsph_dist_set_planes: slab thinner than the halo distance";
    return 1;
  }
  D.lo = lo;
  D.hi = hi;
  h->halo_prev_valid = false;  // the next messages also carry the particles that change slab: full-capacity transfers
  h->halo_full_msgs = 2;       // ... and the step after that too (see dist_status_apply)
  return 0;
}

int spsph_local_counts(spsph_handle *h, int32_t *nloc3) {
  if (!h || !nloc3) return 1;
  for (int k = 0; k < 3; ++k) nloc3[k] = h->nloc[k];
  return 0;
}

int spsph_dist_flags(spsph_handle *h, int32_t *flags) {
  if (!h || !flags) return 1;
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (!h->dist) {
    for (int i = 0; i < h->hp.ntotal2; ++i) flags[i] = 1;
    return 0;
  }
  CUDA_TRY(cudaMemcpy(flags, h->lflag, (size_t)h->hp.ntotal2 * sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

int spsph_destroy(spsph_handle *h) {
  if (!h) return 0;
  if (h->stream) {
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
  }
  if (h->comm && h->p_ncclCommDestroy) h->p_ncclCommDestroy(h->comm);
  for (void *q : h->allocs) cudaFree(q);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  cudaFree(h->L.idx0);
  cudaFree(h->L.w0);
  cudaFree(h->L.gx0);
  cudaFree(h->L.gy0);
  cudaFree(h->L.h0lo);
  cudaFree(h->L.h0hi);
  cudaFree(h->L.xC);
  cudaFree(h->L.yC);
  cudaFree(h->L.hC);
  cudaFree(h->L.idxC);
  cudaFree(h->L.wC);
  cudaFree(h->L.gxC);
  cudaFree(h->L.gyC);
  cudaFree(h->L.idxD);
  cudaFree(h->L.wD);
  cudaFree(h->TL.code0);
  cudaFree(h->TL.codeS);
  cudaFree(h->TL.w0);
  cudaFree(h->TL.gx0);
  cudaFree(h->TL.gy0);
  cudaFree(h->TL.gxC);
  cudaFree(h->TL.gyC);
  cudaFree(h->rows_buf);
  cudaFree(h->rows_ids);
  if (h->tstat_h) cudaFreeHost(h->tstat_h);
  if (h->status_h) cudaFreeHost(h->status_h);
  if (h->dist_h) cudaFreeHost(h->dist_h);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->stream2) cudaStreamDestroy(h->stream2);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

}  // extern "C"
