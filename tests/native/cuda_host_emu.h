// TEST INFRASTRUCTURE. Serial host emulation of the CUDA device side for tests/native/*_host.cpp: every kernel
// thread runs to completion on the host, one after the other (emu_launch), so only kernels whose threads do not
// cooperate can be emulated -- the per-particle sweeps and list-walking kernels of csrc/step_kernels.cuh qualify (their
// warp-level operations only size loops or pick between arithmetically identical code paths); the neighbour build
// (block scans, shared-memory queues, warp reductions) does not and stays behind `#ifndef SPSPH_HOST_EMU`.
// Device memory is host memory, streams and events are no-ops.
#pragma once
#define SPSPH_HOST_EMU 1
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include <cuda_runtime.h>  // types only (double2, float2, dim3, cudaError_t, ...); no CUDA library is linked

struct EmuDim {
  unsigned x = 1, y = 1, z = 1;
};
static thread_local EmuDim emu_blockIdx, emu_blockDim, emu_threadIdx, emu_gridDim;
#define blockIdx emu_blockIdx
#define blockDim emu_blockDim
#define threadIdx emu_threadIdx
#define gridDim emu_gridDim
#undef __launch_bounds__
#define __launch_bounds__(...)
#define __shared__ static

// ---- intrinsics -------------------------------------------------------------------------------------------------
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __fsqrt_rn(float x) { return sqrtf(x); }
static inline double __drcp_rn(double x) { return 1.0 / x; }
static inline float __int_as_float(int v) {
  float f;
  std::memcpy(&f, &v, 4);
  return f;
}
static inline int __float_as_int(float f) {
  int v;
  std::memcpy(&v, &f, 4);
  return v;
}
static inline double __hiloint2double(int hi, int lo) {
  const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double d;
  std::memcpy(&d, &u, 8);
  return d;
}
static inline int __double2hiint(double d) {
  uint64_t u;
  std::memcpy(&u, &d, 8);
  return (int)(u >> 32);
}
static inline int __double2loint(double d) {
  uint64_t u;
  std::memcpy(&u, &d, 8);
  return (int)(u & 0xffffffffu);
}
template <class T>
static inline T __ldcs(const T *p) { return *p; }
template <class T>
static inline T __ldg(const T *p) { return *p; }
// one lane at a time: a warp-wide operation sees only the calling lane
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int, int = 32) { return v; }
template <class T>
static inline T __shfl_sync(unsigned, T v, int, int = 32) { return v; }
template <class T>
static inline T __shfl_down_sync(unsigned, T v, int, int = 32) { return v; }
template <class T>
static inline T __shfl_up_sync(unsigned, T v, int, int = 32) { return v; }
static inline int __any_sync(unsigned, int p) { return p != 0; }
static inline int __all_sync(unsigned, int p) { return p != 0; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __syncthreads() {}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
template <class T>
static inline T atomicAdd(T *p, T v) {
  const T o = *p;
  *p = o + v;
  return o;
}
template <class T>
static inline T atomicMax(T *p, T v) {
  const T o = *p;
  *p = std::max(o, v);
  return o;
}
template <class T>
static inline T atomicMin(T *p, T v) {
  const T o = *p;
  *p = std::min(o, v);
  return o;
}
using std::max;
using std::min;

// ---- kernel launch: k<<<grid, block, smem, stream>>>(args) is rewritten to emu_launch(grid, block, k, args) ----------
template <class K, class... A>
static inline void emu_launch(unsigned grid, unsigned block, K kernel, A... args) {
  emu_gridDim = EmuDim{grid, 1, 1};
  emu_blockDim = EmuDim{block, 1, 1};
  for (unsigned b = 0; b < grid; ++b)
    for (unsigned t = 0; t < block; ++t) {
      emu_blockIdx = EmuDim{b, 0, 0};
      emu_threadIdx = EmuDim{t, 0, 0};
      kernel(args...);
    }
}

// ---- runtime: device memory is host memory ---------------------------------------------------------------------------
#ifdef SPSPH_EMU_RUNTIME
static inline cudaError_t emu_ok() { return cudaSuccess; }
#define cudaGetDeviceCount(p) (*(p) = 1, cudaSuccess)
#define cudaSetDevice(d) emu_ok()
#define cudaGetLastError() emu_ok()
#define cudaGetErrorString(e) "emulated"
// cudaMalloc does not initialise: SPSPH_EMU_POISON=1 fills fresh "device" memory with 0xFF bytes (NaN as a double or a
// float, -1 as an int) so that a kernel reading memory nobody wrote shows up in the parity tests (an initcheck)
template <class T>
static inline cudaError_t emu_malloc(T **p, size_t n) {
  static const bool poison = std::getenv("SPSPH_EMU_POISON") != nullptr;
  *p = (T *)std::malloc(n ? n : 1);
  if (*p) std::memset((void *)*p, poison ? 0xFF : 0, n ? n : 1);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
#define cudaMalloc(p, n) emu_malloc(p, n)
#define cudaMallocHost(p, n) emu_malloc(p, n)
#define cudaFree(p) (std::free((void *)(p)), cudaSuccess)
#define cudaFreeHost(p) (std::free((void *)(p)), cudaSuccess)
#define cudaMemcpy(d, s, n, k) (std::memmove((void *)(d), (const void *)(s), n), cudaSuccess)
#define cudaMemcpyAsync(d, s, n, k, st) (std::memmove((void *)(d), (const void *)(s), n), cudaSuccess)
#define cudaMemset(d, v, n) (std::memset((void *)(d), v, n), cudaSuccess)
#define cudaMemsetAsync(d, v, n, st) (std::memset((void *)(d), v, n), cudaSuccess)
#define cudaStreamCreateWithFlags(s, f) (*(s) = nullptr, cudaSuccess)
#define cudaStreamDestroy(s) emu_ok()
#define cudaStreamSynchronize(s) emu_ok()
#define cudaStreamWaitEvent(s, e, f) emu_ok()
#define cudaEventCreate(e) (*(e) = nullptr, cudaSuccess)
#define cudaEventCreateWithFlags(e, f) (*(e) = nullptr, cudaSuccess)
#define cudaEventDestroy(e) emu_ok()
#define cudaEventRecord(e, s) emu_ok()
#define cudaEventSynchronize(e) emu_ok()
#define cudaEventElapsedTime(ms, a, b) (*(ms) = 0.f, cudaSuccess)
#endif
