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
#ifndef SPSPH_EMU_SIMT
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
#else
// ---- SIMT mode (SPSPH_EMU_SIMT): the threads of a block are fibers (ucontext) that switch at every warp / block
// collective, so shuffles, votes, __syncwarp and __syncthreads have their real meaning; blocks still run one after
// the other. A collective is: deposit my value, warp barrier, read the other lanes, warp barrier.
#include <vector>
#if !defined(__x86_64__)
#error "the SIMT emulation switches fibers with a few lines of x86-64 assembly"
#endif
// fiber switch: callee-saved registers and the stack pointer (ucontext's swapcontext costs two signal-mask system
// calls per switch, which made a time step take 25 s)
extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch,.-emu_switch
)");
struct EmuBlock {
  int nthreads = 0, cur = -1;
  std::vector<void *> sp;  // saved stack pointer of every fiber
  std::vector<char> done;
  std::vector<char *> stack;
  void *sched_sp = nullptr;
  // warp barriers: arrivals and generation per warp; block barrier likewise
  int warr[64] = {0}, wgen[64] = {0}, walive[64] = {0};
  int barr = 0, bgen = 0, balive = 0;
  uint64_t wbuf[64][32];
};
static EmuBlock emu_blk;
static inline void emu_yield() { emu_switch(&emu_blk.sp[emu_blk.cur], emu_blk.sched_sp); }
static inline int emu_tid() { return emu_blk.cur; }
static inline void emu_warp_barrier() {
  const int w = emu_tid() >> 5;
  const int gen = emu_blk.wgen[w];
  if (++emu_blk.warr[w] >= emu_blk.walive[w]) {
    emu_blk.warr[w] = 0;
    ++emu_blk.wgen[w];
    return;
  }
  while (emu_blk.wgen[w] == gen) emu_yield();
}
static inline void emu_block_barrier() {
  const int gen = emu_blk.bgen;
  if (++emu_blk.barr >= emu_blk.balive) {
    emu_blk.barr = 0;
    ++emu_blk.bgen;
    return;
  }
  while (emu_blk.bgen == gen) emu_yield();
}
// a thread that returns leaves its warp and block: barriers waiting for it must be released
static inline void emu_thread_exit() {
  const int w = emu_tid() >> 5;
  --emu_blk.walive[w];
  --emu_blk.balive;
  if (emu_blk.walive[w] > 0 && emu_blk.warr[w] >= emu_blk.walive[w]) {
    emu_blk.warr[w] = 0;
    ++emu_blk.wgen[w];
  }
  if (emu_blk.balive > 0 && emu_blk.barr >= emu_blk.balive) {
    emu_blk.barr = 0;
    ++emu_blk.bgen;
  }
}
template <class T>
static inline uint64_t emu_bits(T v) {
  uint64_t u = 0;
  std::memcpy(&u, &v, sizeof(T));
  return u;
}
template <class T>
static inline T emu_from_bits(uint64_t u) {
  T v;
  std::memcpy(&v, &u, sizeof(T));
  return v;
}
// exchange: every lane deposits v; returns the value of lane `src` (own value if src is out of range or that lane has
// exited, as the hardware returns an undefined value there and the kernels never use it)
template <class T>
static inline T emu_exchange(T v, int src) {
  const int w = emu_tid() >> 5, lane = emu_tid() & 31;
  emu_blk.wbuf[w][lane] = emu_bits(v);
  emu_warp_barrier();
  const int t = (w << 5) + src;
  T r = v;
  if (src >= 0 && src < 32 && t < emu_blk.nthreads && !emu_blk.done[t]) r = emu_from_bits<T>(emu_blk.wbuf[w][src]);
  emu_warp_barrier();
  return r;
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emu_exchange(v, (emu_tid() & 31) ^ m); }
template <class T>
static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu_exchange(v, src & 31); }
template <class T>
static inline T __shfl_down_sync(unsigned, T v, int d, int = 32) {
  const int lane = emu_tid() & 31;
  return emu_exchange(v, lane + d < 32 ? lane + d : lane);
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, int d, int = 32) {
  const int lane = emu_tid() & 31;
  return emu_exchange(v, lane - d >= 0 ? lane - d : lane);
}
static inline unsigned __ballot_sync(unsigned, int p) {
  const int w = emu_tid() >> 5, lane = emu_tid() & 31;
  emu_blk.wbuf[w][lane] = p ? 1 : 0;
  emu_warp_barrier();
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) {
    const int t = (w << 5) + l;
    if (t < emu_blk.nthreads && !emu_blk.done[t] && emu_blk.wbuf[w][l]) m |= 1u << l;
  }
  emu_warp_barrier();
  return m;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, !p) == 0; }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp_barrier(); }
static inline void __syncthreads() { emu_block_barrier(); }
#endif
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
template <class T>
static inline T atomicOr(T *p, T v) {
  const T o = *p;
  *p = o | v;
  return o;
}
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

#ifndef SPSPH_EMU_SIMT
// ---- kernel launch: k<<<grid, block, smem, stream>>>(args) is rewritten to emu_launch(grid, block, k, args) ----------
template <class K, class... A>
static inline void emu_launch(dim3 grid, unsigned block, K kernel, A... args) {
  emu_gridDim = EmuDim{grid.x, grid.y, 1};
  emu_blockDim = EmuDim{block, 1, 1};
  for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned b = 0; b < grid.x; ++b)
      for (unsigned t = 0; t < block; ++t) {
        emu_blockIdx = EmuDim{b, by, 0};
        emu_threadIdx = EmuDim{t, 0, 0};
        kernel(args...);
      }
}

#else
// ---- kernel launch, SIMT mode: one fiber per thread of the block, round-robin between collectives --------------------
#include <functional>
static std::function<void()> emu_thread_body;
static void emu_fiber_entry() {
  emu_thread_body();
  emu_blk.done[emu_blk.cur] = 1;
  emu_thread_exit();
  emu_switch(&emu_blk.sp[emu_blk.cur], emu_blk.sched_sp);  // never resumed
  __builtin_trap();
}
static inline void emu_run_block(unsigned block) {
  constexpr size_t STACK = 256 * 1024;
  EmuBlock &B = emu_blk;
  B.nthreads = (int)block;
  if (B.sp.size() < block) {
    B.sp.resize(block);
    B.done.resize(block);
    while (B.stack.size() < block) B.stack.push_back((char *)std::malloc(STACK));
  }
  for (int w = 0; w < 64; ++w) B.warr[w] = B.wgen[w] = B.walive[w] = 0;
  B.barr = B.bgen = 0;
  B.balive = (int)block;
  for (unsigned t = 0; t < block; ++t) {
    B.done[t] = 0;
    ++B.walive[t >> 5];
    // initial frame: six callee-saved registers, the entry address emu_switch "returns" to, a fake return address
    // (the entry function then sees the stack alignment of an ordinary call)
    uintptr_t top = ((uintptr_t)B.stack[t] + STACK) & ~(uintptr_t)15;
    void **f = (void **)top;
    *--f = nullptr;
    *--f = (void *)&emu_fiber_entry;
    for (int r = 0; r < 6; ++r) *--f = nullptr;
    B.sp[t] = (void *)f;
  }
  int remaining = (int)block;
  while (remaining > 0) {
    remaining = 0;
    for (unsigned t = 0; t < block; ++t) {
      if (B.done[t]) continue;
      B.cur = (int)t;
      emu_threadIdx = EmuDim{t, 0, 0};
      emu_switch(&B.sched_sp, B.sp[t]);
      if (!B.done[t]) ++remaining;
    }
  }
}
template <class K, class... A>
static inline void emu_launch(dim3 grid, unsigned block, K kernel, A... args) {
  emu_gridDim = EmuDim{grid.x, grid.y, 1};
  emu_blockDim = EmuDim{block, 1, 1};
  emu_thread_body = [&]() { kernel(args...); };
  for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
      emu_blockIdx = EmuDim{bx, by, 0};
      emu_run_block(block);
    }
}
#endif

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
