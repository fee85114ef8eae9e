// cuda_emu_full.h — TEST INFRASTRUCTURE: host model of the CUDA execution for a -DB200_EMULATE build of the WHOLE
// library (gtsam_b200/csrc/engine.cu + kernels.cuh compiled by g++ into tests/emu/_build/libgtsam_b200_emu.so, see
// tests/test_library_emulation.py).  Included by kernels.cuh right after <cuda_runtime.h> when B200_EMULATE is defined.
//
//  * a launch runs the blocks of the grid one after the other (x fastest; ascending block ids are dependency-safe for
//    the one kernel whose blocks wait on each other through flags), and inside a block every CUDA thread is a FIBER
//    (ucontext) of one host thread, scheduled round-robin: it runs until it finishes or reaches a barrier;
//    __syncthreads() is a block-wide barrier, __syncwarp() and the __shfl_*_sync exchanges a per-warp barrier, so
//    warp-synchronous code runs with its real data flow; a thread that returns drops out of both barriers (as exited
//    threads do on the device).  Deterministic: the same schedule, hence the same atomic order, every run;
//  * atomics are plain read-modify-writes (one host thread), cp.async is a synchronous copy, __shared__ is a
//    function-local static (one block at a time), dynamic shared memory one host buffer per launch;
//  * the CUDA runtime is tests/emu/cuda_fake_runtime.cpp: device memory = host memory, streams are synchronous.
// Not modelled: memory ordering subtleties, bank conflicts, occupancy, tensor instructions (the DMMA update kernel is
// never launched in this build), anything about speed.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <type_traits>
#include <vector>

#undef __global__
#undef __device__
#undef __host__
#undef __shared__
#undef __constant__
#undef __forceinline__
#undef __launch_bounds__
#undef __restrict__
#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __constant__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__

// ---- cooperative scheduling: every CUDA thread of the running block is a fiber (ucontext) of ONE host thread -----------
#include <ucontext.h>
namespace b200_emu {
struct Fiber {
  ucontext_t ctx;
  uint3 tid;
  bool done = false;
  uint64_t xbuf = 0;     // 64-bit exchange slot (warp shuffles)
};
struct Sched {
  std::vector<Fiber> fibers;
  std::vector<char> stacks;
  ucontext_t main;
  unsigned nt = 0, cur = 0;
  int block_alive = 0, block_arrived = 0;
  unsigned block_gen = 0;
  std::vector<int> warp_alive, warp_arrived;
  std::vector<unsigned> warp_gen;
  std::vector<unsigned char> dyn_smem;
  void (*entry)(void*) = nullptr;
  void* entry_arg = nullptr;
};
inline Sched& sched() { static Sched s; return s; }
inline void yield() { Sched& s = sched(); swapcontext(&s.fibers[s.cur].ctx, &s.main); }
inline void block_sync() {
  Sched& s = sched();
  const unsigned g = s.block_gen;
  if (++s.block_arrived == s.block_alive) { s.block_gen++; s.block_arrived = 0; return; }
  while (s.block_gen == g) yield();
}
inline void warp_sync() {
  Sched& s = sched();
  const unsigned w = s.fibers[s.cur].tid.x >> 5, g = s.warp_gen[w];
  if (++s.warp_arrived[w] == s.warp_alive[w]) { s.warp_gen[w]++; s.warp_arrived[w] = 0; return; }
  while (s.warp_gen[w] == g) yield();
}
inline void fiber_main() {
  Sched& s = sched();
  s.entry(s.entry_arg);
  Fiber& f = s.fibers[s.cur];
  f.done = true;       // an exited thread no longer takes part in any barrier: release the others if they all wait
  const unsigned w = f.tid.x >> 5;
  s.block_alive--; s.warp_alive[w]--;
  if (s.block_alive > 0 && s.block_arrived == s.block_alive) { s.block_gen++; s.block_arrived = 0; }
  if (s.warp_alive[w] > 0 && s.warp_arrived[w] == s.warp_alive[w]) { s.warp_gen[w]++; s.warp_arrived[w] = 0; }
  swapcontext(&f.ctx, &s.main);
}
}  // namespace b200_emu

#define threadIdx (b200_emu::sched().fibers[b200_emu::sched().cur].tid)
static uint3 blockIdx;
static dim3 blockDim, gridDim;

static inline void __syncthreads() { b200_emu::block_sync(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { b200_emu::warp_sync(); }
static inline void __threadfence() {}

template <class T>
static inline T emu_exchange(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  b200_emu::Sched& s = b200_emu::sched();
  const unsigned t = s.fibers[s.cur].tid.x;
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  s.fibers[t].xbuf = bits;
  b200_emu::warp_sync();
  const uint64_t got = s.fibers[(t & ~31u) | ((unsigned)src_lane & 31u)].xbuf;
  b200_emu::warp_sync();
  T r;
  memcpy(&r, &got, sizeof(T));
  return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int lane) { return emu_exchange(v, lane); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int mask) { return emu_exchange(v, (int)((threadIdx.x & 31u) ^ (unsigned)mask)); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned delta) {
  const unsigned lane = threadIdx.x & 31u;
  return emu_exchange(v, (int)(lane + delta < 32 ? lane + delta : lane));
}

template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *(const volatile T*)p; }

static inline double atomicAdd(double* p, double v) { const double o = *p; *p = o + v; return o; }
static inline int atomicAdd(int* p, int v) { const int o = *p; *p = o + v; return o; }
static inline int atomicMax(int* p, int v) { const int o = *p; if (v > o) *p = v; return o; }
static inline int atomicExch(int* p, int v) { const int o = *p; *p = v; return o; }
static inline unsigned atomicInc(unsigned* p, unsigned lim) { const unsigned o = *p; *p = o >= lim ? 0 : o + 1; return o; }

static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
using std::max;
using std::min;

namespace b200_emu {
// run one launch: fn() is the kernel call with its arguments bound
template <class F>
static void run(dim3 grid, dim3 block, size_t smem, bool descending_x, F&& fn) {
  constexpr size_t kStack = 256 << 10;
  gridDim = grid; blockDim = block;
  Sched& s = sched();
  const unsigned nt = block.x * block.y * block.z, nw = (nt + 31) / 32;
  s.nt = nt;
  if (s.fibers.size() < nt) s.fibers.resize(nt);
  if (s.stacks.size() < (size_t)nt * kStack) s.stacks.resize((size_t)nt * kStack);
  s.dyn_smem.assign(smem + 64, 0);
  s.entry = [](void* p) { (*static_cast<typename std::remove_reference<F>::type*>(p))(); };
  s.entry_arg = (void*)&fn;
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned k = 0; k < grid.x; k++) {
        const unsigned bx = descending_x ? grid.x - 1 - k : k;
        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
        s.block_alive = (int)nt; s.block_arrived = 0; s.block_gen = 0;
        s.warp_alive.assign(nw, 0); s.warp_arrived.assign(nw, 0); s.warp_gen.assign(nw, 0);
        for (unsigned t = 0; t < nt; t++) {
          Fiber& f = s.fibers[t];
          f.done = false; f.tid.x = t; f.tid.y = 0; f.tid.z = 0;
          s.warp_alive[t >> 5]++;
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = s.stacks.data() + (size_t)t * kStack;
          f.ctx.uc_stack.ss_size = kStack;
          f.ctx.uc_link = &s.main;
          makecontext(&f.ctx, (void (*)())fiber_main, 0);
        }
        unsigned remaining = nt;
        while (remaining) {      // round robin: a fiber runs until it finishes or has to wait at a barrier
          remaining = 0;
          for (unsigned t = 0; t < nt; t++) {
            if (s.fibers[t].done) continue;
            s.cur = t;
            swapcontext(&s.main, &s.fibers[t].ctx);
            if (!s.fibers[t].done) remaining++;
          }
        }
      }
}
inline void* dyn_smem() { return sched().dyn_smem.data(); }
}  // namespace b200_emu
