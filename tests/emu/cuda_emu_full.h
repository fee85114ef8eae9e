// cuda_emu_full.h — TEST INFRASTRUCTURE: host model of the CUDA execution for a -DB200_EMULATE build of the WHOLE
// library (gtsam_b200/csrc/engine.cu + kernels.cuh compiled by g++ into tests/emu/_build/libgtsam_b200_emu.so, see
// tests/test_library_emulation.py).  Included by kernels.cuh right after <cuda_runtime.h> when B200_EMULATE is defined.
//
//  * a launch runs the blocks of the grid one after the other (x fastest; ascending block ids are dependency-safe for
//    the one kernel whose blocks wait on each other through flags), and inside a block every CUDA thread is a FIBER
//    (its own stack, a 20-instruction context switch) of one host thread, scheduled round-robin: it runs until it finishes or reaches a barrier;
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
#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>
#include <map>
#include <utility>
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

// ---- cooperative scheduling: every CUDA thread of the running block is a fiber of ONE host thread -----------
// (a hand-written x86-64 context switch: swapcontext() saves and restores the signal mask with two system calls per
// switch, and a kernel with one barrier per pivot switches millions of times)
namespace b200_emu {
struct Ctx { void* rsp = nullptr; };
extern "C" void b200_emu_switch(Ctx* from, Ctx* to);
asm(R"(
.text
.globl b200_emu_switch
.type b200_emu_switch,@function
b200_emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  subq $8, %rsp
  stmxcsr (%rsp)
  fnstcw 4(%rsp)
  movq %rsp, (%rdi)
  movq (%rsi), %rsp
  ldmxcsr (%rsp)
  fldcw 4(%rsp)
  addq $8, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size b200_emu_switch, .-b200_emu_switch
)");
struct Fiber {
  Ctx ctx;
  uint3 tid;
  bool done = false;
  uint64_t xbuf = 0;     // 64-bit exchange slot (warp shuffles)
};
struct Sched {
  std::vector<Fiber> fibers;
  std::vector<char> stacks;
  Ctx main;
  unsigned nt = 0, cur = 0;
  int block_alive = 0, block_arrived = 0;
  unsigned block_gen = 0;
  std::vector<int> warp_alive, warp_arrived;
  std::vector<unsigned> warp_gen;
  std::vector<unsigned char> dyn_smem;
  std::vector<unsigned> order;
  void (*entry)(void*) = nullptr;
  void* entry_arg = nullptr;
};
inline Sched& sched() { static Sched s; return s; }
inline void yield() { Sched& s = sched(); b200_emu_switch(&s.fibers[s.cur].ctx, &s.main); }
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
  b200_emu_switch(&f.ctx, &s.main);   // never resumed
  __builtin_trap();
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
static inline int atomicMin(int* p, int v) { const int o = *p; if (v < o) *p = v; return o; }
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
          // initial frame: what b200_emu_switch pops (mxcsr / x87 control word, six callee-saved registers) and the
          // address it returns to; the stack pointer after that `ret` is 8 modulo 16, as at any function entry
          uintptr_t top = (uintptr_t)(s.stacks.data() + (size_t)(t + 1) * kStack);
          top &= ~(uintptr_t)15;
          uint64_t* sp = (uint64_t*)(top - 72);
          sp[0] = 0x037F00001F80ull;        // mxcsr = 0x1F80 (bytes 0-3), x87 control word = 0x037F (bytes 4-5)
          for (int q = 1; q <= 6; q++) sp[q] = 0;
          sp[7] = (uint64_t)(uintptr_t)(void (*)())fiber_main;
          sp[8] = 0;
          f.ctx.rsp = sp;
        }
        unsigned remaining = nt;
        // round robin: a fiber runs until it finishes or has to wait at a barrier.  B200_EMU_ORDER picks the order of a
        // sweep — ascending thread index (default), "reverse", or "shuffle[:seed]" (a new permutation every sweep): code
        // between two barriers must not depend on it, so a missing __syncthreads / __syncwarp (or reliance on warp
        // lock-step) shows up as a result that changes with the order
        static const int order_mode = [] { const char* e = getenv("B200_EMU_ORDER"); return !e ? 0 : !strncmp(e, "reverse", 7) ? 1 : !strncmp(e, "shuffle", 7) ? 2 : 0; }();
        static uint64_t rng = [] { const char* e = getenv("B200_EMU_ORDER"); const char* c = e ? strchr(e, ':') : nullptr; return c ? strtoull(c + 1, nullptr, 10) * 2654435761ull + 88172645463325252ull : 88172645463325252ull; }();
        if (order_mode && s.order.size() != nt) { s.order.resize(nt); for (unsigned t = 0; t < nt; t++) s.order[t] = t; }
        while (remaining) {
          remaining = 0;
          if (order_mode == 2)
            for (unsigned t = nt; t > 1; t--) {
              rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
              std::swap(s.order[t - 1], s.order[rng % t]);
            }
          for (unsigned k = 0; k < nt; k++) {
            const unsigned t = order_mode == 0 ? k : order_mode == 1 ? nt - 1 - k : s.order[k];
            if (s.fibers[t].done) continue;
            s.cur = t;
            b200_emu_switch(&s.main, &s.fibers[t].ctx);
            if (!s.fibers[t].done) remaining++;
          }
        }
      }
}
inline std::map<const void*, long>*& counts_ptr() { static std::map<const void*, long>* p = nullptr; return p; }
inline void* dyn_smem() { return sched().dyn_smem.data(); }
// kernel coverage of the emulated scenarios: with B200_EMU_TRACE_FILE set, every process appends "<mangled name> <launches>"
// lines at exit (tests/emu/kernel_coverage.py compares them with the kernels of the GPU build)
inline void count_launch(const void* fn) {
  static const char* path = getenv("B200_EMU_TRACE_FILE");
  if (!path) return;
  static std::map<const void*, long>* counts = [] {
    auto* m = new std::map<const void*, long>();
    atexit([] {
      FILE* f = fopen(getenv("B200_EMU_TRACE_FILE"), "a");
      if (!f) return;
      for (auto& kv : *counts_ptr()) {
        Dl_info info;
        fprintf(f, "%s %ld\n", dladdr(kv.first, &info) && info.dli_sname ? info.dli_sname : "?", kv.second);
      }
      fclose(f);
    });
    return m;
  }();
  counts_ptr() = counts;
  (*counts)[fn]++;
}
}  // namespace b200_emu
