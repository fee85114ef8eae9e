// cuda_emu_lockstep.h — TEST INFRASTRUCTURE: host model of ONE-BLOCK-AT-A-TIME CUDA execution with real
// __syncthreads(): every thread of a block is a std::thread, __syncthreads() is a std::barrier, __shared__ arrays are
// function-local statics (one copy per block because blocks run one after the other).  Enough for block-cooperative
// kernels that use no warp shuffles (marginal_path_kernel, marginal_joint_kernel); see cuda_emu_prelude.h for the
// cheaper sequential model.
#pragma once
#include <barrier>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static
#define B200_JACOBIAN_MAX_ARITY 8

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local emu_dim3 threadIdx;
static emu_dim3 blockIdx, blockDim, gridDim;
static std::unique_ptr<std::barrier<>> emu_barrier;
static inline void __syncthreads() { emu_barrier->arrive_and_wait(); }

#define EMU_LAUNCH_LOCKSTEP(kernel, grid, block, ...)                                   \
  do {                                                                                  \
    gridDim.x = (grid); blockDim.x = (block);                                           \
    for (unsigned b_ = 0; b_ < gridDim.x; b_++) {                                       \
      blockIdx.x = b_;                                                                  \
      emu_barrier.reset(new std::barrier<>((std::ptrdiff_t)blockDim.x));                \
      std::vector<std::thread> pool_;                                                   \
      for (unsigned t_ = 0; t_ < blockDim.x; t_++)                                      \
        pool_.emplace_back([&, t_]() { threadIdx.x = t_; kernel(__VA_ARGS__); });       \
      for (auto& th_ : pool_) th_.join();                                               \
    }                                                                                   \
  } while (0)
