// cuda_emu_prelude.h — TEST INFRASTRUCTURE: just enough of the CUDA execution model to run the library's
// simple one-thread-per-factor kernels ON THE HOST, verbatim (their text is extracted from
// gtsam_b200/csrc/kernels.cuh by tests/test_kernel_emulation.py and compiled after this prelude).
//
// Model: blocks run one after the other; inside a block the threads run one after the other in DESCENDING
// threadIdx order, each to completion.  That is faithful for kernels whose only intra-block communication is the
// block_sum / finish_sum reduction pair at the end (thread 0 runs last and sees the block total), which is all the
// emulated kernels use; atomics are plain read-modify-writes.  Warp shuffles, __syncthreads in the body, cp.async
// and tensor instructions are NOT modelled: kernels that use them are validated on hardware only.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static
#define B200_JACOBIAN_MAX_ARITY 8

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
static emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

static inline double atomicAdd(double* p, double v) { const double o = *p; *p += v; return o; }
static inline void pdl_sync() {}

// block reduction: result valid in thread 0 (= the last thread to run)
static double emu_block_acc[8];
static int emu_block_call;
template <int NT>
static inline double block_sum(double v, double*) {
  double& a = emu_block_acc[emu_block_call++];
  a += v;
  return threadIdx.x == 0 ? a : 0.0;
}
// deterministic cross-block sum: out = (accumulate ? out : 0) + sum over blocks of thread 0's block_value
static inline void finish_sum(double block_value, double*, unsigned*, double* out, int accumulate, double*) {
  if (threadIdx.x != 0) return;
  if (blockIdx.x == 0) *out = (accumulate ? *out : 0.0) + block_value;
  else *out += block_value;
}

#define EMU_LAUNCH(kernel, grid, block, ...)                                           \
  do {                                                                                 \
    gridDim.x = (grid); blockDim.x = (block);                                          \
    for (unsigned b_ = 0; b_ < gridDim.x; b_++) {                                      \
      blockIdx.x = b_;                                                                 \
      for (int k_ = 0; k_ < 8; k_++) emu_block_acc[k_] = 0.0;                          \
      for (int t_ = (int)blockDim.x - 1; t_ >= 0; t_--) {                              \
        threadIdx.x = (unsigned)t_;                                                    \
        emu_block_call = 0;                                                            \
        kernel(__VA_ARGS__);                                                           \
      }                                                                                \
    }                                                                                  \
  } while (0)
