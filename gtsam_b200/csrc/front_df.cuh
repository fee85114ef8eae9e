// front_df.cuh — dense partial Cholesky of every non-leaf front in ONE launch, as a tile dataflow.
//
// Replaces, per front, gtsam::choleskyPartial (gtsam/base/cholesky.cpp:107-158: LLT of the frontal block,
// TRSM of the row panel, SYRK of the trailing block), the split into the conditional
// (gtsam/linear/HessianFactor.cpp:459-487) and the extend-add of the Schur complement into the parent
// (HessianFactor::updateHessian, HessianFactor.cpp:348-374) — and, across fronts, the leaves-to-root walk of
// gtsam/inference/ClusterTree-inst.h:218-265.
//
// Round 1 walked the tree level by level with two launches per 32 pivot columns (panel + update); the chain of
// kernel boundaries (25 us per 32 pivots) was 47-70 % of every workload at 0.1-7 % of the FP64 peak.  Here:
//
//  * a front (f pivots, n = f + s + 1 columns, col-major upper triangle, ld = n) is cut by ONE list of block
//    boundaries used for rows and columns alike: pivot blocks [0,32) [32,64) ... [32(K-1), f) (the last one may
//    be short), then trailing blocks [f, f+32) ... — no block mixes pivot and trailing rows;
//  * a TILE = 4 consecutive row blocks x 1 column block (128 x 32) of the upper trapezoid is owned by one CTA of
//    4 warps for the whole factorisation: warp w keeps its 32 x 32 block of (minus) the Schur complement in DMMA
//    accumulators (mma.sync.m8n8k4.f64, 32 FP64 registers per thread) — C never round-trips through memory;
//  * pivot step k: the CTA that holds the diagonal block factors it (one warp, matrix in registers, shuffles) and
//    publishes R_kk; every CTA holding rows of pivot block k solves its 32 x 32 piece X = R_kk^-T C (one warp,
//    lane = column), writes it to its final place in the front ([R S d] IS the published panel) and raises the
//    piece's flag; every tile below waits for the two pieces it needs (its rows' and its columns'), stages them in
//    shared memory and applies the rank-32 update on the tensor pipe;
//  * rows >= f are the Schur complement: added straight into the parent front (FP64 red.add), after which the CTA
//    bumps the parent's arrival counter; a parent's tiles start when all its children's tiles have arrived, so the
//    whole tree — every level — is one launch;
//  * CTAs take their tile from an atomic ticket: tiles are ordered (level, front, column block, row tile) and every
//    dependency points to a smaller ticket, i.e. to a CTA that is already resident or finished, so the spin-waits
//    cannot deadlock whatever the grid size; waits are bounded and a timeout aborts the whole launch (fail flag),
//    never hangs the GPU.
//
// Memory model: a piece is written with plain stores, then __syncthreads, __threadfence and a release store of
// its flag by one thread; consumers poll with ld.acquire.gpu, __syncthreads, and read the piece through L2
// (cp.async / ld.global.cg).  Un-final front entries are only ever read with ld.global.cg (never allocated in L1),
// so an L1 line can only hold final values.
#pragma once

namespace b200 {

constexpr int kDfB = 32;        // block size (rows and columns)
constexpr int kDfTR = 4;        // row blocks per tile = warps per CTA
constexpr int kDfLd = 36;       // staged piece: [column][pivot row], 36 doubles per column (conflict-free DMMA fragments)
constexpr int kDfThreads = 32 * kDfTR;
constexpr int kDfSpinLimit = 1 << 21;   // x ~100 ns of back-off: ~0.3 s, then abort
constexpr int kDfLdR = kDfB + 2;   // R_kk rows 16-byte aligned (128-bit broadcast loads in the TRSM / Cholesky)
constexpr int kDfSmemBytes = ((1 + kDfTR) * kDfB * kDfLd + kDfB * (kDfB + 2) + kDfB * kDfLdR + kDfB) * 8;   // 63 744 B: 3 CTAs per SM

struct DfView {
  const int4* tasks;       // (front, column block, row tile, unused), ticket order
  int ntasks;
  int* ctrl;               // [0] ticket of this launch, [1] abort
  int* flags;              // piece flags, per front K x NB
  const int* flag_off;     // per clique: offset into flags
  int* done;               // per clique: tiles of children that finished their extend-add
  const int* expect;       // per clique: how many arrivals to wait for before loading
  unsigned long long* trace;   // B200_DF_TRACE: 32 globaltimer stamps per task (nullptr: off)
};

#if defined(B200_EMULATE)
#define DF_STAMP(code) do {} while (0)
#else
__device__ __forceinline__ unsigned long long df_now() { unsigned long long x; asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(x)); return x; }
// stamp = (event code << 56) | ns, appended to the task's row of the trace (thread 0 only)
#define DF_STAMP(code) do { if (v.trace && tid == 0 && tr_n < 32) v.trace[(size_t)s_task * 32 + tr_n++] = ((unsigned long long)(code) << 56) | (df_now() & 0x00ffffffffffffffull); } while (0)
#endif

#ifdef B200_EMULATE
__device__ __forceinline__ int df_ld_acquire(const int* p) { return *(const volatile int*)p; }
__device__ __forceinline__ void df_st_release(int* p, int v) { *(volatile int*)p = v; }
__device__ __forceinline__ void df_backoff() {}
#else
__device__ __forceinline__ int df_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void df_st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void df_backoff() { __nanosleep(32); }
#endif
#ifdef B200_EMULATE
__device__ __forceinline__ int df_ld_relaxed(const int* p) { return *(const volatile int*)p; }
__device__ __forceinline__ void df_fence_acquire() {}
#else
__device__ __forceinline__ int df_ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void df_fence_acquire() { asm volatile("fence.acq_rel.gpu;\n" ::: "memory"); }
#endif

// bounded wait for *p >= target; false when the launch is being aborted
__device__ __forceinline__ bool df_wait(const int* p, int target, int* ctrl) {
  int spins = 0;
  while (df_ld_relaxed(p) < target) {     // relaxed polls, ONE acquire fence once the value is there
    df_backoff();
    if ((++spins & 255) == 0) {
      if (df_ld_relaxed(ctrl + 1) != 0) return false;
      if (spins >= kDfSpinLimit) { atomicExch(ctrl + 1, 1); return false; }
    }
  }
  df_fence_acquire();
  return true;
}

__device__ __forceinline__ int df_beg(int b, int f, int K) { return b < K ? kDfB * b : f + kDfB * (b - K); }
__device__ __forceinline__ int df_size(int b, int f, int n, int K) {
  const int beg = df_beg(b, f, K), lim = b < K ? f : n;
  return min(kDfB, lim - beg);
}

// stage piece (pivot block rows [r0, r0+nr), columns [c0, c0+nc)) of the front into dst[x * kDfLd + p]; zero padded
__device__ __forceinline__ void df_stage(double* dst, const double* M, int n, int r0, int nr, int c0, int nc, int tid, int nthreads) {
  for (int e = tid; e < kDfB * kDfB; e += nthreads) {
    const int p = e & 31, x = e >> 5;
    double* d = dst + x * kDfLd + p;
    if (p < nr && x < nc) cp_async8(d, M + (r0 + p) + (size_t)(c0 + x) * n);
    else *d = 0.0;
  }
}

__global__ void __launch_bounds__(kDfThreads, 3)
front_df_kernel(TreeView t, DfView v, Scalars* sc) {
  pdl_sync();
  B200_DYN_SMEM(double, df_smem);                                  // kDfSmemBytes, carved up:
  double* Pc = df_smem;                                            // column piece (k, j): the B operand, [column][pivot row]
  double (*Pr)[kDfB * kDfLd] = (double (*)[kDfB * kDfLd])(df_smem + kDfB * kDfLd);   // row pieces (k, 4r + w): the A operands
  double (*Dg)[kDfB + 1] = (double (*)[kDfB + 1])(df_smem + (1 + kDfTR) * kDfB * kDfLd);   // pivot rows out of the accumulators ((kDfB + 2) * kDfB reserved: keeps Rk 16-byte aligned)
  double (*Rk)[kDfLdR] = (double (*)[kDfLdR])(df_smem + (1 + kDfTR) * kDfB * kDfLd + kDfB * (kDfB + 2));   // R_kk
  double* invd = df_smem + (1 + kDfTR) * kDfB * kDfLd + kDfB * (kDfB + 2) + kDfB * kDfLdR;
  __shared__ int s_task, s_ok;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, q = lane & 3;
  if (tid == 0) { s_task = atomicAdd(v.ctrl, 1); s_ok = 1; }
  __syncthreads();
  if (s_task >= v.ntasks) return;
  int tr_n = 0; (void)tr_n;
  DF_STAMP(1);    // start
  const int4 task = v.tasks[s_task];
  const int c = task.x, j = task.y, r = task.z;
  const int f = t.nf[c], n = f + t.ns[c] + 1;
  const int K = (f + kDfB - 1) / kDfB;                    // pivot blocks
  const int NB = K + (n - f + kDfB - 1) / kDfB;           // all blocks
  double* M = t.arena + t.off[c];
  int* flags = v.flags + v.flag_off[c];                   // [k * NB + block]
  const int par = t.parent[c];
  // ---- children first: every tile of every child has added its Schur complement ----
  if (v.expect[c] > 0) {
    if (tid == 0 && !df_wait(v.done + c, v.expect[c], v.ctrl)) s_ok = 0;
    __syncthreads();
    if (!s_ok) return;
  }
  DF_STAMP(2);    // children arrived
  // ---- this warp's block of -C into the accumulators ----
  const int i = kDfTR * r + w;                            // this warp's row block
  const bool wvalid = i <= j && i < NB;
  const int rb = wvalid ? df_beg(i, f, K) : 0, rs = wvalid ? df_size(i, f, n, K) : 0;
  const int cb = df_beg(j, f, K), cs = df_size(j, f, n, K);
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int lr = 8 * a + g, lc = 8 * b + 2 * q + h;
        double val = 0.0;
        if (lr < rs && lc < cs && rb + lr <= cb + lc) val = -__ldcg(M + (rb + lr) + (size_t)(cb + lc) * n);
        acc[a][b][h] = val;
      }
  const int last_rb = min(min(kDfTR * r + kDfTR - 1, j), NB - 1);   // last row block of the tile
  const bool diag_tile = j < K && last_rb == j;           // holds the diagonal block of pivot column j: factored after the loop
  const int kend = min(K, last_rb + 1) - (diag_tile ? 1 : 0);
  for (int k = 0; k < kend; k++) {
    const int wb = k - kDfTR * r;                         // the warp that holds pivot block k (< 0: above the tile)
    const int kb = kDfB * k, ks = min(kDfB, f - kb);      // pivot rows
    if (wb >= 0) {
      // ================= pivot rows live in this tile: X = R_kk^-T C, publish piece (k, j) =================
      if (tid == 0 && !df_wait(flags + k * NB + k, 1, v.ctrl)) s_ok = 0;
      if (w > wb && wvalid && i != j && lane == 0 && !df_wait(flags + k * NB + i, 1, v.ctrl)) s_ok = 0;
      if (w == wb) {
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int b = 0; b < 4; b++)
#pragma unroll
            for (int h = 0; h < 2; h++) Dg[8 * a + g][8 * b + 2 * q + h] = -acc[a][b][h];
      }
      __syncthreads();
      if (!s_ok) return;
      DF_STAMP(3);
      for (int e = tid; e < kDfB * kDfB; e += kDfThreads) {       // R_kk (upper), identity beyond the pivots
        const int p = e & 31, x = e >> 5;
        Rk[p][x] = (p <= x && x < ks) ? __ldcg(M + (kb + p) + (size_t)(kb + x) * n) : ((p == x) ? 1.0 : 0.0);
      }
      if (w > wb && wvalid && i != j) df_stage(Pr[w], M, n, kb, ks, rb, rs, lane, 32);
      cp_async_commit();
      __syncthreads();
      if (tid < kDfB) invd[tid] = 1.0 / Rk[tid][tid];
      __syncthreads();
      DF_STAMP(4);
      if (w == wb) {
        double x[kDfB];
#pragma unroll
        for (int p = 0; p < kDfB; p++) x[p] = (p < ks && lane < cs) ? Dg[p][lane] : 0.0;
        // forward substitution, right-looking: once x[qq] is final every later row takes its term — independent
        // FMAs (the left-looking dot products were one 496-long dependent chain: 2.5 us per piece on the B200)
#pragma unroll
        for (int qq = 0; qq < kDfB; qq++) {
          x[qq] *= invd[qq];
#pragma unroll
          for (int p = qq + 1; p < kDfB; p++) x[p] -= Rk[qq][p] * x[qq];
        }
#pragma unroll
        for (int p = 0; p < kDfB; p++) {
          Pc[lane * kDfLd + p] = x[p];
          if (p < ks && lane < cs) M[(kb + p) + (size_t)(cb + lane) * n] = x[p];
        }
      }
      cp_async_wait<0>();
      __syncthreads();
      DF_STAMP(5);
      if (tid == 0) df_st_release(flags + k * NB + j, 1);   // release = fence + store
      DF_STAMP(6);
    } else {
      // ================= pivot block above the tile: fetch the two pieces =================
      if (tid == 0 && !df_wait(flags + k * NB + j, 1, v.ctrl)) s_ok = 0;
      if (wvalid && i != j && lane == 0 && !df_wait(flags + k * NB + i, 1, v.ctrl)) s_ok = 0;
      __syncthreads();
      if (!s_ok) return;
      DF_STAMP(7);
      df_stage(Pc, M, n, kb, ks, cb, cs, tid, kDfThreads);
      if (wvalid && i != j) df_stage(Pr[w], M, n, kb, ks, rb, rs, lane, 32);
      cp_async_commit();
      cp_async_wait<0>();
      __syncthreads();
      DF_STAMP(8);
    }
    // ================= rank-32 update of the rows below the pivot block: -C += P_r^T P_c =================
    if (wvalid && i > k) {
      const double* A = (i == j) ? Pc : Pr[w];
#pragma unroll
      for (int k4 = 0; k4 < kDfB; k4 += 4) {
        double af[4], bf[4];
#pragma unroll
        for (int a = 0; a < 4; a++) af[a] = A[(8 * a + g) * kDfLd + k4 + q];
#pragma unroll
        for (int b = 0; b < 4; b++) bf[b] = Pc[(8 * b + g) * kDfLd + k4 + q];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int b = 0; b < 4; b++) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
      }
    }
    __syncthreads();   // the staged pieces are overwritten by the next step
    DF_STAMP(9);
  }
  if (diag_tile) {
    const int k = j, wb = k - kDfTR * r, kb = kDfB * k, ks = min(kDfB, f - kb);
      // ================= diagonal block: Cholesky by one warp, publish R_kk =================
      DF_STAMP(10);
      if (w == wb) {
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int b = 0; b < 4; b++)
#pragma unroll
            for (int h = 0; h < 2; h++) Dg[8 * a + g][8 * b + 2 * q + h] = -acc[a][b][h];
        __syncwarp();
        // Cholesky of the 32 x 32 block by ONE warp: lane j keeps column j in registers.  Step kk: the pivot comes from a
        // per-lane running diagonal (one shuffle), row kk is scaled and written to shared memory once, every lane then
        // subtracts r_ki r_kj with r_ki read as 128-bit broadcasts.  (Round 1 fetched every r_ki with its own pair of
        // shuffles under a predicate: 250 instructions per pivot, 10-13 us per block on the B200 — the critical path
        // of the whole solve.)  Entries below the diagonal are never read: they are left to accumulate garbage.
        double col[kDfB];
#pragma unroll
        for (int ii = 0; ii < kDfB; ii++)
          col[ii] = (lane < ks && ii <= lane) ? Dg[ii][lane] : ((ii == lane) ? 1.0 : 0.0);
        double dg = lane < ks ? Dg[lane][lane] : 1.0;
        double* rowbuf = &Rk[0][0];        // 2 x 32 doubles (R_kk staging is not in use in a diagonal tile)
        bool notpd = false;
#pragma unroll
        for (int kk = 0; kk < kDfB; kk++) {
          const double akk = __shfl_sync(0xffffffffu, dg, kk);
          if (kk < ks && !(akk > 0.0)) notpd = true;
          const double rinv = rsqrt(akk);                 // one rsqrt (<= 1 ulp) instead of sqrt + divide on the pivot chain
          const double rr = (lane == kk) ? akk * rinv : col[kk] * rinv;
          col[kk] = rr;
          dg -= rr * rr;                                  // the next pivots: no round trip through shared memory
          double* rb_ = rowbuf + (kk & 1) * kDfB;
          rb_[lane] = rr;
          __syncwarp();
#pragma unroll
          for (int ii = kk + 1; ii < kDfB; ii++) col[ii] -= rb_[ii] * rr;
        }
#pragma unroll
        for (int ii = 0; ii < kDfB; ii++) Dg[ii][lane] = (ii <= lane) ? col[ii] : 0.0;
        __syncwarp();
        if (lane == 0) {
          bool bad = notpd;     // (the pivots are broadcast: every lane saw the same)
          // the reference's underconstrained test on the last two pivots (gtsam/base/cholesky.cpp:144-157)
          if (k == K - 1) {
            if (f >= 2) {
              const double r2 = ks >= 2 ? Dg[ks - 2][ks - 2] : __ldcg(M + (f - 2) + (size_t)(f - 2) * n);
              if (!(dexp(r2) - dexp(Dg[ks - 1][ks - 1]) < 12)) bad = true;
            } else if (!(dexp(Dg[0][0]) > -12)) bad = true;
          }
          if (bad) atomicMax(&sc->fail_code, INT_MAX - c);
        }
        if (lane < ks) {
#pragma unroll
          for (int ii = 0; ii < kDfB; ii++)
            if (ii <= lane) M[(kb + ii) + (size_t)(kb + lane) * n] = col[ii];
        }
      }
      __syncthreads();
      DF_STAMP(11);
      if (tid == 0) df_st_release(flags + k * NB + k, 1);
      DF_STAMP(12);
    // (a pivot column has no trailing rows: nothing to extend-add)
  } else if (wvalid && i >= K) {
    // ---- trailing rows: the Schur complement goes straight into the parent (or stays, for a root) ----
    if (par >= 0) {
      double* P = t.arena + t.off[par];
      const int pn = t.nf[par] + t.ns[par] + 1;
      const int* map = t.ea_map + t.ea_ptr[c];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int lr = 8 * a + g, lc = 8 * b + 2 * q + h;
            if (lr < rs && lc < cs && rb + lr <= cb + lc) {
              const int pi = map[rb + lr - f], pj = map[cb + lc - f];
              const int lo = pi < pj ? pi : pj, hi = pi < pj ? pj : pi;
              atomicAdd(P + lo + (size_t)hi * pn, -acc[a][b][h]);
            }
          }
    } else {
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int lr = 8 * a + g, lc = 8 * b + 2 * q + h;
            if (lr < rs && lc < cs && rb + lr <= cb + lc) M[(rb + lr) + (size_t)(cb + lc) * n] = -acc[a][b][h];
          }
    }
  }
  DF_STAMP(13);
  if (par >= 0) {
    __syncthreads();
    if (tid == 0) { __threadfence(); atomicAdd(v.done + par, 1); }
  }
  DF_STAMP(14);
}

}  // namespace b200
