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

#ifdef B200_EMULATE
#define B200_NOINLINE
#else
#define B200_NOINLINE __noinline__
#endif

namespace b200 {

constexpr int kDfB = 32;        // block size (rows and columns)
constexpr int kDfTR = 4;        // row blocks per tile = warps per CTA
constexpr int kDfLd = 36;       // staged piece: [column][pivot row], 36 doubles per column (conflict-free DMMA fragments)
constexpr int kDfThreads = 32 * kDfTR;
constexpr int kDfSpinLimit = 1 << 21;   // x ~100 ns of back-off: ~0.3 s, then abort
constexpr int kDfLdR = kDfB + 2;                                   // R_kk rows 16-byte aligned
// dynamic shared memory, in doubles: [Pc | Pr[4] | Dg 32 x 33 (+pad) | Rk 32 x 34 | invd 32 | rowbuf 2 x 32]
constexpr int kDfOffPr = kDfB * kDfLd, kDfOffDg = (1 + kDfTR) * kDfB * kDfLd, kDfOffRk = kDfOffDg + kDfB * (kDfB + 2),
              kDfOffInvd = kDfOffRk + kDfB * kDfLdR, kDfOffRow = kDfOffInvd + kDfB;
constexpr int kDfSmemBytes = (kDfOffRow + 2 * kDfB) * 8;           // 64 256 B: 3 CTAs per SM

struct DfView {
  const int4* tasks;       // (front, column block, row tile, unused), ticket order
  int ntasks;
  int* ctrl;               // [0] ticket of this launch, [1] abort
  int* flags;              // piece flags, per front K x NB
  const int* flag_off;     // per clique: offset into flags
  int* done;               // per clique: tiles of children that finished their extend-add
  const int* expect;       // per clique: how many arrivals to wait for before loading
  unsigned long long* trace;   // B200_DF_TRACE: 32 globaltimer stamps per task (nullptr: off)
  double* winv;            // inverses of the factored 32 x 32 diagonal blocks (back-substitution): per front winv_off + 1024 k
  const int64_t* winv_off;
  int warm_ctas;           // the first wave of CTAs (one per resident slot) warms the instruction caches before the grid dependency
};

#if defined(B200_EMULATE)
static inline long long clock64() { return 0; }
#define DF_STAMP(code) do {} while (0)
#define DF_CYC(code, cyc) do {} while (0)
#else
__device__ __forceinline__ unsigned long long df_now() { unsigned long long x; asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(x)); return x; }
// stamp = (event code << 56) | ns, appended to the task's row of the trace (thread 0 only)
#define DF_CYC(code, cyc) do { if (v.trace && lane == 0 && tr_n < 32) v.trace[(size_t)s_task * 32 + tr_n++] = ((unsigned long long)(code) << 56) | ((unsigned long long)(cyc) & 0x00ffffffffffffffull); } while (0)
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

// The two latency-critical routines of the pivot chain, each run by ONE warp on a 32 x 32 block in shared memory
// (D[row][col], lane = column) with the block in REGISTERS while it works.  Measured alone on the B200
// (profiles/micro/lat_bench.cu): Cholesky 6200 cycles (3.2 us), triangular solve 1870 cycles (0.95 us); dependent DFMA
// 8 cycles, dependent 64-bit shuffle 63 cycles, dependent DMMA 26 cycles.  Two things were tried and measured worse:
// round 1's Cholesky with one pair of shuffles per r_ki (250 instructions per pivot: 10-13 us), and versions with the
// block left in shared memory and rolled loops (compact code, but every term a load-FMA-store chain: 15 us each).
// The register versions are ~2500 / ~1000 fully unrolled instructions executed once per tile: inlined into the kernel
// they ran instruction-fetch bound (17 us for the first Cholesky after an L2 flush, 4.6 us after; 4.4 us per solve), so
// they are real functions (__noinline__: one copy of the code) and the first wave of CTAs runs both once on dummy
// data BEFORE griddepcontrol.wait — while the previous kernel of the stream is still draining — which pulls the code
// into L2 and into the instruction caches of every SM.
//
// Cholesky: the pivot comes from a per-lane running diagonal (one shuffle: the pivot chain never waits on shared
// memory), row kk is scaled and written to shared memory once, every lane subtracts r_ki r_kj from its column with
// r_ki read as 128-bit broadcasts.  Entries below the diagonal are never read (they accumulate garbage).  ks < 32: rows
// and columns >= ks are padding (identity).  Returns true when a pivot was not positive.
// (Both take the block as an OFFSET into the kernel's dynamic shared memory, not as a pointer: a pointer argument of a
// non-inlined function is generic, its loads can neither be LDS nor be merged into 128-bit broadcasts.)
template <int MINB>   // (one copy per kernel variant: each is compiled under that variant's register budget)
__device__ B200_NOINLINE bool df_chol32(int d_off, int ks, int lane) {
  B200_DYN_SMEM(double, df_smem);
  double (*D)[kDfB + 1] = (double (*)[kDfB + 1])(df_smem + d_off);
  double* rowbuf = df_smem + kDfOffRow;
  double col[kDfB];
#pragma unroll
  for (int ii = 0; ii < kDfB; ii++)
    col[ii] = (lane < ks && ii <= lane) ? D[ii][lane] : ((ii == lane) ? 1.0 : 0.0);
  double dg = lane < ks ? D[lane][lane] : 1.0;
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
  for (int ii = 0; ii < kDfB; ii++) D[ii][lane] = (ii <= lane) ? col[ii] : 0.0;
  __syncwarp();
  return notpd;
}

// X = R^-T C in place (D holds C on entry, X on exit; lane = column of the piece): forward substitution, right-looking —
// once x[qq] is final every later row takes its term, independent FMAs (the left-looking dot products are one 496-long
// dependent chain).  Rk[row][col] = R_kk with rows 16-byte aligned (kDfLdR), invd = 1 / diag.
template <int MINB>
__device__ B200_NOINLINE void df_trsm32(int d_off, int lane) {
  B200_DYN_SMEM(double, df_smem);
  double (*D)[kDfB + 1] = (double (*)[kDfB + 1])(df_smem + d_off);
  const double* Rk = df_smem + kDfOffRk;
  const double* invd = df_smem + kDfOffInvd;
  double x[kDfB];
#pragma unroll
  for (int p = 0; p < kDfB; p++) x[p] = D[p][lane];
#pragma unroll
  for (int qq = 0; qq < kDfB; qq++) {
    x[qq] *= invd[qq];
#pragma unroll
    for (int p = qq + 1; p < kDfB; p++) x[p] -= Rk[qq * kDfLdR + p] * x[qq];
  }
#pragma unroll
  for (int p = 0; p < kDfB; p++) D[p][lane] = x[p];
  __syncwarp();
}

// MINB = resident CTAs per SM the variant is compiled for: 3 (168 registers: more tiles in flight, the throughput variant
// for trees with thousands of tiles) or 2 (252 registers: the solve and the Cholesky schedule better — 3000 vs 5300 and
// 7700 vs 8100 cycles measured in the kernel — the latency variant for small trees).
// W = R^-1 of the factored diagonal block (D holds R, upper), written to global memory column-major (wout[j * 32 + i] =
// W[i][j]): lane = column j of W, right-looking back substitution in registers (2700 cycles measured alone).  It runs AFTER
// the block has been published — off the pivot chain — and turns the 32 x 32 triangular solves of back-substitution
// (32 dependent shuffle steps each) into matrix-vector products.
template <int MINB>
__device__ B200_NOINLINE void df_inv32(int d_off, int lane, double* wout) {
  B200_DYN_SMEM(double, df_smem);
  double (*D)[kDfB + 1] = (double (*)[kDfB + 1])(df_smem + d_off);
  double s[kDfB];   // s[i] = e_j[i] - sum_{m > i} R[i][m] w[m]
#pragma unroll
  for (int i = 0; i < kDfB; i++) s[i] = (i == lane) ? 1.0 : 0.0;
  const double dinv = 1.0 / D[lane][lane];
#pragma unroll
  for (int m = kDfB - 1; m >= 0; m--) {
    s[m] *= __shfl_sync(0xffffffffu, dinv, m);      // w[m] (zero for m > lane)
#pragma unroll
    for (int i = 0; i < m; i++) s[i] -= D[i][m] * s[m];
  }
#pragma unroll
  for (int i = 0; i < kDfB; i++) wout[lane * kDfB + i] = s[i];
}

template <int MINB>
__global__ void __launch_bounds__(kDfThreads, MINB)
front_df_kernel(TreeView t, DfView v, Scalars* sc) {
  B200_DYN_SMEM(double, df_smem);                                  // kDfSmemBytes, carved up:
  double* Pc = df_smem;                                            // column piece (k, j): the B operand, [column][pivot row]
  double (*Pr)[kDfB * kDfLd] = (double (*)[kDfB * kDfLd])(df_smem + kDfOffPr);   // row pieces (k, 4r + w): the A operands
  double (*Dg)[kDfB + 1] = (double (*)[kDfB + 1])(df_smem + kDfOffDg);           // pivot rows out of the accumulators
  double* Rk = df_smem + kDfOffRk;                                 // R_kk: Rk[row * kDfLdR + col]
  double* invd = df_smem + kDfOffInvd;
  __shared__ int s_task, s_ok, s_ready;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, q = lane & 3;
#ifndef B200_EMULATE
  if (blockIdx.x < (unsigned)v.warm_ctas) {
    // instruction-cache warm-up (see df_chol32): both routines once on an identity block, shared memory only, before
    // the grid dependency is resolved; every warp, so that each SM sub-partition has fetched the code
    double (*Dw)[kDfB + 1] = (double (*)[kDfB + 1])Pr[w];
    for (int ii = 0; ii < kDfB; ii++) { Dw[ii][lane] = (ii == lane) ? 1.0 : 0.0; if (w == 0) Rk[ii * kDfLdR + lane] = (ii == lane) ? 1.0 : 0.0; }
    if (tid < kDfB) invd[tid] = 1.0;
    __syncthreads();
    df_trsm32<MINB>(kDfOffPr + w * kDfB * kDfLd, lane);
    if (w == 0) (void)df_chol32<MINB>(kDfOffPr, kDfB, lane);
    __syncthreads();
  }
#endif
  pdl_sync();
  if (tid == 0) { s_task = atomicAdd(v.ctrl, 1); s_ok = 1; s_ready = 1 << 30; }
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
    if (!s_ok) { if (tid == 0) atomicExch(&sc->df_abort, 1); return; }
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
        // branch-free: 32 independent loads in flight (an entry outside the block reads M[0] and is dropped).  With the load
        // under its predicate the compiler kept each negation next to its load inside a branch region: 32 L2 round trips
        // in sequence, ~7 us per tile in the trace of the 10M-factor graph (round 2).
        const int lr = 8 * a + g, lc = 8 * b + 2 * q + h;
        const bool in = lr < rs && lc < cs && rb + lr <= cb + lc;
        const double val = __ldcg(M + (in ? (size_t)(rb + lr) + (size_t)(cb + lc) * n : (size_t)0));
        acc[a][b][h] = in ? -val : 0.0;
      }
  const int last_rb = min(min(kDfTR * r + kDfTR - 1, j), NB - 1);   // last row block of the tile
  const bool diag_tile = j < K && last_rb == j;           // holds the diagonal block of pivot column j: factored after the loop
  const int kend = min(K, last_rb + 1) - (diag_tile ? 1 : 0);
  // One look at the flags of EVERY step (5 per step: the column piece or R_kk, one row piece per warp), all in flight at
  // once while the loads of C above are: the steps before the first unpublished piece need no polling.  (Trace of the
  // 10M-factor graph, round 2: half of the steps found their pieces published and still paid ~1 us of poll + fence +
  // barrier each; a tile of a front whose chain is ahead runs straight through now.)
  {
    int first_missing = 1 << 30;
    for (int e = tid; e < 5 * kend; e += kDfThreads) {
      const int k = e / 5, which = e - 5 * k;
      int fl = -1;
      if (which == 0) fl = k * NB + (k - kDfTR * r >= 0 ? k : j);
      else { const int iw = kDfTR * r + which - 1; if (iw <= j && iw < NB && iw > k && iw != j) fl = k * NB + iw; }
      if (fl >= 0 && df_ld_relaxed(flags + fl) < 1) first_missing = min(first_missing, k);
    }
    if (first_missing < (1 << 30)) atomicMin(&s_ready, first_missing);
    df_fence_acquire();
    __syncthreads();
  }
  const int ready = s_ready;                              // steps k < ready: every piece this tile reads is published and visible
  for (int k = 0; k < kend; k++) {
    const int wb = k - kDfTR * r;                         // the warp that holds pivot block k (< 0: above the tile)
    const int kb = kDfB * k, ks = min(kDfB, f - kb);      // pivot rows
    const bool need_row = wvalid && i > k && i != j;      // this warp updates rows below the pivot block with a piece of another tile
    if (wb >= 0) {
      // ================= pivot rows live in this tile: X = R_kk^-T C, publish piece (k, j) =================
      if (k >= ready && tid == 0 && !df_wait(flags + k * NB + k, 1, v.ctrl)) s_ok = 0;
      if (w == wb) {
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int b = 0; b < 4; b++)
#pragma unroll
            for (int h = 0; h < 2; h++) Dg[8 * a + g][8 * b + 2 * q + h] = -acc[a][b][h];
      }
      __syncthreads();
      if (!s_ok) { if (tid == 0) atomicExch(&sc->df_abort, 1); return; }
      DF_STAMP(3);
      for (int e = tid; e < kDfB * kDfB; e += kDfThreads) {       // R_kk (upper): one batch of cp.async; zero beyond the pivots
        const int p = e & 31, x = e >> 5;
        double* d = Rk + p * kDfLdR + x;
        if (p <= x && x < ks) cp_async8(d, M + (kb + p) + (size_t)(kb + x) * n);
        else *d = 0.0;
      }
      cp_async_commit();
      cp_async_wait<0>();
      __syncthreads();
      if (tid < kDfB) invd[tid] = tid < ks ? 1.0 / Rk[tid * kDfLdR + tid] : 1.0;
      __syncthreads();
      DF_STAMP(4);
      if (w == wb) {
        const long long c0_ = clock64();
        df_trsm32<MINB>(kDfOffDg, lane);
        if (wb == 0) DF_CYC(21, clock64() - c0_);   // (only warp 0 shares the trace cursor with thread 0)
#pragma unroll 8
        for (int p = 0; p < kDfB; p++) {
          const double xp = (p < ks && lane < cs) ? Dg[p][lane] : 0.0;
          Pc[lane * kDfLd + p] = xp;
          if (p < ks && lane < cs) M[(kb + p) + (size_t)(cb + lane) * n] = xp;
        }
      }
      __syncthreads();
      DF_STAMP(5);
      if (tid == 0) df_st_release(flags + k * NB + j, 1);   // release = fence + store
      DF_STAMP(6);
      // only now the pieces of the other tiles (the rows below the pivot block): the TRSM above never waits for them
      if (need_row) {
        if (k >= ready && lane == 0 && !df_wait(flags + k * NB + i, 1, v.ctrl)) s_ok = 0;
        __syncwarp();
        df_stage(Pr[w], M, n, kb, ks, rb, rs, lane, 32);
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
      }
    } else {
      // ================= pivot block above the tile: fetch the two pieces =================
      if (k >= ready) {
        if (tid == 0 && !df_wait(flags + k * NB + j, 1, v.ctrl)) s_ok = 0;
        if (need_row && lane == 0 && !df_wait(flags + k * NB + i, 1, v.ctrl)) s_ok = 0;
        __syncthreads();
        if (!s_ok) { if (tid == 0) atomicExch(&sc->df_abort, 1); return; }
      }
      DF_STAMP(7);
      df_stage(Pc, M, n, kb, ks, cb, cs, tid, kDfThreads);
      if (need_row) df_stage(Pr[w], M, n, kb, ks, rb, rs, lane, 32);
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
    if (!s_ok) { if (tid == 0) atomicExch(&sc->df_abort, 1); return; }
    DF_STAMP(9);
  }
  if (diag_tile) {
    // ================= diagonal block: Cholesky by one warp, publish R_kk =================
    const int k = j, wb = k - kDfTR * r, kb = kDfB * k, ks = min(kDfB, f - kb);
    DF_STAMP(10);
    if (w == wb) {
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
          for (int h = 0; h < 2; h++) Dg[8 * a + g][8 * b + 2 * q + h] = -acc[a][b][h];
      __syncwarp();
      const long long c0_ = clock64();
      const bool notpd = df_chol32<MINB>(kDfOffDg, ks, lane);
      if (wb == 0) DF_CYC(20, clock64() - c0_);
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
#pragma unroll 8
        for (int ii = 0; ii < kDfB; ii++)
          if (ii <= lane) M[(kb + ii) + (size_t)(kb + lane) * n] = Dg[ii][lane];
      }
    }
    __syncthreads();
    DF_STAMP(11);
    if (tid == 0) df_st_release(flags + k * NB + k, 1);
    DF_STAMP(12);
    if (w == wb && v.winv) df_inv32<MINB>(kDfOffDg, lane, v.winv + v.winv_off[c] + (size_t)k * kDfB * kDfB);
    // (a pivot column has no trailing rows: nothing to extend-add)
  } else if (wvalid && i >= K) {
    // ---- trailing rows: the Schur complement goes straight into the parent (or stays, for a root) ----
    if (par >= 0) {
      double* P = t.arena + t.off[par];
      const int pn = t.nf[par] + t.ns[par] + 1;
      const int* map = t.ea_map + t.ea_ptr[c];
      // the 4 row and 8 column slots of this thread's 32 entries first (12 independent loads, one latency), then the adds
      // back to back.  (Trace of the 10M-factor graph, round 2: with the map looked up per entry the extend-add held the
      // slot 11 us per tile, as long as four rank-32 updates.)
      int mr[4], mc[4][2];
#pragma unroll
      for (int a = 0; a < 4; a++) mr[a] = (8 * a + g < rs) ? map[rb + 8 * a + g - f] : -1;
#pragma unroll
      for (int b = 0; b < 4; b++)
#pragma unroll
        for (int h = 0; h < 2; h++) mc[b][h] = (8 * b + 2 * q + h < cs) ? map[cb + 8 * b + 2 * q + h - f] : -1;
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int lr = 8 * a + g, lc = 8 * b + 2 * q + h;
            if (lr < rs && lc < cs && rb + lr <= cb + lc) {
              const int pi = mr[a], pj = mc[b][h];
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

// packed view <-> full packed Values (b200_set_values_view / b200_get_values_view)
__global__ void __launch_bounds__(256) values_view_kernel(double* values, double* packed, const int* __restrict__ idx, int64_t n, int to_values) {
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= n) return;
  if (to_values == 2) packed[idx[e]] = values[idx[e]];       // same position in a full-size buffer (b200_get_values_all)
  else if (to_values) values[idx[e]] = packed[e];
  else packed[e] = values[idx[e]];
}

// Sharded LM try: [lin_err0, lin_err_delta, new_error | fail code of rank r | nan code of rank r] <-> Scalars (one SUM all-reduce).
constexpr int kMaxRanksMerged = 16;
__global__ void scalars_pack_kernel(Scalars* sc, double* red, int rank, int world, int unpack) {
  pdl_sync();
  if (threadIdx.x != 0) return;
  if (!unpack) {
    red[0] = sc->lin_err0; red[1] = sc->lin_err_delta; red[2] = sc->new_error;
    for (int r = 0; r < world; r++) { red[3 + r] = r == rank ? (double)sc->fail_code : 0.0; red[3 + world + r] = r == rank ? (double)sc->nan_code : 0.0; }
  } else {
    sc->lin_err0 = red[0]; sc->lin_err_delta = red[1]; sc->new_error = red[2];
    int fc = 0, nc = 0;
    for (int r = 0; r < world; r++) { fc = max(fc, (int)red[3 + r]); nc = max(nc, (int)red[3 + world + r]); }
    sc->fail_code = fc; sc->nan_code = nc;
  }
}

// Sharded solve, distributed top: the solutions of the top fronts a rank owns travel to the other ranks in a packed vector.
// scatter = 0: topx <- delta for the fronts this rank owns (the others leave zeros); scatter = 1: delta <- topx (all fronts).
__global__ void __launch_bounds__(128) top_x_kernel(TreeView t, const int* __restrict__ cliques, const int* __restrict__ xoff,
                                                    const int* __restrict__ owned, double* delta, double* topx, int scatter) {
  pdl_sync();
  const int c = cliques[blockIdx.x];
  if (!scatter && !owned[blockIdx.x]) return;
  const int f = t.nf[c];
  const int* di = t.didx + t.didx_ptr[c];
  double* x = topx + xoff[blockIdx.x];
  for (int i = threadIdx.x; i < f; i += 128) {
    if (scatter) delta[di[i]] = x[i];
    else x[i] = delta[di[i]];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Roofline denominators measured on the device the bench runs on (MEASURED_PEAKS.json holds HBM and bf16 only):
// register-resident FP64 throughput of the tensor path (mma.sync.m8n8k4.f64, 512 flop per warp instruction) and of
// the FMA pipe (64 flop per warp instruction), 16 independent accumulator chains per warp, every SM full.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(int iters, int use_dmma, double* sink) {
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
  double af[4], bf[4];
#pragma unroll
  for (int a = 0; a < 4; a++) { af[a] = 1e-3 * (threadIdx.x + a); bf[a] = 1e-3 * (blockIdx.x + a); }
  if (use_dmma) {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
  } else {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) { acc[a][b][0] = fma(af[a], bf[b], acc[a][b][0]); acc[a][b][1] = fma(bf[a], af[b], acc[a][b][1]); }
    }
  }
  double s = 0;
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) s += acc[a][b][0] + acc[a][b][1];
  if (s == 12345.678) *sink = s;   // keeps the loop alive
}

}  // namespace b200
