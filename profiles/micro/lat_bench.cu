// Micro-benchmarks behind the design of front_df_kernel's critical path: one warp, clock64 around
//  (a) a chain of dependent DFMA / DMUL / DMMA / rsqrt / shuffles (latency per dependent op),
//  (b) the 32 x 32 Cholesky variants, (c) the 32 x 32 triangular solves (register right-looking, DMMA with inverse).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_bench lat_bench.cu ; ./lat_bench
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <cmath>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void lat_kernel(double* out, long long* cyc, double seed) {
  const int lane = threadIdx.x;
  double x = seed + lane * 1e-3, y = 1.0000001, z = 0.5;
  long long t0, t1;
  // dependent DFMA chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) x = fma(x, y, z);
  t1 = clock64();
  if (lane == 0) cyc[0] = (t1 - t0);
  // dependent DMUL chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) x = x * y;
  t1 = clock64();
  if (lane == 0) cyc[1] = (t1 - t0);
  // dependent DMMA chain
  double d0 = x, d1 = y;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; i++) dmma(d0, d1, 1e-3, 1e-3);
  t1 = clock64();
  if (lane == 0) cyc[2] = (t1 - t0);
  // dependent rsqrt chain
  double r = fabs(x) + 2.0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 32; i++) r = rsqrt(r) + 1.5;
  t1 = clock64();
  if (lane == 0) cyc[3] = (t1 - t0);
  // dependent shuffle (64-bit) chain
  double s = r;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; i++) s = __shfl_sync(0xffffffffu, s, (lane + 1) & 31);
  t1 = clock64();
  if (lane == 0) cyc[4] = (t1 - t0);
  // 16 independent DFMA chains (throughput of one warp)
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = s + i;
  t0 = clock64();
#pragma unroll
  for (int it = 0; it < 64; it++)
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = fma(a[i], y, z);
  t1 = clock64();
  if (lane == 0) cyc[5] = (t1 - t0);
  double acc = d0 + d1 + s + x;
#pragma unroll
  for (int i = 0; i < 16; i++) acc += a[i];
  // 1/x chain
  double q = acc * 1e-300 + 3.0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 32; i++) q = 1.0 / q + 2.0;
  t1 = clock64();
  if (lane == 0) cyc[6] = (t1 - t0);
  out[lane] = acc + q;
}

// ---- 32 x 32 Cholesky, one warp, lane = column in registers, pivot row through shared memory (front_df v2) ----
__global__ void chol_reg_kernel(const double* A, double* R, long long* cyc) {
  __shared__ double Dg[32][33];
  __shared__ __align__(16) double rowbuf[2][32];
  const int lane = threadIdx.x;
  for (int i = 0; i < 32; i++) Dg[i][lane] = A[i * 32 + lane];
  __syncwarp();
  for (int rep = 0; rep < 3; rep++) {
    long long t0 = clock64();
    double col[32];
#pragma unroll
    for (int ii = 0; ii < 32; ii++) col[ii] = (ii <= lane) ? Dg[ii][lane] : 0.0;
    double dg = Dg[lane][lane];
#pragma unroll
    for (int kk = 0; kk < 32; kk++) {
      const double akk = __shfl_sync(0xffffffffu, dg, kk);
      const double rinv = rsqrt(akk);
      const double rr = (lane == kk) ? akk * rinv : col[kk] * rinv;
      col[kk] = rr;
      dg -= rr * rr;
      double* rb_ = rowbuf[kk & 1];
      rb_[lane] = rr;
      __syncwarp();
#pragma unroll
      for (int ii = kk + 1; ii < 32; ii++) col[ii] -= rb_[ii] * rr;
    }
    long long t1 = clock64();
    if (lane == 0) cyc[rep] = t1 - t0;
#pragma unroll
    for (int ii = 0; ii < 32; ii++) R[ii * 32 + lane] = (ii <= lane) ? col[ii] : 0.0;
    __syncwarp();
  }
}

// ---- TRSM X = R^-T C, lane = column, registers, right-looking (front_df v2) ----
__global__ void trsm_reg_kernel(const double* Rm, const double* C, double* X, long long* cyc) {
  __shared__ __align__(16) double Rk[32][34];
  __shared__ double Dg[32][33];
  __shared__ double invd[32];
  const int lane = threadIdx.x;
  for (int i = 0; i < 32; i++) { Rk[i][lane] = Rm[i * 32 + lane]; Dg[i][lane] = C[i * 32 + lane]; }
  __syncwarp();
  invd[lane] = 1.0 / Rk[lane][lane];
  __syncwarp();
  for (int rep = 0; rep < 3; rep++) {
    long long t0 = clock64();
    double x[32];
#pragma unroll
    for (int p = 0; p < 32; p++) x[p] = Dg[p][lane];
#pragma unroll
    for (int qq = 0; qq < 32; qq++) {
      x[qq] *= invd[qq];
#pragma unroll
      for (int p = qq + 1; p < 32; p++) x[p] -= Rk[qq][p] * x[qq];
    }
    long long t1 = clock64();
    if (lane == 0) cyc[rep] = t1 - t0;
#pragma unroll
    for (int p = 0; p < 32; p++) X[p * 32 + lane] = x[p];
    __syncwarp();
  }
}

// ---- TRSM as a GEMM with the explicit inverse W = R^-1: X = W^T C on the tensor path, one warp, 4 x 4 blocks of 8 x 8 ----
__global__ void trsm_dmma_kernel(const double* Wm, const double* C, double* X, long long* cyc) {
  __shared__ double Ws[32 * 36];   // Ws[col * 36 + row] = W[row][col]
  __shared__ double Cs[32 * 36];   // Cs[col * 36 + row] = C[row][col]
  const int lane = threadIdx.x, g = lane >> 2, q = lane & 3;
  for (int i = 0; i < 32; i++) { Ws[lane * 36 + i] = Wm[i * 32 + lane]; Cs[lane * 36 + i] = C[i * 32 + lane]; }
  __syncwarp();
  for (int rep = 0; rep < 3; rep++) {
    long long t0 = clock64();
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
    // X[i][x] = sum_k W[k][i] C[k][x]; W upper: k <= i.  A[i][k] = W[k][i] = Ws[i * 36 + k]; B[k][x] = Cs[x * 36 + k]
#pragma unroll
    for (int k4 = 0; k4 < 32; k4 += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int a = 0; a < 4; a++) af[a] = Ws[(8 * a + g) * 36 + k4 + q];
#pragma unroll
      for (int b = 0; b < 4; b++) bf[b] = Cs[(8 * b + g) * 36 + k4 + q];
#pragma unroll
      for (int a = 0; a < 4; a++)
        if (8 * a + 7 >= k4) {
#pragma unroll
          for (int b = 0; b < 4; b++) dmma(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
    }
    long long t1 = clock64();
    if (lane == 0) cyc[rep] = t1 - t0;
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++)
#pragma unroll
        for (int h = 0; h < 2; h++) X[(8 * a + g) * 32 + 8 * b + 2 * q + h] = acc[a][b][h];
    __syncwarp();
  }
}

// ---- explicit inverse of the upper-triangular factor, W = R^-1: lane = column j of W, right-looking back substitution ----
__global__ void inv_reg_kernel(const double* Rm, double* Wm, long long* cyc) {
  __shared__ __align__(16) double Rt[32][34];   // Rt[m][i] = R[i][m]: column m of R contiguous over the rows i
  __shared__ double invd[32];
  const int lane = threadIdx.x;
  for (int i = 0; i < 32; i++) Rt[lane][i] = Rm[i * 32 + lane];
  __syncwarp();
  invd[lane] = 1.0 / Rt[lane][lane];
  __syncwarp();
  for (int rep = 0; rep < 3; rep++) {
    long long t0 = clock64();
    double s[32];   // s[i] = e_j[i] - sum_{m > i} R[i][m] w[m]
#pragma unroll
    for (int i = 0; i < 32; i++) s[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int m = 31; m >= 0; m--) {
      s[m] *= invd[m];            // w[m] (zero for m > lane)
#pragma unroll
      for (int i = 0; i < m; i++) s[i] -= Rt[m][i] * s[m];
    }
    long long t1 = clock64();
    if (lane == 0) cyc[rep] = t1 - t0;
#pragma unroll
    for (int i = 0; i < 32; i++) Wm[i * 32 + lane] = s[i];
    __syncwarp();
  }
}

int main() {
  const int n = 32;
  std::vector<double> A(n * n), C(n * n);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) { A[i * n + j] = (i == j ? 40.0 : 0.0) + 1.0 / (1 + abs(i - j)); C[i * n + j] = sin(i * 0.3 + j * 0.7); }
  double *dA, *dR, *dC, *dX, *dW, *dX2, *dout;
  long long* dc;
  cudaMalloc(&dA, n * n * 8); cudaMalloc(&dR, n * n * 8); cudaMalloc(&dC, n * n * 8); cudaMalloc(&dX, n * n * 8);
  cudaMalloc(&dW, n * n * 8); cudaMalloc(&dX2, n * n * 8); cudaMalloc(&dout, 32 * 8); cudaMalloc(&dc, 64 * 8);
  cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dC, C.data(), n * n * 8, cudaMemcpyHostToDevice);
  long long h[16];
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("SM clock (attr) %d kHz\n", clk);
  for (int rep = 0; rep < 2; rep++) {
    lat_kernel<<<1, 32>>>(dout, dc, 1.0);
    cudaMemcpy(h, dc, 8 * 8, cudaMemcpyDeviceToHost);
    printf("run %d: DFMA dep %.1f cyc | DMUL dep %.1f | DMMA dep %.1f | rsqrt+add dep %.1f | shfl64 dep %.1f | DFMA x16 indep %.2f cyc/instr | 1/x+add dep %.1f\n", rep,
           h[0] / 256.0, h[1] / 256.0, h[2] / 64.0, h[3] / 32.0, h[4] / 64.0, h[5] / (64.0 * 16), h[6] / 32.0);
  }
  chol_reg_kernel<<<1, 32>>>(dA, dR, dc);
  cudaMemcpy(h, dc, 3 * 8, cudaMemcpyDeviceToHost);
  printf("chol32 registers + smem row: %lld %lld %lld cycles\n", h[0], h[1], h[2]);
  trsm_reg_kernel<<<1, 32>>>(dR, dC, dX, dc);
  cudaMemcpy(h, dc, 3 * 8, cudaMemcpyDeviceToHost);
  printf("trsm32 registers right-looking: %lld %lld %lld cycles\n", h[0], h[1], h[2]);
  inv_reg_kernel<<<1, 32>>>(dR, dW, dc);
  cudaMemcpy(h, dc, 3 * 8, cudaMemcpyDeviceToHost);
  printf("inverse32 registers: %lld %lld %lld cycles\n", h[0], h[1], h[2]);
  trsm_dmma_kernel<<<1, 32>>>(dW, dC, dX2, dc);
  cudaMemcpy(h, dc, 3 * 8, cudaMemcpyDeviceToHost);
  printf("trsm32 as DMMA GEMM with W: %lld %lld %lld cycles\n", h[0], h[1], h[2]);
  std::vector<double> X(n * n), X2(n * n), R(n * n);
  cudaMemcpy(X.data(), dX, n * n * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(X2.data(), dX2, n * n * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(R.data(), dR, n * n * 8, cudaMemcpyDeviceToHost);
  double e = 0, e2 = 0, nrm = 0;
  for (int i = 0; i < n * n; i++) { e = fmax(e, fabs(X[i] - X2[i])); nrm = fmax(nrm, fabs(X[i])); }
  // check R^T R = A
  for (int i = 0; i < n; i++)
    for (int j = i; j < n; j++) { double s = 0; for (int k = 0; k <= i; k++) s += R[k * n + i] * R[k * n + j]; e2 = fmax(e2, fabs(s - A[i * n + j])); }
  printf("check: |X_reg - X_dmma| = %.3e (|X| %.3e), |R^T R - A| = %.3e, err %s\n", e, nrm, e2, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
