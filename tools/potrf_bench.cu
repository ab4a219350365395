// Micro-benchmark of 32x32 fp64 Cholesky variants for the dense-solver critical path (one CTA, clock64 timing),
// plus raw dependent-chain latencies (DFMA, rsqrt, shuffle, barrier) on the target GPU.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o tools/_build/potrf_bench tools/potrf_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../mcptam_b200/csrc/ba_solve.cu"

using namespace mcp;

#define NREP 64

template <int V>
__global__ void __launch_bounds__(256) k_bench(const double* A, double* Lout, long long* cycles)
{
  __shared__ double S[TB * TLD];
  __shared__ double X[TB * TLD];
  __shared__ double xs[64 + 256 + 64];
  __shared__ int s_bad;
  const int tid = threadIdx.x;
  long long tot = 0;
  for (int rep = 0; rep < NREP; rep++) {
    for (int e = tid; e < TB * TB; e += 256) S[(e >> 5) * TLD + (e & 31)] = A[e];
    if (tid == 0) s_bad = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (V == 0) potrf32_blocked(S, xs, xs + 64, xs + 128, tid, &s_bad);
    else if (V == 1) { if (tid < 32) potrf32_warp(S, xs, tid); __syncthreads(); }
    else if (V == 2) potrf32_panel<false>(S, xs, tid, &s_bad);
    else if (V == 3) potrf32_panel<true>(S, xs, tid, &s_bad);
    else if (V == 4) potrf32_cols<true>(S, xs, xs + 64, tid, &s_bad);
    else if (V == 5) potrf32_cols<false>(S, xs, xs + 64, tid, &s_bad);
    else if (V == 6) { potrf32_panel<true>(S, xs, tid, &s_bad); inverse32_block(S, X, xs, xs + 64, tid); }
    else if (V == 7) potrf32_panel<true, 4>(S, xs, tid, &s_bad);
    else if (V == 8) potrf32_panel<true, 16>(S, xs, tid, &s_bad);
    __syncthreads();
    const long long t1 = clock64();
    if (rep >= 4) tot += t1 - t0;
    __syncthreads();
  }
  for (int e = tid; e < TB * TB; e += 256) { const int r = e >> 5, c = e & 31; Lout[e] = (c <= r) ? S[r * TLD + c] : 0.0; }
  if (tid == 0) cycles[V] = tot / (NREP - 4);
}

__global__ void k_latency(double seed, long long* out, double* sink)
{
  __shared__ double sm[64];
  const int tid = threadIdx.x;
  double x = seed + 1e-9 * tid, y = 1.0000001;
  long long t0, t1;
  // DFMA chain
  __syncthreads(); t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) { x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); }
  t1 = clock64(); if (tid == 0) out[0] = (t1 - t0) / 256;
  // rsqrt(double) chain
  x = fabs(x) + 1.5; __syncthreads(); t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) { x = rsqrt(x) + 1.5; x = rsqrt(x) + 1.5; x = rsqrt(x) + 1.5; x = rsqrt(x) + 1.5; }
  t1 = clock64(); if (tid == 0) out[1] = (t1 - t0) / 256;
  // fast rsqrt chain
  __syncthreads(); t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) { x = fast_rsqrt(x) + 1.5; x = fast_rsqrt(x) + 1.5; x = fast_rsqrt(x) + 1.5; x = fast_rsqrt(x) + 1.5; }
  t1 = clock64(); if (tid == 0) out[2] = (t1 - t0) / 256;
  // double shuffle chain
  __syncthreads(); t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) { x = __shfl_sync(0xffffffffu, x, (i + 1) & 31); x = __shfl_sync(0xffffffffu, x, (i + 5) & 31); x = __shfl_sync(0xffffffffu, x, (i + 9) & 31); x = __shfl_sync(0xffffffffu, x, (i + 3) & 31); }
  t1 = clock64(); if (tid == 0) out[3] = (t1 - t0) / 256;
  // __syncthreads chain
  __syncthreads(); t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) { __syncthreads(); __syncthreads(); __syncthreads(); __syncthreads(); }
  t1 = clock64(); if (tid == 0) out[4] = (t1 - t0) / 256;
  // smem store -> barrier -> load round trip
  __syncthreads(); t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; i++) { if (tid == (i & 31)) sm[i & 63] = x; __syncthreads(); x += sm[i & 63]; }
  t1 = clock64(); if (tid == 0) out[5] = (t1 - t0) / 256;
  // sqrt + division chain
  x = fabs(x) + 2.0; __syncthreads(); t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) { x = 1.0 / x + 1.5; x = 1.0 / x + 1.5; x = 1.0 / x + 1.5; x = 1.0 / x + 1.5; }
  t1 = clock64(); if (tid == 0) out[6] = (t1 - t0) / 256;
  // smem store -> syncwarp -> broadcast load (warp-local)
  __syncthreads(); t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; i++) { if ((tid & 31) == (i & 31)) sm[(tid >> 5) * 8 + (i & 7)] = x; __syncwarp(); x += sm[(tid >> 5) * 8 + (i & 7)]; __syncwarp(); }
  t1 = clock64(); if (tid == 0) out[7] = (t1 - t0) / 256;
  sink[tid] = x;
}

int main()
{
  const int n = TB;
  std::vector<double> A(n * n), M(n * n);
  srand(1);
  for (auto& v : M) v = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      double s = 0;
      for (int k = 0; k < n; k++) s += M[i * n + k] * M[j * n + k];
      A[i * n + j] = 1e4 * s + (i == j ? 1e3 : 0.0);
    }
  double *dA, *dL, *dsink; long long *dc, *dlat;
  cudaMalloc(&dA, sizeof(double) * n * n); cudaMalloc(&dL, sizeof(double) * n * n); cudaMalloc(&dc, sizeof(long long) * 16);
  cudaMalloc(&dlat, sizeof(long long) * 16); cudaMalloc(&dsink, sizeof(double) * 256);
  cudaMemcpy(dA, A.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice);
  cudaMemset(dc, 0, sizeof(long long) * 16);
  const char* names[] = { "V0 blocked (current: 8-lane panel chain + inverse8 + X + trailing)", "V1 one warp, register rows", "V2 panel chain over all rows, rsqrt()",
                          "V3 panel chain over all rows, fast rsqrt", "V4 8 warps x 4 columns, 1 barrier / pivot, fast rsqrt", "V5 same, rsqrt()", "V6 = V3 + inverse32_block", "V7 panel width 4, fast rsqrt", "V8 panel width 16, fast rsqrt" };
  std::vector<double> L(n * n);
  for (int v = 0; v < 9; v++) {
    switch (v) {
      case 0: k_bench<0><<<1, 256>>>(dA, dL, dc); break;
      case 1: k_bench<1><<<1, 256>>>(dA, dL, dc); break;
      case 2: k_bench<2><<<1, 256>>>(dA, dL, dc); break;
      case 3: k_bench<3><<<1, 256>>>(dA, dL, dc); break;
      case 4: k_bench<4><<<1, 256>>>(dA, dL, dc); break;
      case 5: k_bench<5><<<1, 256>>>(dA, dL, dc); break;
      case 6: k_bench<6><<<1, 256>>>(dA, dL, dc); break;
      case 7: k_bench<7><<<1, 256>>>(dA, dL, dc); break;
      case 8: k_bench<8><<<1, 256>>>(dA, dL, dc); break;
    }
    cudaError_t e = cudaDeviceSynchronize();
    long long c[16];
    cudaMemcpy(c, dc, sizeof(c), cudaMemcpyDeviceToHost);
    cudaMemcpy(L.data(), dL, sizeof(double) * n * n, cudaMemcpyDeviceToHost);
    double err = 0, nrm = 0;
    for (int i = 0; i < n; i++)
      for (int j = 0; j <= i; j++) {
        double s = 0;
        for (int k = 0; k <= j; k++) s += L[i * n + k] * L[j * n + k];
        err = fmax(err, fabs(s - A[i * n + j])); nrm = fmax(nrm, fabs(A[i * n + j]));
      }
    printf("%-75s %7lld cycles  (%.2f us @1.965GHz)  max|LL^T-A|/max|A| = %.2e  %s\n", names[v], c[v], c[v] / 1965.0, err / nrm, cudaGetErrorString(e));
  }
  k_latency<<<1, 256>>>(1.0, dlat, dsink);
  cudaDeviceSynchronize();
  long long lat[16];
  cudaMemcpy(lat, dlat, sizeof(lat), cudaMemcpyDeviceToHost);
  printf("latency (cycles, dependent chain): DFMA %lld  rsqrt(double) %lld  fast_rsqrt %lld  shfl(double) %lld  __syncthreads(256) %lld  STS+bar+LDS %lld  1/x %lld  STS+syncwarp+LDS %lld\n",
         lat[0], lat[1] , lat[2], lat[3], lat[4], lat[5], lat[6], lat[7]);
  return 0;
}
