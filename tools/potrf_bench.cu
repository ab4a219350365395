// Micro-benchmark of 32x32 fp64 Cholesky variants for the dense-solver critical path (one CTA, clock64 timing),
// plus raw dependent-chain latencies (DFMA, rsqrt, shuffle, barrier) on the target GPU.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o tools/_build/potrf_bench tools/potrf_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../mcptam_b200/csrc/ba_solve.cu"
namespace mcp { bool pdl_enabled() { return false; } }

using namespace mcp;

#define NREP 64

template <int PW, int MODE>
__device__ __forceinline__ void potrf32_pairs_x(double* __restrict__ S, double* __restrict__ rinv, double* __restrict__ wbuf, int tid, int* bad)
{
  const int lane = tid & 31, wid = tid >> 5;
#pragma unroll 1
  for (int o = 0; o < TB; o += PW) {
    if (wid == 0) {
      double a[PW];
#pragma unroll
      for (int c = 0; c < PW; c++) a[c] = S[lane * TLD + o + c];
      bool isbad = false;
      double myrinv = 0.0;
      double App = __shfl_sync(0xffffffffu, a[0], o), Aqp = __shfl_sync(0xffffffffu, a[0], o + 1), Aqq = __shfl_sync(0xffffffffu, a[1], o + 1);
      double R0 = 0, R1 = 0, R2 = 0, R3 = 0, R4 = 0, R5 = 0, R6 = 0;
      if (PW > 2) {
        R0 = __shfl_sync(0xffffffffu, a[0], o + 2); R1 = __shfl_sync(0xffffffffu, a[1], o + 2);
        R2 = __shfl_sync(0xffffffffu, a[0], o + 3); R3 = __shfl_sync(0xffffffffu, a[1], o + 3);
        R4 = __shfl_sync(0xffffffffu, a[2], o + 2); R5 = __shfl_sync(0xffffffffu, a[2], o + 3); R6 = __shfl_sync(0xffffffffu, a[3], o + 3);
      }
#pragma unroll
      for (int t = 0; t < PW / 2; t++) {
        const int p = o + 2 * t, q = p + 1;
        double det = fma(App, Aqq, -Aqp * Aqp);
        {
          const bool ok1 = (App > 1.0e-290) && (App < 1.0e290);
          isbad |= !ok1;
          App = ok1 ? App : 1.0;
          const bool ok2 = ok1 && (det > 1.0e-290 * App) && (det < 1.0e290) ;
          isbad |= !ok2;
          det = ok2 ? det : App;
        }
        const double r1 = fast_rsqrt(App), r2 = fast_rsqrt(det);
        const double l11 = App * r1, l21 = Aqp * r1, i22 = r2 * l11;
        double x0 = a[2 * t] * r1;
        double x1 = fma(-x0, l21, a[2 * t + 1]) * i22;
        x0 = (lane >= p) ? x0 : 0.0;
        x1 = (lane >= q) ? x1 : 0.0;
        a[2 * t] = x0; a[2 * t + 1] = x1;
        myrinv = (lane == p) ? r1 : ((lane == q) ? i22 : myrinv);
        if (t + 1 < PW / 2) {
          const double Lp0 = R0 * r1, Lq0 = R2 * r1;
          const double Lp1 = fma(-Lp0, l21, R1) * i22, Lq1 = fma(-Lq0, l21, R3) * i22;
          App = fma(-Lp1, Lp1, fma(-Lp0, Lp0, R4));
          Aqp = fma(-Lq1, Lp1, fma(-Lq0, Lp0, R5));
          Aqq = fma(-Lq1, Lq1, fma(-Lq0, Lq0, R6));
          if (MODE < 2) {
          double* xb = wbuf + 64 * (t & 1);
          *reinterpret_cast<double2*>(xb + 2 * lane) = make_double2(x0, x1);
          __syncwarp();
#pragma unroll
          for (int c = 2 * t + 2; c < PW; c++) {
            const double2 y = *reinterpret_cast<const double2*>(xb + 2 * (o + c));
            a[c] = fma(-x1, y.y, fma(-x0, y.x, a[c]));
          }
          }
          if (MODE < 1 && t + 2 < PW / 2) {
            R0 = __shfl_sync(0xffffffffu, a[2 * t + 2], p + 4); R1 = __shfl_sync(0xffffffffu, a[2 * t + 3], p + 4);
            R2 = __shfl_sync(0xffffffffu, a[2 * t + 2], q + 4); R3 = __shfl_sync(0xffffffffu, a[2 * t + 3], q + 4);
            R4 = __shfl_sync(0xffffffffu, a[2 * t + 4], p + 4); R5 = __shfl_sync(0xffffffffu, a[2 * t + 4], q + 4);
            R6 = __shfl_sync(0xffffffffu, a[2 * t + 5], q + 4);
          }
        }
      }
      if (isbad && lane == 0) *bad = 1;
      if (lane >= o && lane < o + PW) rinv[lane] = myrinv;
#pragma unroll
      for (int c = 0; c < PW; c++) S[lane * TLD + o + c] = a[c];
    }
    __syncthreads();
    const int nrow = TB - PW - o;                       // rows / columns right of the panel
    if (nrow > 0) {
      for (int e = tid; e < nrow * nrow; e += 256) {
        const int i = e / nrow, j = e - i * nrow;
        if (j > i) continue;
        const double* xi = S + (o + PW + i) * TLD + o;
        const double* xj = S + (o + PW + j) * TLD + o;
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
        for (int k = 0; k < PW; k += 2) { acc0 = fma(xi[k], xj[k], acc0); acc1 = fma(xi[k + 1], xj[k + 1], acc1); }
        S[(o + PW + i) * TLD + o + PW + j] -= acc0 + acc1;
      }
      __syncthreads();
    }
  }
}


template <int V>
__global__ void __launch_bounds__(256) k_bench(const double* A, double* Lout, long long* cycles)
{
  __shared__ __align__(16) double S[TB * TLD];
  __shared__ __align__(16) double X[TB * TLD];
  __shared__ __align__(16) double xs[64 + 256 + 128];
  __shared__ int s_bad;
  const int tid = threadIdx.x;
  long long tot = 0;
  for (int rep = 0; rep < NREP; rep++) {
    for (int e = tid; e < TB * TB; e += 256) S[(e >> 5) * TLD + (e & 31)] = A[e];
    if (tid == 0) s_bad = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (V == 0) potrf32_panel<true, 16>(S, xs, tid, &s_bad);
    else if (V == 1) potrf32_pairs<16>(S, xs, xs + 64, tid, &s_bad);
    else if (V == 2) potrf32_pairs<32>(S, xs, xs + 64, tid, &s_bad);
    else if (V == 3) potrf32_pairs<8>(S, xs, xs + 64, tid, &s_bad);
    else if (V == 4) { potrf32_pairs<16>(S, xs, xs + 64, tid, &s_bad); inverse32_block(S, X, xs, xs + 192, tid); }
    else if (V == 5) { trsm32_rt(X, S, tid); }
    else if (V == 6) potrf32_pairs_x<16, 1>(S, xs, xs + 64, tid, &s_bad);
    else if (V == 7) potrf32_pairs_x<16, 2>(S, xs, xs + 64, tid, &s_bad);
    else if (V == 8) potrf32_pairs_x<32, 2>(S, xs, xs + 64, tid, &s_bad);
    else if (V == 9) potrf32_lazy(S, xs, xs + 64, tid, &s_bad);
    else if (V == 10) { potrf32_lazy(S, xs, xs + 64, tid, &s_bad); blockinv8_inplace(S, xs, tid); __syncthreads(); }
    else if (V == 11) { trsm32_blk(X, S, tid); }
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
  // shuffle with a data-dependent source lane (cannot be folded)
  { int src = tid & 31; __syncthreads(); t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) { x = __shfl_sync(0xffffffffu, x, src); src = (src + (int)(x > 1e300)) & 31; x = __shfl_sync(0xffffffffu, x, src + 0); src = (src + (int)(x > 1e300)) & 31; x = __shfl_sync(0xffffffffu, x, src); src = (src + (int)(x > 1e300)) & 31; x = __shfl_sync(0xffffffffu, x, src); src = (src + (int)(x > 1e300)) & 31; }
    t1 = clock64(); if (tid == 0) out[8] = (t1 - t0) / 256; }
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
  const char* names[] = { "V0 round-1 chain: 16-column panels, one pivot per step, fast rsqrt", "V1 pivot pairs, 16-column panels (shipped)", "V2 pivot pairs, one 32-column panel",
                          "V3 pivot pairs, 8-column panels", "V4 = V1 + inverse32_block", "V5 trsm32_rt only (garbage in, timing only)", "V6 timing only: V1 without the look-ahead re-fetch", "V7 timing only: V1 chain alone (no rank-2 update)", "V8 timing only: 32-col chain alone", "V9 potrf32_lazy (shipped)", "V10 = V9 + blockinv8_inplace (result is not L: error column meaningless)", "V11 trsm32_blk only (garbage in, timing only)" };
  std::vector<double> L(n * n);
  for (int v = 0; v < 12; v++) {
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
      case 9: k_bench<9><<<1, 256>>>(dA, dL, dc); break;
      case 10: k_bench<10><<<1, 256>>>(dA, dL, dc); break;
      case 11: k_bench<11><<<1, 256>>>(dA, dL, dc); break;
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
  printf("shfl(double, data-dependent lane) + compare: %lld cycles\n", lat[8]);
  return 0;
}
