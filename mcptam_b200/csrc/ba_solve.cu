// ba_solve.cu — dense solve of the damped Schur-reduced camera system  (H0 + lambda I - Sm) dc = gc - rm.
//
// Replaces CHOLMOD inside g2o (reference src/ChainBundle.cc:1156) for the pose block.  The matrix is small
// (6N = 294 at 200 KF, 744 at 1000 KF), so the factorisation is latency bound: it is organised as a tile
// dataflow (32x32 fp64 tiles, left-looking): every tile of L is one task owned by one persistent CTA which
// accumulates  A_ij - sum_k L_ik L_jk^T  as the tiles of earlier block columns become ready (acquire/release
// flags in global memory, L2-resident), then either factors it (diagonal: register Cholesky by one warp +
// explicit inverse) or multiplies with the inverse diagonal factor (off-diagonal).  The right-hand side rides
// along as an extra block row, so the forward substitution is part of the same dataflow; the CTA that retires
// the last task does the backward substitution, the SE3 pose update (VertexPoseSE3::oplusImpl,
// src/ChainBundle.cc:82-86) and g2o's computeScale() partial sums.
#include "ba_types.cuh"

namespace mcp {

constexpr int TB = 32;            // tile edge
constexpr int TLD = TB + 1;       // padded smem stride

__device__ __forceinline__ int ld_acquire(const int* p)
{
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v)
{
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ size_t tile_index(int i, int j) { return (size_t)i * (i + 1) / 2 + j; }

// acc(2x2 per thread) -= As * Bs^T over the 32-wide k range
__device__ __forceinline__ void tile_mm_sub(const double* As, const double* Bs, int ty, int tx, double acc[2][2])
{
#pragma unroll 8
  for (int k = 0; k < TB; k++) {
    const double a0 = As[ty * TLD + k], a1 = As[(ty + 16) * TLD + k];
    const double b0 = Bs[tx * TLD + k], b1 = Bs[(tx + 16) * TLD + k];
    acc[0][0] -= a0 * b0; acc[0][1] -= a0 * b1;
    acc[1][0] -= a1 * b0; acc[1][1] -= a1 * b1;
  }
}


// Cholesky of a 32x32 SPD tile held in shared memory (stride TLD), lower triangle in/out, upper zeroed.
// One warp, lane = row, the row lives in registers; column j is broadcast with shuffles.
// rinv[j] = 1 / L[j][j].  Returns true if a non-positive pivot was met (pivot replaced by 1).
__device__ __noinline__ bool potrf32_warp(double* S, double* rinv, int lane)
{
  double a[TB];
#pragma unroll
  for (int c = 0; c < TB; c++) a[c] = (c <= lane) ? S[lane * TLD + c] : 0.0;
  bool bad = false;
  // software pipelined: the pivot of column jj+1 (and its rsqrt) is produced right after the first
  // rank-1 update of column jj, before the remaining 30-jj column updates are issued.
  double dj = __shfl_sync(0xffffffffu, a[0], 0);
  if (!(dj > 0.0) || !(dj < 1.0e300)) { bad = true; dj = 1.0; }
  double ri = rsqrt(dj);
#pragma unroll
  for (int jj = 0; jj < TB; jj++) {
    if (lane == jj) { rinv[jj] = ri; a[jj] = dj * ri; }
    else if (lane > jj) a[jj] *= ri;
    double dn = 1.0, rn = 1.0;
    if (jj + 1 < TB) {
      const double v1 = __shfl_sync(0xffffffffu, a[jj], jj + 1);
      if (lane >= jj + 1) a[jj + 1] -= a[jj] * v1;
      dn = __shfl_sync(0xffffffffu, a[jj + 1], jj + 1);
      if (!(dn > 0.0) || !(dn < 1.0e300)) { bad = true; dn = 1.0; }
      rn = rsqrt(dn);
    }
#pragma unroll
    for (int c = jj + 2; c < TB; c++) {
      const double v = __shfl_sync(0xffffffffu, a[jj], c);
      if (lane >= c) a[c] -= a[jj] * v;
    }
    dj = dn; ri = rn;
  }
#pragma unroll
  for (int c = 0; c < TB; c++) S[lane * TLD + c] = a[c];
  return bad;
}


// Blocked Cholesky of the 32x32 tile in shared memory S (stride TLD) by the whole CTA (256 threads):
// four 8-column panels; the 8x8 diagonal block is factored (and inverted) in registers by 8 lanes of warp 0,
// the panel below and the trailing update are small matrix products spread over all threads.  The dependent
// chain is 32 pivots (rsqrt) long instead of 32 pivots + 496 serial rank-1 column updates in one warp.
// rinv[32]: reciprocal diagonal; linv8: 64 doubles scratch; tmp: >= 256 doubles scratch.  Returns via *bad.
__device__ __forceinline__ void potrf32_blocked(double* S, double* rinv, double* linv8, double* tmp, int tid, int* bad)
{
  const int lane = tid & 31, wid = tid >> 5;
  for (int b = 0; b < 4; b++) {
    const int o = 8 * b;
    if (wid == 0) {
      const int r = lane & 7;
      double a[8];
#pragma unroll
      for (int c = 0; c < 8; c++) a[c] = (c <= r) ? S[(o + r) * TLD + o + c] : 0.0;
      // branch-free pivot chain: l_jj = d_j * rsqrt(d_j) is the same multiply as the column scaling, bad pivots are
      // handled with selects (no divergence / reconvergence inside the dependent chain)
      bool isbad = false;
      double dj = __shfl_sync(0xffffffffu, a[0], 0);
      {
        const bool ok = (dj > 0.0) && (dj < 1.0e300);
        isbad |= !ok;
        dj = ok ? dj : 1.0;
      }
      double ri = rsqrt(dj);
#pragma unroll
      for (int jj = 0; jj < 8; jj++) {
        a[jj] = (r >= jj) ? a[jj] * ri : a[jj];
        if (lane == jj) rinv[o + jj] = ri;
        double rn = 1.0;
        if (jj + 1 < 8) {
          const double v1 = __shfl_sync(0xffffffffu, a[jj], jj + 1);
          a[jj + 1] = (r >= jj + 1) ? a[jj + 1] - a[jj] * v1 : a[jj + 1];
          double dn = __shfl_sync(0xffffffffu, a[jj + 1], jj + 1);
          const bool ok = (dn > 0.0) && (dn < 1.0e300);
          isbad |= !ok;
          dn = ok ? dn : 1.0;
          rn = rsqrt(dn);
        }
#pragma unroll
        for (int c = jj + 2; c < 8; c++) {
          const double v = __shfl_sync(0xffffffffu, a[jj], c);
          a[c] = (r >= c) ? a[c] - a[jj] * v : a[c];
        }
        ri = rn;
      }
      if (isbad && lane == 0) *bad = 1;
      if (lane < 8) {
#pragma unroll
        for (int c = 0; c < 8; c++) S[(o + r) * TLD + o + c] = a[c];
      }
      __syncwarp();
      if (lane < 8) {            // inverse of the 8x8 factor: lane = column
        const int c = lane;
        double x[8];
#pragma unroll
        for (int rr = 0; rr < 8; rr++) {
          double sacc = (rr == c) ? 1.0 : 0.0;
#pragma unroll
          for (int k = 0; k < rr; k++) sacc -= S[(o + rr) * TLD + o + k] * ((k >= c) ? x[k] : 0.0);
          x[rr] = (rr >= c) ? sacc * rinv[o + rr] : 0.0;
        }
#pragma unroll
        for (int rr = 0; rr < 8; rr++) linv8[rr * 8 + c] = x[rr];
      }
    }
    __syncthreads();
    const int nrow = 24 - o;                       // rows below the diagonal block
    if (nrow > 0) {
      // panel: X[i][c] = sum_{k<=c} A[i][o+k] * Linv[c][k]
      double xv = 0.0;
      const int pi = tid >> 3, pc = tid & 7;
      if (pi < nrow) {
        const double* arow = S + (o + 8 + pi) * TLD + o;
#pragma unroll
        for (int k = 0; k < 8; k++) xv += arow[k] * linv8[pc * 8 + k];
      }
      __syncthreads();
      if (pi < nrow) S[(o + 8 + pi) * TLD + o + pc] = xv;
      __syncthreads();
      // trailing update of the lower triangle: A[i][j] -= sum_k X[i][k] X[j][k]
      for (int e = tid; e < nrow * nrow; e += 256) {
        const int i = e / nrow, j = e - i * nrow;
        if (j > i) continue;
        const double* xi = S + (o + 8 + i) * TLD + o;
        const double* xj = S + (o + 8 + j) * TLD + o;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) acc += xi[k] * xj[k];
        S[(o + 8 + i) * TLD + o + 8 + j] -= acc;
      }
      __syncthreads();
    }
  }
  // zero the strict upper triangle (the tile is consumed as a full 32x32 lower-triangular factor)
  for (int e = tid; e < TB * TB; e += 256) { const int r = e >> 5, c = e & 31; if (c > r) S[r * TLD + c] = 0.0; }
  (void)tmp;
  __syncthreads();
}

// 1/sqrt(d) for a validated pivot: MUFU.RSQ64H seed (rsqrt.approx.ftz.f64, ~2^-20) + one cubic Newton step
// y (1 + e/2 + 3e^2/8), e = 1 - d y^2  ->  relative error ~2^-58, branch free (rsqrt() carries special-case
// branches that keep the compiler from interleaving the pivot chain with the column updates).
__device__ __forceinline__ double fast_rsqrt(double d)
{
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double dy = d * y;
  const double e = fma(-dy, y, 1.0);
  const double p = fma(0.375, e, 0.5);
  return fma(y * e, p, y);
}

// Cholesky of the 32x32 tile in shared memory S (stride TLD) by the whole CTA (256 threads), four 8-column panels.
// Warp 0 (lane = row) runs the pivot chain of a panel over ALL rows below the diagonal, so the panel of L leaves
// the chain finished (no separate 8x8 inverse / triangular solve); the trailing update of the remaining columns is
// a rank-8 update spread over all threads.  Two barriers per panel.  rinv[32]: reciprocal diagonal.
template <bool FAST, int PW = 8>
__device__ __forceinline__ void potrf32_panel(double* S, double* rinv, int tid, int* bad)
{
  const int lane = tid & 31, wid = tid >> 5;
#pragma unroll 1
  for (int b = 0; b < TB / PW; b++) {
    const int o = PW * b;
    if (wid == 0) {
      double a[PW];
#pragma unroll
      for (int c = 0; c < PW; c++) a[c] = S[lane * TLD + o + c];
      bool isbad = false;
      double dj = __shfl_sync(0xffffffffu, a[0], o);
      {
        const bool ok = (dj > 1.0e-290) && (dj < 1.0e290);
        isbad |= !ok;
        dj = ok ? dj : 1.0;
      }
      double ri = FAST ? fast_rsqrt(dj) : rsqrt(dj);
#pragma unroll
      for (int jj = 0; jj < PW; jj++) {
        a[jj] = (lane >= o + jj) ? a[jj] * ri : 0.0;          // rows above the diagonal: strict upper triangle := 0
        if (lane == o + jj) rinv[o + jj] = ri;
        double rn = 1.0;
        if (jj + 1 < PW) {
          const double v1 = __shfl_sync(0xffffffffu, a[jj], o + jj + 1);
          a[jj + 1] = fma(-a[jj], v1, a[jj + 1]);
          double dn = __shfl_sync(0xffffffffu, a[jj + 1], o + jj + 1);
          const bool ok = (dn > 1.0e-290) && (dn < 1.0e290);
          isbad |= !ok;
          dn = ok ? dn : 1.0;
          rn = FAST ? fast_rsqrt(dn) : rsqrt(dn);
        }
#pragma unroll
        for (int c = jj + 2; c < PW; c++) {
          const double v = __shfl_sync(0xffffffffu, a[jj], o + c);
          a[c] = fma(-a[jj], v, a[c]);
        }
        ri = rn;
      }
      if (isbad && lane == 0) *bad = 1;
#pragma unroll
      for (int c = 0; c < PW; c++) S[lane * TLD + o + c] = a[c];
    }
    __syncthreads();
    const int nrow = TB - PW - o;                       // rows / columns right of the panel
    if (nrow > 0) {
      // trailing update of the lower triangle: A[i][j] -= sum_k L[i][o+k] L[j][o+k]
      for (int e = tid; e < nrow * nrow; e += 256) {
        const int i = e / nrow, j = e - i * nrow;
        if (j > i) continue;
        const double* xi = S + (o + PW + i) * TLD + o;
        const double* xj = S + (o + PW + j) * TLD + o;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < PW; k++) acc = fma(xi[k], xj[k], acc);
        S[(o + PW + i) * TLD + o + PW + j] -= acc;
      }
      __syncthreads();
    }
  }
  // rows < o of later panels were zeroed by the chain's own select; nothing else to clear
}

// Cholesky of the 32x32 tile by 8 warps: warp w owns columns 4w..4w+3, lane = row, the tile lives in registers.
// Per pivot: the owning warp scales its column and posts it to shared memory, ONE barrier, every warp applies the
// rank-1 update to its own (at most four) columns.  colbuf: 64 doubles (double buffered column).
template <bool FAST>
__device__ __forceinline__ void potrf32_cols(double* S, double* rinv, double* colbuf, int tid, int* bad)
{
  const int lane = tid & 31, wid = tid >> 5;
  double a[4];
#pragma unroll
  for (int q = 0; q < 4; q++) a[q] = S[lane * TLD + 4 * wid + q];
  bool isbad = false;
#pragma unroll
  for (int j = 0; j < TB; j++) {
    const int ow = j >> 2, oq = j & 3;
    double* cb = colbuf + 32 * (j & 1);
    if (wid == ow) {
      double dj = __shfl_sync(0xffffffffu, a[oq], j);
      const bool ok = (dj > 1.0e-290) && (dj < 1.0e290);
      isbad |= !ok;
      dj = ok ? dj : 1.0;
      const double ri = FAST ? fast_rsqrt(dj) : rsqrt(dj);
      a[oq] = (lane >= j) ? a[oq] * ri : 0.0;
      cb[lane] = a[oq];
      if (lane == j) rinv[j] = ri;
    }
    __syncthreads();
    if (wid >= ow) {
      const double lr = cb[lane];
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (4 * wid + q > j) a[q] = fma(-lr, cb[4 * wid + q], a[q]);
    }
  }
  if (isbad && lane == 0) *bad = 1;
#pragma unroll
  for (int q = 0; q < 4; q++) S[lane * TLD + 4 * wid + q] = a[q];
  __syncthreads();
}

// Inverse of the lower-triangular 32x32 factor L (shared, stride TLD) into X (shared, stride TLD), by
// recursive 2x2 blocking: [A 0; B C]^-1 = [A^-1 0; -C^-1 B A^-1, C^-1] with 8x8 leaves.  256 threads.
// rinv: reciprocals of the diagonal of L; tmp: >= 256 doubles of scratch.
__device__ __forceinline__ void inverse32_block(const double* L, double* X, const double* rinv, double* tmp, int tid)
{
  for (int e = tid; e < TB * TLD; e += 256) X[e] = 0.0;
  __syncthreads();
  if (tid < 32) {
    // leaf: column c of the inverse of diagonal 8x8 block b
    const int b = tid >> 3, c = tid & 7, o = 8 * b;
    double x[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
      double s = (r == c) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < r; k++) s -= L[(o + r) * TLD + o + k] * ((k >= c) ? x[k] : 0.0);
      x[r] = (r >= c) ? s * rinv[o + r] : 0.0;
    }
#pragma unroll
    for (int r = 0; r < 8; r++) X[(o + r) * TLD + o + c] = x[r];
  }
  __syncthreads();
  // level 1: 8x8 off-diagonal blocks of the two 16x16 diagonal blocks:  X10 = -D1inv * (L10 * D0inv)
  if (tid < 128) {
    const int h = tid >> 6, e = tid & 63, r = e >> 3, c = e & 7, o = 16 * h;
    double m = 0.0;
#pragma unroll
    for (int k = 0; k < 8; k++) m += L[(o + 8 + r) * TLD + o + k] * X[(o + k) * TLD + o + c];
    tmp[tid] = m;
  }
  __syncthreads();
  if (tid < 128) {
    const int h = tid >> 6, e = tid & 63, r = e >> 3, c = e & 7, o = 16 * h;
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < 8; k++) v += X[(o + 8 + r) * TLD + o + 8 + k] * tmp[h * 64 + k * 8 + c];
    X[(o + 8 + r) * TLD + o + c] = -v;
  }
  __syncthreads();
  // level 2: 16x16 off-diagonal block:  X[16:32,0:16] = -Cinv * (B * Ainv)
  {
    const int r = tid >> 4, c = tid & 15;
    double m = 0.0;
#pragma unroll
    for (int k = 0; k < 16; k++) m += L[(16 + r) * TLD + k] * X[k * TLD + c];
    tmp[tid] = m;
  }
  __syncthreads();
  {
    const int r = tid >> 4, c = tid & 15;
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < 16; k++) v += X[(16 + r) * TLD + 16 + k] * tmp[k * 16 + c];
    __syncthreads();
    X[(16 + r) * TLD + c] = -v;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) k_chol_solve(BaDev d, int epoch, int base0, int base1)
{
  pdl_prologue(false);
  __shared__ double As[TB * TLD];
  __shared__ double Bs[TB * TLD];
  __shared__ double xs[32 * TB + 64];     // backsolve vector (up to 32 block rows) + scratch
  __shared__ double red[32];
  __shared__ int s_task;
  const int n = d.nc;
  const int T = (n + TB - 1) / TB;
  const int n_tasks = T * (T + 1) / 2 + T;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int ty = tid >> 4, tx = tid & 15;
  BaCtrl* ctrl = d.ctrl;
  const double lambda = trial_lambda(d);
  double* Lt = d.L;
  double* Linv = d.Linv;
  int* ready = d.flags;
  int* inv_ready = d.flags + n_tasks;
  int* ctr = d.flags + n_tasks + T;      // [0] task counter, [1] done counter, [2] fail

  for (;;) {
    if (tid == 0) s_task = atomicAdd(&ctr[0], 1) - base0;
    __syncthreads();
    const int task = s_task;
    if (task >= n_tasks) break;
    // task -> (i, j): column-major over block columns; rows j..T (row T = right-hand side)
    int j = 0, rem = task;
    while (rem >= T + 1 - j) { rem -= T + 1 - j; j++; }
    const int i = j + rem;
    const bool is_rhs = (i == T);
    unsigned long long t0 = 0, t1 = 0, t2 = 0, t3p = 0, t4p = 0;
    if (d.dbg && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    // ---- initial value ------------------------------------------------------------------------------
    double acc[2][2];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
      for (int b = 0; b < 2; b++) {
        const int r = ty + 16 * a, c = tx + 16 * b;
        const int gc_ = TB * j + c;
        double v = 0.0;
        if (is_rhs) {
          if (r == 0 && gc_ < n) v = d.gc[gc_] - d.rm[gc_];
        } else {
          const int gr = TB * i + r;
          if (gr < n && gc_ < n) {
            if (gc_ <= gr) {
              v = d.H0[(size_t)gc_ * n + gr] - d.Sm[(size_t)gc_ * n + gr];
              if (gr == gc_) v += lambda;
            }
          } else if (gr == gc_) v = 1.0;       // identity padding
        }
        acc[a][b] = v;
      }
    // ---- left-looking updates -----------------------------------------------------------------------
    if (i == j && j > 0) {
      // Diagonal tile: the critical path of the factorisation runs  potrf(j-1) -> L_{j,j-1} -> potrf(j).
      // Instead of waiting for the separate (j, j-1) task to publish L_{j,j-1}, this CTA accumulates that
      // tile as well and applies Linv_{j-1} itself, so the chain waits on inv_ready[j-1] only.
      double acc2[2][2];
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++) {
          const int gr = TB * j + ty + 16 * a, gc_ = TB * (j - 1) + tx + 16 * b;
          acc2[a][b] = (gr < n) ? d.H0[(size_t)gc_ * n + gr] - d.Sm[(size_t)gc_ * n + gr] : 0.0;
        }
      for (int k = 0; k < j - 1; k++) {
        const size_t ta = tile_index(j, k), tb = tile_index(j - 1, k);
        if (tid == 0) {
          while (ld_acquire(&ready[ta]) != epoch) { }
          while (ld_acquire(&ready[tb]) != epoch) { }
        }
        __syncthreads();
        const double* ga = Lt + ta * (TB * TB);
        const double* gb = Lt + tb * (TB * TB);
        for (int e = tid; e < TB * TB; e += 256) {
          const int r = e >> 5, c = e & 31;
          As[r * TLD + c] = __ldcg(ga + e);
          Bs[r * TLD + c] = __ldcg(gb + e);
        }
        __syncthreads();
        tile_mm_sub(As, Bs, ty, tx, acc2);
        tile_mm_sub(As, As, ty, tx, acc);
        __syncthreads();
      }
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++) As[(ty + 16 * a) * TLD + tx + 16 * b] = acc2[a][b];
      // X = acc2 * L_{j-1,j-1}^-T by forward substitution against the factor itself (published with the reciprocal
      // pivots on its diagonal) -- the chain does not wait for the explicit inverse.  8 threads per row of X,
      // thread (row, g) keeps columns g, g+8, g+16, g+24; the solved column is broadcast with one shuffle.
      if (tid == 0) { while (ld_acquire(&ready[tile_index(j - 1, j - 1)]) != epoch) { } }
      __syncthreads();
      const double* gd = Lt + tile_index(j - 1, j - 1) * (TB * TB);
      for (int e = tid; e < TB * TB; e += 256) Bs[(e >> 5) * TLD + (e & 31)] = __ldcg(gd + e);
      const int xr = 4 * wid + (lane >> 3), xg = lane & 7;
      double a4[4];
#pragma unroll
      for (int q = 0; q < 4; q++) a4[q] = As[xr * TLD + xg + 8 * q];
      __syncthreads();
#pragma unroll
      for (int c = 0; c < TB; c++) {
        const double xv = __shfl_sync(0xffffffffu, a4[c >> 3] * Bs[c * TLD + c], (lane & 24) | (c & 7));
        if (xg == (c & 7)) As[xr * TLD + c] = xv;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int k = xg + 8 * q;
          if (8 * q + 7 > c) a4[q] = (k > c) ? fma(-Bs[k * TLD + c], xv, a4[q]) : a4[q];
        }
      }
      __syncthreads();
      tile_mm_sub(As, As, ty, tx, acc);        // acc -= X X^T
      __syncthreads();
    } else
    for (int k = 0; k < j; k++) {
      const size_t ta = tile_index(i, k), tb = tile_index(j, k);
      if (tid == 0) {
        while (ld_acquire(&ready[ta]) != epoch) { }
        while (ld_acquire(&ready[tb]) != epoch) { }
      }
      __syncthreads();
      const double* ga = Lt + ta * (TB * TB);
      const double* gb = Lt + tb * (TB * TB);
      for (int e = tid; e < TB * TB; e += 256) {
        const int r = e >> 5, c = e & 31;
        As[r * TLD + c] = __ldcg(ga + e);
        Bs[r * TLD + c] = __ldcg(gb + e);
      }
      __syncthreads();
      tile_mm_sub(As, Bs, ty, tx, acc);
      __syncthreads();
    }
    if (d.dbg && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (i == j) {
      // ---- diagonal tile: Cholesky in registers (warp 0, lane = row), then explicit inverse ------------
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++) As[(ty + 16 * a) * TLD + tx + 16 * b] = acc[a][b];
      __syncthreads();
      {
        __shared__ int s_bad;
        if (tid == 0) s_bad = 0;
        __syncthreads();
        potrf32_panel<true, 16>(As, xs + 32 * TB, tid, &s_bad);
        if (tid == 0 && s_bad) atomicExch(&ctr[2], epoch);
      }
      if (d.dbg && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t3p));
      // publish the factor first (reciprocal pivots on the diagonal): the next diagonal task solves against it directly
      double* gl = Lt + tile_index(j, j) * (TB * TB);
      for (int e = tid; e < TB * TB; e += 256) {
        const int r = e >> 5, c = e & 31;
        gl[e] = (r == c) ? xs[32 * TB + r] : As[r * TLD + c];
      }
      __syncthreads();
      if (tid == 0) { __threadfence(); st_release(&ready[tile_index(j, j)], epoch); }
      // explicit inverse for the off-diagonal tiles of this block column and the back substitution
      inverse32_block(As, Bs, xs + 32 * TB, xs, tid);
      if (d.dbg && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t4p));
      __syncthreads();
      double* gi = Linv + (size_t)j * (TB * TB);
      for (int e = tid; e < TB * TB; e += 256) gi[e] = Bs[(e >> 5) * TLD + (e & 31)];
      __syncthreads();
      if (tid == 0) { __threadfence(); st_release(&inv_ready[j], epoch); }
    } else {
      // ---- off-diagonal / rhs tile: X = acc * Linv_j^T ------------------------------------------------
      if (tid == 0) { while (ld_acquire(&inv_ready[j]) != epoch) { } }
      __syncthreads();
      const double* gi = Linv + (size_t)j * (TB * TB);
      for (int e = tid; e < TB * TB; e += 256) Bs[(e >> 5) * TLD + (e & 31)] = __ldcg(gi + e);
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++) As[(ty + 16 * a) * TLD + tx + 16 * b] = acc[a][b];
      __syncthreads();
      double out[2][2] = { { 0, 0 }, { 0, 0 } };
#pragma unroll 8
      for (int k = 0; k < TB; k++) {
        const double a0 = As[ty * TLD + k], a1 = As[(ty + 16) * TLD + k];
        const double b0 = Bs[tx * TLD + k], b1 = Bs[(tx + 16) * TLD + k];
        out[0][0] += a0 * b0; out[0][1] += a0 * b1;
        out[1][0] += a1 * b0; out[1][1] += a1 * b1;
      }
      double* gx = Lt + tile_index(i, j) * (TB * TB);
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++) gx[(ty + 16 * a) * TB + tx + 16 * b] = out[a][b];
      __syncthreads();
      if (tid == 0) { __threadfence(); st_release(&ready[tile_index(i, j)], epoch); }
    }
    // ---- retire --------------------------------------------------------------------------------------
    if (d.dbg && tid == 0) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t2));
      double* o = d.dbg + 8 * (size_t)task;
      o[0] = i; o[1] = j; o[2] = (double)(t0 % 1000000000ull); o[3] = (double)(t1 % 1000000000ull); o[4] = (double)(t2 % 1000000000ull); o[5] = blockIdx.x; o[6] = (double)(t3p % 1000000000ull); o[7] = (double)(t4p % 1000000000ull);
    }
    __syncthreads();
    if (tid == 0) s_task = atomicAdd(&ctr[1], 1) - base1;
    __syncthreads();
    if (s_task != n_tasks - 1) continue;
    // ======== last task retired: backward substitution, pose update, scalars (this CTA only) ============
    __threadfence();
    for (int k = T - 1; k >= 0; k--) {
      const int c = lane;
      double* scratch = xs + 32 * TB;            // 64 doubles
      // prefetch what does not depend on the x blocks still being formed: four rows of Linv_k and y_k
      const double* gi = Linv + (size_t)k * (TB * TB);
      double li[4];
#pragma unroll
      for (int q = 0; q < 4; q++) li[q] = __ldcg(gi + (wid * 4 + q) * TB + c);
      const double yk = (wid == 0) ? __ldcg(Lt + tile_index(T, k) * (TB * TB) + c) : 0.0;   // row 0 of the rhs tile
      // s[c] = sum_{i>k} sum_r L_ik[r][c] * x_i[r]   (rows split over the warps, loads batched four tiles deep)
      double part = 0.0;
      int ii = k + 1;
      for (; ii + 3 < T; ii += 4) {
        double v[16];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const double* g = Lt + tile_index(ii + u, k) * (TB * TB);
#pragma unroll
          for (int q = 0; q < 4; q++) v[u * 4 + q] = __ldcg(g + (wid + 8 * q) * TB + c);
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
          for (int q = 0; q < 4; q++) part += v[u * 4 + q] * xs[(ii + u) * TB + wid + 8 * q];
      }
      for (; ii < T; ii++) {
        const double* g = Lt + tile_index(ii, k) * (TB * TB);
        double v[4];
#pragma unroll
        for (int q = 0; q < 4; q++) v[q] = __ldcg(g + (wid + 8 * q) * TB + c);
#pragma unroll
        for (int q = 0; q < 4; q++) part += v[q] * xs[ii * TB + wid + 8 * q];
      }
      __syncthreads();
      As[wid * TLD + c] = part;
      __syncthreads();
      if (wid == 0) {
        double sacc = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) sacc += As[w * TLD + c];
        scratch[c] = yk - sacc;
      }
      __syncthreads();
      {
        double p2 = 0;                             // (Linv^T t)[c], rows 4*wid .. 4*wid+3
#pragma unroll
        for (int q = 0; q < 4; q++) p2 += li[q] * scratch[wid * 4 + q];
        Bs[wid * TLD + c] = p2;
      }
      __syncthreads();
      if (wid == 0) {
        double sacc = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) sacc += Bs[w * TLD + c];
        xs[k * TB + c] = sacc;
      }
      __syncthreads();
    }
    if (d.dbg && tid == 0) { unsigned long long t3; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t3)); d.dbg[8 * (size_t)n_tasks] = (double)(t3 % 1000000000ull); }
    const int ok = (ld_acquire(&ctr[2]) != epoch) && ctrl->solve_ok[d.cand];
    double sc = 0, sq = 0;
    for (int e = tid; e < n; e += 256) {
      const double xi = ok ? xs[e] : 0.0;
      d.dc[e] = xi;
      sc += xi * (lambda * xi + d.gc[e]);
      sq += xi * xi;
    }
    sc = warp_sum(sc); sq = warp_sum(sq);
    __syncthreads();
    if (lane == 0) { red[wid] = sc; red[8 + wid] = sq; }
    __syncthreads();
    if (tid == 0) {
      double a = 0, b = 0;
      for (int w = 0; w < 8; w++) { a += red[w]; b += red[8 + w]; }
      ctrl->scale[d.cand] = a; ctrl->sumsq[d.cand] = b; ctrl->solve_ok[d.cand] = ok;
    }
    __syncthreads();
    const int cur = ctrl->cur;
    for (int p = tid; p < d.n_pose; p += 256) {
      Se3 Tm;
      const double* src = d.pose[cur] + 12 * (size_t)p;
#pragma unroll
      for (int q = 0; q < 9; q++) Tm.R[q] = src[q];
#pragma unroll
      for (int q = 0; q < 3; q++) Tm.t[q] = src[9 + q];
      const int v = d.pose_var[p];
      if (v >= 0) {
        double mu[6];
#pragma unroll
        for (int q = 0; q < 6; q++) mu[q] = ok ? xs[6 * v + q] : 0.0;
        Se3 E, O;
        se3_exp(mu, E);
        se3_mul(E, Tm, O);
        Tm = O;
      }
      se3_store(d.pose[trial_buffer(d, cur)] + 12 * (size_t)p, Tm);
    }
    // fall through to the next (failing) grab so that every CTA consumes exactly one id >= n_tasks
  }
}

size_t chol_tiles_doubles(int nc) { const int T = (nc + TB - 1) / TB; return (size_t)(T * (T + 1) / 2 + T) * TB * TB; }
size_t chol_inv_doubles(int nc) { const int T = (nc + TB - 1) / TB; return (size_t)T * TB * TB; }
size_t chol_flag_ints(int nc) { const int T = (nc + TB - 1) / TB; return (size_t)(T * (T + 1) / 2 + T) + T + 8; }
int chol_max_n() { return 32 * TB; }

// The task and done counters are never reset: launch number `epoch` (1-based) consumes exactly
// n_tasks + grid increments of the task counter and n_tasks of the done counter.
void launch_chol_solve(const BaDev& d, int epoch, int n_sms, cudaStream_t s)
{
  const int T = (d.nc + TB - 1) / TB;
  const int n_tasks = T * (T + 1) / 2 + T;
  int grid = n_tasks < n_sms ? n_tasks : n_sms;      // all CTAs must be co-resident (spin-wait dataflow)
  if (grid < 1) grid = 1;
  launch_chain(k_chol_solve, dim3(grid), dim3(256), 0, s, d, epoch, (epoch - 1) * (n_tasks + grid), (epoch - 1) * n_tasks);
}

}  // namespace mcp
