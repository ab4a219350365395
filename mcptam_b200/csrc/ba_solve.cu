// ba_solve.cu — dense solve of the damped Schur-reduced camera system  (H0 + lambda I - Sm) dc = gc - rm.
//
// Replaces CHOLMOD inside g2o (reference src/ChainBundle.cc:1156) for the pose block.  The matrix is small
// (6N = 294 at 200 KF, 744 at 1000 KF), so the factorisation is bound by its dependency chain, not by flops or
// bandwidth.  It is organised as a tile dataflow (32x32 fp64 tiles, left-looking) over persistent CTAs:
//
//   * every tile of L is one task; its owner accumulates  A_ij - sum_k L_ik L_jk^T  on the fp64 tensor cores
//     (mma.sync.m8n8k4.f64, operand fragments loaded straight from L2) as the tiles of earlier block columns appear;
//   * tiles are handed over in "LL" form: every double travels as two self-validating 8-byte words {half, epoch},
//     so a consumer polls the data itself -- no release fence, no separate ready flag, no second round trip on the
//     critical path (one L2 store-to-load latency per hand-off instead of three);
//   * a diagonal task also forms the sub-diagonal tile L_{j,j-1} itself (triangular solve against L_{j-1,j-1} by
//     substitution), so the chain  potrf(j-1) -> L_{j,j-1} -> potrf(j)  has ONE hand-off per block column;
//   * the 32x32 Cholesky of a diagonal tile runs in one warp with TWO pivots per dependent step: both reciprocal
//     square roots of a 2x2 pivot block (of a and of the 2x2 determinant) are independent, and the head of the next
//     pivot block is advanced redundantly in every lane, which takes the shuffles off the chain;
//   * the right-hand side rides along as an extra block row (forward substitution inside the same dataflow); the CTA
//     that retires the last task does the backward substitution, the SE3 pose update (VertexPoseSE3::oplusImpl,
//     src/ChainBundle.cc:82-86) and g2o's computeScale() partial sums.
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

// Inverse of the lower-triangular 32x32 factor L (shared, stride TLD) into X (shared, stride TLD), by
// recursive 2x2 blocking: [A 0; B C]^-1 = [A^-1 0; -C^-1 B A^-1, C^-1] with 8x8 leaves.  256 threads.
// rinv: reciprocals of the diagonal of L; tmp: >= 256 doubles of scratch.
template <bool LEAF_GIVEN = false>
__device__ __forceinline__ void inverse32_block(const double* L, double* X, const double* rinv, double* tmp, int tid)
{
  for (int e = tid; e < TB * TLD; e += 256) X[e] = 0.0;
  __syncthreads();
  if (LEAF_GIVEN) {
    // L arrives in hand-over form: its diagonal 8x8 blocks already are their inverses
    const int b = tid >> 6, r = (tid >> 3) & 7, c = tid & 7, o = 8 * b;
    X[(o + r) * TLD + o + c] = (c <= r) ? L[(o + r) * TLD + o + c] : 0.0;
  } else if (tid < 32) {
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


// ---------------------------------------------------------------------------------------------
// LL tiles: element e of a tile is the 16-byte line {lo32(v), epoch, hi32(v), epoch}.  Each 8-byte half carries its
// own tag, so a line torn between two launches never validates; the tag is the launch number of this buffer.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ll_store(uint4* line, double v, unsigned epoch)
{
  const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(line), "r"(lo), "r"(epoch), "r"(hi), "r"(epoch) : "memory");
}
__device__ __forceinline__ uint4 ll_load(const uint4* line)
{
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(line) : "memory");
  return v;
}
__device__ __forceinline__ bool ll_ok(const uint4& v, unsigned epoch) { return v.y == epoch && v.w == epoch; }
__device__ __forceinline__ double ll_value(const uint4& v) { return __hiloint2double((int)v.z, (int)v.x); }

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// One warp: C(8 x 16, two m8n8 fragments) -= A(rows m0..m0+7 of tile ta) * B(rows n0..n0+15 of tile tb)^T over the 32-wide
// k range, both operands read as LL lines from L2 (polled until the producer's lines of this launch have landed).
// NB == 2 adds a second B tile (tb2 -> c2) that shares the A fragments (diagonal task: the sub-diagonal tile).
template <int NB>
__device__ __forceinline__ void ll_mm_sub(const uint4* __restrict__ ta, const uint4* __restrict__ tb, const uint4* __restrict__ tb2,
                                          int m0, int n0, int lane, unsigned epoch, double (&c)[2][2], double (&c2)[2][2], bool patient)
{
  const int g = lane >> 2, t = lane & 3;
  const uint4* pa = ta + (m0 + g) * TB + t;
  const uint4* pb = tb + (n0 + g) * TB + t;
  const uint4* pb2 = (NB == 2) ? tb2 + (n0 + g) * TB + t : nullptr;
  if (patient) {
    // a task that was claimed ahead of its inputs waits on ONE line per operand tile (a sector per poll, not the
    // whole fragment set), then validates everything it loads as usual
    if (lane < 2) { const uint4* probe = (lane == 0 ? ta : tb) + TB * TB - 1; while (!ll_ok(ll_load(probe), epoch)) __nanosleep(100); }
    __syncwarp();
  }
  constexpr int KS = (NB == 2) ? 2 : 4;          // k4-steps per batch of loads (register budget: 2 CTAs / SM)
#pragma unroll 1
  for (int part = 0; part < 8 / KS; part++) {
    uint4 va[KS], vb[KS][2], vc[KS][2];
    for (;;) {
      bool ok = true;
#pragma unroll
      for (int s = 0; s < KS; s++) {
        const int ko = 4 * (KS * part + s);
        va[s] = ll_load(pa + ko);
        vb[s][0] = ll_load(pb + ko);
        vb[s][1] = ll_load(pb + 8 * TB + ko);
        if (NB == 2) { vc[s][0] = ll_load(pb2 + ko); vc[s][1] = ll_load(pb2 + 8 * TB + ko); }
      }
#pragma unroll
      for (int s = 0; s < KS; s++) {
        ok = ok && ll_ok(va[s], epoch) && ll_ok(vb[s][0], epoch) && ll_ok(vb[s][1], epoch);
        if (NB == 2) ok = ok && ll_ok(vc[s][0], epoch) && ll_ok(vc[s][1], epoch);
      }
      if (__all_sync(0xffffffffu, ok)) break;
      if (patient) __nanosleep(64);
    }
#pragma unroll
    for (int s = 0; s < KS; s++) {
      const double a = -ll_value(va[s]);
      dmma884(c[0], a, ll_value(vb[s][0]));
      dmma884(c[1], a, ll_value(vb[s][1]));
      if (NB == 2) { dmma884(c2[0], a, ll_value(vc[s][0])); dmma884(c2[1], a, ll_value(vc[s][1])); }
    }
  }
}

// c(8 x 16 per warp) -= X(rows m0..) * X(rows n0..)^T with X in shared memory (stride TLD)
__device__ __forceinline__ void smem_mm_sub(const double* X, int m0, int n0, int lane, double (&c)[2][2])
{
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int s = 0; s < 8; s++) {
    const double a = -X[(m0 + g) * TLD + 4 * s + t];
    dmma884(c[0], a, X[(n0 + g) * TLD + 4 * s + t]);
    dmma884(c[1], a, X[(n0 + 8 + g) * TLD + 4 * s + t]);
  }
}

// whole tile (256 threads) from its LL lines into shared memory, polling until every line carries this launch's tag
template <bool PATIENT>
__device__ __forceinline__ void ll_tile_to_smem(const uint4* __restrict__ tl, double* S, int tid, unsigned epoch)
{
  if (PATIENT) {
    if (tid == 0) { while (!ll_ok(ll_load(tl + TB * TB - 1), epoch)) __nanosleep(100); }
    __syncthreads();
  }
  uint4 v[4];
  for (;;) {
    bool ok = true;
#pragma unroll
    for (int q = 0; q < 4; q++) v[q] = ll_load(tl + tid + 256 * q);
#pragma unroll
    for (int q = 0; q < 4; q++) ok = ok && ll_ok(v[q], epoch);
    if (ok) break;
  }
#pragma unroll
  for (int q = 0; q < 4; q++) { const int e = tid + 256 * q; S[(e >> 5) * TLD + (e & 31)] = ll_value(v[q]); }
}

// X := X * L^-T for the 32x32 tile X (shared, stride TLD) against the lower-triangular factor Lf (shared, stride TLD,
// RECIPROCAL pivots on its diagonal), by forward substitution over the columns: 8 threads per row of X, thread
// (row, g) keeps columns g, g+8, g+16, g+24; the solved column is broadcast with one shuffle.  256 threads.
__device__ __forceinline__ void trsm32_rt(double* __restrict__ X, const double* __restrict__ Lf, int tid)
{
  const int lane = tid & 31, wid = tid >> 5;
  const int xr = 4 * wid + (lane >> 3), xg = lane & 7;
  double a4[4], out[4];
#pragma unroll
  for (int q = 0; q < 4; q++) a4[q] = X[xr * TLD + xg + 8 * q];
#pragma unroll
  for (int c = 0; c < TB; c++) {
    const double xv = __shfl_sync(0xffffffffu, a4[c >> 3] * Lf[c * TLD + c], (lane & 24) | (c & 7));
    if (xg == (c & 7)) out[c >> 3] = xv;          // results stay in registers: no store inside the dependent chain
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int k = xg + 8 * q;
      if (8 * q + 7 > c) a4[q] = (k > c) ? fma(-Lf[k * TLD + c], xv, a4[q]) : a4[q];
    }
  }
#pragma unroll
  for (int q = 0; q < 4; q++) X[xr * TLD + xg + 8 * q] = out[q];
}

// The four 8x8 diagonal blocks of the factor L (shared, stride TLD, zero strict upper triangle) are replaced by their
// inverses (lower triangular again; the reciprocal pivots end up on the diagonal).  This is the form in which a diagonal
// tile is handed on: a consumer's triangular solve then needs no division and no 32-step substitution chain.
// Warp 0 only (tid < 32): thread (b, c) solves column c of block b.
__device__ __forceinline__ void blockinv8_inplace(double* S, const double* rinv, int tid)
{
  if (tid >= 32) return;
  const int b = tid >> 3, c = tid & 7, o = 8 * b;
  double x[8];
#pragma unroll
  for (int r = 0; r < 8; r++) {
    double s0 = (r == c) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k < r; k++) {
      const double t = S[(o + r) * TLD + o + k] * ((k >= c) ? x[k] : 0.0);
      if (k & 1) s1 -= t; else s0 -= t;
    }
    x[r] = (r >= c) ? (s0 + s1) * rinv[o + r] : 0.0;
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 8; r++) if (r >= c) S[(o + r) * TLD + o + c] = x[r];
}

// X := X * L^-T for the 32x32 tile X (shared, stride TLD); Lf is the factor in hand-over form (blockinv8_inplace).
// 8 threads per row of X (one 8-lane group), four block steps:  x_b = (b_b - sum_{a<b} x_a L_ba^T) D_b^-T; inside a
// step the eight partial results are exchanged with independent (pipelined) shuffles.  256 threads, no barrier inside.
__device__ __forceinline__ void trsm32_blk(double* __restrict__ X, const double* __restrict__ Lf, int tid)
{
  const int lane = tid & 31, wid = tid >> 5;
  const int xr = 4 * wid + (lane >> 3), g = lane & 7, gb = lane & 24;
  double bq[4], out[4];
#pragma unroll
  for (int q = 0; q < 4; q++) bq[q] = X[xr * TLD + 8 * q + g];
#pragma unroll
  for (int b = 0; b < 4; b++) {
    double tg[8];
#pragma unroll
    for (int k = 0; k < 8; k++) tg[k] = __shfl_sync(0xffffffffu, bq[b], gb | k);
    const double* dr = Lf + (8 * b + g) * TLD + 8 * b;        // row g of D_b^-1 (zeros right of the diagonal)
    double x0 = 0.0, x1 = 0.0;
#pragma unroll
    for (int k = 0; k < 8; k += 2) { x0 = fma(tg[k], dr[k], x0); x1 = fma(tg[k + 1], dr[k + 1], x1); }
    const double x = x0 + x1;
    out[b] = x;
    if (b < 3) {
      double xa[8];
#pragma unroll
      for (int k = 0; k < 8; k++) xa[k] = __shfl_sync(0xffffffffu, x, gb | k);
#pragma unroll
      for (int q = b + 1; q < 4; q++) {
        const double* lr = Lf + (8 * q + g) * TLD + 8 * b;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k += 2) { s0 = fma(xa[k], lr[k], s0); s1 = fma(xa[k + 1], lr[k + 1], s1); }
        bq[q] -= s0 + s1;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; q++) X[xr * TLD + 8 * q + g] = out[q];
}

// Cholesky of the 32x32 tile in shared memory S (stride TLD), lower triangle in, L out (strict upper triangle zeroed),
// rinv[32] = reciprocal pivots.  Warp 0 (lane = row) runs the pivot chain panel by panel (PW columns in registers);
// pivots are taken two at a time:
//     [a b; b c] = [l11 0; l21 l22][..]^T,  r1 = rsqrt(a), r2 = rsqrt(a c - b^2)   (independent of each other)
//     l11 = a r1, l21 = b r1, 1/l22 = r2 l11
// and the three entries that head the NEXT pivot block are advanced redundantly in every lane from seven replicated
// raw values (fetched by shuffles that overlap the rsqrt latency), so that the dependent chain per pivot pair is
// det -> rsqrt -> a handful of FMAs, with no shuffle or shared-memory round trip on it.  The rank-2 update of the other
// panel columns reads (l_c0, l_c1) pairs from a warp-private staging area (one LDS.128 per column).  The trailing
// update between panels is spread over the whole CTA.  wbuf: >= 128 doubles, 16-byte aligned.
template <int PW>
__device__ __forceinline__ void potrf32_pairs(double* __restrict__ S, double* __restrict__ rinv, double* __restrict__ wbuf, int tid, int* bad)
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
          double* xb = wbuf + 64 * (t & 1);
          *reinterpret_cast<double2*>(xb + 2 * lane) = make_double2(x0, x1);
          __syncwarp();
#pragma unroll
          for (int c = 2 * t + 2; c < PW; c++) {
            const double2 y = *reinterpret_cast<const double2*>(xb + 2 * (o + c));
            a[c] = fma(-x1, y.y, fma(-x0, y.x, a[c]));
          }
          if (t + 2 < PW / 2) {
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

// The pivot-pair chain of potrf32_pairs over ONE 8-column panel held in registers (warp 0, lane = row; columns o..o+7).
__device__ __forceinline__ void pair_chain8(double (&a)[8], int o, int lane, double* __restrict__ wbuf, double& myrinv, bool& isbad)
{
  double App = __shfl_sync(0xffffffffu, a[0], o), Aqp = __shfl_sync(0xffffffffu, a[0], o + 1), Aqq = __shfl_sync(0xffffffffu, a[1], o + 1);
  double R0 = __shfl_sync(0xffffffffu, a[0], o + 2), R1 = __shfl_sync(0xffffffffu, a[1], o + 2);
  double R2 = __shfl_sync(0xffffffffu, a[0], o + 3), R3 = __shfl_sync(0xffffffffu, a[1], o + 3);
  double R4 = __shfl_sync(0xffffffffu, a[2], o + 2), R5 = __shfl_sync(0xffffffffu, a[2], o + 3), R6 = __shfl_sync(0xffffffffu, a[3], o + 3);
#pragma unroll
  for (int t = 0; t < 4; t++) {
    const int p = o + 2 * t, q = p + 1;
    double det = fma(App, Aqq, -Aqp * Aqp);
    {
      const bool ok1 = (App > 1.0e-290) && (App < 1.0e290);
      isbad |= !ok1;
      App = ok1 ? App : 1.0;
      const bool ok2 = ok1 && (det > 1.0e-290 * App) && (det < 1.0e290);
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
    if (t < 3) {
      const double Lp0 = R0 * r1, Lq0 = R2 * r1;
      const double Lp1 = fma(-Lp0, l21, R1) * i22, Lq1 = fma(-Lq0, l21, R3) * i22;
      App = fma(-Lp1, Lp1, fma(-Lp0, Lp0, R4));
      Aqp = fma(-Lq1, Lp1, fma(-Lq0, Lp0, R5));
      Aqq = fma(-Lq1, Lq1, fma(-Lq0, Lq0, R6));
      double* xb = wbuf + 64 * (t & 1);
      *reinterpret_cast<double2*>(xb + 2 * lane) = make_double2(x0, x1);
      __syncwarp();
#pragma unroll
      for (int c = 2 * t + 2; c < 8; c++) {
        const double2 y = *reinterpret_cast<const double2*>(xb + 2 * (o + c));
        a[c] = fma(-x1, y.y, fma(-x0, y.x, a[c]));
      }
      if (t < 2) {
        R0 = __shfl_sync(0xffffffffu, a[2 * t + 2], p + 4); R1 = __shfl_sync(0xffffffffu, a[2 * t + 3], p + 4);
        R2 = __shfl_sync(0xffffffffu, a[2 * t + 2], q + 4); R3 = __shfl_sync(0xffffffffu, a[2 * t + 3], q + 4);
        R4 = __shfl_sync(0xffffffffu, a[2 * t + 4], p + 4); R5 = __shfl_sync(0xffffffffu, a[2 * t + 4], q + 4);
        R6 = __shfl_sync(0xffffffffu, a[2 * t + 5], q + 4);
      }
    }
  }
}

// Cholesky of the 32x32 tile (same contract as potrf32_pairs) with the trailing update taken off the chain: after
// panel b (8 columns) warp 0 brings only the NEXT panel's columns up to date, in registers, and continues with the
// pivot chain; warps 1..7 apply the rest of panel b's update (columns >= 8b+16) meanwhile.  One barrier per panel.
__device__ __forceinline__ void potrf32_lazy(double* __restrict__ S, double* __restrict__ rinv, double* __restrict__ wbuf, int tid, int* bad)
{
  const int lane = tid & 31, wid = tid >> 5;
  double a[8];
  bool isbad = false;
  if (wid == 0) {
#pragma unroll
    for (int c = 0; c < 8; c++) a[c] = S[lane * TLD + c];
  }
#pragma unroll 1
  for (int b = 0; b < 4; b++) {
    const int o = 8 * b;
    if (wid == 0) {
      double myrinv = 0.0;
      pair_chain8(a, o, lane, wbuf, myrinv, isbad);
      if (lane >= o && lane < o + 8) rinv[lane] = myrinv;
#pragma unroll
      for (int c = 0; c < 8; c++) S[lane * TLD + o + c] = a[c];
    }
    __syncthreads();                    // panel b is visible; the deferred part of panel b-1 is complete
    if (b == 3) break;
    if (wid == 0) {
      double an[8];
#pragma unroll
      for (int c = 0; c < 8; c++) {
        const double* lj = S + (o + 8 + c) * TLD + o;
        double s0 = S[lane * TLD + o + 8 + c], s1 = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k += 2) { s0 = fma(-a[k], lj[k], s0); s1 = fma(-a[k + 1], lj[k + 1], s1); }
        an[c] = s0 + s1;
      }
#pragma unroll
      for (int c = 0; c < 8; c++) a[c] = an[c];
    } else {
      const int c0 = o + 16, nr = TB - c0;            // columns c0..31, rows >= column
      for (int e = tid - 32; e < nr * nr; e += 224) {
        const int i = e / nr, j = e - i * nr;
        if (j > i) continue;
        const double* xi = S + (c0 + i) * TLD + o;
        const double* xj = S + (c0 + j) * TLD + o;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k += 2) { s0 = fma(xi[k], xj[k], s0); s1 = fma(xi[k + 1], xj[k + 1], s1); }
        S[(c0 + i) * TLD + c0 + j] -= s0 + s1;
      }
    }
  }
  if (wid == 0 && isbad && lane == 0) *bad = 1;
  __syncthreads();
}

#ifndef MCP_CHOL_PW
#define MCP_CHOL_PW 8
#endif

// c(8 x 16 per warp) -= X(rows m0.., shared memory) * B(rows n0.. of an LL tile)^T
__device__ __forceinline__ void smem_ll_mm_sub(const double* X, const uint4* __restrict__ tb, int m0, int n0, int lane, unsigned epoch, double (&c)[2][2])
{
  const int g = lane >> 2, t = lane & 3;
  const uint4* pb = tb + (n0 + g) * TB + t;
#pragma unroll 1
  for (int part = 0; part < 2; part++) {
    uint4 vb[4][2];
    for (;;) {
      bool ok = true;
#pragma unroll
      for (int s = 0; s < 4; s++) { vb[s][0] = ll_load(pb + 16 * part + 4 * s); vb[s][1] = ll_load(pb + 8 * TB + 16 * part + 4 * s); }
#pragma unroll
      for (int s = 0; s < 4; s++) ok = ok && ll_ok(vb[s][0], epoch) && ll_ok(vb[s][1], epoch);
      if (__all_sync(0xffffffffu, ok)) break;
    }
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const double a = -X[(m0 + g) * TLD + 16 * part + 4 * s + t];
      dmma884(c[0], a, ll_value(vb[s][0]));
      dmma884(c[1], a, ll_value(vb[s][1]));
    }
  }
}

// Worker accumulation of the task that also prepares the next-but-one diagonal step (tile (i, j) with i = j + 2):
//   cw -= L_ik L_jk^T  (its own tile),  cp -= L_ik L_ik^T  (diagonal tile (i, i)),  cs -= L_ik L_{i-1,k}^T  (tile (i, i-1))
__device__ __forceinline__ void ll_mm_sub3(const uint4* __restrict__ ti, const uint4* __restrict__ tj, const uint4* __restrict__ tm,
                                           int m0, int n0, int lane, unsigned epoch, double (&cw)[2][2], double (&cp)[2][2], double (&cs)[2][2])
{
  const int g = lane >> 2, t = lane & 3;
  const uint4* pa = ti + (m0 + g) * TB + t;
  const uint4* p1 = tj + (n0 + g) * TB + t;
  const uint4* p2 = ti + (n0 + g) * TB + t;
  const uint4* p3 = tm + (n0 + g) * TB + t;
  if (lane < 3) { const uint4* probe = (lane == 0 ? ti : lane == 1 ? tj : tm) + TB * TB - 1; while (!ll_ok(ll_load(probe), epoch)) __nanosleep(100); }
  __syncwarp();
#pragma unroll 1
  for (int part = 0; part < 4; part++) {
    uint4 va[2], v1[2][2], v2[2][2], v3[2][2];
    for (;;) {
      bool ok = true;
#pragma unroll
      for (int s = 0; s < 2; s++) {
        const int ko = 4 * (2 * part + s);
        va[s] = ll_load(pa + ko);
        v1[s][0] = ll_load(p1 + ko); v1[s][1] = ll_load(p1 + 8 * TB + ko);
        v2[s][0] = ll_load(p2 + ko); v2[s][1] = ll_load(p2 + 8 * TB + ko);
        v3[s][0] = ll_load(p3 + ko); v3[s][1] = ll_load(p3 + 8 * TB + ko);
      }
#pragma unroll
      for (int s = 0; s < 2; s++)
        ok = ok && ll_ok(va[s], epoch) && ll_ok(v1[s][0], epoch) && ll_ok(v1[s][1], epoch) && ll_ok(v2[s][0], epoch) && ll_ok(v2[s][1], epoch) &&
             ll_ok(v3[s][0], epoch) && ll_ok(v3[s][1], epoch);
      if (__all_sync(0xffffffffu, ok)) break;
      __nanosleep(64);
    }
#pragma unroll
    for (int s = 0; s < 2; s++) {
      const double a = -ll_value(va[s]);
      dmma884(cw[0], a, ll_value(v1[s][0])); dmma884(cw[1], a, ll_value(v1[s][1]));
      dmma884(cp[0], a, ll_value(v2[s][0])); dmma884(cp[1], a, ll_value(v2[s][1]));
      dmma884(cs[0], a, ll_value(v3[s][0])); dmma884(cs[1], a, ll_value(v3[s][1]));
    }
  }
}

// worker tasks of block column j, in this order: tiles (i, j) for i = j+2 .. T-1 (the first one also prepares diagonal
// step j+2), the rhs tile (T, j), the explicit inverse of L_jj (back substitution only)
__device__ __forceinline__ int chol_col_tasks(int T, int j) { return 2 + (T - j - 2 > 0 ? T - j - 2 : 0); }
__device__ __forceinline__ size_t chol_prep_tile(int T, int j, int which) { return (size_t)T * (T + 1) / 2 + T + 2 * (size_t)j + which; }

// entry (r, c) of the damped reduced camera system, tile-local padding: identity
__device__ __forceinline__ double chol_entry(const BaDev& d, int n, int gr, int gc_, double lambda)
{
  if (gr < n && gc_ < n) {
    if (gc_ > gr) return 0.0;
    const double v = d.H0[(size_t)gc_ * n + gr] - d.Sm[(size_t)gc_ * n + gr];
    return (gr == gc_) ? v + lambda : v;
  }
  return (gr == gc_) ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(256) k_chol_solve(BaDev d, unsigned epoch, int base0, int base1)
{
  pdl_prologue(false);
  __shared__ __align__(16) double As[TB * TLD];
  __shared__ __align__(16) double Bs[TB * TLD];
  __shared__ __align__(16) double xs[32 * TB + 64 + 128];     // backsolve vector (up to 32 block rows) + rinv/scratch (64) + potrf staging (128)
  __shared__ double red[32];
  __shared__ int s_task;
  __shared__ int s_bad;
  const int n = d.nc;
  const int T = (n + TB - 1) / TB;
  int n_tasks = 0;
  for (int j = 0; j < T; j++) n_tasks += chol_col_tasks(T, j);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int m0 = 8 * (wid & 3), n0 = 16 * (wid >> 2);          // this warp's 8 x 16 block of the tile (two m8n8 fragments)
  const int fr = m0 + (lane >> 2), fc = n0 + 2 * (lane & 3);   // fragment element (fr, fc + 8 f + {0,1})
  BaCtrl* ctrl = d.ctrl;
  const double lambda = trial_lambda(d);
  double* Lt = d.L;
  double* Linv = d.Linv;
  uint4* Lll = d.Lll;
  int* ctr = d.flags;      // [0] task counter, [1] done counter, [2] fail
  double* rinv = xs + 32 * TB;
  double* wbuf = xs + 32 * TB + 64;

  if (blockIdx.x == 0) {
    // ================= chain CTA: every diagonal tile, in order; the factor of the previous step never leaves shared memory ============
    //   step j:  P(j) = {A_jj - sum_{k<=j-2} L_jk L_jk^T,  A_{j,j-1} - sum_{k<=j-2} L_jk L_{j-1,k}^T}   (prepared by the (j, j-2) worker)
    //            X = P_s L_{j-1,j-1}^-T  -> tile (j, j-1);   L_jj = chol(P_d - X X^T)
    for (int j = 0; j < T; j++) {
      unsigned long long t0 = 0, t1 = 0, t2 = 0;
      double acc[2][2];
      if (j < 2) {
#pragma unroll
        for (int f = 0; f < 2; f++)
#pragma unroll
          for (int e = 0; e < 2; e++) acc[f][e] = chol_entry(d, n, TB * j + fr, TB * j + fc + 8 * f + e, lambda);
        if (j == 1)
#pragma unroll
          for (int q = 0; q < 4; q++) { const int e = tid + 256 * q, r = e >> 5, c = e & 31; Bs[r * TLD + c] = (TB + r < n) ? chol_entry(d, n, TB + r, c, lambda) : 0.0; }
      } else {
        const uint4* pd = Lll + chol_prep_tile(T, j, 0) * (TB * TB);
        uint4 v[2][2];
        for (;;) {
          bool ok = true;
#pragma unroll
          for (int f = 0; f < 2; f++)
#pragma unroll
            for (int e = 0; e < 2; e++) { v[f][e] = ll_load(pd + fr * TB + fc + 8 * f + e); ok = ok && ll_ok(v[f][e], epoch); }
          if (ok) break;
        }
#pragma unroll
        for (int f = 0; f < 2; f++)
#pragma unroll
          for (int e = 0; e < 2; e++) acc[f][e] = ll_value(v[f][e]);
        ll_tile_to_smem<false>(Lll + chol_prep_tile(T, j, 1) * (TB * TB), Bs, tid, epoch);
      }
      if (d.dbg && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      if (j > 0) {
        __syncthreads();
        trsm32_blk(Bs, As, tid);                 // As: L_{j-1,j-1} in hand-over form (inverted 8x8 diagonal blocks)
        __syncthreads();
        uint4* gl = Lll + tile_index(j, j - 1) * (TB * TB);
        double* gp = Lt + tile_index(j, j - 1) * (TB * TB);
#pragma unroll
        for (int q = 0; q < 4; q++) { const int e = tid + 256 * q; const double v = Bs[(e >> 5) * TLD + (e & 31)]; ll_store(gl + e, v, epoch); gp[e] = v; }
        smem_mm_sub(Bs, m0, n0, lane, acc);
      }
      if (d.dbg && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
#pragma unroll
      for (int f = 0; f < 2; f++) { As[fr * TLD + fc + 8 * f] = acc[f][0]; As[fr * TLD + fc + 8 * f + 1] = acc[f][1]; }
      if (tid == 0) s_bad = 0;
      __syncthreads();
      potrf32_lazy(As, rinv, wbuf, tid, &s_bad);
      if (tid == 0 && s_bad) atomicExch(&ctr[2], (int)epoch);
      blockinv8_inplace(As, rinv, tid);
      __syncthreads();
      {
        uint4* gl = Lll + tile_index(j, j) * (TB * TB);
#pragma unroll
        for (int q = 0; q < 4; q++) { const int e = tid + 256 * q; ll_store(gl + e, As[(e >> 5) * TLD + (e & 31)], epoch); }
      }
      if (d.dbg && tid == 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t2));
        double* o = d.dbg + 8 * (size_t)(n_tasks + 1 + j);
        o[0] = j; o[1] = j; o[2] = (double)(t0 % 1000000000ull); o[3] = (double)(t1 % 1000000000ull); o[4] = (double)(t2 % 1000000000ull); o[5] = 0;
      }
    }
    __threadfence();
    return;
  }

  for (;;) {
    if (tid == 0) s_task = atomicAdd(&ctr[0], 1) - base0;
    __syncthreads();
    const int task = s_task;
    if (task < 0 || task >= n_tasks) break;
    int j = 0, rem = task;
    while (rem >= chol_col_tasks(T, j)) { rem -= chol_col_tasks(T, j); j++; }
    const int cnt = chol_col_tasks(T, j);
    const bool is_inv = (rem == cnt - 1);
    const int i = (rem == cnt - 2) ? T : j + 2 + rem;
    const bool is_rhs = (i == T);
    const bool is_prep = !is_inv && !is_rhs && rem == 0;          // tile (j+2, j): also prepares diagonal step j+2
    unsigned long long t0 = 0, t1 = 0, t2 = 0;
    if (d.dbg && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    if (is_inv) {
      // ---- explicit inverse of L_jj (only the back substitution reads it) -----------------------------------
      ll_tile_to_smem<true>(Lll + tile_index(j, j) * (TB * TB), As, tid, epoch);
      __syncthreads();
      inverse32_block<true>(As, Bs, rinv, xs, tid);
      double* gi = Linv + (size_t)j * (TB * TB);
      for (int e = tid; e < TB * TB; e += 256) gi[e] = Bs[(e >> 5) * TLD + (e & 31)];
    } else {
      // ---- initial value, in fragment layout ---------------------------------------------------------------
      double acc[2][2], accp[2][2], accs[2][2];
#pragma unroll
      for (int f = 0; f < 2; f++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int c = fc + 8 * f + e;
          if (is_rhs) acc[f][e] = (fr == 0 && TB * j + c < n) ? d.gc[TB * j + c] - d.rm[TB * j + c] : 0.0;
          else acc[f][e] = (TB * i + fr < n) ? chol_entry(d, n, TB * i + fr, TB * j + c, lambda) : 0.0;
          accp[f][e] = is_prep ? chol_entry(d, n, TB * i + fr, TB * i + c, lambda) : 0.0;
          accs[f][e] = (is_prep && TB * i + fr < n) ? chol_entry(d, n, TB * i + fr, TB * (i - 1) + c, lambda) : 0.0;
        }
      // ---- left-looking updates on the tensor cores ---------------------------------------------------------
      if (is_prep) {
        for (int k = 0; k < j; k++)
          ll_mm_sub3(Lll + tile_index(i, k) * (TB * TB), Lll + tile_index(j, k) * (TB * TB), Lll + tile_index(i - 1, k) * (TB * TB),
                     m0, n0, lane, epoch, acc, accp, accs);
      } else {
        for (int k = 0; k < j; k++)
          ll_mm_sub<1>(Lll + tile_index(i, k) * (TB * TB), Lll + tile_index(j, k) * (TB * TB), nullptr, m0, n0, lane, epoch, acc, accp, true);
      }
      if (d.dbg && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      // ---- X = acc * L_jj^-T ------------------------------------------------------------------------------
#pragma unroll
      for (int f = 0; f < 2; f++) { As[fr * TLD + fc + 8 * f] = acc[f][0]; As[fr * TLD + fc + 8 * f + 1] = acc[f][1]; }
      ll_tile_to_smem<true>(Lll + tile_index(j, j) * (TB * TB), Bs, tid, epoch);
      __syncthreads();
      trsm32_blk(As, Bs, tid);
      __syncthreads();
      uint4* gl = Lll + tile_index(i, j) * (TB * TB);
      double* gp = Lt + tile_index(i, j) * (TB * TB);
#pragma unroll
      for (int q = 0; q < 4; q++) { const int e = tid + 256 * q; const double v = As[(e >> 5) * TLD + (e & 31)]; ll_store(gl + e, v, epoch); gp[e] = v; }
      if (is_prep) {
        // P(i): fold this tile in, then the sub-diagonal tile (i-1, j) the chain CTA is publishing about now
        smem_mm_sub(As, m0, n0, lane, accp);
        smem_ll_mm_sub(As, Lll + tile_index(i - 1, j) * (TB * TB), m0, n0, lane, epoch, accs);
        uint4* pd = Lll + chol_prep_tile(T, i, 0) * (TB * TB);
        uint4* ps = Lll + chol_prep_tile(T, i, 1) * (TB * TB);
#pragma unroll
        for (int f = 0; f < 2; f++)
#pragma unroll
          for (int e = 0; e < 2; e++) { ll_store(ps + fr * TB + fc + 8 * f + e, accs[f][e], epoch); ll_store(pd + fr * TB + fc + 8 * f + e, accp[f][e], epoch); }
      }
    }
    // ---- retire --------------------------------------------------------------------------------------
    if (d.dbg && tid == 0) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t2));
      double* o = d.dbg + 8 * (size_t)task;
      o[0] = is_inv ? -1 : i; o[1] = j; o[2] = (double)(t0 % 1000000000ull); o[3] = (double)(t1 % 1000000000ull); o[4] = (double)(t2 % 1000000000ull); o[5] = blockIdx.x;
    }
    __threadfence();                          // plain copies (L, Linv) of this task visible before it counts as done
    __syncthreads();
    if (tid == 0) s_task = atomicAdd(&ctr[1], 1) - base1;
    __syncthreads();
    if (s_task != n_tasks - 1) continue;
    // ======== last task retired: backward substitution, pose update, scalars (this CTA only) ============
    __threadfence();
    // s_k := y_k (row 0 of the rhs tiles), then for k = T-1 .. 0:  x_k = Linv_k^T s_k ;  s_m -= L_km^T x_k  (m < k)
    for (int e = tid; e < T * TB; e += 256) xs[e] = __ldcg(Lt + tile_index(T, e >> 5) * (TB * TB) + (e & 31));
    __syncthreads();
    for (int k = T - 1; k >= 0; k--) {
      const int c = lane;
      const double* gi = Linv + (size_t)k * (TB * TB);
      double li[4];
#pragma unroll
      for (int q = 0; q < 4; q++) li[q] = __ldcg(gi + (wid * 4 + q) * TB + c);
      // this warp's first tile of block row k (independent of x_k): m = k-1-wid
      const int mfirst = k - 1 - wid;
      double lt[TB];
      if (mfirst >= 0) {
        const double* g = Lt + tile_index(k, mfirst) * (TB * TB);
#pragma unroll
        for (int r = 0; r < TB; r++) lt[r] = __ldcg(g + r * TB + c);
      }
      {
        double p2 = 0;                             // (Linv_k^T s_k)[c], rows 4*wid .. 4*wid+3
#pragma unroll
        for (int q = 0; q < 4; q++) p2 = fma(li[q], xs[k * TB + wid * 4 + q], p2);
        As[wid * TLD + c] = p2;
      }
      __syncthreads();
      if (wid == 0) {
        double sacc = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) sacc += As[w * TLD + c];
        xs[k * TB + c] = sacc;
      }
      __syncthreads();
      for (int m = mfirst; m >= 0; m -= 8) {
        if (m != mfirst) {
          const double* g = Lt + tile_index(k, m) * (TB * TB);
#pragma unroll
          for (int r = 0; r < TB; r++) lt[r] = __ldcg(g + r * TB + c);
        }
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
        for (int r = 0; r < TB; r += 4) {
          s0 = fma(lt[r], xs[k * TB + r], s0); s1 = fma(lt[r + 1], xs[k * TB + r + 1], s1);
          s2 = fma(lt[r + 2], xs[k * TB + r + 2], s2); s3 = fma(lt[r + 3], xs[k * TB + r + 3], s3);
        }
        xs[m * TB + c] -= (s0 + s1) + (s2 + s3);
      }
      __syncthreads();
    }
    if (d.dbg && tid == 0) { unsigned long long t3; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t3)); d.dbg[8 * (size_t)n_tasks] = (double)(t3 % 1000000000ull); }
    const int ok = (ld_acquire(&ctr[2]) != (int)epoch) && ctrl->solve_ok[d.cand];
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
    // fall through to the next (failing) grab so that every worker CTA consumes exactly one id >= n_tasks
  }
}

static int chol_n_tasks(int nc) { const int T = (nc + TB - 1) / TB; int n = 0; for (int j = 0; j < T; j++) n += 2 + (T - j - 2 > 0 ? T - j - 2 : 0); return n; }
size_t chol_tiles_doubles(int nc) { const int T = (nc + TB - 1) / TB; return (size_t)(T * (T + 1) / 2 + T) * TB * TB; }
size_t chol_ll_bytes(int nc) { const int T = (nc + TB - 1) / TB; return (size_t)(T * (T + 1) / 2 + T + 2 * T) * TB * TB * sizeof(uint4); }
size_t chol_inv_doubles(int nc) { const int T = (nc + TB - 1) / TB; return (size_t)T * TB * TB; }
size_t chol_flag_ints(int nc) { (void)nc; return 8; }
int chol_max_n() { return 32 * TB; }
int chol_task_count(int nc) { return chol_n_tasks(nc); }

// The task and done counters are never reset.  A launch consumes n_tasks + workers increments of the task counter (every
// worker CTA ends on one failing grab; block 0 is the chain CTA and takes none) and n_tasks of the done counter; the
// caller keeps the running total of the former in *task_base (the number of workers may differ between launches).
void launch_chol_solve(const BaDev& d, int epoch, int max_ctas, int* task_base, cudaStream_t s)
{
  const int n_tasks = chol_n_tasks(d.nc);
  int workers = n_tasks < max_ctas - 1 ? n_tasks : max_ctas - 1;
  if (workers < 1) workers = 1;
  launch_chain(k_chol_solve, dim3(1 + workers), dim3(256), 0, s, d, (unsigned)epoch, *task_base, (epoch - 1) * n_tasks);
  *task_base += n_tasks + workers;
}

}  // namespace mcp
