// ba_schur.cu — per-point Schur complement as a gather:  Sm = sum_p W_p (V_p + lambda I)^-1 W_p^T.
//
// The scatter formulation (one warp per point, ~1300 fp64 atomics per point into the dense 6N x 6N matrix) is
// bound by L2 atomic throughput.  Here the reduction is turned around: at load time every (point, pose-slot
// pair) incidence is bucketed by the 6x6 block (a,b) of the reduced camera system it contributes to
// ("co-visibility lists", built on the device).  Per LM trial
//   k_schur_y      Y_s = W_s (V_p + lambda I)^-1 for every (point, slot) s      (coalesced, no reduction)
//   k_schur_pairs  one warp per block-pair chunk: lanes stride the incidence list, each lane accumulates a
//                  private 6x6 block  Y_A W_B^T  in registers, one warp reduction, 36 adds per chunk.
// Reads are 144-byte contiguous records served from L2; there is no per-incidence atomic.
#include <algorithm>
#include <cstdlib>

#include "ba_types.cuh"
#include "ba_vinv.cuh"

namespace mcp {

__device__ __forceinline__ int pair_id(int a, int b, int npv) { return a * npv - (a * (a - 1)) / 2 + (b - a); }

// ---- load-time construction of the co-visibility lists ------------------------------------------------
// one thread per (point, slot) entry x: the incidences (x, y >= x) of its point
__global__ void k_pair_count(BaDev d, int* __restrict__ cnt)
{
  for (int s = d.slot_lo + blockIdx.x * blockDim.x + threadIdx.x; s < d.slot_hi; s += gridDim.x * blockDim.x) {
    const int s_end = d.pt_slot_off[d.slot_pt[s] + 1], a = d.slot_var[s];
    for (int y = s; y < s_end; y++) atomicAdd(&cnt[pair_id(a, d.slot_var[y], d.n_pose_var)], 1);
  }
}
__global__ void k_pair_fill(BaDev d, int* __restrict__ cursor, int2* __restrict__ inc)
{
  for (int s = d.slot_lo + blockIdx.x * blockDim.x + threadIdx.x; s < d.slot_hi; s += gridDim.x * blockDim.x) {
    const int s_end = d.pt_slot_off[d.slot_pt[s] + 1], a = d.slot_var[s];
    for (int y = s; y < s_end; y++) {
      const int pos = atomicAdd(&cursor[pair_id(a, d.slot_var[y], d.n_pose_var)], 1);
      inc[pos] = make_int2(s, y);
    }
  }
}

// one block: exclusive scan of the per-pair incidence counts (in place: cnt becomes the fill cursor of k_pair_fill,
// cnt[n_pairs] the total), the work items {block row, block col, begin, end} of <= 128 incidences in (row, col, begin)
// order, and their number -- on the device, so that mcp_ba_load never waits for a read-back
__global__ void __launch_bounds__(1024) k_pair_items(BaDev d, int* __restrict__ cnt, int4* __restrict__ items, int* __restrict__ n_items_out)
{
  __shared__ int sh_inc[1024], sh_it[1024];
  const int npv = d.n_pose_var, n_pairs = npv * (npv + 1) / 2;
  const int t = threadIdx.x, per = (n_pairs + 1023) / 1024;
  const int lo = min(t * per, n_pairs), hi = min(lo + per, n_pairs);
  int s_inc = 0, s_it = 0;
  for (int i = lo; i < hi; i++) { const int c = cnt[i]; s_inc += c; s_it += (c + 127) >> 7; }
  sh_inc[t] = s_inc; sh_it[t] = s_it;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int a = t >= o ? sh_inc[t - o] : 0, b = t >= o ? sh_it[t - o] : 0;
    __syncthreads();
    sh_inc[t] += a; sh_it[t] += b;
    __syncthreads();
  }
  int off_inc = sh_inc[t] - s_inc, off_it = sh_it[t] - s_it;
  int a = 0, rem = lo;
  while (a < npv && rem >= npv - a) { rem -= npv - a; a++; }
  int b = a + rem;
  for (int i = lo; i < hi; i++) {
    const int c = cnt[i];
    cnt[i] = off_inc;
    for (int q = 0; q < c; q += 128) items[off_it++] = make_int4(a, b, off_inc + q, off_inc + min(q + 128, c));
    off_inc += c;
    if (++b == npv) { a++; b = a; }
  }
  if (t == 1023) { cnt[n_pairs] = sh_inc[1023]; *n_items_out = sh_it[1023]; }
}

// ---- per trial --------------------------------------------------------------------------------------
// one thread per (point, slot): Y = W Vinv (18 doubles), z = Y g_p (6 doubles)
__global__ void __launch_bounds__(256) k_schur_y(BaDev d)
{
  const double lambda = trial_lambda(d);
  const int s_lo = d.slot_lo, s_hi = d.slot_hi;
  int fail = 0;
  for (int s = s_lo + blockIdx.x * blockDim.x + threadIdx.x; s < s_hi; s += gridDim.x * blockDim.x) {
    const int p = d.slot_pt[s];
    double V6[6], gp[3], Vi[9];
#pragma unroll
    for (int i = 0; i < 6; i++) V6[i] = d.V[6 * (size_t)p + i];
#pragma unroll
    for (int i = 0; i < 3; i++) gp[i] = d.gp[3 * (size_t)p + i];
    if (!inv3_sym_s(V6, lambda, Vi)) fail = 1;
    const double2* w2 = reinterpret_cast<const double2*>(d.W + 18 * (size_t)s);
    double w[18];
#pragma unroll
    for (int i = 0; i < 9; i++) { const double2 t = w2[i]; w[2 * i] = t.x; w[2 * i + 1] = t.y; }
    double y[24];
#pragma unroll
    for (int r = 0; r < 6; r++) {
#pragma unroll
      for (int c = 0; c < 3; c++) y[r * 3 + c] = w[r * 3] * Vi[c] + w[r * 3 + 1] * Vi[3 + c] + w[r * 3 + 2] * Vi[6 + c];
      y[18 + r] = y[r * 3] * gp[0] + y[r * 3 + 1] * gp[1] + y[r * 3 + 2] * gp[2];
    }
    double2* y2 = reinterpret_cast<double2*>(d.Y + 24 * (size_t)s);
#pragma unroll
    for (int i = 0; i < 12; i++) y2[i] = make_double2(y[2 * i], y[2 * i + 1]);
  }
  if (fail) atomicExch(&d.ctrl->solve_ok[d.cand < 0 ? 0 : d.cand], 0);
}

// One warp per work item {block row a, block col b, begin, end}.  Incidences are processed in groups of G:
// the warp gathers the Y_A (192 B: Y + z) and W_B (144 B) records of the group with coalesced 16-byte pieces
// into shared memory, then every lane accumulates the output entries it owns:
//   lane l < 32 : S[r][c], (r,c) = (l / 6, l % 6);  lanes 0..3 additionally own entries 32..35;
//   lanes 4..9  : rm[a][lane-4] on diagonal items (sum of z).
constexpr int SG = 32;                       // incidences per group
constexpr int SREC = 42;                     // doubles per staged incidence: Y(18) z(6) W(18)
__global__ void __launch_bounds__(128) k_schur_pairs(BaDev d)
{
  __shared__ __align__(16) double stage[4][SG * SREC];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int nc = d.nc;
  double* st = stage[wid];
  const int r0 = lane / 6, c0 = lane - 6 * r0;           // entry `lane`
  const int e1 = 32 + (lane & 3), r1 = e1 / 6, c1 = e1 - 6 * r1;
  const int n_items = __ldg(d.n_items_dev);
  for (int it = gw; it < n_items; it += nw) {
    const int4 item = d.items[it];
    const bool diag = item.x == item.y;
    double acc0 = 0.0, acc1 = 0.0, accz = 0.0;
    for (int g0 = item.z; g0 < item.w; g0 += SG) {
      const int ng = min(SG, item.w - g0);
      int2 ab = make_int2(0, 0);
      if (lane < ng) ab = d.inc[g0 + lane];
      // 21 pieces of 16 B per incidence: 12 of the Y record, 9 of the W record
      const int npiece = ng * 21;
      // all 21 gathers of the group are issued back to back (static unroll), then staged to shared memory
      double2 piece[21];
#pragma unroll
      for (int k = 0; k < 21; k++) {
        const int e = k * 32 + lane;
        const int q = min(e / 21, ng - 1), part = e - (e / 21) * 21;
        const int sa = __shfl_sync(0xffffffffu, ab.x, q), sb = __shfl_sync(0xffffffffu, ab.y, q);
        const double2* src = part < 12 ? reinterpret_cast<const double2*>(d.Y + 24 * (size_t)sa) + part
                                       : reinterpret_cast<const double2*>(d.W + 18 * (size_t)sb) + (part - 12);
        piece[k] = (e < npiece) ? __ldcg(src) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int k = 0; k < 21; k++) {
        const int e = k * 32 + lane;
        if (e < npiece) reinterpret_cast<double2*>(st + (e / 21) * SREC)[e - (e / 21) * 21] = piece[k];
      }
      __syncwarp();
      for (int q = 0; q < ng; q++) {
        const double* y = st + q * SREC;
        const double* w = y + 24;
        acc0 += y[r0 * 3] * w[c0 * 3] + y[r0 * 3 + 1] * w[c0 * 3 + 1] + y[r0 * 3 + 2] * w[c0 * 3 + 2];
        if (lane < 4) acc1 += y[r1 * 3] * w[c1 * 3] + y[r1 * 3 + 1] * w[c1 * 3 + 1] + y[r1 * 3 + 2] * w[c1 * 3 + 2];
        else if (diag && lane < 10) accz += y[18 + lane - 4];
      }
      __syncwarp();
    }
    double* base = d.Sm + (size_t)(6 * item.x) * nc + 6 * item.y;
    atomicAdd(base + (size_t)r0 * nc + c0, acc0);
    if (lane < 4) atomicAdd(base + (size_t)r1 * nc + c1, acc1);
    else if (diag && lane < 10) atomicAdd(d.rm + 6 * item.x + lane - 4, accz);
  }
}

// ---------------------------------------------------------------------------------------------
// k_schur_pairs_tma: same work items, B200-native data path.
//   * every lane issues two cp.async.bulk (TMA 1-D bulk) copies for "its" incidence of the group — the 192-byte
//     Y record and the 144-byte W record — straight from L2 into the warp's staging buffer; completion is counted
//     in bytes on an mbarrier (no per-piece address arithmetic, no register staging), double buffered;
//   * the 6x6 block  sum_q Y_q W_q^T  is a true contraction over k = (incidence, 3):  D(8x8) += A(8x4) B(4x8) with
//     fp64 tensor cores (mma.sync m8n8k4, rows/cols 6,7 are zero padding), 24 DMMAs per 32 incidences.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int TG = 32;                       // incidences per group (K = 96)
constexpr int TW = 4;                        // warps per block
__global__ void __launch_bounds__(TW * 32) k_schur_pairs_tma(BaDev d)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double* stage = reinterpret_cast<double*>(smem_raw) + (size_t)wid * 2 * TG * SREC;          // [2][TG][SREC]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)TW * 2 * TG * SREC * sizeof(double)) + wid * 2;
  if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int nc = d.nc;
  const int g = lane >> 2, kk = lane & 3;              // fragment coordinates
  // per-phase offsets inside a 4-incidence (12 k) block: k = 4p + kk -> (incidence, component)
  int offA[3], offB[3];
#pragma unroll
  for (int p = 0; p < 3; p++) {
    const int k = 4 * p + kk, q = k / 3, t = k - 3 * q;
    offA[p] = q * SREC + g * 3 + t;                    // Y_q[g][t]
    offB[p] = q * SREC + 24 + g * 3 + t;               // W_q[g][t]
  }
  const bool gvalid = g < 6;
  unsigned phase[2] = { 0u, 0u };
  const int n_items = __ldg(d.n_items_dev);
  for (int it = gw; it < n_items; it += nw) {
    const int4 item = d.items[it];
    const bool diag = item.x == item.y;
    const int n_groups = (item.w - item.z + TG - 1) / TG;
    double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0, accz = 0.0;      // two accumulator pairs: halves the dependent DMMA chain
    // prologue: issue group 0 into buffer 0
    auto issue = [&](int grp, int buf) {
      const int g0 = item.z + grp * TG, ng = min(TG, item.w - g0);
      double* st = stage + (size_t)buf * TG * SREC;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (lane == 0) mbar_expect_tx(&bars[buf], (unsigned)ng * (192u + 144u));
      __syncwarp();
      if (lane < ng) {
        const int2 ab = d.inc[g0 + lane];
        bulk_g2s(st + lane * SREC, d.Y + 24 * (size_t)ab.x, 192u, &bars[buf]);
        bulk_g2s(st + lane * SREC + 24, d.W + 18 * (size_t)ab.y, 144u, &bars[buf]);
      } else {
        for (int i = 0; i < SREC; i++) st[lane * SREC + i] = 0.0;          // zero padding of a partial group
      }
    };
    issue(0, 0);
    for (int grp = 0; grp < n_groups; grp++) {
      const int buf = grp & 1;
      if (grp + 1 < n_groups) issue(grp + 1, buf ^ 1);
      mbar_wait(&bars[buf], phase[buf]);
      phase[buf] ^= 1u;
      __syncwarp();
      const double* st = stage + (size_t)buf * TG * SREC;
#pragma unroll
      for (int m = 0; m < TG / 4; m++) {
        const double* blk = st + m * 4 * SREC;
#pragma unroll
        for (int p = 0; p < 3; p++) {
          const double a = gvalid ? blk[offA[p]] : 0.0;
          const double b = gvalid ? blk[offB[p]] : 0.0;
          if (m & 1) dmma_m8n8k4(e0, e1, a, b); else dmma_m8n8k4(c0, c1, a, b);
        }
      }
      if (diag && lane < 6) {
#pragma unroll 8
        for (int q = 0; q < TG; q++) accz += st[q * SREC + 18 + lane];
      }
      __syncwarp();
    }
    c0 += e0; c1 += e1;
    // D fragment: row g, cols 2*kk, 2*kk+1
    if (g < 6) {
      double* base = d.Sm + (size_t)(6 * item.x + g) * nc + 6 * item.y;
      if (2 * kk < 6) atomicAdd(base + 2 * kk, c0);
      if (2 * kk + 1 < 6) atomicAdd(base + 2 * kk + 1, c1);
    }
    if (diag && lane < 6) atomicAdd(d.rm + 6 * item.x + lane, accz);
  }
}

// ---------------------------------------------------------------------------------------------
// Multi-candidate Schur reduction.  The speculative LM candidates of one trial round differ only in lambda, i.e. in
// the 3x3 inverse (V_p + lambda_c I)^-1 of every point; W, the co-visibility lists and the work items are shared.
// Three separate pair kernels (one per candidate stream) each saturate the copy engine and serialise (timeline:
// 3 x ~52 us on the critical path of the round).  Here ONE pass over the incidence lists serves all candidates:
//   k_schur_vinv_multi   per point: R_p = [Vinv_0 | Vinv_1 | Vinv_2 | u_0 | u_1 | u_2], u_c = Vinv_c g_p   (288 B)
//   k_schur_pairs_multi  per incidence three bulk copies -- W_A, W_B (144 B each) and R_p -- instead of two per
//                        candidate; Y_A^(c) = W_A Vinv_c is formed in registers on the way into the fp64 tensor-core
//                        MMA (same expression as k_schur_y: identical rounding), z_A^(c) = W_A u_c on diagonal items.
// ---------------------------------------------------------------------------------------------
constexpr int MG_CA = 16;                    // incidences per group of the cp.async variant
constexpr int MRECD = 72;                    // doubles per staged incidence: W_A(18) W_B(18) R(36)

template <int NC>
__global__ void __launch_bounds__(256) k_schur_vinv_multi(BaDev d, SchurMulti mc)
{
  pdl_prologue();
  schur_vinv_body<NC>(d, mc, blockIdx.x, gridDim.x);
}

template <int NC, int MG>
__global__ void __launch_bounds__(TW * 32) k_schur_pairs_multi(BaDev d, SchurMulti mc)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double* stage = reinterpret_cast<double*>(smem_raw) + (size_t)wid * 2 * MG * MRECD;          // [2][MG][MRECD]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)TW * 2 * MG * MRECD * sizeof(double)) + wid * 2;
  if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int nc = d.nc;
  const int g = lane >> 2, kk = lane & 3;              // fragment coordinates
  // per-phase offsets inside a 4-incidence (12 k) block: k = 4p + kk -> (incidence q, component t)
  int offW[3], offV[3], offB[3];
#pragma unroll
  for (int p = 0; p < 3; p++) {
    const int k = 4 * p + kk, q = k / 3, t = k - 3 * q;
    offW[p] = q * MRECD + g * 3;                       // W_A,q[g][0..2]
    offV[p] = q * MRECD + 36 + 3 * t;                  // Vinv_c,q[t][0..2] (+ 9 c; symmetric: row t = column t)
    offB[p] = q * MRECD + 18 + g * 3 + t;              // W_B,q[g][t]
  }
  const bool gvalid = g < 6;
  unsigned phase[2] = { 0u, 0u };
  const int n_items = __ldg(d.n_items_dev);
  for (int it = gw; it < n_items; it += nw) {
    const int4 item = d.items[it];
    const bool diag = item.x == item.y;
    const int n_groups = (item.w - item.z + MG - 1) / MG;
    double acc[NC][4], accz[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) { acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.0; accz[c] = 0.0; }
    auto issue = [&](int grp, int buf) {
      const int g0 = item.z + grp * MG, ng = min(MG, item.w - g0);
      double* st = stage + (size_t)buf * MG * MRECD;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (lane == 0) mbar_expect_tx(&bars[buf], (unsigned)ng * (unsigned)(MRECD * sizeof(double)));
      __syncwarp();
      if (lane < ng) {
        const int2 ab = d.inc[g0 + lane];
        const int pt = __ldg(d.slot_pt + ab.x);
        bulk_g2s(st + lane * MRECD, d.W + 18 * (size_t)ab.x, 144u, &bars[buf]);
        bulk_g2s(st + lane * MRECD + 18, d.W + 18 * (size_t)ab.y, 144u, &bars[buf]);
        bulk_g2s(st + lane * MRECD + 36, mc.R + RPD * (size_t)pt, (unsigned)(RPD * sizeof(double)), &bars[buf]);
      } else if (lane < MG) {
        for (int i = 0; i < MRECD; i++) st[lane * MRECD + i] = 0.0;        // zero padding of a partial group
      }
    };
    issue(0, 0);
    for (int grp = 0; grp < n_groups; grp++) {
      const int buf = grp & 1;
      if (grp + 1 < n_groups) issue(grp + 1, buf ^ 1);
      mbar_wait(&bars[buf], phase[buf]);
      phase[buf] ^= 1u;
      __syncwarp();
      const double* st = stage + (size_t)buf * MG * MRECD;
#pragma unroll
      for (int m = 0; m < MG / 4; m++) {
        const double* blk = st + m * 4 * MRECD;
#pragma unroll
        for (int p = 0; p < 3; p++) {
          const double w0 = blk[offW[p]], w1 = blk[offW[p] + 1], w2 = blk[offW[p] + 2];
          const double b = gvalid ? blk[offB[p]] : 0.0;
#pragma unroll
          for (int c = 0; c < NC; c++) {
            const double* v = blk + offV[p] + 9 * c;
            double a = w0 * v[0] + w1 * v[1] + w2 * v[2];
            a = gvalid ? a : 0.0;
            dmma_m8n8k4(acc[c][2 * (m & 1)], acc[c][2 * (m & 1) + 1], a, b);
          }
        }
      }
      if (diag && lane < 6) {
#pragma unroll 4
        for (int q = 0; q < MG; q++) {
          const double* rec = st + q * MRECD;
          const double w0 = rec[lane * 3], w1 = rec[lane * 3 + 1], w2 = rec[lane * 3 + 2];
#pragma unroll
          for (int c = 0; c < NC; c++) accz[c] += w0 * rec[63 + 3 * c] + w1 * rec[64 + 3 * c] + w2 * rec[65 + 3 * c];
        }
      }
      __syncwarp();
    }
    // D fragment: row g, cols 2*kk, 2*kk+1
#pragma unroll
    for (int c = 0; c < NC; c++) {
      if (g < 6) {
        double* base = mc.Sm[c] + (size_t)(6 * item.x + g) * nc + 6 * item.y;
        if (2 * kk < 6) atomicAdd(base + 2 * kk, acc[c][0] + acc[c][2]);
        if (2 * kk + 1 < 6) atomicAdd(base + 2 * kk + 1, acc[c][1] + acc[c][3]);
      }
      if (diag && lane < 6) atomicAdd(mc.rm[c] + 6 * item.x + lane, accz[c]);
    }
  }
}

// The same reduction with the staging done by the warp itself: 16-byte cp.async (LDGSTS) pieces, 32 per
// instruction, instead of three bulk copies per incidence.  The bulk-copy version runs at the copy engine's
// per-operation rate (~1 small copy per 20 cycles per SM: 1.16 M copies -> ~90 us at 200 KF / 10 k points) although
// the bytes are a fraction of what the L2 can deliver; 36 pieces per incidence through the LSU cost 18 warp
// instructions per group of 16 incidences.
template <int NC>
__global__ void __launch_bounds__(TW * 32, 3) k_schur_pairs_multi_ca(BaDev d, SchurMulti mc)
{
  pdl_prologue();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double* stage = reinterpret_cast<double*>(smem_raw) + (size_t)wid * 2 * MG_CA * MRECD;          // [2][MG_CA][MRECD]
  // per staged incidence the three source pointers (W_A, W_B, R_p), section-major: [2][3][MG_CA]
  unsigned long long* ptrs = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)TW * 2 * MG_CA * MRECD * sizeof(double)) + wid * 2 * 3 * MG_CA;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int nc = d.nc;
  const int g = lane >> 2, kk = lane & 3;              // fragment coordinates
  int offW[3], offV[3], offB[3];
#pragma unroll
  for (int p = 0; p < 3; p++) {
    const int k = 4 * p + kk, q = k / 3, t = k - 3 * q;
    offW[p] = q * MRECD + g * 3;
    offV[p] = q * MRECD + 36 + 3 * t;
    offB[p] = q * MRECD + 18 + g * 3 + t;
  }
  const bool gvalid = g < 6;
  const unsigned long long Wb = reinterpret_cast<unsigned long long>(d.W), Rb = reinterpret_cast<unsigned long long>(mc.R);
  const int n_items = __ldg(d.n_items_dev);
  // the first item of a warp is its index, further ones come from a device counter (cleared by k_schur_vinv_multi):
  // items range from 1 to 128 incidences, a static round-robin leaves a tail
  (void)nw;
  for (int it = gw; it < n_items;) {
    const int4 item = d.items[it];
    {
      int nx = 0;
      if (lane == 0) nx = atomicAdd(mc.next_item, 1);
      it = __shfl_sync(0xffffffffu, nx, 0);
    }
    const bool diag = item.x == item.y;
    const int n_groups = (item.w - item.z + MG_CA - 1) / MG_CA;
    double acc[NC][4], accz[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) { acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.0; accz[c] = 0.0; }
    // lane q < MG_CA: indices {slot A, slot B, point} of incidence q of a group, -1 beyond the end of the item.  They
    // are fetched one group ahead of the copies that need them (two dependent L2 round trips, hidden by the compute).
    auto load_idx = [&](int grp) -> int3 {
      int3 o = make_int3(-1, 0, 0);
      const int e = item.z + grp * MG_CA + lane;
      if (lane < MG_CA && grp < n_groups && e < item.w) {
        const int2 ab = d.inc[e];
        o = make_int3(ab.x, ab.y, __ldg(d.slot_pt + ab.x));
      }
      return o;
    };
    auto issue = [&](const int3 o, int buf) {
      unsigned long long* pt = ptrs + buf * 3 * MG_CA;
      if (lane < MG_CA) {
        const bool on = o.x >= 0;
        pt[lane] = on ? Wb + 144ull * (unsigned)o.x : 0ull;
        pt[MG_CA + lane] = on ? Wb + 144ull * (unsigned)o.y : 0ull;
        pt[2 * MG_CA + lane] = on ? Rb + 288ull * (unsigned)o.z : 0ull;
      }
      __syncwarp();
      const unsigned dst0 = smem_u32(stage + (size_t)buf * MG_CA * MRECD);
      int q = lane / 36, pc = lane - 36 * q;             // piece j = 32 i + lane of the group: incidence q, 16-byte piece pc
#pragma unroll 6
      for (int i = 0; i < MG_CA * 36 / 32; i++) {
        const int sec = (pc >= 9) + (pc >= 18);          // 0: W_A, 1: W_B, 2: R_p
        const unsigned long long base = pt[sec * MG_CA + q];
        const unsigned long long src = base + (unsigned)(16 * (pc - 9 * sec));
        const unsigned nbytes = base ? 16u : 0u;         // 0: zero fill (padding of a partial group)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + (unsigned)(q * MRECD * 8 + pc * 16)), "l"(base ? src : Wb), "r"(nbytes) : "memory");
        pc += 32;
        if (pc >= 36) { pc -= 36; q++; }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int3 o_next = load_idx(0);
    issue(o_next, 0);
    o_next = load_idx(1);
    for (int grp = 0; grp < n_groups; grp++) {
      const int buf = grp & 1;
      if (grp + 1 < n_groups) {
        issue(o_next, buf ^ 1);
        o_next = load_idx(grp + 2);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncwarp();
      const double* st = stage + (size_t)buf * MG_CA * MRECD;
#pragma unroll
      for (int m = 0; m < MG_CA / 4; m++) {
        const double* blk = st + m * 4 * MRECD;
#pragma unroll
        for (int p = 0; p < 3; p++) {
          const double w0 = blk[offW[p]], w1 = blk[offW[p] + 1], w2 = blk[offW[p] + 2];
          const double b = gvalid ? blk[offB[p]] : 0.0;
#pragma unroll
          for (int c = 0; c < NC; c++) {
            const double* v = blk + offV[p] + 9 * c;
            double a = w0 * v[0] + w1 * v[1] + w2 * v[2];
            a = gvalid ? a : 0.0;
            dmma_m8n8k4(acc[c][2 * (m & 1)], acc[c][2 * (m & 1) + 1], a, b);
          }
        }
      }
      if (diag && lane < 6) {
#pragma unroll 4
        for (int q = 0; q < MG_CA; q++) {
          const double* rec = st + q * MRECD;
          const double w0 = rec[lane * 3], w1 = rec[lane * 3 + 1], w2 = rec[lane * 3 + 2];
#pragma unroll
          for (int c = 0; c < NC; c++) accz[c] += w0 * rec[63 + 3 * c] + w1 * rec[64 + 3 * c] + w2 * rec[65 + 3 * c];
        }
      }
      __syncwarp();
    }
#pragma unroll
    for (int c = 0; c < NC; c++) {
      if (g < 6) {
        double* base = mc.Sm[c] + (size_t)(6 * item.x + g) * nc + 6 * item.y;
        if (2 * kk < 6) atomicAdd(base + 2 * kk, acc[c][0] + acc[c][2]);
        if (2 * kk + 1 < 6) atomicAdd(base + 2 * kk + 1, acc[c][1] + acc[c][3]);
      }
      if (diag && lane < 6) atomicAdd(mc.rm[c] + 6 * item.x + lane, accz[c]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k_schur_rows: row-wise formulation.  The pair kernels above fetch two small records per (point, pose pair)
// incidence -- 870 k bulk copies of 144-192 bytes per launch at 200 KF / 10 k points, and the kernel runs at the
// TMA's issue rate, not at L2 bandwidth.  Here a work item is ONE pose variable a and a run of the (point, slot)
// entries that observe it.  Per entry the warp fetches Y_a (192 B) and the W records of the point's slots from a to
// the end of the point -- contiguous, because the slots of a point are stored in ascending pose order -- with two
// bulk copies, and adds the blocks Y_a W_b^T for every b >= a of that point with one fp64 tensor-core MMA each
// (m8n8k4: rows/cols 6,7 and k = 3 are padding).  The blocks of row a accumulate in a per-warp shared-memory strip
// (every lane owns fixed fragment positions: no synchronisation), flushed with one atomic per touched entry.
// One tenth of the copies, the same bytes.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_schur_rows(BaDev d)
{
  const int nblk = d.rs_nblk;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const size_t acc_bytes = (size_t)(nblk + 8) * 48 * sizeof(double);
  const size_t per_warp = acc_bytes + 2 * RS_BYTES + 2 * 64 * sizeof(int) + 2 * sizeof(unsigned long long);
  unsigned char* base = smem_raw + per_warp * wid;
  double* acc = reinterpret_cast<double*>(base);
  unsigned char* stg = base + acc_bytes;
  int* sv = reinterpret_cast<int*>(stg + 2 * RS_BYTES);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sv + 128);
  if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  for (int i = lane; i < (nblk + 8) * 48; i += 32) acc[i] = 0.0;
  __syncwarp();
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int nc = d.nc;
  const int g = lane >> 2, kk = lane & 3;
  const bool fvalid = g < 6 && kk < 3;
  unsigned phase[2] = { 0u, 0u };

  for (int it = gw; it < d.n_rs_items; it += nw) {
    const int4 item = d.rs_items[it];
    const int a = item.x, n_grp = item.z - item.y;
    double accz = 0.0;
    // lane k < count of a group holds entry k: {slot, number of slots to the end of the point}
    auto load_ent = [&](int grp, int& cnt) -> int2 {
      int2 e = make_int2(0, 0);
      cnt = 0;
      if (grp < n_grp) {
        const int gd = __ldg(d.rs_grp + item.y + grp);
        cnt = gd & 15;
        if (lane < cnt) e = __ldg(d.rs_ent + (gd >> 4) + lane);
      }
      return e;
    };
    // byte offset of entry `lane` inside the staging buffer / index of its first block in the sv list
    auto scan = [&](int2 e, int cnt, int& off, int& pre) {
      const int sz = lane < cnt ? 192 + 144 * e.y : 0, nb = lane < cnt ? e.y : 0;
      int io = sz, ip = nb;
#pragma unroll
      for (int o = 1; o < RS_MAXE; o <<= 1) {
        const int to = __shfl_up_sync(0xffffffffu, io, o), tp = __shfl_up_sync(0xffffffffu, ip, o);
        if (lane >= o) { io += to; ip += tp; }
      }
      off = io - sz; pre = ip - nb;
    };
    // issue the copies of one group; returns the block indices (pose variable - a) this lane stages later
    auto issue = [&](int2 e, int cnt, int buf, int& v0, int& v1) {
      int off, pre;
      scan(e, cnt, off, pre);
      const int tot_bytes = __shfl_sync(0xffffffffu, off + (lane < cnt ? 192 + 144 * e.y : 0), cnt - 1);
      const int tot_nb = __shfl_sync(0xffffffffu, pre + (lane < cnt ? e.y : 0), cnt - 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (lane == 0) mbar_expect_tx(&bars[buf], (unsigned)tot_bytes);
      __syncwarp();
      if (lane < cnt) {
        unsigned char* dst = stg + (size_t)buf * RS_BYTES + off;
        bulk_g2s(dst, d.Y + 24 * (size_t)e.x, 192u, &bars[buf]);
        bulk_g2s(dst + 192, d.W + 18 * (size_t)e.x, 144u * (unsigned)e.y, &bars[buf]);
      }
      // block index of (entry k, j): slot_var[slot_k + j] - a, flattened in entry order
      v0 = 0; v1 = 0;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int o = lane + 32 * h;
        int sl = -1;
#pragma unroll
        for (int k = 0; k < RS_MAXE; k++) {
          const int pk = __shfl_sync(0xffffffffu, pre, k), nk = __shfl_sync(0xffffffffu, e.y, k), sk = __shfl_sync(0xffffffffu, e.x, k);
          if (k < cnt && o >= pk && o < pk + nk) sl = sk + (o - pk);
        }
        const int v = (sl >= 0 && o < tot_nb) ? __ldg(d.slot_var + sl) - a : 0;
        if (h == 0) v0 = v; else v1 = v;
      }
    };

    int cnt_cur, cnt_nxt;
    int2 e_cur = load_ent(0, cnt_cur);
    int2 e_nxt = load_ent(1, cnt_nxt);
    int v0c, v1c, v0n = 0, v1n = 0;
    issue(e_cur, cnt_cur, 0, v0c, v1c);
    for (int grp = 0; grp < n_grp; grp++) {
      const int buf = grp & 1;
      if (grp + 1 < n_grp) issue(e_nxt, cnt_nxt, buf ^ 1, v0n, v1n);
      int cnt_n2;
      const int2 e_n2 = load_ent(grp + 2, cnt_n2);
      int off, pre;
      scan(e_cur, cnt_cur, off, pre);
      sv[buf * 64 + lane] = v0c; sv[buf * 64 + 32 + lane] = v1c;
      mbar_wait(&bars[buf], phase[buf]);
      phase[buf] ^= 1u;
      __syncwarp();
      const unsigned char* sb = stg + (size_t)buf * RS_BYTES;
      for (int k = 0; k < cnt_cur; k++) {
        const int ok = __shfl_sync(0xffffffffu, off, k), pk = __shfl_sync(0xffffffffu, pre, k), nk = __shfl_sync(0xffffffffu, e_cur.y, k);
        const double* Yb = reinterpret_cast<const double*>(sb + ok);
        const double av = fvalid ? Yb[g * 3 + kk] : 0.0;
        if (lane < 6) accz += Yb[18 + lane];
        // the blocks of one entry are distinct (a point lists a pose once): eight read-modify-write updates of the
        // accumulator strip are batched so that their shared-memory round trips overlap (unused lanes of a batch
        // go to the spare blocks behind the strip)
        for (int j0 = 0; j0 < nk; j0 += 8) {
          double bv[8];
          double2 c[8];
          double2* cp[8];
#pragma unroll
          for (int u = 0; u < 8; u++) {
            const bool on = j0 + u < nk;
            const double* Wj = Yb + 24 + 18 * (j0 + u);
            bv[u] = (fvalid && on) ? Wj[g * 3 + kk] : 0.0;
            const int blk = on ? sv[buf * 64 + pk + j0 + u] : nblk + u;
            cp[u] = reinterpret_cast<double2*>(acc + (size_t)blk * 48 + g * 8 + 2 * kk);
          }
#pragma unroll
          for (int u = 0; u < 8; u++) c[u] = fvalid ? *cp[u] : make_double2(0.0, 0.0);
#pragma unroll
          for (int u = 0; u < 8; u++) dmma_m8n8k4(c[u].x, c[u].y, av, bv[u]);
#pragma unroll
          for (int u = 0; u < 8; u++) if (fvalid) *cp[u] = c[u];
        }
      }
      __syncwarp();
      e_cur = e_nxt; cnt_cur = cnt_nxt; v0c = v0n; v1c = v1n;
      e_nxt = e_n2; cnt_nxt = cnt_n2;
    }
    // flush row a: blocks (a, a + blk)
    const int nb_row = min(nblk, d.n_pose_var - a);
    for (int e = lane; e < nb_row * 36; e += 32) {
      const int blk = e / 36, rc = e - 36 * blk, r = rc / 6, c = rc - 6 * r;
      double* ap = acc + (size_t)blk * 48 + r * 8 + c;
      const double v = *ap;
      if (v != 0.0) { atomicAdd(d.Sm + (size_t)(6 * a + r) * nc + 6 * (a + blk) + c, v); *ap = 0.0; }
    }
    if (lane < 6) atomicAdd(d.rm + 6 * a + lane, accz);
    __syncwarp();
  }
}

// ChainBundle's point-depth covariance (src/ChainBundle.cc:1401-1448; [3P] SparseOptimizer::computeMarginals on the
// undamped Hessian of the last buildSystem): (H^-1)_pp = V^-1 + Y^T S^-1 Y with Y = W V^-1 (lambda = 0) and
// S = H_cc - sum W V^-1 W^T.  Only the (2,2) entry (radial direction) of every non-fixed point is needed.
// The reference only attempts this with < 3 movable poses, so S is at most 12 x 12: one block, S^-1 by thread 0.
constexpr int MARG_MAXN = 32;
__global__ void __launch_bounds__(256) k_marginals(BaDev d, double* __restrict__ cov)
{
  __shared__ double S[MARG_MAXN * MARG_MAXN], Si[MARG_MAXN * MARG_MAXN];
  __shared__ int s_fail;
  const int n = d.nc;
  if (threadIdx.x == 0) s_fail = (n > MARG_MAXN) ? 1 : 0;
  for (int e = threadIdx.x; e < n * n && n <= MARG_MAXN; e += blockDim.x) {
    const int r = e / n, c = e - r * n;
    const int lo = r < c ? r : c, hi = r < c ? c : r;                 // the upper triangle is the valid one
    S[e] = d.H0[(size_t)lo * n + hi] - d.Sm[(size_t)lo * n + hi];
  }
  __syncthreads();
  if (threadIdx.x == 0 && !s_fail && n > 0) {
    // Cholesky S = L L^T in place (lower), then S^-1 = L^-T L^-1 column by column
    for (int j = 0; j < n && !s_fail; j++) {
      double dj = S[j * n + j];
      for (int k = 0; k < j; k++) dj -= S[j * n + k] * S[j * n + k];
      if (!(dj > 0.0) || !isfinite(dj)) { s_fail = 1; break; }
      const double l = sqrt(dj);
      S[j * n + j] = l;
      for (int i = j + 1; i < n; i++) {
        double v = S[i * n + j];
        for (int k = 0; k < j; k++) v -= S[i * n + k] * S[j * n + k];
        S[i * n + j] = v / l;
      }
    }
    for (int c = 0; c < n && !s_fail; c++) {
      double y[MARG_MAXN];
      for (int i = 0; i < n; i++) {
        double v = (i == c) ? 1.0 : 0.0;
        for (int k = 0; k < i; k++) v -= S[i * n + k] * y[k];
        y[i] = v / S[i * n + i];
      }
      for (int i = n - 1; i >= 0; i--) {
        double v = y[i];
        for (int k = i + 1; k < n; k++) v -= S[k * n + i] * y[k];
        y[i] = v / S[i * n + i];
      }
      for (int i = 0; i < n; i++) Si[i * n + c] = y[i];
    }
  }
  __syncthreads();
  int fail = 0;
  if (!s_fail)
    for (int p = threadIdx.x; p < d.n_pt; p += blockDim.x) {
      const int pv = d.pt_var[p];
      if (pv < 0) continue;
      double V6[6], Vi[9];
#pragma unroll
      for (int i = 0; i < 6; i++) V6[i] = d.V[6 * (size_t)p + i];
      if (!inv3_sym_s(V6, 0.0, Vi)) fail = 1;
      double c22 = Vi[8];
      const int s0 = d.pt_slot_off[p], K = d.pt_slot_off[p + 1] - s0;
      for (int a = 0; a < K; a++) {
        const int va = d.slot_var[s0 + a];
        double ya[6];
#pragma unroll
        for (int r = 0; r < 6; r++) ya[r] = d.Y[24 * (size_t)(s0 + a) + 3 * r + 2];
        for (int b = 0; b < K; b++) {
          const int vb = d.slot_var[s0 + b];
#pragma unroll
          for (int r = 0; r < 6; r++) {
            double t = 0.0;
#pragma unroll
            for (int c = 0; c < 6; c++) t += Si[(6 * va + r) * n + 6 * vb + c] * d.Y[24 * (size_t)(s0 + b) + 3 * c + 2];
            c22 += ya[r] * t;
          }
        }
      }
      cov[pv] = c22;
    }
  if (fail) atomicExch(&s_fail, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    d.ctrl->marg_fail = (s_fail || !d.ctrl->solve_ok[0]) ? 1 : 0;
    d.ctrl->solve_ok[0] = 1;
  }
}
void launch_marginals(const BaDev& d, double* cov, cudaStream_t s) { k_marginals<<<1, 256, 0, s>>>(d, cov); }

void launch_pair_count(const BaDev& d, int* cnt, cudaStream_t s) { k_pair_count<<<148 * 4, 256, 0, s>>>(d, cnt); }
void launch_pair_fill(const BaDev& d, int* cursor, int2* inc, cudaStream_t s) { k_pair_fill<<<148 * 4, 256, 0, s>>>(d, cursor, inc); }
void launch_pair_items(const BaDev& d, int* cnt, int4* items, int* n_items_out, cudaStream_t s) { k_pair_items<<<1, 1024, 0, s>>>(d, cnt, items, n_items_out); }
// all candidates of a trial round in one pass (mc.n_cand = 2 or 3); the per-candidate Sm / rm must be zeroed before
// the opt-in shared-memory size of a kernel is a per-device attribute: set it the first time a kernel is launched on
// each device of the process (a handle may live on any device)
static bool first_launch_on_device(bool (&seen)[64])
{
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if (seen[dev]) return false;
  seen[dev] = true;
  return true;
}

template <int NC, int MGT>
static void launch_pairs_multi_tma(const BaDev& d, const SchurMulti& mc, cudaStream_t s)
{
  const size_t smem = (size_t)TW * 2 * MGT * MRECD * sizeof(double) + TW * 2 * sizeof(unsigned long long);
  static bool seen[64] = { false };
  if (first_launch_on_device(seen)) cudaFuncSetAttribute(k_schur_pairs_multi<NC, MGT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int per_sm = (int)std::min<size_t>(6, (220 * 1024) / smem);
  int g2 = (d.max_items + TW - 1) / TW;
  if (g2 < 1) g2 = 1;
  if (g2 > 148 * per_sm) g2 = 148 * per_sm;
  k_schur_pairs_multi<NC, MGT><<<g2, TW * 32, smem, s>>>(d, mc);
}

static int pairs_multi_grid(const BaDev& d)
{
  int g2 = (d.max_items + TW - 1) / TW;
  if (g2 < 1) g2 = 1;
  if (g2 > 148 * 3) g2 = 148 * 3;
  return g2;
}
// first_dynamic_item of the cp.async pair kernel (every warp starts on its own index): needed by whoever forms the point records
void schur_multi_prepare(const BaDev& d, SchurMulti& mc) { mc.first_dynamic_item = pairs_multi_grid(d) * TW; }

void launch_schur_multi(const BaDev& d, const SchurMulti& mc_in, cudaStream_t s, bool records_ready)
{
  // MCP_BA_SCHUR_STAGE: "ca" (default) = 16-byte cp.async pieces through the LSU; "tma8" / "tma16" = three bulk
  // copies per incidence, groups of 8 / 16 incidences (measured slower: profiles/README.md)
  static const int stage_mode = [] {
    const char* e = getenv("MCP_BA_SCHUR_STAGE");
    if (!e || e[0] != 't') return 2;
    return (e[1] && e[2] && e[3] == '1') ? 1 : 0;
  }();
  SchurMulti mc = mc_in;
  const int g2 = pairs_multi_grid(d);
  mc.first_dynamic_item = g2 * TW;
  if (stage_mode != 2) records_ready = false;          // the bulk-copy variants do not use the work counter: keep their own launch
  if (!records_ready) {
    int g1 = (d.p_hi - d.p_lo + 255) / 256;
    if (g1 < 148) g1 = 148;
    if (mc.n_cand == 2) launch_chain(k_schur_vinv_multi<2>, dim3(g1), dim3(256), 0, s, d, mc);
    else launch_chain(k_schur_vinv_multi<3>, dim3(g1), dim3(256), 0, s, d, mc);
  }
  if (stage_mode == 0) {
    if (mc.n_cand == 2) launch_pairs_multi_tma<2, 8>(d, mc, s); else launch_pairs_multi_tma<3, 8>(d, mc, s);
    return;
  }
  if (stage_mode == 1) {
    if (mc.n_cand == 2) launch_pairs_multi_tma<2, 16>(d, mc, s); else launch_pairs_multi_tma<3, 16>(d, mc, s);
    return;
  }
  const size_t smem_ca = (size_t)TW * 2 * MG_CA * MRECD * sizeof(double) + (size_t)TW * 2 * 3 * MG_CA * sizeof(unsigned long long);
  static bool seen_ca[64] = { false };
  if (first_launch_on_device(seen_ca)) {
    cudaFuncSetAttribute(k_schur_pairs_multi_ca<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ca);
    cudaFuncSetAttribute(k_schur_pairs_multi_ca<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ca);
  }
  if (mc.n_cand == 2) launch_chain(k_schur_pairs_multi_ca<2>, dim3(g2), dim3(TW * 32), smem_ca, s, d, mc);
  else launch_chain(k_schur_pairs_multi_ca<3>, dim3(g2), dim3(TW * 32), smem_ca, s, d, mc);
}

void launch_schur_gather(const BaDev& d, cudaStream_t s)
{
  const int nslots = d.slot_hi - d.slot_lo;
  int g1 = (nslots + 255) / 256;
  if (g1 < 1) g1 = 1;
  if (g1 > 148 * 8) g1 = 148 * 8;
  k_schur_y<<<g1, 256, 0, s>>>(d);
  if (d.schur_mode == 0) {
    const size_t per_warp = (size_t)(d.rs_nblk + 8) * 48 * sizeof(double) + 2 * RS_BYTES + 2 * 64 * sizeof(int) + 2 * sizeof(unsigned long long);
    int warps = 4;
    while (warps > 1 && per_warp * warps > 100 * 1024) warps >>= 1;
    const size_t smem = per_warp * warps;
    cudaFuncSetAttribute(k_schur_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     // size varies per problem; non-default path
    int g = (d.n_rs_items + warps - 1) / warps;
    if (g < 1) g = 1;
    k_schur_rows<<<g, warps * 32, smem, s>>>(d);
    return;
  }
  const int use_v1 = d.schur_mode == 2;
  int g2 = (d.max_items + 3) / 4;                       // upper bound; the kernels read the count from the device
  if (g2 < 1) g2 = 1;
  if (use_v1) {
    if (g2 > 148 * 5) g2 = 148 * 5;
    k_schur_pairs<<<g2, 128, 0, s>>>(d);
  } else {
    const size_t smem = (size_t)TW * 2 * TG * SREC * sizeof(double) + TW * 2 * sizeof(unsigned long long);
    static bool seen_tma[64] = { false };
    if (first_launch_on_device(seen_tma)) cudaFuncSetAttribute(k_schur_pairs_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (g2 > 148 * 2) g2 = 148 * 2;
    k_schur_pairs_tma<<<g2, TW * 32, smem, s>>>(d);
  }
}

}  // namespace mcp
