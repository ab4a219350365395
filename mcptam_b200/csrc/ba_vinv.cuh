// ba_vinv.cuh — per-point records of the multi-candidate Schur reduction (shared by ba_schur.cu and ba_kernels.cu).
//   R_p = [Vinv_0 | Vinv_1 | Vinv_2 | u_0 | u_1 | u_2],  Vinv_c = (V_p + lambda_c I)^-1,  u_c = Vinv_c g_p     (288 B)
// lambda_c follows g2o's rejection schedule (lambda, lambda*ni, lambda*ni*2ni, ...; src/ChainBundle.cc:1009-1118 drives
// g2o::OptimizationAlgorithmLevenberg).  The body also clears the candidates' [Sm | rm] accumulators and re-arms the work
// counter of k_schur_pairs_multi_ca.  It runs either as k_schur_vinv_multi (a trial round after a rejection) or inside
// extra blocks of k_pose_blocks (the round that follows a linearisation: lambda is already known, so the records are
// formed next to the pose blocks and the vinv launch leaves the critical path).
#pragma once
#include "ba_types.cuh"

namespace mcp {

constexpr int RPD = 36;                      // doubles per point record

__device__ __forceinline__ bool inv3_sym_s(const double* V6, double lambda, double* Vi)
{
  const double a = V6[0] + lambda, b = V6[1], c = V6[2], dd = V6[3] + lambda, e = V6[4], f = V6[5] + lambda;
  const double c00 = dd * f - e * e, c01 = c * e - b * f, c02 = b * e - c * dd;
  const double det = a * c00 + b * c01 + c * c02;
  const double id = 1.0 / det;
  Vi[0] = c00 * id; Vi[1] = c01 * id; Vi[2] = c02 * id;
  Vi[3] = Vi[1]; Vi[4] = (a * f - c * c) * id; Vi[5] = (b * c - a * e) * id;
  Vi[6] = Vi[2]; Vi[7] = Vi[5]; Vi[8] = (a * dd - b * b) * id;
  return (a > 0) && (a * dd - b * b > 0) && (det > 0) && isfinite(id);
}

// bid / nblk: index of this block among the blocks that run the body
template <int NC>
__device__ __forceinline__ void schur_vinv_body(const BaDev& d, const SchurMulti& mc, int bid, int nblk)
{
  double lam[NC];
  {
    double l = d.ctrl->lambda, ni = d.ctrl->ni;
#pragma unroll
    for (int c = 0; c < NC; c++) { lam[c] = l; l *= ni; ni *= 2; }
  }
  if (bid == 0 && threadIdx.x == 0) *mc.next_item = mc.first_dynamic_item;
  // the reduced systems accumulate by atomics: clear them here instead of one memset per candidate in the stream
#pragma unroll
  for (int c = 0; c < NC; c++)
    if (mc.zero_mask & (1 << c))
      for (size_t i = (size_t)bid * blockDim.x + threadIdx.x; i < mc.sm_doubles; i += (size_t)nblk * blockDim.x) mc.Sm[c][i] = 0.0;
  int fail = 0;
  for (int p = d.p_lo + bid * blockDim.x + threadIdx.x; p < d.p_hi; p += nblk * blockDim.x) {
    if (d.pt_var[p] < 0) continue;
    double V6[6], gp[3];
#pragma unroll
    for (int i = 0; i < 6; i++) V6[i] = d.V[6 * (size_t)p + i];
#pragma unroll
    for (int i = 0; i < 3; i++) gp[i] = d.gp[3 * (size_t)p + i];
    double* R = mc.R + RPD * (size_t)p;
#pragma unroll
    for (int c = 0; c < NC; c++) {
      double Vi[9];
      if (!inv3_sym_s(V6, lam[c], Vi)) fail |= 1 << c;
#pragma unroll
      for (int i = 0; i < 9; i++) R[9 * c + i] = Vi[i];
#pragma unroll
      for (int r = 0; r < 3; r++) R[27 + 3 * c + r] = Vi[3 * r] * gp[0] + Vi[3 * r + 1] * gp[1] + Vi[3 * r + 2] * gp[2];
    }
#pragma unroll
    for (int c = NC; c < 3; c++) {
#pragma unroll
      for (int i = 0; i < 9; i++) R[9 * c + i] = 0.0;
#pragma unroll
      for (int r = 0; r < 3; r++) R[27 + 3 * c + r] = 0.0;
    }
  }
#pragma unroll
  for (int c = 0; c < NC; c++)
    if (fail & (1 << c)) atomicExch(&d.ctrl->solve_ok[c], 0);
}

}  // namespace mcp
