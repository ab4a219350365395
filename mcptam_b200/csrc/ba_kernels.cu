// ba_kernels.cu — sm_100a kernels of the ChainBundle LM bundle adjuster.
//
//   k_linearize         eight lanes per map point: per-measurement TaylorCamera reprojection, 2x6 / 2x3
//                       Jacobians (reference src/ChainBundle.cc:376-397, 449-685), robust weights
//                       (:871-897), per-point 3x3 / 6x3 blocks in shared memory, pose-pose blocks by fp64
//                       atomics.  (The Schur reduction for each lambda lives in ba_schur.cu.)
//   k_select_cluster /  exact upper median of |chi2| (radix select) -> Huber / Tukey sigma^2
//   k_sel_pass          (include/mcptam/MEstimator.h:109-126,194-204; src/ChainBundle.cc:810-833).
//   k_lambda_init       g2o computeLambdaInit: 1e-5 * max diagonal.
//   k_solve             dense Cholesky of the damped reduced camera system + pose update (:82-86).
//   k_backsub_eval      per point back-substitution, VertexRelPoint::oplusImpl (:237-281) and the trial
//                       error evaluation.
//   k_lm_control        accept / reject, lambda schedule, convergence actions (:1009-1118).
#include "ba_types.cuh"
#include "ba_vinv.cuh"
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>

namespace mcp {

// ---------------------------------------------------------------------------------------------
// shared device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void robustify(const BaCtrl* __restrict__ c, double e2, double& rho0, double& rho1)
{
  if (!c->use_robust) { rho0 = e2; rho1 = 1.0; return; }
  if (e2 <= c->sigma_sq_lim) { rho0 = fabs(e2); rho1 = 1.0; }
  else { const double e = sqrt(e2); rho0 = 2 * c->sigma_lim * e - c->sigma_sq_lim; rho1 = c->sigma_lim / e; }
}
// the same with the kernel's own copy of the Huber parameters (k_linearize may have picked them itself)
__device__ __forceinline__ void robustify(bool use_robust, double sigma_sq_lim, double sigma_lim, double e2, double& rho0, double& rho1)
{
  if (!use_robust) { rho0 = e2; rho1 = 1.0; return; }
  if (e2 <= sigma_sq_lim) { rho0 = fabs(e2); rho1 = 1.0; }
  else { const double e = sqrt(e2); rho0 = 2 * sigma_lim * e - sigma_sq_lim; rho1 = sigma_lim / e; }
}

struct PtCtx {
  Se3 Bs;          // source MKF pose (base from world)
  double RCcs[9];  // rotation of the source cam-from-base link (identity for 1-link chains)
  double pw[3];    // point in world frame
  double qs[3];    // point in the source MKF frame
  double prel[3];  // point in the source camera frame (the estimate)
};

__device__ __forceinline__ void load_pt_ctx(const BaDev& d, const double* pose, const int4 pi,
                                            const double* prel, PtCtx& c)
{
  se3_load(pose + 12 * (size_t)pi.x, c.Bs);
  c.prel[0] = prel[0]; c.prel[1] = prel[1]; c.prel[2] = prel[2];
  if (pi.y >= 0) {
    Se3 C;
    se3_load(pose + 12 * (size_t)pi.y, C);
#pragma unroll
    for (int i = 0; i < 9; i++) c.RCcs[i] = C.R[i];
    se3_apply_inv(C, prel, c.qs);
  } else {
#pragma unroll
    for (int i = 0; i < 9; i++) c.RCcs[i] = (i % 4 == 0) ? 1.0 : 0.0;
    c.qs[0] = prel[0]; c.qs[1] = prel[1]; c.qs[2] = prel[2];
  }
  se3_apply_inv(c.Bs, c.qs, c.pw);
}

// Tangent basis of VertexRelPoint (src/ChainBundle.cc:595-620): columns are the camera-frame motion of the
// point for d_beta, d_alpha, d_rho.  M row-major 3x3.
__device__ __forceinline__ void point_tangent(const double* p, double* M)
{
  const double len = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  const double rho = 1.0 / len;
  const double dir[3] = { p[0] * rho, p[1] * rho, p[2] * rho };
  double axis[3] = { dir[1], -dir[0], 0.0 };   // dir ^ (0,0,1)
  const double an = sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
  const double angle = asin(an);
  axis[0] = axis[0] / an * angle; axis[1] = axis[1] / an * angle; axis[2] = axis[2] / an * angle;
  double Rp[9];
  so3_exp(axis, Rp);
  double rpp[3];
  m3_vec(Rp, p, rpp);
  const double g0[3] = { 0.0, -rpp[2], rpp[1] };   // SO3 generator 0 on Rp*p
  const double g1[3] = { rpp[2], 0.0, -rpp[0] };   // generator 1
  double c0[3], c1[3];
  m3t_vec(Rp, g0, c0);
  m3t_vec(Rp, g1, c1);
#pragma unroll
  for (int r = 0; r < 3; r++) { M[r * 3 + 0] = c0[r]; M[r * 3 + 1] = c1[r]; M[r * 3 + 2] = -1 * p[r] / rho; }
}

// VertexRelPoint::oplusImpl (src/ChainBundle.cc:237-281)
__device__ __forceinline__ void point_oplus(const double* est, const double* upd, double* out)
{
  const double dist_before = sqrt(est[0] * est[0] + est[1] * est[1] + est[2] * est[2]);
  const double rho_before = 1.0 / dist_before;
  const double dir[3] = { est[0] * rho_before, est[1] * rho_before, est[2] * rho_before };
  double axis[3] = { dir[1], -dir[0], 0.0 };
  const double an = sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
  const double angle = asin(an);
  axis[0] = axis[0] / an * angle; axis[1] = axis[1] / an * angle; axis[2] = axis[2] / an * angle;
  double Rp[9], Ru[9], M1[9], M2[9], RpT[9];
  so3_exp(axis, Rp);
  const double w[3] = { upd[0], upd[1], 0.0 };
  so3_exp(w, Ru);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) RpT[i * 3 + j] = Rp[j * 3 + i];
  m3_mul(RpT, Ru, M1);
  m3_mul(M1, Rp, M2);
  double c[3];
  m3_vec(M2, dir, c);
  const double s = 1 / (rho_before + upd[2]);
  out[0] = s * c[0]; out[1] = s * c[1]; out[2] = s * c[2];
  const double dist_after = sqrt(out[0] * out[0] + out[1] * out[1] + out[2] * out[2]);
  if (dist_after > 1e5) { const double f = 1e5 / dist_after; out[0] *= f; out[1] *= f; out[2] *= f; }
  if (dist_after < 1e-5) { const double f = 1e-5 / dist_after; out[0] *= f; out[1] *= f; out[2] *= f; }
}

// Per-measurement geometry: residual and the three 2x3 pixel-motion maps
//   A  : motion in the observing MKF frame  -> pixel     (J_obs = -A  * Gamma(q))
//   A2 : motion in the source MKF frame     -> pixel     (J_src = +A2 * Gamma(qs))
//   A3 : motion in the source camera frame  -> pixel     (J_pt  = -A3 * M)
struct MeasGeom {
  double e[2];
  double q[3];
  double A[6], A2[6], A3[6];
};

template <bool WITH_JAC>
__device__ __forceinline__ void meas_geometry(const DevCam* cams, const double* pose, const PtCtx& c,
                                              const int4 ma, const double2 z, MeasGeom& g)
{
  Se3 Bm;
  se3_load(pose + 12 * (size_t)ma.x, Bm);
  se3_apply(Bm, c.pw, g.q);
  double v[3];
  double RC[9];
  if (ma.y >= 0) {
    Se3 C;
    se3_load(pose + 12 * (size_t)ma.y, C);
    se3_apply(C, g.q, v);
#pragma unroll
    for (int i = 0; i < 9; i++) RC[i] = C.R[i];
  } else {
    v[0] = g.q[0]; v[1] = g.q[1]; v[2] = g.q[2];
#pragma unroll
    for (int i = 0; i < 9; i++) RC[i] = (i % 4 == 0) ? 1.0 : 0.0;
  }
  double px[2], G[6];
  cam_project(cams[ma.z], v, px, WITH_JAC ? G : nullptr);
  g.e[0] = z.x - px[0];
  g.e[1] = z.y - px[1];
  if (WITH_JAC) {
    // A = G * RC
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
      for (int k = 0; k < 3; k++) g.A[r * 3 + k] = G[r * 3] * RC[k] + G[r * 3 + 1] * RC[3 + k] + G[r * 3 + 2] * RC[6 + k];
    // T = A * R_Bm ; A2 = T * R_Bs^T ; A3 = A2 * R_Ccs^T
    double T[6];
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
      for (int k = 0; k < 3; k++) T[r * 3 + k] = g.A[r * 3] * Bm.R[k] + g.A[r * 3 + 1] * Bm.R[3 + k] + g.A[r * 3 + 2] * Bm.R[6 + k];
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
      for (int k = 0; k < 3; k++) g.A2[r * 3 + k] = T[r * 3] * c.Bs.R[k * 3] + T[r * 3 + 1] * c.Bs.R[k * 3 + 1] + T[r * 3 + 2] * c.Bs.R[k * 3 + 2];
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
      for (int k = 0; k < 3; k++) g.A3[r * 3 + k] = g.A2[r * 3] * c.RCcs[k * 3] + g.A2[r * 3 + 1] * c.RCcs[k * 3 + 1] + g.A2[r * 3 + 2] * c.RCcs[k * 3 + 2];
  }
}

// The same for k_linearize with the per-point context in shared memory (one copy per 8-lane group): 33 doubles that would
// otherwise stay in registers across the whole measurement loop next to 30 accumulators (the kernel sat at the 255-register
// limit with spills; with the context in shared memory it fits 168 registers = 12 warps per SM).
//   cs[0..2] point in the world frame, cs[3..11] R of the source MKF pose, cs[12..20] R of the source cam-from-base link,
//   cs[21..23] point in the source MKF frame, cs[24..32] tangent basis M (point_tangent)
constexpr int CTXD = 33;
__device__ __forceinline__ void meas_geometry_s(const DevCam* cams, const double* pose, const double* cs, const int4 ma, const double2 z, MeasGeom& g)
{
  Se3 Bm;
  se3_load(pose + 12 * (size_t)ma.x, Bm);
  {
    const double pw[3] = { cs[0], cs[1], cs[2] };
    se3_apply(Bm, pw, g.q);
  }
  double v[3];
  double RC[9];
  if (ma.y >= 0) {
    Se3 C;
    se3_load(pose + 12 * (size_t)ma.y, C);
    se3_apply(C, g.q, v);
#pragma unroll
    for (int i = 0; i < 9; i++) RC[i] = C.R[i];
  } else {
    v[0] = g.q[0]; v[1] = g.q[1]; v[2] = g.q[2];
#pragma unroll
    for (int i = 0; i < 9; i++) RC[i] = (i % 4 == 0) ? 1.0 : 0.0;
  }
  double px[2], G[6];
  cam_project(cams[ma.z], v, px, G);
  g.e[0] = z.x - px[0];
  g.e[1] = z.y - px[1];
#pragma unroll
  for (int r = 0; r < 2; r++)
#pragma unroll
    for (int k = 0; k < 3; k++) g.A[r * 3 + k] = G[r * 3] * RC[k] + G[r * 3 + 1] * RC[3 + k] + G[r * 3 + 2] * RC[6 + k];
  double T[6];
#pragma unroll
  for (int r = 0; r < 2; r++)
#pragma unroll
    for (int k = 0; k < 3; k++) T[r * 3 + k] = g.A[r * 3] * Bm.R[k] + g.A[r * 3 + 1] * Bm.R[3 + k] + g.A[r * 3 + 2] * Bm.R[6 + k];
  const double* BsR = cs + 3;
#pragma unroll
  for (int r = 0; r < 2; r++)
#pragma unroll
    for (int k = 0; k < 3; k++) g.A2[r * 3 + k] = T[r * 3] * BsR[k * 3] + T[r * 3 + 1] * BsR[k * 3 + 1] + T[r * 3 + 2] * BsR[k * 3 + 2];
  const double* RCcs = cs + 12;
#pragma unroll
  for (int r = 0; r < 2; r++)
#pragma unroll
    for (int k = 0; k < 3; k++) g.A3[r * 3 + k] = g.A2[r * 3] * RCcs[k * 3] + g.A2[r * 3 + 1] * RCcs[k * 3 + 1] + g.A2[r * 3 + 2] * RCcs[k * 3 + 2];
}

// J (2x6, row-major) = sign * A * Gamma(q),  Gamma(q) = [ I3 | e_k x q ]  (TooN SE3 generator field)
__device__ __forceinline__ void pose_jac(const double* A, const double* q, double sign, double* J)
{
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const double a0 = A[r * 3], a1 = A[r * 3 + 1], a2 = A[r * 3 + 2];
    J[r * 6 + 0] = sign * a0;
    J[r * 6 + 1] = sign * a1;
    J[r * 6 + 2] = sign * a2;
    J[r * 6 + 3] = sign * (-a1 * q[2] + a2 * q[1]);
    J[r * 6 + 4] = sign * (a0 * q[2] - a2 * q[0]);
    J[r * 6 + 5] = sign * (-a0 * q[1] + a1 * q[0]);
  }
}

__device__ __forceinline__ bool inv3_sym(const double* V6, double lambda, double* Vi)
{
  // V6 = {v00, v01, v02, v11, v12, v22}; returns inverse (full 3x3 row-major) of V + lambda I
  const double a = V6[0] + lambda, b = V6[1], c = V6[2], dd = V6[3] + lambda, e = V6[4], f = V6[5] + lambda;
  const double c00 = dd * f - e * e, c01 = c * e - b * f, c02 = b * e - c * dd;
  const double det = a * c00 + b * c01 + c * c02;
  const double id = 1.0 / det;
  Vi[0] = c00 * id; Vi[1] = c01 * id; Vi[2] = c02 * id;
  Vi[3] = Vi[1]; Vi[4] = (a * f - c * c) * id; Vi[5] = (b * c - a * e) * id;
  Vi[6] = Vi[2]; Vi[7] = Vi[5]; Vi[8] = (a * dd - b * b) * id;
  // SPD check (leading minors) – a failed point makes the whole solve fail, as CHOLMOD would.
  return (a > 0) && (a * dd - b * b > 0) && (det > 0) && isfinite(id);
}

__device__ __forceinline__ double block_sum(double v, double* red /*>= 32 doubles of smem*/)
{
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double r = 0;
  if (wid == 0) {
    r = (lane < (int)(blockDim.x >> 5)) ? red[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;  // valid in warp 0
}

// ---------------------------------------------------------------------------------------------
// Work distribution of the per-point kernels: a map point has ~8 measurements on average, so a full warp per
// point leaves three quarters of the lanes idle in the per-measurement geometry.  Each point gets a group of
// LG = 8 lanes (four points per warp); points are visited through d.pt_order (sorted by measurement count,
// heaviest first) so that the four groups of a warp run the same number of iterations and the block scheduler
// sees the long units first.
// ---------------------------------------------------------------------------------------------
// The pose array and the camera models are read with data-dependent indices by every measurement; staged in shared
// memory the dependent load costs ~30 cycles instead of an L2 round trip (the per-point kernels are latency bound).
__device__ __forceinline__ void stage_pose_cams(const BaDev& d, const double* pose_g, double* sm, const double*& pose, const DevCam*& cams)
{
  if (d.stage_doubles == 0) { pose = pose_g; cams = d.cams; return; }
  const int np = d.n_pose * 12, ncd = d.n_cam * (int)(sizeof(DevCam) / 8);
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(sm);
  const unsigned long long* pg = reinterpret_cast<const unsigned long long*>(pose_g);
  const unsigned long long* cg = reinterpret_cast<const unsigned long long*>(d.cams);
  for (int i = threadIdx.x; i < np; i += blockDim.x) dst[i] = pg[i];
  for (int i = threadIdx.x; i < ncd; i += blockDim.x) dst[np + i] = cg[i];
  __syncthreads();
  pose = sm;
  cams = reinterpret_cast<const DevCam*>(sm + np);
}

constexpr int LG = 8;            // lanes per map point
constexpr int PPW = 32 / LG;     // points per warp

__device__ __forceinline__ double group_sum(double v, unsigned gmask)
{
#pragma unroll
  for (int o = LG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// k_linearize
// ---------------------------------------------------------------------------------------------
template <int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_linearize(BaDev d)
{
  pdl_prologue();
  extern __shared__ __align__(16) double smem[];
  __shared__ double red[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int gl = lane & (LG - 1), grp = lane / LG;
  const unsigned gmask = ((1u << LG) - 1u) << (grp * LG);
  if (lookahead_skip(d)) return;
  double* Wsm = smem + d.stage_doubles + (size_t)(wid * PPW + grp) * ((size_t)d.max_slots * 18 + CTXD);
  double* cs = Wsm + (size_t)d.max_slots * 18;             // this group's point context
  const BaCtrl* ctrl = d.ctrl;
  const int cur = ctrl->cur;
  // Huber parameters of this linearisation.  A look-ahead launch (d.pick_sigma) follows an accepted trial whose state is the
  // new linearisation point: its sigma^2 was computed next to the trial (k_select_cluster mode 3) and is adopted here by
  // every block for itself -- RobustKernelData::RecomputeNow -- block 0 also records it in the control block.
  __shared__ double s_sig[2];
  if (threadIdx.x == 0) {
    double lim2 = ctrl->sigma_sq_lim, lim = ctrl->sigma_lim;
    if (d.pick_sigma && ctrl->accepted) {
      const double raw = d.spec_sigma[ctrl->acc_cand];
      lim2 = raw < ctrl->min_sigma_sq ? ctrl->min_sigma_sq : raw;
      lim = sqrt(lim2);
      if (blockIdx.x == 0) { BaCtrl* cw = d.ctrl; cw->sigma_sq_raw = raw; cw->sigma_sq_lim = lim2; cw->sigma_lim = lim; cw->med_hint = d.spec_med[ctrl->acc_cand]; }
    }
    s_sig[0] = lim2; s_sig[1] = lim;
  }
  __syncthreads();
  const double sig_sq_lim = s_sig[0], sig_lim = s_sig[1];
  const bool use_robust = ctrl->use_robust != 0;
  const double* pose;
  const DevCam* cams;
  stage_pose_cams(d, d.pose[cur], smem, pose, cams);
  const double* __restrict__ ptv = d.pt[cur];
  const int nc = d.nc;
  const int n_local = d.p_hi - d.p_lo;
  double chi_acc = 0.0;

  for (int u = blockIdx.x * nw + wid; u * PPW < n_local; u += gridDim.x * nw) {
    const int idx = u * PPW + grp;
    if (idx >= n_local) continue;
    const int p = d.pt_order[d.p_lo + idx];
    const int4 pi = d.pt_info[p];
    const int pvar = d.pt_var[p];
    const double prel[3] = { ptv[3 * (size_t)p], ptv[3 * (size_t)p + 1], ptv[3 * (size_t)p + 2] };
    {
      PtCtx c;
      load_pt_ctx(d, pose, pi, prel, c);
      double M[9];
      point_tangent(prel, M);
      if (gl == 0) {
#pragma unroll
        for (int i = 0; i < 3; i++) { cs[i] = c.pw[i]; cs[21 + i] = c.qs[i]; }
#pragma unroll
        for (int i = 0; i < 9; i++) { cs[3 + i] = c.Bs.R[i]; cs[12 + i] = c.RCcs[i]; cs[24 + i] = M[i]; }
      }
    }
    const double* M = cs + 24;
    const int s0 = d.pt_slot_off[p], K = d.pt_slot_off[p + 1] - s0;
    for (int i = gl; i < K * 18; i += LG) Wsm[i] = 0.0;
    __syncwarp(gmask);

    // sums over the measurements of this point
    double P3[6] = { 0, 0, 0, 0, 0, 0 }, t3[3] = { 0, 0, 0 };
    double Q[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 }, P2[6] = { 0, 0, 0, 0, 0, 0 }, t2[3] = { 0, 0, 0 };
    const int m0 = d.pt_meas_off[p], m1 = d.pt_meas_off[p + 1];
    for (int m = m0 + gl; m < m1; m += LG) {
      const int4 ma = d.meas_a[m];
      const int4 mbi = d.meas_b[m];
      const double2 z = d.meas_xy[m];
      const double info = d.meas_info[m];
      MeasGeom g;
      meas_geometry_s(cams, pose, cs, ma, z, g);
      double chi2 = info * (g.e[0] * g.e[0] + g.e[1] * g.e[1]);
      if (pvar < 0 && use_robust) chi2 = -chi2;                   // src/ChainBundle.cc:413-414
      double rho0, rho1;
      robustify(use_robust, sig_sq_lim, sig_lim, chi2, rho0, rho1);
      chi_acc += rho0;
      const double w = rho1 * info;
      const double we0 = w * g.e[0], we1 = w * g.e[1];
      if (pvar >= 0) {
        // P3 += w A3^T A3 ; t3 += w A3^T e
        P3[0] += w * (g.A3[0] * g.A3[0] + g.A3[3] * g.A3[3]);
        P3[1] += w * (g.A3[0] * g.A3[1] + g.A3[3] * g.A3[4]);
        P3[2] += w * (g.A3[0] * g.A3[2] + g.A3[3] * g.A3[5]);
        P3[3] += w * (g.A3[1] * g.A3[1] + g.A3[4] * g.A3[4]);
        P3[4] += w * (g.A3[1] * g.A3[2] + g.A3[4] * g.A3[5]);
        P3[5] += w * (g.A3[2] * g.A3[2] + g.A3[5] * g.A3[5]);
#pragma unroll
        for (int k = 0; k < 3; k++) t3[k] += g.A3[k] * we0 + g.A3[3 + k] * we1;
      }
      const bool has_src = mbi.z != 0;
      if (has_src) {
#pragma unroll
        for (int r = 0; r < 3; r++) {
          t2[r] += g.A2[r] * we0 + g.A2[3 + r] * we1;
          if (pvar >= 0) {
#pragma unroll
            for (int k = 0; k < 3; k++) Q[r * 3 + k] += w * (g.A2[r] * g.A3[k] + g.A2[3 + r] * g.A3[3 + k]);
          }
        }
        P2[0] += w * (g.A2[0] * g.A2[0] + g.A2[3] * g.A2[3]);
        P2[1] += w * (g.A2[0] * g.A2[1] + g.A2[3] * g.A2[4]);
        P2[2] += w * (g.A2[0] * g.A2[2] + g.A2[3] * g.A2[5]);
        P2[3] += w * (g.A2[1] * g.A2[1] + g.A2[4] * g.A2[4]);
        P2[4] += w * (g.A2[1] * g.A2[2] + g.A2[4] * g.A2[5]);
        P2[5] += w * (g.A2[2] * g.A2[2] + g.A2[5] * g.A2[5]);
      }
      const int vo = mbi.x;
      if (vo >= 0) {
        // The pose-pose blocks (U_oo, g_o and the observer/source cross block) are NOT accumulated here: 63 fp64
        // global atomics per measurement made this kernel atomic-throughput bound.  The measurement leaves a
        // record instead and k_pose_blocks reduces the records pose block by pose block.
        double2* rec = reinterpret_cast<double2*>(d.mrec + (size_t)MREC * m);
        rec[0] = make_double2(g.A[0], g.A[1]); rec[1] = make_double2(g.A[2], g.A[3]); rec[2] = make_double2(g.A[4], g.A[5]);
        rec[3] = make_double2(g.q[0], g.q[1]); rec[4] = make_double2(g.q[2], w); rec[5] = make_double2(we0, we1);
        if (has_src) {
          rec[6] = make_double2(g.A2[0], g.A2[1]); rec[7] = make_double2(g.A2[2], g.A2[3]); rec[8] = make_double2(g.A2[4], g.A2[5]);
          rec[9] = make_double2(cs[21], cs[22]); rec[10] = make_double2(cs[23], (double)vo);
        }
        if (pvar >= 0) {
          double Jo[12];
          pose_jac(g.A, g.q, -1.0, Jo);
          // J_pt = -A3 * M ; W_obs = w Jo^T J_pt
          double Jp[6];
#pragma unroll
          for (int r = 0; r < 2; r++)
#pragma unroll
            for (int k = 0; k < 3; k++) Jp[r * 3 + k] = -(g.A3[r * 3] * M[k] + g.A3[r * 3 + 1] * M[3 + k] + g.A3[r * 3 + 2] * M[6 + k]);
          double* Wo = Wsm + mbi.y * 18;
#pragma unroll
          for (int r = 0; r < 6; r++)
#pragma unroll
            for (int k = 0; k < 3; k++) atomicAdd(Wo + r * 3 + k, w * (Jo[r] * Jp[k] + Jo[6 + r] * Jp[3 + k]));
        }
      }
    }
    // group-reduce the point sums (every lane of the group ends up with the totals)
#pragma unroll
    for (int i = 0; i < 6; i++) { P3[i] = group_sum(P3[i], gmask); P2[i] = group_sum(P2[i], gmask); }
#pragma unroll
    for (int i = 0; i < 3; i++) { t3[i] = group_sum(t3[i], gmask); t2[i] = group_sum(t2[i], gmask); }
#pragma unroll
    for (int i = 0; i < 9; i++) Q[i] = group_sum(Q[i], gmask);

    // point block: V = M^T P3 M, gp = M^T t3   (J_pt = -A3 M, b_p = -sum w J_pt^T e)
    double V6[6] = { 0, 0, 0, 0, 0, 0 }, gp[3] = { 0, 0, 0 };
    if (pvar >= 0) {
      const double P[9] = { P3[0], P3[1], P3[2], P3[1], P3[3], P3[4], P3[2], P3[4], P3[5] };
      double PM[9];
      m3_mul(P, M, PM);
      V6[0] = M[0] * PM[0] + M[3] * PM[3] + M[6] * PM[6];
      V6[1] = M[0] * PM[1] + M[3] * PM[4] + M[6] * PM[7];
      V6[2] = M[0] * PM[2] + M[3] * PM[5] + M[6] * PM[8];
      V6[3] = M[1] * PM[1] + M[4] * PM[4] + M[7] * PM[7];
      V6[4] = M[1] * PM[2] + M[4] * PM[5] + M[7] * PM[8];
      V6[5] = M[2] * PM[2] + M[5] * PM[5] + M[8] * PM[8];
      m3t_vec(M, t3, gp);
    }
    // source-pose blocks: U_ss = Gs^T P2 Gs, g_s = -Gs^T t2, W_src = -Gs^T (Q M)
    const int vs = pi.z;
    const bool any_src = (P2[0] != 0.0) || (P2[3] != 0.0) || (P2[5] != 0.0);
    if (vs >= 0 && any_src) {
      // Gs (3x6) = [I | o_k(qs)], o_0=(0,-q2,q1) o_1=(q2,0,-q0) o_2=(-q1,q0,0)
      const double q0 = cs[21], q1 = cs[22], q2 = cs[23];
      const double Gs[18] = { 1, 0, 0, 0, q2, -q1, 0, 1, 0, -q2, 0, q0, 0, 0, 1, q1, -q0, 0 };
      const double P[9] = { P2[0], P2[1], P2[2], P2[1], P2[3], P2[4], P2[2], P2[4], P2[5] };
      for (int e = gl; e < 27; e += LG) {
        if (e < 21) {
          // upper-triangle entry (r,cc) of the 6x6
          int r = 0, k = e;
          while (k >= 6 - r) { k -= 6 - r; r++; }
          const int cc = r + k;
          double acc = 0;
#pragma unroll
          for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) acc += Gs[i * 6 + r] * P[i * 3 + j] * Gs[j * 6 + cc];
          atomicAdd(d.H0 + (size_t)(6 * vs + r) * nc + 6 * vs + cc, acc);
        } else {
          const int r = e - 21;
          atomicAdd(d.gc + 6 * vs + r, -(Gs[r] * t2[0] + Gs[6 + r] * t2[1] + Gs[12 + r] * t2[2]));
        }
      }
      __syncwarp(gmask);
      if (pvar >= 0 && pi.w >= 0) {
        double X[9];
        m3_mul(Q, M, X);
        for (int e = gl; e < 18; e += LG) {
          const int r = e / 3, k = e - r * 3;
          Wsm[pi.w * 18 + e] += -(Gs[r] * X[k] + Gs[6 + r] * X[3 + k] + Gs[12 + r] * X[6 + k]);
        }
      }
    }
    __syncwarp(gmask);
    if (pvar >= 0) {
      double* Wg = d.W + (size_t)s0 * 18;
      for (int i = gl; i < K * 18; i += LG) Wg[i] = Wsm[i];
      if (gl < 6) d.V[6 * (size_t)p + gl] = V6[gl];
      if (gl < 3) d.gp[3 * (size_t)p + gl] = gp[gl];
    }
    __syncwarp(gmask);
  }
  const double tot = block_sum(chi_acc, red);
  if (threadIdx.x == 0) d.part[PART_CUR_CHI * MAX_PARTIALS + blockIdx.x] = tot;
}

// ---------------------------------------------------------------------------------------------
// k_pose_blocks: pose-pose part of the normal equations (g2o constructQuadraticForm [3P]) from the
// per-measurement records of k_linearize.  One warp per work item {block row, block col, begin, end}: a diagonal
// item sums U_vv (upper triangle) and g_v over measurements observed from pose v, an off-diagonal item sums the
// observer/source cross block of one pose pair.  Lane = measurement; the 27 / 36 sums are warp-reduced and added
// with one atomic per entry per item (items of one block are <= 128 measurements each).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_pose_blocks(BaDev d, SchurMulti mc, int n_pb_blocks, double* zero_ptr, size_t zero_n)
{
  pdl_prologue();
  if (lookahead_skip(d)) return;
  // the accumulators of the NEXT linearisation (the other half of the double-buffered [H0 | gc | Sm | rm]) are cleared here,
  // off the critical path: nothing reads them any more once this linearisation has been accepted for execution
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < zero_n; i += (size_t)gridDim.x * blockDim.x) zero_ptr[i] = 0.0;
  if ((int)blockIdx.x >= n_pb_blocks) {
    // extra blocks: the point records of the trial round that follows this linearisation (ba_vinv.cuh)
    if (mc.n_cand == 2) schur_vinv_body<2>(d, mc, blockIdx.x - n_pb_blocks, gridDim.x - n_pb_blocks);
    else schur_vinv_body<3>(d, mc, blockIdx.x - n_pb_blocks, gridDim.x - n_pb_blocks);
    return;
  }
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (n_pb_blocks * blockDim.x) >> 5;
  const int nc = d.nc;
  for (int it = gw; it < d.n_pb_items; it += nw) {
    const int4 item = d.pb_items[it];
    const int lo = item.x, hi = item.y;
    const bool diag = lo == hi;
    double acc[36];
#pragma unroll
    for (int i = 0; i < 36; i++) acc[i] = 0.0;
    for (int e = item.z + lane; e < item.w; e += 32) {
      const int m = d.pb_idx[e];
      const double2* rec = reinterpret_cast<const double2*>(d.mrec + (size_t)MREC * m);
      const double2 r0 = rec[0], r1 = rec[1], r2 = rec[2], r3 = rec[3], r4 = rec[4], r5 = rec[5];
      const double A[6] = { r0.x, r0.y, r1.x, r1.y, r2.x, r2.y };
      const double q[3] = { r3.x, r3.y, r4.x };
      const double w = r4.y, we0 = r5.x, we1 = r5.y;
      double Jo[12];
      pose_jac(A, q, -1.0, Jo);
      if (diag) {
        int t = 0;
#pragma unroll
        for (int r = 0; r < 6; r++) {
#pragma unroll
          for (int cc = r; cc < 6; cc++) acc[t++] += w * (Jo[r] * Jo[cc] + Jo[6 + r] * Jo[6 + cc]);
        }
#pragma unroll
        for (int r = 0; r < 6; r++) acc[21 + r] -= Jo[r] * we0 + Jo[6 + r] * we1;
      } else {
        const double2 r6 = rec[6], r7 = rec[7], r8 = rec[8], r9 = rec[9], r10 = rec[10];
        const double A2[6] = { r6.x, r6.y, r7.x, r7.y, r8.x, r8.y };
        const double qs[3] = { r9.x, r9.y, r10.x };
        double Js[12];
        pose_jac(A2, qs, 1.0, Js);
        const bool obs_first = ((int)r10.y == lo);          // block row = the smaller pose variable
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
          for (int cc = 0; cc < 6; cc++) {
            const double v = obs_first ? (Jo[r] * Js[cc] + Jo[6 + r] * Js[6 + cc]) : (Js[r] * Jo[cc] + Js[6 + r] * Jo[6 + cc]);
            acc[r * 6 + cc] += w * v;
          }
      }
    }
#pragma unroll
    for (int i = 0; i < 36; i++) {
      if (diag && i >= 27) break;
      acc[i] = warp_sum(acc[i]);
    }
    // lane l adds entries l and l + 32
#pragma unroll
    for (int i = 0; i < 36; i++) {
      if ((i & 31) != lane) continue;
      if (diag) {
        if (i < 21) {
          int r = 0, k = i;
          while (k >= 6 - r) { k -= 6 - r; r++; }
          atomicAdd(d.H0 + (size_t)(6 * lo + r) * nc + 6 * lo + r + k, acc[i]);
        } else if (i < 27) atomicAdd(d.gc + 6 * lo + (i - 21), acc[i]);
      } else {
        const int r = i / 6, cc = i - 6 * r;
        atomicAdd(d.H0 + (size_t)(6 * lo + r) * nc + 6 * hi + cc, acc[i]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k_backsub_eval: apply==1: point back-substitution + oplus into the trial buffers, then error eval of
// the trial state.  apply==0: error eval of state `which` only.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_backsub_eval(BaDev d, int apply, int which_in, double* err_out)
{
  pdl_prologue();
  __shared__ double red[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int gl = lane & (LG - 1), grp = lane / LG;
  const unsigned gmask = ((1u << LG) - 1u) << (grp * LG);
  const BaCtrl* ctrl = d.ctrl;
  const int cur = ctrl->cur;
  const int dst = apply ? trial_buffer(d, cur) : (which_in < 0 ? cur : which_in);
  extern __shared__ __align__(16) double smem[];
  const double* pose;
  const DevCam* cams;
  stage_pose_cams(d, d.pose[dst], smem, pose, cams);
  const double lambda = trial_lambda(d);
  const int n_local = d.p_hi - d.p_lo;
  double chi_acc = 0, scale_acc = 0, sumsq_acc = 0;
  for (int u = blockIdx.x * nw + wid; u * PPW < n_local; u += gridDim.x * nw) {
    const int idx = u * PPW + grp;
    if (idx >= n_local) continue;
    const int p = d.pt_order[d.p_lo + idx];
    const int4 pi = d.pt_info[p];
    const int pvar = d.pt_var[p];
    double pnew[3];
    if (apply) {
      const double* pold = d.pt[cur] + 3 * (size_t)p;
      const double po[3] = { pold[0], pold[1], pold[2] };
      if (pvar >= 0) {
        const int s0 = d.pt_slot_off[p], K = d.pt_slot_off[p + 1] - s0;
        const double* Wg = d.W + (size_t)s0 * 18;
        const int* svar = d.slot_var + s0;
        double t[3] = { 0, 0, 0 };
        for (int i = gl; i < K * 18; i += LG) {
          const int a = i / 18, rem = i - a * 18, r = rem / 3, k = rem - r * 3;
          const double v = Wg[i] * d.dc[6 * svar[a] + r];
          t[0] += (k == 0) ? v : 0.0; t[1] += (k == 1) ? v : 0.0; t[2] += (k == 2) ? v : 0.0;
        }
        t[0] = group_sum(t[0], gmask); t[1] = group_sum(t[1], gmask); t[2] = group_sum(t[2], gmask);
        double V6[6], gp[3], Vi[9];
#pragma unroll
        for (int i = 0; i < 6; i++) V6[i] = d.V[6 * (size_t)p + i];
#pragma unroll
        for (int i = 0; i < 3; i++) gp[i] = d.gp[3 * (size_t)p + i];
        inv3_sym(V6, lambda, Vi);
        const double rr[3] = { gp[0] - t[0], gp[1] - t[1], gp[2] - t[2] };
        double dp[3];
        m3_vec(Vi, rr, dp);
        if (!ctrl->solve_ok[d.cand]) { dp[0] = dp[1] = dp[2] = 0.0; }
        point_oplus(po, dp, pnew);
        if (gl == 0) {
#pragma unroll
          for (int i = 0; i < 3; i++) {
            scale_acc += dp[i] * (lambda * dp[i] + gp[i]);
            sumsq_acc += dp[i] * dp[i];
          }
        }
      } else { pnew[0] = po[0]; pnew[1] = po[1]; pnew[2] = po[2]; }
      if (gl < 3) d.pt[dst][3 * (size_t)p + gl] = pnew[gl];
    } else {
      const double* pp = d.pt[dst] + 3 * (size_t)p;
      pnew[0] = pp[0]; pnew[1] = pp[1]; pnew[2] = pp[2];
    }
    PtCtx c;
    load_pt_ctx(d, pose, pi, pnew, c);
    const int m0 = d.pt_meas_off[p], m1 = d.pt_meas_off[p + 1];
    for (int m = m0 + gl; m < m1; m += LG) {
      const int4 ma = d.meas_a[m];
      MeasGeom g;
      meas_geometry<false>(cams, pose, c, ma, d.meas_xy[m], g);
      double chi2 = d.meas_info[m] * (g.e[0] * g.e[0] + g.e[1] * g.e[1]);
      if (pvar < 0 && ctrl->use_robust) chi2 = -chi2;
      d.chi2[dst][m] = chi2;
      if (err_out) { err_out[2 * (size_t)ma.w] = g.e[0]; err_out[2 * (size_t)ma.w + 1] = g.e[1]; }
      double rho0, rho1;
      robustify(ctrl, chi2, rho0, rho1);
      chi_acc += rho0;
    }
  }
  const double a = block_sum(chi_acc, red);
  const double b = block_sum(scale_acc, red);
  const double c2 = block_sum(sumsq_acc, red);
  if (threadIdx.x == 0) {
    d.part[PART_TMP_CHI * MAX_PARTIALS + blockIdx.x] = a;
    d.part[PART_SCALE * MAX_PARTIALS + blockIdx.x] = b;
    d.part[PART_SUMSQ * MAX_PARTIALS + blockIdx.x] = c2;
  }
}

// ---------------------------------------------------------------------------------------------
// k_sel_pass: exact element [n/2] of sorted |chi2| by MSB-first radix select (6 digits of 11 bits).
// One launch per digit, many blocks; the last block to finish a pass (ticket counter) scans the global
// histogram, narrows (prefix, rank) for the next pass and re-arms the histogram.  The last block of the
// last pass turns the median into the Huber (mode 0) or Tukey (mode 1) sigma^2
// (include/mcptam/MEstimator.h:109-126,194-204; src/ChainBundle.cc:810-833, 1376-1383).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sel_pass(BaDev d, int which_in, int pass, int mode)
{
  if (lookahead_skip(d)) return;
  __shared__ unsigned hist[SEL_BINS];
  __shared__ unsigned wsum[8];
  __shared__ int s_last;
  BaCtrl* ctrl = d.ctrl;
  const int which = which_in < 0 ? ctrl->cur : which_in;
  const double* __restrict__ v = d.chi2[which];
  const int n = d.n_meas;
  unsigned long long* state = d.sel_state;            // [2*pass] = prefix, [2*pass+1] = rank
  const unsigned long long prefix = pass == 0 ? 0ull : __ldcg(&state[2 * pass]);
  const int shift = 63 - SEL_BITS * (pass + 1);       // 52, 41, 30, 19, 8, -3
  for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const int i = base + threadIdx.x;
    bool match = false;
    unsigned dig = 0;
    if (i < n) {
      const unsigned long long key = (unsigned long long)__double_as_longlong(fabs(v[i]));
      unsigned long long hi;
      if (shift >= 0) { hi = key >> (shift + SEL_BITS); dig = (unsigned)(key >> shift) & (SEL_BINS - 1); }
      else { hi = key >> (SEL_BITS + shift); dig = (unsigned)(key << (-shift)) & (SEL_BINS - 1); }
      match = (pass == 0) || (hi == prefix);
    }
    const unsigned act = __ballot_sync(0xffffffffu, match);
    if (match) {
      const unsigned peers = __match_any_sync(act, dig);
      if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[dig], (unsigned)__popc(peers));
    }
  }
  __syncthreads();
  unsigned* ghist = d.sel_hist + pass * SEL_BINS;
  for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) if (hist[i]) atomicAdd(&ghist[i], hist[i]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&d.sel_done[pass], 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // ---- last block: locate the bin holding the wanted rank ------------------------------------------
  const unsigned rank = pass == 0 ? (unsigned)(n / 2) : (unsigned)__ldcg(&state[2 * pass + 1]);
  unsigned loc[8], tsum = 0;
#pragma unroll
  for (int q = 0; q < 8; q++) { loc[q] = __ldcg(&ghist[threadIdx.x * 8 + q]); tsum += loc[q]; ghist[threadIdx.x * 8 + q] = 0; }
  // exclusive scan of the 256 per-thread sums
  unsigned incl = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
  if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
  __syncthreads();
  unsigned woff = 0;
  for (int w = 0; w < (int)(threadIdx.x >> 5); w++) woff += wsum[w];
  const unsigned excl = woff + incl - tsum;
  if (rank >= excl && rank < excl + tsum) {
    unsigned acc = excl; int b = 0;
    for (b = 0; b < 8; b++) { if (acc + loc[b] > rank) break; acc += loc[b]; }
    const unsigned bin = threadIdx.x * 8 + b;
    const unsigned long long np = (shift >= 0) ? ((prefix << SEL_BITS) | bin) : ((prefix << (SEL_BITS + shift)) | (bin >> (-shift)));
    if (pass + 1 < SEL_PASSES) { state[2 * (pass + 1)] = np; state[2 * (pass + 1) + 1] = rank - acc; }
    else {
      const double med = __longlong_as_double((long long)np);
      const size_t denom = (size_t)n * 2 - 6;                     // size_t arithmetic as in the reference
      double s = 1.4826 * (1 + 5.0 / (double)denom) * sqrt(med);
      if (mode == 2) ctrl->median_out = med;
      else if (mode == 0) {
        { const double st = 4.6851 * s; double t = st * st; if (t < ctrl->min_sigma_sq) t = ctrl->min_sigma_sq; ctrl->tukey_sigma_sq = t; }   // same median: the Tukey sigma^2 comes along
        s = 1.345 * s;
        ctrl->sigma_sq_raw = s * s;
        ctrl->sigma_sq_lim = ctrl->sigma_sq_raw < ctrl->min_sigma_sq ? ctrl->min_sigma_sq : ctrl->sigma_sq_raw;
        ctrl->sigma_lim = sqrt(ctrl->sigma_sq_lim);
      } else {
        s = 4.6851 * s;
        double t = s * s;
        if (t < ctrl->min_sigma_sq) t = ctrl->min_sigma_sq;
        ctrl->tukey_sigma_sq = t;
      }
    }
  }
  if (threadIdx.x == 0) d.sel_done[pass] = 0;
}

// ---------------------------------------------------------------------------------------------
// k_select_cluster: the same exact radix select in ONE launch for n <= SELC_CAP values.  One thread-block cluster
// of 8 CTAs; every thread keeps its <= 16 keys in registers for all six digits, each CTA histograms into its own
// shared memory, and after a cluster barrier every CTA sums the eight histograms through distributed shared
// memory and narrows (prefix, rank) redundantly -- no global-memory round trips between the passes.
// ---------------------------------------------------------------------------------------------
constexpr int SELC_CTAS = 8, SELC_THREADS = 1024, SELC_K = 16;
constexpr int SELC_CAP = SELC_CTAS * SELC_THREADS * SELC_K;      // 131072

constexpr int SELC_COPIES = 8;                                   // histogram replicas per CTA (lane & 7)
constexpr size_t SELC_SMEM = sizeof(unsigned) * (size_t)(SELC_COPIES + 2) * SEL_BINS;

__global__ void __cluster_dims__(SELC_CTAS, 1, 1) __launch_bounds__(SELC_THREADS) k_select_cluster(BaDev d, int which_in, int mode, double* zero_ptr,
                                                                                                   size_t zero_n)
{
  pdl_prologue();
  if (lookahead_skip(d)) return;                                  // uniform over the whole cluster
  // look-ahead launches: the accumulators of the next linearisation are cleared here instead of by a kernel of their own
  for (size_t i = (size_t)blockIdx.x * SELC_THREADS + threadIdx.x; i < zero_n; i += (size_t)SELC_CTAS * SELC_THREADS) zero_ptr[i] = 0.0;
  namespace cg = cooperative_groups;
  // The first digit is the exponent: a handful of hot bins.  Same-address shared-memory atomics serialise, so every
  // CTA keeps SELC_COPIES replicas of the histogram (replica = lane & 7, interleaved so that replicas of one bin sit in
  // different banks) and folds them into `red` before the cluster-wide sum.
  extern __shared__ __align__(16) unsigned sel_smem[];
  unsigned* copies = sel_smem;                                   // [SEL_BINS][SELC_COPIES]
  unsigned* redh = sel_smem + SELC_COPIES * SEL_BINS;            // [2][SEL_BINS]
  __shared__ unsigned wsum[32];
  __shared__ unsigned long long s_prefix;
  __shared__ unsigned s_rank, s_count;
  __shared__ int s_prefix_shift;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned crank = cluster.block_rank();
  BaCtrl* ctrl = d.ctrl;
  // which_in -2: the trial state of this launch's candidate (speculative sigma of the state that becomes current if the
  // candidate is accepted; the control kernel has not run yet, so ctrl->cur is still the linearisation point)
  const int which = which_in == -2 ? trial_buffer(d, ctrl->cur) : (which_in < 0 ? ctrl->cur : which_in);
  const double* __restrict__ v = d.chi2[which];
  const int n = d.n_meas;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int gtid = (int)crank * SELC_THREADS + tid;
  const unsigned rep = lane & (SELC_COPIES - 1);
  unsigned long long key[SELC_K];
#pragma unroll
  for (int k = 0; k < SELC_K; k++) {
    const int i = gtid + k * (SELC_CTAS * SELC_THREADS);
    key[k] = (i < n) ? (unsigned long long)__double_as_longlong(fabs(v[i])) : ~0ull;   // ~0: never matches (sign bit)
  }
  unsigned long long prefix = 0;
  unsigned rank = (unsigned)(n / 2);
  bool done = false;
  // ---- bracket first ------------------------------------------------------------------------------------------------------
  // The median moves little from one LM state to the next, so the one of the current state (ctrl->med_hint) brackets the
  // wanted one: every CTA counts its keys below [0.9, 1.1) x hint and lists the keys inside (shared memory); after ONE
  // cluster barrier CTA 0 pulls the few thousand listed keys into registers (distributed shared memory) and finishes the
  // exact selection among them alone, on key - lo digits, with block barriers only.  If the wanted rank is not inside the
  // bracket (or the lists overflow) nothing is lost: the keys are still in registers and the full passes below run.
  bool fast_done = false;
  unsigned long long fast_bits = 0;
  {
    const double hint = ctrl->med_hint;
    if (mode != 2 && hint > 0.0 && hint < 1e300) {                   // uniform over the cluster (mode 2 selects over other data)
      __shared__ unsigned s_cnt, s_below, s_fcount;
      __shared__ unsigned long long s_fkey;
      unsigned long long* list = reinterpret_cast<unsigned long long*>(copies);     // SELC_LIST keys (the replica area is idle now)
      constexpr unsigned SELC_LIST = SELC_COPIES * SEL_BINS / 2;      // 8192
      const unsigned long long lo_b = (unsigned long long)__double_as_longlong(hint * 0.9), hi_b = (unsigned long long)__double_as_longlong(hint * 1.1);
      if (tid == 0) { s_cnt = 0; s_below = 0; }
      __syncthreads();
      unsigned below = 0, inb = 0;
#pragma unroll
      for (int k = 0; k < SELC_K; k++) { below += key[k] < lo_b; inb |= (unsigned)(key[k] >= lo_b && key[k] < hi_b) << k; }
      // warp-aggregated append: one shared-memory atomic per warp
      const unsigned mine = __popc(inb);
      unsigned incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
      const unsigned wtot = __shfl_sync(0xffffffffu, incl, 31);
      unsigned wbase = 0;
      if (lane == 31 && wtot) wbase = atomicAdd(&s_cnt, wtot);
      wbase = __shfl_sync(0xffffffffu, wbase, 31);
      unsigned at = wbase + incl - mine;
#pragma unroll
      for (int k = 0; k < SELC_K; k++)
        if (inb & (1u << k)) { if (at < SELC_LIST) list[at] = key[k] - lo_b; at++; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
      if (lane == 0 && below) atomicAdd(&s_below, below);
      cluster.sync();
      // totals (every CTA reads the eight counter pairs: the verdict is uniform)
      unsigned cnt_r[SELC_CTAS], tot = 0, bel = 0;
      bool overflow = false;
#pragma unroll
      for (int r = 0; r < SELC_CTAS; r++) {
        cnt_r[r] = *cluster.map_shared_rank(&s_cnt, r);
        bel += *cluster.map_shared_rank(&s_below, r);
        overflow |= cnt_r[r] > SELC_LIST;
        tot += cnt_r[r];
      }
      const unsigned want = (unsigned)(n / 2);
      const bool ok = !overflow && tot <= SELC_LIST && want >= bel && want - bel < tot;
      if (ok) {
        unsigned long long rel[SELC_LIST / SELC_THREADS];             // 8 listed keys per thread of CTA 0
        if (crank == 0) {
#pragma unroll
          for (int j = 0; j < (int)(SELC_LIST / SELC_THREADS); j++) {
            unsigned e = (unsigned)tid + (unsigned)j * SELC_THREADS;
            rel[j] = ~0ull;
            if (e < tot) {
              int r = 0;
              while (e >= cnt_r[r]) { e -= cnt_r[r]; r++; }
              rel[j] = cluster.map_shared_rank(list, r)[e];
            }
          }
        }
        cluster.sync();                                             // the lists have been read: the other CTAs may leave
        if (crank != 0) return;
        // exact selection of rank (want - bel) among the listed keys, MSB first on the bits of (hi - lo)
        const unsigned long long R = hi_b - lo_b;
        const int nb = 64 - __clzll((long long)R);
        unsigned long long fpre = 0;
        unsigned frank = want - bel;
        int consumed = 0;
        bool single = false;
        unsigned* fh = redh;
        while (consumed < nb && !single) {
          const int take = min(SEL_BITS, nb - consumed), shift = nb - consumed - take;
          *reinterpret_cast<uint2*>(fh + 2 * tid) = make_uint2(0u, 0u);
          __syncthreads();
#pragma unroll
          for (int j = 0; j < (int)(SELC_LIST / SELC_THREADS); j++)
            if (rel[j] != ~0ull && (rel[j] >> (shift + take)) == fpre) atomicAdd(&fh[(unsigned)(rel[j] >> shift) & ((1u << take) - 1u)], 1u);
          __syncthreads();
          const uint2 cc = *reinterpret_cast<const uint2*>(fh + 2 * tid);
          const unsigned tsum = cc.x + cc.y;
          unsigned inc2 = tsum;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, inc2, o); if (lane >= o) inc2 += t; }
          if (lane == 31) wsum[wid] = inc2;
          __syncthreads();
          unsigned wv = wsum[lane], winc = wv;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
          const unsigned woff = __shfl_sync(0xffffffffu, winc - wv, wid);
          const unsigned excl = woff + inc2 - tsum;
          if (frank >= excl && frank < excl + tsum) {
            const bool first = frank < excl + cc.x;
            s_fkey = (fpre << take) | (unsigned long long)(first ? 2 * tid : 2 * tid + 1);
            s_rank = frank - (first ? excl : excl + cc.x);
            s_fcount = first ? cc.x : cc.y;
          }
          __syncthreads();
          fpre = s_fkey; frank = s_rank;
          consumed += take;
          single = s_fcount == 1u;
          __syncthreads();
        }
        // fpre = the leading `consumed` bits of the wanted key - lo; all bits if the digits ran out, else exactly one listed key has them
        if (consumed < nb) {
#pragma unroll
          for (int j = 0; j < (int)(SELC_LIST / SELC_THREADS); j++)
            if (rel[j] != ~0ull && (rel[j] >> (nb - consumed)) == fpre) s_fkey = rel[j];
          __syncthreads();
          fpre = s_fkey;
        }
        fast_done = true;
        fast_bits = fpre + lo_b;
      }
    }
  }
  for (int pass = 0; pass < SEL_PASSES && !done && !fast_done; pass++) {
    unsigned* h = redh + (pass & 1) * SEL_BINS;
    const int shift = 63 - SEL_BITS * (pass + 1);       // 52, 41, 30, 19, 8, -3
    if (pass == 0) {
      // every key takes part and the digit is the exponent (hot bins): replicated histogram, folded afterwards
      for (int i = tid; i < SELC_COPIES * SEL_BINS; i += SELC_THREADS) copies[i] = 0;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < SELC_K; k++)
        if (key[k] != ~0ull) atomicAdd(&copies[(unsigned)(key[k] >> 52) * SELC_COPIES + rep], 1u);
      __syncthreads();
      const uint4* cp = reinterpret_cast<const uint4*>(copies + (size_t)2 * tid * SELC_COPIES);
      const uint4 a0 = cp[0], a1 = cp[1], b0 = cp[2], b1 = cp[3];
      uint2 f;
      f.x = a0.x + a0.y + a0.z + a0.w + a1.x + a1.y + a1.z + a1.w;
      f.y = b0.x + b0.y + b0.z + b0.w + b1.x + b1.y + b1.z + b1.w;
      *reinterpret_cast<uint2*>(h + 2 * tid) = f;
    } else {
      // later digits are mantissa bits (well spread) and only the keys under the current prefix take part
      *reinterpret_cast<uint2*>(h + 2 * tid) = make_uint2(0u, 0u);
      __syncthreads();
#pragma unroll
      for (int k = 0; k < SELC_K; k++) {
        const unsigned long long hi = (shift >= 0) ? (key[k] >> (shift + SEL_BITS)) : (key[k] >> (SEL_BITS + shift));
        if (hi == prefix) {
          const unsigned dig = (shift >= 0) ? ((unsigned)(key[k] >> shift) & (SEL_BINS - 1)) : ((unsigned)(key[k] << (-shift)) & (SEL_BINS - 1));
          atomicAdd(&h[dig], 1u);
        }
      }
    }
    cluster.sync();
    // global counts of bins 2*tid, 2*tid+1
    unsigned c0 = 0, c1 = 0;
#pragma unroll
    for (int r = 0; r < SELC_CTAS; r++) {
      const unsigned* rh = cluster.map_shared_rank(h, r);
      const uint2 t = *reinterpret_cast<const uint2*>(rh + 2 * tid);
      c0 += t.x; c1 += t.y;
    }
    const unsigned tsum = c0 + c1;
    unsigned incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    // exclusive offset of this warp: scan of the 32 warp totals by every warp
    unsigned wv = wsum[lane], winc = wv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
    const unsigned woff = __shfl_sync(0xffffffffu, winc - wv, wid);
    const unsigned excl = woff + incl - tsum;
    if (rank >= excl && rank < excl + tsum) {
      const bool first = rank < excl + c0;
      const unsigned bin = first ? 2 * tid : 2 * tid + 1;
      const unsigned acc = first ? excl : excl + c0;
      s_prefix = (shift >= 0) ? ((prefix << SEL_BITS) | bin) : ((prefix << (SEL_BITS + shift)) | (bin >> (-shift)));
      s_rank = rank - acc;
      s_count = first ? c0 : c1;
      s_prefix_shift = shift;                              // key >> shift == prefix for the keys under it (shift >= 0 here when used)
    }
    __syncthreads();
    prefix = s_prefix; rank = s_rank;
    done = (s_count == 1u) && (pass + 1 < SEL_PASSES);      // a single key is left under the prefix: it is the answer
  }
  // `prefix` holds the leading bits of the wanted key.  If the passes ran to the end it is the whole key; otherwise
  // exactly one key in the cluster starts with it, and the thread that owns it finishes the job.
  unsigned long long med_bits = fast_done ? fast_bits : prefix;
  bool owner = (crank == 0 && tid == 0);
  if (done && !fast_done) {
    owner = false;
#pragma unroll
    for (int k = 0; k < SELC_K; k++) {
      // number of prefix bits after p passes: 11p + 1 leading (sign) bit => compare key >> (63 - 11p) ... done by length
      if (key[k] != ~0ull && (key[k] >> s_prefix_shift) == prefix) { owner = true; med_bits = key[k]; }
    }
  }
  if (!fast_done) cluster.sync();                       // nobody leaves while its histogram may still be read
  if (owner) {
    const double med = __longlong_as_double((long long)med_bits);
    if (mode == 0) ctrl->med_hint = med;
    else if (mode == 3) d.spec_med[d.cand] = med;
    const size_t denom = (size_t)n * 2 - 6;                     // size_t arithmetic as in the reference
    double s = 1.4826 * (1 + 5.0 / (double)denom) * sqrt(med);
    if (mode == 2) ctrl->median_out = med;                      // plain upper median (src/ChainBundle.cc:1434)
    else if (mode == 3) { s = 1.345 * s; d.spec_sigma[d.cand] = s * s; }
    else if (mode == 0) {
      { const double st = 4.6851 * s; double t = st * st; if (t < ctrl->min_sigma_sq) t = ctrl->min_sigma_sq; ctrl->tukey_sigma_sq = t; }   // same median: the Tukey sigma^2 comes along
      s = 1.345 * s;
      ctrl->sigma_sq_raw = s * s;
      ctrl->sigma_sq_lim = ctrl->sigma_sq_raw < ctrl->min_sigma_sq ? ctrl->min_sigma_sq : ctrl->sigma_sq_raw;
      ctrl->sigma_lim = sqrt(ctrl->sigma_sq_lim);
    } else {
      s = 4.6851 * s;
      double t = s * s;
      if (t < ctrl->min_sigma_sq) t = ctrl->min_sigma_sq;
      ctrl->tukey_sigma_sq = t;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k_select_grid: the same exact radix select for SELC_CAP < n <= SELG_CAP values in ONE cooperative launch: one CTA per
// SM, every thread keeps its <= 16 keys in registers for all six digits; per digit the CTAs add their shared-memory
// histogram into the global one, meet at a grid barrier, and every CTA narrows (prefix, rank) redundantly.  Replaces six
// dependent launches of k_sel_pass (the 1000 KF map: 86 -> ~35 us).
// ---------------------------------------------------------------------------------------------
constexpr int SELG_THREADS = 1024, SELG_K = 16;
__global__ void __launch_bounds__(SELG_THREADS, 1) k_select_grid(BaDev d, int which_in, int mode, double* zero_ptr, size_t zero_n)
{
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  if (lookahead_skip(d)) return;                                  // uniform over the grid
  for (size_t i = (size_t)blockIdx.x * SELG_THREADS + threadIdx.x; i < zero_n; i += (size_t)gridDim.x * SELG_THREADS) zero_ptr[i] = 0.0;
  __shared__ unsigned hist[SEL_BINS];
  __shared__ unsigned wsum[32];
  __shared__ unsigned long long s_prefix;
  __shared__ unsigned s_rank, s_count;
  __shared__ int s_prefix_shift;
  BaCtrl* ctrl = d.ctrl;
  const int which = which_in < 0 ? ctrl->cur : which_in;
  const double* __restrict__ v = d.chi2[which];
  const int n = d.n_meas;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int gtid = blockIdx.x * SELG_THREADS + tid, gsz = gridDim.x * SELG_THREADS;
  unsigned long long key[SELG_K];
#pragma unroll
  for (int k = 0; k < SELG_K; k++) {
    const int i = gtid + k * gsz;
    key[k] = (i < n) ? (unsigned long long)__double_as_longlong(fabs(v[i])) : ~0ull;
  }
  unsigned long long prefix = 0;
  unsigned rank = (unsigned)(n / 2);
  bool done = false;
  for (int pass = 0; pass < SEL_PASSES && !done; pass++) {
    unsigned* gh = d.sel_hist + pass * SEL_BINS;
    const int shift = 63 - SEL_BITS * (pass + 1);
    for (int i = tid; i < SEL_BINS; i += SELG_THREADS) hist[i] = 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SELG_K; k++) {
      const unsigned long long hi = (shift >= 0) ? (key[k] >> (shift + SEL_BITS)) : (key[k] >> (SEL_BITS + shift));
      const bool match = (key[k] != ~0ull) && (pass == 0 || hi == prefix);
      const unsigned dig = (shift >= 0) ? ((unsigned)(key[k] >> shift) & (SEL_BINS - 1)) : ((unsigned)(key[k] << (-shift)) & (SEL_BINS - 1));
      const unsigned act = __ballot_sync(0xffffffffu, match);
      if (match) {
        const unsigned peers = __match_any_sync(act, dig);
        if (lane == __ffs(peers) - 1) atomicAdd(&hist[dig], (unsigned)__popc(peers));
      }
    }
    __syncthreads();
    for (int i = tid; i < SEL_BINS; i += SELG_THREADS) if (hist[i]) atomicAdd(&gh[i], hist[i]);
    grid.sync();
    // global counts of bins 2*tid, 2*tid+1
    const uint2 cc = __ldcg(reinterpret_cast<const uint2*>(gh) + tid);
    const unsigned c0 = cc.x, c1 = cc.y, tsum = c0 + c1;
    unsigned incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    unsigned wv = wsum[lane], winc = wv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
    const unsigned woff = __shfl_sync(0xffffffffu, winc - wv, wid);
    const unsigned excl = woff + incl - tsum;
    if (rank >= excl && rank < excl + tsum) {
      const bool first = rank < excl + c0;
      const unsigned bin = first ? 2 * tid : 2 * tid + 1;
      const unsigned acc = first ? excl : excl + c0;
      s_prefix = (shift >= 0) ? ((prefix << SEL_BITS) | bin) : ((prefix << (SEL_BITS + shift)) | (bin >> (-shift)));
      s_rank = rank - acc;
      s_count = first ? c0 : c1;
      s_prefix_shift = shift;
    }
    __syncthreads();
    prefix = s_prefix; rank = s_rank;
    done = (s_count == 1u) && (pass + 1 < SEL_PASSES);
  }
  unsigned long long med_bits = prefix;
  bool owner = (blockIdx.x == 0 && tid == 0);
  if (done) {
    owner = false;
#pragma unroll
    for (int k = 0; k < SELG_K; k++)
      if (key[k] != ~0ull && (key[k] >> s_prefix_shift) == prefix) { owner = true; med_bits = key[k]; }
  }
  grid.sync();                                          // every CTA has read the last histogram: re-arm them for the next call
  for (int i = gtid; i < SEL_PASSES * SEL_BINS; i += gsz) d.sel_hist[i] = 0;
  if (owner) {
    const double med = __longlong_as_double((long long)med_bits);
    const size_t denom = (size_t)n * 2 - 6;
    double s = 1.4826 * (1 + 5.0 / (double)denom) * sqrt(med);
    if (mode == 2) ctrl->median_out = med;
    else if (mode == 0) {
      { const double st = 4.6851 * s; double t = st * st; if (t < ctrl->min_sigma_sq) t = ctrl->min_sigma_sq; ctrl->tukey_sigma_sq = t; }   // same median: the Tukey sigma^2 comes along
      s = 1.345 * s;
      ctrl->sigma_sq_raw = s * s;
      ctrl->sigma_sq_lim = ctrl->sigma_sq_raw < ctrl->min_sigma_sq ? ctrl->min_sigma_sq : ctrl->sigma_sq_raw;
      ctrl->sigma_lim = sqrt(ctrl->sigma_sq_lim);
    } else {
      s = 4.6851 * s;
      double t = s * s;
      if (t < ctrl->min_sigma_sq) t = ctrl->min_sigma_sq;
      ctrl->tukey_sigma_sq = t;
    }
  }
}

// Tukey outlier flags (src/ChainBundle.cc:1385-1398)
// Tukey outliers (src/ChainBundle.cc:1384-1399) as a compact list: outlier_flags[0] = count (cleared by the caller),
// outlier_flags[1..] = the ORIGINAL indices of the flagged measurements, in no particular order (the host sorts them)
__global__ void k_tukey_flags(BaDev d)
{
  const double ts = d.ctrl->tukey_sigma_sq;
  const double* __restrict__ v = d.chi2[d.ctrl->cur];
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < d.n_meas; m += gridDim.x * blockDim.x) {
    const double a = fabs(v[m]);
    const double sq = a > ts ? 0.0 : 1.0 - (a / ts);
    if (sq * sq == 0.0) d.outlier_flags[1 + atomicAdd(&d.outlier_flags[0], 1)] = d.meas_a[m].w;
  }
}

// sum of the robustified chi2 of a state whose errors are already in d.chi2[which] (this rank's measurement range): the
// "BEFORE" / "AFTER" figures of a Compute (src/ChainBundle.cc:1317-1345) need the errors for the sigma selection first and
// the robust sum afterwards -- a pass over the stored values instead of a second reprojection of every measurement
__global__ void __launch_bounds__(256) k_robust_sum(BaDev d, int which_in)
{
  __shared__ double red[32];
  const BaCtrl* ctrl = d.ctrl;
  const double* __restrict__ v = d.chi2[which_in < 0 ? ctrl->cur : which_in];
  double acc = 0;
  for (int m = d.m_lo + blockIdx.x * blockDim.x + threadIdx.x; m < d.m_hi; m += gridDim.x * blockDim.x) {
    double rho0, rho1;
    robustify(ctrl, v[m], rho0, rho1);
    acc += rho0;
  }
  const double a = block_sum(acc, red);
  if (threadIdx.x == 0) {
    d.part[PART_TMP_CHI * MAX_PARTIALS + blockIdx.x] = a;
    d.part[PART_SCALE * MAX_PARTIALS + blockIdx.x] = 0.0;
    d.part[PART_SUMSQ * MAX_PARTIALS + blockIdx.x] = 0.0;
  }
}

// g2o computeLambdaInit: tau * max |H_jj| over poses (H0 diagonal) and points (V diagonal)
__global__ void __launch_bounds__(1024) k_lambda_init(BaDev d)
{
  __shared__ double red[32];
  double mx = 0;
  for (int i = threadIdx.x; i < d.nc; i += blockDim.x) mx = fmax(mx, fabs(d.H0[(size_t)i * d.nc + i]));
  for (int p = d.p_lo + threadIdx.x; p < d.p_hi; p += blockDim.x) {
    if (d.pt_var[p] < 0) continue;
    const double* V = d.V + 6 * (size_t)p;
    mx = fmax(mx, fmax(fabs(V[0]), fmax(fabs(V[3]), fabs(V[5]))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); i++) mx = fmax(mx, red[i]);
    d.part[PART_MAXDIAG * MAX_PARTIALS] = mx;
  }
}
__global__ void k_lambda_apply(BaDev d)
{
  BaCtrl* c = d.ctrl;
  if (c->need_lambda_init) {
    c->max_diag = d.part[PART_MAXDIAG * MAX_PARTIALS];
    c->lambda = c->user_lambda > 0 ? c->user_lambda : 1e-5 * c->max_diag;
    c->ni = 2;
    c->need_lambda_init = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// k_lm_control: [3P] OptimizationAlgorithmLevenberg::solve trial bookkeeping + post-iteration actions.
// n_cand speculative candidates were evaluated concurrently (candidate c used lambda after c rejections); they are
// consumed strictly in g2o's order, so the accepted step and the lambda/ni sequence are those of the sequential
// algorithm.  red_in != nullptr: sums already reduced over the ranks (multi-GPU); else reduce partials here.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_lm_control(BaDev d, CandParts parts, int n_cand, int n_part_lin, int n_part_bs,
                                                   const double* red_in, int first_trial)
{
  pdl_prologue();
  __shared__ double s_sum[MAX_CAND][4];
  __shared__ double s_abort;
  if (threadIdx.x == 0) s_abort = 0.0;
  __syncthreads();
  if (red_in) {
    // red_in = { cur_chi, (tmp_chi, scale, sumsq) per candidate }
    if (threadIdx.x == 0) {
      s_sum[0][0] = red_in[0];
      for (int cnd = 0; cnd < n_cand; cnd++) { s_sum[cnd][1] = red_in[1 + 3 * cnd]; s_sum[cnd][2] = red_in[2 + 3 * cnd]; s_sum[cnd][3] = red_in[3 + 3 * cnd]; }
      s_abort = red_in[1 + 3 * n_cand];
    }
  } else {
    // every partial sum is loaded before the first reduction (one round of independent loads instead of ten
    // dependent ones), then the 1 + 3 n_cand values are reduced together
    double v[1 + 3 * MAX_CAND];
#pragma unroll
    for (int k = 0; k < 1 + 3 * MAX_CAND; k++) v[k] = 0.0;
    for (int i = threadIdx.x; i < n_part_lin; i += blockDim.x) v[0] += d.part[PART_CUR_CHI * MAX_PARTIALS + i];
#pragma unroll
    for (int cnd = 0; cnd < MAX_CAND; cnd++) {
      if (cnd < n_cand) {
        const double* part = parts.p[cnd];
        for (int i = threadIdx.x; i < n_part_bs; i += blockDim.x) {
          v[1 + 3 * cnd] += part[PART_TMP_CHI * MAX_PARTIALS + i];
          v[2 + 3 * cnd] += part[PART_SCALE * MAX_PARTIALS + i];
          v[3 + 3 * cnd] += part[PART_SUMSQ * MAX_PARTIALS + i];
        }
      }
    }
    __shared__ double s_red[1 + 3 * MAX_CAND][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 1 + 3 * MAX_CAND; k++) {
      const double w = warp_sum(v[k]);
      if (lane == 0) s_red[k][wid] = w;
    }
    __syncthreads();
    if (threadIdx.x < 1 + 3 * MAX_CAND) {
      double t = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s_red[threadIdx.x][w];
      if (threadIdx.x == 0) s_sum[0][0] = t;
      else s_sum[(threadIdx.x - 1) / 3][1 + (threadIdx.x - 1) % 3] = t;
    }
  }
  // The bookkeeping below is a chain of dependent read-modify-writes of the control block: run it on a shared-memory copy
  // (one coalesced load, one coalesced store) instead of ~40 serial round trips to global memory.
  __shared__ BaCtrl s_ctrl;
  {
    const int* src = reinterpret_cast<const int*>(d.ctrl);
    int* dst = reinterpret_cast<int*>(&s_ctrl);
    for (int i = threadIdx.x; i < (int)(sizeof(BaCtrl) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
  BaCtrl* c = &s_ctrl;
  if (red_in) c->abort_agreed = s_abort;
  if (first_trial) { c->current_chi = s_sum[0][0]; c->lin_chi = s_sum[0][0]; }
  const int cur0 = c->cur;
  bool again = true;
  double temp_raw = 0, sumsq = 0;
  int used = 0;
  for (int cnd = 0; cnd < n_cand && again; cnd++) {
    used++;
    temp_raw = s_sum[cnd][1];
    double temp_chi = temp_raw;
    if (!c->solve_ok[cnd]) temp_chi = 1.7976931348623157e308;
    c->temp_chi = temp_raw;
    double rho = c->current_chi - temp_chi;
    double scale = c->scale[cnd] + s_sum[cnd][2];
    scale += 1e-3;
    rho /= scale;
    sumsq = c->sumsq[cnd] + s_sum[cnd][3];
    if (rho > 0 && isfinite(temp_chi)) {
      const double t = 2 * rho - 1;
      double alpha = 1. - t * t * t;
      alpha = fmin(alpha, 2. / 3.);
      const double sf = fmax(1. / 3., alpha);
      c->lambda *= sf;          // c->lambda was already advanced by the rejections of the earlier candidates
      c->ni = 2;
      c->current_chi = temp_chi;
      c->cur = (cur0 + 1 + cnd) % N_STATE;
      c->accepted = 1;
      c->acc_cand = cnd;
    } else {
      c->lambda *= c->ni;
      c->ni *= 2;
      c->accepted = 0;
    }
    c->rho = rho;
    c->qmax++;
    again = (rho < 0) && (c->qmax < c->max_trials);
    if (!again && (c->qmax == c->max_trials || rho == 0)) c->terminate = 1;
  }
  c->cand_used = used;
  for (int q = 0; q < MAX_CAND; q++) c->solve_ok[q] = 1;
  c->stop_trials = again ? 0 : 1;
  if (!again) {
    c->iter++;
    // CheckConvergedUpdateMagAction (src/ChainBundle.cc:1009-1047)
    const double rms = sqrt(sumsq / c->dim);
    if (rms < c->rms_limit) c->conv_mag = 1;
    // CheckConvergedResidualAction (:1091-1118): robust chi2 of the errors currently held by the edges
    const double curchi = temp_raw;
    const double pct = (c->last_chi2 - curchi) / c->last_chi2;
    if (pct >= 0 && pct <= c->pct_limit) c->conv_res = 1;
    else if (curchi == 0) c->conv_res = 1;
    c->last_chi2 = curchi;
    c->total_trials += c->qmax;
    c->qmax = 0;
  }
  }
  __syncthreads();
  {
    int* dst = reinterpret_cast<int*>(d.ctrl);
    const int* src = reinterpret_cast<const int*>(&s_ctrl);
    for (int i = threadIdx.x; i < (int)(sizeof(BaCtrl) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
  }
}

// Multi-GPU exchange buffers: the reduced camera system is symmetric and only its upper block triangle is ever written
// (element (r, c), c <= r, lives at [c * n + r]), so the all-reduce carries the packed triangle plus the vector that
// follows the matrix in memory (gc + the scalar sums, or rm) -- half the bytes of the square.
//   packed[c * n - c (c - 1) / 2 + (r - c)] = full[c * n + r],   tail: packed[n (n + 1) / 2 + i] = full[n * n + i]
struct TriPackArgs { double* full[MAX_CAND]; };
__global__ void __launch_bounds__(256) k_tri_pack(TriPackArgs a, double* __restrict__ packed_all, int n, int tail, int unpack)
{
  const int c = blockIdx.y;
  const size_t ntri = (size_t)n * (n + 1) / 2;
  double* full = a.full[blockIdx.z];
  double* packed = packed_all + (ntri + tail) * blockIdx.z;
  if (c < n) {
    const size_t po = (size_t)c * n - (size_t)c * (c - 1) / 2;
    for (int r = c + blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
      if (unpack) full[(size_t)c * n + r] = packed[po + (r - c)];
      else packed[po + (r - c)] = full[(size_t)c * n + r];
    }
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < tail; i += gridDim.x * blockDim.x) {
      if (unpack) full[(size_t)n * n + i] = packed[ntri + i];
      else packed[ntri + i] = full[(size_t)n * n + i];
    }
  }
}
// `count` matrices [full | tail] <-> consecutive packed records of n (n + 1) / 2 + tail doubles, one launch
void launch_tri_pack(double* const* full, int count, double* packed, int n, int tail, bool unpack, cudaStream_t s)
{
  if (n <= 0 || count <= 0) return;
  TriPackArgs a;
  for (int q = 0; q < MAX_CAND; q++) a.full[q] = full[q < count ? q : 0];
  k_tri_pack<<<dim3((n + 255) / 256, n + 1, count), 256, 0, s>>>(a, packed, n, tail, unpack ? 1 : 0);
}

// sums the per-block partials into out[0..3] = {cur_chi, tmp_chi, scale, sumsq} (multi-GPU path)
__global__ void __launch_bounds__(256) k_reduce_partials(BaDev d, int n_part_lin, int n_part_bs, double* out, const volatile double* host_word, int word_slot)
{
  // (multi-GPU) this rank's view of the caller's abort flag rides along: read from pinned host memory, summed with the trial sums
  if (host_word && threadIdx.x == 0) out[word_slot] = *host_word;
  __shared__ double red[32];
  double a = 0, b = 0, c = 0, e = 0;
  for (int i = threadIdx.x; i < n_part_lin; i += blockDim.x) a += d.part[PART_CUR_CHI * MAX_PARTIALS + i];
  for (int i = threadIdx.x; i < n_part_bs; i += blockDim.x) {
    b += d.part[PART_TMP_CHI * MAX_PARTIALS + i];
    c += d.part[PART_SCALE * MAX_PARTIALS + i];
    e += d.part[PART_SUMSQ * MAX_PARTIALS + i];
  }
  a = block_sum(a, red); b = block_sum(b, red); c = block_sum(c, red); e = block_sum(e, red);
  if (threadIdx.x == 0) { if (n_part_lin > 0) out[0] = a; if (n_part_bs > 0) { out[1] = b; out[2] = c; out[3] = e; } }
}

// Debug: explicit Jacobians per measurement in the caller's order: Jobs(12) Jsrc(12) Jpt(6)
__global__ void k_debug_jacobians(BaDev d, double* out)
{
  const int cur = d.ctrl->cur;
  const double* pose = d.pose[cur];
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < d.n_meas; m += gridDim.x * blockDim.x) {
    const int4 ma = d.meas_a[m], mbi = d.meas_b[m];
    const int p = mbi.w;
    const int4 pi = d.pt_info[p];
    const double* pp = d.pt[cur] + 3 * (size_t)p;
    const double prel[3] = { pp[0], pp[1], pp[2] };
    PtCtx c;
    load_pt_ctx(d, pose, pi, prel, c);
    MeasGeom g;
    meas_geometry<true>(d.cams, pose, c, ma, d.meas_xy[m], g);
    double* o = out + 30 * (size_t)ma.w;
    for (int i = 0; i < 30; i++) o[i] = 0;
    if (mbi.x >= 0) pose_jac(g.A, g.q, -1.0, o);
    if (mbi.z) pose_jac(g.A2, c.qs, 1.0, o + 12);
    if (d.pt_var[p] >= 0) {
      double M[9];
      point_tangent(prel, M);
      for (int r = 0; r < 2; r++)
        for (int k = 0; k < 3; k++) o[24 + r * 3 + k] = -(g.A3[r * 3] * M[k] + g.A3[r * 3 + 1] * M[3 + k] + g.A3[r * 3 + 2] * M[6 + k]);
    }
  }
}

// gathers the full update vector (movable poses then movable points) for mcp_ba_lm_step
__global__ void k_gather_delta(BaDev d, double* out)
{
  const double lambda = trial_lambda(d);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  for (int i = tid; i < d.nc; i += nt) out[i] = d.dc[i];
  for (int p = tid; p < d.n_pt; p += nt) {
    const int pv = d.pt_var[p];
    if (pv < 0) continue;
    const int s0 = d.pt_slot_off[p], K = d.pt_slot_off[p + 1] - s0;
    double t[3] = { 0, 0, 0 };
    for (int a = 0; a < K; a++)
      for (int r = 0; r < 6; r++)
        for (int k = 0; k < 3; k++) t[k] += d.W[(size_t)(s0 + a) * 18 + r * 3 + k] * d.dc[6 * d.slot_var[s0 + a] + r];
    double V6[6], gp[3], Vi[9];
    for (int i = 0; i < 6; i++) V6[i] = d.V[6 * (size_t)p + i];
    for (int i = 0; i < 3; i++) gp[i] = d.gp[3 * (size_t)p + i];
    inv3_sym(V6, lambda, Vi);
    const double rr[3] = { gp[0] - t[0], gp[1] - t[1], gp[2] - t[2] };
    double dp[3];
    m3_vec(Vi, rr, dp);
    for (int k = 0; k < 3; k++) out[d.nc + 3 * pv + k] = dp[k];
  }
}

// ---------------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------------
static int per_point_grid(const BaDev& d, int warps)
{
  const int units = (d.p_hi - d.p_lo + PPW - 1) / PPW;     // one warp per PPW points
  int g = (units + warps - 1) / warps;
  if (g < 1) g = 1;
  if (g > MAX_PARTIALS) g = MAX_PARTIALS;
  return g;
}

// Register-allocation variants of k_linearize (the kernel is latency bound; fewer registers = more resident warps).
// MCP_BA_LIN_VARIANT: 1 (default) = 128 threads x 3 blocks/SM (168 registers; used when three blocks' shared memory fit,
// else variant 0), 0 = 256 threads, 1 block/SM (214 registers), 2 = 256 threads x 2 blocks/SM (128 registers, spills).
static int lin_variant()
{
  static const int v = [] { const char* e = getenv("MCP_BA_LIN_VARIANT"); return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1; }();
  return v;
}
// zeroes the linearisation accumulators [H0 | gc | red] (a memset that honours the look-ahead predicate)
__global__ void k_zero_acc(BaDev d, double* acc, size_t n, int pick_sigma)
{
  pdl_prologue();
  if (lookahead_skip(d)) return;
  if (pick_sigma && blockIdx.x == 0 && threadIdx.x == 0) {
    // the outer iteration ended by accepting candidate acc_cand: its trial state is the new linearisation point and the
    // Huber sigma^2 of that state was computed next to the trial (k_select_cluster mode 3).  RobustKernelData::RecomputeNow.
    BaCtrl* c = d.ctrl;
    if (c->accepted) {
      const double raw = d.spec_sigma[c->acc_cand];
      c->med_hint = d.spec_med[c->acc_cand];
      c->sigma_sq_raw = raw;
      c->sigma_sq_lim = raw < c->min_sigma_sq ? c->min_sigma_sq : raw;
      c->sigma_lim = sqrt(c->sigma_sq_lim);
    }
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) acc[i] = 0.0;
}
void launch_zero_acc(const BaDev& d, double* acc, size_t n, cudaStream_t s, int pick_sigma) { launch_chain(k_zero_acc, dim3(148), dim3(512), 0, s, d, acc, n, pick_sigma); }
// mc != nullptr: the launch also forms the point records of the next trial round (lambda must already be valid on the device)
static void launch_pose_blocks(const BaDev& d, cudaStream_t s, const SchurMulti* mc, double* zero_ptr, size_t zero_n)
{
  if (d.n_pb_items <= 0 && !mc && !zero_n) return;
  int g = (d.n_pb_items + 3) / 4;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 0) g = 0;
  SchurMulti m;
  memset(&m, 0, sizeof(m));
  int extra = 0;
  if (mc) {
    m = *mc;
    extra = (d.p_hi - d.p_lo + 127) / 128;
    if (extra < 148) extra = 148;
    if (extra > 148 * 4) extra = 148 * 4;
  }
  if (g + extra == 0) extra = 148;                     // (only the clearing to do)
  launch_chain(k_pose_blocks, dim3(g + extra), dim3(128), 0, s, d, m, g, zero_ptr, zero_n);
}
int launch_linearize(const BaDev& d, int warps, size_t smem, cudaStream_t s, const SchurMulti* mc, double* zero_ptr, size_t zero_n)
{
  const size_t stage = sizeof(double) * d.stage_doubles;
  if (lin_variant() == 1 && warps >= 4 && smem / warps * 4 + stage <= 74 * 1024) {
    const int g = per_point_grid(d, 4);
    launch_chain(k_linearize<128, 3>, dim3(g), dim3(128), smem / warps * 4 + stage, s, d);
    launch_pose_blocks(d, s, mc, zero_ptr, zero_n);
    return g;
  }
  const int g = per_point_grid(d, warps);
  if (lin_variant() == 2 && warps == 8 && smem + stage <= 100 * 1024) launch_chain(k_linearize<256, 2>, dim3(g), dim3(256), smem + stage, s, d);
  else launch_chain(k_linearize<256, 1>, dim3(g), dim3(warps * 32), smem + stage, s, d);
  launch_pose_blocks(d, s, mc, zero_ptr, zero_n);
  return g;
}
int launch_backsub_eval(const BaDev& d, int apply, int which, double* err_out, cudaStream_t s)
{
  const int g = per_point_grid(d, 8);
  launch_chain(k_backsub_eval, dim3(g), dim3(256), sizeof(double) * d.stage_doubles, s, d, apply, which, err_out);
  return g;
}
int launch_select_sigma(const BaDev& d, int which, int mode, cudaStream_t s, double* zero_ptr, size_t zero_n)
{
  static const bool multi_launch = [] { const char* e = getenv("MCP_BA_SELECT_MULTI"); return e && e[0] == '1'; }();
  if (d.n_meas > 0 && d.n_meas <= SELC_CAP && !multi_launch) { launch_chain(k_select_cluster, dim3(SELC_CTAS), dim3(SELC_THREADS), SELC_SMEM, s, d, which, mode, zero_ptr, zero_n); return 1; }
  {
    // one cooperative launch, one CTA per SM, keys in registers
    static const int n_sms = [] { int dev = 0, v = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); return v > 0 ? v : 1; }();
    if (!multi_launch && (long long)d.n_meas <= (long long)n_sms * SELG_THREADS * SELG_K) {
      BaDev dd = d;
      void* args[] = { &dd, &which, &mode, &zero_ptr, &zero_n };
      if (cudaLaunchCooperativeKernel((void*)k_select_grid, dim3(n_sms), dim3(SELG_THREADS), args, 0, s) == cudaSuccess) return 1;
      (void)cudaGetLastError();                      // fall back to the multi-launch path
    }
  }
  if (zero_n) launch_zero_acc(d, zero_ptr, zero_n, s, 0);
  int grid = (d.n_meas + 2047) / 2048;
  if (grid < 1) grid = 1;
  if (grid > 148) grid = 148;
  for (int pass = 0; pass < SEL_PASSES; pass++) k_sel_pass<<<grid, 256, 0, s>>>(d, which, pass, mode);
  return SEL_PASSES;
}
// Huber sigma^2 of state buffer `which` (candidate d.cand's trial state) into d.spec_sigma[d.cand]; one-launch kernel only
bool select_spec_possible(const BaDev& d) { return d.n_meas > 0 && d.n_meas <= SELC_CAP; }
void launch_select_spec(const BaDev& d, int which, cudaStream_t s)
{
  launch_chain(k_select_cluster, dim3(SELC_CTAS), dim3(SELC_THREADS), SELC_SMEM, s, d, which, 3, (double*)nullptr, (size_t)0);
}
__global__ void __launch_bounds__(512) k_load_init(LoadInit li)
{
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < li.n_zero; r++) {
    const size_t n16 = li.zero_bytes[r] / 16, tail = (li.zero_bytes[r] - 16 * n16) / 4;
    uint4* p16 = reinterpret_cast<uint4*>(li.zero_ptr[r]);
    for (size_t i = gtid; i < n16; i += gsz) p16[i] = make_uint4(0u, 0u, 0u, 0u);
    unsigned* p4 = reinterpret_cast<unsigned*>(p16 + n16);
    if (gtid < tail) p4[gtid] = 0u;
  }
  for (size_t i = gtid; i < li.pose_doubles; i += gsz) {
    const double v = li.pose0[i];
#pragma unroll
    for (int k = 0; k < N_STATE; k++) li.pose[k][i] = v;
  }
  for (size_t i = gtid; i < li.pt_doubles; i += gsz) {
    const double v = li.pt0[i];
#pragma unroll
    for (int k = 0; k < N_STATE; k++) li.pt[k][i] = v;
  }
}
void launch_load_init(const LoadInit& li, cudaStream_t s) { k_load_init<<<148, 512, 0, s>>>(li); }
void launch_tukey_flags(const BaDev& d, cudaStream_t s) { k_tukey_flags<<<148, 256, 0, s>>>(d); }
int launch_robust_sum(const BaDev& d, int which, cudaStream_t s)
{
  const int g = 148 < MAX_PARTIALS ? 148 : MAX_PARTIALS;
  k_robust_sum<<<g, 256, 0, s>>>(d, which);
  return g;
}
void launch_lambda_init(const BaDev& d, cudaStream_t s) { k_lambda_init<<<1, 1024, 0, s>>>(d); }
void launch_lambda_apply(const BaDev& d, cudaStream_t s) { k_lambda_apply<<<1, 1, 0, s>>>(d); }
void launch_lm_control(const BaDev& d, const CandParts& parts, int n_cand, int n_lin, int n_bs, const double* red_in, int first_trial, cudaStream_t s)
{
  launch_chain(k_lm_control, dim3(1), dim3(256), 0, s, d, parts, n_cand, n_lin, n_bs, red_in, first_trial);
}
void launch_reduce_partials(const BaDev& d, int n_lin, int n_bs, double* out, cudaStream_t s, const double* host_word, int word_slot)
{
  k_reduce_partials<<<1, 256, 0, s>>>(d, n_lin, n_bs, out, host_word, word_slot);
}
void launch_debug_jacobians(const BaDev& d, double* out, cudaStream_t s) { k_debug_jacobians<<<148, 128, 0, s>>>(d, out); }
void launch_gather_delta(const BaDev& d, double* out, cudaStream_t s) { k_gather_delta<<<148, 128, 0, s>>>(d, out); }

int stage_doubles_for(int n_pose, int n_cam)
{
  const size_t n = (size_t)n_pose * 12 + (size_t)n_cam * (sizeof(DevCam) / 8);
  return n * sizeof(double) <= 40 * 1024 ? (int)n : 0;
}

int configure_kernels(int max_slots, int stage_doubles, int* warps_out, size_t* smem_out)
{
  // shared memory per warp: the W blocks of its PPW points, max_slots x 18 doubles each, plus their contexts
  const size_t per_warp = ((size_t)max_slots * 18 + CTXD) * sizeof(double) * PPW;
  int warps = 8;
  const size_t stage = sizeof(double) * (size_t)stage_doubles;
  while (warps > 1 && per_warp * warps + stage > 200 * 1024) warps >>= 1;
  const size_t smem = per_warp * warps;
  if (smem + stage > 200 * 1024) return -1;
  if (cudaFuncSetAttribute(k_select_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SELC_SMEM) != cudaSuccess) return -2;
  if (cudaFuncSetAttribute(k_linearize<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)) != cudaSuccess) return -2;
  if (cudaFuncSetAttribute(k_linearize<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(100 * 1024)) != cudaSuccess) return -2;
  if (cudaFuncSetAttribute(k_linearize<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(74 * 1024)) != cudaSuccess) return -2;
  *warps_out = warps;
  *smem_out = smem;
  return 0;
}

}  // namespace mcp
