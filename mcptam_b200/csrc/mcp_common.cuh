// mcp_common.cuh — shared device math for the BA and front-end kernels (sm_100a).
// fp64 throughout the BA path (north-star: "fp64 Jacobians").
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mcptam_b200.h"

namespace mcp {

void set_last_error(const char* fmt, ...);

#define MCP_CUDA_CHECK(expr)                                                                    \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      mcp::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MCP_ERR_CUDA;                                                                      \
    }                                                                                           \
  } while (0)

// Programmatic dependent launch (PDL): the kernels of the per-round chain are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so that the next kernel's launch and prologue overlap the tail of
// the current one.  Every such kernel starts with pdl_prologue(): it lets ITS dependents launch early and then waits
// until the kernel before it in the stream has completed and flushed its memory.  The wait is the first statement on
// every path (also before early returns): a kernel that skipped it could finish before its predecessor and break the
// transitive ordering of the chain.  Without the launch attribute both instructions are no-ops.
// early_dependents = false for long-running narrow kernels (k_chol_solve): CTAs of the dependent kernel that become
// resident early sit on registers and thread slots until this kernel ends, which would keep the other candidates'
// solvers (other streams) off the SMs.
__device__ __forceinline__ void pdl_prologue(bool early_dependents = true)
{
  if (early_dependents) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
bool pdl_enabled();           // MCP_BA_PDL (ba_api.cu)

// kernel<<<grid, block, smem, s>>>(args...) with the PDL attribute when enabled
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Device-side camera (McpTaylorCam plus the derivative-polynomial coefficients of
// TaylorCamera::RefreshParams, src/TaylorCamera.cc:107-110).
struct DevCam {
  double poly[5];
  double dmod[5];
  double center[2];
  double affine[4];
  double image_size[2];
  double min_theta, theta_mean, theta_std;
  int n_inv;
  int pad_;
  double inv[32];
};

struct Se3 {
  double R[9];
  double t[3];
};

__device__ __forceinline__ void m3_vec(const double* A, const double* v, double* o)
{
  const double x = A[0] * v[0] + A[1] * v[1] + A[2] * v[2];
  const double y = A[3] * v[0] + A[4] * v[1] + A[5] * v[2];
  const double z = A[6] * v[0] + A[7] * v[1] + A[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
__device__ __forceinline__ void m3t_vec(const double* A, const double* v, double* o)
{
  const double x = A[0] * v[0] + A[3] * v[1] + A[6] * v[2];
  const double y = A[1] * v[0] + A[4] * v[1] + A[7] * v[2];
  const double z = A[2] * v[0] + A[5] * v[1] + A[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
__device__ __forceinline__ void m3_mul(const double* A, const double* B, double* C)
{
  double T[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) T[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
#pragma unroll
  for (int i = 0; i < 9; i++) C[i] = T[i];
}
// C = A * B^T
__device__ __forceinline__ void m3_mul_bt(const double* A, const double* B, double* C)
{
  double T[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) T[i * 3 + j] = A[i * 3] * B[j * 3] + A[i * 3 + 1] * B[j * 3 + 1] + A[i * 3 + 2] * B[j * 3 + 2];
#pragma unroll
  for (int i = 0; i < 9; i++) C[i] = T[i];
}
__device__ __forceinline__ void se3_apply(const Se3& T, const double* p, double* o)
{
  double q[3];
  m3_vec(T.R, p, q);
  o[0] = q[0] + T.t[0]; o[1] = q[1] + T.t[1]; o[2] = q[2] + T.t[2];
}
// o = R^T (p - t)
__device__ __forceinline__ void se3_apply_inv(const Se3& T, const double* p, double* o)
{
  const double d[3] = { p[0] - T.t[0], p[1] - T.t[1], p[2] - T.t[2] };
  m3t_vec(T.R, d, o);
}
__device__ __forceinline__ void se3_mul(const Se3& a, const Se3& b, Se3& o)
{
  Se3 r;
  m3_mul(a.R, b.R, r.R);
  m3_vec(a.R, b.t, r.t);
  r.t[0] += a.t[0]; r.t[1] += a.t[1]; r.t[2] += a.t[2];
  o = r;
}
__device__ __forceinline__ void se3_load(const double* p, Se3& T)
{
  // 12 doubles = 96 B, 16 B aligned: three double2 x2 loads
  const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const double2 v = q[i];          // generic load: the pose array may be staged in shared memory
    if (2 * i < 9) T.R[2 * i] = v.x; else T.t[2 * i - 9] = v.x;
    if (2 * i + 1 < 9) T.R[2 * i + 1] = v.y; else T.t[2 * i + 1 - 9] = v.y;
  }
}
__device__ __forceinline__ void se3_store(double* p, const Se3& T)
{
#pragma unroll
  for (int i = 0; i < 9; i++) p[i] = T.R[i];
#pragma unroll
  for (int i = 0; i < 3; i++) p[9 + i] = T.t[i];
}

// TooN rodrigues_so3_exp / SO3::exp / SE3::exp restated (same series thresholds as TooN so3.h/se3.h)
__device__ __forceinline__ void rodrigues(const double* w, double A, double B, double* R)
{
  const double wx2 = w[0] * w[0], wy2 = w[1] * w[1], wz2 = w[2] * w[2];
  R[0] = 1.0 - B * (wy2 + wz2);
  R[4] = 1.0 - B * (wx2 + wz2);
  R[8] = 1.0 - B * (wx2 + wy2);
  { const double a = A * w[2], b = B * (w[0] * w[1]); R[1] = b - a; R[3] = b + a; }
  { const double a = A * w[1], b = B * (w[0] * w[2]); R[2] = b + a; R[6] = b - a; }
  { const double a = A * w[0], b = B * (w[1] * w[2]); R[5] = b - a; R[7] = b + a; }
}
__device__ __forceinline__ void so3_exp(const double* w, double* R)
{
  const double one_6th = 1.0 / 6.0, one_20th = 1.0 / 20.0;
  const double theta_sq = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double theta = sqrt(theta_sq);
  double A, B;
  if (theta_sq < 1e-8) { A = 1.0 - one_6th * theta_sq; B = 0.5; }
  else if (theta_sq < 1e-6) { B = 0.5 - 0.25 * one_6th * theta_sq; A = 1.0 - theta_sq * one_6th * (1.0 - one_20th * theta_sq); }
  else { const double inv_theta = 1.0 / theta; double s, c; sincos(theta, &s, &c); A = s * inv_theta; B = (1 - c) * (inv_theta * inv_theta); }
  rodrigues(w, A, B, R);
}
__device__ __forceinline__ void se3_exp(const double* mu, Se3& T)
{
  const double one_6th = 1.0 / 6.0, one_20th = 1.0 / 20.0;
  const double* w = mu + 3;
  const double theta_sq = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double theta = sqrt(theta_sq);
  double A, B;
  const double cr[3] = { w[1] * mu[2] - w[2] * mu[1], w[2] * mu[0] - w[0] * mu[2], w[0] * mu[1] - w[1] * mu[0] };
  if (theta_sq < 1e-8) {
    A = 1.0 - one_6th * theta_sq; B = 0.5;
    for (int i = 0; i < 3; i++) T.t[i] = mu[i] + 0.5 * cr[i];
  } else {
    double Cc;
    if (theta_sq < 1e-6) { Cc = one_6th * (1.0 - one_20th * theta_sq); A = 1.0 - theta_sq * Cc; B = 0.5 - 0.25 * one_6th * theta_sq; }
    else { const double inv_theta = 1.0 / theta; double s, c; sincos(theta, &s, &c); A = s * inv_theta; B = (1 - c) * (inv_theta * inv_theta); Cc = (1 - A) * (inv_theta * inv_theta); }
    const double wcr[3] = { w[1] * cr[2] - w[2] * cr[1], w[2] * cr[0] - w[0] * cr[2], w[0] * cr[1] - w[1] * cr[0] };
    for (int i = 0; i < 3; i++) T.t[i] = mu[i] + B * cr[i] + Cc * wcr[i];
  }
  rodrigues(w, A, B, T.R);
}

__device__ __forceinline__ double polyval5(const double* c, double x)
{
  double val = 0;
#pragma unroll
  for (int i = 4; i > 0; i--) { val += c[i]; val *= x; }
  return val + c[0];
}

// Project + pixel Jacobian w.r.t. the camera-frame point.
//   px = TaylorCamera::Project(v)                            src/TaylorCamera.cc:202-287
//   G  = GetProjectionDerivs() * [dTheta ; dPhi]  (2x3)      src/TaylorCamera.cc:353-383, 617-669
// Returns the invalid flag.
__device__ __forceinline__ bool cam_project(const DevCam& cam, const double* v, double* px, double* G)
{
  const double n2 = v[0] * v[0] + v[1] * v[1];
  const double norm = sqrt(n2);
  double theta, rho, cphi, sphi;
  if (norm == 0) { theta = 1.5707963267948966; rho = 0; cphi = 0; sphi = 0; }
  else {
    theta = atan(v[2] / norm);
    const double xs = (theta - cam.theta_mean) / cam.theta_std;
    double val = 0;
    for (int i = cam.n_inv - 1; i > 0; i--) { val += cam.inv[i]; val *= xs; }
    rho = val + cam.inv[0];
    cphi = v[0] / norm; sphi = v[1] / norm;
  }
  bool invalid = theta < cam.min_theta;
  const double u = cphi * rho, w = sphi * rho;
  px[0] = cam.affine[0] * u + cam.affine[1] * w + cam.center[0];
  px[1] = cam.affine[2] * u + cam.affine[3] * w + cam.center[1];
  if (!(px[0] >= 0 && px[0] < cam.image_size[0] && px[1] >= 0 && px[1] < cam.image_size[1])) invalid = true;
  if (G) {
    const double wv = polyval5(cam.poly, rho);
    const double drho = (rho * rho + wv * wv) / polyval5(cam.dmod, rho);
    const double dth0 = cphi * drho, dth1 = sphi * drho;
    const double dph0 = -sphi * rho, dph1 = cphi * rho;
    const double D00 = cam.affine[0] * dth0 + cam.affine[1] * dth1;
    const double D10 = cam.affine[2] * dth0 + cam.affine[3] * dth1;
    const double D01 = cam.affine[0] * dph0 + cam.affine[1] * dph1;
    const double D11 = cam.affine[2] * dph0 + cam.affine[3] * dph1;
    double dth[3], dph[3];
    const double x = v[0], y = v[1], z = v[2];
    const double z2 = z * z, nn2 = norm * norm, n3 = nn2 * norm;
    if (norm == 0) { dth[0] = dth[1] = dth[2] = 0; dph[0] = dph[1] = dph[2] = 0; }
    else {
      const double den = n3 + norm * z2;
      dth[0] = -z * x / den; dth[1] = -z * y / den; dth[2] = norm / (nn2 + z2);
      const double xy2 = x * x + y * y;
      dph[0] = -y / xy2; dph[1] = x / xy2; dph[2] = 0;
    }
#pragma unroll
    for (int k = 0; k < 3; k++) { G[k] = D00 * dth[k] + D01 * dph[k]; G[3 + k] = D10 * dth[k] + D11 * dph[k]; }
  }
  return invalid;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace mcp
