/*
 * ba_oracle.c — CPU restatement of the ChainBundle Levenberg–Marquardt bundle adjuster.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Pinned against src/ChainBundle.cc / TaylorCamera.cc / MEstimator.h compiled from the reference (tests/test_oracle_vs_ref.py); g2o / CHOLMOD restated [3P].
 *
 * Follows (file:line in /root/reference):
 *   src/ChainBundle.cc:82-86      VertexPoseSE3::oplusImpl          -> pose_oplus
 *   src/ChainBundle.cc:120-150    PoseChainHelper::UpdateTransforms -> chain_transforms
 *   src/ChainBundle.cc:157-199    PoseChainHelper::MoveTogether     -> move_together
 *   src/ChainBundle.cc:237-281    VertexRelPoint::oplusImpl         -> point_oplus
 *   src/ChainBundle.cc:376-417    EdgeChainMeas::computeError/chi2  -> meas_error
 *   src/ChainBundle.cc:449-685    EdgeChainMeas::linearizeOplus     -> meas_jacobians
 *   src/ChainBundle.cc:810-897    RobustKernelData / RobustKernelAdaptive
 *   src/ChainBundle.cc:1009-1118  convergence actions
 *   src/ChainBundle.cc:1305-1451  ChainBundle::Compute
 *   src/TaylorCamera.cc:202-287, 353-383, 472-485, 617-669
 *   include/mcptam/MEstimator.h:84-126,194-204
 * [3P] g2o OptimizationAlgorithmLevenberg::solve / SparseOptimizer::optimize /
 *      BaseMultiEdge::constructQuadraticForm and TooN SE3/SO3 exp are restated from their
 *      published sources (g2o ~2013 as packaged by ros-hydro-libg2o; TooN 2.x se3.h/so3.h).
 */
#include "oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double R[9]; double t[3]; } se3_t;

/* ---------- small linear algebra ---------- */
static void m3_mul(const double* A, const double* B, double* C)
{
  double T[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      T[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j];
  memcpy(C, T, sizeof(T));
}
static void m3_vec(const double* A, const double* v, double* o)
{
  double x = A[0] * v[0] + A[1] * v[1] + A[2] * v[2];
  double y = A[3] * v[0] + A[4] * v[1] + A[5] * v[2];
  double z = A[6] * v[0] + A[7] * v[1] + A[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static void m3t_vec(const double* A, const double* v, double* o)
{
  double x = A[0] * v[0] + A[3] * v[1] + A[6] * v[2];
  double y = A[1] * v[0] + A[4] * v[1] + A[7] * v[2];
  double z = A[2] * v[0] + A[5] * v[1] + A[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static void cross3(const double* a, const double* b, double* o)
{
  double x = a[1] * b[2] - a[2] * b[1];
  double y = a[2] * b[0] - a[0] * b[2];
  double z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
static void se3_identity(se3_t* T)
{
  memset(T, 0, sizeof(*T));
  T->R[0] = T->R[4] = T->R[8] = 1.0;
}
/* o = a * b   (TooN SE3 product: R = Ra Rb, t = Ra tb + ta) */
static void se3_mul(const se3_t* a, const se3_t* b, se3_t* o)
{
  se3_t r;
  m3_mul(a->R, b->R, r.R);
  m3_vec(a->R, b->t, r.t);
  r.t[0] += a->t[0]; r.t[1] += a->t[1]; r.t[2] += a->t[2];
  *o = r;
}
static void se3_inv(const se3_t* a, se3_t* o)
{
  se3_t r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.R[i * 3 + j] = a->R[j * 3 + i];
  m3_vec(r.R, a->t, r.t);
  r.t[0] = -r.t[0]; r.t[1] = -r.t[1]; r.t[2] = -r.t[2];
  *o = r;
}
static void se3_apply(const se3_t* T, const double* p, double* o)
{
  double q[3];
  m3_vec(T->R, p, q);
  o[0] = q[0] + T->t[0]; o[1] = q[1] + T->t[1]; o[2] = q[2] + T->t[2];
}

/* [3P] TooN so3.h rodrigues_so3_exp */
static void rodrigues(const double* w, double A, double B, double* R)
{
  {
    const double wx2 = w[0] * w[0], wy2 = w[1] * w[1], wz2 = w[2] * w[2];
    R[0] = 1.0 - B * (wy2 + wz2);
    R[4] = 1.0 - B * (wx2 + wz2);
    R[8] = 1.0 - B * (wx2 + wy2);
  }
  { const double a = A * w[2], b = B * (w[0] * w[1]); R[1] = b - a; R[3] = b + a; }
  { const double a = A * w[1], b = B * (w[0] * w[2]); R[2] = b + a; R[6] = b - a; }
  { const double a = A * w[0], b = B * (w[1] * w[2]); R[5] = b - a; R[7] = b + a; }
}
/* [3P] TooN SO3<>::exp */
void ora_so3_exp(const double* w, double* R)
{
  static const double one_6th = 1.0 / 6.0, one_20th = 1.0 / 20.0;
  const double theta_sq = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double theta = sqrt(theta_sq);
  double A, B;
  if (theta_sq < 1e-8) {
    A = 1.0 - one_6th * theta_sq;
    B = 0.5;
  } else if (theta_sq < 1e-6) {
    B = 0.5 - 0.25 * one_6th * theta_sq;
    A = 1.0 - theta_sq * one_6th * (1.0 - one_20th * theta_sq);
  } else {
    const double inv_theta = 1.0 / theta;
    A = sin(theta) * inv_theta;
    B = (1 - cos(theta)) * (inv_theta * inv_theta);
  }
  rodrigues(w, A, B, R);
}
/* [3P] TooN SE3<>::exp, mu = (translation part, rotation part) */
static void se3_exp(const double* mu, se3_t* T)
{
  static const double one_6th = 1.0 / 6.0, one_20th = 1.0 / 20.0;
  const double* w = mu + 3;
  const double theta_sq = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double theta = sqrt(theta_sq);
  double A, B, cr[3];
  cross3(w, mu, cr);
  if (theta_sq < 1e-8) {
    A = 1.0 - one_6th * theta_sq;
    B = 0.5;
    for (int i = 0; i < 3; i++) T->t[i] = mu[i] + 0.5 * cr[i];
  } else {
    double C, wcr[3];
    if (theta_sq < 1e-6) {
      C = one_6th * (1.0 - one_20th * theta_sq);
      A = 1.0 - theta_sq * C;
      B = 0.5 - 0.25 * one_6th * theta_sq;
    } else {
      const double inv_theta = 1.0 / theta;
      A = sin(theta) * inv_theta;
      B = (1 - cos(theta)) * (inv_theta * inv_theta);
      C = (1 - A) * (inv_theta * inv_theta);
    }
    cross3(w, cr, wcr);
    for (int i = 0; i < 3; i++) T->t[i] = mu[i] + B * cr[i] + C * wcr[i];
  }
  rodrigues(w, A, B, T->R);
}
void ora_se3_exp(const double* mu6, double* Rt12)
{
  se3_t T;
  se3_exp(mu6, &T);
  memcpy(Rt12, T.R, 9 * sizeof(double));
  memcpy(Rt12 + 9, T.t, 3 * sizeof(double));
}
/* [3P] TooN generator_field(i, pos) for SE3 on a homogeneous point (pos,1) and SO3 on pos */
static void gen_field(int i, const double* p, double* o)
{
  o[0] = o[1] = o[2] = 0.0;
  if (i < 3) { o[i] = 1.0; return; }
  const int k = i - 3;
  o[(k + 1) % 3] = -p[(k + 2) % 3];
  o[(k + 2) % 3] = p[(k + 1) % 3];
}

/* ---------- Taylor camera ---------- */
/* src/TaylorCamera.cc:472-485 */
static double polyval(const double* c, int n, double x)
{
  double val = 0;
  for (int i = n - 1; i > 0; i--) { val += c[i]; val *= x; }
  val += c[0];
  return val;
}
typedef struct { double rho, cosphi, sinphi; int invalid; } proj_cache_t;
/* src/TaylorCamera.cc:202-287 (live mode with inverse polynomial) */
static void cam_project(const OraTaylorCam* cam, const double* v, double* px, proj_cache_t* pc)
{
  const double norm = sqrt(v[0] * v[0] + v[1] * v[1]);
  double theta;
  if (norm == 0) theta = M_PI_2; else theta = atan(v[2] / norm);
  pc->invalid = (theta < cam->min_theta);
  if (norm == 0) {
    pc->rho = 0; pc->cosphi = 0; pc->sinphi = 0;
  } else {
    pc->rho = polyval(cam->inv_poly, cam->n_inv, (theta - cam->theta_mean) / cam->theta_std);
    pc->cosphi = v[0] / norm;
    pc->sinphi = v[1] / norm;
  }
  const double u = pc->cosphi * pc->rho, w = pc->sinphi * pc->rho;
  px[0] = cam->affine[0] * u + cam->affine[1] * w + cam->center[0];
  px[1] = cam->affine[2] * u + cam->affine[3] * w + cam->center[1];
  /* util::PointInRectangle, include/mcptam/Utility.h:230-237 */
  if (!(px[0] >= 0 && px[0] < cam->image_size[0] && px[1] >= 0 && px[1] < cam->image_size[1])) pc->invalid = 1;
}
/* src/TaylorCamera.cc:353-383; D row-major 2x2 : columns = d/dtheta, d/dphi */
static void cam_derivs(const OraTaylorCam* cam, const proj_cache_t* pc, double* D)
{
  double dm[5] = { -cam->poly[0], cam->poly[1], cam->poly[2], 2 * cam->poly[3], 3 * cam->poly[4] };
  const double w = polyval(cam->poly, 5, pc->rho);
  const double drho = (pc->rho * pc->rho + w * w) / polyval(dm, 5, pc->rho);
  const double dth0 = pc->cosphi * drho, dth1 = pc->sinphi * drho;
  const double dph0 = -pc->sinphi * pc->rho, dph1 = pc->cosphi * pc->rho;
  D[0] = cam->affine[0] * dth0 + cam->affine[1] * dth1;
  D[2] = cam->affine[2] * dth0 + cam->affine[3] * dth1;
  D[1] = cam->affine[0] * dph0 + cam->affine[1] * dph1;
  D[3] = cam->affine[2] * dph0 + cam->affine[3] * dph1;
}
/* src/TaylorCamera.cc:617-669 */
void ora_cam_sphere_deriv(const double* v, double* dth, double* dph)
{
  const double x = v[0], y = v[1], z = v[2];
  const double x2 = x * x, y2 = y * y, z2 = z * z;
  const double n = sqrt(x * x + y * y), n2 = n * n, n3 = n2 * n;
  if (n == 0) { dth[0] = dth[1] = dth[2] = 0; }
  else { dth[0] = -z * x / (n3 + n * z2); dth[1] = -z * y / (n3 + n * z2); dth[2] = n / (n2 + z2); }
  if (x == 0 && y == 0) { dph[0] = dph[1] = dph[2] = 0; }
  else { dph[0] = -y / (x2 + y2); dph[1] = x / (x2 + y2); dph[2] = 0; }
}
int ora_cam_project(const OraTaylorCam* cam, const double* p3, double* px2, double* derivs4)
{
  proj_cache_t pc;
  cam_project(cam, p3, px2, &pc);
  if (derivs4) cam_derivs(cam, &pc, derivs4);
  return pc.invalid;
}
/* src/TaylorCamera.cc:319-346 */
void ora_cam_unproject(const OraTaylorCam* cam, const double* px, double* ray)
{
  /* mm2AffineInv = opts::M2Inverse(mm2Affine) (SmallMatrixOpts.h:66-77: entries times 1/det), then
     mv2LastDistCam = mm2AffineInv * (v2ImFrame - mv2Center)  (src/TaylorCamera.cc:322) -- same operation order, so that
     the result is bit-identical to the reference's (tests/test_oracle_vs_ref.py) */
  const double det = cam->affine[0] * cam->affine[3] - cam->affine[1] * cam->affine[2];
  const double idet = 1.0 / det;
  const double i00 = cam->affine[3] * idet, i01 = -cam->affine[1] * idet, i10 = -cam->affine[2] * idet, i11 = cam->affine[0] * idet;
  const double dx = px[0] - cam->center[0], dy = px[1] - cam->center[1];
  const double u = i00 * dx + i01 * dy;
  const double v = i10 * dx + i11 * dy;
  const double rho = sqrt(u * u + v * v);
  ray[0] = u; ray[1] = v; ray[2] = polyval(cam->poly, 5, rho);
  const double n = sqrt(ray[0] * ray[0] + ray[1] * ray[1] + ray[2] * ray[2]);
  ray[0] /= n; ray[1] /= n; ray[2] /= n;
}

/* ---------- M-estimators ---------- */
static int cmp_double(const void* a, const void* b)
{
  const double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}
static double median_upper(const double* v, int n)
{
  double* tmp = (double*)malloc(sizeof(double) * (size_t)n);
  memcpy(tmp, v, sizeof(double) * (size_t)n);
  qsort(tmp, (size_t)n, sizeof(double), cmp_double);
  const double m = tmp[n / 2];
  free(tmp);
  return m;
}
/* include/mcptam/MEstimator.h:194-204.  (size()*2-6 is size_t arithmetic in the reference.) */
double ora_huber_sigma_sq(const double* v, int n)
{
  const double med = median_upper(v, n);
  const size_t denom = (size_t)n * 2 - 6;
  double s = 1.4826 * (1 + 5.0 / (double)denom) * sqrt(med);
  s = 1.345 * s;
  return s * s;
}
/* include/mcptam/MEstimator.h:109-126 */
double ora_tukey_sigma_sq(const double* v, int n)
{
  const double med = median_upper(v, n);
  const size_t denom = (size_t)n * 2 - 6;
  double s = 1.4826 * (1 + 5.0 / (double)denom) * sqrt(med);
  s = 4.6851 * s;
  return s * s;
}

/* ---------- problem ---------- */
struct OraBa {
  int use_robust, use_tukey;
  int n_cam; OraTaylorCam* cam;
  int n_pose; se3_t* pose; uint8_t* pose_fixed; int* pose_var;   /* var index or -1 */
  int n_pt; double* pt; int32_t* pt_chain; uint8_t* pt_fixed; int* pt_var;
  int n_meas; double* meas_xy; int32_t* meas_chain; int32_t* meas_pt; double* meas_noise; int32_t* meas_cam;
  int n_pose_var, n_pt_var;
  /* CSR measurements by point */
  int* pt_meas_off; int* pt_meas_idx;
  /* working */
  double* err; double* chi2;
  double sigma_sq_raw, sigma_sq_lim, sigma_lim;
  int recompute_sigma;
  double last_chi2;          /* CheckConvergedResidualAction::_dLastChi2 (persists across Compute) */
  double lambda; double ni;
  int converged, hit_max;
  int total_trials;
  double max_cov;
  int n_outliers; int32_t* outliers;
  /* system storage */
  double* x;   /* update vector, size 6*npv + 3*nptv */
  double* b;   /* rhs */
};

OraBa* ora_ba_create(int use_robust, int use_tukey)
{
  OraBa* h = (OraBa*)calloc(1, sizeof(OraBa));
  h->use_robust = use_robust;
  h->use_tukey = use_tukey;
  h->last_chi2 = DBL_MAX;
  h->max_cov = DBL_MAX;
  return h;
}
static void free_problem(OraBa* h)
{
  free(h->pose); free(h->pose_fixed); free(h->pose_var);
  free(h->pt); free(h->pt_chain); free(h->pt_fixed); free(h->pt_var);
  free(h->meas_xy); free(h->meas_chain); free(h->meas_pt); free(h->meas_noise); free(h->meas_cam);
  free(h->pt_meas_off); free(h->pt_meas_idx);
  free(h->err); free(h->chi2); free(h->outliers); free(h->x); free(h->b);
  h->pose = NULL; h->pose_fixed = NULL; h->pose_var = NULL; h->pt = NULL; h->pt_chain = NULL;
  h->pt_fixed = NULL; h->pt_var = NULL; h->meas_xy = NULL; h->meas_chain = NULL; h->meas_pt = NULL;
  h->meas_noise = NULL; h->meas_cam = NULL; h->pt_meas_off = NULL; h->pt_meas_idx = NULL;
  h->err = NULL; h->chi2 = NULL; h->outliers = NULL; h->x = NULL; h->b = NULL;
}
void ora_ba_destroy(OraBa* h)
{
  if (!h) return;
  free_problem(h);
  free(h->cam);
  free(h);
}
int ora_ba_set_cameras(OraBa* h, int n_cam, const OraTaylorCam* cams)
{
  free(h->cam);
  h->cam = (OraTaylorCam*)malloc(sizeof(OraTaylorCam) * (size_t)n_cam);
  memcpy(h->cam, cams, sizeof(OraTaylorCam) * (size_t)n_cam);
  h->n_cam = n_cam;
  return 0;
}
#define DUP(dst, src, n, T) do { dst = (T*)malloc(sizeof(T) * (size_t)((n) > 0 ? (n) : 1)); memcpy(dst, src, sizeof(T) * (size_t)(n)); } while (0)
int ora_ba_load(OraBa* h, int n_pose, const double* pose_Rt, const uint8_t* pose_fixed,
                int n_pt, const double* pt_xyz, const int32_t* pt_chain, const uint8_t* pt_fixed,
                int n_meas, const double* meas_xy, const int32_t* meas_chain,
                const int32_t* meas_pt, const double* meas_noise, const int32_t* meas_cam)
{
  free_problem(h);
  h->n_pose = n_pose; h->n_pt = n_pt; h->n_meas = n_meas;
  h->pose = (se3_t*)malloc(sizeof(se3_t) * (size_t)(n_pose > 0 ? n_pose : 1));
  for (int i = 0; i < n_pose; i++) {
    memcpy(h->pose[i].R, pose_Rt + 12 * i, 9 * sizeof(double));
    memcpy(h->pose[i].t, pose_Rt + 12 * i + 9, 3 * sizeof(double));
  }
  DUP(h->pose_fixed, pose_fixed, n_pose, uint8_t);
  DUP(h->pt, pt_xyz, 3 * n_pt, double);
  DUP(h->pt_chain, pt_chain, 2 * n_pt, int32_t);
  DUP(h->pt_fixed, pt_fixed, n_pt, uint8_t);
  DUP(h->meas_xy, meas_xy, 2 * n_meas, double);
  DUP(h->meas_chain, meas_chain, 2 * n_meas, int32_t);
  DUP(h->meas_pt, meas_pt, n_meas, int32_t);
  DUP(h->meas_noise, meas_noise, n_meas, double);
  DUP(h->meas_cam, meas_cam, n_meas, int32_t);
  h->pose_var = (int*)malloc(sizeof(int) * (size_t)(n_pose > 0 ? n_pose : 1));
  h->pt_var = (int*)malloc(sizeof(int) * (size_t)(n_pt > 0 ? n_pt : 1));
  h->n_pose_var = 0; h->n_pt_var = 0;
  for (int i = 0; i < n_pose; i++) h->pose_var[i] = pose_fixed[i] ? -1 : h->n_pose_var++;
  for (int i = 0; i < n_pt; i++) h->pt_var[i] = pt_fixed[i] ? -1 : h->n_pt_var++;
  for (int i = 0; i < n_meas; i++) {
    if (meas_pt[i] < 0 || meas_pt[i] >= n_pt) return -2;
    if (meas_cam[i] < 0 || meas_cam[i] >= h->n_cam) return -2;
    if (meas_chain[2 * i] < 0 || meas_chain[2 * i] >= n_pose || meas_chain[2 * i + 1] >= n_pose) return -2;
  }
  for (int i = 0; i < n_pt; i++)
    if (pt_chain[2 * i] < 0 || pt_chain[2 * i] >= n_pose || pt_chain[2 * i + 1] >= n_pose) return -2;
  /* CSR by point */
  h->pt_meas_off = (int*)calloc((size_t)n_pt + 1, sizeof(int));
  h->pt_meas_idx = (int*)malloc(sizeof(int) * (size_t)(n_meas > 0 ? n_meas : 1));
  for (int i = 0; i < n_meas; i++) h->pt_meas_off[meas_pt[i] + 1]++;
  for (int i = 0; i < n_pt; i++) h->pt_meas_off[i + 1] += h->pt_meas_off[i];
  int* cur = (int*)malloc(sizeof(int) * (size_t)(n_pt > 0 ? n_pt : 1));
  memcpy(cur, h->pt_meas_off, sizeof(int) * (size_t)n_pt);
  for (int i = 0; i < n_meas; i++) h->pt_meas_idx[cur[meas_pt[i]]++] = i;
  free(cur);
  h->err = (double*)calloc((size_t)(2 * n_meas + 2), sizeof(double));
  h->chi2 = (double*)calloc((size_t)(n_meas + 1), sizeof(double));
  h->outliers = (int32_t*)calloc((size_t)(n_meas + 1), sizeof(int32_t));
  const int dim = 6 * h->n_pose_var + 3 * h->n_pt_var;
  h->x = (double*)calloc((size_t)dim + 1, sizeof(double));
  h->b = (double*)calloc((size_t)dim + 1, sizeof(double));
  return 0;
}
int ora_ba_get_poses(const OraBa* h, double* o)
{
  for (int i = 0; i < h->n_pose; i++) { memcpy(o + 12 * i, h->pose[i].R, 72); memcpy(o + 12 * i + 9, h->pose[i].t, 24); }
  return 0;
}
int ora_ba_get_points(const OraBa* h, double* o) { memcpy(o, h->pt, sizeof(double) * 3 * (size_t)h->n_pt); return 0; }
int ora_ba_set_poses(OraBa* h, const double* o)
{
  for (int i = 0; i < h->n_pose; i++) { memcpy(h->pose[i].R, o + 12 * i, 72); memcpy(h->pose[i].t, o + 12 * i + 9, 24); }
  return 0;
}
int ora_ba_set_points(OraBa* h, const double* o) { memcpy(h->pt, o, sizeof(double) * 3 * (size_t)h->n_pt); return 0; }
int ora_ba_get_outliers(const OraBa* h, int32_t* idx, int cap)
{
  int n = h->n_outliers < cap ? h->n_outliers : cap;
  memcpy(idx, h->outliers, sizeof(int32_t) * (size_t)n);
  return h->n_outliers;
}

/* ---------- chains ---------- */
typedef struct {
  int n;            /* chain length 1 or 2 */
  int id[2];
  se3_t first[2];   /* _vTransforms[i].first : pose_i * ... * pose_0 */
  double secondR[2][9]; /* rotation of _vTransforms[i].second : pose_last*...*pose_{i+1} */
} chain_t;
/* src/ChainBundle.cc:120-150 */
static void chain_transforms(const OraBa* h, const int32_t* ids, chain_t* c)
{
  c->n = ids[1] >= 0 ? 2 : 1;
  c->id[0] = ids[0]; c->id[1] = ids[1];
  se3_t acc; se3_identity(&acc);
  for (int i = 0; i < c->n; i++) { se3_mul(&h->pose[ids[i]], &acc, &acc); c->first[i] = acc; }
  se3_t back; se3_identity(&back);
  for (int i = c->n - 1; i >= 0; i--) {
    memcpy(c->secondR[i], back.R, sizeof(back.R));
    se3_mul(&back, &h->pose[ids[i]], &back);
  }
}
/* src/ChainBundle.cc:157-199 */
static int move_together(const OraBa* h, const chain_t* self, const chain_t* other, int depth)
{
  int furthest = -1;
  for (;;) {
    const int t = furthest + 1;
    if (self->n <= t || other->n <= t) break;
    if (self->id[t] != other->id[t]) break;
    furthest = t;
    if (furthest == depth) return 1;
  }
  if (furthest == -1) return 0;
  for (int i = furthest; i <= depth; i++)
    if (!h->pose_fixed[self->id[i]]) return 0;
  return 1;
}

/* src/ChainBundle.cc:376-417.  Returns chi2 (signed as in the reference). */
static double meas_error(const OraBa* h, int m, double* e, double* v3cam_out, proj_cache_t* pc_out,
                         chain_t* obs_out, chain_t* src_out)
{
  chain_t obs, src;
  const int p = h->meas_pt[m];
  chain_transforms(h, h->meas_chain + 2 * m, &obs);
  chain_transforms(h, h->pt_chain + 2 * p, &src);
  se3_t srcinv;
  se3_inv(&src.first[src.n - 1], &srcinv);
  double glob[3], vcam[3], px[2];
  se3_apply(&srcinv, h->pt + 3 * p, glob);                 /* estimateInGlobalCartesian :320-323 */
  se3_apply(&obs.first[obs.n - 1], glob, vcam);
  proj_cache_t pc;
  cam_project(&h->cam[h->meas_cam[m]], vcam, px, &pc);
  e[0] = h->meas_xy[2 * m] - px[0];
  e[1] = h->meas_xy[2 * m + 1] - px[1];
  const double info = 1.0 / sqrt(h->meas_noise[m]);        /* :1244-1245 */
  double val = e[0] * (info * e[0]) + e[1] * (info * e[1]);
  if (h->pt_fixed[p] && h->use_robust) val *= -1;          /* :413-414 */
  if (v3cam_out) memcpy(v3cam_out, vcam, 24);
  if (pc_out) *pc_out = pc;
  if (obs_out) *obs_out = obs;
  if (src_out) *src_out = src;
  return val;
}

/* src/ChainBundle.cc:449-685.  J_obs[i] / J_src[i]: 2x6 row-major for chain link i; J_pt 2x3. */
static void meas_jacobians(const OraBa* h, int m, double J_obs[2][12], double J_src[2][12], double J_pt[6])
{
  double e[2], vcam[3];
  proj_cache_t pc;
  chain_t obs, src;
  meas_error(h, m, e, vcam, &pc, &obs, &src);
  const int p = h->meas_pt[m];
  double D[4], dth[3], dph[3];
  cam_derivs(&h->cam[h->meas_cam[m]], &pc, D);
  ora_cam_sphere_deriv(vcam, dth, dph);
  se3_t srcinv;
  se3_inv(&src.first[src.n - 1], &srcinv);
  double glob[3];
  se3_apply(&srcinv, h->pt + 3 * p, glob);
  memset(J_obs, 0, sizeof(double) * 24);
  memset(J_src, 0, sizeof(double) * 24);
  memset(J_pt, 0, sizeof(double) * 6);

  for (int i = 0; i < obs.n; i++) {                         /* :485-532 */
    if (h->pose_fixed[obs.id[i]]) continue;
    if (move_together(h, &obs, &src, i)) continue;
    double base[3];
    se3_apply(&obs.first[i], glob, base);
    for (int k = 0; k < 6; k++) {
      double mb[3], mc[3];
      gen_field(k, base, mb);
      m3_vec(obs.secondR[i], mb, mc);
      const double s0 = dth[0] * mc[0] + dth[1] * mc[1] + dth[2] * mc[2];
      const double s1 = dph[0] * mc[0] + dph[1] * mc[1] + dph[2] * mc[2];
      J_obs[i][k] = -1 * (D[0] * s0 + D[1] * s1);
      J_obs[i][6 + k] = -1 * (D[2] * s0 + D[3] * s1);
    }
  }
  for (int i = 0; i < src.n; i++) {                         /* :535-586 */
    if (h->pose_fixed[src.id[i]]) continue;
    if (move_together(h, &src, &obs, i)) continue;
    double base[3];
    se3_apply(&src.first[i], glob, base);
    se3_t inv_i, cfb;
    se3_inv(&src.first[i], &inv_i);
    se3_mul(&obs.first[obs.n - 1], &inv_i, &cfb);           /* :567 */
    for (int k = 0; k < 6; k++) {
      double mb[3], mc[3];
      gen_field(k, base, mb);
      mb[0] = -mb[0]; mb[1] = -mb[1]; mb[2] = -mb[2];
      m3_vec(cfb.R, mb, mc);
      const double s0 = dth[0] * mc[0] + dth[1] * mc[1] + dth[2] * mc[2];
      const double s1 = dph[0] * mc[0] + dph[1] * mc[1] + dph[2] * mc[2];
      J_src[i][k] = -1 * (D[0] * s0 + D[1] * s1);
      J_src[i][6 + k] = -1 * (D[2] * s0 + D[3] * s1);
    }
  }
  if (!h->pt_fixed[p]) {                                    /* :589-684 */
    const double* pc3 = h->pt + 3 * p;
    const double len = sqrt(pc3[0] * pc3[0] + pc3[1] * pc3[1] + pc3[2] * pc3[2]);
    const double rho = 1.0 / len;
    double dir[3] = { pc3[0] * rho, pc3[1] * rho, pc3[2] * rho };
    const double ez[3] = { 0, 0, 1 };
    double axis[3];
    cross3(dir, ez, axis);
    const double an = sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
    const double angle = asin(an);
    axis[0] = axis[0] / an * angle; axis[1] = axis[1] / an * angle; axis[2] = axis[2] / an * angle;
    double Rp[9];
    ora_so3_exp(axis, Rp);
    double rpp[3], g0[3], g1[3], M[9]; /* M columns = motion vectors */
    m3_vec(Rp, pc3, rpp);
    gen_field(3, rpp, g0);
    gen_field(4, rpp, g1);
    double c0[3], c1[3], c2[3];
    m3t_vec(Rp, g0, c0);
    m3t_vec(Rp, g1, c1);
    c2[0] = -1 * pc3[0] / rho; c2[1] = -1 * pc3[1] / rho; c2[2] = -1 * pc3[2] / rho;
    for (int r = 0; r < 3; r++) { M[r * 3 + 0] = c0[r]; M[r * 3 + 1] = c1[r]; M[r * 3 + 2] = c2[r]; }
    se3_t cfs;
    se3_mul(&obs.first[obs.n - 1], &srcinv, &cfs);          /* :659 */
    double RM[9];
    m3_mul(cfs.R, M, RM);
    for (int k = 0; k < 3; k++) {
      const double mv[3] = { RM[0 * 3 + k], RM[1 * 3 + k], RM[2 * 3 + k] };
      const double s0 = dth[0] * mv[0] + dth[1] * mv[1] + dth[2] * mv[2];
      const double s1 = dph[0] * mv[0] + dph[1] * mv[1] + dph[2] * mv[2];
      J_pt[k] = -1 * (D[0] * s0 + D[1] * s1);
      J_pt[3 + k] = -1 * (D[2] * s0 + D[3] * s1);
    }
  }
}
int ora_ba_jacobians(OraBa* h, int meas, double* J_obs, double* J_src, double* J_pt)
{
  double a[2][12], b[2][12];
  meas_jacobians(h, meas, a, b, J_pt);
  memcpy(J_obs, a, sizeof(a));
  memcpy(J_src, b, sizeof(b));
  return 0;
}

/* src/ChainBundle.cc:82-86 */
static void pose_oplus(se3_t* T, const double* d6)
{
  se3_t E;
  se3_exp(d6, &E);
  se3_mul(&E, T, T);
}
/* src/ChainBundle.cc:237-281 */
static void point_oplus(double* est, const double* upd)
{
  const double dist_before = sqrt(est[0] * est[0] + est[1] * est[1] + est[2] * est[2]);
  const double rho_before = 1.0 / dist_before;
  double dir[3] = { est[0] * rho_before, est[1] * rho_before, est[2] * rho_before };
  const double ez[3] = { 0, 0, 1 };
  double axis[3];
  cross3(dir, ez, axis);
  const double an = sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
  const double angle = asin(an);
  axis[0] = axis[0] / an * angle; axis[1] = axis[1] / an * angle; axis[2] = axis[2] / an * angle;
  double Rp[9], Ru[9];
  ora_so3_exp(axis, Rp);
  const double w[3] = { upd[0], upd[1], 0 };
  ora_so3_exp(w, Ru);
  double a[3], b[3], c[3];
  /* (Rp^-1 * exp * Rp) * dir : SO3 products are formed left to right in the reference expression */
  double RpT[9], M1[9], M2[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) RpT[i * 3 + j] = Rp[j * 3 + i];
  m3_mul(RpT, Ru, M1);
  m3_mul(M1, Rp, M2);
  m3_vec(M2, dir, c);
  (void)a; (void)b;
  const double s = 1 / (rho_before + upd[2]);
  est[0] = s * c[0]; est[1] = s * c[1]; est[2] = s * c[2];
  const double dist_after = sqrt(est[0] * est[0] + est[1] * est[1] + est[2] * est[2]);
  if (dist_after > 1e5) { const double f = 1e5 / dist_after; est[0] *= f; est[1] *= f; est[2] *= f; }
  if (dist_after < 1e-5) { const double f = 1e-5 / dist_after; est[0] *= f; est[1] *= f; est[2] *= f; }
}
int ora_ba_oplus_pose(OraBa* h, int id, const double* d6) { pose_oplus(&h->pose[id], d6); return 0; }
int ora_ba_oplus_point(OraBa* h, int id, const double* d3) { point_oplus(h->pt + 3 * id, d3); return 0; }

/* computeActiveErrors */
static void compute_errors(OraBa* h)
{
  for (int m = 0; m < h->n_meas; m++) h->chi2[m] = meas_error(h, m, h->err + 2 * m, NULL, NULL, NULL, NULL);
}
int ora_ba_eval(OraBa* h, double* err_xy, double* chi2)
{
  compute_errors(h);
  if (err_xy) memcpy(err_xy, h->err, sizeof(double) * 2 * (size_t)h->n_meas);
  if (chi2) memcpy(chi2, h->chi2, sizeof(double) * (size_t)h->n_meas);
  return 0;
}
/* RobustKernelData::RecomputeNow, src/ChainBundle.cc:810-833 (sdMinMEstimatorSigma^2 = 0.25, :1136,1148) */
static void recompute_sigma(OraBa* h)
{
  h->recompute_sigma = 0;
  double* a = (double*)malloc(sizeof(double) * (size_t)(h->n_meas > 0 ? h->n_meas : 1));
  for (int m = 0; m < h->n_meas; m++) a[m] = fabs(h->chi2[m]);
  h->sigma_sq_raw = ora_huber_sigma_sq(a, h->n_meas);
  free(a);
  h->sigma_sq_lim = h->sigma_sq_raw < 0.25 ? 0.25 : h->sigma_sq_raw;
  h->sigma_lim = sqrt(h->sigma_sq_lim);
}
/* RobustKernelAdaptive::robustify, :871-897 */
static void robustify(OraBa* h, double e2, double* rho)
{
  if (h->recompute_sigma) recompute_sigma(h);
  if (e2 <= h->sigma_sq_lim) { rho[0] = fabs(e2); rho[1] = 1.; rho[2] = 0.; }
  else {
    const double e = sqrt(e2);
    rho[0] = 2 * h->sigma_lim * e - h->sigma_sq_lim;
    rho[1] = h->sigma_lim / e;
    rho[2] = -0.5 * rho[1] / e2;
  }
}
/* [3P] SparseOptimizer::activeRobustChi2 */
static double robust_chi2(OraBa* h)
{
  double chi = 0, rho[3];
  for (int m = 0; m < h->n_meas; m++) {
    if (h->use_robust) { robustify(h, h->chi2[m], rho); chi += rho[0]; }
    else chi += h->chi2[m];
  }
  return chi;
}

/* ---------- linear system ---------- */
/* In-place lower Cholesky of dense n x n (row-major, lower triangle used). returns 0 ok */
static int chol_dense(double* A, int n)
{
  for (int j = 0; j < n; j++) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; k++) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0)) return -1;
    d = sqrt(d);
    A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = A[(size_t)i * n + j];
      const double* ai = A + (size_t)i * n;
      const double* aj = A + (size_t)j * n;
      for (int k = 0; k < j; k++) s -= ai[k] * aj[k];
      A[(size_t)i * n + j] = s / d;
    }
  }
  return 0;
}
static void chol_solve(const double* L, int n, double* x)
{
  for (int i = 0; i < n; i++) {
    double s = x[i];
    for (int k = 0; k < i; k++) s -= L[(size_t)i * n + k] * x[k];
    x[i] = s / L[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = x[i];
    for (int k = i + 1; k < n; k++) s -= L[(size_t)k * n + i] * x[k];
    x[i] = s / L[(size_t)i * n + i];
  }
}
static int inv3_sym(const double* V, double* Vi)
{
  /* Cholesky-based 3x3 SPD inverse */
  double L[9];
  memcpy(L, V, sizeof(L));
  if (chol_dense(L, 3)) return -1;
  for (int c = 0; c < 3; c++) {
    double e[3] = { 0, 0, 0 };
    e[c] = 1;
    chol_solve(L, 3, e);
    Vi[0 * 3 + c] = e[0]; Vi[1 * 3 + c] = e[1]; Vi[2 * 3 + c] = e[2];
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------------------------------
   solve_mode 3: the FULL non-marginalised system, as the reference configures g2o (BlockSolverX + LinearSolverCholmod with
   every point vertex setMarginalized(false), src/ChainBundle.cc:1150-1158, 1218).  CHOLMOD is not in this image; this is
   a general block-sparse right-looking Cholesky with a minimum-degree elimination order chosen on the block graph (what
   CHOLMOD's AMD does on this matrix up to tie-breaking): nothing in it knows about "points" and "poses" -- blocks of size
   3 and 6 live in a hash map keyed by (row block, column block), fill blocks are created as they arise, the order comes
   from a degree-bucket queue.  On a bundle-adjustment matrix minimum degree eliminates the point blocks first (their degree
   is the handful of poses they touch, a pose's degree is hundreds of points until they are gone and then the whole pose
   clique), i.e. it performs the Schur complement -- which is why the default restatement (solve_mode 0) is the same
   arithmetic without the bookkeeping.  Used as the second CPU baseline of bench.py and as an independent check of the
   Schur path (tests/test_oracle_cpu.py). */
typedef struct { long long key; double* v; } fs_slot_t;
typedef struct {
  int nb;                 /* blocks */
  const int* bdim;        /* 3 or 6 */
  fs_slot_t* tab; size_t cap, used;
  int** adj; int* nadj; int* capadj;    /* neighbour lists (may hold eliminated nodes; filtered on use) */
  double** chunk; int n_chunk; size_t chunk_used;   /* bump allocator for the blocks (zeroed chunks of FS_CHUNK doubles) */
} fs_t;
#define FS_CHUNK (1u << 20)
static double* fs_alloc(fs_t* f, size_t n)
{
  if (f->n_chunk == 0 || f->chunk_used + n > FS_CHUNK) {
    f->chunk = (double**)realloc(f->chunk, sizeof(double*) * (size_t)(f->n_chunk + 1));
    f->chunk[f->n_chunk++] = (double*)calloc(FS_CHUNK, sizeof(double));
    f->chunk_used = 0;
  }
  double* p = f->chunk[f->n_chunk - 1] + f->chunk_used;
  f->chunk_used += n;
  return p;
}
static size_t fs_hash(long long k, size_t cap) { unsigned long long x = (unsigned long long)k * 0x9E3779B97F4A7C15ull; return (size_t)(x >> 20) & (cap - 1); }
static double* fs_get(fs_t* f, int i, int j, int create)
{
  const long long key = (long long)i * f->nb + j;
  size_t h = fs_hash(key, f->cap);
  for (;;) {
    if (f->tab[h].v == NULL) break;
    if (f->tab[h].key == key) return f->tab[h].v;
    h = (h + 1) & (f->cap - 1);
  }
  if (!create) return NULL;
  f->tab[h].key = key;
  f->tab[h].v = fs_alloc(f, (size_t)f->bdim[i] * f->bdim[j]);
  f->used++;
  return f->tab[h].v;
}
static void fs_link(fs_t* f, int i, int j)
{
  if (f->nadj[i] == f->capadj[i]) { f->capadj[i] = f->capadj[i] ? 2 * f->capadj[i] : 8; f->adj[i] = (int*)realloc(f->adj[i], sizeof(int) * (size_t)f->capadj[i]); }
  f->adj[i][f->nadj[i]++] = j;
}
/* block (i,j) with i != j is stored once, under (min, max), as a bdim[min] x bdim[max] row-major matrix */
static double* fs_off(fs_t* f, int i, int j, int create, int* transposed)
{
  const int a = i < j ? i : j, b = i < j ? j : i;
  *transposed = i > j;
  double* v = fs_get(f, a, b, 0);
  if (!v && create) { v = fs_get(f, a, b, 1); fs_link(f, a, b); fs_link(f, b, a); }
  return v;
}
static int full_sparse_solve(int nb, const int* bdim, const int* boff, fs_t* f, const double* rhs, double* x)
{
  int rc = 0;
  char* gone = (char*)calloc((size_t)nb + 1, 1);
  int* order = (int*)malloc(sizeof(int) * (size_t)(nb + 1));
  int* deg = (int*)calloc((size_t)nb + 1, sizeof(int));
  /* factor storage per eliminated block: its Cholesky factor and the column blocks L_ik */
  double** Lkk = (double**)calloc((size_t)nb + 1, sizeof(double*));
  int** col_i = (int**)calloc((size_t)nb + 1, sizeof(int*));
  double*** col_L = (double***)calloc((size_t)nb + 1, sizeof(double**));
  int* col_n = (int*)calloc((size_t)nb + 1, sizeof(int));
  /* degree buckets (degree = number of remaining neighbours): doubly linked lists */
  int* head = (int*)malloc(sizeof(int) * (size_t)(nb + 2));
  int* nxt = (int*)malloc(sizeof(int) * (size_t)(nb + 1));
  int* prv = (int*)malloc(sizeof(int) * (size_t)(nb + 1));
  for (int d = 0; d <= nb; d++) head[d] = -1;
#define FS_UNLINK(i) do { if (prv[i] >= 0) nxt[prv[i]] = nxt[i]; else head[deg[i]] = nxt[i]; if (nxt[i] >= 0) prv[nxt[i]] = prv[i]; } while (0)
#define FS_PUSH(i) do { prv[i] = -1; nxt[i] = head[deg[i]]; if (head[deg[i]] >= 0) prv[head[deg[i]]] = i; head[deg[i]] = i; } while (0)
  for (int i = nb - 1; i >= 0; i--) { deg[i] = f->nadj[i]; FS_PUSH(i); }       /* (pushed in reverse: ties resolve to the lowest index) */
  int* nbr = (int*)malloc(sizeof(int) * (size_t)(nb + 1));
  char* mark = (char*)calloc((size_t)nb + 1, 1);
  int dmin = 0;
  for (int step = 0; step < nb && !rc; step++) {
    while (dmin <= nb && head[dmin] < 0) dmin++;
    const int k = head[dmin];
    FS_UNLINK(k);
    gone[k] = 1; order[step] = k;
    /* remaining neighbours of k (the lists may repeat a node or hold eliminated ones) */
    int d = 0;
    for (int t = 0; t < f->nadj[k]; t++) { const int i = f->adj[k][t]; if (!gone[i] && !mark[i]) { mark[i] = 1; nbr[d++] = i; } }
    for (int t = 0; t < d; t++) mark[nbr[t]] = 0;
    const int dk = bdim[k];
    double* Akk = fs_get(f, k, k, 0);
    Lkk[k] = fs_alloc(f, (size_t)dk * dk);
    memcpy(Lkk[k], Akk, sizeof(double) * (size_t)dk * dk);
    if (chol_dense(Lkk[k], dk)) { rc = -1; break; }
    col_n[k] = d;
    col_i[k] = (int*)malloc(sizeof(int) * (size_t)(d + 1));
    col_L[k] = (double**)malloc(sizeof(double*) * (size_t)(d + 1));
    for (int t = 0; t < d; t++) {
      const int i = nbr[t], di = bdim[i];
      int tr;
      const double* Aik = fs_off(f, i, k, 0, &tr);               /* A_ik (di x dk); stored under (min,max) */
      double* L = fs_alloc(f, (size_t)di * dk);
      for (int r = 0; r < di; r++) {
        /* row r of L_ik solves  L_ik Lkk^T = A_ik  (forward substitution along the row) */
        for (int c = 0; c < dk; c++) {
          double sacc = tr ? Aik[(size_t)c * di + r] : Aik[(size_t)r * dk + c];
          for (int q = 0; q < c; q++) sacc -= L[(size_t)r * dk + q] * Lkk[k][(size_t)c * dk + q];
          L[(size_t)r * dk + c] = sacc / Lkk[k][(size_t)c * dk + c];
        }
      }
      col_i[k][t] = i; col_L[k][t] = L;
    }
    /* Schur update of the remaining sub-matrix: A_ij -= L_ik L_jk^T for every pair of neighbours (fill where absent) */
    for (int t = 0; t < d; t++) {
      const int i = nbr[t], di = bdim[i];
      const double* Li = col_L[k][t];
      double* Aii = fs_get(f, i, i, 0);
      for (int r = 0; r < di; r++)
        for (int c = 0; c < di; c++) {
          double sacc = 0;
          for (int q = 0; q < dk; q++) sacc += Li[(size_t)r * dk + q] * Li[(size_t)c * dk + q];
          Aii[(size_t)r * di + c] -= sacc;
        }
      for (int u = t + 1; u < d; u++) {
        const int j = nbr[u], dj = bdim[j];
        const double* Lj = col_L[k][u];
        int tr;
        const int before = f->nadj[i];
        double* Aij = fs_off(f, i, j, 1, &tr);
        if (f->nadj[i] != before) {                                /* a fill block: both degrees grow */
          FS_UNLINK(i); deg[i]++; FS_PUSH(i);
          FS_UNLINK(j); deg[j]++; FS_PUSH(j);
        }
        for (int r = 0; r < di; r++)
          for (int c = 0; c < dj; c++) {
            double sacc = 0;
            for (int q = 0; q < dk; q++) sacc += Li[(size_t)r * dk + q] * Lj[(size_t)c * dk + q];
            if (tr) Aij[(size_t)c * di + r] -= sacc; else Aij[(size_t)r * dj + c] -= sacc;
          }
      }
      /* i lost the neighbour k */
      FS_UNLINK(i); deg[i]--; FS_PUSH(i);
      if (deg[i] < dmin) dmin = deg[i];
    }
  }
  if (!rc) {
    /* forward: y_k = Lkk^-1 b_k, b_i -= L_ik y_k; backward in reverse order */
    double* y = (double*)malloc(sizeof(double) * (size_t)(boff[nb] + 1));
    memcpy(y, rhs, sizeof(double) * (size_t)boff[nb]);
    for (int step = 0; step < nb; step++) {
      const int k = order[step], dk = bdim[k];
      double* yk = y + boff[k];
      for (int r = 0; r < dk; r++) { double sacc = yk[r]; for (int q = 0; q < r; q++) sacc -= Lkk[k][(size_t)r * dk + q] * yk[q]; yk[r] = sacc / Lkk[k][(size_t)r * dk + r]; }
      for (int t = 0; t < col_n[k]; t++) {
        const int i = col_i[k][t], di = bdim[i];
        const double* L = col_L[k][t];
        double* yi = y + boff[i];
        for (int r = 0; r < di; r++) { double sacc = 0; for (int q = 0; q < dk; q++) sacc += L[(size_t)r * dk + q] * yk[q]; yi[r] -= sacc; }
      }
    }
    for (int step = nb - 1; step >= 0; step--) {
      const int k = order[step], dk = bdim[k];
      double* yk = y + boff[k];
      for (int t = 0; t < col_n[k]; t++) {
        const int i = col_i[k][t], di = bdim[i];
        const double* L = col_L[k][t];
        const double* xi = y + boff[i];
        for (int q = 0; q < dk; q++) { double sacc = 0; for (int r = 0; r < di; r++) sacc += L[(size_t)r * dk + q] * xi[r]; yk[q] -= sacc; }
      }
      for (int r = dk - 1; r >= 0; r--) { double sacc = yk[r]; for (int q = r + 1; q < dk; q++) sacc -= Lkk[k][(size_t)q * dk + r] * yk[q]; yk[r] = sacc / Lkk[k][(size_t)r * dk + r]; }
    }
    memcpy(x, y, sizeof(double) * (size_t)boff[nb]);
    free(y);
  }
#undef FS_UNLINK
#undef FS_PUSH
  for (int k = 0; k < nb; k++) { free(col_i[k]); free(col_L[k]); }
  free(gone); free(order); free(deg); free(Lkk); free(col_i); free(col_L); free(col_n); free(head); free(nxt); free(prv); free(nbr); free(mark);
  return rc;
}

#define MAXSLOT 256
typedef struct { int nslot; int var[MAXSLOT]; double W[MAXSLOT][18]; double V[9]; double bp[3]; } ptblk_t;

/* Build the normal equations at the current state (errors must be current) and solve with lambda.
   Fills h->x and h->b.  Returns 0 ok / -1 solver failure.
   [3P] g2o BlockSolver::buildSystem + BaseMultiEdge::constructQuadraticForm (first-order robust weight). */
static int build_and_solve(OraBa* h, double lambda, int solve_mode)
{
  const int npv = h->n_pose_var, nptv = h->n_pt_var;
  const int nc = 6 * npv, dim = nc + 3 * nptv;
  double* Hcc = (double*)calloc((size_t)nc * (size_t)nc + 1, sizeof(double));
  double* bc = h->b;
  double* bpall = h->b + nc;
  memset(h->b, 0, sizeof(double) * (size_t)dim);
  ptblk_t* blk = (ptblk_t*)malloc(sizeof(ptblk_t));
  double* Hfull = NULL;
  if (solve_mode == 1 || solve_mode == 2) Hfull = (double*)calloc((size_t)dim * (size_t)dim + 1, sizeof(double));
  double* V_all = solve_mode == 3 ? (double*)calloc((size_t)nptv * 9 + 1, sizeof(double)) : NULL;   /* point blocks of the full system */
  /* per-point storage for Schur back-substitution */
  double* Vinv_all = (double*)calloc((size_t)nptv * 9 + 1, sizeof(double));
  int* slot_off = (int*)calloc((size_t)h->n_pt + 1, sizeof(int));
  int cap = 0;
  for (int p = 0; p < h->n_pt; p++) { slot_off[p] = cap; cap += 2 * (h->pt_meas_off[p + 1] - h->pt_meas_off[p]) + 2; }
  slot_off[h->n_pt] = cap;
  int* slot_var = (int*)malloc(sizeof(int) * (size_t)(cap + 1));
  double* slot_W = (double*)calloc((size_t)cap * 18 + 1, sizeof(double));
  int* slot_cnt = (int*)calloc((size_t)h->n_pt + 1, sizeof(int));
  int fail = 0;

  for (int p = 0; p < h->n_pt; p++) {
    blk->nslot = 0;
    memset(blk->V, 0, sizeof(blk->V));
    memset(blk->bp, 0, sizeof(blk->bp));
    for (int q = h->pt_meas_off[p]; q < h->pt_meas_off[p + 1]; q++) {
      const int m = h->pt_meas_idx[q];
      double Jo[2][12], Js[2][12], Jp[6], rho[3] = { 0, 1, 0 };
      meas_jacobians(h, m, Jo, Js, Jp);
      if (h->use_robust) robustify(h, h->chi2[m], rho);
      const double info = 1.0 / sqrt(h->meas_noise[m]);
      const double w = rho[1] * info;                       /* robustInformation = rho1 * Omega */
      const double* e = h->err + 2 * m;
      /* gather pose jacobians by variable (sum if a variable occurs twice) */
      int nv = 0, vars[4];
      double J[4][12];
      const int32_t* oc = h->meas_chain + 2 * m;
      const int32_t* sc = h->pt_chain + 2 * p;
      for (int side = 0; side < 2; side++)
        for (int i = 0; i < 2; i++) {
          const int id = side == 0 ? oc[i] : sc[i];
          if (id < 0 || h->pose_var[id] < 0) continue;
          const double* Jsrc = side == 0 ? Jo[i] : Js[i];
          int k;
          for (k = 0; k < nv; k++) if (vars[k] == h->pose_var[id]) break;
          if (k == nv) { vars[nv] = h->pose_var[id]; memset(J[nv], 0, sizeof(J[nv])); nv++; }
          for (int t = 0; t < 12; t++) J[k][t] += Jsrc[t];
        }
      /* pose-pose and pose rhs */
      for (int a = 0; a < nv; a++) {
        for (int r = 0; r < 6; r++) bc[6 * vars[a] + r] -= w * (J[a][r] * e[0] + J[a][6 + r] * e[1]);
        for (int c2 = 0; c2 < nv; c2++)
          for (int r = 0; r < 6; r++)
            for (int c = 0; c < 6; c++)
              Hcc[(size_t)(6 * vars[a] + r) * nc + 6 * vars[c2] + c] += w * (J[a][r] * J[c2][c] + J[a][6 + r] * J[c2][6 + c]);
      }
      if (h->pt_var[p] >= 0) {
        for (int r = 0; r < 3; r++) {
          blk->bp[r] -= w * (Jp[r] * e[0] + Jp[3 + r] * e[1]);
          for (int c = 0; c < 3; c++) blk->V[r * 3 + c] += w * (Jp[r] * Jp[c] + Jp[3 + r] * Jp[3 + c]);
        }
        for (int a = 0; a < nv; a++) {
          int s;
          for (s = 0; s < blk->nslot; s++) if (blk->var[s] == vars[a]) break;
          if (s == blk->nslot) {
            if (blk->nslot >= MAXSLOT) { fail = 1; break; }
            blk->var[s] = vars[a]; memset(blk->W[s], 0, sizeof(blk->W[s])); blk->nslot++;
          }
          for (int r = 0; r < 6; r++)
            for (int c = 0; c < 3; c++) blk->W[s][r * 3 + c] += w * (J[a][r] * Jp[c] + J[a][6 + r] * Jp[3 + c]);
        }
      }
    }
    if (h->pt_var[p] < 0) continue;
    const int pv = h->pt_var[p];
    memcpy(bpall + 3 * pv, blk->bp, sizeof(blk->bp));
    if (blk->nslot > slot_off[p + 1] - slot_off[p]) { fail = 1; continue; }
    slot_cnt[p] = blk->nslot;
    for (int s = 0; s < blk->nslot; s++) {
      slot_var[slot_off[p] + s] = blk->var[s];
      memcpy(slot_W + (size_t)(slot_off[p] + s) * 18, blk->W[s], sizeof(blk->W[s]));
    }
    if (solve_mode == 3) {
      memcpy(V_all + 9 * (size_t)pv, blk->V, sizeof(blk->V));
    } else if (solve_mode >= 1) {
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) Hfull[(size_t)(nc + 3 * pv + r) * dim + nc + 3 * pv + c] = blk->V[r * 3 + c];
      for (int s = 0; s < blk->nslot; s++)
        for (int r = 0; r < 6; r++)
          for (int c = 0; c < 3; c++) {
            Hfull[(size_t)(6 * blk->var[s] + r) * dim + nc + 3 * pv + c] = blk->W[s][r * 3 + c];
            Hfull[(size_t)(nc + 3 * pv + c) * dim + 6 * blk->var[s] + r] = blk->W[s][r * 3 + c];
          }
    } else {
      double Vl[9], Vi[9];
      memcpy(Vl, blk->V, sizeof(Vl));
      Vl[0] += lambda; Vl[4] += lambda; Vl[8] += lambda;
      if (inv3_sym(Vl, Vi)) { fail = 1; continue; }
      memcpy(Vinv_all + 9 * (size_t)pv, Vi, sizeof(Vi));
      /* Schur complement: S -= W Vinv W^T ; r -= W Vinv bp */
      for (int a = 0; a < blk->nslot; a++) {
        double Y[18];
        for (int r = 0; r < 6; r++)
          for (int c = 0; c < 3; c++)
            Y[r * 3 + c] = blk->W[a][r * 3 + 0] * Vi[0 * 3 + c] + blk->W[a][r * 3 + 1] * Vi[1 * 3 + c] + blk->W[a][r * 3 + 2] * Vi[2 * 3 + c];
        for (int r = 0; r < 6; r++)
          bc[6 * blk->var[a] + r] -= Y[r * 3 + 0] * blk->bp[0] + Y[r * 3 + 1] * blk->bp[1] + Y[r * 3 + 2] * blk->bp[2];
        for (int b2 = 0; b2 < blk->nslot; b2++)
          for (int r = 0; r < 6; r++)
            for (int c = 0; c < 6; c++)
              Hcc[(size_t)(6 * blk->var[a] + r) * nc + 6 * blk->var[b2] + c] -=
                  Y[r * 3 + 0] * blk->W[b2][c * 3 + 0] + Y[r * 3 + 1] * blk->W[b2][c * 3 + 1] + Y[r * 3 + 2] * blk->W[b2][c * 3 + 2];
      }
    }
  }

  int rc = 0;
  if (fail) rc = -1;
  if (!rc && solve_mode == 2) {
    /* [3P] SparseOptimizer::computeMarginals on the stored (undamped) Hessian: (H^-1)_pp(2,2) per point, dense.
       (H^-1)_ii = || L^-1 e_i ||^2 with H = L L^T.  Median (element [n/2]) as src/ChainBundle.cc:1420-1436. */
    for (int r = 0; r < nc; r++)
      for (int c = 0; c < nc; c++) Hfull[(size_t)r * dim + c] = Hcc[(size_t)r * nc + c];
    if (chol_dense(Hfull, dim)) rc = -1;
    else {
      double* cov = (double*)malloc(sizeof(double) * (size_t)(nptv + 1));
      double* y = (double*)malloc(sizeof(double) * (size_t)(dim + 1));
      for (int pv = 0; pv < nptv; pv++) {
        const int i0 = nc + 3 * pv + 2;
        double acc = 0;
        for (int i = i0; i < dim; i++) {
          double v = (i == i0) ? 1.0 : 0.0;
          for (int k = i0; k < i; k++) v -= Hfull[(size_t)i * dim + k] * y[k];
          y[i] = v / Hfull[(size_t)i * dim + i];
          acc += y[i] * y[i];
        }
        cov[pv] = acc;
      }
      qsort(cov, (size_t)nptv, sizeof(double), cmp_double);
      h->max_cov = nptv > 0 ? cov[nptv / 2] : DBL_MAX;
      free(cov); free(y);
    }
  } else if (!rc && solve_mode == 3) {
    /* block graph of the full damped system: npv pose blocks (6) then nptv point blocks (3) */
    const int nb = npv + nptv;
    int* bdim = (int*)malloc(sizeof(int) * (size_t)(nb + 1));
    int* boff = (int*)malloc(sizeof(int) * (size_t)(nb + 2));
    for (int b = 0; b < nb; b++) { bdim[b] = b < npv ? 6 : 3; boff[b] = b < npv ? 6 * b : nc + 3 * (b - npv); }
    boff[nb] = dim;
    fs_t f;
    memset(&f, 0, sizeof(f));
    f.nb = nb; f.bdim = bdim;
    size_t want = 4 * ((size_t)nb + (size_t)cap + (size_t)npv * npv) + 64;
    f.cap = 1; while (f.cap < want) f.cap <<= 1;
    f.tab = (fs_slot_t*)calloc(f.cap, sizeof(fs_slot_t));
    f.adj = (int**)calloc((size_t)nb + 1, sizeof(int*)); f.nadj = (int*)calloc((size_t)nb + 1, sizeof(int)); f.capadj = (int*)calloc((size_t)nb + 1, sizeof(int));
    for (int a = 0; a < npv; a++) {
      double* D = fs_get(&f, a, a, 1);
      for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) D[r * 6 + c] = Hcc[(size_t)(6 * a + r) * nc + 6 * a + c];
      for (int r = 0; r < 6; r++) D[r * 6 + r] += lambda;
      for (int b2 = a + 1; b2 < npv; b2++) {
        int any = 0;
        for (int r = 0; r < 6 && !any; r++) for (int c = 0; c < 6; c++) if (Hcc[(size_t)(6 * a + r) * nc + 6 * b2 + c] != 0.0) { any = 1; break; }
        if (!any) continue;
        int tr;
        double* O = fs_off(&f, a, b2, 1, &tr);
        for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) O[r * 6 + c] = Hcc[(size_t)(6 * a + r) * nc + 6 * b2 + c];
      }
    }
    for (int p = 0; p < h->n_pt; p++) {
      if (h->pt_var[p] < 0) continue;
      const int pv = h->pt_var[p], node = npv + pv;
      double* D = fs_get(&f, node, node, 1);
      memcpy(D, V_all + 9 * (size_t)pv, sizeof(double) * 9);
      D[0] += lambda; D[4] += lambda; D[8] += lambda;
      for (int s2 = 0; s2 < slot_cnt[p]; s2++) {
        int tr;
        double* O = fs_off(&f, slot_var[slot_off[p] + s2], node, 1, &tr);      /* pose index < point index: stored 6 x 3 as W */
        memcpy(O, slot_W + (size_t)(slot_off[p] + s2) * 18, sizeof(double) * 18);
      }
    }
    rc = full_sparse_solve(nb, bdim, boff, &f, h->b, h->x);
    for (int i = 0; i < f.n_chunk; i++) free(f.chunk[i]);
    free(f.chunk);
    for (int b = 0; b < nb; b++) free(f.adj[b]);
    free(f.tab); free(f.adj); free(f.nadj); free(f.capadj); free(bdim); free(boff);
  } else if (!rc && solve_mode == 1) {
    for (int r = 0; r < nc; r++)
      for (int c = 0; c < nc; c++) Hfull[(size_t)r * dim + c] = Hcc[(size_t)r * nc + c];
    for (int i = 0; i < dim; i++) Hfull[(size_t)i * dim + i] += lambda;
    memcpy(h->x, h->b, sizeof(double) * (size_t)dim);
    if (chol_dense(Hfull, dim)) rc = -1; else chol_solve(Hfull, dim, h->x);
  } else if (!rc) {
    /* note: bc currently holds the *reduced* rhs; keep the original b for computeScale */
    double* red = (double*)malloc(sizeof(double) * (size_t)(nc + 1));
    memcpy(red, bc, sizeof(double) * (size_t)nc);
    /* restore original pose rhs: recompute by adding back the Schur part */
    for (int p = 0; p < h->n_pt; p++) {
      if (h->pt_var[p] < 0) continue;
      const int pv = h->pt_var[p];
      const double* Vi = Vinv_all + 9 * (size_t)pv;
      const double* bp = bpall + 3 * pv;
      double t[3] = { Vi[0] * bp[0] + Vi[1] * bp[1] + Vi[2] * bp[2], Vi[3] * bp[0] + Vi[4] * bp[1] + Vi[5] * bp[2],
                      Vi[6] * bp[0] + Vi[7] * bp[1] + Vi[8] * bp[2] };
      for (int s = 0; s < slot_cnt[p]; s++) {
        const double* W = slot_W + (size_t)(slot_off[p] + s) * 18;
        const int v = slot_var[slot_off[p] + s];
        for (int r = 0; r < 6; r++) bc[6 * v + r] += W[r * 3 + 0] * t[0] + W[r * 3 + 1] * t[1] + W[r * 3 + 2] * t[2];
      }
    }
    for (int i = 0; i < nc; i++) Hcc[(size_t)i * nc + i] += lambda;
    if (nc > 0 && chol_dense(Hcc, nc)) rc = -1;
    if (!rc) {
      if (nc > 0) chol_solve(Hcc, nc, red);
      memcpy(h->x, red, sizeof(double) * (size_t)nc);
      for (int p = 0; p < h->n_pt; p++) {
        if (h->pt_var[p] < 0) continue;
        const int pv = h->pt_var[p];
        const double* Vi = Vinv_all + 9 * (size_t)pv;
        double t[3] = { bpall[3 * pv], bpall[3 * pv + 1], bpall[3 * pv + 2] };
        for (int s = 0; s < slot_cnt[p]; s++) {
          const double* W = slot_W + (size_t)(slot_off[p] + s) * 18;
          const double* dc = h->x + 6 * slot_var[slot_off[p] + s];
          for (int c = 0; c < 3; c++)
            for (int r = 0; r < 6; r++) t[c] -= W[r * 3 + c] * dc[r];
        }
        for (int r = 0; r < 3; r++) h->x[nc + 3 * pv + r] = Vi[r * 3 + 0] * t[0] + Vi[r * 3 + 1] * t[1] + Vi[r * 3 + 2] * t[2];
      }
    }
    free(red);
  }
  if (rc) memset(h->x, 0, sizeof(double) * (size_t)dim);
  free(Hcc); free(blk); free(Hfull); free(V_all); free(Vinv_all); free(slot_off); free(slot_var); free(slot_W); free(slot_cnt);
  return rc;
}

/* [3P] OptimizationAlgorithmLevenberg::computeLambdaInit needs max |H_jj| before damping */
static double max_diagonal(OraBa* h)
{
  double mx = 0;
  const int npv = h->n_pose_var;
  double* dpose = (double*)calloc((size_t)6 * npv + 1, sizeof(double));
  for (int p = 0; p < h->n_pt; p++) {
    double dpt[3] = { 0, 0, 0 };
    for (int q = h->pt_meas_off[p]; q < h->pt_meas_off[p + 1]; q++) {
      const int m = h->pt_meas_idx[q];
      double Jo[2][12], Js[2][12], Jp[6], rho[3] = { 0, 1, 0 };
      meas_jacobians(h, m, Jo, Js, Jp);
      if (h->use_robust) robustify(h, h->chi2[m], rho);
      const double w = rho[1] / sqrt(h->meas_noise[m]);
      int nv = 0, vars[4];
      double J[4][12];
      const int32_t* oc = h->meas_chain + 2 * m;
      const int32_t* sc = h->pt_chain + 2 * p;
      for (int side = 0; side < 2; side++)
        for (int i = 0; i < 2; i++) {
          const int id = side == 0 ? oc[i] : sc[i];
          if (id < 0 || h->pose_var[id] < 0) continue;
          const double* Jsrc = side == 0 ? Jo[i] : Js[i];
          int k;
          for (k = 0; k < nv; k++) if (vars[k] == h->pose_var[id]) break;
          if (k == nv) { vars[nv] = h->pose_var[id]; memset(J[nv], 0, sizeof(J[nv])); nv++; }
          for (int t = 0; t < 12; t++) J[k][t] += Jsrc[t];
        }
      for (int a = 0; a < nv; a++)
        for (int r = 0; r < 6; r++) dpose[6 * vars[a] + r] += w * (J[a][r] * J[a][r] + J[a][6 + r] * J[a][6 + r]);
      for (int r = 0; r < 3; r++) dpt[r] += w * (Jp[r] * Jp[r] + Jp[3 + r] * Jp[3 + r]);
    }
    if (h->pt_var[p] >= 0)
      for (int r = 0; r < 3; r++) if (fabs(dpt[r]) > mx) mx = fabs(dpt[r]);
  }
  for (int i = 0; i < 6 * npv; i++) if (fabs(dpose[i]) > mx) mx = fabs(dpose[i]);
  free(dpose);
  return mx;
}

static void apply_update(OraBa* h)
{
  const int nc = 6 * h->n_pose_var;
  for (int i = 0; i < h->n_pose; i++)
    if (h->pose_var[i] >= 0) pose_oplus(&h->pose[i], h->x + 6 * h->pose_var[i]);
  for (int p = 0; p < h->n_pt; p++)
    if (h->pt_var[p] >= 0) point_oplus(h->pt + 3 * p, h->x + nc + 3 * h->pt_var[p]);
}

int ora_ba_lm_step(OraBa* h, double lambda, double sigma_sq, int solve_mode, double* delta,
                   double* sigma_sq_used, double* rchi2)
{
  compute_errors(h);
  if (sigma_sq < 0) h->recompute_sigma = 1;
  else { h->recompute_sigma = 0; h->sigma_sq_raw = sigma_sq; h->sigma_sq_lim = sigma_sq < 0.25 ? 0.25 : sigma_sq; h->sigma_lim = sqrt(h->sigma_sq_lim); }
  const double chi = robust_chi2(h);
  if (rchi2) *rchi2 = chi;
  if (sigma_sq_used) *sigma_sq_used = h->sigma_sq_raw;
  const int rc = build_and_solve(h, lambda, solve_mode);
  const int dim = 6 * h->n_pose_var + 3 * h->n_pt_var;
  if (delta) memcpy(delta, h->x, sizeof(double) * (size_t)dim);
  return rc;
}

/* ChainBundle::Compute, src/ChainBundle.cc:1305-1451 around [3P] SparseOptimizer::optimize and
   OptimizationAlgorithmLevenberg::solve */
int ora_ba_compute(OraBa* h, volatile const uint8_t* abort_ext, int n_iter, double user_lambda,
                   int solve_mode, OraBaStats* st)
{
  const int npv = h->n_pose_var, nptv = h->n_pt_var;
  const int dim = 6 * npv + 3 * nptv;
  uint8_t local_abort = 0;      /* the reference shares one flag: external requests and convergence */
  int conv_mag = 0, conv_res = 0;
  h->n_outliers = 0;
  h->hit_max = 0; h->converged = 0;
  memset(st, 0, sizeof(*st));

  compute_errors(h);            /* :1318-1320 */
  h->recompute_sigma = 1;
  st->chi2_before = robust_chi2(h);

  h->total_trials = 0;
  int counter = 0;
  int ok = 1;
  se3_t* pose_bak = (se3_t*)malloc(sizeof(se3_t) * (size_t)(h->n_pose > 0 ? h->n_pose : 1));
  double* pt_bak = (double*)malloc(sizeof(double) * 3 * (size_t)(h->n_pt > 0 ? h->n_pt : 1));

  for (int it = 0; it < n_iter && !(local_abort || (abort_ext && *abort_ext)) && ok; it++) {
    h->recompute_sigma = 1;                                   /* preIteration: UpdateSigmaSquaredAction */
    compute_errors(h);
    double current_chi = robust_chi2(h);
    double temp_chi = current_chi;
    if (it == 0) {
      h->lambda = user_lambda > 0 ? user_lambda : 1e-5 * max_diagonal(h);
      h->ni = 2;
    }
    double rho = 0;
    int qmax = 0;
    do {
      memcpy(pose_bak, h->pose, sizeof(se3_t) * (size_t)h->n_pose);   /* push */
      memcpy(pt_bak, h->pt, sizeof(double) * 3 * (size_t)h->n_pt);
      /* the linearisation point is the pushed state, errors must be those of that state */
      if (qmax > 0) { compute_errors(h); }
      const int ok2 = (build_and_solve(h, h->lambda, solve_mode) == 0);
      apply_update(h);
      compute_errors(h);
      temp_chi = robust_chi2(h);
      if (!ok2) temp_chi = DBL_MAX;
      rho = current_chi - temp_chi;
      double scale = 0;
      for (int j = 0; j < dim; j++) scale += h->x[j] * (h->lambda * h->x[j] + h->b[j]);
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && isfinite(temp_chi)) {
        double alpha = 1. - pow((2 * rho - 1), 3);
        alpha = alpha < (2. / 3.) ? alpha : (2. / 3.);
        const double sf = alpha > (1. / 3.) ? alpha : (1. / 3.);
        h->lambda *= sf;
        h->ni = 2;
        current_chi = temp_chi;
      } else {
        h->lambda *= h->ni;
        h->ni *= 2;
        memcpy(h->pose, pose_bak, sizeof(se3_t) * (size_t)h->n_pose);  /* pop */
        memcpy(h->pt, pt_bak, sizeof(double) * 3 * (size_t)h->n_pt);
      }
      qmax++;
    } while (rho < 0 && qmax < 100 && !(local_abort || (abort_ext && *abort_ext)));
    if (qmax == 100 || rho == 0) ok = 0;                      /* Terminate */
    ++counter;
    /* postIteration actions */
    {                                                         /* CheckConvergedUpdateMagAction :1009-1047 */
      double ss = 0;
      for (int i = 0; i < dim; i++) ss += h->x[i] * h->x[i];
      const double rms = sqrt(ss / dim);
      if (rms < 1e-10) { conv_mag = 1; local_abort = 1; }
    }
    {                                                         /* CheckConvergedResidualAction :1091-1118 */
      const double cur = robust_chi2(h);
      const double pct = (h->last_chi2 - cur) / h->last_chi2;
      if (pct >= 0 && pct <= 1e-10) { conv_res = 1; local_abort = 1; }
      else if (cur == 0) { conv_res = 1; local_abort = 1; }
      h->last_chi2 = cur;
    }
    h->total_trials += qmax;                                  /* UpdateTotalIterationsAction */
  }
  /* marginals (:1401-1448): only with <3 movable poses, on the Hessian of the LAST buildSystem, i.e. the state the
     final iteration was linearised at (pose_bak / pt_bak), with that iteration's robust weights */
  h->max_cov = 0;
  if (counter > 0 && npv < 3) {
    if (nptv == 0) h->max_cov = DBL_MAX;
    else if (dim <= 2500) {
      se3_t* pose_fin = (se3_t*)malloc(sizeof(se3_t) * (size_t)h->n_pose);
      double* pt_fin = (double*)malloc(sizeof(double) * 3 * (size_t)h->n_pt);
      memcpy(pose_fin, h->pose, sizeof(se3_t) * (size_t)h->n_pose); memcpy(pt_fin, h->pt, sizeof(double) * 3 * (size_t)h->n_pt);
      memcpy(h->pose, pose_bak, sizeof(se3_t) * (size_t)h->n_pose); memcpy(h->pt, pt_bak, sizeof(double) * 3 * (size_t)h->n_pt);
      h->recompute_sigma = 0;
      compute_errors(h);
      if (build_and_solve(h, 0.0, 2)) h->max_cov = 0;         /* computeMarginals() failed */
      memcpy(h->pose, pose_fin, sizeof(se3_t) * (size_t)h->n_pose); memcpy(h->pt, pt_fin, sizeof(double) * 3 * (size_t)h->n_pt);
      free(pose_fin); free(pt_fin);
    }
  }
  free(pose_bak); free(pt_bak);

  h->hit_max = (counter == n_iter);
  compute_errors(h);                                          /* :1340-1343 */
  h->recompute_sigma = 1;
  st->chi2_after = robust_chi2(h);
  h->converged = (conv_mag || conv_res);
  const int ext_abort_flag = (local_abort || (abort_ext && *abort_ext));
  const int external_abort = ext_abort_flag && !h->converged;

  st->iterations = counter;
  st->total_trials = h->total_trials;
  st->converged = h->converged;
  st->hit_max_iter = h->hit_max;
  st->sigma_sq = h->sigma_sq_raw;
  st->lambda = h->lambda;
  st->mean_chi2 = h->n_meas ? st->chi2_after / h->n_meas : 0;
  st->max_cov = h->max_cov;
  if (counter == 0 && !external_abort) return -1;
  if (counter == 0 && ext_abort_flag) return 0;

  if (h->use_tukey) {                                         /* :1368-1399 */
    double* a = (double*)malloc(sizeof(double) * (size_t)(h->n_meas > 0 ? h->n_meas : 1));
    for (int m = 0; m < h->n_meas; m++) a[m] = fabs(h->chi2[m]);
    double ts = ora_tukey_sigma_sq(a, h->n_meas);
    if (ts < 0.25) ts = 0.25;
    for (int m = 0; m < h->n_meas; m++) {
      /* Tukey::Weight == 0  <=>  SquareRootWeight == 0  (MEstimator.h:84-96) */
      double sq = a[m] > ts ? 0.0 : 1.0 - (a[m] / ts);
      if (sq * sq == 0) h->outliers[h->n_outliers++] = m;
    }
    free(a);
  }
  st->max_cov = h->max_cov;                                   /* computed above, before the AFTER block */
  st->n_outliers = h->n_outliers;
  return counter;
}
