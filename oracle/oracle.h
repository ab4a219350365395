/*
 * oracle.h — CPU restatement of the mcptam hot paths.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under oracle/ is part of the product.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library, and only
 * as the checker / the timed CPU baseline.  The product path (mcptam_b200/csrc) never
 * links, loads or calls it.
 *
 * PARITY: PINNED AGAINST THE REFERENCE'S OWN CODE for everything the reference itself implements.  The reference
 * (aharmat/mcptam @ ae54e1b) ships no tests, golden vectors or fixtures, and its build (ROS, TooN, libCVD, g2o,
 * SuiteSparse) is not available here -- but its hot-path translation units compile, UNMODIFIED and where they lie under
 * /root/reference, against minimal stand-ins for the third-party headers (oracle/ref_shim/, oracle/build_ref.py ->
 * the .so files under oracle/_ref): include/mcptam/MEstimator.h, LevelHelpers.h, SmallMatrixOpts.h, src/ShiTomasi.cc, src/MiniPatch.cc,
 * src/TaylorCamera.cc, src/PatchFinder.cc (SSE and scalar ZMSSD) and src/ChainBundle.cc (vertices, edges, pose-chain
 * helpers, adaptive Huber kernel, convergence actions, Compute with its Tukey pass).  tests/test_oracle_vs_ref.py holds
 * the oracle to them: integers and same-order fp64 bit-for-bit, the rest to 1e-12 .. 1e-7 (stated per test).
 * STILL [3P]-FROM-MEMORY (behaviour of code that is not in the reference tree, restated from the published algorithms,
 * in the oracle AND in the stand-ins): g2o's Levenberg-Marquardt loop and linear solve, CHOLMOD, libCVD's halfSample /
 * fast_corner_detect_10 / fast_corner_score_10 / fast_nonmax / transform / sample, TooN's SE3 / Cholesky / SVD, Eigen's
 * polynomial solver.  Every function below cites the reference file:line it restates.
 *
 * What the oracle is ADDITIONALLY pinned against (tests/test_oracle_cpu.py, tests/test_host_cpu.py, tests/test_epipolar.py):
 *   - the reference's own validation device: analytic vs central-difference Jacobians (src/ChainBundle.cc:688-740);
 *   - itself, through different algorithms: Schur vs dense full-system solve, closed-form vs bisection FAST score,
 *     brute-force vs fast detector, marginal covariances vs a numpy inverse, fast_nonmax vs its definition;
 *   - independent code: OpenCV's FAST 9_16 detector (ring / strictness / border semantics) and matchTemplate
 *     (MiniPatch SSD exactly, PatchFinder ZMSSD up to its integer truncation), scipy's
 *     least_squares (the state the LM driver converges to is the least-squares optimum of the same residuals),
 *     numpy (one LM step = the solution of the Huber-weighted normal equations assembled outside the oracle; SE3 exp
 *     = scipy's matrix exponential of the twist; Shi-Tomasi = cv2.cornerMinEigenVal up to the documented scale),
 *     and the separately written C++ TaylorCamera mirror (inverse-polynomial fit, projection, derivatives);
 *   - committed golden vectors (tests/golden/), which guard against regressions of the oracle itself.
 * The build that CHECKS is -O2 -ffp-contract=off (Makefile); bench.py TIMES a -O3 -march=native build of the same
 * sources (MCP_ORACLE_FAST=1, oracle.py), which is never used as a checker.
 */
#ifndef MCPTAM_ORACLE_H
#define MCPTAM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same memory layout as McpTaylorCam in include/mcptam_b200.h (kept separate on purpose). */
typedef struct OraTaylorCam {
  double poly[5];       /* a0, 0, a2, a3, a4            src/TaylorCamera.cc:98-102 */
  double center[2];     /* mv2Center                    src/TaylorCamera.cc:130-131 */
  double affine[4];     /* mm2Affine row-major          src/TaylorCamera.cc:183-186 */
  double image_size[2]; /* mv2ImageSize */
  double min_theta;     /* mdMinTheta                   src/TaylorCamera.cc:152 */
  double theta_mean;    /* mdThetaMean                  src/TaylorCamera.cc:559 */
  double theta_std;     /* mdThetaStd                   src/TaylorCamera.cc:563 */
  int32_t n_inv;        /* number of inverse-poly coefficients (degree+1) */
  int32_t pad_;
  double inv_poly[32];  /* mvxPolyInvCoeffs, coefficient of x^0 first */
} OraTaylorCam;

typedef struct OraBaStats {
  int32_t iterations;        /* outer LM iterations run (return value of Compute) */
  int32_t total_trials;      /* mnTotalIterations: sum of levenbergIterations */
  int32_t converged;         /* mbConverged */
  int32_t hit_max_iter;      /* mbHitMaxIterations */
  int32_t n_outliers;
  int32_t pad_;
  double sigma_sq;           /* GetSigmaSquared(): raw Huber sigma^2 of the final state */
  double mean_chi2;          /* GetMeanChiSquared() */
  double lambda;             /* GetLambda() */
  double max_cov;            /* GetMaxCov() */
  double chi2_before;        /* robust chi2 before optimisation */
  double chi2_after;         /* robust chi2 after optimisation  */
} OraBaStats;

typedef struct OraBa OraBa;

/* ---- bundle adjustment ---------------------------------------------------------- */
OraBa* ora_ba_create(int use_robust, int use_tukey);
void ora_ba_destroy(OraBa* h);
int ora_ba_set_cameras(OraBa* h, int n_cam, const OraTaylorCam* cams);
/* pose_Rt: 12 doubles per pose (row-major R, then t).  chains: 2 ids each, -1 pad.
   ids index the pose array (0-based).  meas_noise = dNoiseSigmaSquared of AddMeas. */
int ora_ba_load(OraBa* h, int n_pose, const double* pose_Rt, const uint8_t* pose_fixed,
                int n_pt, const double* pt_xyz, const int32_t* pt_chain, const uint8_t* pt_fixed,
                int n_meas, const double* meas_xy, const int32_t* meas_chain,
                const int32_t* meas_pt, const double* meas_noise, const int32_t* meas_cam);
/* solve_mode: 0 = per-point Schur complement + dense pose solve, 1 = dense full system */
int ora_ba_compute(OraBa* h, volatile const uint8_t* abort_flag, int n_iter, double user_lambda,
                   int solve_mode, OraBaStats* stats);
int ora_ba_get_poses(const OraBa* h, double* pose_Rt);
int ora_ba_get_points(const OraBa* h, double* pt_xyz);
int ora_ba_set_poses(OraBa* h, const double* pose_Rt);
int ora_ba_set_points(OraBa* h, const double* pt_xyz);
int ora_ba_get_outliers(const OraBa* h, int32_t* meas_idx, int cap);

/* building blocks exposed for unit tests */
int ora_ba_eval(OraBa* h, double* err_xy /*2*n_meas*/, double* chi2 /*n_meas*/);
/* Jacobians of one measurement: J_obs[2][12]  (chain link 0 then 1, 2x6 each, row-major 2x6),
   J_src[2][12], J_pt[6] (2x3).  Links that are fixed / absent are returned as zeros. */
int ora_ba_jacobians(OraBa* h, int meas, double* J_obs, double* J_src, double* J_pt);
int ora_ba_oplus_pose(OraBa* h, int pose_id, const double* d6);
int ora_ba_oplus_point(OraBa* h, int pt_id, const double* d3);
/* One LM trial from the current state with given lambda and Huber sigma^2 (sigma_sq<0: recompute).
   Writes the update (movable poses in id order, 6 each; then movable points in id order, 3 each)
   to delta (does not apply it).  Returns <0 if the solve failed. */
int ora_ba_lm_step(OraBa* h, double lambda, double sigma_sq, int solve_mode, double* delta,
                   double* sigma_sq_used, double* robust_chi2);
double ora_huber_sigma_sq(const double* abs_chi2, int n);
double ora_tukey_sigma_sq(const double* abs_chi2, int n);

/* camera model */
int ora_cam_project(const OraTaylorCam* cam, const double* p3, double* px2, double* derivs4);
void ora_cam_sphere_deriv(const double* p3, double* dtheta3, double* dphi3);
void ora_cam_unproject(const OraTaylorCam* cam, const double* px2, double* ray3);
void ora_se3_exp(const double* mu6, double* Rt12);
void ora_so3_exp(const double* w3, double* R9);

/* ---- front end (oracle/fe_oracle.c) ------------------------------------------- */
void ora_halfsample(const uint8_t* in, int w, int h, int in_stride, uint8_t* out, int out_stride);
/* FAST-10 detection at threshold b (raster order).  Returns number of corners, writes up to cap. */
int ora_fast10_detect(const uint8_t* im, int w, int h, int stride, int b, int32_t* xy, int cap);
/* Definition-level (brute force over all 16 arcs) detector used to pin the fast detector above */
int ora_fast10_detect_bruteforce(const uint8_t* im, int w, int h, int stride, int b, int32_t* xy, int cap);
/* generic N-of-16 brute force (N=9 is cross-checked against cv2 in the CPU tests) */
int ora_fastN_detect_bruteforce(const uint8_t* im, int w, int h, int stride, int b, int n_arc, int32_t* xy, int cap);
void ora_fast10_score(const uint8_t* im, int stride, const int32_t* xy, int n, int b, int32_t* scores);
void ora_fast10_score_bisect(const uint8_t* im, int stride, const int32_t* xy, int n, int b, int32_t* scores);
/* One pyramid level of KeyFrame::MakeKeyFrame_Lite: detect(5)+score, histogram, adaptive
   threshold, mask filter, row LUT.  Returns number of kept corners. */
int ora_level_corners(const uint8_t* im, int w, int h, int stride, const uint8_t* mask, int mask_stride,
                      int adaptive, int fixed_thresh, int32_t* xy, int cap, int32_t* fast_freq31,
                      int32_t* fast_thresh, int32_t* row_lut);
double ora_shitomasi(const uint8_t* im, int stride, int half_box, int x, int y);

typedef struct OraPatchReq {
  int32_t src_level_w, src_level_h;   /* unused by oracle calls taking explicit images */
  double warp_inv[4];                 /* mm2WarpInverse, row-major (CalcSearchLevelAndWarpMatrix) */
  int32_t src_cx, src_cy;             /* point.mirCenter in the source level */
  int32_t search_level;               /* mnSearchLevel */
  int32_t pred_x, pred_y;             /* ir(td.mv2Image): predicted L0 position */
  int32_t range;                      /* nRange (L0 pixels) */
  int32_t subpix_its;                 /* 0: none */
  int32_t exhaustive;
} OraPatchReq;

/* 8x8 warped template via CVD::transform restatement.  Returns nOutside. */
int ora_patch_template(const uint8_t* src, int w, int h, int stride, const double* m2 /*2x2*/,
                       double cx, double cy, uint8_t* templ64);
int ora_zmssd(const uint8_t* im, int w, int h, int stride, const uint8_t* templ64, int tsum, int tsumsq,
              int x, int y, int max_ssd);
/* FindPatchCoarse: returns found flag; best position (level coords) and score */
int ora_find_patch_coarse(const uint8_t* im, int w, int h, int stride, const int32_t* corners_xy,
                          int n_corners, const int32_t* row_lut, const uint8_t* templ64, int level,
                          int pred_x, int pred_y, int range, int exhaustive, int32_t* best_xy,
                          int32_t* score);
/* MakeSubPixTemplate + IterateSubPixToConvergence.  pos_io: L0 position in/out. returns converged */
int ora_subpix(const uint8_t* im, int w, int h, int stride, const uint8_t* templ64, int level,
               double* pos_io, int max_its);
int ora_minipatch_ssd(const uint8_t* im, int w, int h, int stride, const uint8_t* patch81, int x, int y);
int ora_minipatch_find(const uint8_t* im, int w, int h, int stride, const uint8_t* patch81,
                       const int32_t* corners_xy, int n_corners, const int32_t* row_lut_or_null,
                       int n_lut, int range, int32_t* pos_io);

/* KeyFrame::MakeKeyFrame_Rest candidate generation (src/KeyFrame.cc:363-531); fast_nonmax restated from libCVD [3P] */
int ora_fast_old_score(const uint8_t* im, int stride, int x, int y, int barrier);
int ora_fast_nonmax(const uint8_t* im, int w, int h, int stride, const int32_t* cxy, int nc, int barrier, int strict, uint8_t* keep);
int ora_keyframe_rest_level(const uint8_t* im, int w, int h, int stride, const int32_t* cxy, int nc, const int32_t* lut, int fast_thresh,
                            int use_shi, int use_thresh, double top_fraction, double thresh, int nonmax_strict,
                            const uint8_t* prev_im, const int32_t* prev_cxy, int prev_nc, const int32_t* prev_lut, int n_prev,
                            int32_t* out_xy, double* out_score, int cap, int32_t* n_max);

int ora_search_patches_batch(const uint8_t* const* src_pyr, const uint8_t* const* tgt_pyr, const int* widths, const int* heights,
                             const int32_t* const* corners, const int* n_corners, const int32_t* const* luts, int n,
                             const int32_t* req_i, const double* m2, double* found_xy, int32_t* found_flag);
int ora_calc_jacobian(const OraTaylorCam* cam, const double* base_Rt, const double* cfb_Rt, const double* pw, double* px2, double* derivs4, double* J12);
int ora_pose_update(int n, const double* found_xy, const double* image_xy, const double* sqrt_inv_noise, const double* jac12,
                    const int32_t* found, int estimator, double override_sigma, double* mu6, double* sigma_sq_out, int32_t* outlier);
int ora_project_point(const OraTaylorCam* cam, const double* pose_Rt, const double* pw, const double* right_w, const double* down_w,
                      double* px2, double* derivs4, double* warp_inv4, double* v3cam, int* in_image);

#ifdef __cplusplus
}
#endif
#endif
