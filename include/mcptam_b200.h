/*
 * mcptam_b200.h — C ABI of the B200-native replacement for MCPTAM's two data-parallel hot paths.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types, never throws.
 * Every entry point returns an int status (MCP_OK = 0, <0 = error; mcp_ba_compute returns the
 * ChainBundle::Compute convention).  A handle owns its device memory and one CUDA stream, and is
 * used by exactly one host thread (the MapMaker thread for BA, the Tracker thread for the front
 * end); the BA abort flag may be written by another thread exactly as in the reference.
 *
 * Reference interfaces replaced (file:line in aharmat/mcptam @ ae54e1b):
 *   mcp_ba_create/destroy   <- ChainBundle::ChainBundle / ~ChainBundle   src/ChainBundle.cc:1139-1195
 *   mcp_ba_set_cameras      <- TaylorCameraMap& ctor argument            include/mcptam/ChainBundle.h:106
 *   mcp_ba_load             <- AddPose / AddPoint / AddMeas (batched)    src/ChainBundle.cc:1198-1281
 *   mcp_ba_compute          <- ChainBundle::Compute                      src/ChainBundle.cc:1305-1451
 *   mcp_ba_get_poses/points <- GetPose / GetPoint                        src/ChainBundle.cc:1453-1463
 *   mcp_ba_get_outliers     <- GetOutlierMeasurements                    src/ChainBundle.cc:1465-1468
 *   McpBaStats fields       <- Converged/TotalIterations/GetSigmaSquared/GetMeanChiSquared/
 *                              GetMaxCov/GetLambda                       include/mcptam/ChainBundle.h:150-179
 *   mcp_fe_make_keyframe    <- KeyFrame::MakeKeyFrame_Lite               src/KeyFrame.cc:145-361
 *   mcp_fe_search_patches   <- Tracker::SearchForPoints loop body        src/Tracker.cc:1299-1377
 *                              (PatchFinder::MakeTemplateCoarseCont, FindPatchCoarse,
 *                               MakeSubPixTemplate, IterateSubPixToConvergence; src/PatchFinder.cc)
 *   mcp_fe_shitomasi        <- FindShiTomasiScoreAtPoint                 src/ShiTomasi.cc:34-63
 *   mcp_fe_minipatch_find   <- MiniPatch::FindPatch / SSDAtPoint         src/MiniPatch.cc:34-113
 */
#ifndef MCPTAM_B200_H
#define MCPTAM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCP_OK 0
#define MCP_ERR_INVALID (-101)      /* bad argument / index out of range */
#define MCP_ERR_CUDA (-102)         /* CUDA runtime error, see mcp_last_error() */
#define MCP_ERR_UNSUPPORTED (-103)  /* e.g. movable second chain link (calibration BA), chain length > 2 */
#define MCP_ERR_NO_DEVICE (-104)    /* no CUDA device: there is no CPU fallback */
#define MCP_ERR_NCCL (-105)
#define MCP_ERR_STATE (-106)        /* called out of order (e.g. compute before load) */

/* Thread-local description of the last error returned on this thread. */
const char* mcp_last_error(void);
/* ABI version of this library (bumped on any signature change). */
int mcp_abi_version(void);

/* ------------------------------------------------------------------------------------------
 * Taylor (Scaramuzza) camera model, all derived quantities precomputed on the host exactly as
 * TaylorCamera::RefreshParams does (src/TaylorCamera.cc:84-198).
 * ------------------------------------------------------------------------------------------ */
typedef struct McpTaylorCam {
  double poly[5];       /* mv5PolyCoeffs: a0, 0, a2, a3, a4 */
  double center[2];     /* mv2Center */
  double affine[4];     /* mm2Affine, row-major */
  double image_size[2]; /* mv2ImageSize */
  double min_theta;     /* mdMinTheta */
  double theta_mean;    /* mdThetaMean */
  double theta_std;     /* mdThetaStd */
  int32_t n_inv;        /* number of inverse polynomial coefficients (<= 31, MAX_INV_DEGREE+1) */
  int32_t pad_;
  double inv_poly[32];  /* mvxPolyInvCoeffs, coefficient of x^0 first */
} McpTaylorCam;

/* ------------------------------------------------------------------------------------------
 * Bundle adjustment (ChainBundle)
 * ------------------------------------------------------------------------------------------ */
typedef struct McpBa McpBa;

typedef struct McpBaConfig {
  int32_t use_robust;                /* ChainBundle ctor bUseRobust */
  int32_t use_tukey;                 /* ChainBundle ctor bUseTukey */
  int32_t verbose;
  int32_t max_trials_after_failure;  /* ChainBundle::snMaxTrialsAfterFailure, default 100 */
  double update_pct_limit;           /* sdUpdatePercentConvergenceLimit, default 1e-10 */
  double update_rms_limit;           /* sdUpdateRMSConvergenceLimit, default 1e-10 */
  double min_sigma;                  /* sdMinMEstimatorSigma, default 0.5 */
  int32_t device;                    /* CUDA device ordinal, -1 = current */
  int32_t pad_;
} McpBaConfig;

typedef struct McpBaStats {
  int32_t iterations;    /* outer LM iterations run (the Compute return value when >= 0) */
  int32_t total_trials;  /* ChainBundle::TotalIterations() */
  int32_t converged;     /* ChainBundle::Converged() */
  int32_t hit_max_iter;
  int32_t n_outliers;
  int32_t pad_;
  double sigma_sq;       /* GetSigmaSquared() */
  double mean_chi2;      /* GetMeanChiSquared() */
  double lambda;         /* GetLambda() */
  double max_cov;        /* GetMaxCov() */
  double chi2_before;
  double chi2_after;
  double gpu_ms;         /* device time of the optimise loop (CUDA events on the handle's stream) */
  int32_t kernel_launches; /* kernels launched by this call */
  int32_t pad2_;
} McpBaStats;

void mcp_ba_default_config(McpBaConfig* cfg);
int mcp_ba_create(const McpBaConfig* cfg, McpBa** out);
int mcp_ba_destroy(McpBa* h);
int mcp_ba_set_cameras(McpBa* h, int32_t n_cam, const McpTaylorCam* cams);

/* Batched AddPose/AddPoint/AddMeas.  All arrays are caller-owned host memory, copied to the device.
 *   pose_Rt     n_pose x 12 doubles: row-major rotation then translation (TooN::SE3 layout)
 *   pose_fixed  n_pose bytes
 *   pt_xyz      n_pt x 3: v3PointInCam, the point in its source camera frame
 *   pt_chain    n_pt x 2 pose indices (0-based into pose_*): [MKF pose, cam-from-base pose], -1 pads
 *               a 1-link chain (fixed points use [world pose, -1])
 *   meas_chain  n_meas x 2: chain of the observing camera
 *   meas_noise  n_meas: dNoiseSigmaSquared (= LevelScale^2, src/BundleAdjusterMulti.cc:196)
 *   meas_cam    n_meas: index into the camera array (the reference's camera-name string)
 * Only chain link 0 may be movable (the hot path: BundleAdjusterMulti/Single); a movable link 1
 * (BundleAdjusterCalib) returns MCP_ERR_UNSUPPORTED.
 * The caller's arrays have been consumed when the call returns (they may be reused at once), but the
 * device may still be working on the load: every later call on the handle is ordered behind it, and a
 * device-side failure of the load surfaces from that later call (MCP_BA_LOAD_SYNC=1 makes the load wait). */
int mcp_ba_load(McpBa* h, int32_t n_pose, const double* pose_Rt, const uint8_t* pose_fixed,
                int32_t n_pt, const double* pt_xyz, const int32_t* pt_chain, const uint8_t* pt_fixed,
                int32_t n_meas, const double* meas_xy, const int32_t* meas_chain,
                const int32_t* meas_pt, const double* meas_noise, const int32_t* meas_cam);

/* Returns the number of LM iterations run (>0), 0 if aborted before the first step, <0 on failure
 * (-1: the reference's "map is probably corrupt"; other negatives are MCP_ERR_*).
 * abort_flag may be NULL; it is polled between LM trials. */
int mcp_ba_compute(McpBa* h, volatile const uint8_t* abort_flag, int32_t n_iter, double user_lambda,
                   McpBaStats* stats);
int mcp_ba_get_poses(McpBa* h, double* pose_Rt);
int mcp_ba_get_points(McpBa* h, double* pt_xyz);
/* Writes up to cap measurement indices (caller's original order, ascending); returns the total count. */
int mcp_ba_get_outliers(McpBa* h, int32_t* meas_idx, int32_t cap);
/* Overwrite the current estimate (poses and points) without re-marshalling the graph. */
int mcp_ba_set_state(McpBa* h, const double* pose_Rt, const double* pt_xyz);
/* Restore the estimate given to mcp_ba_load (device-to-device) and forget LM history. */
int mcp_ba_reset_state(McpBa* h);

/* Multi-GPU: every rank loads the same problem; map points (with all their measurements) are
 * partitioned across ranks, the Schur-reduced camera system is all-reduced with NCCL.
 * nccl_unique_id is the 128-byte ncclUniqueId created on rank 0. */
int mcp_ba_comm_init(McpBa* h, const void* nccl_unique_id, int32_t rank, int32_t world);
int mcp_nccl_unique_id(void* out128);
/* The point partition used for `world` ranks: part_pt[world+1] boundaries (pure host code, no GPU needed). */
int mcp_ba_partition(int32_t n_pt, const int32_t* meas_pt, int32_t n_meas, int32_t world, int32_t* part_pt);

/* Test / diagnostic hooks (not part of the reference surface). */
int mcp_ba_eval(McpBa* h, double* err_xy, double* chi2);                 /* original measurement order */
int mcp_ba_debug_jacobians(McpBa* h, double* J30);                       /* per measurement: Jobs(12) Jsrc(12) Jpt(6) */
int mcp_ba_lm_step(McpBa* h, double lambda, double sigma_sq, double* delta, double* sigma_sq_used,
                   double* robust_chi2);                                 /* one trial, update returned not applied */
typedef struct McpBaTiming {   /* accumulated over the last mcp_ba_compute, milliseconds / counts */
  double ms_select, ms_linearize, ms_schur, ms_solve, ms_backsub, ms_control, ms_other;
  int32_t n_select, n_linearize, n_schur, n_solve, n_backsub, n_control, n_other, pad_;
} McpBaTiming;
/* out == NULL arms a per-task timestamp trace of the dense solver; a later call with out != NULL reads it. */
int mcp_ba_debug_solve_trace(McpBa* h, double* out, int32_t cap_doubles);
/* The CUDA stream (cudaStream_t) all work of this handle is enqueued on, for external event timing. */
int mcp_ba_get_stream(McpBa* h, void** cuda_stream);
int mcp_ba_set_profiling(McpBa* h, int32_t enable);   /* per-kernel CUDA events (adds sync overhead) */
int mcp_ba_get_timing(McpBa* h, McpBaTiming* out);

/* ------------------------------------------------------------------------------------------
 * Front end (KeyFrame::MakeKeyFrame_Lite + PatchFinder)
 * ------------------------------------------------------------------------------------------ */
#define MCP_LEVELS 4            /* include/mcptam/KeyFrame.h:85 */
#define MCP_MIN_FAST_THRESH 5   /* :88 */
#define MCP_MAX_FAST_THRESH 30  /* :89 */

typedef struct McpFe McpFe;

typedef struct McpFeConfig {
  int32_t width, height;        /* level-0 image size */
  int32_t adaptive_thresh;      /* KeyFrame::sbAdaptiveThresh (default 1) */
  int32_t max_corners_per_level;/* capacity of the per-level corner buffers (level 0; halves per level) */
  int32_t max_keyframes;        /* resident keyframe pyramids (current + map keyframes used as patch sources) */
  int32_t max_patches;          /* capacity of one mcp_fe_search_patches call */
  int32_t device;
  int32_t halfsample_round;     /* 0: truncating mean (CVD generic template, default) 1: round to nearest */
  int32_t transform_round;      /* 0: truncating float->byte in CVD::sample (default) 1: +0.5 */
  int32_t pad_;
} McpFeConfig;

typedef struct McpLevelOut {
  int32_t width, height;
  int32_t n_corners;            /* after threshold + mask filtering; raster order */
  int32_t fast_thresh;          /* Level::nFastThresh */
  int32_t fast_freq[31];        /* Level::vFastFrequency[0..30] */
  int32_t n_corners_total;      /* corners the level really has; > n_corners means max_corners_per_level truncated the list (raise it) */
  uint8_t* image;               /* optional host out: width*height bytes, may be NULL */
  int32_t* corners_xy;          /* optional host out: 2*cap ints, may be NULL */
  int32_t corners_cap;
  int32_t pad2_;
  int32_t* row_lut;             /* optional host out: height ints (Level::vCornerRowLUT), may be NULL */
  uint8_t* last_mask;           /* optional host out: width*height bytes, Level::lastMask (src/KeyFrame.cc:242), may be NULL */
} McpLevelOut;

typedef struct McpPatchReq {
  int32_t src_kf;               /* resident keyframe slot holding point.mpPatchSourceKF's pyramid */
  int32_t src_level;            /* point.mnSourceLevel */
  int32_t src_cx, src_cy;       /* point.mirCenter */
  double warp_inv[4];           /* PatchFinder::mm2WarpInverse after CalcSearchLevelAndWarpMatrix */
  int32_t search_level;         /* PatchFinder::mnSearchLevel (GetLevel()) */
  int32_t pred_x, pred_y;       /* CVD::ir(td.mv2Image) */
  int32_t range;                /* nRange, level-0 pixels */
  int32_t subpix_its;           /* nSubPixIts (0 = none) */
  int32_t exhaustive;           /* bExhaustive || point.mbFixed; 2: no coarse search -- pred_x / pred_y are a coarse match
                                   (irBest, search-level coordinates) and only the sub-pixel iteration runs from it
                                   (SetSubPixPos + IterateSubPixToConvergence, src/MapMakerServerBase.cc:832-846) */
} McpPatchReq;

typedef struct McpPatchRes {
  int32_t template_bad;         /* finder.TemplateBad() */
  int32_t found;                /* td.mbFound */
  int32_t did_subpix;           /* td.mbDidSubPix */
  int32_t score;                /* nScore of FindPatchCoarse */
  int32_t coarse_x, coarse_y;   /* irBest in search-level coordinates */
  double found_x, found_y;      /* td.mv2Found (level-0 coordinates) */
  int32_t n_candidates;         /* corners tested (nValidCorners) */
  int32_t pad_;
} McpPatchRes;

void mcp_fe_default_config(McpFeConfig* cfg);
int mcp_fe_create(const McpFeConfig* cfg, McpFe** out);
int mcp_fe_destroy(McpFe* h);
/* Optional fixed mask (KeyFrame::SetMask): level-0 mask, half-sampled internally; NULL clears it. */
int mcp_fe_set_mask(McpFe* h, const uint8_t* mask, int32_t stride);
/* bGlareMasking of KeyFrame::MakeKeyFrame_Lite (src/KeyFrame.cc:214-242; include/mcptam/KeyFrame.h:186): when enabled, the
 * following mcp_fe_make_keyframe calls AND the internal mask with the glare mask -- 0 wherever a pixel brighter than 245 lies
 * within five dilations by OpenCV's 5x5 MORPH_ELLIPSE element -- before the corners are filtered. */
int mcp_fe_set_glare_masking(McpFe* h, int32_t enable);
/* Builds the 4-level pyramid + corners of one image into resident keyframe slot `slot`
 * (H2D copy of the image, D2H copy of the corner lists/LUTs when out pointers are given). */
int mcp_fe_make_keyframe(McpFe* h, int32_t slot, const uint8_t* img, int32_t stride, McpLevelOut out[MCP_LEVELS]);
/* One call per (camera, vTD): searches the pyramid in slot target_kf. */
int mcp_fe_search_patches(McpFe* h, int32_t target_kf, int32_t n, const McpPatchReq* req, McpPatchRes* res);
/* Debug/parity: the 8x8 templates generated for the last search call (n*64 bytes). */
int mcp_fe_get_templates(McpFe* h, int32_t n, uint8_t* templ);
int mcp_fe_shitomasi(McpFe* h, int32_t kf, int32_t level, int32_t n, const int32_t* xy, double* scores);
/* MiniPatch: sample 9x9 patches at src positions in (kf_src, level), find them among the FAST corners
 * of (kf_dst, level) within +-range.  found[i] = 1 and pos_out updated, else 0. */
int mcp_fe_minipatch_find(McpFe* h, int32_t kf_src, int32_t kf_dst, int32_t level, int32_t n,
                          const int32_t* src_xy, const int32_t* start_xy, int32_t range, int32_t* pos_out,
                          int32_t* found);
/* KeyFrame::MakeKeyFrame_Rest candidate generation (src/KeyFrame.cc:363-531) for the pyramid resident in `slot`:
 * CVD::fast_nonmax at Level::nFastThresh, in_image_with_border(10), FAST or Shi-Tomasi scoring, "percent" (top
 * fraction of the descending sort) or "thresh" selection, then the MiniPatch stable-point test against the oldest
 * stored previous frame (Level::imagePrev[0] / vCornersPrev[0], here another resident slot).  Fills Level::vCandidates. */
typedef struct McpCandidate {
  int32_t x, y;                 /* Candidate::irLevelPos */
  double score;                 /* Candidate::dSTScore */
} McpCandidate;
typedef struct McpRestConfig {
  int32_t use_shi;              /* KeyFrame::ssCandidateType == "shi" (default "fast": 0) */
  int32_t use_thresh;           /* KeyFrame::ssCandidateCriterion == "thresh" (default "percent": 0) */
  double top_fraction;          /* KeyFrame::sdCandidateTopFraction (0.8) */
  double thresh;                /* KeyFrame::sdCandidateThresh (70) */
  int32_t nonmax_strict;        /* 0: libCVD nonmax_suppression (dropped only by a strictly higher neighbour, default);
                                   1: nonmax_suppression_strict (dropped by a higher-or-equal neighbour) */
  int32_t prev_slot;            /* resident slot holding imagePrev[0] / vCornersPrev[0]; -1: no history (no pruning) */
  int32_t n_prev;               /* imagePrev.size(): the MiniPatch search range is 10 * n_prev */
  int32_t pad_;
} McpRestConfig;
typedef struct McpRestLevelOut {
  int32_t n_max;                /* vScoresAndMaxCorners.size() */
  int32_t n_selected;           /* candidates before the stable-point test */
  int32_t n_candidates;         /* vCandidates.size() */
  int32_t cap;                  /* capacity of cand[] (entries beyond it are dropped) */
  McpCandidate* cand;           /* host out, may be NULL */
} McpRestLevelOut;
void mcp_fe_default_rest_config(McpRestConfig* cfg);
int mcp_fe_make_keyframe_rest(McpFe* h, int32_t slot, const McpRestConfig* cfg, McpRestLevelOut out[MCP_LEVELS]);
/* Debug: FAST score map of one level (0 = no corner at b=5, else fast_corner_score_10). */
int mcp_fe_debug_scores(McpFe* h, int32_t slot, int32_t level, uint8_t* out);
/* FindPVS building block: TrackerData::Project + GetDerivsUnsafe (include/mcptam/TrackerData.h:102-129) and
 * PatchFinder::CalcSearchLevelAndWarpMatrix (src/PatchFinder.cc:69-122) for n map points against one camera pose. */
typedef struct McpProjRes {
  double px[2];          /* td.mv2Image */
  double cam_derivs[4];  /* td.mm2CamDerivs, row-major */
  double warp_inv[4];    /* PatchFinder::mm2WarpInverse, row-major */
  double v3cam[3];       /* td.mv3Cam */
  int32_t in_image;      /* td.mbInImage (camera valid and inside the image, '>' size as in the reference) */
  int32_t search_level;  /* GetLevel(), -1 = template bad */
} McpProjRes;
int mcp_fe_set_camera(McpFe* h, const McpTaylorCam* cam);
/* cam_from_world: 12 doubles (row-major R, t); world_xyz / pixel_right_w / pixel_down_w: n x 3 doubles (host) */
int mcp_fe_project_points(McpFe* h, const double* cam_from_world, int32_t n, const double* world_xyz,
                          const double* pixel_right_w, const double* pixel_down_w, McpProjRes* out);
/* Tracker pose update ("next" row): TrackerData::ProjectAndDerivs + CalcJacobian (include/mcptam/TrackerData.h:102-178)
 * and Tracker::CalcPoseUpdate (src/Tracker.cc:1386-1511, TooN WLS<6> with prior 100). */
typedef struct McpJacRes {
  double px[2];          /* td.mv2Image */
  double jac[12];        /* td.mm26Jacobian, row-major 2x6, w.r.t. the MKF base pose */
  int32_t in_image, pad_;
} McpJacRes;
typedef struct McpPoseMeas {
  double found[2];       /* td.mv2Found */
  double image[2];       /* td.mv2Image */
  double sqrt_inv_noise; /* td.mdSqrtInvNoise */
  double jac[12];        /* td.mm26Jacobian */
  int32_t found_flag;    /* td.mbFound */
  int32_t pad_;
} McpPoseMeas;
typedef struct McpPoseUpdate {
  double mu[6];          /* wls.get_mu(): the 6-vector pose update */
  double sigma_sq;       /* M-estimator sigma^2 used */
  double c_inv[36];      /* wls.get_C_inv() (for the pose covariance) */
  int32_t n_inliers;     /* measurements with non-zero weight */
  int32_t n_valid;       /* found measurements */
} McpPoseUpdate;
int mcp_fe_calc_jacobians(McpFe* h, const double* base_from_world, const double* cam_from_base, int32_t n,
                          const double* world_xyz, McpJacRes* out);
/* estimator: 0 Tukey, 1 Cauchy, 2 Huber (Tracker::sMEstimatorName); override_sigma <= 0: estimate from the data.
 * outlier[i] = 1 where the weight is exactly zero (may be NULL). */
int mcp_fe_pose_update(McpFe* h, int32_t n, const McpPoseMeas* meas, int32_t estimator, double override_sigma,
                       McpPoseUpdate* out, int32_t* outlier);
typedef struct McpFeTiming { double ms_pyramid, ms_fast, ms_compact, ms_search, ms_other; int32_t n_launches, pad_; } McpFeTiming;
int mcp_fe_get_timing(McpFe* h, McpFeTiming* out);

#ifdef __cplusplus
}
#endif
#endif /* MCPTAM_B200_H */
