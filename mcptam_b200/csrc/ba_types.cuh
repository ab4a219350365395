// ba_types.cuh — device-side data layout of one bundle-adjustment problem.
#pragma once

#include "mcp_common.cuh"

namespace mcp {

constexpr int SEL_BITS = 11;
constexpr int SEL_BINS = 1 << SEL_BITS;   // 2048
constexpr int SEL_PASSES = 6;             // 6 x 11 bits >= 63 significant bits of |chi2|
constexpr int MAX_PARTIALS = 4096;        // per-block partial sums (grid size cap for the per-point kernels)
constexpr int RS_BYTES = 6144;            // staging bytes per buffer of k_schur_rows
constexpr int RS_MAXE = 8;                // entries (point, pose slot) per group
constexpr int MREC = 22;                  // doubles per measurement record: A(6) q(3) w we0 we1 A2(6) qs(3) obs-var
constexpr int MAX_CAND = 4;               // speculative LM candidates evaluated concurrently
constexpr int N_STATE = MAX_CAND + 1;     // accepted state + one trial buffer per candidate

// LM control block: lives in device memory, mirrored into pinned host memory after every trial.
// Restates the state of g2o::OptimizationAlgorithmLevenberg + the ChainBundle actions
// (reference src/ChainBundle.cc:904-1126).
struct BaCtrl {
  double lambda, ni;
  double sigma_sq_raw, sigma_sq_lim, sigma_lim;   // RobustKernelData (src/ChainBundle.cc:810-833)
  double current_chi, temp_chi;
  double scale[MAX_CAND], sumsq[MAX_CAND];                     // computeScale(), sum x^2 of the last solve, per speculative candidate
  double max_diag;
  double last_chi2;                               // CheckConvergedResidualAction::_dLastChi2
  double rho;
  double min_sigma_sq, pct_limit, rms_limit, user_lambda;
  double tukey_sigma_sq;
  double lin_chi;     // robust chi2 at the linearisation point of the current outer iteration
  int cur;            // index of the accepted state buffers
  int accepted;       // last trial accepted
  int stop_trials;    // trial loop of this outer iteration is over
  int terminate;      // solver returned Terminate (qmax hit / rho == 0)
  int qmax;           // trials in this outer iteration
  int solve_ok[MAX_CAND];    // per candidate
  int iter;           // outer iterations completed in this Compute
  int conv_mag, conv_res;
  int total_trials;
  int max_trials, use_robust;
  int dim;            // 6*n_pose_var + 3*n_pt_var (global)
  int need_lambda_init;
  int n_outliers;
  int sel_n;          // number of values in the selection (global measurement count)
  int sel_rank;       // n/2
  int cand_used;      // candidates consumed by the last k_lm_control
  int acc_cand;       // index of the candidate the last k_lm_control accepted (valid if accepted)
  int pad_acc;
  int marg_fail;      // computeMarginals() failed (singular block)
  double median_out;  // plain upper median of the last mode-2 selection (point-depth covariances)
  double med_hint;     // median |chi2| of the current state (0: unknown): centre of the bracket k_select_cluster tries first
  double abort_agreed; // multi-GPU: sum over the ranks of their abort flags as of the last trial round (> 0: everybody stops)
};

struct BaDev {
  const DevCam* cams;
  int n_pose, n_pt, n_meas, n_pose_var, n_pt_var, nc, n_slots, max_slots;
  int n_cam, stage_doubles;      // cameras; doubles of shared memory used to stage poses + cameras (0: read from global)
  int p_lo, p_hi;                // local point range (multi-GPU shard)
  int m_lo, m_hi;                // local measurement range (sorted order)
  const int* pose_var;           // [n_pose] variable index or -1
  const int4* pt_info;           // [n_pt] {src link0 pose id, src link1 pose id | -1, src var | -1, src slot | -1}
  const int* pt_var;             // [n_pt] point variable index or -1 (fixed)
  const int* pt_order;           // [n_pt] visiting order: each rank's point range sorted by measurement count, heaviest first
  const int* pt_meas_off;        // [n_pt+1]
  const int* pt_slot_off;        // [n_pt+1]
  const int* slot_var;           // [n_slots] pose variable of each (point, slot), ascending within a point
  const double2* meas_xy;        // [n_meas] sorted by point
  const double* meas_info;       // [n_meas] 1/sqrt(dNoiseSigmaSquared)
  const int4* meas_a;            // [n_meas] {obs link0 pose id, obs link1 pose id | -1, camera, original index}
  const int4* meas_b;            // [n_meas] {obs var | -1 (no obs Jacobian), obs slot | -1, has_src_jac, point id}
  double* pose[N_STATE];              // [n_pose*12]  accepted state + one trial buffer per speculative candidate
  double* pt[N_STATE];               // [n_pt*3]
  double* chi2[N_STATE];              // [n_meas] signed as EdgeChainMeas::chi2
  int cand, ahead;               // speculative candidate index of this launch (c: lambda after c rejections); ahead = 1: this launch
                                 // was enqueued before the host knew the trial loop ended (see next_iteration_started)
  int pick_sigma, pad_pick;      // look-ahead linearisation: take the Huber sigma^2 of the accepted candidate from spec_sigma (k_linearize prologue)
  double* V;                     // [n_pt*6]  upper triangle of J_pt^T W J_pt
  double* gp;                    // [n_pt*3]
  double* W;                     // [n_slots*18] 6x3 row-major
  double* Y;                     // [n_slots*24] W (V + lambda I)^-1 (18) followed by Y g_p (6)
  const int* slot_pt;            // [n_slots] owning point of each slot
  int slot_lo, slot_hi;          // local slot range (multi-GPU shard)
  const int2* inc;               // co-visibility incidences {slot A, slot B}, bucketed by block pair
  const int4* items;             // work items {block row, block col, begin, end} into inc
  const int* n_items_dev;        // number of work items (written by k_pair_items at load time)
  int max_items, pad_items;      // host-side upper bound of it (grid sizing)
  const int2* rs_ent;            // row-wise Schur lists: {slot, slots from it to the end of its point}, sorted by pose variable
  const int* rs_grp;             // groups of <= RS_MAXE entries that fit one staging buffer: (first entry << 4) | count
  const int4* rs_items;          // work items {pose variable a, first group, end group, 0}
  int n_rs_items, schur_mode;    // schur_mode 0: row-wise (k_schur_rows), 1: pair gathers with TMA, 2: staged pair gathers
  int rs_nblk, pad_rs;           // blocks per accumulator strip: 1 + the largest (pose variable - a) inside one point
  double* mrec;                  // [n_meas][MREC] per-measurement record written by k_linearize for k_pose_blocks
  const int* pb_idx;             // measurement indices bucketed by pose block: (v,v) diagonal blocks, then (lo,hi) observer/source pairs
  const int4* pb_items;          // work items {block row, block col, begin, end} into pb_idx
  int n_pb_items, pad_pb;
  double* H0;                    // [nc*nc] upper block triangle, pose-pose normal matrix (no damping)
  double* Sm;                    // [nc*nc] upper block triangle, sum_p W (V+lambda I)^-1 W^T
  double* gc;                    // [nc]
  double* rm;                    // [nc]   sum_p W (V+lambda I)^-1 g_p
  double* dc;                    // [nc]   pose update
  double* L;                     // Cholesky factor, 32x32 tiles (lower block triangle + rhs block row)
  double* Linv;                  // inverse diagonal tiles
  uint4* Lll;                    // the same tiles as self-validating 16-byte lines {lo, epoch, hi, epoch} (ba_solve.cu)
  int* flags;                    // dataflow ready flags + counters (ba_solve.cu)
  double* part;                  // [8][MAX_PARTIALS] per-block partial sums
  unsigned* sel_hist;            // [SEL_PASSES][SEL_BINS] radix-select histograms (self re-arming)
  unsigned* sel_done;            // [SEL_PASSES] ticket counters
  unsigned long long* sel_state; // [SEL_PASSES][2] prefix, rank
  double* spec_sigma;            // [MAX_CAND] Huber sigma^2 (raw) of every candidate's trial state, computed speculatively (mode 3)
  double* spec_med;              // [MAX_CAND] the medians they come from
  BaCtrl* ctrl;
  int* outlier_flags;            // [n_meas] (sorted order)
  double* dbg;                   // optional debug output
};

// lambda of the LM trial this launch belongs to: candidate c is the trial g2o would run after rejecting candidates
// 0..c-1 (each rejection: lambda *= ni, ni *= 2)
__device__ __forceinline__ double trial_lambda(const BaDev& d)
{
  if (d.cand < 0) return 0.0;                      // marginals pass: the undamped Hessian
  double l = d.ctrl->lambda, ni = d.ctrl->ni;
  for (int c = 0; c < d.cand; c++) { l *= ni; ni *= 2; }
  return l;
}
// Look-ahead launches (next outer iteration's sigma / linearisation enqueued right behind k_lm_control, before the host
// has read the control block) run only if that k_lm_control closed the trial loop without ending the optimisation.
__device__ __forceinline__ bool lookahead_skip(const BaDev& d)
{
  if (!d.ahead) return false;
  const BaCtrl* c = d.ctrl;
  return !(c->stop_trials && !c->terminate && !c->conv_mag && !c->conv_res);
}
__device__ __forceinline__ int trial_buffer(const BaDev& d, int cur) { return (cur + 1 + d.cand) % N_STATE; }

// per-candidate outputs of the multi-candidate Schur reduction and the shared per-point record array
struct SchurMulti {
  double* Sm[3]; double* rm[3]; double* R;
  int n_cand, zero_mask;          // bit c of zero_mask: k_schur_vinv_multi clears candidate c's [Sm | rm] first
  size_t sm_doubles;              // doubles in one [Sm | rm] buffer
  int* next_item;                 // work counter of k_schur_pairs_multi_ca, set to first_dynamic_item by k_schur_vinv_multi
  int first_dynamic_item, pad;    // = number of warps of the pair kernel (every warp starts on its own index)
};

struct CandParts { const double* p[MAX_CAND]; };

// one launch at the end of mcp_ba_load instead of ~25 memset / device-to-device copy calls: ranges to clear (16-byte aligned
// buffers, sizes in bytes, multiples of 4) and the initial state to replicate into the N_STATE state buffers
constexpr int LOAD_ZERO_MAX = 20;
struct LoadInit {
  void* zero_ptr[LOAD_ZERO_MAX];
  unsigned long long zero_bytes[LOAD_ZERO_MAX];
  int n_zero, pad;
  const double* pose0; const double* pt0;
  double* pose[N_STATE]; double* pt[N_STATE];
  unsigned long long pose_doubles, pt_doubles;
};   // per-candidate partial-sum arrays handed to k_lm_control

enum PartialRow { PART_CUR_CHI = 0, PART_MAXDIAG = 1, PART_TMP_CHI = 2, PART_SCALE = 3, PART_SUMSQ = 4 };

}  // namespace mcp
