// ba_api.cu — C ABI of the bundle adjuster (include/mcptam_b200.h): host marshalling of the map into the
// flat device layout, the LM driver loop (restating g2o's SparseOptimizer::optimize around the device
// kernels) and the NCCL exchange of the Schur-reduced camera system for point-sharded multi-GPU runs.
//
// Reference: ChainBundle::{AddPose,AddPoint,AddMeas,Compute,...}  src/ChainBundle.cc:1198-1488
//            BundleAdjusterMulti::BundleAdjust (marshalling)        src/BundleAdjusterMulti.cc:55-203
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#include "ba_prep.hpp"
#include "ba_types.cuh"

namespace mcp {

static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Programmatic dependent launch of the per-round kernel chain (mcp_common.cuh).  MCP_BA_PDL=0/1 overrides; the default is
// on for single-GPU processes (the whole GPU test suite runs with it) and off once a handle of the process joins a
// communicator (the multi-GPU path has NCCL kernels inside the chain and was not measured with it).
static std::atomic<bool> g_pdl_multi_gpu{ false };
bool pdl_enabled()
{
  static const int env = [] { const char* e = getenv("MCP_BA_PDL"); return e && e[0] ? (e[0] != '0' ? 1 : 0) : -1; }();
  if (env >= 0) return env == 1;
  return !g_pdl_multi_gpu.load(std::memory_order_relaxed);
}

// launchers defined in ba_kernels.cu
int launch_linearize(const BaDev& d, int warps, size_t smem, cudaStream_t s, const SchurMulti* mc = nullptr, double* zero_ptr = nullptr, size_t zero_n = 0);
int launch_backsub_eval(const BaDev& d, int apply, int which, double* err_out, cudaStream_t s);
int launch_select_sigma(const BaDev& d, int which, int mode, cudaStream_t s, double* zero_ptr = nullptr, size_t zero_n = 0);
void launch_tukey_flags(const BaDev& d, cudaStream_t s);
void launch_load_init(const LoadInit& li, cudaStream_t s);
int launch_robust_sum(const BaDev& d, int which, cudaStream_t s);
void launch_lambda_init(const BaDev& d, cudaStream_t s);
void launch_lambda_apply(const BaDev& d, cudaStream_t s);
void launch_chol_solve(const BaDev& d, int epoch, int max_ctas, int* task_base, cudaStream_t s);
size_t chol_tiles_doubles(int nc);
size_t chol_inv_doubles(int nc);
size_t chol_ll_bytes(int nc);
size_t chol_flag_ints(int nc);
int chol_max_n();
void launch_lm_control(const BaDev& d, const CandParts& parts, int n_cand, int n_lin, int n_bs, const double* red_in, int first_trial, cudaStream_t s);
void launch_reduce_partials(const BaDev& d, int n_lin, int n_bs, double* out, cudaStream_t s, const double* host_word = nullptr, int word_slot = 0);
void launch_debug_jacobians(const BaDev& d, double* out, cudaStream_t s);
void launch_gather_delta(const BaDev& d, double* out, cudaStream_t s);
int configure_kernels(int max_slots, int stage_doubles, int* warps_out, size_t* smem_out);
int stage_doubles_for(int n_pose, int n_cam);
void launch_pair_count(const BaDev& d, int* cnt, cudaStream_t s);
void launch_pair_fill(const BaDev& d, int* cursor, int2* inc, cudaStream_t s);
void launch_pair_items(const BaDev& d, int* cnt, int4* items, int* n_items_out, cudaStream_t s);
void launch_schur_gather(const BaDev& d, cudaStream_t s);
void launch_schur_multi(const BaDev& d, const SchurMulti& mc, cudaStream_t s, bool records_ready = false);
void schur_multi_prepare(const BaDev& d, SchurMulti& mc);
void launch_marginals(const BaDev& d, double* cov, cudaStream_t s);
void launch_zero_acc(const BaDev& d, double* acc, size_t n, cudaStream_t s, int pick_sigma = 0);
bool select_spec_possible(const BaDev& d);
void launch_select_spec(const BaDev& d, int which, cudaStream_t s);
void launch_tri_pack(double* const* full, int count, double* packed, int n, int tail, bool unpack, cudaStream_t s);
struct P2pPeers { uint4* base[8]; };
static inline size_t p2p_slice(size_t cnt, int world) { return (cnt + world - 1) / world; }   // as in ba_p2p.cu
void launch_p2p_allreduce(double* buf, size_t cnt, const P2pPeers& peers, size_t box_off, int rank, int world, unsigned tag, cudaStream_t s);
void launch_p2p_allgather(const BaDev& d, double* arr_or_null, int stride, const int* bounds, const P2pPeers& peers, size_t box_off, int rank, int world,
                          unsigned tag, cudaStream_t s);

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes)
  {
    if (bytes <= cap && p) return MCP_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    const size_t want = bytes + bytes / 4 + 256;          // pooled across BundleAdjust calls
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { set_last_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); return MCP_ERR_CUDA; }
    cap = want;
    return MCP_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// pinned host storage for the marshalling output (ba_prep.hpp): uploads from it are truly asynchronous
static void* pinned_alloc(size_t bytes)
{
  void* p = nullptr;
  return cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? p : nullptr;
}
static void pinned_release(void* p) { cudaFreeHost(p); }
static const PrepAlloc g_pinned_alloc = { pinned_alloc, pinned_release };

enum Cat { C_SELECT = 0, C_LIN, C_SCHUR, C_SOLVE, C_BACKSUB, C_CONTROL, C_OTHER, C_N };

}  // namespace mcp

using namespace mcp;

struct McpBa {
  McpBaConfig cfg;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<DevCam> cams;
  bool loaded = false;
  BaDev d;
  int lin_warps = 8;
  size_t lin_smem = 0;
  // device buffers (pooled)
  DevBuf b_cams, b_pose_var, b_pt_info, b_pt_var, b_pt_order, b_pt_meas_off, b_pt_slot_off, b_slot_var, b_meas_xy, b_meas_info,
      b_meas_a, b_meas_b, b_pose[N_STATE], b_pt[N_STATE], b_chi2[N_STATE], b_V, b_gp, b_W, b_acc, b_dc, b_L, b_part, b_ctrl, b_flags,
      b_pose0, b_pt0, b_tmp, b_Linv, b_Lll, b_cflags, b_dbg, b_sel, b_Y, b_slot_pt, b_inc, b_items, b_paircnt, b_mrec, b_pb_idx, b_pb_items, b_rs_ent, b_rs_grp, b_rs_items, b_R, b_pack;
  // speculative LM candidates 1..n_spec-1 (lambda after that many rejections), one extra stream each
  struct Cand {
    DevBuf b_acc, b_dc, b_L, b_Linv, b_Lll, b_cflags, b_part, b_Y;
    BaDev d;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_done = nullptr, ev_schur = nullptr;
    int chol_epoch = 0, chol_task_base = 0;
  } cand[MAX_CAND];               // [0] unused (candidate 0 lives in the handle's own buffers)
  cudaEvent_t ev_ready = nullptr, ev_red = nullptr, ev_ctrl = nullptr;
  // the co-visibility lists of a load are built on aux_stream next to the first evaluation / selection / linearisation of the
  // Compute that follows; the first Schur reduction waits for ev_pairs
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_up = nullptr, ev_pairs = nullptr;
  bool pairs_pending = false;
  cudaStream_t sel_stream[MAX_CAND] = { nullptr, nullptr, nullptr, nullptr };   // speculative sigma of every candidate's trial state
  cudaEvent_t ev_bs[MAX_CAND] = { nullptr, nullptr, nullptr, nullptr }, ev_sel[MAX_CAND] = { nullptr, nullptr, nullptr, nullptr };
  cudaStream_t copy_stream = nullptr;   // control-block read-back that does not queue behind look-ahead kernels
  int n_spec_multi = 3;           // candidates per round when sharded over several GPUs (one grouped all-reduce per round)
  int n_spec = 3;                 // candidates per round (1 = no speculation)
  bool fuse_schur = true;         // one multi-candidate Schur pass per round (MCP_BA_FUSE_SCHUR=0: one pass per candidate)
  int spec_rounds = 0, spec_used = 0;
  int chol_epoch = 0, chol_task_base = 0, n_sms = 148;
  size_t acc_doubles = 0, off_H0 = 0, off_gc = 0, off_red = 0, off_Sm = 0, off_rm = 0;
  BaCtrl* ctrl_host = nullptr;   // pinned
  double* abort_word = nullptr;  // pinned [4]: this rank's abort flag for the next round, the all-reduced one of the last round, chi2 before / after
  int* flags_host = nullptr;     // pinned, n_meas
  size_t flags_cap = 0;
  BaPrep prep;                   // host marshalling output in pinned memory, pooled across loads
  std::unique_ptr<HostPool> pool; // host threads of the marshalling passes
  std::vector<int> meas_orig;    // sorted position -> original index
  std::vector<int32_t> outliers;
  // multi-GPU
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  std::vector<int> part_pt, part_meas;   // world+1 boundaries
  // peer-memory exchanges (ba_p2p.cu): this rank's exchange buffer, the peers' buffers as mapped through CUDA IPC, one region
  // (two parities) per kind of exchange, call counters
  enum { X_H0 = 0, X_SM, X_RED, X_CHI, X_N };
  bool p2p = false;
  uint4* xbuf = nullptr;
  size_t xbuf_lines = 0, x_off[X_N] = { 0, 0, 0, 0 }, x_lines[X_N] = { 0, 0, 0, 0 };
  unsigned x_calls[X_N] = { 0, 0, 0, 0 }, x_tag = 0;
  P2pPeers peers;
  // MCP_BA_TIMELINE=1: start/stop events around the launches of every stream (does not serialise the streams);
  // dumped to stderr at the end of mcp_ba_compute as 'TL name stream start_us stop_us'
  struct Tl { const char* name; int sid; cudaEvent_t a, b; };
  std::vector<Tl> tl;
  bool timeline = false;
  // profiling
  bool profiling = false;
  McpBaTiming timing;
  struct Ev { int cat; cudaEvent_t a, b; };
  std::vector<Ev> evs;
  int launches = 0;
  double last_chi2_init = 1.7976931348623157e308;
};

extern "C" {

static int p2p_setup(McpBa* h, size_t n_meas, size_t ncp);
static void p2p_release(McpBa* h);

const char* mcp_last_error(void) { return g_err; }
int mcp_abi_version(void) { return 1; }

void mcp_ba_default_config(McpBaConfig* c)
{
  memset(c, 0, sizeof(*c));
  c->use_robust = 1;
  c->use_tukey = 1;
  c->max_trials_after_failure = 100;   // src/ChainBundle.cc:1133
  c->update_pct_limit = 1e-10;         // :1134
  c->update_rms_limit = 1e-10;         // :1135
  c->min_sigma = 0.5;                  // :1136
  c->device = -1;
}

static int ba_create_impl(const McpBaConfig* cfg, McpBa* h)
{
  if (cfg) h->cfg = *cfg; else mcp_ba_default_config(&h->cfg);
  if (h->cfg.device >= 0) { MCP_CUDA_CHECK(cudaSetDevice(h->cfg.device)); }
  MCP_CUDA_CHECK(cudaGetDevice(&h->device));
  { int sms = 0; if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device) == cudaSuccess && sms > 0) h->n_sms = sms; }
  MCP_CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  MCP_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
  MCP_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_red, cudaEventDisableTiming));
  MCP_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_ctrl, cudaEventDisableTiming));
  MCP_CUDA_CHECK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  MCP_CUDA_CHECK(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
  MCP_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_up, cudaEventDisableTiming));
  MCP_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_pairs, cudaEventDisableTiming));
  for (int q = 0; q < MAX_CAND; q++) {
    MCP_CUDA_CHECK(cudaStreamCreateWithFlags(&h->sel_stream[q], cudaStreamNonBlocking));
    MCP_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_bs[q], cudaEventDisableTiming));
    MCP_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_sel[q], cudaEventDisableTiming));
  }
  for (int q = 1; q < MAX_CAND; q++) {
    MCP_CUDA_CHECK(cudaStreamCreateWithFlags(&h->cand[q].stream, cudaStreamNonBlocking));
    MCP_CUDA_CHECK(cudaEventCreateWithFlags(&h->cand[q].ev_done, cudaEventDisableTiming));
    MCP_CUDA_CHECK(cudaEventCreateWithFlags(&h->cand[q].ev_schur, cudaEventDisableTiming));
  }
  // MCP_BA_SPECULATE = number of LM candidates evaluated per round (1 or 0: sequential trials)
  { const char* e = getenv("MCP_BA_SPECULATE"); if (e && e[0]) { int v = atoi(e); h->n_spec = v < 1 ? 1 : (v > MAX_CAND ? MAX_CAND : v); } }
  { const char* e = getenv("MCP_BA_FUSE_SCHUR"); h->fuse_schur = !(e && e[0] == '0'); }
  { const char* e = getenv("MCP_BA_SPECULATE_MULTI"); if (e && e[0]) { int v = atoi(e); h->n_spec_multi = v < 1 ? 1 : (v > MAX_CAND ? MAX_CAND : v); } }
  MCP_CUDA_CHECK(cudaEventCreate(&h->ev0));
  MCP_CUDA_CHECK(cudaEventCreate(&h->ev1));
  MCP_CUDA_CHECK(cudaMallocHost(&h->ctrl_host, sizeof(BaCtrl)));
  memset(h->ctrl_host, 0, sizeof(BaCtrl));
  MCP_CUDA_CHECK(cudaMallocHost(&h->abort_word, 4 * sizeof(double)));
  h->abort_word[0] = h->abort_word[1] = h->abort_word[2] = h->abort_word[3] = 0;
  memset(&h->d, 0, sizeof(h->d));
  memset(&h->timing, 0, sizeof(h->timing));
  return MCP_OK;
}

int mcp_ba_destroy(McpBa* h);
int mcp_ba_create(const McpBaConfig* cfg, McpBa** out)
{
  if (!out) { set_last_error("mcp_ba_create: out is NULL"); return MCP_ERR_INVALID; }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_last_error("mcp_ba_create: no CUDA device available (this library has no CPU fallback)");
    return MCP_ERR_NO_DEVICE;
  }
  McpBa* h = new McpBa();
  const int rc = ba_create_impl(cfg, h);
  if (rc != MCP_OK) { mcp_ba_destroy(h); return rc; }       // streams, events and pinned buffers created so far are released
  *out = h;
  return MCP_OK;
}

int mcp_ba_destroy(McpBa* h)
{
  if (!h) return MCP_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  DevBuf* all[] = { &h->b_cams, &h->b_pose_var, &h->b_pt_info, &h->b_pt_var, &h->b_pt_order, &h->b_pt_meas_off, &h->b_pt_slot_off,
                    &h->b_slot_var, &h->b_meas_xy, &h->b_meas_info, &h->b_meas_a, &h->b_meas_b, &h->b_V, &h->b_gp, &h->b_W,
                    &h->b_acc, &h->b_dc, &h->b_L, &h->b_part, &h->b_ctrl, &h->b_flags, &h->b_pose0, &h->b_pt0, &h->b_tmp, &h->b_Linv, &h->b_Lll, &h->b_cflags, &h->b_dbg, &h->b_sel, &h->b_Y, &h->b_slot_pt, &h->b_inc, &h->b_items, &h->b_paircnt, &h->b_mrec, &h->b_pb_idx, &h->b_pb_items, &h->b_rs_ent, &h->b_rs_grp, &h->b_rs_items, &h->b_R, &h->b_pack };
  for (DevBuf* b : all) b->release();
  for (int k = 0; k < N_STATE; k++) { h->b_pose[k].release(); h->b_pt[k].release(); h->b_chi2[k].release(); }
  for (int q = 1; q < MAX_CAND; q++) {
    McpBa::Cand& cq = h->cand[q];
    if (cq.stream) cudaStreamSynchronize(cq.stream);
    DevBuf* cb[] = { &cq.b_acc, &cq.b_dc, &cq.b_L, &cq.b_Linv, &cq.b_Lll, &cq.b_cflags, &cq.b_part, &cq.b_Y };
    for (DevBuf* b : cb) b->release();
    if (cq.stream) cudaStreamDestroy(cq.stream);
    if (cq.ev_done) cudaEventDestroy(cq.ev_done);
    if (cq.ev_schur) cudaEventDestroy(cq.ev_schur);
  }
  if (h->ctrl_host) cudaFreeHost(h->ctrl_host);
  if (h->abort_word) cudaFreeHost(h->abort_word);
  if (h->flags_host) cudaFreeHost(h->flags_host);
  h->prep.free_all(g_pinned_alloc);
  p2p_release(h);
  if (h->comm) ncclCommDestroy(h->comm);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->ev_ready) cudaEventDestroy(h->ev_ready);
  if (h->ev_red) cudaEventDestroy(h->ev_red);
  if (h->ev_ctrl) cudaEventDestroy(h->ev_ctrl);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
  if (h->ev_up) cudaEventDestroy(h->ev_up);
  if (h->ev_pairs) cudaEventDestroy(h->ev_pairs);
  for (int q = 0; q < MAX_CAND; q++) {
    if (h->sel_stream[q]) { cudaStreamSynchronize(h->sel_stream[q]); cudaStreamDestroy(h->sel_stream[q]); }
    if (h->ev_bs[q]) cudaEventDestroy(h->ev_bs[q]);
    if (h->ev_sel[q]) cudaEventDestroy(h->ev_sel[q]);
  }
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return MCP_OK;
}

int mcp_ba_set_cameras(McpBa* h, int32_t n_cam, const McpTaylorCam* cams)
{
  if (!h || n_cam <= 0 || !cams) { set_last_error("mcp_ba_set_cameras: bad arguments"); return MCP_ERR_INVALID; }
  h->cams.resize(n_cam);
  for (int i = 0; i < n_cam; i++) {
    const McpTaylorCam& s = cams[i];
    if (s.n_inv < 2 || s.n_inv > 32) { set_last_error("camera %d: n_inv=%d out of range", i, s.n_inv); return MCP_ERR_INVALID; }
    DevCam& c = h->cams[i];
    memset(&c, 0, sizeof(c));
    memcpy(c.poly, s.poly, sizeof(c.poly));
    // mv5PolyDerivModCoeffs, src/TaylorCamera.cc:107-110
    c.dmod[0] = -s.poly[0]; c.dmod[1] = s.poly[1]; c.dmod[2] = s.poly[2]; c.dmod[3] = 2 * s.poly[3]; c.dmod[4] = 3 * s.poly[4];
    memcpy(c.center, s.center, sizeof(c.center));
    memcpy(c.affine, s.affine, sizeof(c.affine));
    memcpy(c.image_size, s.image_size, sizeof(c.image_size));
    c.min_theta = s.min_theta; c.theta_mean = s.theta_mean; c.theta_std = s.theta_std;
    c.n_inv = s.n_inv;
    memcpy(c.inv, s.inv_poly, sizeof(double) * 32);
  }
  cudaSetDevice(h->device);
  int rc = h->b_cams.ensure(sizeof(DevCam) * n_cam);
  if (rc) return rc;
  h->loaded = false;                                   // the camera table may have moved: a problem must be (re)loaded against it
  MCP_CUDA_CHECK(cudaMemcpyAsync(h->b_cams.p, h->cams.data(), sizeof(DevCam) * n_cam, cudaMemcpyHostToDevice, h->stream));
  MCP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return MCP_OK;
}

static int upload(McpBa* h, DevBuf& b, const void* src, size_t bytes)
{
  int rc = b.ensure(bytes ? bytes : 16);
  if (rc) return rc;
  if (bytes) MCP_CUDA_CHECK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, h->stream));
  return MCP_OK;
}

int mcp_ba_load(McpBa* h, int32_t n_pose, const double* pose_Rt, const uint8_t* pose_fixed, int32_t n_pt,
                const double* pt_xyz, const int32_t* pt_chain, const uint8_t* pt_fixed, int32_t n_meas,
                const double* meas_xy, const int32_t* meas_chain, const int32_t* meas_pt, const double* meas_noise,
                const int32_t* meas_cam)
{
  if (!h) { set_last_error("mcp_ba_load: NULL handle"); return MCP_ERR_INVALID; }
  if (h->cams.empty()) { set_last_error("mcp_ba_load: call mcp_ba_set_cameras first"); return MCP_ERR_STATE; }
  if (n_pose <= 0 || n_pt < 0 || n_meas < 0 || !pose_Rt || !pose_fixed || (n_pt && (!pt_xyz || !pt_chain || !pt_fixed)) ||
      (n_meas && (!meas_xy || !meas_chain || !meas_pt || !meas_noise || !meas_cam))) {
    set_last_error("mcp_ba_load: bad arguments");
    return MCP_ERR_INVALID;
  }
  h->loaded = false;
  // mcp_ba_load returns without waiting for the device: whatever an earlier load / compute left in flight must be over before
  // the pooled staging arrays are rewritten (normally nothing is: every Compute ends synchronised)
  MCP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  if (h->pairs_pending) { MCP_CUDA_CHECK(cudaStreamSynchronize(h->aux_stream)); h->pairs_pending = false; }
  const int n_cam = (int)h->cams.size();
  cudaSetDevice(h->device);
  // MCP_BA_SCHUR: 1 (default) pair gathers with TMA, 2 staged pair gathers, 0 row-wise (k_schur_rows; measured slower
  // than the pair kernel at cfg2 -- profiles/README.md -- kept selectable and parity-tested)
  int schur_mode = 1;
  { const char* e = getenv("MCP_BA_SCHUR"); if (e && e[0]) schur_mode = atoi(e); if (getenv("MCP_BA_SCHUR_V1") && getenv("MCP_BA_SCHUR_V1")[0] == '1') schur_mode = 2; }
  if (schur_mode < 0 || schur_mode > 2) schur_mode = 1;

  // host marshalling (ba_prep.hpp): linear passes into pinned staging that is pooled across calls
  static const bool trace = getenv("MCP_BA_LOAD_TRACE") && getenv("MCP_BA_LOAD_TRACE")[0] == '1';
  const auto t_begin = std::chrono::steady_clock::now();
  BaPrep& pr = h->prep;
  {
    // host threads for the marshalling passes: MCP_BA_HOST_THREADS, else three quarters of the cores (<= 12) shared between the ranks
    int want = 0;
    if (const char* e = getenv("MCP_BA_HOST_THREADS")) want = atoi(e);
    if (want <= 0) want = std::min(12, std::max(1, (int)std::thread::hardware_concurrency() * 3 / (4 * std::max(h->world, 1))));
    if (!h->pool || h->pool->size() != want) h->pool.reset(new HostPool(want));
    if (n_meas >= pr.par_min_meas) h->pool->prewake();            // (the workers are spinning by the time the first pass is dispatched)
  }
  int rc = MCP_OK;
#define UP(buf, arr) if (rc == MCP_OK) rc = upload(h, buf, (arr).p, (arr).bytes())
  // called by ba_prepare as soon as the measurement / point / slot arrays are final: their upload overlaps the
  // construction of the work lists
  const std::function<void()> upload_points = [&]() {
    UP(h->b_pose_var, pr.pose_var); UP(h->b_pt_info, pr.pt_info); UP(h->b_pt_var, pr.pt_var);
    UP(h->b_pt_meas_off, pr.pt_meas_off); UP(h->b_pt_slot_off, pr.pt_slot_off); UP(h->b_slot_var, pr.slot_var); UP(h->b_slot_pt, pr.slot_pt);
    UP(h->b_meas_xy, pr.meas_xy); UP(h->b_meas_info, pr.meas_info); UP(h->b_meas_a, pr.meas_a); UP(h->b_meas_b, pr.meas_b);
  };
  int prc = ba_prepare(pr, g_pinned_alloc, n_cam, n_pose, pose_fixed, n_pt, pt_chain, pt_fixed, n_meas, meas_xy, meas_chain,
                       meas_pt, meas_noise, meas_cam, h->rank, h->world, schur_mode == 0, h->pool.get(), &upload_points);
  if (prc == PREP_OK && schur_mode == 0 && pr.max_slots > 32) {          // an entry must fit one staging buffer
    schur_mode = 1;
    pr.rs_ent.n = pr.rs_grp.n = pr.rs_items.n = 0;
  }
  if (prc != PREP_OK) {
    cudaStreamSynchronize(h->stream);                                    // uploads from the staging may be in flight
    set_last_error("%s", pr.err);
    return prc == PREP_INVALID ? MCP_ERR_INVALID : prc == PREP_UNSUPPORTED ? MCP_ERR_UNSUPPORTED : MCP_ERR_CUDA;
  }
  if (rc != MCP_OK) return rc;
  const int npv = pr.npv, nptv = pr.nptv, n_slots = pr.n_slots, max_slots = pr.max_slots, rs_nblk = pr.rs_nblk;
  const int nc = 6 * npv;
  if (nc > chol_max_n()) {
    set_last_error("mcp_ba_load: %d movable poses exceed the dense solver capacity (%d rows)", npv, chol_max_n());
    return MCP_ERR_UNSUPPORTED;
  }
  const int stage_doubles = stage_doubles_for(n_pose, n_cam);
  if (configure_kernels(max_slots, stage_doubles, &h->lin_warps, &h->lin_smem) != 0) {
    set_last_error("mcp_ba_load: a point is observed from %d movable keyframes, more than the kernel supports", max_slots);
    return MCP_ERR_UNSUPPORTED;
  }
  h->part_pt = pr.part_pt; h->part_meas = pr.part_meas;
  h->meas_orig.swap(pr.meas_orig);
  const auto t_prep = std::chrono::steady_clock::now();

  UP(h->b_pt_order, pr.pt_order); UP(h->b_pb_idx, pr.pb_idx); UP(h->b_pb_items, pr.pb_items);
  UP(h->b_rs_ent, pr.rs_ent); UP(h->b_rs_grp, pr.rs_grp); UP(h->b_rs_items, pr.rs_items);
#undef UP
  if (rc != MCP_OK) return rc;
  if ((rc = h->b_mrec.ensure(sizeof(double) * MREC * (size_t)std::max(n_meas, 1)))) return rc;
  // initial state: one copy from the caller's (pageable) arrays, replicated on the device
  const size_t pose_bytes = sizeof(double) * 12 * (size_t)n_pose, pt_bytes = sizeof(double) * 3 * (size_t)n_pt;
  if ((rc = upload(h, h->b_pose0, pose_Rt, pose_bytes))) return rc;
  if ((rc = upload(h, h->b_pt0, pt_xyz, pt_bytes))) return rc;
  // (everything that has to be cleared or replicated on the device is collected here and done by ONE kernel, k_load_init)
  LoadInit li;
  memset(&li, 0, sizeof(li));
  auto zero = [&](void* ptr, size_t bytes) { if (bytes && li.n_zero < LOAD_ZERO_MAX) { li.zero_ptr[li.n_zero] = ptr; li.zero_bytes[li.n_zero] = bytes; li.n_zero++; return true; } return bytes == 0; };
  bool zok = true;
  for (int k = 0; k < N_STATE; k++) {
    if ((rc = h->b_pose[k].ensure(pose_bytes))) return rc;
    if ((rc = h->b_pt[k].ensure(pt_bytes ? pt_bytes : 16))) return rc;
    if ((rc = h->b_chi2[k].ensure(sizeof(double) * (size_t)std::max(n_meas, 1)))) return rc;
    li.pose[k] = h->b_pose[k].as<double>(); li.pt[k] = h->b_pt[k].as<double>();
  }
  li.pose0 = h->b_pose0.as<double>(); li.pt0 = h->b_pt0.as<double>();
  li.pose_doubles = 12 * (size_t)n_pose; li.pt_doubles = 3 * (size_t)n_pt;
  if ((rc = h->b_V.ensure(sizeof(double) * 6 * (size_t)std::max(n_pt, 1)))) return rc;
  if ((rc = h->b_gp.ensure(sizeof(double) * 3 * (size_t)std::max(n_pt, 1)))) return rc;
  if ((rc = h->b_W.ensure(sizeof(double) * 18 * (size_t)std::max(n_slots, 1)))) return rc;
  if ((rc = h->b_Y.ensure(sizeof(double) * 24 * (size_t)std::max(n_slots, 1)))) return rc;
  if ((rc = h->b_R.ensure(sizeof(double) * (36 * (size_t)std::max(n_pt, 1) + 2)))) return rc;      // + the work counter
  // accumulators: [H0 | gc | red(8) | Sm | rm]
  const size_t ncp = (size_t)std::max(nc, 1);
  h->off_H0 = 0; h->off_gc = ncp * ncp; h->off_red = h->off_gc + ncp; h->off_Sm = h->off_red + 16; h->off_rm = h->off_Sm + ncp * ncp;
  h->acc_doubles = h->off_rm + ncp;
  if ((rc = h->b_acc.ensure(sizeof(double) * 2 * h->acc_doubles))) return rc;          // two halves: see run_compute (double-buffered linearisation)
  if (h->world > 1 && (rc = h->b_pack.ensure(sizeof(double) * MAX_CAND * (ncp * (ncp + 1) / 2 + ncp + 16)))) return rc;
  if (h->world > 1 && (rc = p2p_setup(h, (size_t)n_meas, ncp))) return rc;
  if ((rc = h->b_dc.ensure(sizeof(double) * ncp))) return rc;
  if ((rc = h->b_L.ensure(sizeof(double) * chol_tiles_doubles(nc)))) return rc;
  if ((rc = h->b_Linv.ensure(sizeof(double) * chol_inv_doubles(nc)))) return rc;
  if ((rc = h->b_cflags.ensure(sizeof(int) * chol_flag_ints(nc)))) return rc;
  zok = zok && zero(h->b_cflags.p, sizeof(int) * chol_flag_ints(nc));
  // LL tiles are tagged with the launch number (chol_epoch restarts at 1): stale tags of an earlier problem must go
  if ((rc = h->b_Lll.ensure(chol_ll_bytes(nc)))) return rc;
  zok = zok && zero(h->b_Lll.p, chol_ll_bytes(nc));
  h->chol_epoch = 0; h->chol_task_base = 0;
  {
    const size_t sel_bytes = sizeof(unsigned) * (SEL_PASSES * SEL_BINS + 16) + sizeof(unsigned long long) * 2 * (SEL_PASSES + 1) + sizeof(double) * 2 * MAX_CAND;
    if ((rc = h->b_sel.ensure(sel_bytes))) return rc;
    zok = zok && zero(h->b_sel.p, sel_bytes);
  }
  if ((rc = h->b_part.ensure(sizeof(double) * 8 * MAX_PARTIALS))) return rc;
  if ((rc = h->b_ctrl.ensure(sizeof(BaCtrl)))) return rc;
  if ((rc = h->b_flags.ensure(sizeof(int) * ((size_t)std::max(n_meas, 1) + 16)))) return rc;
  if ((size_t)n_meas > h->flags_cap) {
    if (h->flags_host) cudaFreeHost(h->flags_host);
    h->flags_host = nullptr;
    MCP_CUDA_CHECK(cudaMallocHost(&h->flags_host, sizeof(int) * (size_t)(n_meas + n_meas / 4 + 16)));
    h->flags_cap = (size_t)(n_meas + n_meas / 4 + 16);
  }
  zok = zok && zero(h->b_acc.p, sizeof(double) * 2 * h->acc_doubles);
  zok = zok && zero(h->b_dc.p, sizeof(double) * ncp);
  zok = zok && zero(h->b_part.p, sizeof(double) * 8 * MAX_PARTIALS);

  BaDev& d = h->d;
  memset(&d, 0, sizeof(d));
  d.cams = h->b_cams.as<DevCam>();
  d.n_pose = n_pose; d.n_pt = n_pt; d.n_meas = n_meas; d.n_pose_var = npv; d.n_pt_var = nptv; d.nc = nc;
  d.n_slots = n_slots; d.max_slots = max_slots; d.n_cam = n_cam; d.stage_doubles = stage_doubles;
  d.p_lo = h->part_pt[h->rank]; d.p_hi = h->part_pt[h->rank + 1];
  d.m_lo = h->part_meas[h->rank]; d.m_hi = h->part_meas[h->rank + 1];
  d.pose_var = h->b_pose_var.as<int>(); d.pt_info = h->b_pt_info.as<int4>(); d.pt_var = h->b_pt_var.as<int>(); d.pt_order = h->b_pt_order.as<int>();
  d.pt_meas_off = h->b_pt_meas_off.as<int>(); d.pt_slot_off = h->b_pt_slot_off.as<int>(); d.slot_var = h->b_slot_var.as<int>();
  d.meas_xy = h->b_meas_xy.as<double2>(); d.meas_info = h->b_meas_info.as<double>();
  d.meas_a = h->b_meas_a.as<int4>(); d.meas_b = h->b_meas_b.as<int4>();
  for (int k = 0; k < N_STATE; k++) { d.pose[k] = h->b_pose[k].as<double>(); d.pt[k] = h->b_pt[k].as<double>(); d.chi2[k] = h->b_chi2[k].as<double>(); }
  d.V = h->b_V.as<double>(); d.gp = h->b_gp.as<double>(); d.W = h->b_W.as<double>(); d.Y = h->b_Y.as<double>();
  d.slot_pt = h->b_slot_pt.as<int>(); d.slot_lo = pr.pt_slot_off[d.p_lo]; d.slot_hi = pr.pt_slot_off[d.p_hi];
  double* acc = h->b_acc.as<double>();
  d.H0 = acc + h->off_H0; d.gc = acc + h->off_gc; d.Sm = acc + h->off_Sm; d.rm = acc + h->off_rm;
  d.dc = h->b_dc.as<double>(); d.L = h->b_L.as<double>(); d.Linv = h->b_Linv.as<double>(); d.Lll = h->b_Lll.as<uint4>(); d.flags = h->b_cflags.as<int>();
  d.rs_ent = h->b_rs_ent.as<int2>(); d.rs_grp = h->b_rs_grp.as<int>(); d.rs_items = h->b_rs_items.as<int4>(); d.n_rs_items = (int)pr.rs_items.n;
  d.schur_mode = schur_mode; d.rs_nblk = rs_nblk;
  d.mrec = h->b_mrec.as<double>(); d.pb_idx = h->b_pb_idx.as<int>(); d.pb_items = h->b_pb_items.as<int4>(); d.n_pb_items = (int)pr.pb_items.n;
  d.sel_state = h->b_sel.as<unsigned long long>();
  d.sel_hist = reinterpret_cast<unsigned*>(d.sel_state + 2 * (SEL_PASSES + 1));
  d.sel_done = d.sel_hist + SEL_PASSES * SEL_BINS; d.part = h->b_part.as<double>();
  d.spec_sigma = reinterpret_cast<double*>(d.sel_done + 16);
  d.spec_med = d.spec_sigma + MAX_CAND;
  d.ctrl = h->b_ctrl.as<BaCtrl>(); d.outlier_flags = h->b_flags.as<int>();

  for (int q = 1; q < MAX_CAND; q++) {
    McpBa::Cand& cq = h->cand[q];
    cq.d = d;
    cq.d.cand = q;
    cq.chol_epoch = 0; cq.chol_task_base = 0;
    if (q >= h->n_spec) continue;
    const size_t sm_doubles = h->acc_doubles - h->off_Sm;
    if ((rc = cq.b_acc.ensure(sizeof(double) * sm_doubles))) return rc;
    if ((rc = cq.b_dc.ensure(sizeof(double) * ncp))) return rc;
    if ((rc = cq.b_L.ensure(sizeof(double) * chol_tiles_doubles(nc)))) return rc;
    if ((rc = cq.b_Linv.ensure(sizeof(double) * chol_inv_doubles(nc)))) return rc;
    if ((rc = cq.b_cflags.ensure(sizeof(int) * chol_flag_ints(nc)))) return rc;
    if ((rc = cq.b_Lll.ensure(chol_ll_bytes(nc)))) return rc;
    zok = zok && zero(cq.b_Lll.p, chol_ll_bytes(nc));
    if ((rc = cq.b_part.ensure(sizeof(double) * 8 * MAX_PARTIALS))) return rc;
    if ((rc = cq.b_Y.ensure(sizeof(double) * 24 * (size_t)std::max(n_slots, 1)))) return rc;
    zok = zok && zero(cq.b_cflags.p, sizeof(int) * chol_flag_ints(nc));
    zok = zok && zero(cq.b_dc.p, sizeof(double) * ncp);
    zok = zok && zero(cq.b_part.p, sizeof(double) * 8 * MAX_PARTIALS);
    cq.d.Sm = cq.b_acc.as<double>(); cq.d.rm = cq.d.Sm + (h->off_rm - h->off_Sm);
    cq.d.dc = cq.b_dc.as<double>(); cq.d.L = cq.b_L.as<double>(); cq.d.Linv = cq.b_Linv.as<double>(); cq.d.Lll = cq.b_Lll.as<uint4>();
    cq.d.flags = cq.b_cflags.as<int>(); cq.d.part = cq.b_part.as<double>(); cq.d.Y = cq.b_Y.as<double>();
  }
  BaCtrl& c = *h->ctrl_host;
  memset(&c, 0, sizeof(c));
  c.cur = 0;
  c.last_chi2 = 1.7976931348623157e308;   // CheckConvergedResidualAction ctor, src/ChainBundle.cc:1068
  c.min_sigma_sq = h->cfg.min_sigma * h->cfg.min_sigma;
  c.sigma_sq_raw = 0; c.sigma_sq_lim = c.min_sigma_sq; c.sigma_lim = h->cfg.min_sigma;
  c.pct_limit = h->cfg.update_pct_limit; c.rms_limit = h->cfg.update_rms_limit;
  c.max_trials = h->cfg.max_trials_after_failure; c.use_robust = h->cfg.use_robust;
  c.dim = 6 * npv + 3 * nptv;
  for (int q = 0; q < MAX_CAND; q++) c.solve_ok[q] = 1;
  c.sel_n = n_meas; c.sel_rank = n_meas / 2;
  MCP_CUDA_CHECK(cudaMemcpyAsync(d.ctrl, &c, sizeof(c), cudaMemcpyHostToDevice, h->stream));
  if (!zok) { set_last_error("mcp_ba_load: internal: too many clear ranges"); return MCP_ERR_CUDA; }
  launch_load_init(li, h->stream);
  if (schur_mode != 0) {
    // co-visibility lists for the pair-gather Schur kernels (ba_schur.cu), built entirely on the device: count per
    // block pair, scan + work items (k_pair_items), fill.  Only their sizes (upper bounds) are known to the host.
    const int n_pairs = npv * (npv + 1) / 2;
    const size_t max_items = (size_t)(pr.n_inc / 128) + (size_t)n_pairs + 1;
    if ((rc = h->b_paircnt.ensure(sizeof(int) * (size_t)(n_pairs + 2)))) return rc;
    if ((rc = h->b_inc.ensure(sizeof(int2) * (size_t)std::max<long long>(pr.n_inc, 1)))) return rc;
    if ((rc = h->b_items.ensure(sizeof(int4) * max_items + 16))) return rc;
    // on aux_stream, behind everything uploaded so far: ≈ 0.1 ms of device work that overlaps the start of the next Compute
    MCP_CUDA_CHECK(cudaEventRecord(h->ev_up, h->stream));
    MCP_CUDA_CHECK(cudaStreamWaitEvent(h->aux_stream, h->ev_up, 0));
    MCP_CUDA_CHECK(cudaMemsetAsync(h->b_paircnt.p, 0, sizeof(int) * (size_t)(n_pairs + 2), h->aux_stream));
    int* n_items_dev = reinterpret_cast<int*>(h->b_items.as<int4>() + max_items);
    launch_pair_count(d, h->b_paircnt.as<int>(), h->aux_stream);
    launch_pair_items(d, h->b_paircnt.as<int>(), h->b_items.as<int4>(), n_items_dev, h->aux_stream);
    launch_pair_fill(d, h->b_paircnt.as<int>(), h->b_inc.as<int2>(), h->aux_stream);
    MCP_CUDA_CHECK(cudaEventRecord(h->ev_pairs, h->aux_stream));
    h->pairs_pending = true;
    d.inc = h->b_inc.as<int2>(); d.items = h->b_items.as<int4>(); d.n_items_dev = n_items_dev; d.max_items = (int)max_items;
    for (int q = 1; q < MAX_CAND; q++) { h->cand[q].d.inc = d.inc; h->cand[q].d.items = d.items; h->cand[q].d.n_items_dev = d.n_items_dev; h->cand[q].d.max_items = d.max_items; }
  }
  const auto t_enq = std::chrono::steady_clock::now();
  // no wait for the device here (MCP_BA_LOAD_SYNC=1 / the trace restore it): the uploads come from pinned staging owned by the
  // handle, the caller's arrays have been consumed, and every later call is ordered behind this work on the handle's stream
  static const bool load_sync = getenv("MCP_BA_LOAD_SYNC") && getenv("MCP_BA_LOAD_SYNC")[0] == '1';
  if (trace || load_sync) { MCP_CUDA_CHECK(cudaStreamSynchronize(h->stream)); MCP_CUDA_CHECK(cudaStreamSynchronize(h->aux_stream)); }
  MCP_CUDA_CHECK(cudaGetLastError());
  if (trace) {
    const auto t_end = std::chrono::steady_clock::now();
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    fprintf(stderr, "mcp_ba_load: marshal %.3f ms, enqueue %.3f ms, drain %.3f ms\n", ms(t_begin, t_prep), ms(t_prep, t_enq), ms(t_enq, t_end));
  }
  h->outliers.clear();
  h->loaded = true;
  return MCP_OK;
}

// ------------------------------------------------------------------------------------------
// driver loop
// ------------------------------------------------------------------------------------------
#define NCCL_CHECK(expr)                                                                     \
  do {                                                                                       \
    ncclResult_t _r = (expr);                                                                \
    if (_r != ncclSuccess) { set_last_error("%s failed: %s", #expr, ncclGetErrorString(_r)); return MCP_ERR_NCCL; } \
  } while (0)


// ------------------------------------------------------------------------------------------
// peer-memory exchange buffers (ba_p2p.cu)
// ------------------------------------------------------------------------------------------
static void p2p_release(McpBa* h)
{
  if (!h->xbuf) return;
  cudaStreamSynchronize(h->stream);
  for (int p = 0; p < h->world && p < 8; p++)
    if (p != h->rank && h->peers.base[p]) cudaIpcCloseMemHandle(h->peers.base[p]);
  cudaFree(h->xbuf);
  h->xbuf = nullptr; h->xbuf_lines = 0; h->p2p = false;
  memset(&h->peers, 0, sizeof(h->peers));
}

// (Re)allocates the exchange buffer when the problem outgrew it and maps every peer's buffer into this process.  Collective:
// every rank takes the same decisions (same problem sizes), the IPC handles travel through one NCCL all-gather, and a rank
// that cannot map a peer makes ALL ranks fall back to the NCCL path.  MCP_BA_P2P=0 disables.
static int p2p_setup(McpBa* h, size_t n_meas, size_t ncp)
{
  static const bool off = [] { const char* e = getenv("MCP_BA_P2P"); return e && e[0] == '0'; }();
  if (off || h->world > 8) { h->p2p = false; return MCP_OK; }
  const int W = h->world;
  const size_t ntri = ncp * (ncp + 1) / 2;
  const size_t cnt[McpBa::X_N] = { ntri + ncp + 16, (size_t)MAX_CAND * (ntri + ncp), 16, n_meas };
  size_t lines[McpBa::X_N], total = 0;
  for (int k = 0; k < McpBa::X_N; k++) {
    lines[k] = (k == McpBa::X_CHI) ? cnt[k] : 2 * (size_t)W * p2p_slice(cnt[k], W);
    lines[k] = (lines[k] + 63) & ~(size_t)63;
    total += 2 * lines[k];
  }
  bool same = h->xbuf && total <= h->xbuf_lines;
  for (int k = 0; k < McpBa::X_N && same; k++) same = (lines[k] == h->x_lines[k]);
  if (same) return MCP_OK;                                   // pooled across loads of the same shape
  p2p_release(h);
  MCP_CUDA_CHECK(cudaMalloc(&h->xbuf, total * sizeof(uint4)));
  MCP_CUDA_CHECK(cudaMemsetAsync(h->xbuf, 0, total * sizeof(uint4), h->stream));
  h->xbuf_lines = total;
  size_t o = 0;
  for (int k = 0; k < McpBa::X_N; k++) { h->x_off[k] = o; h->x_lines[k] = lines[k]; o += 2 * lines[k]; h->x_calls[k] = 0; }
  // exchange the IPC handles
  cudaIpcMemHandle_t mine;
  int fail = cudaIpcGetMemHandle(&mine, h->xbuf) == cudaSuccess ? 0 : 1;
  if (fail) (void)cudaGetLastError();
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  unsigned char* stage = nullptr;
  MCP_CUDA_CHECK(cudaMalloc(&stage, 64 * (size_t)W + 16));
  MCP_CUDA_CHECK(cudaMemcpyAsync(stage + 64 * (size_t)h->rank, &mine, 64, cudaMemcpyHostToDevice, h->stream));
  NCCL_CHECK(ncclAllGather(stage + 64 * (size_t)h->rank, stage, 64, ncclChar, h->comm, h->stream));
  std::vector<cudaIpcMemHandle_t> all((size_t)W);
  MCP_CUDA_CHECK(cudaMemcpyAsync(all.data(), stage, 64 * (size_t)W, cudaMemcpyDeviceToHost, h->stream));
  MCP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  memset(&h->peers, 0, sizeof(h->peers));
  for (int p = 0; p < W && !fail; p++) {
    if (p == h->rank) { h->peers.base[p] = h->xbuf; continue; }
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, all[(size_t)p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { (void)cudaGetLastError(); fail = 1; break; }
    h->peers.base[p] = reinterpret_cast<uint4*>(ptr);
  }
  // every rank must agree on the path
  int* flag = reinterpret_cast<int*>(stage + 64 * (size_t)W);
  MCP_CUDA_CHECK(cudaMemcpyAsync(flag, &fail, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  NCCL_CHECK(ncclAllReduce(flag, flag, 1, ncclInt, ncclMax, h->comm, h->stream));
  int any = 0;
  MCP_CUDA_CHECK(cudaMemcpyAsync(&any, flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  MCP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  cudaFree(stage);
  if (any) {
    static bool told = false;
    if (!told) { fprintf(stderr, "mcptam_b200: peer memory could not be mapped on every rank; the exchanges use NCCL\n"); told = true; }
    p2p_release(h);
    return MCP_OK;
  }
  h->p2p = true;
  return MCP_OK;
}

// one exchange = one region, alternating parity, a tag that never repeats
static inline size_t p2p_next(McpBa* h, int region, unsigned* tag)
{
  const unsigned par = h->x_calls[region]++ & 1u;
  *tag = ++h->x_tag;
  return h->x_off[region] + (size_t)par * h->x_lines[region];
}

namespace {

struct Prof {
  McpBa* h; int cat; cudaEvent_t a = nullptr, b = nullptr;
  Prof(McpBa* h_, int cat_) : h(h_), cat(cat_)
  {
    h->launches++;
    if (h->profiling) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, h->stream); }
  }
  ~Prof() { if (h->profiling) { cudaEventRecord(b, h->stream); h->evs.push_back({ cat, a, b }); } }
};

struct TlScope {
  McpBa* h; cudaStream_t s; cudaEvent_t b = nullptr;
  TlScope(McpBa* h_, const char* name, int sid, cudaStream_t s_) : h(h_), s(s_)
  {
    if (!h->timeline) return;
    cudaEvent_t a;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, s);
    h->tl.push_back({ name, sid, a, b });
  }
  ~TlScope() { if (b) cudaEventRecord(b, s); }
};

int sync_ctrl(McpBa* h)
{
  MCP_CUDA_CHECK(cudaMemcpyAsync(h->ctrl_host, h->d.ctrl, sizeof(BaCtrl), cudaMemcpyDeviceToHost, h->stream));
  MCP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return MCP_OK;
}
int push_ctrl(McpBa* h)
{
  MCP_CUDA_CHECK(cudaMemcpyAsync(h->d.ctrl, h->ctrl_host, sizeof(BaCtrl), cudaMemcpyHostToDevice, h->stream));
  return MCP_OK;
}

// make chi2[which] / pt[which] complete on every rank (each rank owns a contiguous range)
int allgather_ranges(McpBa* h, double* base, const std::vector<int>& bounds, int stride)
{
  if (h->world == 1) return MCP_OK;
  if (h->p2p && stride == 1 && (size_t)bounds[h->world] <= h->x_lines[McpBa::X_CHI]) {
    // chi2 shards: one peer-memory kernel per rank
    unsigned tag;
    const size_t off = p2p_next(h, McpBa::X_CHI, &tag);
    launch_p2p_allgather(h->d, base, 1, bounds.data(), h->peers, off, h->rank, h->world, tag, h->stream);
    return MCP_OK;
  }
  NCCL_CHECK(ncclGroupStart());
  for (int r = 0; r < h->world; r++) {
    const size_t off = (size_t)bounds[r] * stride, cnt = (size_t)(bounds[r + 1] - bounds[r]) * stride;
    if (cnt) NCCL_CHECK(ncclBroadcast(base + off, base + off, cnt, ncclDouble, r, h->comm, h->stream));
  }
  NCCL_CHECK(ncclGroupEnd());
  return MCP_OK;
}

int eval_state(McpBa* h, int which, double* err_out)
{
  { Prof p(h, C_BACKSUB); launch_backsub_eval(h->d, 0, which, err_out, h->stream); }
  return MCP_OK;
}

}  // namespace

static int run_compute(McpBa* h, volatile const uint8_t* abort_flag, int n_iter, double user_lambda, McpBaStats* st,
                       bool single_step, double step_lambda, double step_sigma_sq)
{
  BaDev& d = h->d;
  BaCtrl& c = *h->ctrl_host;
  cudaStream_t s = h->stream;
  double* const acc_base = h->b_acc.as<double>();
  double* acc = acc_base;
  double* red = acc_base + h->off_red;                 // (the scalar slots of half 0 serve both halves)
  const bool multi = h->world > 1;
  int rc;
  // [H0 | gc | Sm | rm] exists twice.  With look-ahead the next outer iteration linearises into the other half, which
  // k_pose_blocks cleared an iteration earlier: no clearing kernel between k_lm_control and k_linearize.
  int ab = 0;
  auto set_acc = [&](int b) {
    acc = acc_base + (size_t)b * h->acc_doubles;
    d.H0 = acc + h->off_H0; d.gc = acc + h->off_gc; d.Sm = acc + h->off_Sm; d.rm = acc + h->off_rm;
    for (int q = 1; q < MAX_CAND; q++) { h->cand[q].d.H0 = d.H0; h->cand[q].d.gc = d.gc; }
  };
  set_acc(0);
  // Single GPU: the caller's flag is polled between trial rounds like the reference polls it between iterations.
  // Several ranks: every rank must take the same branch (each round issues collectives), so the flag is only ever acted
  // on through a value all ranks agreed on -- it rides on the per-round all-reduce of the trial sums (and on one
  // all-reduce before the first round).
  bool agreed_abort = false;
  auto local_flag = [&]() { return abort_flag && *abort_flag; };
  auto aborted = [&]() { return multi ? agreed_abort : local_flag(); };

  { const char* e = getenv("MCP_BA_TIMELINE"); h->timeline = e && e[0] == '1'; }
  if ((rc = sync_ctrl(h))) return rc;
  c.need_lambda_init = 1; c.user_lambda = user_lambda; c.iter = 0; c.conv_mag = 0; c.conv_res = 0; c.total_trials = 0;
  c.terminate = 0; c.qmax = 0; for (int q = 0; q < MAX_CAND; q++) c.solve_ok[q] = 1; c.stop_trials = 0; c.accepted = 0; c.n_outliers = 0; c.abort_agreed = 0; c.med_hint = 0;
  if ((rc = push_ctrl(h))) return rc;
  h->outliers.clear();

  // errors of the initial state, sigma, "BEFORE" chi2 (src/ChainBundle.cc:1317-1323)
  eval_state(h, -1, nullptr);
  if ((rc = allgather_ranges(h, d.chi2[c.cur], h->part_meas, 1))) return rc;
  if (h->cfg.use_robust && !(single_step && step_sigma_sq >= 0)) { Prof p(h, C_SELECT); h->launches += launch_select_sigma(d, -1, 0, s) - 1; }
  if (single_step && step_sigma_sq >= 0) {
    if ((rc = sync_ctrl(h))) return rc;
    c.sigma_sq_raw = step_sigma_sq;
    c.sigma_sq_lim = step_sigma_sq < c.min_sigma_sq ? c.min_sigma_sq : step_sigma_sq;
    c.sigma_lim = std::sqrt(c.sigma_sq_lim);
    if ((rc = push_ctrl(h))) return rc;
  }
  if (single_step) {
    if ((rc = sync_ctrl(h))) return rc;
    c.lambda = step_lambda; c.need_lambda_init = 0; c.ni = 2;
    if ((rc = push_ctrl(h))) return rc;
  }
  int n_bs = 0, n_lin = 0;
  // (the errors are in d.chi2 already: the robust sum is a pass over them, not a second reprojection)
  n_bs = launch_robust_sum(d, -1, s); h->launches++;
  { Prof p(h, C_OTHER); launch_reduce_partials(d, 0, n_bs, red, s); }
  if (multi) NCCL_CHECK(ncclAllReduce(red + 1, red + 1, 1, ncclDouble, ncclSum, h->comm, s));
  // the figure is only reported: it travels to pinned memory in stream order and is read when the Compute ends
  MCP_CUDA_CHECK(cudaMemcpyAsync(&h->abort_word[2], red + 1, sizeof(double), cudaMemcpyDeviceToHost, s));
  if (single_step) { MCP_CUDA_CHECK(cudaStreamSynchronize(s)); if (st) st->chi2_before = h->abort_word[2]; }

  if (multi) {
    // red[15] is free between rounds
    const double mine = local_flag() ? 1.0 : 0.0;
    double all = 0;
    MCP_CUDA_CHECK(cudaMemcpyAsync(red + 15, &mine, sizeof(double), cudaMemcpyHostToDevice, s));
    NCCL_CHECK(ncclAllReduce(red + 15, red + 15, 1, ncclDouble, ncclSum, h->comm, s));
    MCP_CUDA_CHECK(cudaMemcpyAsync(&all, red + 15, sizeof(double), cudaMemcpyDeviceToHost, s));
    MCP_CUDA_CHECK(cudaStreamSynchronize(s));
    agreed_abort = all > 0;
  }
  MCP_CUDA_CHECK(cudaEventRecord(h->ev0, s));
  int counter = 0;
  bool ok = true, local_abort = false;
  // Look-ahead: most trial rounds end their outer iteration, so the next iteration's sigma selection and linearisation
  // are enqueued right behind k_lm_control -- predicated on the device on that kernel's verdict -- and run while the
  // host is still reading the control block back.  (Single GPU only: the multi-GPU path has collectives in between.)
  static const bool ahead_env = [] { const char* e = getenv("MCP_BA_LOOKAHEAD"); return !(e && e[0] == '0'); }();
  const bool can_look_ahead = ahead_env && !multi && !single_step && !h->profiling;
  // Speculative sigma: the Huber sigma^2 of every candidate's trial state is computed next to the trial (own stream), so the
  // next outer iteration does not start with a median selection on its critical path.  MCP_BA_SPEC_SIGMA=0 disables.
  static const bool spec_env = [] { const char* e = getenv("MCP_BA_SPEC_SIGMA"); return !(e && e[0] == '0'); }();
  const bool spec_sigma = spec_env && can_look_ahead && h->cfg.use_robust && select_spec_possible(d);
  int spec_pending = 0;                                   // candidates whose speculative selection is in flight
  bool next_iteration_started = false;
  if (can_look_ahead) MCP_CUDA_CHECK(cudaMemsetAsync(acc_base + h->acc_doubles, 0, sizeof(double) * h->acc_doubles, s));   // half 1 starts clean
  // speculate on the trials g2o would run after rejections of this one (lambda * ni, * 2ni, ...): most outer
  // iterations reject their first trial(s), so the candidates are evaluated concurrently on side streams
  const int n_cand = (single_step || h->profiling) ? 1 : (multi ? std::min(h->n_spec, h->n_spec_multi) : h->n_spec);
  // 2 or 3 candidates: ONE pass over the co-visibility lists reduces every candidate's camera system
  // (k_schur_pairs_multi); otherwise one reduction per candidate on its own stream
  const bool fused = h->fuse_schur && d.schur_mode == 1 && n_cand >= 2 && n_cand <= 3;
  const size_t sm_doubles = h->acc_doubles - h->off_Sm;
  auto schur_args = [&](bool first_round) {
    SchurMulti mc;
    memset(&mc, 0, sizeof(mc));
    mc.n_cand = n_cand; mc.R = h->b_R.as<double>(); mc.sm_doubles = sm_doubles;
    mc.next_item = reinterpret_cast<int*>(mc.R + 36 * (size_t)std::max(d.n_pt, 1));
    mc.Sm[0] = d.Sm; mc.rm[0] = d.rm;
    mc.zero_mask = first_round ? 0 : 1;                  // candidate 0 shares the accumulator zeroed before the linearisation
    for (int q = 1; q < n_cand; q++) { mc.Sm[q] = h->cand[q].d.Sm; mc.rm[q] = h->cand[q].d.rm; mc.zero_mask |= 1 << q; }
    schur_multi_prepare(d, mc);
    return mc;
  };
  // Fold: when lambda is already valid on the device (every linearisation but the first), the point records of the first
  // trial round are formed by extra blocks of k_pose_blocks and k_schur_vinv_multi leaves the critical path.
  static const bool fold_env = [] { const char* e = getenv("MCP_BA_FOLD_VINV"); return !(e && e[0] == '0'); }();
  bool records_ready = false;
  for (int it = 0; it < n_iter && !local_abort && !aborted() && ok; it++) {
    if (!next_iteration_started) {
      if (it > 0) {
        { TlScope t(h, "ag_chi2", 0, s); if ((rc = allgather_ranges(h, d.chi2[c.cur], h->part_meas, 1))) return rc; }
        if (h->cfg.use_robust) { Prof p(h, C_SELECT); TlScope t(h, "select", 0, s); h->launches += launch_select_sigma(d, -1, 0, s) - 1; }
      }
      MCP_CUDA_CHECK(cudaMemsetAsync(acc, 0, sizeof(double) * h->acc_doubles, s));
      const bool fold = fold_env && fused && !c.need_lambda_init;
      const SchurMulti mcf = schur_args(true);
      { Prof p(h, C_LIN); TlScope t(h, "linearize", 0, s); n_lin = launch_linearize(d, h->lin_warps, h->lin_smem, s, fold ? &mcf : nullptr); h->launches++; }
      records_ready = fold;
    }
    next_iteration_started = false;
    if (multi) {
      launch_reduce_partials(d, n_lin, 0, red, s); h->launches++;
      {
        // [H0 | gc | scalar sums] travel as the packed upper triangle + tail (k_tri_pack)
        TlScope t(h, "ar_H0", 0, s);
        const int tail = (int)(h->off_Sm - h->off_gc);
        const size_t cnt = (size_t)d.nc * (d.nc + 1) / 2 + tail;
        double* pk = h->b_pack.as<double>();
        double* src[1] = { acc };
        launch_tri_pack(src, 1, pk, d.nc, tail, false, s);
        if (h->p2p) { unsigned tag; const size_t off = p2p_next(h, McpBa::X_H0, &tag); launch_p2p_allreduce(pk, cnt, h->peers, off, h->rank, h->world, tag, s); }
        else NCCL_CHECK(ncclAllReduce(pk, pk, cnt, ncclDouble, ncclSum, h->comm, s));
        launch_tri_pack(src, 1, pk, d.nc, tail, true, s);
        h->launches += 2;
      }
    }
    if (c.need_lambda_init) {
      { Prof p(h, C_OTHER); launch_lambda_init(d, s); }
      if (multi) NCCL_CHECK(ncclAllReduce(d.part + PART_MAXDIAG * MAX_PARTIALS, d.part + PART_MAXDIAG * MAX_PARTIALS, 1, ncclDouble, ncclMax, h->comm, s));
      { Prof p(h, C_OTHER); launch_lambda_apply(d, s); }
      c.need_lambda_init = 0;
    }
    bool first = true;
    for (;;) {
      if (h->pairs_pending) { MCP_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_pairs, 0)); h->pairs_pending = false; }   // (lists of this load)
      CandParts parts;
      for (int q = 0; q < MAX_CAND; q++) parts.p[q] = d.part;
      for (int q = 0; q < spec_pending; q++) MCP_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_sel[q], 0));   // (their chi2 buffers are about to be rewritten)
      spec_pending = 0;
      if (!first && !fused) MCP_CUDA_CHECK(cudaMemsetAsync(d.Sm, 0, sizeof(double) * sm_doubles, s));
      if (n_cand > 1 && !fused) MCP_CUDA_CHECK(cudaEventRecord(h->ev_ready, s));
      if (fused) {
        const SchurMulti mc = schur_args(first);
        const bool ready = first && records_ready;
        records_ready = false;
        { Prof p(h, C_SCHUR); TlScope t(h, "schur", 0, s); launch_schur_multi(d, mc, s, ready); h->launches += ready ? 1 : 2; }
        if (!multi) MCP_CUDA_CHECK(cudaEventRecord(h->ev_ready, s));
      } else {
        Prof p(h, C_SCHUR); TlScope t(h, "schur", 0, s); launch_schur_gather(d, s); h->launches++;
      }
      if (multi) {
        // every candidate's reduced camera system goes through ONE grouped all-reduce (one NCCL launch) per round
        for (int q = 1; q < n_cand && !fused; q++) {
          McpBa::Cand& cq = h->cand[q];
          MCP_CUDA_CHECK(cudaStreamWaitEvent(cq.stream, h->ev_ready, 0));
          MCP_CUDA_CHECK(cudaMemsetAsync(cq.d.Sm, 0, sizeof(double) * sm_doubles, cq.stream));
          launch_schur_gather(cq.d, cq.stream);
          MCP_CUDA_CHECK(cudaEventRecord(cq.ev_schur, cq.stream));
          MCP_CUDA_CHECK(cudaStreamWaitEvent(s, cq.ev_schur, 0));
          h->launches += 2;
        }
        {
          // every candidate's [Sm | rm] as packed triangle + tail, ONE all-reduce for all of them
          TlScope t(h, "ar_Sm", 0, s);
          const size_t cnt = (size_t)d.nc * (d.nc + 1) / 2 + d.nc;
          double* pk = h->b_pack.as<double>();
          double* src[MAX_CAND];
          for (int q = 0; q < n_cand; q++) src[q] = q ? h->cand[q].d.Sm : d.Sm;
          launch_tri_pack(src, n_cand, pk, d.nc, d.nc, false, s);
          if (h->p2p) { unsigned tag; const size_t off = p2p_next(h, McpBa::X_SM, &tag); launch_p2p_allreduce(pk, cnt * n_cand, h->peers, off, h->rank, h->world, tag, s); }
          else NCCL_CHECK(ncclAllReduce(pk, pk, cnt * n_cand, ncclDouble, ncclSum, h->comm, s));
          launch_tri_pack(src, n_cand, pk, d.nc, d.nc, true, s);
          h->launches += 2;
        }
        if (n_cand > 1) MCP_CUDA_CHECK(cudaEventRecord(h->ev_red, s));
      }
      { Prof p(h, C_SOLVE); TlScope t(h, "solve", 0, s); launch_chol_solve(d, ++h->chol_epoch, h->n_sms / n_cand, &h->chol_task_base, s); }
      { Prof p(h, C_BACKSUB); TlScope t(h, "backsub", 0, s); n_bs = launch_backsub_eval(d, 1, -1, nullptr, s); }
      if (spec_sigma) {
        MCP_CUDA_CHECK(cudaEventRecord(h->ev_bs[0], s));
        MCP_CUDA_CHECK(cudaStreamWaitEvent(h->sel_stream[0], h->ev_bs[0], 0));
        { TlScope t(h, "spec_select", 4, h->sel_stream[0]); launch_select_spec(d, (c.cur + 1) % N_STATE, h->sel_stream[0]); }
        MCP_CUDA_CHECK(cudaEventRecord(h->ev_sel[0], h->sel_stream[0]));
        h->launches++;
      }
      for (int q = 1; q < n_cand; q++) {
        McpBa::Cand& cq = h->cand[q];
        parts.p[q] = cq.d.part;
        if (multi) {
          MCP_CUDA_CHECK(cudaStreamWaitEvent(cq.stream, h->ev_red, 0));
        } else {
          MCP_CUDA_CHECK(cudaStreamWaitEvent(cq.stream, h->ev_ready, 0));
          if (!fused) {
            MCP_CUDA_CHECK(cudaMemsetAsync(cq.d.Sm, 0, sizeof(double) * sm_doubles, cq.stream));
            { TlScope t(h, "schur", q, cq.stream); launch_schur_gather(cq.d, cq.stream); }
            h->launches += 2;
          }
        }
        { TlScope t(h, "solve", q, cq.stream); launch_chol_solve(cq.d, ++cq.chol_epoch, h->n_sms / n_cand, &cq.chol_task_base, cq.stream); }
        { TlScope t(h, "backsub", q, cq.stream); launch_backsub_eval(cq.d, 1, -1, nullptr, cq.stream); }
        h->launches += 2;
        if (spec_sigma) {
          MCP_CUDA_CHECK(cudaEventRecord(h->ev_bs[q], cq.stream));
          MCP_CUDA_CHECK(cudaStreamWaitEvent(h->sel_stream[q], h->ev_bs[q], 0));
          { TlScope t(h, "spec_select", 4 + q, h->sel_stream[q]); launch_select_spec(cq.d, (c.cur + 1 + q) % N_STATE, h->sel_stream[q]); }
          MCP_CUDA_CHECK(cudaEventRecord(h->ev_sel[q], h->sel_stream[q]));
          h->launches++;
        }
        if (multi) { launch_reduce_partials(cq.d, 0, n_bs, red + 3 * q, cq.stream); h->launches++; }
        MCP_CUDA_CHECK(cudaEventRecord(cq.ev_done, cq.stream));
        h->spec_rounds++;
      }
      if (multi) { h->abort_word[0] = local_flag() ? 1.0 : 0.0; launch_reduce_partials(d, 0, n_bs, red, s, h->abort_word, 1 + 3 * n_cand); h->launches++; }
      for (int q = 1; q < n_cand; q++) MCP_CUDA_CHECK(cudaStreamWaitEvent(s, h->cand[q].ev_done, 0));
      if (multi) {
        // the trial sums of every candidate + this rank's view of the abort flag (slot 1 + 3 n_cand)
        TlScope t(h, "ar_red", 0, s);
        if (h->p2p) { unsigned tag; const size_t off = p2p_next(h, McpBa::X_RED, &tag); launch_p2p_allreduce(red + 1, (size_t)(3 * n_cand + 1), h->peers, off, h->rank, h->world, tag, s); }
        else NCCL_CHECK(ncclAllReduce(red + 1, red + 1, 3 * n_cand + 1, ncclDouble, ncclSum, h->comm, s));
      }
      { Prof p(h, C_CONTROL); TlScope t(h, "control", 0, s); launch_lm_control(d, parts, n_cand, n_lin, n_bs, multi ? red : nullptr, first ? 1 : 0, s); }
      first = false;
      if (spec_sigma) spec_pending = n_cand;
      const bool ahead = can_look_ahead && it + 1 < n_iter;
      if (ahead) {
        // the control block is read back on a side stream so that the copy does not queue behind the look-ahead kernels
        MCP_CUDA_CHECK(cudaEventRecord(h->ev_ctrl, s));
        MCP_CUDA_CHECK(cudaStreamWaitEvent(h->copy_stream, h->ev_ctrl, 0));
        MCP_CUDA_CHECK(cudaMemcpyAsync(h->ctrl_host, h->d.ctrl, sizeof(BaCtrl), cudaMemcpyDeviceToHost, h->copy_stream));
        // The next iteration linearises into the other accumulator half (clean since the iteration before this one) and its
        // k_pose_blocks clears the half this iteration used.  With speculative sigma the linearisation itself adopts the
        // accepted candidate's sigma^2: nothing runs between k_lm_control and k_linearize.
        double* const used_half = acc;
        set_acc(ab ^ 1);
        BaDev d_ahead = d;
        d_ahead.ahead = 1;
        if (spec_sigma) {
          for (int q = 0; q < n_cand; q++) MCP_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_sel[q], 0));
          d_ahead.pick_sigma = 1;
        } else if (h->cfg.use_robust) { TlScope t(h, "select", 0, s); h->launches += launch_select_sigma(d_ahead, -1, 0, s); }
        {
          const bool fold = fold_env && fused;
          const SchurMulti mcf = schur_args(true);
          TlScope t(h, "linearize", 0, s);
          n_lin = launch_linearize(d_ahead, h->lin_warps, h->lin_smem, s, fold ? &mcf : nullptr, used_half, h->acc_doubles);
          records_ready = fold;
        }
        h->launches += 2;
        MCP_CUDA_CHECK(cudaStreamSynchronize(h->copy_stream));
        // (the device skipped these launches unless this round closed the iteration by accepting a trial)
        if (c.stop_trials && !c.terminate && !c.conv_mag && !c.conv_res) ab ^= 1; else set_acc(ab);
      } else if ((rc = sync_ctrl(h))) return rc;
      if (multi) agreed_abort = agreed_abort || c.abort_agreed > 0;
      if (c.cand_used > 1) h->spec_used++;
      if (single_step) break;
      if (c.stop_trials) { next_iteration_started = ahead && !c.terminate && !c.conv_mag && !c.conv_res; break; }
      if (aborted()) {
        // the trial loop ends on terminate(); close the outer iteration bookkeeping like g2o does
        c.iter++; c.total_trials += c.qmax; c.qmax = 0;
        if ((rc = push_ctrl(h))) return rc;
        break;
      }
    }
    if (single_step) return MCP_OK;
    counter++;
    if (c.terminate) ok = false;
    if (c.conv_mag || c.conv_res) local_abort = true;
  }
  for (int q = 0; q < spec_pending; q++) MCP_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_sel[q], 0));
  MCP_CUDA_CHECK(cudaEventRecord(h->ev1, s));
  if (h->timeline) {
    cudaDeviceSynchronize();
    for (const McpBa::Tl& e : h->tl) {
      float ta = 0, tb = 0;
      cudaEventElapsedTime(&ta, h->ev0, e.a); cudaEventElapsedTime(&tb, h->ev0, e.b);
      fprintf(stderr, "TL %s %d %.1f %.1f\n", e.name, e.sid, 1e3 * ta, 1e3 * tb);
      cudaEventDestroy(e.a); cudaEventDestroy(e.b);
    }
    h->tl.clear();
  }

  // "AFTER" block (src/ChainBundle.cc:1338-1345): errors + sigma of the final state
  eval_state(h, -1, nullptr);
  if ((rc = allgather_ranges(h, d.chi2[c.cur], h->part_meas, 1))) return rc;
  if (h->cfg.use_robust) { Prof p(h, C_SELECT); h->launches += launch_select_sigma(d, -1, 0, s) - 1; }
  n_bs = launch_robust_sum(d, -1, s); h->launches++;
  launch_reduce_partials(d, 0, n_bs, red, s); h->launches++;
  if (multi) NCCL_CHECK(ncclAllReduce(red + 1, red + 1, 1, ncclDouble, ncclSum, h->comm, s));
  MCP_CUDA_CHECK(cudaMemcpyAsync(&h->abort_word[3], red + 1, sizeof(double), cudaMemcpyDeviceToHost, s));
  if ((rc = allgather_ranges(h, d.pt[c.cur], h->part_pt, 3))) return rc;
  // Tukey outliers (src/ChainBundle.cc:1368-1399) are flagged before the one host synchronisation of this block.  With the
  // Huber kernel on, the selection above already left the Tukey sigma^2 (same values, same median); the flags come back as
  // a compact list of original indices -- its head rides in the same copy as a count, the tail follows only if it is long
  const bool want_tukey = h->cfg.use_tukey && d.n_meas > 0;
  const int head = std::min(d.n_meas, 4095);
  if (want_tukey) {
    if (!h->cfg.use_robust) { Prof p(h, C_SELECT); h->launches += launch_select_sigma(d, -1, 1, s) - 1; }
    MCP_CUDA_CHECK(cudaMemsetAsync(d.outlier_flags, 0, sizeof(int), s));
    { Prof p(h, C_OTHER); launch_tukey_flags(d, s); }
    MCP_CUDA_CHECK(cudaMemcpyAsync(h->flags_host, d.outlier_flags, sizeof(int) * (size_t)(1 + head), cudaMemcpyDeviceToHost, s));
  }
  if ((rc = sync_ctrl(h))) return rc;
  const double chi_after = h->abort_word[3];
  if (st) st->chi2_before = h->abort_word[2];

  const bool converged = c.conv_mag || c.conv_res;
  const bool abort_now = local_abort || aborted();
  const bool external_abort = abort_now && !converged;
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  if (st) {
    st->iterations = counter; st->total_trials = c.total_trials; st->converged = converged ? 1 : 0;
    st->hit_max_iter = (counter == n_iter) ? 1 : 0; st->sigma_sq = c.sigma_sq_raw; st->lambda = c.lambda;
    st->chi2_after = chi_after; st->mean_chi2 = d.n_meas ? chi_after / d.n_meas : 0; st->max_cov = 1.7976931348623157e308;
    st->gpu_ms = ms; st->kernel_launches = h->launches; st->n_outliers = 0;
  }
  if (counter == 0 && !external_abort) return -1;
  if (counter == 0 && abort_now) return 0;

  if (want_tukey) {
    const int n_out = h->flags_host[0];
    if (n_out > head) {
      MCP_CUDA_CHECK(cudaMemcpyAsync(h->flags_host + 1 + head, d.outlier_flags + 1 + head, sizeof(int) * (size_t)(n_out - head), cudaMemcpyDeviceToHost, s));
      MCP_CUDA_CHECK(cudaStreamSynchronize(s));
    }
    h->outliers.assign(h->flags_host + 1, h->flags_host + 1 + n_out);
    std::sort(h->outliers.begin(), h->outliers.end());
  }
  // median point-depth covariance (src/ChainBundle.cc:1401-1448): attempted only with < 3 movable poses
  double max_cov = 0;                               // the reference's "computeMarginals() failed" value
  if (d.n_pose_var < 3 && !multi) {
    if (d.n_pt_var == 0) max_cov = 1.7976931348623157e308;
    else {
      if ((rc = h->b_tmp.ensure(sizeof(double) * ((size_t)d.n_pt_var + 2)))) return rc;
      BaDev dm = d;
      dm.cand = -1;                                 // lambda = 0: the stored Hessian of the last linearisation
      MCP_CUDA_CHECK(cudaMemsetAsync(d.Sm, 0, sizeof(double) * (h->acc_doubles - h->off_Sm), s));
      launch_schur_gather(dm, s);
      launch_marginals(dm, h->b_tmp.as<double>(), s);
      BaDev ds = d;
      ds.n_meas = d.n_pt_var;
      for (int k = 0; k < N_STATE; k++) ds.chi2[k] = h->b_tmp.as<double>();
      h->launches += 3 + launch_select_sigma(ds, 0, 2, s);
      if ((rc = sync_ctrl(h))) return rc;
      max_cov = c.marg_fail ? 0.0 : c.median_out;
    }
  }
  if (st) { st->n_outliers = (int)h->outliers.size(); st->max_cov = max_cov; st->kernel_launches = h->launches; }
  return counter;
}

static void resolve_profile(McpBa* h)
{
  memset(&h->timing, 0, sizeof(h->timing));
  double* ms[C_N] = { &h->timing.ms_select, &h->timing.ms_linearize, &h->timing.ms_schur, &h->timing.ms_solve,
                      &h->timing.ms_backsub, &h->timing.ms_control, &h->timing.ms_other };
  int32_t* cnt[C_N] = { &h->timing.n_select, &h->timing.n_linearize, &h->timing.n_schur, &h->timing.n_solve,
                        &h->timing.n_backsub, &h->timing.n_control, &h->timing.n_other };
  for (auto& e : h->evs) {
    float t = 0;
    cudaEventSynchronize(e.b);
    cudaEventElapsedTime(&t, e.a, e.b);
    *ms[e.cat] += t; (*cnt[e.cat])++;
    cudaEventDestroy(e.a); cudaEventDestroy(e.b);
  }
  h->evs.clear();
}

int mcp_ba_compute(McpBa* h, volatile const uint8_t* abort_flag, int32_t n_iter, double user_lambda, McpBaStats* stats)
{
  if (!h || !h->loaded) { set_last_error("mcp_ba_compute: no problem loaded"); return MCP_ERR_STATE; }
  cudaSetDevice(h->device);
  McpBaStats local;
  if (!stats) stats = &local;
  memset(stats, 0, sizeof(*stats));
  h->launches = 0;
  const int rc = run_compute(h, abort_flag, n_iter, user_lambda, stats, false, 0, 0);
  if (h->profiling) resolve_profile(h);
  if (rc < -1) cudaStreamSynchronize(h->stream);
  return rc;
}

int mcp_ba_lm_step(McpBa* h, double lambda, double sigma_sq, double* delta, double* sigma_sq_used, double* robust_chi2)
{
  if (!h || !h->loaded) { set_last_error("mcp_ba_lm_step: no problem loaded"); return MCP_ERR_STATE; }
  if (h->world > 1) { set_last_error("mcp_ba_lm_step: single-GPU diagnostic only"); return MCP_ERR_UNSUPPORTED; }
  cudaSetDevice(h->device);
  McpBaStats st;
  memset(&st, 0, sizeof(st));
  const int cur_before = h->ctrl_host->cur;
  int rc = run_compute(h, nullptr, 1, -1.0, &st, true, lambda, sigma_sq);
  if (rc) return rc;
  BaCtrl& c = *h->ctrl_host;
  const size_t dim = (size_t)c.dim;
  if ((rc = h->b_tmp.ensure(sizeof(double) * (dim + 1)))) return rc;
  // the update belongs to the pre-step lambda: restore it for the gather
  const double lam_after = c.lambda;
  c.lambda = lambda; c.cur = cur_before;
  if ((rc = push_ctrl(h))) return rc;
  launch_gather_delta(h->d, h->b_tmp.as<double>(), h->stream);
  if (delta) MCP_CUDA_CHECK(cudaMemcpyAsync(delta, h->b_tmp.p, sizeof(double) * dim, cudaMemcpyDeviceToHost, h->stream));
  MCP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  (void)lam_after;
  if (sigma_sq_used) *sigma_sq_used = c.sigma_sq_raw;
  if (robust_chi2) *robust_chi2 = c.lin_chi;
  // leave the estimate untouched (cur restored) and reset the LM history
  c.last_chi2 = 1.7976931348623157e308; c.qmax = 0; c.iter = 0;
  if ((rc = push_ctrl(h))) return rc;
  MCP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return MCP_OK;
}

int mcp_ba_get_poses(McpBa* h, double* out)
{
  if (!h || !h->loaded || !out) { set_last_error("mcp_ba_get_poses: bad arguments"); return MCP_ERR_INVALID; }
  cudaSetDevice(h->device);
  MCP_CUDA_CHECK(cudaMemcpyAsync(out, h->d.pose[h->ctrl_host->cur], sizeof(double) * 12 * (size_t)h->d.n_pose, cudaMemcpyDeviceToHost, h->stream));
  MCP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return MCP_OK;
}
int mcp_ba_get_points(McpBa* h, double* out)
{
  if (!h || !h->loaded || !out) { set_last_error("mcp_ba_get_points: bad arguments"); return MCP_ERR_INVALID; }
  cudaSetDevice(h->device);
  MCP_CUDA_CHECK(cudaMemcpyAsync(out, h->d.pt[h->ctrl_host->cur], sizeof(double) * 3 * (size_t)h->d.n_pt, cudaMemcpyDeviceToHost, h->stream));
  MCP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return MCP_OK;
}
int mcp_ba_get_outliers(McpBa* h, int32_t* idx, int32_t cap)
{
  if (!h) { set_last_error("mcp_ba_get_outliers: NULL handle"); return MCP_ERR_INVALID; }
  const int n = (int)h->outliers.size();
  if (idx && cap > 0) memcpy(idx, h->outliers.data(), sizeof(int32_t) * (size_t)std::min(n, (int)cap));
  return n;
}
int mcp_ba_set_state(McpBa* h, const double* pose_Rt, const double* pt_xyz)
{
  if (!h || !h->loaded) { set_last_error("mcp_ba_set_state: no problem loaded"); return MCP_ERR_STATE; }
  cudaSetDevice(h->device);
  const int cur = h->ctrl_host->cur;
  for (int k = 0; k < N_STATE; k++) {
    if (pose_Rt) MCP_CUDA_CHECK(cudaMemcpyAsync(h->d.pose[k], pose_Rt, sizeof(double) * 12 * (size_t)h->d.n_pose, cudaMemcpyHostToDevice, h->stream));
    if (pt_xyz && (k == cur)) MCP_CUDA_CHECK(cudaMemcpyAsync(h->d.pt[k], pt_xyz, sizeof(double) * 3 * (size_t)h->d.n_pt, cudaMemcpyHostToDevice, h->stream));
  }
  MCP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return MCP_OK;
}
int mcp_ba_reset_state(McpBa* h)
{
  if (!h || !h->loaded) { set_last_error("mcp_ba_reset_state: no problem loaded"); return MCP_ERR_STATE; }
  cudaSetDevice(h->device);
  for (int k = 0; k < N_STATE; k++) {
    MCP_CUDA_CHECK(cudaMemcpyAsync(h->d.pose[k], h->b_pose0.p, sizeof(double) * 12 * (size_t)h->d.n_pose, cudaMemcpyDeviceToDevice, h->stream));
    MCP_CUDA_CHECK(cudaMemcpyAsync(h->d.pt[k], h->b_pt0.p, sizeof(double) * 3 * (size_t)h->d.n_pt, cudaMemcpyDeviceToDevice, h->stream));
  }
  BaCtrl& c = *h->ctrl_host;
  int rc;
  if ((rc = sync_ctrl(h))) return rc;
  c.last_chi2 = 1.7976931348623157e308; c.cur = 0;
  if ((rc = push_ctrl(h))) return rc;
  MCP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return MCP_OK;
}

int mcp_ba_eval(McpBa* h, double* err_xy, double* chi2)
{
  if (!h || !h->loaded) { set_last_error("mcp_ba_eval: no problem loaded"); return MCP_ERR_STATE; }
  cudaSetDevice(h->device);
  int rc;
  const size_t n = (size_t)h->d.n_meas;
  if ((rc = h->b_tmp.ensure(sizeof(double) * (2 * n + 2)))) return rc;
  if ((rc = sync_ctrl(h))) return rc;
  launch_backsub_eval(h->d, 0, -1, h->b_tmp.as<double>(), h->stream);
  if ((rc = allgather_ranges(h, h->d.chi2[h->ctrl_host->cur], h->part_meas, 1))) return rc;
  if (err_xy) MCP_CUDA_CHECK(cudaMemcpyAsync(err_xy, h->b_tmp.p, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, h->stream));
  std::vector<double> sorted(n);
  MCP_CUDA_CHECK(cudaMemcpyAsync(sorted.data(), h->d.chi2[h->ctrl_host->cur], sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  MCP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  if (chi2) for (size_t q = 0; q < n; q++) chi2[h->meas_orig[q]] = sorted[q];
  return MCP_OK;
}

int mcp_ba_debug_jacobians(McpBa* h, double* J30)
{
  if (!h || !h->loaded || !J30) { set_last_error("mcp_ba_debug_jacobians: bad arguments"); return MCP_ERR_INVALID; }
  cudaSetDevice(h->device);
  int rc;
  const size_t n = (size_t)h->d.n_meas;
  if ((rc = h->b_tmp.ensure(sizeof(double) * (30 * n + 2)))) return rc;
  if ((rc = sync_ctrl(h))) return rc;
  launch_debug_jacobians(h->d, h->b_tmp.as<double>(), h->stream);
  MCP_CUDA_CHECK(cudaMemcpyAsync(J30, h->b_tmp.p, sizeof(double) * 30 * n, cudaMemcpyDeviceToHost, h->stream));
  MCP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return MCP_OK;
}

/* Debug: per-task timestamps of the next k_chol_solve launches (8 doubles per task: i, j, t_start, t_deps, t_end, cta) */
int mcp_ba_debug_solve_trace(McpBa* h, double* out, int32_t cap_doubles)
{
  if (!h || !h->loaded) return MCP_ERR_STATE;
  cudaSetDevice(h->device);
  const size_t nd = 8 * 2048;
  if (!out) {   // arm
    int rc = h->b_dbg.ensure(sizeof(double) * nd);
    if (rc) return rc;
    MCP_CUDA_CHECK(cudaMemsetAsync(h->b_dbg.p, 0, sizeof(double) * nd, h->stream));
    h->d.dbg = h->b_dbg.as<double>();
    return MCP_OK;
  }
  if (!h->d.dbg) return MCP_ERR_STATE;
  const size_t n = std::min((size_t)cap_doubles, nd);
  MCP_CUDA_CHECK(cudaMemcpyAsync(out, h->b_dbg.p, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  MCP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  h->d.dbg = nullptr;
  return MCP_OK;
}
int mcp_ba_get_stream(McpBa* h, void** out) { if (!h || !out) return MCP_ERR_INVALID; *out = (void*)h->stream; return MCP_OK; }
int mcp_ba_set_profiling(McpBa* h, int32_t enable) { if (!h) return MCP_ERR_INVALID; h->profiling = enable != 0; return MCP_OK; }
int mcp_ba_get_timing(McpBa* h, McpBaTiming* out) { if (!h || !out) return MCP_ERR_INVALID; *out = h->timing; return MCP_OK; }

int mcp_nccl_unique_id(void* out128)
{
  if (!out128) return MCP_ERR_INVALID;
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  NCCL_CHECK(ncclGetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  return MCP_OK;
}

int mcp_ba_partition(int32_t n_pt, const int32_t* meas_pt, int32_t n_meas, int32_t world, int32_t* part_pt)
{
  if (n_pt < 0 || n_meas < 0 || world < 1 || !part_pt || (n_meas && !meas_pt)) { set_last_error("mcp_ba_partition: bad arguments"); return MCP_ERR_INVALID; }
  std::vector<int> off(n_pt + 1, 0);
  for (int m = 0; m < n_meas; m++) {
    if (meas_pt[m] < 0 || meas_pt[m] >= n_pt) { set_last_error("mcp_ba_partition: point index out of range"); return MCP_ERR_INVALID; }
    off[meas_pt[m] + 1]++;
  }
  for (int p = 0; p < n_pt; p++) off[p + 1] += off[p];
  partition_points(off.data(), n_pt, world, part_pt);
  return MCP_OK;
}

int mcp_ba_comm_init(McpBa* h, const void* nccl_unique_id, int32_t rank, int32_t world)
{
  if (!h || world < 1 || rank < 0 || rank >= world) { set_last_error("mcp_ba_comm_init: bad arguments"); return MCP_ERR_INVALID; }
  if (h->loaded) { set_last_error("mcp_ba_comm_init: must be called before mcp_ba_load"); return MCP_ERR_STATE; }
  cudaSetDevice(h->device);
  if (h->comm) { ncclCommDestroy(h->comm); h->comm = nullptr; }
  h->rank = rank; h->world = world;
  if (world == 1) return MCP_OK;
  if (!nccl_unique_id) { set_last_error("mcp_ba_comm_init: unique id required for world > 1"); return MCP_ERR_INVALID; }
  g_pdl_multi_gpu.store(true, std::memory_order_relaxed);
  ncclUniqueId id;
  memcpy(&id, nccl_unique_id, sizeof(id));
  NCCL_CHECK(ncclCommInitRank(&h->comm, world, id, rank));
  return MCP_OK;
}

}  // extern "C"
