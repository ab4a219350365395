// fe_api.cu — C ABI of the front end (include/mcptam_b200.h): resident keyframe pyramids on the device,
// KeyFrame::MakeKeyFrame_Lite (src/KeyFrame.cc:145-361) and the batched Tracker::SearchForPoints body
// (src/Tracker.cc:1299-1377) as kernels; host<->device copies through pinned staging buffers.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda.h>

#include "fe_types.cuh"

namespace mcp {
void fe_launch_pyramid(const FeKf& kf, int rnd, cudaStream_t s);
int fe_launch_fast(const FeKf& kf, int adaptive, cudaStream_t s);
bool fe_fast_fused();
bool fe_zero_copy();
void fe_launch_glare_mask(const FeKf& kf, const FeMasks& out, cudaStream_t s);
void fe_launch_patch_search(const FeDev& fe, const void* tmaps, bool all_fit_window, int target, int n, const McpPatchReq* req, McpPatchRes* res, uint8_t* templ, cudaStream_t s);
void fe_launch_shitomasi(const FeLevel& L, int n, const int2* xy, double* out, cudaStream_t s);
void fe_launch_calc_jacobians(const DevCam& cam, const Se3& B, const Se3& Cb, int n, const double* pw, McpJacRes* out, cudaStream_t s);
void fe_launch_pose_update(int n, const McpPoseMeas* meas, int estimator, double override_sigma, double* e2buf, McpPoseUpdate* res, int* outlier, cudaStream_t s);
void fe_launch_project(const DevCam& cam, const Se3& T, int n, const double* pw, const double* rw, const double* dw, McpProjRes* out, cudaStream_t s);
void fe_launch_minipatch(const FeLevel& S, const FeLevel& T, int n_corners, int n, const int2* src, const int2* start, int range,
                         int2* pos, int* found, cudaStream_t s);
void fe_launch_rest_level(const FeKf& kf, const FeKf* prev, int level, int n_hint, int n_prev_corners, int adaptive, int strict, int use_shi,
                          int use_thresh, double top_fraction, double thresh, int n_prev, int* flag, double* score, McpCandidate* cand,
                          McpCandidate* cand_out, int2* cand_xy, int2* pos1, int2* pos2, int* f1, int* f2, int* ctr, cudaStream_t s);
}  // namespace mcp

using namespace mcp;

struct McpFe {
  McpFeConfig cfg;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
  std::vector<FeKf> kf_host;       // per slot (device pointers inside)
  std::vector<bool> kf_valid;
  FeKf* kf_dev = nullptr;
  uint8_t* pool = nullptr;         // one device allocation for everything
  size_t pool_bytes = 0;
  size_t out_block_bytes = 0;      // per slot: meta + luts + corners (contiguous, copied to the host in one go)
  std::vector<uint8_t*> out_block_dev;
  uint8_t* mask_dev[MCP_LEVELS] = { nullptr, nullptr, nullptr, nullptr };
  std::vector<CUtensorMap> tmaps;  // one per (slot, level): 2-D tile descriptors of the level images (TMA window loads); empty: not available
  uint8_t* gmask_dev[MCP_LEVELS] = { nullptr, nullptr, nullptr, nullptr };   // internal mask AND glare mask of the last frame (Level::lastMask)
  bool has_mask = false, glare = false;
  uint8_t* stage_img = nullptr;    // pinned
  uint8_t* stage_out = nullptr;    // pinned
  McpPatchReq* req_host = nullptr; // pinned
  McpPatchRes* res_host = nullptr; // pinned
  McpPatchReq* req_dev = nullptr;
  McpPatchRes* res_dev = nullptr;
  uint8_t* templ_dev = nullptr;
  int2* xy_dev = nullptr; int2* xy2_dev = nullptr; int2* pos_dev = nullptr; int* found_dev = nullptr; double* sc_dev = nullptr;
  void* aux_host = nullptr;        // pinned scratch for shitomasi / minipatch
  int last_n = 0;
  DevCam cam; bool has_cam = false;
  double* proj_in = nullptr; McpProjRes* proj_out = nullptr; size_t proj_cap = 0;
  void* pu_buf = nullptr; size_t pu_cap = 0;
  uint8_t* rest_buf = nullptr; void* rest_host = nullptr;   // MakeKeyFrame_Rest scratch (lazy)
  McpFeTiming timing;
  int lw[MCP_LEVELS], lh[MCP_LEVELS], lp[MCP_LEVELS];
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" {

void mcp_fe_default_config(McpFeConfig* c)
{
  memset(c, 0, sizeof(*c));
  c->width = 640; c->height = 480;
  c->adaptive_thresh = 1;             // KeyFrame::sbAdaptiveThresh, src/KeyFrame.cc:70
  c->max_corners_per_level = 8192;
  c->max_keyframes = 8;
  c->max_patches = 4096;
  c->device = -1;
}

int mcp_fe_create(const McpFeConfig* cfg, McpFe** out)
{
  if (!out) { set_last_error("mcp_fe_create: out is NULL"); return MCP_ERR_INVALID; }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_last_error("mcp_fe_create: no CUDA device available (this library has no CPU fallback)");
    return MCP_ERR_NO_DEVICE;
  }
  McpFe* h = new McpFe();
  if (cfg) h->cfg = *cfg; else mcp_fe_default_config(&h->cfg);
  const McpFeConfig& c = h->cfg;
  if (c.width < 64 || c.height < 64 || c.max_keyframes < 1 || c.max_corners_per_level < 16 || c.max_patches < 1) {
    set_last_error("mcp_fe_create: bad configuration");
    delete h;
    return MCP_ERR_INVALID;
  }
  if (c.device >= 0) MCP_CUDA_CHECK(cudaSetDevice(c.device));
  MCP_CUDA_CHECK(cudaGetDevice(&h->device));
  MCP_CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  for (auto& e : h->ev) MCP_CUDA_CHECK(cudaEventCreate(&e));
  int w = c.width, hh = c.height;
  for (int l = 0; l < MCP_LEVELS; l++) { h->lw[l] = w; h->lh[l] = hh; h->lp[l] = (int)align_up((size_t)w + 8, 128); w /= 2; hh /= 2; }
  // ---- carve one pool ---------------------------------------------------------------------------------
  const int S = c.max_keyframes, cap = c.max_corners_per_level;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 256); return o; };
  std::vector<size_t> o_img(S * MCP_LEVELS), o_score(S * MCP_LEVELS), o_rowcount(S * MCP_LEVELS), o_hist(S), o_out(S);
  size_t lut_total = 0;
  for (int l = 0; l < MCP_LEVELS; l++) lut_total += align_up(sizeof(int) * h->lh[l], 256);
  h->out_block_bytes = align_up(sizeof(FeMeta), 256) + lut_total + (size_t)MCP_LEVELS * align_up(sizeof(int2) * cap, 256);
  for (int s = 0; s < S; s++) {
    for (int l = 0; l < MCP_LEVELS; l++) {
      o_img[s * MCP_LEVELS + l] = take((size_t)h->lp[l] * (h->lh[l] + 2) + 64);
      o_score[s * MCP_LEVELS + l] = take((size_t)h->lp[l] * h->lh[l]);
      o_rowcount[s * MCP_LEVELS + l] = take(sizeof(int) * (32 * (size_t)h->lh[l] + 32));
    }
    o_hist[s] = take(sizeof(unsigned) * 32 * MCP_LEVELS);
    o_out[s] = take(h->out_block_bytes);
  }
  size_t o_mask[MCP_LEVELS];
  for (int l = 0; l < MCP_LEVELS; l++) o_mask[l] = take((size_t)h->lp[l] * h->lh[l]);
  size_t o_gmask[MCP_LEVELS];
  for (int l = 0; l < MCP_LEVELS; l++) o_gmask[l] = take((size_t)h->lp[l] * h->lh[l]);
  const size_t o_kf = take(sizeof(FeKf) * S);
  const size_t o_req = take(sizeof(McpPatchReq) * c.max_patches);
  const size_t o_res = take(sizeof(McpPatchRes) * c.max_patches);
  const size_t o_templ = take((size_t)64 * c.max_patches);
  const size_t o_xy = take(sizeof(int2) * c.max_patches), o_xy2 = take(sizeof(int2) * c.max_patches);
  const size_t o_pos = take(sizeof(int2) * c.max_patches), o_found = take(sizeof(int) * c.max_patches), o_sc = take(sizeof(double) * c.max_patches);
  h->pool_bytes = off;
  MCP_CUDA_CHECK(cudaMalloc(&h->pool, h->pool_bytes));
  MCP_CUDA_CHECK(cudaMemsetAsync(h->pool, 0, h->pool_bytes, h->stream));
  h->kf_host.resize(S);
  h->kf_valid.assign(S, false);
  h->out_block_dev.resize(S);
  for (int l = 0; l < MCP_LEVELS; l++) { h->mask_dev[l] = h->pool + o_mask[l]; h->gmask_dev[l] = h->pool + o_gmask[l]; }
  for (int s = 0; s < S; s++) {
    FeKf& k = h->kf_host[s];
    memset(&k, 0, sizeof(k));
    uint8_t* ob = h->pool + o_out[s];
    h->out_block_dev[s] = ob;
    k.meta = reinterpret_cast<FeMeta*>(ob);
    size_t oo = align_up(sizeof(FeMeta), 256);
    int tiles = 0, rows = 0;
    for (int l = 0; l < MCP_LEVELS; l++) {
      FeLevel& L = k.lv[l];
      L.w = h->lw[l]; L.h = h->lh[l]; L.pitch = h->lp[l];
      L.img = h->pool + o_img[s * MCP_LEVELS + l];
      L.score = h->pool + o_score[s * MCP_LEVELS + l];
      L.rowcount = reinterpret_cast<int*>(h->pool + o_rowcount[s * MCP_LEVELS + l]);
      L.rowhist = L.rowcount;
      L.hist = reinterpret_cast<unsigned*>(h->pool + o_hist[s]) + 32 * l;
      L.mask = nullptr;
      L.row_lut = reinterpret_cast<int*>(ob + oo);
      oo += align_up(sizeof(int) * L.h, 256);
      k.tile_off[l] = tiles; k.row_off[l] = rows;
      tiles += ((L.w + 31) / 32) * ((L.h + 7) / 8);
      rows += L.h;
    }
    for (int l = 0; l < MCP_LEVELS; l++) { k.lv[l].corners = reinterpret_cast<int2*>(ob + oo); oo += align_up(sizeof(int2) * cap, 256); }
    k.tile_off[MCP_LEVELS] = tiles; k.row_off[MCP_LEVELS] = rows;
    // old-style fixed thresholds (src/KeyFrame.cc:322-341)
    k.fixed_thresh[0] = 10; k.fixed_thresh[1] = 15; k.fixed_thresh[2] = 15; k.fixed_thresh[3] = 10;
    k.corner_cap = cap;
  }
  h->kf_dev = reinterpret_cast<FeKf*>(h->pool + o_kf);
  MCP_CUDA_CHECK(cudaMemcpyAsync(h->kf_dev, h->kf_host.data(), sizeof(FeKf) * S, cudaMemcpyHostToDevice, h->stream));
  {
    // 2-D TMA descriptors (80 x 48 byte boxes) of every resident level image: the patch search fetches a patch's whole
    // footprint with one tile copy.  The encoder is a driver entry point (no libcuda link); MCP_FE_TMA=0 or any failure keeps
    // the global-memory kernel.
    static const bool off = [] { const char* e = getenv("MCP_FE_TMA"); return e && e[0] == '0'; }();
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (!off && cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn && qres == cudaDriverEntryPointSuccess) {
      std::vector<CUtensorMap> maps((size_t)S * MCP_LEVELS);
      bool ok = true;
      for (int s = 0; s < S && ok; s++)
        for (int l = 0; l < MCP_LEVELS && ok; l++) {
          const FeLevel& L = h->kf_host[s].lv[l];
          const cuuint64_t dims[2] = { (cuuint64_t)L.w, (cuuint64_t)L.h }, strides[1] = { (cuuint64_t)L.pitch };
          const cuuint32_t box[2] = { 80, 48 }, estr[2] = { 1, 1 };
          ok = reinterpret_cast<EncodeFn>(fn)(&maps[(size_t)s * MCP_LEVELS + l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, L.img, dims, strides, box, estr,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
      if (ok) h->tmaps.swap(maps);
    } else (void)cudaGetLastError();
  }
  h->req_dev = reinterpret_cast<McpPatchReq*>(h->pool + o_req);
  h->res_dev = reinterpret_cast<McpPatchRes*>(h->pool + o_res);
  h->templ_dev = h->pool + o_templ;
  h->xy_dev = reinterpret_cast<int2*>(h->pool + o_xy); h->xy2_dev = reinterpret_cast<int2*>(h->pool + o_xy2);
  h->pos_dev = reinterpret_cast<int2*>(h->pool + o_pos); h->found_dev = reinterpret_cast<int*>(h->pool + o_found);
  h->sc_dev = reinterpret_cast<double*>(h->pool + o_sc);
  MCP_CUDA_CHECK(cudaMallocHost(&h->stage_img, (size_t)c.width * c.height));
  MCP_CUDA_CHECK(cudaMallocHost(&h->stage_out, h->out_block_bytes));
  MCP_CUDA_CHECK(cudaMallocHost(&h->req_host, sizeof(McpPatchReq) * c.max_patches));
  MCP_CUDA_CHECK(cudaMallocHost(&h->res_host, sizeof(McpPatchRes) * c.max_patches));
  MCP_CUDA_CHECK(cudaMallocHost(&h->aux_host, (size_t)64 * c.max_patches));
  MCP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  memset(&h->timing, 0, sizeof(h->timing));
  *out = h;
  return MCP_OK;
}

int mcp_fe_destroy(McpFe* h)
{
  if (!h) return MCP_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->pool) cudaFree(h->pool);
  if (h->stage_img) cudaFreeHost(h->stage_img);
  if (h->stage_out) cudaFreeHost(h->stage_out);
  if (h->req_host) cudaFreeHost(h->req_host);
  if (h->res_host) cudaFreeHost(h->res_host);
  if (h->aux_host) cudaFreeHost(h->aux_host);
  if (h->proj_in) cudaFree(h->proj_in);
  if (h->proj_out) cudaFree(h->proj_out);
  if (h->pu_buf) cudaFree(h->pu_buf);
  if (h->rest_buf) cudaFree(h->rest_buf);
  if (h->rest_host) cudaFreeHost(h->rest_host);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return MCP_OK;
}

// KeyFrame::SetMask (src/KeyFrame.cc:116-126): level-0 mask, half-sampled down the pyramid
int mcp_fe_set_mask(McpFe* h, const uint8_t* mask, int32_t stride)
{
  if (!h) { set_last_error("mcp_fe_set_mask: NULL handle"); return MCP_ERR_INVALID; }
  cudaSetDevice(h->device);
  h->has_mask = mask != nullptr;
  if (mask) {
    for (int y = 0; y < h->lh[0]; y++) memcpy(h->stage_img + (size_t)y * h->lw[0], mask + (size_t)y * stride, h->lw[0]);
    MCP_CUDA_CHECK(cudaMemcpy2DAsync(h->mask_dev[0], h->lp[0], h->stage_img, h->lw[0], h->lw[0], h->lh[0], cudaMemcpyHostToDevice, h->stream));
    FeKf tmp = h->kf_host[0];
    for (int l = 0; l < MCP_LEVELS; l++) tmp.lv[l].img = h->mask_dev[l];
    fe_launch_pyramid(tmp, h->cfg.halfsample_round, h->stream);
  }
  for (auto& k : h->kf_host)
    for (int l = 0; l < MCP_LEVELS; l++) k.lv[l].mask = mask ? h->mask_dev[l] : nullptr;
  MCP_CUDA_CHECK(cudaMemcpyAsync(h->kf_dev, h->kf_host.data(), sizeof(FeKf) * h->kf_host.size(), cudaMemcpyHostToDevice, h->stream));
  MCP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return MCP_OK;
}

// bGlareMasking of KeyFrame::MakeKeyFrame_Lite (src/KeyFrame.cc:214-242): applies to the following mcp_fe_make_keyframe calls
int mcp_fe_set_glare_masking(McpFe* h, int32_t enable)
{
  if (!h) { set_last_error("mcp_fe_set_glare_masking: NULL handle"); return MCP_ERR_INVALID; }
  h->glare = enable != 0;
  return MCP_OK;
}

int mcp_fe_make_keyframe(McpFe* h, int32_t slot, const uint8_t* img, int32_t stride, McpLevelOut out[MCP_LEVELS])
{
  if (!h || !img || slot < 0 || slot >= (int)h->kf_host.size() || stride < h->lw[0]) {
    set_last_error("mcp_fe_make_keyframe: bad arguments");
    return MCP_ERR_INVALID;
  }
  cudaSetDevice(h->device);
  cudaStream_t s = h->stream;
  const FeKf& kf = h->kf_host[slot];
  const int w = h->lw[0], hh = h->lh[0];
  MCP_CUDA_CHECK(cudaEventRecord(h->ev[0], s));
  {
    // the image goes through the pinned staging buffer in four bands: the upload of a band overlaps the host copy of the next
    const int bands = hh >= 64 ? 4 : 1;
    for (int b = 0; b < bands; b++) {
      const int y0 = (int)((long long)hh * b / bands), y1 = (int)((long long)hh * (b + 1) / bands);
      if (stride == w) memcpy(h->stage_img + (size_t)y0 * w, img + (size_t)y0 * w, (size_t)w * (y1 - y0));
      else for (int y = y0; y < y1; y++) memcpy(h->stage_img + (size_t)y * w, img + (size_t)y * stride, w);
      MCP_CUDA_CHECK(cudaMemcpy2DAsync(kf.lv[0].img + (size_t)y0 * kf.lv[0].pitch, kf.lv[0].pitch, h->stage_img + (size_t)y0 * w, w, w, y1 - y0, cudaMemcpyHostToDevice, s));
    }
  }
  MCP_CUDA_CHECK(cudaEventRecord(h->ev[1], s));
  fe_launch_pyramid(kf, h->cfg.halfsample_round, s);
  MCP_CUDA_CHECK(cudaEventRecord(h->ev[2], s));
  int nl = 0;
  // two-launch FAST: the kernels write meta / row LUT / corners to the device block AND to its pinned host mirror
  const bool mirrored = fe_fast_fused() && fe_zero_copy();
  FeKf kl = kf;
  if (mirrored) kl.host_delta = (long long)(reinterpret_cast<char*>(h->stage_out) - reinterpret_cast<char*>(h->out_block_dev[slot]));
  if (h->glare) {
    // the frame's effective mask = internal mask AND (no pixel > 245 within the dilation footprint); FAST filters against it
    FeMasks gm;
    for (int l = 0; l < MCP_LEVELS; l++) { gm.m[l] = h->gmask_dev[l]; kl.lv[l].mask = h->gmask_dev[l]; }
    fe_launch_glare_mask(kf, gm, s);
    nl = 1 + fe_launch_fast(kl, h->cfg.adaptive_thresh, s);
  } else nl = fe_launch_fast(kl, h->cfg.adaptive_thresh, s);
  MCP_CUDA_CHECK(cudaEventRecord(h->ev[3], s));
  if (!mirrored) MCP_CUDA_CHECK(cudaMemcpyAsync(h->stage_out, h->out_block_dev[slot], h->out_block_bytes, cudaMemcpyDeviceToHost, s));
  MCP_CUDA_CHECK(cudaEventRecord(h->ev[4], s));
  MCP_CUDA_CHECK(cudaStreamSynchronize(s));
  h->kf_valid[slot] = true;
  float t;
  cudaEventElapsedTime(&t, h->ev[1], h->ev[2]); h->timing.ms_pyramid = t;
  cudaEventElapsedTime(&t, h->ev[2], h->ev[3]); h->timing.ms_fast = t;
  cudaEventElapsedTime(&t, h->ev[0], h->ev[1]); h->timing.ms_other = t;
  cudaEventElapsedTime(&t, h->ev[3], h->ev[4]); h->timing.ms_compact = t;    // device->host copy of the results
  h->timing.n_launches = 1 + nl;
  if (out) {
    const FeMeta* meta = reinterpret_cast<const FeMeta*>(h->stage_out);
    size_t oo = align_up(sizeof(FeMeta), 256);
    const int cap = h->cfg.max_corners_per_level;
    size_t lut_off[MCP_LEVELS], cor_off[MCP_LEVELS];
    for (int l = 0; l < MCP_LEVELS; l++) { lut_off[l] = oo; oo += align_up(sizeof(int) * h->lh[l], 256); }
    for (int l = 0; l < MCP_LEVELS; l++) { cor_off[l] = oo; oo += align_up(sizeof(int2) * cap, 256); }
    for (int l = 0; l < MCP_LEVELS; l++) {
      McpLevelOut& o = out[l];
      o.width = h->lw[l]; o.height = h->lh[l];
      o.n_corners = std::min(meta->lv[l].n_corners, cap);
      o.n_corners_total = meta->lv[l].n_corners;            // > n_corners: the level overflowed max_corners_per_level (the tail is lost)
      if (o.corners_cap < 0) o.corners_cap = 0;
      o.fast_thresh = meta->lv[l].fast_thresh;
      memcpy(o.fast_freq, meta->lv[l].fast_freq, sizeof(o.fast_freq));
      if (o.corners_xy) memcpy(o.corners_xy, h->stage_out + cor_off[l], sizeof(int32_t) * 2 * (size_t)std::min(o.n_corners, o.corners_cap));
      if (o.row_lut) memcpy(o.row_lut, h->stage_out + lut_off[l], sizeof(int32_t) * h->lh[l]);
      if (o.last_mask) {
        // Level::lastMask: the mask the corners were filtered with (all 255 when neither an internal mask nor glare masking is on)
        const uint8_t* src = h->glare ? h->gmask_dev[l] : (h->has_mask ? h->mask_dev[l] : nullptr);
        if (src) MCP_CUDA_CHECK(cudaMemcpy2D(o.last_mask, h->lw[l], src, h->lp[l], h->lw[l], h->lh[l], cudaMemcpyDeviceToHost));
        else memset(o.last_mask, 255, (size_t)h->lw[l] * h->lh[l]);
      }
      if (o.image) {
        if (l == 0) memcpy(o.image, h->stage_img, (size_t)w * hh);
        else MCP_CUDA_CHECK(cudaMemcpy2D(o.image, h->lw[l], kf.lv[l].img, kf.lv[l].pitch, h->lw[l], h->lh[l], cudaMemcpyDeviceToHost));
      }
    }
  }
  return MCP_OK;
}

int mcp_fe_search_patches(McpFe* h, int32_t target_kf, int32_t n, const McpPatchReq* req, McpPatchRes* res)
{
  if (!h || n < 0 || (n && (!req || !res)) || target_kf < 0 || target_kf >= (int)h->kf_host.size()) {
    set_last_error("mcp_fe_search_patches: bad arguments");
    return MCP_ERR_INVALID;
  }
  if (n > h->cfg.max_patches) { set_last_error("mcp_fe_search_patches: n=%d exceeds max_patches=%d", n, h->cfg.max_patches); return MCP_ERR_INVALID; }
  if (!h->kf_valid[target_kf]) { set_last_error("mcp_fe_search_patches: target keyframe slot %d is empty", target_kf); return MCP_ERR_STATE; }
  for (int i = 0; i < n; i++) {
    const int sk = req[i].src_kf;
    if (sk < 0 || sk >= (int)h->kf_host.size() || !h->kf_valid[sk]) { set_last_error("mcp_fe_search_patches: request %d: source keyframe slot %d is out of range or empty", i, sk); return MCP_ERR_STATE; }
  }
  if (n == 0) return MCP_OK;
  cudaSetDevice(h->device);
  cudaStream_t s = h->stream;
  memcpy(h->req_host, req, sizeof(McpPatchReq) * (size_t)n);
  MCP_CUDA_CHECK(cudaEventRecord(h->ev[0], s));
  // the kernel reads the requests from the pinned host array (one uniform read per warp) and writes its results there:
  // no copy-engine hop on either side of the launch (MCP_FE_ZEROCOPY=0: staged copies)
  const bool req_direct = fe_zero_copy();
  if (!req_direct) MCP_CUDA_CHECK(cudaMemcpyAsync(h->req_dev, h->req_host, sizeof(McpPatchReq) * (size_t)n, cudaMemcpyHostToDevice, s));
  MCP_CUDA_CHECK(cudaEventRecord(h->ev[1], s));
  FeDev fe;
  fe.kf = h->kf_dev; fe.n_slots = (int)h->kf_host.size(); fe.transform_round = h->cfg.transform_round;
  // the TMA kernel stages an 80 x 48 window around the prediction: disc radius + half patch + sub-pixel drift must fit 24 rows
  bool fits = true;
  for (int i = 0; i < n && fits; i++) {
    const int lv = req[i].search_level;
    if (lv < 0 || lv >= MCP_LEVELS) continue;
    const int nr = (req[i].range + (1 << lv) - 1) >> lv;
    fits = req[i].exhaustive == 2 || (req[i].range >= 0 && nr + 4 + 3 <= 24);
  }
  // the kernel writes its results straight into the pinned (mapped) host array: no device->host copy behind it
  const bool direct = fe_zero_copy();
  fe_launch_patch_search(fe, h->tmaps.empty() ? nullptr : h->tmaps.data(), fits, target_kf, n, req_direct ? h->req_host : h->req_dev, direct ? h->res_host : h->res_dev, h->templ_dev, s);
  MCP_CUDA_CHECK(cudaEventRecord(h->ev[2], s));
  if (!direct) MCP_CUDA_CHECK(cudaMemcpyAsync(h->res_host, h->res_dev, sizeof(McpPatchRes) * (size_t)n, cudaMemcpyDeviceToHost, s));
  MCP_CUDA_CHECK(cudaStreamSynchronize(s));
  memcpy(res, h->res_host, sizeof(McpPatchRes) * (size_t)n);
  float t;
  cudaEventElapsedTime(&t, h->ev[1], h->ev[2]); h->timing.ms_search = t;
  h->timing.n_launches = 1;
  h->last_n = n;
  return MCP_OK;
}

int mcp_fe_get_templates(McpFe* h, int32_t n, uint8_t* templ)
{
  if (!h || !templ || n < 0 || n > h->last_n) { set_last_error("mcp_fe_get_templates: bad arguments"); return MCP_ERR_INVALID; }
  cudaSetDevice(h->device);
  MCP_CUDA_CHECK(cudaMemcpyAsync(templ, h->templ_dev, (size_t)64 * n, cudaMemcpyDeviceToHost, h->stream));
  MCP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return MCP_OK;
}

int mcp_fe_shitomasi(McpFe* h, int32_t kf, int32_t level, int32_t n, const int32_t* xy, double* scores)
{
  if (!h || kf < 0 || kf >= (int)h->kf_host.size() || level < 0 || level >= MCP_LEVELS || n < 0 || n > h->cfg.max_patches || (n && (!xy || !scores))) {
    set_last_error("mcp_fe_shitomasi: bad arguments");
    return MCP_ERR_INVALID;
  }
  if (!h->kf_valid[kf]) { set_last_error("mcp_fe_shitomasi: keyframe slot %d is empty", kf); return MCP_ERR_STATE; }
  if (n == 0) return MCP_OK;
  cudaSetDevice(h->device);
  cudaStream_t s = h->stream;
  memcpy(h->aux_host, xy, sizeof(int32_t) * 2 * (size_t)n);
  MCP_CUDA_CHECK(cudaMemcpyAsync(h->xy_dev, h->aux_host, sizeof(int2) * (size_t)n, cudaMemcpyHostToDevice, s));
  fe_launch_shitomasi(h->kf_host[kf].lv[level], n, h->xy_dev, h->sc_dev, s);
  MCP_CUDA_CHECK(cudaMemcpyAsync(scores, h->sc_dev, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
  MCP_CUDA_CHECK(cudaStreamSynchronize(s));
  return MCP_OK;
}

int mcp_fe_minipatch_find(McpFe* h, int32_t kf_src, int32_t kf_dst, int32_t level, int32_t n, const int32_t* src_xy,
                          const int32_t* start_xy, int32_t range, int32_t* pos_out, int32_t* found)
{
  const int S = h ? (int)h->kf_host.size() : 0;
  if (!h || kf_src < 0 || kf_src >= S || kf_dst < 0 || kf_dst >= S || level < 0 || level >= MCP_LEVELS || n < 0 || n > h->cfg.max_patches ||
      (n && (!src_xy || !start_xy || !pos_out || !found))) {
    set_last_error("mcp_fe_minipatch_find: bad arguments");
    return MCP_ERR_INVALID;
  }
  if (!h->kf_valid[kf_src] || !h->kf_valid[kf_dst]) { set_last_error("mcp_fe_minipatch_find: keyframe slot %d is empty", h->kf_valid[kf_src] ? kf_dst : kf_src); return MCP_ERR_STATE; }
  if (n == 0) return MCP_OK;
  cudaSetDevice(h->device);
  cudaStream_t s = h->stream;
  MCP_CUDA_CHECK(cudaMemcpyAsync(h->xy_dev, src_xy, sizeof(int2) * (size_t)n, cudaMemcpyHostToDevice, s));
  MCP_CUDA_CHECK(cudaMemcpyAsync(h->xy2_dev, start_xy, sizeof(int2) * (size_t)n, cudaMemcpyHostToDevice, s));
  // n_corners of the destination level is read from the device-side meta through the staged copy of the last keyframe
  FeMeta meta;
  MCP_CUDA_CHECK(cudaMemcpyAsync(&meta, h->kf_host[kf_dst].meta, sizeof(FeMeta), cudaMemcpyDeviceToHost, s));
  MCP_CUDA_CHECK(cudaStreamSynchronize(s));
  const int nc = std::min(meta.lv[level].n_corners, h->cfg.max_corners_per_level);
  fe_launch_minipatch(h->kf_host[kf_src].lv[level], h->kf_host[kf_dst].lv[level], nc, n, h->xy_dev, h->xy2_dev, range, h->pos_dev, h->found_dev, s);
  MCP_CUDA_CHECK(cudaMemcpyAsync(pos_out, h->pos_dev, sizeof(int2) * (size_t)n, cudaMemcpyDeviceToHost, s));
  MCP_CUDA_CHECK(cudaMemcpyAsync(found, h->found_dev, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s));
  MCP_CUDA_CHECK(cudaStreamSynchronize(s));
  return MCP_OK;
}

void mcp_fe_default_rest_config(McpRestConfig* c)
{
  memset(c, 0, sizeof(*c));
  c->use_shi = 0; c->use_thresh = 0;            // src/KeyFrame.cc:68-69 ("fast", "percent")
  c->top_fraction = 0.8; c->thresh = 70;        // :67, :64
  c->nonmax_strict = 0;
  c->prev_slot = -1; c->n_prev = 0;
}

int mcp_fe_make_keyframe_rest(McpFe* h, int32_t slot, const McpRestConfig* cfg, McpRestLevelOut out[MCP_LEVELS])
{
  const int S = h ? (int)h->kf_host.size() : 0;
  if (!h || !cfg || !out || slot < 0 || slot >= S || !h->kf_valid[slot] || cfg->prev_slot >= S ||
      (cfg->prev_slot >= 0 && (!h->kf_valid[cfg->prev_slot] || cfg->prev_slot == slot || cfg->n_prev < 1))) {
    set_last_error("mcp_fe_make_keyframe_rest: bad arguments (slot / prev_slot must hold keyframes)");
    return MCP_ERR_INVALID;
  }
  cudaSetDevice(h->device);
  cudaStream_t s = h->stream;
  const size_t cap = (size_t)h->cfg.max_corners_per_level;
  // scratch per level: flag, f1, f2 (int) | score (double) | cand, cand_out (16 B) | cand_xy, pos1, pos2 (int2) | 8 counters
  const size_t per_level = align_up(cap * (3 * sizeof(int) + sizeof(double) + 2 * sizeof(McpCandidate) + 3 * sizeof(int2)) + 8 * 256, 256);
  if (!h->rest_buf) {
    MCP_CUDA_CHECK(cudaMalloc(&h->rest_buf, per_level * MCP_LEVELS));
    MCP_CUDA_CHECK(cudaMallocHost(&h->rest_host, (sizeof(McpCandidate) * cap + 256) * MCP_LEVELS));
  }
  FeMeta meta, meta_prev;
  memset(&meta_prev, 0, sizeof(meta_prev));
  MCP_CUDA_CHECK(cudaMemcpyAsync(&meta, h->kf_host[slot].meta, sizeof(FeMeta), cudaMemcpyDeviceToHost, s));
  if (cfg->prev_slot >= 0) MCP_CUDA_CHECK(cudaMemcpyAsync(&meta_prev, h->kf_host[cfg->prev_slot].meta, sizeof(FeMeta), cudaMemcpyDeviceToHost, s));
  MCP_CUDA_CHECK(cudaStreamSynchronize(s));
  McpCandidate* host_c = reinterpret_cast<McpCandidate*>(h->rest_host);
  int* host_ctr = reinterpret_cast<int*>(host_c + cap * MCP_LEVELS);
  const bool prune = cfg->prev_slot >= 0 && cfg->n_prev > 0;
  for (int l = 0; l < MCP_LEVELS; l++) {
    uint8_t* b = h->rest_buf + per_level * l;
    int* ctr = reinterpret_cast<int*>(b);
    double* score = reinterpret_cast<double*>(b + 256);
    McpCandidate* cand = reinterpret_cast<McpCandidate*>(score + cap);
    McpCandidate* cand_out = cand + cap;
    int2* cand_xy = reinterpret_cast<int2*>(cand_out + cap);
    int2* pos1 = cand_xy + cap; int2* pos2 = pos1 + cap;
    int* flag = reinterpret_cast<int*>(pos2 + cap);
    int* f1 = flag + cap; int* f2 = f1 + cap;
    MCP_CUDA_CHECK(cudaMemsetAsync(ctr, 0, 256, s));
    const int n = std::min(meta.lv[l].n_corners, (int)cap);
    const int np = std::min(meta_prev.lv[l].n_corners, (int)cap);
    fe_launch_rest_level(h->kf_host[slot], prune ? &h->kf_host[cfg->prev_slot] : nullptr, l, n, np, h->cfg.adaptive_thresh, cfg->nonmax_strict,
                         cfg->use_shi, cfg->use_thresh, cfg->top_fraction, cfg->thresh, cfg->n_prev, flag, score, cand, cand_out, cand_xy,
                         pos1, pos2, f1, f2, ctr, s);
    MCP_CUDA_CHECK(cudaMemcpyAsync(host_ctr + 4 * l, ctr, sizeof(int) * 4, cudaMemcpyDeviceToHost, s));
    MCP_CUDA_CHECK(cudaMemcpyAsync(host_c + cap * l, prune ? cand_out : cand, sizeof(McpCandidate) * (size_t)std::max(n, 1), cudaMemcpyDeviceToHost, s));
  }
  MCP_CUDA_CHECK(cudaStreamSynchronize(s));
  for (int l = 0; l < MCP_LEVELS; l++) {
    const int* c = host_ctr + 4 * l;
    out[l].n_max = c[0]; out[l].n_selected = c[1]; out[l].n_candidates = prune ? c[2] : c[1];
    if (out[l].cand && out[l].cap > 0)
      memcpy(out[l].cand, host_c + cap * l, sizeof(McpCandidate) * (size_t)std::min(out[l].n_candidates, out[l].cap));
  }
  return MCP_OK;
}

/* Debug: FAST score map of one level (0 = not a corner at b=5, else fast_corner_score_10), width*height bytes. */
int mcp_fe_debug_scores(McpFe* h, int32_t slot, int32_t level, uint8_t* out)
{
  if (!h || !out || slot < 0 || slot >= (int)h->kf_host.size() || level < 0 || level >= MCP_LEVELS) return MCP_ERR_INVALID;
  cudaSetDevice(h->device);
  const FeLevel& L = h->kf_host[slot].lv[level];
  MCP_CUDA_CHECK(cudaMemcpy2D(out, L.w, L.score, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost));
  return MCP_OK;
}

int mcp_fe_set_camera(McpFe* h, const McpTaylorCam* s)
{
  if (!h || !s || s->n_inv < 2 || s->n_inv > 32) { set_last_error("mcp_fe_set_camera: bad arguments"); return MCP_ERR_INVALID; }
  DevCam& c = h->cam;
  memset(&c, 0, sizeof(c));
  memcpy(c.poly, s->poly, sizeof(c.poly));
  c.dmod[0] = -s->poly[0]; c.dmod[1] = s->poly[1]; c.dmod[2] = s->poly[2]; c.dmod[3] = 2 * s->poly[3]; c.dmod[4] = 3 * s->poly[4];
  memcpy(c.center, s->center, sizeof(c.center)); memcpy(c.affine, s->affine, sizeof(c.affine)); memcpy(c.image_size, s->image_size, sizeof(c.image_size));
  c.min_theta = s->min_theta; c.theta_mean = s->theta_mean; c.theta_std = s->theta_std; c.n_inv = s->n_inv;
  memcpy(c.inv, s->inv_poly, sizeof(double) * 32);
  h->has_cam = true;
  return MCP_OK;
}

int mcp_fe_project_points(McpFe* h, const double* cam_from_world, int32_t n, const double* world_xyz, const double* pixel_right_w,
                          const double* pixel_down_w, McpProjRes* out)
{
  if (!h || !cam_from_world || n < 0 || (n && (!world_xyz || !pixel_right_w || !pixel_down_w || !out))) { set_last_error("mcp_fe_project_points: bad arguments"); return MCP_ERR_INVALID; }
  if (!h->has_cam) { set_last_error("mcp_fe_project_points: call mcp_fe_set_camera first"); return MCP_ERR_STATE; }
  if (n == 0) return MCP_OK;
  cudaSetDevice(h->device);
  cudaStream_t s = h->stream;
  if ((size_t)n > h->proj_cap) {
    if (h->proj_in) cudaFree(h->proj_in);
    if (h->proj_out) cudaFree(h->proj_out);
  if (h->pu_buf) cudaFree(h->pu_buf);
    h->proj_cap = (size_t)n + n / 4 + 64;
    MCP_CUDA_CHECK(cudaMalloc(&h->proj_in, sizeof(double) * 9 * h->proj_cap));
    MCP_CUDA_CHECK(cudaMalloc(&h->proj_out, sizeof(McpProjRes) * h->proj_cap));
  }
  double* pw = h->proj_in; double* rw = pw + 3 * h->proj_cap; double* dw = rw + 3 * h->proj_cap;
  MCP_CUDA_CHECK(cudaMemcpyAsync(pw, world_xyz, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, s));
  MCP_CUDA_CHECK(cudaMemcpyAsync(rw, pixel_right_w, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, s));
  MCP_CUDA_CHECK(cudaMemcpyAsync(dw, pixel_down_w, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, s));
  Se3 T;
  memcpy(T.R, cam_from_world, sizeof(double) * 9); memcpy(T.t, cam_from_world + 9, sizeof(double) * 3);
  fe_launch_project(h->cam, T, n, pw, rw, dw, h->proj_out, s);
  MCP_CUDA_CHECK(cudaMemcpyAsync(out, h->proj_out, sizeof(McpProjRes) * (size_t)n, cudaMemcpyDeviceToHost, s));
  MCP_CUDA_CHECK(cudaStreamSynchronize(s));
  return MCP_OK;
}

static int ensure_pu(McpFe* h, size_t n)
{
  if (n <= h->pu_cap) return MCP_OK;
  if (h->pu_buf) cudaFree(h->pu_buf);
  h->pu_cap = n + n / 4 + 64;
  const size_t per = sizeof(McpPoseMeas) + sizeof(McpJacRes) + sizeof(double) * 4 + sizeof(int);
  MCP_CUDA_CHECK(cudaMalloc(&h->pu_buf, per * h->pu_cap + sizeof(McpPoseUpdate) + 256));
  return MCP_OK;
}

int mcp_fe_calc_jacobians(McpFe* h, const double* base_from_world, const double* cam_from_base, int32_t n, const double* world_xyz, McpJacRes* out)
{
  if (!h || !base_from_world || !cam_from_base || n < 0 || (n && (!world_xyz || !out))) { set_last_error("mcp_fe_calc_jacobians: bad arguments"); return MCP_ERR_INVALID; }
  if (!h->has_cam) { set_last_error("mcp_fe_calc_jacobians: call mcp_fe_set_camera first"); return MCP_ERR_STATE; }
  if (n == 0) return MCP_OK;
  cudaSetDevice(h->device);
  int rc = ensure_pu(h, (size_t)n);
  if (rc) return rc;
  cudaStream_t s = h->stream;
  double* pw = reinterpret_cast<double*>(h->pu_buf);
  McpJacRes* jr = reinterpret_cast<McpJacRes*>(pw + 4 * h->pu_cap);
  MCP_CUDA_CHECK(cudaMemcpyAsync(pw, world_xyz, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, s));
  Se3 B, Cb;
  memcpy(B.R, base_from_world, 72); memcpy(B.t, base_from_world + 9, 24);
  memcpy(Cb.R, cam_from_base, 72); memcpy(Cb.t, cam_from_base + 9, 24);
  fe_launch_calc_jacobians(h->cam, B, Cb, n, pw, jr, s);
  MCP_CUDA_CHECK(cudaMemcpyAsync(out, jr, sizeof(McpJacRes) * (size_t)n, cudaMemcpyDeviceToHost, s));
  MCP_CUDA_CHECK(cudaStreamSynchronize(s));
  return MCP_OK;
}

int mcp_fe_pose_update(McpFe* h, int32_t n, const McpPoseMeas* meas, int32_t estimator, double override_sigma, McpPoseUpdate* out, int32_t* outlier)
{
  if (!h || n < 0 || (n && !meas) || !out || estimator < 0 || estimator > 2) { set_last_error("mcp_fe_pose_update: bad arguments"); return MCP_ERR_INVALID; }
  cudaSetDevice(h->device);
  int rc = ensure_pu(h, (size_t)std::max(n, 1));
  if (rc) return rc;
  cudaStream_t s = h->stream;
  // layout: [pw 4*cap doubles | McpJacRes cap | McpPoseMeas cap | int cap | McpPoseUpdate]; e2 scratch aliases the pw region
  double* e2 = reinterpret_cast<double*>(h->pu_buf);
  McpJacRes* jr = reinterpret_cast<McpJacRes*>(e2 + 4 * h->pu_cap);
  McpPoseMeas* dm = reinterpret_cast<McpPoseMeas*>(jr + h->pu_cap);
  int* dout = reinterpret_cast<int*>(dm + h->pu_cap);
  McpPoseUpdate* dres = reinterpret_cast<McpPoseUpdate*>(reinterpret_cast<char*>(dout + h->pu_cap) + 128 - (reinterpret_cast<uintptr_t>(dout + h->pu_cap) % 128));
  if (n) MCP_CUDA_CHECK(cudaMemcpyAsync(dm, meas, sizeof(McpPoseMeas) * (size_t)n, cudaMemcpyHostToDevice, s));
  fe_launch_pose_update(n, dm, estimator, override_sigma, e2, dres, dout, s);
  MCP_CUDA_CHECK(cudaMemcpyAsync(out, dres, sizeof(McpPoseUpdate), cudaMemcpyDeviceToHost, s));
  if (outlier && n) MCP_CUDA_CHECK(cudaMemcpyAsync(outlier, dout, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s));
  MCP_CUDA_CHECK(cudaStreamSynchronize(s));
  return MCP_OK;
}

int mcp_fe_get_timing(McpFe* h, McpFeTiming* out) { if (!h || !out) return MCP_ERR_INVALID; *out = h->timing; return MCP_OK; }

}  // extern "C"
