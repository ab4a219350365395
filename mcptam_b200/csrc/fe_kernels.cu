// fe_kernels.cu — sm_100a kernels of the per-frame front end (integer / byte work, bit-exact by construction).
//
//   k_halfsample_fused / k_halfsample   CVD::halfSample x3                      (src/KeyFrame.cc:189-190)
//   k_fast_score     fast_corner_detect_10(b=5) + fast_corner_score_10 + score histogram   (:259-275)
//   k_fast_count     adaptive threshold from the histogram (:279-300), per-row corner counts
//   k_fast_scan      exclusive scan of the row counts = Level::vCornerRowLUT (:348-355)
//   k_fast_compact   raster-ordered corner list (mask + threshold filter, :302-312)
//   k_patch_search   one warp per patch: CVD::transform template (src/PatchFinder.cc:135-182), template sums,
//                    FindPatchCoarse over the row LUT / exhaustive disc (:229-355) with ZMSSDAtPoint (:511-658),
//                    MakeSubPixTemplate + IterateSubPixToConvergence (:362-470)
//   k_shitomasi      FindShiTomasiScoreAtPoint (src/ShiTomasi.cc:34-63)
//   k_minipatch      MiniPatch::SampleFromImage / FindPatch / SSDAtPoint (src/MiniPatch.cc:34-122)
#include <cuda.h>

#include "fe_types.cuh"

namespace mcp {

// ---------------------------------------------------------------------------------------------
// pyramid
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned avg4(unsigned a, unsigned b, unsigned c, unsigned d, int rnd) { return (a + b + c + d + (rnd ? 2u : 0u)) >> 2; }

// One thread per 8x8 level-0 block -> 4x4 L1, 2x2 L2, 1 L3 pixel (requires w,h multiples of 8).
__global__ void __launch_bounds__(128) k_halfsample_fused(FeKf kf, int rnd)
{
  const int bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y;
  const int w3 = kf.lv[3].w;
  if (bx >= w3) return;
  const uint8_t* src = kf.lv[0].img + (size_t)(8 * by) * kf.lv[0].pitch + 8 * bx;
  unsigned l1[4][4];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const uint2 a = *reinterpret_cast<const uint2*>(src + (size_t)(2 * r) * kf.lv[0].pitch);
    const uint2 b = *reinterpret_cast<const uint2*>(src + (size_t)(2 * r + 1) * kf.lv[0].pitch);
    const unsigned aw[2] = { a.x, a.y }, bw[2] = { b.x, b.y };
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const unsigned wa = aw[c >> 1] >> (16 * (c & 1)), wb = bw[c >> 1] >> (16 * (c & 1));
      l1[r][c] = avg4(wa & 0xff, (wa >> 8) & 0xff, wb & 0xff, (wb >> 8) & 0xff, rnd);
    }
  }
  uint8_t* d1 = kf.lv[1].img + (size_t)(4 * by) * kf.lv[1].pitch + 4 * bx;
#pragma unroll
  for (int r = 0; r < 4; r++)
    *reinterpret_cast<unsigned*>(d1 + (size_t)r * kf.lv[1].pitch) = l1[r][0] | (l1[r][1] << 8) | (l1[r][2] << 16) | (l1[r][3] << 24);
  unsigned l2[2][2];
#pragma unroll
  for (int r = 0; r < 2; r++)
#pragma unroll
    for (int c = 0; c < 2; c++) l2[r][c] = avg4(l1[2 * r][2 * c], l1[2 * r][2 * c + 1], l1[2 * r + 1][2 * c], l1[2 * r + 1][2 * c + 1], rnd);
  uint8_t* d2 = kf.lv[2].img + (size_t)(2 * by) * kf.lv[2].pitch + 2 * bx;
  *reinterpret_cast<unsigned short*>(d2) = (unsigned short)(l2[0][0] | (l2[0][1] << 8));
  *reinterpret_cast<unsigned short*>(d2 + kf.lv[2].pitch) = (unsigned short)(l2[1][0] | (l2[1][1] << 8));
  kf.lv[3].img[(size_t)by * kf.lv[3].pitch + bx] = (uint8_t)avg4(l2[0][0], l2[0][1], l2[1][0], l2[1][1], rnd);
}
// generic single level (any size): out = in.size()/2
__global__ void k_halfsample(const uint8_t* in, int in_pitch, uint8_t* out, int out_pitch, int ow, int oh, int rnd)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= ow || y >= oh) return;
  const uint8_t* r0 = in + (size_t)(2 * y) * in_pitch + 2 * x;
  const uint8_t* r1 = r0 + in_pitch;
  out[(size_t)y * out_pitch + x] = (uint8_t)avg4(r0[0], r0[1], r1[0], r1[1], rnd);
}

// ---------------------------------------------------------------------------------------------
// FAST-10
// ---------------------------------------------------------------------------------------------
__constant__ int c_ring_dx[16] = { 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1 };
__constant__ int c_ring_dy[16] = { -3, -3, -2, -1, 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3 };

constexpr int FT_W = 32, FT_H = 8;       // output tile
constexpr int FT_SW = FT_W + 6, FT_SH = FT_H + 6;

__device__ __forceinline__ bool has_run10(unsigned m)
{
  unsigned dd = m | (m << 16);
  unsigned r = dd & (dd >> 1);      // runs of 2
  r = r & (r >> 2);                 // 4
  r = r & (r >> 4);                 // 8
  r = r & (dd >> 8) & (dd >> 9);    // 10
  return (r & 0xffffu) != 0;
}

// blocks are assigned to levels by prefix (kf.tile_off[l])
__global__ void __launch_bounds__(FT_W* FT_H) k_fast_score(FeKf kf)
{
  __shared__ uint8_t tile[FT_SH][FT_SW + 2];
  __shared__ unsigned hist[32];
  int l = 0;
  while (l < MCP_LEVELS - 1 && (int)blockIdx.x >= kf.tile_off[l + 1]) l++;
  const FeLevel L = kf.lv[l];
  const int t = blockIdx.x - kf.tile_off[l];
  const int tiles_x = (L.w + FT_W - 1) / FT_W;
  const int x0 = (t % tiles_x) * FT_W, y0 = (t / tiles_x) * FT_H;
  const int tid = threadIdx.y * FT_W + threadIdx.x;
  if (tid < 32) hist[tid] = 0;
  for (int i = tid; i < FT_SH * FT_SW; i += FT_W * FT_H) {
    const int sy = i / FT_SW, sx = i - sy * FT_SW;
    const int gx = x0 + sx - 3, gy = y0 + sy - 3;
    tile[sy][sx] = (gx >= 0 && gy >= 0 && gx < L.w && gy < L.h) ? L.img[(size_t)gy * L.pitch + gx] : 0;
  }
  __syncthreads();
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  int score = 0;
  if (x >= 3 && y >= 3 && x < L.w - 3 && y < L.h - 3) {
    const int sx = threadIdx.x + 3, sy = threadIdx.y + 3;
    const int c = tile[sy][sx];
    const int b = MCP_MIN_FAST_THRESH;
    const int cb = c + b, c_b = c - b;
    int dv[16];
    unsigned br = 0, dk = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const int v = tile[sy + c_ring_dy[i]][sx + c_ring_dx[i]];
      dv[i] = v - c;
      br |= (unsigned)(v > cb) << i;
      dk |= (unsigned)(v < c_b) << i;
    }
    if (has_run10(br) || has_run10(dk)) {
      // fast_corner_score_10: largest t for which the pixel is still a corner = max over arcs of (min |diff|) - 1
      // fast_corner_score_10 exactly as libCVD does it: bisection on the threshold (bmin = b, bmax = 255),
      // each probe re-evaluating the FAST-10 criterion with compare masks only.
      int bmin = b, bmax = 255, tt = (bmax + bmin) / 2;
      for (;;) {
        unsigned mb = 0, md = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) { mb |= (unsigned)(dv[i] > tt) << i; md |= (unsigned)(dv[i] < -tt) << i; }
        if (has_run10(mb) || has_run10(md)) bmin = tt; else bmax = tt;
        if (bmin == bmax - 1 || bmin == bmax) break;
        tt = (bmin + bmax) / 2;
      }
      const int best = bmin;
      score = best;                                             // 5..254
#ifdef MCP_FE_DEBUG
      if (l == 0 && x == 488 && y == 3) { printf("dbg c=%d br=%x dk=%x best=%d dv:", c, br, dk, best); for (int i = 0; i < 16; i++) printf(" %d", dv[i]); printf("\n"); }
#endif
      atomicAdd(&hist[min(score, MCP_MAX_FAST_THRESH)], 1u);
    }
  }
  if (x < L.w && y < L.h) L.score[(size_t)y * L.pitch + x] = (uint8_t)score;
  __syncthreads();
  if (tid < 32 && hist[tid]) atomicAdd(&L.hist[tid], hist[tid]);
}

// adaptive threshold from the capped-score histogram (src/KeyFrame.cc:264-300)
__device__ __forceinline__ int fast_threshold(const unsigned* hist, int w, int h, int* freq_out /*31 or null*/)
{
  int freq[32];
  int acc = 0;
  for (int t = 31; t >= 0; t--) { acc += (t <= MCP_MAX_FAST_THRESH && t >= MCP_MIN_FAST_THRESH) ? (int)hist[t] : 0; freq[t] = (t >= MCP_MIN_FAST_THRESH && t <= MCP_MAX_FAST_THRESH) ? acc : 0; }
  if (freq_out) for (int t = 0; t <= 30; t++) freq_out[t] = freq[t];
  const double target = -1 * (w * h) / 500.0;
  int thr = MCP_MIN_FAST_THRESH;
  for (int t = MCP_MIN_FAST_THRESH; t <= MCP_MAX_FAST_THRESH; ++t) {
    double deriv;
    if (t == MCP_MIN_FAST_THRESH) deriv = freq[t + 1] - freq[t];
    else if (t == MCP_MAX_FAST_THRESH) deriv = freq[t] - freq[t - 1];
    else deriv = (freq[t + 1] - freq[t - 1]) / 2.0;
    thr = t;
    if (deriv > target) break;
  }
  return thr;
}

// one warp per image row (all levels): threshold + mask filter count
__global__ void __launch_bounds__(256) k_fast_count(FeKf kf, int adaptive)
{
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int l = 0;
  while (l < MCP_LEVELS - 1 && row >= kf.row_off[l + 1]) l++;
  if (row >= kf.row_off[MCP_LEVELS]) return;
  const FeLevel L = kf.lv[l];
  const int y = row - kf.row_off[l];
  int thr;
  if (adaptive) {
    thr = 0;
    if (lane == 0) thr = fast_threshold(L.hist, L.w, L.h, y == 0 ? kf.meta->lv[l].fast_freq : nullptr);
    thr = __shfl_sync(0xffffffffu, thr, 0);
  } else {
    thr = kf.fixed_thresh[l];
    if (y == 0 && lane == 0) for (int t = 0; t <= 30; t++) kf.meta->lv[l].fast_freq[t] = 0;
  }
  if (y == 0 && lane == 0) kf.meta->lv[l].fast_thresh = thr;
  int cnt = 0;
  const uint8_t* sc = L.score + (size_t)y * L.pitch;
  const uint8_t* mk = (adaptive && L.mask) ? L.mask + (size_t)y * L.pitch : nullptr;
  for (int x = lane; x < L.w; x += 32) {
    const int s = sc[x];
    cnt += (s >= thr && s > 0 && (!mk || mk[x] == 255)) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) L.rowcount[y] = cnt;
}

// one block per level: exclusive scan of the row counts -> row LUT, total
__global__ void __launch_bounds__(512) k_fast_scan(FeKf kf)
{
  __shared__ int wsum[16];
  __shared__ int carry;
  const FeLevel L = kf.lv[blockIdx.x];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < L.h; base += blockDim.x) {
    const int y = base + threadIdx.x;
    const int v = y < L.h ? L.rowcount[y] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    int off = carry;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) off += wsum[w];
    if (y < L.h) L.row_lut[y] = off + incl - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = off + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) { kf.meta->lv[blockIdx.x].n_corners = carry; kf.meta->lv[blockIdx.x].width = L.w; kf.meta->lv[blockIdx.x].height = L.h; }
}

// one warp per row: ordered compaction
__global__ void __launch_bounds__(256) k_fast_compact(FeKf kf, int adaptive)
{
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int l = 0;
  while (l < MCP_LEVELS - 1 && row >= kf.row_off[l + 1]) l++;
  if (row >= kf.row_off[MCP_LEVELS]) return;
  const FeLevel L = kf.lv[l];
  const int y = row - kf.row_off[l];
  const int thr = kf.meta->lv[l].fast_thresh;
  int pos = L.row_lut[y];
  const uint8_t* sc = L.score + (size_t)y * L.pitch;
  const uint8_t* mk = (adaptive && L.mask) ? L.mask + (size_t)y * L.pitch : nullptr;
  for (int x0 = 0; x0 < L.w; x0 += 32) {
    const int x = x0 + lane;
    bool keep = false;
    if (x < L.w) { const int s = sc[x]; keep = (s >= thr && s > 0 && (!mk || mk[x] == 255)); }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int idx = pos + __popc(m & ((1u << lane) - 1));
      if (idx < kf.corner_cap) L.corners[idx] = make_int2(x, y);
    }
    pos += __popc(m);
  }
}

// ---------------------------------------------------------------------------------------------
// FAST-10 in two launches (default; MCP_FE_FAST_FUSED=0 keeps the four-kernel path above).
//   k_fast_score_rows  per 32x8 tile: the 38x14 footprint is staged with 16-byte loads (aligned 64-byte rows), the
//                      FAST-10 test runs on compare masks, and the score of a corner -- the largest t at which it still is
//                      one, libCVD's fast_corner_score_10 bisection -- is evaluated directly as
//                          max over the 16 arcs of (min over the 10 ring pixels of the arc of +-(p - c)) - 1
//                      (the criterion is monotone in t, so this is what the bisection converges to) with the 16-bit SIMD
//                      min3/max3 instructions, bright and dark arcs side by side in one register.  Besides the level
//                      histogram of capped scores (adaptive threshold, src/KeyFrame.cc:264-300) every tile adds its
//                      corners that pass the mask to a per-ROW histogram; the last tile of a level (ticket counter) turns
//                      the histogram into the threshold, the row histograms into row counts for that threshold, scans
//                      them into Level::vCornerRowLUT and re-arms histogram, row histograms and ticket for the next frame.
//   k_fast_compact4    one warp per row, four pixels per lane: ordered compaction into the raster-ordered corner list.
// Both are launched with programmatic stream serialisation behind the pyramid kernel.
// ---------------------------------------------------------------------------------------------
constexpr int FT_LW = 64;                 // staged bytes per tile row: the aligned span [x0 - 16, x0 + 48)

// the host mirror of a location in the slot's output block (pinned, mapped: the kernels write the results the host reads
// straight into host memory, so no device->host copy follows them)
template <typename T>
__device__ __forceinline__ T* fe_mirror(const FeKf& kf, T* dev) { return reinterpret_cast<T*>(reinterpret_cast<char*>(dev) + kf.host_delta); }

__device__ __forceinline__ unsigned pack_pm(int d) { return ((unsigned)d & 0xffffu) | ((unsigned)(-d) << 16); }   // (d, -d) as s16x2

__device__ __forceinline__ int fast_score_direct(const int (&dv)[16])
{
  unsigned q[16], m3[16];
#pragma unroll
  for (int i = 0; i < 16; i++) q[i] = pack_pm(dv[i]);
#pragma unroll
  for (int i = 0; i < 16; i++) m3[i] = __vimin3_s16x2(q[i], q[(i + 1) & 15], q[(i + 2) & 15]);
  unsigned best = 0x80008000u;            // (-32768, -32768)
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const unsigned m9 = __vimin3_s16x2(m3[i], m3[(i + 3) & 15], m3[(i + 6) & 15]);
    const unsigned m10 = __vimin3_s16x2(m9, q[(i + 9) & 15], q[(i + 9) & 15]);
    best = __vimax3_s16x2(best, m10, m10);
  }
  const int bright = (int)(short)(best & 0xffffu), dark = (int)(short)(best >> 16);
  return max(bright, dark) - 1;
}

__global__ void __launch_bounds__(FT_W* FT_H) k_fast_score_rows(FeKf kf, int adaptive)
{
  pdl_prologue();
  __shared__ __align__(16) uint8_t tile[FT_SH][FT_LW];
  __shared__ unsigned hist[32];
  __shared__ unsigned rowh[FT_H][32];
  __shared__ int s_last, s_thr, carry;
  __shared__ int wsum[8];
  int l = 0;
  while (l < MCP_LEVELS - 1 && (int)blockIdx.x >= kf.tile_off[l + 1]) l++;
  const FeLevel L = kf.lv[l];
  const int t = blockIdx.x - kf.tile_off[l];
  const int tiles_x = (L.w + FT_W - 1) / FT_W;
  const int x0 = (t % tiles_x) * FT_W, y0 = (t / tiles_x) * FT_H;
  const int tid = threadIdx.y * FT_W + threadIdx.x;
  if (tid < 32) hist[tid] = 0;
  rowh[threadIdx.y][threadIdx.x] = 0;
  if (tid < FT_SH * (FT_LW / 16)) {
    const int sy = tid >> 2, ch = tid & 3;
    const int gy = y0 + sy - 3, gx = x0 - 16 + 16 * ch;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (gy >= 0 && gy < L.h && gx >= 0 && gx < L.pitch) v = *reinterpret_cast<const uint4*>(L.img + (size_t)gy * L.pitch + gx);
    *reinterpret_cast<uint4*>(&tile[sy][16 * ch]) = v;
  }
  __syncthreads();
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  const bool use_mask = adaptive && L.mask;
  int score = 0;
  if (x >= 3 && y >= 3 && x < L.w - 3 && y < L.h - 3) {
    const int sx = threadIdx.x + 16, sy = threadIdx.y + 3;
    const int c = tile[sy][sx];
    const int b = MCP_MIN_FAST_THRESH;
    const int cb = c + b, c_b = c - b;
    int dv[16];
    unsigned br = 0, dk = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const int v = tile[sy + c_ring_dy[i]][sx + c_ring_dx[i]];
      dv[i] = v - c;
      br |= (unsigned)(v > cb) << i;
      dk |= (unsigned)(v < c_b) << i;
    }
    if (has_run10(br) || has_run10(dk)) {
      score = fast_score_direct(dv);                             // 5..254
      atomicAdd(&hist[min(score, MCP_MAX_FAST_THRESH)], 1u);
      if (!use_mask || L.mask[(size_t)y * L.pitch + x] == 255) {
        // adaptive: binned by capped score (the threshold is not known yet); fixed threshold: bin 31 = kept
        if (adaptive) atomicAdd(&rowh[threadIdx.y][min(score, MCP_MAX_FAST_THRESH)], 1u);
        else if (score >= kf.fixed_thresh[l]) atomicAdd(&rowh[threadIdx.y][31], 1u);
      }
    }
  }
  if (x < L.w && y < L.h) L.score[(size_t)y * L.pitch + x] = (uint8_t)score;
  __syncthreads();
  if (tid < 32 && hist[tid]) atomicAdd(&L.hist[tid], hist[tid]);
  {
    const unsigned v = rowh[threadIdx.y][threadIdx.x];
    if (v && y0 + (int)threadIdx.y < L.h) atomicAdd(reinterpret_cast<unsigned*>(L.rowhist) + (size_t)(y0 + threadIdx.y) * 32 + threadIdx.x, v);
  }
  __threadfence();
  __syncthreads();
  const int n_tiles = kf.tile_off[l + 1] - kf.tile_off[l];
  if (tid == 0) s_last = (atomicAdd(L.rowhist + (size_t)L.h * 32, 1) == n_tiles - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  // ---- last tile of the level: threshold, row counts, row LUT ---------------------------------------------------
  __threadfence();
  if (tid == 0) {
    unsigned hh[32];
    for (int i = 0; i < 32; i++) hh[i] = __ldcg(&L.hist[i]);
    int thr;
    if (adaptive) thr = fast_threshold(hh, L.w, L.h, kf.meta->lv[l].fast_freq);
    else { thr = kf.fixed_thresh[l]; for (int i = 0; i <= 30; i++) kf.meta->lv[l].fast_freq[i] = 0; }
    kf.meta->lv[l].fast_thresh = thr;
    if (kf.host_delta) {
      FeMetaLevel* mh = &fe_mirror(kf, kf.meta)->lv[l];
      mh->fast_thresh = thr;
      for (int i = 0; i <= 30; i++) mh->fast_freq[i] = kf.meta->lv[l].fast_freq[i];
    }
    s_thr = thr;
    carry = 0;
    L.rowhist[(size_t)L.h * 32] = 0;                             // ticket re-armed
  }
  __syncthreads();
  if (tid < 32) L.hist[tid] = 0;
  const int lo_bin = adaptive ? s_thr : 31;
  for (int base = 0; base < L.h; base += FT_W * FT_H) {
    const int yy = base + tid;
    int v = 0;
    if (yy < L.h) {
      int4* rh = reinterpret_cast<int4*>(L.rowhist + (size_t)yy * 32);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int4 r = __ldcg(rh + k);
        v += (4 * k + 0 >= lo_bin ? r.x : 0) + (4 * k + 1 >= lo_bin ? r.y : 0) + (4 * k + 2 >= lo_bin ? r.z : 0) + (4 * k + 3 >= lo_bin ? r.w : 0);
        rh[k] = make_int4(0, 0, 0, 0);
      }
    }
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += u; }
    if ((tid & 31) == 31) wsum[tid >> 5] = incl;
    __syncthreads();
    int off = carry;
    for (int w = 0; w < (tid >> 5); w++) off += wsum[w];
    if (yy < L.h) {
      L.row_lut[yy] = off + incl - v;
      if (kf.host_delta) fe_mirror(kf, L.row_lut)[yy] = off + incl - v;
    }
    __syncthreads();
    if (tid == FT_W * FT_H - 1) carry = off + incl;
    __syncthreads();
  }
  if (tid == 0) {
    kf.meta->lv[l].n_corners = carry; kf.meta->lv[l].width = L.w; kf.meta->lv[l].height = L.h;
    if (kf.host_delta) { FeMetaLevel* mh = &fe_mirror(kf, kf.meta)->lv[l]; mh->n_corners = carry; mh->width = L.w; mh->height = L.h; }
  }
}

__global__ void __launch_bounds__(256) k_fast_compact4(FeKf kf, int adaptive)
{
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int l = 0;
  while (l < MCP_LEVELS - 1 && row >= kf.row_off[l + 1]) l++;
  const bool row_ok = row < kf.row_off[MCP_LEVELS];
  const FeLevel L = kf.lv[l];
  const int y = row_ok ? row - kf.row_off[l] : 0;
  const int thr = kf.meta->lv[l].fast_thresh;
  int pos = L.row_lut[y];
  // empty rows are skipped (the LUT is an exclusive scan: equal neighbours = no corner in the row)
  const bool work = row_ok && !(y + 1 < L.h && pos == L.row_lut[y + 1]);
  const unsigned* sc = reinterpret_cast<const unsigned*>(L.score + (size_t)y * L.pitch);
  const unsigned* mk = (adaptive && L.mask) ? reinterpret_cast<const unsigned*>(L.mask + (size_t)y * L.pitch) : nullptr;
  const unsigned lt = (1u << lane) - 1u;
  for (int x0 = 0; work && x0 < L.w; x0 += 128) {
    const int x = x0 + 4 * lane;
    unsigned keep = 0;
    if (x < L.w) {
      const unsigned s4 = sc[x >> 2];
      if (s4) {
        const unsigned m4 = mk ? mk[x >> 2] : 0xffffffffu;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int sv = (int)((s4 >> (8 * k)) & 255u);
          const bool ok = sv >= thr && sv > 0 && ((m4 >> (8 * k)) & 255u) == 255u && x + k < L.w;
          keep |= (unsigned)ok << k;
        }
      }
    }
    if (!__any_sync(0xffffffffu, keep != 0)) continue;
    const unsigned b0 = __ballot_sync(0xffffffffu, keep & 1u), b1 = __ballot_sync(0xffffffffu, keep & 2u);
    const unsigned b2 = __ballot_sync(0xffffffffu, keep & 4u), b3 = __ballot_sync(0xffffffffu, keep & 8u);
    int idx = pos + __popc(b0 & lt) + __popc(b1 & lt) + __popc(b2 & lt) + __popc(b3 & lt);
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (keep & (1u << k)) { if (idx < kf.corner_cap) L.corners[idx] = make_int2(x + k, y); idx++; }
    pos += __popc(b0) + __popc(b1) + __popc(b2) + __popc(b3);
  }
  if (!kf.host_delta || !work) return;
  // ---- host mirror: the warp ships its row's segment of the corner list to pinned host memory, 32 corners per store
  // instruction (per-corner stores would cross PCIe as thousands of 8-byte writes) ---------------------------------------------
  __syncwarp();
  const int beg = L.row_lut[y], end = min(pos, kf.corner_cap);
  int2* dst = fe_mirror(kf, L.corners);
  for (int i = beg + lane; i < end; i += 32) dst[i] = __ldcg(&L.corners[i]);
}

// ---------------------------------------------------------------------------------------------
// patch search
// ---------------------------------------------------------------------------------------------
// 8 bytes at an arbitrary address via two aligned 8-byte loads
__device__ __forceinline__ unsigned long long load_u8x8(const uint8_t* p)
{
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const unsigned long long* q = reinterpret_cast<const unsigned long long*>(a & ~(uintptr_t)7);
  const unsigned sh = (unsigned)(a & 7) * 8;
  const unsigned long long lo = __ldg(q);
  if (sh == 0) return lo;
  const unsigned long long hi = __ldg(q + 1);
  return (lo >> sh) | (hi << (64 - sh));
}

// ZMSSDAtPoint (src/PatchFinder.cc:511-658): integer arithmetic, truncating division
__device__ __forceinline__ int zmssd_at(const FeLevel& L, const unsigned long long* trow /*8 rows packed*/, int tsum, int tsumsq,
                                        int x, int y, int max_ssd)
{
  if (!(x >= 4 && y >= 4 && x < L.w - 4 && y < L.h - 4)) return max_ssd + 1;
  int isum = 0, isq = 0, cross = 0;
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const unsigned long long iv = load_u8x8(L.img + (size_t)(y - 4 + r) * L.pitch + (x - 4));
    const unsigned long long tv = trow[r];
#pragma unroll
    for (int c = 0; c < 8; c++) {
      const int n = (int)((iv >> (8 * c)) & 0xff), t = (int)((tv >> (8 * c)) & 0xff);
      isum += n; isq += n * n; cross += n * t;
    }
  }
  const int SA = tsum, SB = isum;
  return ((2 * SA * SB - SA * SA - SB * SB) / 64 + isq + tsumsq - 2 * cross);
}

// CVD::sample<byte,byte>: double bilinear, no fused multiply-add, truncating (or rounding) conversion
__device__ __forceinline__ uint8_t sample_u8(const FeLevel& L, double x, double y, int rnd)
{
  const int lx = (int)x, ly = (int)y;
  x = __dsub_rn(x, (double)lx); y = __dsub_rn(y, (double)ly);
  const uint8_t* p = L.img + (size_t)ly * L.pitch + lx;
  const double a = p[0], b = p[1], c = p[L.pitch], dd = p[L.pitch + 1];
  const double omx = __dsub_rn(1.0, x), omy = __dsub_rn(1.0, y);
  const double top = __dadd_rn(__dmul_rn(omx, a), __dmul_rn(x, b));
  const double bot = __dadd_rn(__dmul_rn(omx, c), __dmul_rn(x, dd));
  double v = __dadd_rn(__dmul_rn(omy, top), __dmul_rn(y, bot));
  if (rnd) v = __dadd_rn(v, 0.5);
  return (uint8_t)v;
}

// TooN::Cholesky<3> (LDL^T) + get_inverse(), no contraction
__device__ __forceinline__ void toon_chol3_inverse(const double* A, double* inv)
{
  double c[9];
#pragma unroll
  for (int i = 0; i < 9; i++) c[i] = A[i];
  bool stop = false;
  for (int col = 0; col < 3 && !stop; col++) {
    double inv_diag = 1;
    for (int row = col; row < 3; row++) {
      double val = c[row * 3 + col];
      for (int col2 = 0; col2 < col; col2++) val = __dsub_rn(val, __dmul_rn(c[col2 * 3 + col], c[row * 3 + col2]));
      if (row == col) { c[row * 3 + col] = val; if (val == 0) { stop = true; break; } inv_diag = __ddiv_rn(1.0, val); }
      else { c[col * 3 + row] = val; c[row * 3 + col] = __dmul_rn(val, inv_diag); }
    }
  }
  for (int k = 0; k < 3; k++) {
    double v[3] = { 0, 0, 0 }, y[3], r[3];
    v[k] = 1;
    for (int i = 0; i < 3; i++) { double val = v[i]; for (int j = 0; j < i; j++) val = __dsub_rn(val, __dmul_rn(c[i * 3 + j], y[j])); y[i] = val; }
    for (int i = 0; i < 3; i++) y[i] = __ddiv_rn(y[i], c[i * 3 + i]);
    for (int i = 2; i >= 0; i--) { double val = y[i]; for (int j = i + 1; j < 3; j++) val = __dsub_rn(val, __dmul_rn(c[j * 3 + i], r[j])); r[i] = val; }
    inv[0 * 3 + k] = r[0]; inv[1 * 3 + k] = r[1]; inv[2 * 3 + k] = r[2];
  }
}

__global__ void __launch_bounds__(128) k_patch_search(FeDev fe, int target_slot, int n, const McpPatchReq* __restrict__ req,
                                                     McpPatchRes* __restrict__ res, uint8_t* __restrict__ templ_out)
{
  __shared__ uint8_t s_t[4][64];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int i = blockIdx.x * (blockDim.x >> 5) + wid;
  if (i >= n) return;
  const McpPatchReq rq = req[i];
  McpPatchRes out;
  out.template_bad = 1; out.found = 0; out.did_subpix = 0; out.score = 0; out.coarse_x = 0; out.coarse_y = 0;
  out.found_x = 0; out.found_y = 0; out.n_candidates = 0; out.pad_ = 0;
  const int max_ssd = 8 * 8 * 250;                                  // src/PatchFinder.cc:44,61
  const bool args_ok = rq.src_kf >= 0 && rq.src_kf < fe.n_slots && rq.src_level >= 0 && rq.src_level < MCP_LEVELS &&
                       rq.search_level >= 0 && rq.search_level < MCP_LEVELS;
  if (!args_ok) { if (lane == 0) res[i] = out; return; }
  // ---- template: MakeTemplateCoarseCont -------------------------------------------------------------
  const FeLevel S = fe.kf[rq.src_kf].lv[rq.src_level];
  // m2 = M2Inverse(mm2WarpInverse) * LevelScale(mnSearchLevel)       (include/mcptam/SmallMatrixOpts.h)
  const double wi0 = rq.warp_inv[0], wi1 = rq.warp_inv[1], wi2 = rq.warp_inv[2], wi3 = rq.warp_inv[3];
  const double det = __dsub_rn(__dmul_rn(wi0, wi3), __dmul_rn(wi1, wi2));
  const double idet = __ddiv_rn(1.0, det);
  const double ls = (double)(1 << rq.search_level);
  const double M0 = __dmul_rn(__dmul_rn(wi3, idet), ls), M1 = __dmul_rn(__dmul_rn(-wi1, idet), ls);
  const double M2 = __dmul_rn(__dmul_rn(-wi2, idet), ls), M3 = __dmul_rn(__dmul_rn(wi0, idet), ls);
  const double ax = M0, ay = M2, dx = M1, dy = M3;                   // across = M.T()[0], down = M.T()[1]
  const double p0x = __dsub_rn((double)rq.src_cx, __dadd_rn(__dmul_rn(M0, 4.0), __dmul_rn(M1, 4.0)));
  const double p0y = __dsub_rn((double)rq.src_cy, __dadd_rn(__dmul_rn(M2, 4.0), __dmul_rn(M3, 4.0)));
  double min_x = p0x, min_y = p0y, max_x = p0x, max_y = p0y;
  if (ax < 0) min_x = __dadd_rn(min_x, __dmul_rn(8.0, ax)); else max_x = __dadd_rn(max_x, __dmul_rn(8.0, ax));
  if (dx < 0) min_x = __dadd_rn(min_x, __dmul_rn(8.0, dx)); else max_x = __dadd_rn(max_x, __dmul_rn(8.0, dx));
  if (ay < 0) min_y = __dadd_rn(min_y, __dmul_rn(8.0, ay)); else max_y = __dadd_rn(max_y, __dmul_rn(8.0, ay));
  if (dy < 0) min_y = __dadd_rn(min_y, __dmul_rn(8.0, dy)); else max_y = __dadd_rn(max_y, __dmul_rn(8.0, dy));
  const bool all_in = (min_x >= 0 && min_y >= 0 && max_x < S.w - 1 && max_y < S.h - 1);
  const double crx = __dsub_rn(dx, __dmul_rn(8.0, ax)), cry = __dsub_rn(dy, __dmul_rn(8.0, ay));
  const double xb = S.w - 1, yb = S.h - 1;
  // the reference accumulates p incrementally (p += across / carriage_return); replay the same sequence
  double px = p0x, py = p0y;
  int n_out = 0;
  uint8_t mine[2] = { 0, 0 };
  // every lane replays the whole position sequence (the additions are order dependent) and keeps the positions of its two
  // template pixels; the bilinear samples are then taken by all lanes at once (two global round trips per warp, not 64)
  double qx[2] = { 0, 0 }, qy[2] = { 0, 0 };
#pragma unroll
  for (int ii = 0; ii < 8; ++ii) {
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int k = ii * 8 + jj;
      const bool me = (k & 31) == lane;
      qx[k >> 5] = me ? px : qx[k >> 5]; qy[k >> 5] = me ? py : qy[k >> 5];
      px = __dadd_rn(px, ax); py = __dadd_rn(py, ay);
    }
    px = __dadd_rn(px, crx); py = __dadd_rn(py, cry);
  }
#pragma unroll
  for (int h2 = 0; h2 < 2; h2++) {
    uint8_t v = 0;
    if (all_in || (0 <= qx[h2] && 0 <= qy[h2] && qx[h2] < xb && qy[h2] < yb)) v = sample_u8(S, qx[h2], qy[h2], fe.transform_round);
    else n_out++;
    mine[h2] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_out += __shfl_xor_sync(0xffffffffu, n_out, o);
  s_t[wid][lane] = mine[0]; s_t[wid][lane + 32] = mine[1];
  __syncwarp();
  if (templ_out) { templ_out[(size_t)i * 64 + lane] = mine[0]; templ_out[(size_t)i * 64 + 32 + lane] = mine[1]; }
  const bool det_ok = isfinite(idet);
  out.template_bad = (n_out > 0 || !det_ok) ? 1 : 0;
  if (out.template_bad) { if (lane == 0) res[i] = out; return; }
  // template sums and packed rows
  int tsum = mine[0] + mine[1], tsq = mine[0] * mine[0] + mine[1] * mine[1];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { tsum += __shfl_xor_sync(0xffffffffu, tsum, o); tsq += __shfl_xor_sync(0xffffffffu, tsq, o); }
  unsigned long long trow[8];
#pragma unroll
  for (int r = 0; r < 8; r++) trow[r] = *reinterpret_cast<const unsigned long long*>(&s_t[wid][r * 8]);

  // ---- FindPatchCoarse --------------------------------------------------------------------------------
  const FeLevel T = fe.kf[target_slot].lv[rq.search_level];
  const int n_corners = min(fe.kf[target_slot].meta->lv[rq.search_level].n_corners, fe.kf[target_slot].corner_cap);
  const int lsc = 1 << rq.search_level;
  const int ipx = rq.pred_x / lsc, ipy = rq.pred_y / lsc;
  const unsigned nRange = ((unsigned)rq.range + lsc - 1) / lsc;
  int nTop = ipy - (int)nRange, nBottomPlusOne = ipy + (int)nRange + 1, nLeft = ipx - (int)nRange;
  const int nRight = ipx + (int)nRange;
  bool early = false;
  if (nTop < 0) nTop = 0;
  if (nTop >= T.h) early = true;
  if (nBottomPlusOne <= 0) early = true;
  if (nLeft < 0) nLeft = 0;
  if (nLeft >= T.w) early = true;
  int best_ssd = max_ssd + 1, best_idx = 0x7fffffff, best_x = 0, best_y = 0, n_valid = 0;
  if (rq.exhaustive == 2) {
    // PatchFinder::SetSubPixPos on a coarse match found earlier (AddPointEpipolar's second pass,
    // src/MapMakerServerBase.cc:822-846): pred_x / pred_y ARE the match in search-level coordinates, no search
    early = false; best_ssd = 0; best_x = rq.pred_x; best_y = rq.pred_y;
  } else if (!early) {
    if (rq.exhaustive) {
      const int y_end = min(nBottomPlusOne, T.h), x_end = min(nRight + 1, T.w);
      const int bw = x_end - nLeft, bh = y_end - nTop;
      const int total = (bw > 0 && bh > 0) ? bw * bh : 0;
      for (int k = lane; k < total; k += 32) {
        const int y = nTop + k / bw, x = nLeft + k % bw;
        if ((unsigned)((ipx - x) * (ipx - x) + (ipy - y) * (ipy - y)) > nRange * nRange) continue;
        n_valid++;
        const int s = zmssd_at(T, trow, tsum, tsq, x, y, max_ssd);
        if (s < best_ssd || (s == best_ssd && k < best_idx)) { best_ssd = s; best_idx = k; best_x = x; best_y = y; }
      }
    } else {
      const int i0 = T.row_lut[nTop];
      const int i1 = nBottomPlusOne >= T.h ? n_corners : min(T.row_lut[nBottomPlusOne], n_corners);
      for (int k = i0 + lane; k < i1; k += 32) {
        const int2 c = T.corners[k];
        if (c.x < nLeft || c.x > nRight) continue;
        if ((unsigned)((ipx - c.x) * (ipx - c.x) + (ipy - c.y) * (ipy - c.y)) > nRange * nRange) continue;
        n_valid++;
        const int s = zmssd_at(T, trow, tsum, tsq, c.x, c.y, max_ssd);
        if (s < best_ssd || (s == best_ssd && k < best_idx)) { best_ssd = s; best_idx = k; best_x = c.x; best_y = c.y; }
      }
    }
    // first-best in list order: lexicographic (ssd, index) minimum; ssd == max_ssd+1 with no index never wins
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int os = __shfl_xor_sync(0xffffffffu, best_ssd, o), oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
      const int ox = __shfl_xor_sync(0xffffffffu, best_x, o), oy = __shfl_xor_sync(0xffffffffu, best_y, o);
      if (os < best_ssd || (os == best_ssd && oi < best_idx)) { best_ssd = os; best_idx = oi; best_x = ox; best_y = oy; }
      n_valid += __shfl_xor_sync(0xffffffffu, n_valid, o);
    }
  }
  out.score = best_ssd;
  out.n_candidates = n_valid;
  const bool found = !early && best_ssd < max_ssd;
  if (!found) { if (early) out.score = max_ssd + 1; if (lane == 0) res[i] = out; return; }
  out.found = 1;
  out.coarse_x = best_x; out.coarse_y = best_y;
  // mv2CoarsePos = LevelZeroPos(irBest, level)   (include/mcptam/LevelHelpers.h:61-64)
  double posx = __dsub_rn(__dmul_rn(__dadd_rn((double)best_x, 0.5), (double)lsc), 0.5);
  double posy = __dsub_rn(__dmul_rn(__dadd_rn((double)best_y, 0.5), (double)lsc), 0.5);
  out.found_x = posx; out.found_y = posy;
  if (rq.subpix_its <= 0) { if (lane == 0) res[i] = out; return; }
  // ---- MakeSubPixTemplate + IterateSubPixToConvergence (mixed fp32/fp64, no contraction) ----------------
  out.did_subpix = 1;
  // lanes 0..31 and +32: inner pixel k = y*6+x (y,x in 1..6) -> jacobians
  float jxv[2] = { 0.f, 0.f }, jyv[2] = { 0.f, 0.f };
  int tv[2] = { 0, 0 };
  double H[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
#pragma unroll
  for (int h2 = 0; h2 < 2; h2++) {
    const int k = lane + 32 * h2;
    if (k < 36) {
      const int yy = 1 + k / 6, xx = 1 + k % 6;
      const uint8_t* t = s_t[wid];
      const double gx = __dmul_rn(0.5, (double)((int)t[yy * 8 + xx + 1] - (int)t[yy * 8 + xx - 1]));
      const double gy = __dmul_rn(0.5, (double)((int)t[(yy + 1) * 8 + xx] - (int)t[(yy - 1) * 8 + xx]));
      jxv[h2] = (float)gx; jyv[h2] = (float)gy; tv[h2] = t[yy * 8 + xx];
      const double g[3] = { gx, gy, 1.0 };
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) H[r * 3 + c] += g[r] * g[c];   // exact (multiples of 1/4)
    }
  }
#pragma unroll
  for (int q = 0; q < 9; q++) { for (int o = 16; o > 0; o >>= 1) H[q] += __shfl_xor_sync(0xffffffffu, H[q], o); }
  double Hinv[9];
  toon_chol3_inverse(H, Hinv);
  double mean_diff = 0.0;
  int converged = 0;
  for (int it = 0; it < rq.subpix_its; it++) {
    const double cxl = __dsub_rn(__ddiv_rn(__dadd_rn(posx, 0.5), (double)lsc), 0.5);    // LevelNPos
    const double cyl = __dsub_rn(__ddiv_rn(__dadd_rn(posy, 0.5), (double)lsc), 0.5);
    const int rx = (int)(cxl > 0.0 ? __dadd_rn(cxl, 0.5) : __dsub_rn(cxl, 0.5));
    const int ry = (int)(cyl > 0.0 ? __dadd_rn(cyl, 0.5) : __dsub_rn(cyl, 0.5));
    if (!(rx >= 5 && ry >= 5 && rx < T.w - 5 && ry < T.h - 5)) { converged = 0; break; }
    const double bx = __dsub_rn(cxl, 4.0), by = __dsub_rn(cyl, 4.0);
    const double dX = __dsub_rn(bx, floor(bx)), dY = __dsub_rn(by, floor(by));
    const float fMixTL = (float)__dmul_rn(__dsub_rn(1.0, dX), __dsub_rn(1.0, dY));
    const float fMixTR = (float)__dmul_rn(dX, __dsub_rn(1.0, dY));
    const float fMixBL = (float)__dmul_rn(__dsub_rn(1.0, dX), dY);
    const float fMixBR = (float)__dmul_rn(dX, dY);
    const int ibx = (int)bx, iby = (int)by;
    double dd[2] = { 0, 0 };
#pragma unroll
    for (int h2 = 0; h2 < 2; h2++) {
      const int k = lane + 32 * h2;
      if (k < 36) {
        const int yy = 1 + k / 6, xx = 1 + k % 6;
        const uint8_t* tl = T.img + (size_t)(iby + yy) * T.pitch + ibx + xx;
        const float fPixel = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(fMixTL, (float)tl[0]), __fmul_rn(fMixTR, (float)tl[1])),
                                                   __fmul_rn(fMixBL, (float)tl[T.pitch])), __fmul_rn(fMixBR, (float)tl[T.pitch + 1]));
        dd[h2] = __dadd_rn((double)__fsub_rn(fPixel, (float)tv[h2]), mean_diff);
      }
    }
    // sequential accumulation in the reference's pixel order (y outer, x inner)
    double acc0 = 0, acc1 = 0, acc2 = 0;
    for (int k = 0; k < 36; k++) {
      const double dk = __shfl_sync(0xffffffffu, dd[k >> 5], k & 31);
      const float jxk = __shfl_sync(0xffffffffu, jxv[k >> 5], k & 31), jyk = __shfl_sync(0xffffffffu, jyv[k >> 5], k & 31);
      acc0 = __dadd_rn(acc0, __dmul_rn(dk, (double)jxk));
      acc1 = __dadd_rn(acc1, __dmul_rn(dk, (double)jyk));
      acc2 = __dadd_rn(acc2, dk);
    }
    double upd[3];
#pragma unroll
    for (int r = 0; r < 3; r++)
      upd[r] = __dadd_rn(__dadd_rn(__dmul_rn(Hinv[r * 3], acc0), __dmul_rn(Hinv[r * 3 + 1], acc1)), __dmul_rn(Hinv[r * 3 + 2], acc2));
    posx = __dsub_rn(posx, __dmul_rn(upd[0], (double)lsc));
    posy = __dsub_rn(posy, __dmul_rn(upd[1], (double)lsc));
    mean_diff = __dsub_rn(mean_diff, upd[2]);
    const double u2 = __dadd_rn(__dmul_rn(upd[0], upd[0]), __dmul_rn(upd[1], upd[1]));
    if (u2 < 0.03 * 0.03) { converged = 1; break; }
  }
  if (!converged) out.found = 0;                                   // src/Tracker.cc:1355-1362
  else { out.found_x = posx; out.found_y = posy; }
  if (lane == 0) res[i] = out;
}

// ---------------------------------------------------------------------------------------------
// k_patch_search_tma: the same search, B200 data path.
//   * The part of the target level a patch can touch -- the search disc plus the 8x8 patch plus the sub-pixel drift -- is ONE
//     80 x 48 byte box (x origin rounded down to 16 bytes).  Lane 0 fetches it with a 2-D TMA tile copy (cp.async.bulk.tensor.2d, SASS UTMALDG; pixels outside the
//     image arrive as zeros) into the warp's shared-memory window and the warp waits on an mbarrier.  Candidate scoring and all
//     sub-pixel iterations then read shared memory: no per-candidate unaligned global loads, no global round trip per iteration.
//   * Candidates are first compacted (ballot) and then scored FOUR AT A TIME, eight lanes per candidate, one template row per
//     lane -- three or thirty candidates keep the warp equally busy.
//   Arithmetic, candidate order and tie-breaks are those of k_patch_search (bit-identical results); a request whose footprint
//   does not fit the box (search range > 17 level pixels) is handled by that kernel's global-memory path.
// ---------------------------------------------------------------------------------------------
constexpr int PSW_W = 80, PSW_H = 48;        // the box: its x origin must be 16-byte aligned (TMA tiled mode), hence 80 = 2 * 24 + 8 + 16 + 8 wide

__device__ __forceinline__ unsigned ps_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

struct PsMaps { CUtensorMap lv[MCP_LEVELS]; };       // the target slot's four level descriptors, passed as a __grid_constant__ parameter
struct PsWin { const uint8_t* win; int x0, y0; const uint8_t* img; int pitch; };
__device__ __forceinline__ int ps_pix(const PsWin& W, int x, int y)
{
  const int wx = x - W.x0, wy = y - W.y0;
  if ((unsigned)wx < (unsigned)PSW_W && (unsigned)wy < (unsigned)PSW_H) return W.win[wy * PSW_W + wx];
  return W.img[(size_t)y * W.pitch + x];                     // (a sub-pixel walk that left the window: never in practice)
}

__global__ void __launch_bounds__(128) k_patch_search_tma(FeDev fe, const __grid_constant__ PsMaps tmaps, int target_slot, int n,
                                                         const McpPatchReq* __restrict__ req, McpPatchRes* __restrict__ res, uint8_t* __restrict__ templ_out)
{
  __shared__ __align__(128) uint8_t s_win[4][PSW_W * PSW_H];
  __shared__ __align__(8) unsigned long long s_bar[4];
  __shared__ __align__(8) uint8_t s_t[4][64];
  __shared__ double s_terms[4][3][36];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int i = blockIdx.x * (blockDim.x >> 5) + wid;
  if (i >= n) return;
  const McpPatchReq rq = req[i];
  McpPatchRes out;
  out.template_bad = 1; out.found = 0; out.did_subpix = 0; out.score = 0; out.coarse_x = 0; out.coarse_y = 0;
  out.found_x = 0; out.found_y = 0; out.n_candidates = 0; out.pad_ = 0;
  const int max_ssd = 8 * 8 * 250;                                  // src/PatchFinder.cc:44,61
  const bool args_ok = rq.src_kf >= 0 && rq.src_kf < fe.n_slots && rq.src_level >= 0 && rq.src_level < MCP_LEVELS &&
                       rq.search_level >= 0 && rq.search_level < MCP_LEVELS;
  if (!args_ok) { if (lane == 0) res[i] = out; return; }
  // ---- the window of the target level: issued first, it lands while the template is being warped --------------------------
  const FeLevel T = fe.kf[target_slot].lv[rq.search_level];
  const int lsc = 1 << rq.search_level;
  const int ipx = rq.exhaustive == 2 ? rq.pred_x : rq.pred_x / lsc, ipy = rq.exhaustive == 2 ? rq.pred_y : rq.pred_y / lsc;
  const unsigned nRange = ((unsigned)rq.range + lsc - 1) / lsc;
  PsWin W;
  W.win = s_win[wid]; W.x0 = ((ipx - 24) >> 4) << 4; W.y0 = ipy - PSW_H / 2; W.img = T.img; W.pitch = T.pitch;   // covers x in [ipx - 24, ipx + 24]
  if (lane == 0) {
    const unsigned bar = ps_smem_u32(&s_bar[wid]);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(PSW_W * PSW_H) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(ps_smem_u32(s_win[wid])), "l"(&tmaps.lv[rq.search_level]), "r"(W.x0), "r"(W.y0), "r"(bar) : "memory");
  }
  __syncwarp();
  // ---- FindPatchCoarse bookkeeping: the corner-list range of the disc's rows and its first 32 corners are requested now
  // (four dependent global round trips) and land while the template is being warped -------------------------------------------
  const int n_corners = min(fe.kf[target_slot].meta->lv[rq.search_level].n_corners, fe.kf[target_slot].corner_cap);
  int nTop = ipy - (int)nRange, nBottomPlusOne = ipy + (int)nRange + 1, nLeft = ipx - (int)nRange;
  const int nRight = ipx + (int)nRange;
  bool early = false;
  if (nTop < 0) nTop = 0;
  if (nTop >= T.h) early = true;
  if (nBottomPlusOne <= 0) early = true;
  if (nLeft < 0) nLeft = 0;
  if (nLeft >= T.w) early = true;
  const bool list_search = !early && rq.exhaustive == 0;
  const int i0 = list_search ? T.row_lut[nTop] : 0;
  const int i1 = list_search ? (nBottomPlusOne >= T.h ? n_corners : min(T.row_lut[nBottomPlusOne], n_corners)) : 0;
  int2 c_next = (i0 + lane < i1) ? T.corners[i0 + lane] : make_int2(0, 0);
  // ---- template: MakeTemplateCoarseCont (unchanged arithmetic) -------------------------------------------------------------
  const FeLevel S = fe.kf[rq.src_kf].lv[rq.src_level];
  const double wi0 = rq.warp_inv[0], wi1 = rq.warp_inv[1], wi2 = rq.warp_inv[2], wi3 = rq.warp_inv[3];
  const double det = __dsub_rn(__dmul_rn(wi0, wi3), __dmul_rn(wi1, wi2));
  const double idet = __ddiv_rn(1.0, det);
  const double ls = (double)lsc;
  const double M0 = __dmul_rn(__dmul_rn(wi3, idet), ls), M1 = __dmul_rn(__dmul_rn(-wi1, idet), ls);
  const double M2 = __dmul_rn(__dmul_rn(-wi2, idet), ls), M3 = __dmul_rn(__dmul_rn(wi0, idet), ls);
  const double ax = M0, ay = M2, dx = M1, dy = M3;
  const double p0x = __dsub_rn((double)rq.src_cx, __dadd_rn(__dmul_rn(M0, 4.0), __dmul_rn(M1, 4.0)));
  const double p0y = __dsub_rn((double)rq.src_cy, __dadd_rn(__dmul_rn(M2, 4.0), __dmul_rn(M3, 4.0)));
  double min_x = p0x, min_y = p0y, max_x = p0x, max_y = p0y;
  if (ax < 0) min_x = __dadd_rn(min_x, __dmul_rn(8.0, ax)); else max_x = __dadd_rn(max_x, __dmul_rn(8.0, ax));
  if (dx < 0) min_x = __dadd_rn(min_x, __dmul_rn(8.0, dx)); else max_x = __dadd_rn(max_x, __dmul_rn(8.0, dx));
  if (ay < 0) min_y = __dadd_rn(min_y, __dmul_rn(8.0, ay)); else max_y = __dadd_rn(max_y, __dmul_rn(8.0, ay));
  if (dy < 0) min_y = __dadd_rn(min_y, __dmul_rn(8.0, dy)); else max_y = __dadd_rn(max_y, __dmul_rn(8.0, dy));
  const bool all_in = (min_x >= 0 && min_y >= 0 && max_x < S.w - 1 && max_y < S.h - 1);
  const double crx = __dsub_rn(dx, __dmul_rn(8.0, ax)), cry = __dsub_rn(dy, __dmul_rn(8.0, ay));
  const double xb = S.w - 1, yb = S.h - 1;
  double px = p0x, py = p0y;
  int n_out = 0;
  uint8_t mine[2] = { 0, 0 };
  // every lane replays the whole position sequence (the additions are order dependent) and keeps the positions of its two
  // template pixels; the bilinear samples are then taken by all lanes at once (two global round trips per warp, not 64)
  double qx[2] = { 0, 0 }, qy[2] = { 0, 0 };
#pragma unroll
  for (int ii = 0; ii < 8; ++ii) {
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int k = ii * 8 + jj;
      const bool me = (k & 31) == lane;
      qx[k >> 5] = me ? px : qx[k >> 5]; qy[k >> 5] = me ? py : qy[k >> 5];
      px = __dadd_rn(px, ax); py = __dadd_rn(py, ay);
    }
    px = __dadd_rn(px, crx); py = __dadd_rn(py, cry);
  }
#pragma unroll
  for (int h2 = 0; h2 < 2; h2++) {
    uint8_t v = 0;
    if (all_in || (0 <= qx[h2] && 0 <= qy[h2] && qx[h2] < xb && qy[h2] < yb)) v = sample_u8(S, qx[h2], qy[h2], fe.transform_round);
    else n_out++;
    mine[h2] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_out += __shfl_xor_sync(0xffffffffu, n_out, o);
  s_t[wid][lane] = mine[0]; s_t[wid][lane + 32] = mine[1];
  __syncwarp();
  if (templ_out) { templ_out[(size_t)i * 64 + lane] = mine[0]; templ_out[(size_t)i * 64 + 32 + lane] = mine[1]; }
  const bool det_ok = isfinite(idet);
  out.template_bad = (n_out > 0 || !det_ok) ? 1 : 0;
  // (every lane must consume the TMA completion before the warp may leave: the copy targets this CTA's shared memory)
  {
    const unsigned bar = ps_smem_u32(&s_bar[wid]);
    unsigned done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar) : "memory");
  }
  if (out.template_bad) { if (lane == 0) res[i] = out; return; }
  int tsum = mine[0] + mine[1], tsq = mine[0] * mine[0] + mine[1] * mine[1];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { tsum += __shfl_xor_sync(0xffffffffu, tsum, o); tsq += __shfl_xor_sync(0xffffffffu, tsq, o); }

  // ---- FindPatchCoarse -----------------------------------------------------------------------------------------------------
  int best_ssd = max_ssd + 1, best_idx = 0x7fffffff, best_x = 0, best_y = 0, n_valid = 0;
  const int sub = lane & 7, grp = lane >> 3;
  // eight lanes score one candidate (lane `sub` = template row), four candidates per pass; k = position in the reference's visiting order
  auto score4 = [&](int cx, int cy, int k, bool valid) {
    int isum = 0, isq = 0, cross = 0;
    const bool inb = valid && (cx >= 4 && cy >= 4 && cx < T.w - 4 && cy < T.h - 4);
    if (inb) {
      const uint8_t* trow = s_t[wid] + 8 * sub;
#pragma unroll
      for (int c = 0; c < 8; c++) {
        const int nn = ps_pix(W, cx - 4 + c, cy - 4 + sub), t = trow[c];
        isum += nn; isq += nn * nn; cross += nn * t;
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) { isum += __shfl_xor_sync(0xffffffffu, isum, o); isq += __shfl_xor_sync(0xffffffffu, isq, o); cross += __shfl_xor_sync(0xffffffffu, cross, o); }
    if (valid && sub == 0) {
      const int SA = tsum, SB = isum;
      const int sc = inb ? ((2 * SA * SB - SA * SA - SB * SB) / 64 + isq + tsq - 2 * cross) : max_ssd + 1;
      if (sc < best_ssd || (sc == best_ssd && k < best_idx)) { best_ssd = sc; best_idx = k; best_x = cx; best_y = cy; }
    }
  };
  if (rq.exhaustive == 2) {
    early = false; best_ssd = 0; best_x = rq.pred_x; best_y = rq.pred_y;
  } else if (!early) {
    if (rq.exhaustive) {
      const int y_end = min(nBottomPlusOne, T.h), x_end = min(nRight + 1, T.w);
      const int bw = x_end - nLeft, bh = y_end - nTop;
      const int total = (bw > 0 && bh > 0) ? bw * bh : 0;
      for (int k0 = 0; k0 < total; k0 += 4) {
        const int k = k0 + grp;
        const int y = nTop + k / max(bw, 1), x = nLeft + k % max(bw, 1);
        const bool valid = k < total && !((unsigned)((ipx - x) * (ipx - x) + (ipy - y) * (ipy - y)) > nRange * nRange);
        if (valid && sub == 0) n_valid++;
        score4(x, y, k, valid);
      }
    } else {
      // 32 corners of the list per step (the next 32 are already in flight); the ones inside the disc are scored four at a
      // time (group g takes the g-th of them)
      for (int k0 = i0; k0 < i1; k0 += 32) {
        const int k = k0 + lane;
        const int2 c = c_next;
        if (k + 32 < i1) c_next = T.corners[k + 32];
        bool ok = false;
        if (k < i1) {
          ok = !(c.x < nLeft || c.x > nRight) && !((unsigned)((ipx - c.x) * (ipx - c.x) + (ipy - c.y) * (ipy - c.y)) > nRange * nRange);
        }
        unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) n_valid += __popc(m);
        while (m) {
          const unsigned srcl = __fns(m, 0, grp + 1);               // lane holding this group's candidate, or ~0u
          const bool valid = srcl != 0xffffffffu;
          const int cx = __shfl_sync(0xffffffffu, c.x, valid ? srcl : 0), cy = __shfl_sync(0xffffffffu, c.y, valid ? srcl : 0);
          score4(cx, cy, k0 + (int)srcl, valid);
          for (int q = 0; q < 4 && m; q++) m &= m - 1;              // the four lowest are done
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int os = __shfl_xor_sync(0xffffffffu, best_ssd, o), oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
      const int ox = __shfl_xor_sync(0xffffffffu, best_x, o), oy = __shfl_xor_sync(0xffffffffu, best_y, o);
      if (os < best_ssd || (os == best_ssd && oi < best_idx)) { best_ssd = os; best_idx = oi; best_x = ox; best_y = oy; }
      n_valid += __shfl_xor_sync(0xffffffffu, n_valid, o);
    }
  }
  out.score = best_ssd;
  out.n_candidates = n_valid;
  const bool found = !early && best_ssd < max_ssd;
  if (!found) { if (early) out.score = max_ssd + 1; if (lane == 0) res[i] = out; return; }
  out.found = 1;
  out.coarse_x = best_x; out.coarse_y = best_y;
  double posx = __dsub_rn(__dmul_rn(__dadd_rn((double)best_x, 0.5), (double)lsc), 0.5);
  double posy = __dsub_rn(__dmul_rn(__dadd_rn((double)best_y, 0.5), (double)lsc), 0.5);
  out.found_x = posx; out.found_y = posy;
  if (rq.subpix_its <= 0) { if (lane == 0) res[i] = out; return; }
  // ---- MakeSubPixTemplate + IterateSubPixToConvergence (mixed fp32/fp64, no contraction), pixels from the window -------------
  out.did_subpix = 1;
  float jxv[2] = { 0.f, 0.f }, jyv[2] = { 0.f, 0.f };
  int tv[2] = { 0, 0 };
  double H[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
#pragma unroll
  for (int h2 = 0; h2 < 2; h2++) {
    const int k = lane + 32 * h2;
    if (k < 36) {
      const int yy = 1 + k / 6, xx = 1 + k % 6;
      const uint8_t* t = s_t[wid];
      const double gx = __dmul_rn(0.5, (double)((int)t[yy * 8 + xx + 1] - (int)t[yy * 8 + xx - 1]));
      const double gy = __dmul_rn(0.5, (double)((int)t[(yy + 1) * 8 + xx] - (int)t[(yy - 1) * 8 + xx]));
      jxv[h2] = (float)gx; jyv[h2] = (float)gy; tv[h2] = t[yy * 8 + xx];
      const double g[3] = { gx, gy, 1.0 };
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) H[r * 3 + c] += g[r] * g[c];
    }
  }
#pragma unroll
  for (int q = 0; q < 9; q++) { for (int o = 16; o > 0; o >>= 1) H[q] += __shfl_xor_sync(0xffffffffu, H[q], o); }
  double Hinv[9];
  toon_chol3_inverse(H, Hinv);
  double mean_diff = 0.0;
  int converged = 0;
  const double inv_lsc = 1.0 / (double)lsc;
  for (int it = 0; it < rq.subpix_its; it++) {
    // LevelNPos: the division by the level scale (a power of two) is exact, so it is the multiplication by its reciprocal
    const double cxl = __dsub_rn(__dmul_rn(__dadd_rn(posx, 0.5), inv_lsc), 0.5);
    const double cyl = __dsub_rn(__dmul_rn(__dadd_rn(posy, 0.5), inv_lsc), 0.5);
    const int rx = (int)(cxl > 0.0 ? __dadd_rn(cxl, 0.5) : __dsub_rn(cxl, 0.5));
    const int ry = (int)(cyl > 0.0 ? __dadd_rn(cyl, 0.5) : __dsub_rn(cyl, 0.5));
    if (!(rx >= 5 && ry >= 5 && rx < T.w - 5 && ry < T.h - 5)) { converged = 0; break; }
    const double bx = __dsub_rn(cxl, 4.0), by = __dsub_rn(cyl, 4.0);
    const double dX = __dsub_rn(bx, floor(bx)), dY = __dsub_rn(by, floor(by));
    const float fMixTL = (float)__dmul_rn(__dsub_rn(1.0, dX), __dsub_rn(1.0, dY));
    const float fMixTR = (float)__dmul_rn(dX, __dsub_rn(1.0, dY));
    const float fMixBL = (float)__dmul_rn(__dsub_rn(1.0, dX), dY);
    const float fMixBR = (float)__dmul_rn(dX, dY);
    const int ibx = (int)bx, iby = (int)by;
    double dd[2] = { 0, 0 };
#pragma unroll
    for (int h2 = 0; h2 < 2; h2++) {
      const int k = lane + 32 * h2;
      if (k < 36) {
        const int yy = 1 + k / 6, xx = 1 + k % 6;
        const int X = ibx + xx, Y = iby + yy;
        const float p00 = (float)ps_pix(W, X, Y), p01 = (float)ps_pix(W, X + 1, Y), p10 = (float)ps_pix(W, X, Y + 1), p11 = (float)ps_pix(W, X + 1, Y + 1);
        const float fPixel = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(fMixTL, p00), __fmul_rn(fMixTR, p01)), __fmul_rn(fMixBL, p10)), __fmul_rn(fMixBR, p11));
        dd[h2] = __dadd_rn((double)__fsub_rn(fPixel, (float)tv[h2]), mean_diff);
      }
    }
    // v3Accum += diff * (jx, jy, 1) over the 36 pixels IN ORDER (fp64, order dependent): every lane forms its pixels' three
    // terms, lanes 0..2 each add one component's 36 terms sequentially from shared memory, the sums are broadcast
#pragma unroll
    for (int h2 = 0; h2 < 2; h2++) {
      const int k = lane + 32 * h2;
      if (k < 36) {
        s_terms[wid][0][k] = __dmul_rn(dd[h2], (double)jxv[h2]);
        s_terms[wid][1][k] = __dmul_rn(dd[h2], (double)jyv[h2]);
        s_terms[wid][2][k] = dd[h2];
      }
    }
    __syncwarp();
    double asum = 0;
    if (lane < 3) {
      const double* tr = s_terms[wid][lane];
#pragma unroll
      for (int k = 0; k < 36; k++) asum = __dadd_rn(asum, tr[k]);
    }
    __syncwarp();
    const double acc0 = __shfl_sync(0xffffffffu, asum, 0), acc1 = __shfl_sync(0xffffffffu, asum, 1), acc2 = __shfl_sync(0xffffffffu, asum, 2);
    double upd[3];
#pragma unroll
    for (int r = 0; r < 3; r++)
      upd[r] = __dadd_rn(__dadd_rn(__dmul_rn(Hinv[r * 3], acc0), __dmul_rn(Hinv[r * 3 + 1], acc1)), __dmul_rn(Hinv[r * 3 + 2], acc2));
    posx = __dsub_rn(posx, __dmul_rn(upd[0], (double)lsc));
    posy = __dsub_rn(posy, __dmul_rn(upd[1], (double)lsc));
    mean_diff = __dsub_rn(mean_diff, upd[2]);
    const double u2 = __dadd_rn(__dmul_rn(upd[0], upd[0]), __dmul_rn(upd[1], upd[1]));
    if (u2 < 0.03 * 0.03) { converged = 1; break; }
  }
  if (!converged) out.found = 0;
  else { out.found_x = posx; out.found_y = posy; }
  if (lane == 0) res[i] = out;
}

// FindShiTomasiScoreAtPoint (src/ShiTomasi.cc:34-63, half box 3)
__device__ __forceinline__ double shitomasi_at(const FeLevel& L, int cx, int cy)
{
  const int hb = 3;
  if (!(cx >= hb + 1 && cy >= hb + 1 && cx < L.w - hb - 1 && cy < L.h - hb - 1)) return 0.0;
  double dXX = 0, dYY = 0, dXY = 0;
  for (int y = cy - hb; y <= cy + hb; y++)
    for (int x = cx - hb; x <= cx + hb; x++) {
      const double dx = (int)L.img[(size_t)y * L.pitch + x + 1] - (int)L.img[(size_t)y * L.pitch + x - 1];
      const double dy = (int)L.img[(size_t)(y + 1) * L.pitch + x] - (int)L.img[(size_t)(y - 1) * L.pitch + x];
      dXX += dx * dx; dYY += dy * dy; dXY += dx * dy;             // exact integers
    }
  const int nPixels = (2 * hb + 1) * (2 * hb + 1);
  dXX = __ddiv_rn(dXX, __dmul_rn(2.0, (double)nPixels));
  dYY = __ddiv_rn(dYY, __dmul_rn(2.0, (double)nPixels));
  dXY = __ddiv_rn(dXY, __dmul_rn(2.0, (double)nPixels));
  const double tr = __dadd_rn(dXX, dYY);
  const double disc = __dsub_rn(__dmul_rn(tr, tr), __dmul_rn(4.0, __dsub_rn(__dmul_rn(dXX, dYY), __dmul_rn(dXY, dXY))));
  return __dmul_rn(0.5, __dsub_rn(tr, sqrt(disc)));
}

// one thread per corner
__global__ void k_shitomasi(FeLevel L, int n, const int2* __restrict__ xy, double* __restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = shitomasi_at(L, xy[i].x, xy[i].y);
}

// MiniPatch: one warp per query.  9x9 SSD, first-best over corners in the +-range box (src/MiniPatch.cc:61-113)
__global__ void __launch_bounds__(128) k_minipatch(FeLevel S, FeLevel T, int n_corners, int n, const int2* __restrict__ src_xy,
                                                  const int2* __restrict__ start_xy, int range, int2* __restrict__ pos_out,
                                                  int* __restrict__ found, const int* __restrict__ n_dev)
{
  __shared__ uint8_t s_p[4][81 + 3];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int i = blockIdx.x * (blockDim.x >> 5) + wid;
  if (n_dev) n = min(n, *n_dev);          // the query count lives on the device (MakeKeyFrame_Rest)
  if (i >= n) return;
  const int2 sp = src_xy[i], st = start_xy[i];
  const bool ok = sp.x >= 4 && sp.y >= 4 && sp.x < S.w - 4 && sp.y < S.h - 4;      // assert in SampleFromImage
  for (int k = lane; k < 81; k += 32) s_p[wid][k] = ok ? S.img[(size_t)(sp.y - 4 + k / 9) * S.pitch + sp.x - 4 + k % 9] : 0;
  __syncwarp();
  const int max_ssd = 9999;
  int best = max_ssd + 1, best_idx = 0x7fffffff, bx = 0, by = 0;
  const int tlx = st.x - range, tly = st.y - range, brx = st.x + range, bry = st.y + range;
  int top = tly;
  if (top < 0) top = 0;
  if (top >= T.h) top = T.h - 1;
  const int i0 = T.row_lut[top];
  const int bot = bry + 1;
  const int i1 = bot >= T.h ? n_corners : (bot < 0 ? 0 : min(T.row_lut[bot], n_corners));
  if (ok)
    for (int k = i0 + lane; k < i1; k += 32) {
      const int2 c = T.corners[k];
      if (c.x < tlx || c.x > brx) continue;
      int s = max_ssd + 1;
      if (c.x >= 4 && c.y >= 4 && c.x < T.w - 4 && c.y < T.h - 4) {
        s = 0;
        for (int r = 0; r < 9; r++) {
          const uint8_t* ip = T.img + (size_t)(c.y - 4 + r) * T.pitch + c.x - 4;
          for (int cc = 0; cc < 9; cc++) { const int dfr = (int)ip[cc] - (int)s_p[wid][r * 9 + cc]; s += dfr * dfr; }
        }
      }
      if (s < best || (s == best && k < best_idx)) { best = s; best_idx = k; bx = c.x; by = c.y; }
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int os = __shfl_xor_sync(0xffffffffu, best, o), oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
    const int ox = __shfl_xor_sync(0xffffffffu, bx, o), oy = __shfl_xor_sync(0xffffffffu, by, o);
    if (os < best || (os == best && oi < best_idx)) { best = os; best_idx = oi; bx = ox; by = oy; }
  }
  if (lane == 0) {
    const int f = (ok && best < max_ssd) ? 1 : 0;
    found[i] = f;
    pos_out[i] = f ? make_int2(bx, by) : st;
  }
}

// FindPVS building block (src/Tracker.cc:663-723): TrackerData::Project + GetDerivsUnsafe
// (include/mcptam/TrackerData.h:102-129) and PatchFinder::CalcSearchLevelAndWarpMatrix (src/PatchFinder.cc:69-122)
// for every map point against one camera pose.  One thread per point.
__global__ void k_project_points(DevCam cam, Se3 T, int n, const double* __restrict__ pw, const double* __restrict__ rw,
                                 const double* __restrict__ dw, McpProjRes* __restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double p[3] = { pw[3 * i], pw[3 * i + 1], pw[3 * i + 2] };
  const double r[3] = { rw[3 * i], rw[3 * i + 1], rw[3 * i + 2] }, dn[3] = { dw[3 * i], dw[3 * i + 1], dw[3 * i + 2] };
  double vc[3], mr[3], md[3];
  se3_apply(T, p, vc);
  m3_vec(T.R, r, mr);
  m3_vec(T.R, dn, md);
  // projection, pixel derivatives w.r.t. (theta, phi) and the sphere derivatives, kept separate like the reference
  const double n2 = vc[0] * vc[0] + vc[1] * vc[1], norm = sqrt(n2);
  double theta, rho, cphi, sphi;
  if (norm == 0) { theta = 1.5707963267948966; rho = 0; cphi = 0; sphi = 0; }
  else {
    theta = atan(vc[2] / norm);
    const double xs = (theta - cam.theta_mean) / cam.theta_std;
    double val = 0;
    for (int q = cam.n_inv - 1; q > 0; q--) { val += cam.inv[q]; val *= xs; }
    rho = val + cam.inv[0];
    cphi = vc[0] / norm; sphi = vc[1] / norm;
  }
  bool invalid = theta < cam.min_theta;
  const double u = cphi * rho, w = sphi * rho;
  McpProjRes o;
  o.px[0] = cam.affine[0] * u + cam.affine[1] * w + cam.center[0];
  o.px[1] = cam.affine[2] * u + cam.affine[3] * w + cam.center[1];
  if (!(o.px[0] >= 0 && o.px[0] < cam.image_size[0] && o.px[1] >= 0 && o.px[1] < cam.image_size[1])) invalid = true;
  o.in_image = (!invalid && !(o.px[0] < 0 || o.px[1] < 0 || o.px[0] > cam.image_size[0] || o.px[1] > cam.image_size[1])) ? 1 : 0;
  const double wv = polyval5(cam.poly, rho);
  const double drho = (rho * rho + wv * wv) / polyval5(cam.dmod, rho);
  const double dth0 = cphi * drho, dth1 = sphi * drho, dph0 = -sphi * rho, dph1 = cphi * rho;
  o.cam_derivs[0] = cam.affine[0] * dth0 + cam.affine[1] * dth1; o.cam_derivs[2] = cam.affine[2] * dth0 + cam.affine[3] * dth1;
  o.cam_derivs[1] = cam.affine[0] * dph0 + cam.affine[1] * dph1; o.cam_derivs[3] = cam.affine[2] * dph0 + cam.affine[3] * dph1;
  double dth[3], dph[3];
  const double x = vc[0], y = vc[1], z = vc[2], z2 = z * z, nn2 = norm * norm, n3 = nn2 * norm;
  if (norm == 0) { dth[0] = dth[1] = dth[2] = 0; dph[0] = dph[1] = dph[2] = 0; }
  else {
    dth[0] = -z * x / (n3 + norm * z2); dth[1] = -z * y / (n3 + norm * z2); dth[2] = norm / (nn2 + z2);
    dph[0] = -y / (x * x + y * y); dph[1] = x / (x * x + y * y); dph[2] = 0;
  }
  const double r0 = dth[0] * mr[0] + dth[1] * mr[1] + dth[2] * mr[2], r1 = dph[0] * mr[0] + dph[1] * mr[1] + dph[2] * mr[2];
  const double d0 = dth[0] * md[0] + dth[1] * md[1] + dth[2] * md[2], d1 = dph[0] * md[0] + dph[1] * md[1] + dph[2] * md[2];
  o.warp_inv[0] = o.cam_derivs[0] * r0 + o.cam_derivs[1] * r1; o.warp_inv[2] = o.cam_derivs[2] * r0 + o.cam_derivs[3] * r1;
  o.warp_inv[1] = o.cam_derivs[0] * d0 + o.cam_derivs[1] * d1; o.warp_inv[3] = o.cam_derivs[2] * d0 + o.cam_derivs[3] * d1;
  double dDet = o.warp_inv[0] * o.warp_inv[3] - o.warp_inv[1] * o.warp_inv[2];
  int level = 0;
  while (dDet > 3 && level < MCP_LEVELS - 1) { level++; dDet *= 0.25; }
  o.search_level = (dDet > 3 || dDet < 0.5 || !isfinite(dDet)) ? -1 : level;
  o.v3cam[0] = vc[0]; o.v3cam[1] = vc[1]; o.v3cam[2] = vc[2];
  out[i] = o;
}

// ---------------------------------------------------------------------------------------------
// Tracker pose update ("next" row f-1): TrackerData::ProjectAndDerivs + CalcJacobian per point
// (include/mcptam/TrackerData.h:102-178) and Tracker::CalcPoseUpdate (src/Tracker.cc:1386-1511).
// ---------------------------------------------------------------------------------------------
__global__ void k_calc_jacobians(DevCam cam, Se3 B, Se3 Cb, int n, const double* __restrict__ pw, McpJacRes* __restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double p[3] = { pw[3 * i], pw[3 * i + 1], pw[3 * i + 2] };
  double vb[3], vc[3], G[6];
  se3_apply(B, p, vb);
  se3_apply(Cb, vb, vc);
  McpJacRes o;
  const bool invalid = cam_project(cam, vc, o.px, G);       // G = D * [dTheta; dPhi]  (2x3)
  o.in_image = (!invalid && !(o.px[0] < 0 || o.px[1] < 0 || o.px[0] > cam.image_size[0] || o.px[1] > cam.image_size[1])) ? 1 : 0;
  o.pad_ = 0;
  double A[6];
#pragma unroll
  for (int r = 0; r < 2; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) A[r * 3 + c] = G[r * 3] * Cb.R[c] + G[r * 3 + 1] * Cb.R[3 + c] + G[r * 3 + 2] * Cb.R[6 + c];
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const double a0 = A[r * 3], a1 = A[r * 3 + 1], a2 = A[r * 3 + 2];
    o.jac[r * 6 + 0] = a0; o.jac[r * 6 + 1] = a1; o.jac[r * 6 + 2] = a2;
    o.jac[r * 6 + 3] = -a1 * vb[2] + a2 * vb[1];
    o.jac[r * 6 + 4] = a0 * vb[2] - a2 * vb[0];
    o.jac[r * 6 + 5] = -a0 * vb[1] + a1 * vb[0];
  }
  out[i] = o;
}

// one block: errors, exact upper median (radix select in shared/global), M-estimator weights, 6x6 WLS with prior
__global__ void __launch_bounds__(1024) k_pose_update(int n, const McpPoseMeas* __restrict__ meas, int estimator, double override_sigma,
                                                     double* __restrict__ e2buf, McpPoseUpdate* __restrict__ res, int* __restrict__ outlier)
{
  __shared__ unsigned hist[2048];
  __shared__ unsigned long long s_prefix;
  __shared__ unsigned s_rank, s_count;
  __shared__ double acc[27][33];
  __shared__ double s_sig2;
  const int tid = threadIdx.x;
  if (tid == 0) s_count = 0;
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) {
    double e2 = -1.0;
    if (meas[i].found_flag) {
      const double ex = meas[i].sqrt_inv_noise * (meas[i].found[0] - meas[i].image[0]);
      const double ey = meas[i].sqrt_inv_noise * (meas[i].found[1] - meas[i].image[1]);
      e2 = ex * ex + ey * ey;
      atomicAdd(&s_count, 1u);
    }
    e2buf[i] = e2;
    outlier[i] = 0;
  }
  __syncthreads();
  const unsigned nv = s_count;
  if (nv == 0) {
    if (tid == 0) { for (int k = 0; k < 6; k++) res->mu[k] = 0; res->sigma_sq = 0; res->n_inliers = 0; res->n_valid = 0; }
    return;
  }
  if (override_sigma > 0) { if (tid == 0) s_sig2 = override_sigma; }
  else {
    if (tid == 0) { s_prefix = 0ull; s_rank = nv / 2; }
    for (int pass = 0; pass < 6; pass++) {
      const int shift = 63 - 11 * (pass + 1);
      for (int i = tid; i < 2048; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      for (int i = tid; i < n; i += blockDim.x) {
        const double v = e2buf[i];
        if (v < 0) continue;
        const unsigned long long key = (unsigned long long)__double_as_longlong(v);
        unsigned long long hi; unsigned dig;
        if (shift >= 0) { hi = key >> (shift + 11); dig = (unsigned)(key >> shift) & 2047u; }
        else { hi = key >> (11 + shift); dig = (unsigned)(key << (-shift)) & 2047u; }
        if (pass == 0 || hi == prefix) atomicAdd(&hist[dig], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        unsigned r = s_rank, a = 0; int b = 0;
        for (b = 0; b < 2048; b++) { if (a + hist[b] > r) break; a += hist[b]; }
        if (b >= 2048) b = 2047;
        s_rank = r - a;
        s_prefix = (shift >= 0) ? ((prefix << 11) | (unsigned)b) : ((prefix << (11 + shift)) | ((unsigned)b >> (-shift)));
      }
      __syncthreads();
    }
    if (tid == 0) {
      const double med = __longlong_as_double((long long)s_prefix);
      double s = 1.4826 * (1 + 5.0 / (double)((size_t)nv * 2 - 6)) * sqrt(med);
      s = (estimator == 2 ? 1.345 : 4.6851) * s;
      s_sig2 = s * s;
    }
  }
  __syncthreads();
  const double sig2 = s_sig2;
  // 21 upper-triangle entries of C_inv + 6 of the vector, per-thread partials
  double part[27];
#pragma unroll
  for (int k = 0; k < 27; k++) part[k] = 0.0;
  int n_in = 0;
  for (int i = tid; i < n; i += blockDim.x) {
    const double esq = e2buf[i];
    if (esq < 0) continue;
    double w;
    if (estimator == 0) { const double sq = esq > sig2 ? 0.0 : 1.0 - (esq / sig2); w = sq * sq; }
    else if (estimator == 1) w = 1.0 / (1.0 + esq / sig2);
    else w = esq < sig2 ? 1.0 : sqrt(sig2 / esq);
    if (w == 0.0) { outlier[i] = 1; continue; }
    n_in++;
    const double sn = meas[i].sqrt_inv_noise;
    const double ex = sn * (meas[i].found[0] - meas[i].image[0]), ey = sn * (meas[i].found[1] - meas[i].image[1]);
#pragma unroll
    for (int r = 0; r < 2; r++) {
      double J[6];
#pragma unroll
      for (int k = 0; k < 6; k++) J[k] = sn * meas[i].jac[6 * r + k];
      const double er = r == 0 ? ex : ey;
      int q = 0;
#pragma unroll
      for (int a = 0; a < 6; a++) {
#pragma unroll
        for (int b = a; b < 6; b++) part[q++] += (J[a] * w) * J[b];
      }
#pragma unroll
      for (int a = 0; a < 6; a++) part[21 + a] += er * (J[a] * w);
    }
  }
  // block reduction (warp shuffles, then 32 warp partials in shared memory)
  const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int k = 0; k < 27; k++) {
    double v = part[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) acc[k][wid] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_in += __shfl_xor_sync(0xffffffffu, n_in, o);
  __shared__ int s_nin[32];
  if (lane == 0) s_nin[wid] = n_in;
  __syncthreads();
  if (tid == 0) {
    const int nwarps = blockDim.x >> 5;
    double Cinv[36], vec[6];
    int q = 0, total_in = 0;
    for (int w2 = 0; w2 < nwarps; w2++) total_in += s_nin[w2];
    for (int a = 0; a < 6; a++)
      for (int b = a; b < 6; b++, q++) { double s = 0; for (int w2 = 0; w2 < nwarps; w2++) s += acc[q][w2]; Cinv[a * 6 + b] = s; Cinv[b * 6 + a] = s; }
    for (int a = 0; a < 6; a++) { double s = 0; for (int w2 = 0; w2 < nwarps; w2++) s += acc[21 + a][w2]; vec[a] = s; Cinv[a * 6 + a] += 100.0; }   // add_prior(100)
    // TooN::Cholesky<6> (LDL^T) backsub
    double c[36];
    for (int i = 0; i < 36; i++) { c[i] = Cinv[i]; res->c_inv[i] = Cinv[i]; }
    for (int col = 0; col < 6; col++) {
      double inv_diag = 1;
      for (int row = col; row < 6; row++) {
        double val = c[row * 6 + col];
        for (int col2 = 0; col2 < col; col2++) val -= c[col2 * 6 + col] * c[row * 6 + col2];
        if (row == col) { c[row * 6 + col] = val; inv_diag = 1 / val; }
        else { c[col * 6 + row] = val; c[row * 6 + col] = val * inv_diag; }
      }
    }
    double y[6], mu[6];
    for (int i = 0; i < 6; i++) { double val = vec[i]; for (int j = 0; j < i; j++) val -= c[i * 6 + j] * y[j]; y[i] = val; }
    for (int i = 0; i < 6; i++) y[i] /= c[i * 6 + i];
    for (int i = 5; i >= 0; i--) { double val = y[i]; for (int j = i + 1; j < 6; j++) val -= c[j * 6 + i] * mu[j]; mu[i] = val; }
    for (int k = 0; k < 6; k++) res->mu[k] = mu[k];
    res->sigma_sq = sig2; res->n_inliers = total_in; res->n_valid = (int)nv;
  }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// Glare mask (KeyFrame::MakeKeyFrame_Lite with bGlareMasking, src/KeyFrame.cc:214-242):
//   cv::dilate(img, 5x5 MORPH_ELLIPSE, 5 iterations) -> cv::threshold(245, 255, THRESH_BINARY_INV) -> bitwise_and with the
//   internal mask.  Dilation commutes with the threshold, and five passes of the 5x5 ellipse (rows +-2: centre pixel only,
//   rows -1..1: 5 wide) are ONE pass of their Minkowski sum: a 21-row shape of half-width 10 for |dy| <= 5 and
//   10 - 2 (|dy| - 5) beyond.  One CTA per 32x32 tile: the tile plus a 10-pixel apron is thresholded into shared memory as
//   the horizontal distance to the nearest bright pixel of its row, then every output pixel scans its 21 rows.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_glare_mask(FeKf kf, FeMasks out)
{
  const int l = blockIdx.z;
  const FeLevel& L = kf.lv[l];
  const int tiles_x = (L.w + 31) / 32;
  if ((int)blockIdx.x >= tiles_x * ((L.h + 31) / 32)) return;
  const int tx0 = (blockIdx.x % tiles_x) * 32, ty0 = (blockIdx.x / tiles_x) * 32;
  __shared__ unsigned char bright[52][52 + 4];
  __shared__ unsigned char dist[52][32];
  for (int e = threadIdx.x; e < 52 * 52; e += 256) {
    const int yy = e / 52, xx = e % 52, gx = tx0 + xx - 10, gy = ty0 + yy - 10;
    bright[yy][xx] = (gx >= 0 && gy >= 0 && gx < L.w && gy < L.h && L.img[(size_t)gy * L.pitch + gx] > 245) ? 1 : 0;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 52 * 32; e += 256) {
    const int yy = e / 32, x = e % 32;
    int dmin = 11;
    for (int dx = 0; dx <= 10; dx++)
      if (bright[yy][10 + x + dx] || bright[yy][10 + x - dx]) { dmin = dx; break; }
    dist[yy][x] = (unsigned char)dmin;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 32 * 32; e += 256) {
    const int y = e / 32, x = e % 32, gx = tx0 + x, gy = ty0 + y;
    if (gx >= L.w || gy >= L.h) continue;
    bool glare = false;
#pragma unroll
    for (int dy = -10; dy <= 10; dy++) {
      const int ady = dy < 0 ? -dy : dy, hw = ady <= 5 ? 10 : 10 - 2 * (ady - 5);
      glare = glare || (dist[10 + y + dy][x] <= hw);
    }
    const unsigned char internal = L.mask ? L.mask[(size_t)gy * L.pitch + gx] : 255;
    out.m[l][(size_t)gy * L.pitch + gx] = glare ? 0 : internal;
  }
}
void fe_launch_glare_mask(const FeKf& kf, const FeMasks& out, cudaStream_t s)
{
  const int tiles0 = ((kf.lv[0].w + 31) / 32) * ((kf.lv[0].h + 31) / 32);
  k_glare_mask<<<dim3(tiles0, 1, MCP_LEVELS), 256, 0, s>>>(kf, out);
}

void fe_launch_pyramid(const FeKf& kf, int rnd, cudaStream_t s)
{
  const bool fused = (kf.lv[0].w % 8 == 0) && (kf.lv[0].h % 8 == 0) && kf.lv[1].w == kf.lv[0].w / 2 && kf.lv[3].w == kf.lv[0].w / 8;
  if (fused) {
    dim3 grid((kf.lv[3].w + 127) / 128, kf.lv[3].h);
    k_halfsample_fused<<<grid, 128, 0, s>>>(kf, rnd);
  } else {
    for (int l = 1; l < MCP_LEVELS; l++) {
      dim3 blk(32, 8), grid((kf.lv[l].w + 31) / 32, (kf.lv[l].h + 7) / 8);
      k_halfsample<<<grid, blk, 0, s>>>(kf.lv[l - 1].img, kf.lv[l - 1].pitch, kf.lv[l].img, kf.lv[l].pitch, kf.lv[l].w, kf.lv[l].h, rnd);
    }
  }
}
bool fe_fast_fused()
{
  static const bool on = [] { const char* e = getenv("MCP_FE_FAST_FUSED"); return !(e && e[0] == '0'); }();
  return on;
}
// MCP_FE_ZEROCOPY=0: results return through device->host copies instead of kernel stores into pinned host memory
bool fe_zero_copy()
{
  static const bool on = [] { const char* e = getenv("MCP_FE_ZEROCOPY"); return !(e && e[0] == '0'); }();
  return on;
}
int fe_launch_fast(const FeKf& kf, int adaptive, cudaStream_t s)
{
  dim3 blk(FT_W, FT_H);
  if (fe_fast_fused()) {
    const int rows = kf.row_off[MCP_LEVELS];
    launch_chain(k_fast_score_rows, dim3(kf.tile_off[MCP_LEVELS]), blk, 0, s, kf, adaptive);
    launch_chain(k_fast_compact4, dim3((rows + 7) / 8), dim3(256), 0, s, kf, adaptive);
    return 2;
  }
  cudaMemsetAsync(kf.lv[0].hist, 0, sizeof(unsigned) * 32 * MCP_LEVELS, s);
  k_fast_score<<<kf.tile_off[MCP_LEVELS], blk, 0, s>>>(kf);
  const int rows = kf.row_off[MCP_LEVELS];
  k_fast_count<<<(rows + 7) / 8, 256, 0, s>>>(kf, adaptive);
  k_fast_scan<<<MCP_LEVELS, 512, 0, s>>>(kf);
  k_fast_compact<<<(rows + 7) / 8, 256, 0, s>>>(kf, adaptive);
  return 4;
}
void fe_launch_patch_search(const FeDev& fe, const void* tmaps, bool all_fit_window, int target, int n, const McpPatchReq* req, McpPatchRes* res, uint8_t* templ, cudaStream_t s)
{
  if (n <= 0) return;
  // tmaps: one CUtensorMap per (slot, level) in HOST memory or NULL (MCP_FE_TMA=0 / descriptor creation failed); all_fit_window: every request's
  // footprint fits the 64 x 48 TMA box (decided on the host from the ranges)
  if (tmaps && all_fit_window) {
    PsMaps m;
    memcpy(&m, reinterpret_cast<const CUtensorMap*>(tmaps) + (size_t)target * MCP_LEVELS, sizeof(m));     // tmaps: HOST array, [slot][level]
    k_patch_search_tma<<<(n + 3) / 4, 128, 0, s>>>(fe, m, target, n, req, res, templ);
  }
  else k_patch_search<<<(n + 3) / 4, 128, 0, s>>>(fe, target, n, req, res, templ);
}
void fe_launch_project(const DevCam& cam, const Se3& T, int n, const double* pw, const double* rw, const double* dw, McpProjRes* out, cudaStream_t s)
{
  if (n > 0) k_project_points<<<(n + 127) / 128, 128, 0, s>>>(cam, T, n, pw, rw, dw, out);
}
void fe_launch_calc_jacobians(const DevCam& cam, const Se3& B, const Se3& Cb, int n, const double* pw, McpJacRes* out, cudaStream_t s)
{
  if (n > 0) k_calc_jacobians<<<(n + 127) / 128, 128, 0, s>>>(cam, B, Cb, n, pw, out);
}
void fe_launch_pose_update(int n, const McpPoseMeas* meas, int estimator, double override_sigma, double* e2buf, McpPoseUpdate* res, int* outlier, cudaStream_t s)
{
  k_pose_update<<<1, 1024, 0, s>>>(n, meas, estimator, override_sigma, e2buf, res, outlier);
}

// ---------------------------------------------------------------------------------------------
// KeyFrame::MakeKeyFrame_Rest candidate generation (src/KeyFrame.cc:363-531), one pyramid level per launch
// ---------------------------------------------------------------------------------------------
// [3P] libCVD old_style_corner_score: the score fast_nonmax suppresses on
__device__ __forceinline__ int fast_old_score(const FeLevel& L, int x, int y, int b)
{
  const uint8_t* p = L.img + (size_t)y * L.pitch + x;
  const int cb = (int)*p + b, c_b = (int)*p - b;
  int sp = 0, sn = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int v = p[c_ring_dy[i] * L.pitch + c_ring_dx[i]];
    sp += (v > cb) ? v - cb : 0;
    sn += (v < c_b) ? c_b - v : 0;
  }
  return sp > sn ? sp : sn;
}

// is (x, y) in Level::vCorners?  (same predicate as k_fast_compact)
__device__ __forceinline__ bool listed_corner(const FeLevel& L, int x, int y, int thr, bool use_mask)
{
  const int sc = L.score[(size_t)y * L.pitch + x];
  return sc >= thr && sc > 0 && (!use_mask || L.mask[(size_t)y * L.pitch + x] == 255);
}

// fast_nonmax (3x3, old-style score) + in_image_with_border(10) + candidate score.  One thread per listed corner.
// ctr[0] += number of maximal corners kept (= vScoresAndMaxCorners.size()).
__global__ void __launch_bounds__(128) k_rest_flags(FeKf kf, int level, int adaptive, int strict, int use_shi, int* __restrict__ flag,
                                                   double* __restrict__ score, int* __restrict__ ctr)
{
  const FeLevel L = kf.lv[level];
  const int n = min(kf.meta->lv[level].n_corners, kf.corner_cap);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int thr = kf.meta->lv[level].fast_thresh;
  const bool use_mask = adaptive && L.mask;
  const int2 c = L.corners[i];
  const int s0 = fast_old_score(L, c.x, c.y, thr);
  bool is_max = true;
#pragma unroll
  for (int dy = -1; dy <= 1; dy++)
#pragma unroll
    for (int dx = -1; dx <= 1; dx++) {
      if (dx == 0 && dy == 0) continue;
      const int x = c.x + dx, y = c.y + dy;           // corners lie >= 3 px inside the image: always addressable
      if (!listed_corner(L, x, y, thr, use_mask)) continue;
      const int o = fast_old_score(L, x, y, thr);
      if (strict ? (o >= s0) : (o > s0)) is_max = false;
    }
  const bool keep = is_max && c.x >= 10 && c.y >= 10 && c.x < L.w - 10 && c.y < L.h - 10;
  flag[i] = keep ? 1 : 0;
  double sc = 0.0;
  if (keep) sc = use_shi ? shitomasi_at(L, c.x, c.y) : (double)L.score[(size_t)c.y * L.pitch + c.x];   // fast_corner_score_10(.., nFastThresh)
  score[i] = sc;
  const unsigned m = __ballot_sync(__activemask(), keep);
  if (keep && (threadIdx.x & 31) == (__ffs(m) - 1)) atomicAdd(&ctr[0], __popc(m));
}

// Candidate selection by rank: "percent": descending (score, y, x) order, first (int)(n_max * top_fraction) kept
// (std::sort on reverse iterators of pair<double, ImageRef>); "thresh": raster order, score > thresh.
// One thread per corner, all corners streamed through shared memory.  ctr[1] = number of candidates.
__global__ void __launch_bounds__(256) k_rest_select(FeKf kf, int level, int use_thresh, double top_fraction, double thresh,
                                                    const int* __restrict__ flag, const double* __restrict__ score,
                                                    McpCandidate* __restrict__ cand, int2* __restrict__ cand_xy, int* __restrict__ ctr)
{
  __shared__ double s_sc[256];
  __shared__ int2 s_xy[256];
  __shared__ int s_fl[256];
  const FeLevel L = kf.lv[level];
  const int n = min(kf.meta->lv[level].n_corners, kf.corner_cap);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_max = ctr[0];
  const int n_use = use_thresh ? n_max : min((int)(n_max * top_fraction), n_max);
  bool mine = false;
  double sc = 0;
  int2 c = make_int2(0, 0);
  if (i < n) { c = L.corners[i]; sc = score[i]; mine = flag[i] != 0 && (!use_thresh || sc > thresh); }
  int rank = 0;
  for (int base = 0; base < n; base += 256) {
    const int j = base + threadIdx.x;
    if (j < n) { s_sc[threadIdx.x] = score[j]; s_xy[threadIdx.x] = L.corners[j]; s_fl[threadIdx.x] = flag[j]; }
    else s_fl[threadIdx.x] = 0;
    __syncthreads();
    if (mine) {
      const int lim = min(256, n - base);
      for (int k = 0; k < lim; k++) {
        if (!s_fl[k]) continue;
        const double os = s_sc[k];
        const int2 oc = s_xy[k];
        bool before;
        if (use_thresh) before = (os > thresh) && (base + k < i);
        else before = (os > sc) || (os == sc && (oc.y > c.y || (oc.y == c.y && oc.x > c.x)));
        rank += before ? 1 : 0;
      }
    }
    __syncthreads();
  }
  if (mine && rank < n_use) {
    McpCandidate o;
    o.x = c.x; o.y = c.y; o.score = sc;
    cand[rank] = o;
    cand_xy[rank] = c;
    if (use_thresh) atomicAdd(&ctr[1], 1);
  }
  if (!use_thresh && i == 0) ctr[1] = n_use;
}

// stable-point pruning (src/KeyFrame.cc:455-527): keep candidate i if both MiniPatch searches succeeded and the
// round trip ends within sqrt(2) px; ordered compaction by one block.  ctr[2] = number kept.
__global__ void __launch_bounds__(1024) k_rest_prune(const McpCandidate* __restrict__ cand, const int2* __restrict__ back_pos,
                                                    const int* __restrict__ f1, const int* __restrict__ f2, McpCandidate* __restrict__ out,
                                                    int* __restrict__ ctr)
{
  __shared__ int wsum[32];
  __shared__ int carry;
  const int n = ctr[1];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    bool keep = false;
    McpCandidate c;
    if (i < n) {
      c = cand[i];
      const int dx = back_pos[i].x - c.x, dy = back_pos[i].y - c.y;
      keep = f1[i] && f2[i] && (dx * dx + dy * dy <= 2);
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) wsum[wid] = __popc(m);
    __syncthreads();
    int off = carry;
    for (int w = 0; w < wid; w++) off += wsum[w];
    if (keep) out[off + __popc(m & ((1u << lane) - 1))] = c;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = off + __popc(m);
    __syncthreads();
  }
  if (threadIdx.x == 0) ctr[2] = carry;
}

// cfg mirrors McpRestConfig; scratch arrays hold >= corner_cap entries each.  n_hint: upper bound of the corner count.
void fe_launch_rest_level(const FeKf& kf, const FeKf* prev, int level, int n_hint, int n_prev_corners, int adaptive, int strict, int use_shi,
                          int use_thresh, double top_fraction, double thresh, int n_prev, int* flag, double* score, McpCandidate* cand,
                          McpCandidate* cand_out, int2* cand_xy, int2* pos1, int2* pos2, int* f1, int* f2, int* ctr, cudaStream_t s)
{
  if (n_hint <= 0) return;
  k_rest_flags<<<(n_hint + 127) / 128, 128, 0, s>>>(kf, level, adaptive, strict, use_shi, flag, score, ctr);
  k_rest_select<<<(n_hint + 255) / 256, 256, 0, s>>>(kf, level, use_thresh, top_fraction, thresh, flag, score, cand, cand_xy, ctr);
  if (prev && n_prev > 0) {
    // back to the oldest stored frame, then forward to the current one (range 10 per stored frame)
    k_minipatch<<<(n_hint + 3) / 4, 128, 0, s>>>(kf.lv[level], prev->lv[level], n_prev_corners, n_hint, cand_xy, cand_xy, 10 * n_prev, pos1, f1, ctr + 1);
    k_minipatch<<<(n_hint + 3) / 4, 128, 0, s>>>(prev->lv[level], kf.lv[level], n_hint, n_hint, pos1, pos1, 10 * n_prev, pos2, f2, ctr + 1);
    k_rest_prune<<<1, 1024, 0, s>>>(cand, pos2, f1, f2, cand_out, ctr);
  }
}

void fe_launch_shitomasi(const FeLevel& L, int n, const int2* xy, double* out, cudaStream_t s)
{
  if (n > 0) k_shitomasi<<<(n + 127) / 128, 128, 0, s>>>(L, n, xy, out);
}
void fe_launch_minipatch(const FeLevel& S, const FeLevel& T, int n_corners, int n, const int2* src, const int2* start, int range,
                         int2* pos, int* found, cudaStream_t s)
{
  if (n > 0) k_minipatch<<<(n + 3) / 4, 128, 0, s>>>(S, T, n_corners, n, src, start, range, pos, found, nullptr);
}

}  // namespace mcp
