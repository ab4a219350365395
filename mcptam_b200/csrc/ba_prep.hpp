// ba_prep.hpp — pure host code (no CUDA): turns the caller's flat map arrays (the arguments of mcp_ba_load) into the
// device layout of the bundle adjuster: measurements sorted by point, per-point lists of the movable poses that see
// the point ("slots"), the per-point visiting order, and the work lists of k_pose_blocks / k_schur_rows.
//
// This is the marshalling half of BundleAdjusterMulti::BundleAdjust (src/BundleAdjusterMulti.cc:83-203: one
// AddPose per keyframe, one AddPoint per map point, one AddMeas per measurement) and of ChainBundle::Compute's
// initializeOptimization (src/ChainBundle.cc:1293).  The reference rebuilds its graph on every BundleAdjust call, so
// this runs once per call and is part of the end-to-end time: every pass is linear (counting sorts, no comparison
// sorts) and all outputs land in caller-provided (pinned) storage that is pooled across calls.
//
// Header-only so that the CPU tests and tools/prep_bench.cpp can compile it without nvcc.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace mcp {

// layout-compatible with CUDA's int4 / int2 / double2
struct alignas(16) PInt4 { int x, y, z, w; };
struct alignas(8) PInt2 { int x, y; };
struct alignas(16) PDouble2 { double x, y; };

struct PrepAlloc {
  void* (*alloc)(size_t bytes);
  void (*release)(void* p);
};

// grow-only array in allocator-provided storage; contents are NOT preserved by resize (every load refills)
template <class T> struct PrepArr {
  T* p = nullptr;
  size_t n = 0, cap = 0;
  bool resize(size_t m, const PrepAlloc& a)
  {
    if (m > cap) {
      if (p) a.release(p);
      cap = m + m / 4 + 64;
      p = static_cast<T*>(a.alloc(cap * sizeof(T)));
      if (!p) { cap = 0; n = 0; return false; }
    }
    n = m;
    return true;
  }
  void free_all(const PrepAlloc& a) { if (p) a.release(p); p = nullptr; n = cap = 0; }
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
  size_t bytes() const { return n * sizeof(T); }
};

constexpr int PREP_PB_CHUNK = 128;        // measurements per k_pose_blocks work item
constexpr int PREP_RS_BYTES = 6144;       // staging bytes per buffer of k_schur_rows (= RS_BYTES)
constexpr int PREP_RS_MAXE = 8;           // entries per group of k_schur_rows (= RS_MAXE)

enum { PREP_OK = 0, PREP_INVALID = -1, PREP_UNSUPPORTED = -2, PREP_NOMEM = -3 };

struct BaPrep {
  // outputs that travel to the device
  PrepArr<int> pose_var;          // [n_pose] index among the movable poses, -1 if fixed
  PrepArr<int> pt_var;            // [n_pt]
  PrepArr<PInt4> pt_info;         // [n_pt] {source pose, second chain link, source pose variable, slot of the source pose}
  PrepArr<int> pt_order;          // [n_pt] visiting order: inside every rank's range heaviest points first (stable)
  PrepArr<int> pt_meas_off;       // [n_pt+1]
  PrepArr<int> pt_slot_off;       // [n_pt+1]
  PrepArr<int> slot_var, slot_pt; // [n_slots] ascending pose variable inside a point
  PrepArr<PDouble2> meas_xy;      // [n_meas] sorted by point
  PrepArr<double> meas_info;      // [n_meas] 1/sqrt(noise)  (src/ChainBundle.cc:1244-1245)
  PrepArr<PInt4> meas_a;          // {observer pose, second link, camera, original index}
  PrepArr<PInt4> meas_b;          // {observer variable, observer slot, has source Jacobian, point}
  PrepArr<int> pb_idx;            // measurement positions bucketed by pose block
  PrepArr<PInt4> pb_items;        // {block row, block col, begin, end} into pb_idx
  PrepArr<PInt2> rs_ent;          // row-wise Schur lists (only when want_rows)
  PrepArr<int> rs_grp;
  PrepArr<PInt4> rs_items;
  // host-side results
  std::vector<int> meas_orig;     // sorted position -> original measurement index
  std::vector<int> part_pt, part_meas;   // world+1 boundaries of the contiguous point partition
  int npv = 0, nptv = 0, n_slots = 0, max_slots = 1, rs_nblk = 1;
  long long n_inc = 0;            // co-visibility incidences of this rank: sum over its points of K(K+1)/2
  char err[256] = "";
  // scratch, kept between calls
  std::vector<int> cursor, stamp, pos, tmp, key_cnt, order;
  std::vector<long long> keys;

  void free_all(const PrepAlloc& a)
  {
    pose_var.free_all(a); pt_var.free_all(a); pt_info.free_all(a); pt_order.free_all(a); pt_meas_off.free_all(a);
    pt_slot_off.free_all(a); slot_var.free_all(a); slot_pt.free_all(a); meas_xy.free_all(a); meas_info.free_all(a);
    meas_a.free_all(a); meas_b.free_all(a); pb_idx.free_all(a); pb_items.free_all(a); rs_ent.free_all(a);
    rs_grp.free_all(a); rs_items.free_all(a);
  }
};

// contiguous, measurement-count-balanced partition of points
inline void partition_points(const int* pt_meas_off, int n_pt, int world, int* part_pt)
{
  const long long total = pt_meas_off[n_pt] + (long long)n_pt * 4;   // weight: measurements + per-point overhead
  int p = 0;
  part_pt[0] = 0;
  for (int r = 1; r < world; r++) {
    const long long target = total * r / world;
    while (p < n_pt && (long long)pt_meas_off[p] + (long long)p * 4 < target) p++;
    part_pt[r] = p;
  }
  part_pt[world] = n_pt;
}

#define MCP_PREP_FAIL(code, ...) do { snprintf(o.err, sizeof(o.err), __VA_ARGS__); return code; } while (0)

inline int ba_prepare(BaPrep& o, const PrepAlloc& al, int n_cam, int n_pose, const uint8_t* pose_fixed, int n_pt,
                      const int32_t* pt_chain, const uint8_t* pt_fixed, int n_meas, const double* meas_xy,
                      const int32_t* meas_chain, const int32_t* meas_pt, const double* meas_noise, const int32_t* meas_cam,
                      int rank, int world, bool want_rows)
{
  o.err[0] = 0;
  const size_t np1 = (size_t)std::max(n_pt, 1), nm1 = (size_t)std::max(n_meas, 1);
  if (!o.pose_var.resize((size_t)n_pose, al) || !o.pt_var.resize(np1, al) || !o.pt_info.resize(np1, al) ||
      !o.pt_order.resize(np1, al) || !o.pt_meas_off.resize((size_t)n_pt + 1, al) || !o.pt_slot_off.resize((size_t)n_pt + 1, al) ||
      !o.slot_var.resize((size_t)n_meas + n_pt + 1, al) || !o.slot_pt.resize((size_t)n_meas + n_pt + 1, al) ||
      !o.meas_xy.resize(nm1, al) || !o.meas_info.resize(nm1, al) || !o.meas_a.resize(nm1, al) || !o.meas_b.resize(nm1, al))
    MCP_PREP_FAIL(PREP_NOMEM, "mcp_ba_load: host staging allocation failed");

  int npv = 0;
  for (int i = 0; i < n_pose; i++) o.pose_var[i] = pose_fixed[i] ? -1 : npv++;
  o.npv = npv;
  for (int p = 0; p < n_pt; p++) {
    const int a = pt_chain[2 * p], b = pt_chain[2 * p + 1];
    if (a < 0 || a >= n_pose || b >= n_pose) MCP_PREP_FAIL(PREP_INVALID, "point %d: chain index out of range", p);
    if (b >= 0 && !pose_fixed[b]) MCP_PREP_FAIL(PREP_UNSUPPORTED, "point %d: movable second chain link is not supported", p);
  }
  // validation + histogram of measurements per point in one pass
  int* pmo = o.pt_meas_off.p;
  std::fill(pmo, pmo + n_pt + 1, 0);
  for (int m = 0; m < n_meas; m++) {
    const int a = meas_chain[2 * m], b = meas_chain[2 * m + 1];
    if (a < 0 || a >= n_pose || b >= n_pose) MCP_PREP_FAIL(PREP_INVALID, "measurement %d: chain index out of range", m);
    if (b >= 0 && !pose_fixed[b]) MCP_PREP_FAIL(PREP_UNSUPPORTED, "measurement %d: movable second chain link is not supported", m);
    if (meas_pt[m] < 0 || meas_pt[m] >= n_pt) MCP_PREP_FAIL(PREP_INVALID, "measurement %d: point index out of range", m);
    if (meas_cam[m] < 0 || meas_cam[m] >= n_cam) MCP_PREP_FAIL(PREP_INVALID, "measurement %d: camera index out of range", m);
    if (!(meas_noise[m] > 0)) MCP_PREP_FAIL(PREP_INVALID, "measurement %d: noise must be > 0", m);
    pmo[meas_pt[m] + 1]++;
  }
  // measurements sorted by point (stable counting sort)
  for (int p = 0; p < n_pt; p++) pmo[p + 1] += pmo[p];
  o.cursor.assign(pmo, pmo + n_pt);
  o.meas_orig.resize((size_t)n_meas);
  for (int m = 0; m < n_meas; m++) o.meas_orig[o.cursor[meas_pt[m]]++] = m;

  int nptv = 0;
  for (int p = 0; p < n_pt; p++) o.pt_var[p] = pt_fixed[p] ? -1 : nptv++;
  o.nptv = nptv;

  // per point: the ascending list of movable poses that carry a Jacobian block of the point (its slots)
  o.stamp.assign((size_t)std::max(npv, 1), -1);
  o.pos.assign((size_t)std::max(npv, 1), 0);
  int n_slots = 0, max_slots = 1;
  std::vector<int>& tmp = o.tmp;
  for (int p = 0; p < n_pt; p++) {
    const int src0 = pt_chain[2 * p], src1 = pt_chain[2 * p + 1];
    const int src_var = o.pose_var[src0];
    const bool movable = o.pt_var[p] >= 0;
    bool any_src = false;
    tmp.clear();
    for (int q = pmo[p]; q < pmo[p + 1]; q++) {
      const int m = o.meas_orig[q];
      const int obs0 = meas_chain[2 * m];
      const bool has_jac = (obs0 != src0);                 // PoseChainHelper::MoveTogether at depth 0
      const int ov = has_jac ? o.pose_var[obs0] : -1;
      const bool has_src = has_jac && src_var >= 0;
      any_src |= has_src;
      if (ov >= 0 && movable && o.stamp[ov] != p) { o.stamp[ov] = p; tmp.push_back(ov); }
      o.meas_a[q] = PInt4{ obs0, meas_chain[2 * m + 1], meas_cam[m], m };
      o.meas_b[q] = PInt4{ ov, -1, has_src ? 1 : 0, p };
      o.meas_xy[q] = PDouble2{ meas_xy[2 * m], meas_xy[2 * m + 1] };
      o.meas_info[q] = 1.0 / std::sqrt(meas_noise[m]);     // src/ChainBundle.cc:1244-1245
    }
    int src_slot = -1;
    o.pt_slot_off[p] = n_slots;
    if (movable) {
      if (any_src && o.stamp[src_var] != p) { o.stamp[src_var] = p; tmp.push_back(src_var); }
      const int K = (int)tmp.size();
      if (K <= 16) {                                       // insertion sort: the lists are short and nearly sorted
        for (int i = 1; i < K; i++) {
          const int v = tmp[i];
          int j = i - 1;
          while (j >= 0 && tmp[j] > v) { tmp[j + 1] = tmp[j]; j--; }
          tmp[j + 1] = v;
        }
      } else {
        std::sort(tmp.begin(), tmp.end());
      }
      for (int i = 0; i < K; i++) { o.slot_var[n_slots + i] = tmp[i]; o.slot_pt[n_slots + i] = p; o.pos[tmp[i]] = i; }
      n_slots += K;
      max_slots = std::max(max_slots, K);
      for (int q = pmo[p]; q < pmo[p + 1]; q++)
        if (o.meas_b[q].x >= 0) o.meas_b[q].y = o.pos[o.meas_b[q].x];
      if (any_src) src_slot = o.pos[src_var];
    }
    o.pt_info[p] = PInt4{ src0, src1, src_var, src_slot };
  }
  o.pt_slot_off[n_pt] = n_slots;
  o.n_slots = n_slots; o.max_slots = max_slots;
  o.slot_var.n = o.slot_pt.n = (size_t)std::max(n_slots, 1);
  if (n_slots == 0) { o.slot_var[0] = 0; o.slot_pt[0] = 0; }

  o.part_pt.assign((size_t)world + 1, 0);
  o.part_meas.assign((size_t)world + 1, 0);
  partition_points(pmo, n_pt, world, o.part_pt.data());
  for (int r = 0; r <= world; r++) o.part_meas[r] = pmo[o.part_pt[r]];

  // visiting order of the per-point kernels: inside every rank's range, heaviest points first (stable counting sort
  // by descending measurement count)
  {
    int max_cnt = 0;
    for (int p = 0; p < n_pt; p++) max_cnt = std::max(max_cnt, pmo[p + 1] - pmo[p]);
    std::vector<int>& cnt = o.key_cnt;
    if (n_pt == 0) o.pt_order[0] = 0;
    for (int r = 0; r < world; r++) {
      const int lo = o.part_pt[r], hi = o.part_pt[r + 1];
      cnt.assign((size_t)max_cnt + 2, 0);
      for (int p = lo; p < hi; p++) cnt[(size_t)(max_cnt - (pmo[p + 1] - pmo[p])) + 1]++;
      for (int c = 0; c <= max_cnt; c++) cnt[c + 1] += cnt[c];
      for (int p = lo; p < hi; p++) o.pt_order[lo + cnt[(size_t)(max_cnt - (pmo[p + 1] - pmo[p]))]++] = p;
    }
  }

  // co-visibility incidences of this rank (sizes the pair lists built on the device)
  {
    long long n_inc = 0;
    for (int p = o.part_pt[rank]; p < o.part_pt[rank + 1]; p++) {
      const long long K = o.pt_slot_off[p + 1] - o.pt_slot_off[p];
      n_inc += K * (K + 1) / 2;
    }
    o.n_inc = n_inc;
  }

  // work lists of k_pose_blocks: this rank's measurements bucketed by the pose block they contribute to
  // ((v,v): observed from movable pose v; (lo,hi): observer / source pair), cut into items of <= 128 measurements.
  // Counting sort over the npv^2 block keys; inside a block the measurements keep ascending position.
  {
    const int m_lo = o.part_meas[rank], m_hi = o.part_meas[rank + 1];
    const size_t n_keys = (size_t)std::max(npv, 1) * (size_t)std::max(npv, 1);
    std::vector<int>& cnt = o.key_cnt;
    cnt.assign(n_keys + 1, 0);
    size_t n_ent = 0;
    for (int q = m_lo; q < m_hi; q++) {
      const int vo = o.meas_b[q].x;
      if (vo < 0) continue;
      cnt[(size_t)vo * npv + vo + 1]++;
      n_ent++;
      if (o.meas_b[q].z) {
        const int vs = o.pt_info[o.meas_b[q].w].z;
        cnt[(size_t)std::min(vo, vs) * npv + std::max(vo, vs) + 1]++;
        n_ent++;
      }
    }
    size_t n_items = 0;
    for (size_t k = 0; k < n_keys; k++) { n_items += ((size_t)cnt[k + 1] + PREP_PB_CHUNK - 1) / PREP_PB_CHUNK; cnt[k + 1] += cnt[k]; }
    if (!o.pb_idx.resize(std::max(n_ent, (size_t)1), al) || !o.pb_items.resize(std::max(n_items, (size_t)1), al))
      MCP_PREP_FAIL(PREP_NOMEM, "mcp_ba_load: host staging allocation failed");
    size_t it = 0;
    for (size_t k = 0; k < n_keys; k++) {
      const int b = cnt[k], e = cnt[k + 1];
      const int lo = (int)(k / (size_t)std::max(npv, 1)), hi = (int)(k % (size_t)std::max(npv, 1));
      for (int s = b; s < e; s += PREP_PB_CHUNK) o.pb_items[it++] = PInt4{ lo, hi, s, std::min(s + PREP_PB_CHUNK, e) };
    }
    o.pb_items.n = it;                                     // 0 items is legal (nothing movable is observed)
    for (int q = m_lo; q < m_hi; q++) {
      const int vo = o.meas_b[q].x;
      if (vo < 0) continue;
      o.pb_idx[cnt[(size_t)vo * npv + vo]++] = q;
      if (o.meas_b[q].z) {
        const int vs = o.pt_info[o.meas_b[q].w].z;
        o.pb_idx[cnt[(size_t)std::min(vo, vs) * npv + std::max(vo, vs)]++] = q;
      }
    }
    o.pb_idx.n = n_ent;
  }

  // work lists of k_schur_rows (MCP_BA_SCHUR=0): this rank's (point, slot) entries sorted by pose variable; entry =
  // {slot, number of slots from it to the end of its point}; groups of entries that fit one staging buffer; items =
  // runs of groups of one pose variable sized so that every resident warp gets about one item
  o.rs_nblk = 1;
  o.rs_ent.n = o.rs_grp.n = o.rs_items.n = 0;
  if (want_rows) {
    const int s_lo = o.pt_slot_off[o.part_pt[rank]], s_hi = o.pt_slot_off[o.part_pt[rank + 1]];
    const size_t n_e = (size_t)std::max(s_hi - s_lo, 0);
    std::vector<int>& cnt = o.key_cnt;
    cnt.assign((size_t)std::max(npv, 1) + 1, 0);
    for (int s = s_lo; s < s_hi; s++) cnt[(size_t)o.slot_var[s] + 1]++;
    for (int v = 0; v < npv; v++) cnt[v + 1] += cnt[v];
    o.order.resize(n_e);
    for (int s = s_lo; s < s_hi; s++) o.order[cnt[o.slot_var[s]]++] = s;
    const int target = std::max(16, (int)((n_e + 148 * 8 - 1) / (148 * 8)));     // entries per item
    if (!o.rs_ent.resize(n_e + 1, al) || !o.rs_grp.resize(n_e + 1, al) || !o.rs_items.resize(n_e / 16 + (size_t)npv + 2, al))
      MCP_PREP_FAIL(PREP_NOMEM, "mcp_ba_load: host staging allocation failed");
    size_t ne = 0, ng = 0, ni = 0, i = 0;
    while (i < n_e) {
      size_t j = i;
      const int a = o.slot_var[o.order[i]];
      while (j < n_e && o.slot_var[o.order[j]] == a) j++;
      for (size_t b = i; b < j; b += (size_t)target) {
        const size_t e_end = std::min(b + (size_t)target, j);
        const int g_begin = (int)ng;
        size_t k = b;
        while (k < e_end) {
          int bytes = 0, c = 0;
          const int first = (int)ne;
          while (k < e_end && c < PREP_RS_MAXE) {
            const int sidx = o.order[k];
            const int nb = o.pt_slot_off[o.slot_pt[sidx] + 1] - sidx;
            if (c > 0 && bytes + 192 + 144 * nb > PREP_RS_BYTES) break;
            bytes += 192 + 144 * nb;
            o.rs_ent[ne++] = PInt2{ sidx, nb };
            o.rs_nblk = std::max(o.rs_nblk, o.slot_var[sidx + nb - 1] - a + 1);
            c++; k++;
          }
          o.rs_grp[ng++] = (first << 4) | c;
        }
        o.rs_items[ni++] = PInt4{ a, g_begin, (int)ng, 0 };
      }
      i = j;
    }
    o.rs_ent.n = ne; o.rs_grp.n = ng; o.rs_items.n = ni;
  }
  return PREP_OK;
}

#undef MCP_PREP_FAIL

}  // namespace mcp
