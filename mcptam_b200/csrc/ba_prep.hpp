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
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace mcp {

// A few persistent host threads for the marshalling passes.  run(f) calls f(tid) for tid in [0, size()) -- tid 0 on the
// caller -- and returns when all are done.  Workers spin briefly between the back-to-back passes of one load and sleep
// on a condition variable otherwise.
class HostPool {
 public:
  explicit HostPool(int n_threads) : n_(std::max(n_threads, 1))
  {
    for (int t = 1; t < n_; t++) workers_.emplace_back([this, t] { loop(t); });
  }
  ~HostPool()
  {
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
    cv_.notify_all();
    for (std::thread& w : workers_) w.join();
  }
  HostPool(const HostPool&) = delete;
  HostPool& operator=(const HostPool&) = delete;
  int size() const { return n_; }
  void run(const std::function<void(int)>& f)
  {
    if (n_ == 1) { f(0); return; }
    const unsigned long long next = gen_.load(std::memory_order_relaxed) + 1;
    slot_[next & 1].store(&f, std::memory_order_release);
    remaining_.store(n_ - 1, std::memory_order_relaxed);
    { std::lock_guard<std::mutex> lk(mu_); gen_.store(next, std::memory_order_release); }
    cv_.notify_all();
    f(0);
    while (remaining_.load(std::memory_order_acquire) > 0) cpu_relax();
  }
  // Takes the workers out of their sleep without giving them work: called at the start of a load, so that the first pass
  // does not pay the wake-up latency of a condition variable (the workers then spin for kSpinUs waiting for the passes).
  void prewake()
  {
    if (n_ == 1) return;
    const unsigned long long next = gen_.load(std::memory_order_relaxed) + 1;
    slot_[next & 1].store(nullptr, std::memory_order_release);
    { std::lock_guard<std::mutex> lk(mu_); gen_.store(next, std::memory_order_release); }
    cv_.notify_all();
  }

 private:
  static constexpr int kSpinUs = 500;       // longer than the serial stretches between two passes of one load
  static void cpu_relax()
  {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
  }
  void loop(int tid)
  {
    unsigned long long last = 0;
    for (;;) {
      bool got = false;
      const auto t0 = std::chrono::steady_clock::now();
      for (int spin = 0;; spin++) {
        if (gen_.load(std::memory_order_acquire) != last) { got = true; break; }
        cpu_relax();
        if ((spin & 63) == 63 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(kSpinUs)) break;
      }
      if (!got) {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || gen_.load(std::memory_order_acquire) != last; });
        if (stop_) return;
      }
      // the job of generation g sits in slot g & 1, written before g is published and rewritten only while g + 2 is being
      // prepared: the slot value is taken for g only if g is still the current generation after the read (otherwise retry
      // with the newer one -- generations without a job, prewake(), can follow each other without waiting for the workers)
      const unsigned long long cand = gen_.load(std::memory_order_acquire);
      const std::function<void(int)>* job = slot_[cand & 1].load(std::memory_order_acquire);
      if (gen_.load(std::memory_order_acquire) != cand) continue;
      last = cand;
      if (job) {
        (*job)(tid);
        remaining_.fetch_sub(1, std::memory_order_release);
      }
    }
  }
  int n_;
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::atomic<unsigned long long> gen_{ 0 };
  std::atomic<int> remaining_{ 0 };
  std::atomic<const std::function<void(int)>*> slot_[2] = { { nullptr }, { nullptr } };
  bool stop_ = false;
};

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
  std::vector<int> pb_cnt;        // pose-block key histograms / cursors, one row per thread
  std::vector<int> thr_unsorted;  // per thread: its range of point indices is not non-decreasing
  std::vector<int> part_pt, part_meas;   // world+1 boundaries of the contiguous point partition
  int npv = 0, nptv = 0, n_slots = 0, max_slots = 1, rs_nblk = 1;
  long long n_inc = 0;            // co-visibility incidences of this rank: sum over its points of K(K+1)/2
  char err[256] = "";
  int par_min_meas = 16384;       // below this many measurements the passes run on the calling thread only
  // scratch, kept between calls
  std::vector<int> key_cnt, order, slot_tmp, thr_bad, thr_max, thr_maxcnt, thr_lo, thr_hist, thr_keysum, thr_itemsum, key_tot;
  std::vector<long long> thr_inc;

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

// `after_points` (may be NULL) is called once the per-measurement, per-point and slot arrays are final -- the caller
// starts their upload while the work lists are still being built.
// `pool` (may be NULL) runs the three heavy passes -- validation + histogram, the per-point pass and the pose-block
// bucketing -- on several host threads; the result is identical for any thread count.
inline int ba_prepare(BaPrep& o, const PrepAlloc& al, int n_cam, int n_pose, const uint8_t* pose_fixed, int n_pt,
                      const int32_t* pt_chain, const uint8_t* pt_fixed, int n_meas, const double* meas_xy,
                      const int32_t* meas_chain, const int32_t* meas_pt, const double* meas_noise, const int32_t* meas_cam,
                      int rank, int world, bool want_rows, HostPool* pool = nullptr,
                      const std::function<void()>* after_points = nullptr)
{
  o.err[0] = 0;
  static const bool prep_trace = getenv("MCP_PREP_TRACE") != nullptr;
  auto prep_t0 = std::chrono::steady_clock::now();
#define PREP_TICK(name) do { if (prep_trace) { auto t1 = std::chrono::steady_clock::now(); fprintf(stderr, "PREP %-14s %8.1f us\n", name, std::chrono::duration<double, std::micro>(t1 - prep_t0).count()); prep_t0 = t1; } } while (0)
  const size_t np1 = (size_t)std::max(n_pt, 1), nm1 = (size_t)std::max(n_meas, 1);
  if (!o.pose_var.resize((size_t)n_pose, al) || !o.pt_var.resize(np1, al) || !o.pt_info.resize(np1, al) ||
      !o.pt_order.resize(np1, al) || !o.pt_meas_off.resize((size_t)n_pt + 1, al) || !o.pt_slot_off.resize((size_t)n_pt + 1, al) ||
      !o.slot_var.resize((size_t)n_meas + n_pt + 1, al) || !o.slot_pt.resize((size_t)n_meas + n_pt + 1, al) ||
      !o.meas_xy.resize(nm1, al) || !o.meas_info.resize(nm1, al) || !o.meas_a.resize(nm1, al) || !o.meas_b.resize(nm1, al))
    MCP_PREP_FAIL(PREP_NOMEM, "mcp_ba_load: host staging allocation failed");
  const int T = (pool && n_meas >= o.par_min_meas) ? pool->size() : 1;
  auto par = [&](const std::function<void(int)>& f) { if (T > 1) pool->run(f); else f(0); };

  int npv = 0;
  for (int i = 0; i < n_pose; i++) o.pose_var[i] = pose_fixed[i] ? -1 : npv++;
  o.npv = npv;
  for (int p = 0; p < n_pt; p++) {
    const int a = pt_chain[2 * p], b = pt_chain[2 * p + 1];
    if (a < 0 || a >= n_pose || b >= n_pose) MCP_PREP_FAIL(PREP_INVALID, "point %d: chain index out of range", p);
    if (b >= 0 && !pose_fixed[b]) MCP_PREP_FAIL(PREP_UNSUPPORTED, "point %d: movable second chain link is not supported", p);
  }
  // validation of the measurements in one parallel pass; the first offending measurement (in index order) is reported, as a
  // sequential scan would.
  int* pmo = o.pt_meas_off.p;
  o.thr_bad.assign((size_t)T * 2, -1);
  o.thr_unsorted.assign((size_t)T, 0);
  // The reference's own loop adds the measurements keyframe by keyframe (src/BundleAdjusterMulti.cc:168-199): that order goes
  // through the parallel counting sort below.  A caller that emits them point by point needs no sort at all -- the validation
  // pass also finds out whether the point indices are non-decreasing, and only if they are not does the sort run.
  par([&](int t) {
    const int lo = (int)((long long)n_meas * t / T), hi = (int)((long long)n_meas * (t + 1) / T);
    int prev = lo > 0 ? meas_pt[lo - 1] : -1, unsorted = 0;
    for (int m = lo; m < hi; m++) {
      const int a = meas_chain[2 * m], b = meas_chain[2 * m + 1];
      int bad = 0;
      if (a < 0 || a >= n_pose || b >= n_pose) bad = 1;
      else if (b >= 0 && !pose_fixed[b]) bad = 2;
      else if (meas_pt[m] < 0 || meas_pt[m] >= n_pt) bad = 3;
      else if (meas_cam[m] < 0 || meas_cam[m] >= n_cam) bad = 4;
      else if (!(meas_noise[m] > 0)) bad = 5;
      if (bad) { o.thr_bad[2 * t] = m; o.thr_bad[2 * t + 1] = bad; return; }
      unsorted |= meas_pt[m] < prev;
      prev = meas_pt[m];
    }
    o.thr_unsorted[t] = unsorted;
  });
  for (int t = 0; t < T; t++) {
    const int m = o.thr_bad[2 * t];
    if (m < 0) continue;
    switch (o.thr_bad[2 * t + 1]) {
      case 1: MCP_PREP_FAIL(PREP_INVALID, "measurement %d: chain index out of range", m);
      case 2: MCP_PREP_FAIL(PREP_UNSUPPORTED, "measurement %d: movable second chain link is not supported", m);
      case 3: MCP_PREP_FAIL(PREP_INVALID, "measurement %d: point index out of range", m);
      case 4: MCP_PREP_FAIL(PREP_INVALID, "measurement %d: camera index out of range", m);
      default: MCP_PREP_FAIL(PREP_INVALID, "measurement %d: noise must be > 0", m);
    }
  }
  PREP_TICK("validate");
  bool point_major = true;
  for (int t = 0; t < T; t++) point_major = point_major && !o.thr_unsorted[t];
  o.meas_orig.resize((size_t)n_meas);
  if (point_major) {
    // pmo[p] = first measurement whose point index is >= p: every thread fills the offsets of the points that begin in its range
    par([&](int t) {
      const int lo = (int)((long long)n_meas * t / T), hi = (int)((long long)n_meas * (t + 1) / T);
      int prev = lo > 0 ? meas_pt[lo - 1] : -1;
      int* orig = o.meas_orig.data();
      for (int m = lo; m < hi; m++) {
        const int p = meas_pt[m];
        for (int pp = prev + 1; pp <= p; pp++) pmo[pp] = m;
        prev = p;
        orig[m] = m;
      }
    });
    for (int pp = (n_meas > 0 ? meas_pt[n_meas - 1] : -1) + 1; pp <= n_pt; pp++) pmo[pp] = n_meas;
    PREP_TICK("offsets(sorted)");
  } else {
    // Stable parallel sort by point in two levels, so that every thread only ever writes memory of its own:
    //   1. source thread s splits its contiguous range of measurements by the DESTINATION thread d that owns the point
    //      (points are dealt to the threads in equal index ranges) into the segment (d, s) of a temporary index array --
    //      T sequential write streams per thread;
    //   2. destination thread d walks its segments in source order (= ascending measurement index: stable), counts per
    //      point, scans, and places the measurement indices -- all inside its own slice of the output.
    // (A one-level scatter by source range makes the threads' 4-byte targets interleave inside cache lines: measured 4-8x slower.)
    auto owner = [&](int p) { return (int)((long long)p * T / std::max(n_pt, 1)); };
    o.thr_hist.assign((size_t)T * (size_t)T + 1, 0);                      // [s][d] counts, then segment cursors
    par([&](int t) {
      int* c = o.thr_hist.data() + (size_t)t * T;
      const int lo = (int)((long long)n_meas * t / T), hi = (int)((long long)n_meas * (t + 1) / T);
      for (int m = lo; m < hi; m++) c[owner(meas_pt[m])]++;
    });
    o.thr_keysum.assign((size_t)T + 1, 0);                               // bucket starts per destination thread
    {
      int run = 0;
      for (int d2 = 0; d2 < T; d2++) {
        o.thr_keysum[d2] = run;
        for (int s2 = 0; s2 < T; s2++) { int& c = o.thr_hist[(size_t)s2 * T + d2]; const int v = c; c = run; run += v; }
      }
      o.thr_keysum[T] = run;
    }
    o.slot_tmp.resize((size_t)n_meas + n_pt + 1);                        // (reused below for the provisional slot lists)
    par([&](int t) {
      int* c = o.thr_hist.data() + (size_t)t * T;
      int* const tmp_m = o.slot_tmp.data();
      const int lo = (int)((long long)n_meas * t / T), hi = (int)((long long)n_meas * (t + 1) / T);
      for (int m = lo; m < hi; m++) tmp_m[c[owner(meas_pt[m])]++] = m;
    });
    PREP_TICK("offsets");
    par([&](int t) {
      // points owned by thread t: [p0, p1); its bucket: tmp_m[b0, b1)
      int p0 = 0, p1 = 0;
      { // inverse of owner(): smallest p with owner(p) >= t
        auto first_of = [&](int d2) { long long v = ((long long)d2 * std::max(n_pt, 1) + T - 1) / T; return (int)std::min<long long>(v, n_pt); };
        p0 = first_of(t); p1 = first_of(t + 1);
      }
      const int b0 = o.thr_keysum[t], b1 = o.thr_keysum[t + 1];
      const int* const tmp_m = o.slot_tmp.data();
      int* orig = o.meas_orig.data();
      for (int p = p0; p < p1; p++) pmo[p] = 0;
      for (int e = b0; e < b1; e++) pmo[meas_pt[tmp_m[e]]]++;
      int run = b0;
      for (int p = p0; p < p1; p++) { const int v = pmo[p]; pmo[p] = run; run += v; }
      // place (pmo[p] doubles as the cursor and is restored by shifting back afterwards)
      for (int e = b0; e < b1; e++) { const int m = tmp_m[e]; orig[pmo[meas_pt[m]]++] = m; }
      for (int p = p1 - 1; p > p0; p--) pmo[p] = pmo[p - 1];
      if (p1 > p0) pmo[p0] = b0;
    });
    pmo[n_pt] = n_meas;
  }
  PREP_TICK("scatter");
  int nptv = 0;
  for (int p = 0; p < n_pt; p++) o.pt_var[p] = pt_fixed[p] ? -1 : nptv++;
  o.nptv = nptv;

  o.part_pt.assign((size_t)world + 1, 0);
  o.part_meas.assign((size_t)world + 1, 0);
  partition_points(pmo, n_pt, world, o.part_pt.data());
  for (int r = 0; r <= world; r++) o.part_meas[r] = pmo[o.part_pt[r]];
  // pose-block keys of this rank's measurements are counted inside the per-point pass (one histogram row per thread)
  const int pb_lo = o.part_meas[rank], pb_hi = o.part_meas[rank + 1];
  const size_t pb_nv = (size_t)std::max(npv, 1), pb_keys = pb_nv * pb_nv;
  o.pb_cnt.assign(pb_keys * (size_t)T, 0);

  // per point: the ascending list of movable poses that carry a Jacobian block of the point (its slots).  Threads take
  // measurement-balanced point ranges; the lists first go to a provisional place (point p at pmo[p] + p: a point has
  // at most one slot per measurement plus its source pose), then an exclusive scan of the list lengths places them.
  o.slot_tmp.resize((size_t)n_meas + n_pt + 1);
  o.thr_max.assign((size_t)T, 1);
  o.thr_maxcnt.assign((size_t)T, 0);
  o.thr_inc.assign((size_t)T, 0);
  const int inc_lo = o.part_pt[rank], inc_hi = o.part_pt[rank + 1];
  o.thr_lo.assign((size_t)T + 1, n_pt);
  for (int t = 0; t < T; t++)
    o.thr_lo[t] = (int)(std::lower_bound(pmo, pmo + n_pt, (int)((long long)n_meas * t / T)) - pmo);
  o.thr_lo[0] = 0;
  par([&](int t) {
    // thread-private scratch and local copies of every pointer: the loop below must not reload them through the closure
    std::vector<int> stamp((size_t)std::max(npv, 1), -1), pos((size_t)std::max(npv, 1), 0), tmp;
    tmp.reserve(64);
    const int* const pose_var_ = o.pose_var.p;
    const int* const pt_var_ = o.pt_var.p;
    const int* const orig_ = o.meas_orig.data();
    const int* const off_ = pmo;
    PInt4* const ma = o.meas_a.p;
    PInt4* const mb = o.meas_b.p;
    PDouble2* const mxy = o.meas_xy.p;
    double* const minfo = o.meas_info.p;
    PInt4* const pinfo = o.pt_info.p;
    int* const slot_cnt = o.pt_slot_off.p;
    int* const slot_tmp = o.slot_tmp.data();
    const double* const in_xy = meas_xy;
    const double* const in_noise = meas_noise;
    const int32_t* const in_chain = meas_chain;
    const int32_t* const in_cam = meas_cam;
    const int32_t* const in_ptchain = pt_chain;
    const int p_lo = o.thr_lo[t], p_hi = o.thr_lo[t + 1];
    int* const kc = o.pb_cnt.data() + pb_keys * (size_t)t;
    int max_slots = 1, max_cnt_t = 0;
    long long inc_t = 0;
    double last_noise = -1.0, last_info = 0.0;
    for (int p = p_lo; p < p_hi; p++) {
      const int src0 = in_ptchain[2 * p], src1 = in_ptchain[2 * p + 1];
      const int src_var = pose_var_[src0];
      const bool movable = pt_var_[p] >= 0;
      const int q_lo = off_[p], q_hi = off_[p + 1];
      max_cnt_t = std::max(max_cnt_t, q_hi - q_lo);
      bool any_src = false;
      tmp.clear();
      for (int q = q_lo; q < q_hi; q++) {
        const int m = orig_[q];
        const int obs0 = in_chain[2 * m];
        const bool has_jac = (obs0 != src0);               // PoseChainHelper::MoveTogether at depth 0
        const int ov = has_jac ? pose_var_[obs0] : -1;
        const bool has_src = has_jac && src_var >= 0;
        any_src |= has_src;
        if (ov >= 0 && movable && stamp[ov] != p) { stamp[ov] = p; tmp.push_back(ov); }
        ma[q] = PInt4{ obs0, in_chain[2 * m + 1], in_cam[m], m };
        mb[q] = PInt4{ ov, -1, has_src ? 1 : 0, p };
        if (ov >= 0 && q >= pb_lo && q < pb_hi) {
          kc[(size_t)ov * pb_nv + ov]++;
          if (has_src) kc[(size_t)std::min(ov, src_var) * pb_nv + std::max(ov, src_var)]++;
        }
        mxy[q] = PDouble2{ in_xy[2 * m], in_xy[2 * m + 1] };
        const double nz = in_noise[m];
        if (nz != last_noise) { last_noise = nz; last_info = 1.0 / std::sqrt(nz); }
        minfo[q] = last_info;                              // 1/sqrt(noise), src/ChainBundle.cc:1244-1245
      }
      int src_slot = -1, K = 0;
      if (movable) {
        if (any_src && stamp[src_var] != p) { stamp[src_var] = p; tmp.push_back(src_var); }
        K = (int)tmp.size();
        int* const tv = tmp.data();
        if (K <= 16) {                                     // insertion sort: the lists are short and nearly sorted
          for (int i = 1; i < K; i++) {
            const int v = tv[i];
            int j = i - 1;
            while (j >= 0 && tv[j] > v) { tv[j + 1] = tv[j]; j--; }
            tv[j + 1] = v;
          }
        } else {
          std::sort(tmp.begin(), tmp.end());
        }
        int* dst = slot_tmp + q_lo + p;
        for (int i = 0; i < K; i++) { dst[i] = tv[i]; pos[tv[i]] = i; }
        max_slots = std::max(max_slots, K);
        if (p >= inc_lo && p < inc_hi) inc_t += (long long)K * (K + 1) / 2;     // co-visibility incidences of this rank
        for (int q = q_lo; q < q_hi; q++)
          if (mb[q].x >= 0) mb[q].y = pos[mb[q].x];
        if (any_src) src_slot = pos[src_var];
      }
      slot_cnt[p + 1] = K;
      pinfo[p] = PInt4{ src0, src1, src_var, src_slot };
    }
    o.thr_max[t] = max_slots;
    o.thr_maxcnt[t] = max_cnt_t;
    o.thr_inc[t] = inc_t;
  });
  PREP_TICK("per-point");
  o.pt_slot_off[0] = 0;
  for (int p = 0; p < n_pt; p++) o.pt_slot_off[p + 1] += o.pt_slot_off[p];
  const int n_slots = o.pt_slot_off[n_pt];
  int max_slots = 1;
  for (int t = 0; t < T; t++) max_slots = std::max(max_slots, o.thr_max[t]);
  par([&](int t) {
    for (int p = o.thr_lo[t]; p < o.thr_lo[t + 1]; p++) {
      const int* src = o.slot_tmp.data() + pmo[p] + p;
      const int s0 = o.pt_slot_off[p], K = o.pt_slot_off[p + 1] - s0;
      for (int i = 0; i < K; i++) { o.slot_var[s0 + i] = src[i]; o.slot_pt[s0 + i] = p; }
    }
  });
  o.n_slots = n_slots; o.max_slots = max_slots;
  o.slot_var.n = o.slot_pt.n = (size_t)std::max(n_slots, 1);
  if (n_slots == 0) { o.slot_var[0] = 0; o.slot_pt[0] = 0; }
  PREP_TICK("slots");
  if (after_points) (*after_points)();
  PREP_TICK("after_points");

  // visiting order of the per-point kernels: inside every rank's range, heaviest points first (stable counting sort
  // by descending measurement count)
  {
    int max_cnt = 0;
    for (int t = 0; t < T; t++) max_cnt = std::max(max_cnt, o.thr_maxcnt[t]);
    std::vector<int>& cnt = o.key_cnt;
    if (n_pt == 0) o.pt_order[0] = 0;
    const size_t nb = (size_t)max_cnt + 1;               // bins: descending measurement count
    for (int r = 0; r < world; r++) {
      const int lo = o.part_pt[r], hi = o.part_pt[r + 1];
      // stable parallel counting sort: one histogram row per thread over its contiguous share of the rank's points
      cnt.assign(nb * (size_t)T, 0);
      par([&](int t) {
        int* c = cnt.data() + nb * (size_t)t;
        const int a = lo + (int)((long long)(hi - lo) * t / T), b = lo + (int)((long long)(hi - lo) * (t + 1) / T);
        for (int p = a; p < b; p++) c[max_cnt - (pmo[p + 1] - pmo[p])]++;
      });
      int run = lo;
      for (size_t bin = 0; bin < nb; bin++)
        for (int t = 0; t < T; t++) { int& c = cnt[nb * (size_t)t + bin]; const int v = c; c = run; run += v; }
      par([&](int t) {
        int* c = cnt.data() + nb * (size_t)t;
        int* const out = o.pt_order.p;
        const int a = lo + (int)((long long)(hi - lo) * t / T), b = lo + (int)((long long)(hi - lo) * (t + 1) / T);
        for (int p = a; p < b; p++) out[c[max_cnt - (pmo[p + 1] - pmo[p])]++] = p;
      });
    }
  }

  // co-visibility incidences of this rank (sizes the pair lists built on the device): summed in the per-point pass
  {
    long long n_inc = 0;
    for (int t = 0; t < T; t++) n_inc += o.thr_inc[t];
    o.n_inc = n_inc;
  }

  PREP_TICK("order+inc");
  // work lists of k_pose_blocks: this rank's measurements bucketed by the pose block they contribute to
  // ((v,v): observed from movable pose v; (lo,hi): observer / source pair), cut into items of <= 128 measurements.
  // Counting sort over the npv^2 block keys; inside a block the measurements keep ascending position: the per-point pass
  // above left one histogram row per thread (thread t = the contiguous positions of its points), the cursors are laid out
  // key-major / thread-minor and the fill pass walks the same ranges.
  {
    const size_t nv = pb_nv, n_keys = pb_keys;
    std::vector<int>& cnt = o.pb_cnt;
    // (key x thread) exclusive scan in two parallel passes over key ranges: totals per key and per range, then the cursors
    // and the work items of every range from the range's base
    o.key_tot.resize(n_keys);
    o.thr_keysum.assign((size_t)T, 0);
    o.thr_itemsum.assign((size_t)T, 0);
    par([&](int t) {
      const size_t k0 = n_keys * (size_t)t / T, k1 = n_keys * (size_t)(t + 1) / T;
      int ents = 0, items = 0;
      for (size_t k = k0; k < k1; k++) {
        int tot = 0;
        for (int u = 0; u < T; u++) tot += cnt[n_keys * (size_t)u + k];
        o.key_tot[k] = tot;
        ents += tot; items += (tot + PREP_PB_CHUNK - 1) / PREP_PB_CHUNK;
      }
      o.thr_keysum[t] = ents; o.thr_itemsum[t] = items;
    });
    size_t n_ent = 0, n_items = 0;
    for (int t = 0; t < T; t++) {
      const int e = o.thr_keysum[t], i2 = o.thr_itemsum[t];
      o.thr_keysum[t] = (int)n_ent; o.thr_itemsum[t] = (int)n_items;
      n_ent += (size_t)e; n_items += (size_t)i2;
    }
    if (!o.pb_idx.resize(std::max(n_ent, (size_t)1), al) || !o.pb_items.resize(std::max(n_items, (size_t)1), al))
      MCP_PREP_FAIL(PREP_NOMEM, "mcp_ba_load: host staging allocation failed");
    PREP_TICK("pb:cursors");
    par([&](int t) {
      const size_t k0 = n_keys * (size_t)t / T, k1 = n_keys * (size_t)(t + 1) / T;
      int run = o.thr_keysum[t];
      size_t it = (size_t)o.thr_itemsum[t];
      for (size_t k = k0; k < k1; k++) {
        const int b = run, e = run + o.key_tot[k];
        for (int u = 0; u < T; u++) { int& c = cnt[n_keys * (size_t)u + k]; const int v = c; c = run; run += v; }
        const int lo = (int)(k / nv), hi = (int)(k % nv);
        for (int s2 = b; s2 < e; s2 += PREP_PB_CHUNK) o.pb_items[it++] = PInt4{ lo, hi, s2, std::min(s2 + PREP_PB_CHUNK, e) };
      }
    });
    o.pb_items.n = n_items;                                // 0 items is legal (nothing movable is observed)
    PREP_TICK("pb:items");
    par([&](int t) {
      int* c = cnt.data() + n_keys * (size_t)t;
      int* const out = o.pb_idx.p;
      const PInt4* const mb = o.meas_b.p;
      const PInt4* const pinfo = o.pt_info.p;
      const int lo = std::max(pmo[o.thr_lo[t]], pb_lo), hi = std::min(pmo[o.thr_lo[t + 1]], pb_hi);
      for (int q = lo; q < hi; q++) {
        const int vo = mb[q].x;
        if (vo < 0) continue;
        out[c[(size_t)vo * nv + vo]++] = q;
        if (mb[q].z) {
          const int vs = pinfo[mb[q].w].z;
          out[c[(size_t)std::min(vo, vs) * nv + std::max(vo, vs)]++] = q;
        }
      }
    });
    o.pb_idx.n = n_ent;
  }

  PREP_TICK("pose-blocks");
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
#undef PREP_TICK

}  // namespace mcp
