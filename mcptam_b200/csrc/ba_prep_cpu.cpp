// ba_prep_cpu.cpp — CPU-only test/benchmark shim around ba_prep.hpp (built with g++ into libmcptam_prep.so).
// Lets the `-m "not gpu"` tests check the host marshalling of mcp_ba_load (ordering, slot lists, work lists, multi-rank
// partition) without a device, and lets tools/ time it.  Not part of the product ABI.
#include <chrono>
#include <cstring>

#include "ba_prep.hpp"

namespace {
void* host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void host_release(void* p) { free(p); }
mcp::BaPrep g_prep;
const mcp::PrepAlloc g_alloc = { host_alloc, host_release };
mcp::HostPool* g_pool = nullptr;
}

extern "C" {

struct McpPrepView {
  const int *pose_var, *pt_var, *pt_info, *pt_order, *pt_meas_off, *pt_slot_off, *slot_var, *slot_pt;
  const double *meas_xy, *meas_info;
  const int *meas_a, *meas_b, *pb_idx, *pb_items, *rs_ent, *rs_grp, *rs_items, *meas_orig, *part_pt, *part_meas;
  long long n_pb_idx, n_pb_items, n_rs_ent, n_rs_grp, n_rs_items, n_inc;
  int npv, nptv, n_slots, max_slots, rs_nblk, pad;
  const char* err;
};

int mcp_prep_run(int n_cam, int n_pose, const uint8_t* pose_fixed, int n_pt, const int32_t* pt_chain, const uint8_t* pt_fixed,
                 int n_meas, const double* meas_xy, const int32_t* meas_chain, const int32_t* meas_pt, const double* meas_noise,
                 const int32_t* meas_cam, int rank, int world, int want_rows, int reps, int threads, int par_min_meas,
                 double* ms_out, McpPrepView* v)
{
  if (!g_pool || g_pool->size() != (threads > 0 ? threads : 1)) { delete g_pool; g_pool = new mcp::HostPool(threads); }
  if (par_min_meas >= 0) g_prep.par_min_meas = par_min_meas;
  int rc = 0;
  double best = 1e300;
  for (int r = 0; r < (reps > 0 ? reps : 1); r++) {
    g_pool->prewake();                                    // as mcp_ba_load does
    const auto t0 = std::chrono::steady_clock::now();
    rc = mcp::ba_prepare(g_prep, g_alloc, n_cam, n_pose, pose_fixed, n_pt, pt_chain, pt_fixed, n_meas, meas_xy, meas_chain,
                         meas_pt, meas_noise, meas_cam, rank, world, want_rows != 0, g_pool);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    best = ms < best ? ms : best;
    if (rc) break;
  }
  if (ms_out) *ms_out = best;
  if (v) {
    const mcp::BaPrep& o = g_prep;
    memset(v, 0, sizeof(*v));
    v->err = o.err;
    if (rc == 0) {
      v->pose_var = o.pose_var.p; v->pt_var = o.pt_var.p; v->pt_info = &o.pt_info.p->x; v->pt_order = o.pt_order.p;
      v->pt_meas_off = o.pt_meas_off.p; v->pt_slot_off = o.pt_slot_off.p; v->slot_var = o.slot_var.p; v->slot_pt = o.slot_pt.p;
      v->meas_xy = &o.meas_xy.p->x; v->meas_info = o.meas_info.p; v->meas_a = &o.meas_a.p->x; v->meas_b = &o.meas_b.p->x;
      v->pb_idx = o.pb_idx.p; v->pb_items = &o.pb_items.p->x;
      v->rs_ent = o.rs_ent.p ? &o.rs_ent.p->x : nullptr; v->rs_grp = o.rs_grp.p; v->rs_items = o.rs_items.p ? &o.rs_items.p->x : nullptr;
      v->meas_orig = o.meas_orig.data(); v->part_pt = o.part_pt.data(); v->part_meas = o.part_meas.data();
      v->n_pb_idx = (long long)o.pb_idx.n; v->n_pb_items = (long long)o.pb_items.n; v->n_rs_ent = (long long)o.rs_ent.n;
      v->n_rs_grp = (long long)o.rs_grp.n; v->n_rs_items = (long long)o.rs_items.n; v->n_inc = o.n_inc;
      v->npv = o.npv; v->nptv = o.nptv; v->n_slots = o.n_slots; v->max_slots = o.max_slots; v->rs_nblk = o.rs_nblk;
    }
  }
  return rc;
}

}  // extern "C"
