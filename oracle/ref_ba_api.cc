// oracle/_ref: the reference's src/ChainBundle.cc, compiled UNMODIFIED (it is #included so that its file-local vertex /
// edge / robust-kernel / action classes are reachable) against the g2o stand-in under oracle/ref_shim/g2o, behind C entry
// points that mirror oracle/ba_oracle.c.  TEST INFRASTRUCTURE ONLY (tests/test_oracle_vs_ref.py).
// What runs here is the reference's own code for: PoseChainHelper::UpdateTransforms, EdgeChainMeas::computeError / chi2 /
// linearizeOplus, VertexPoseSE3 / VertexRelPoint::oplusImpl, RobustKernelData / RobustKernelAdaptive, the convergence
// actions, ChainBundle::AddPose/AddPoint/AddMeas/Compute (Tukey outliers, return codes).  [3P] g2o's LM loop and linear
// solve are the stand-in's dense restatement.
#include <ChainBundle.cc>

#include <cstdint>
#include <cstring>
#include <string>

namespace {
struct RefBa : ChainBundle {
  RefBa(TaylorCameraMap& cams, bool robust, bool tukey) : ChainBundle(cams, robust, tukey, false) {}
  g2o::SparseOptimizer* opt() { return mpOptimizer; }
  RobustKernelData* robust() { return mpRobustKernelData; }
};
struct Handle {
  TaylorCameraMap cams;
  RefBa* ba;
  int n_pose, n_pt, n_meas;
  bool abort_flag;
  std::vector<int> meas_pt;
  Handle() : ba(0), n_pose(0), n_pt(0), n_meas(0), abort_flag(false) {}
};
std::string cam_name(int c) { char b[32]; snprintf(b, sizeof(b), "cam%03d", c); return b; }
TooN::SE3<> unpack(const double* p)
{
  TooN::Matrix<3> R;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R(i, j) = p[3 * i + j];
  return TooN::SE3<>(TooN::SO3<>(R), TooN::makeVector(p[9], p[10], p[11]));
}
void pack(const TooN::SE3<>& T, double* p)
{
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) p[3 * i + j] = T.get_rotation().get_matrix()(i, j);
  for (int i = 0; i < 3; i++) p[9 + i] = T.get_translation()[i];
}
}  // namespace

extern "C" {

void* ref_ba_create(int n_cam, const double* params9, const int32_t* sizes6, int use_robust, int use_tukey)
{
  Handle* h = new Handle;
  for (int c = 0; c < n_cam; c++) {
    TooN::Vector<9> p;
    for (int i = 0; i < 9; i++) p[i] = params9[9 * c + i];
    const int32_t* s = sizes6 + 6 * c;
    h->cams.insert(std::make_pair(cam_name(c), TaylorCamera(p, CVD::ImageRef(s[0], s[1]), CVD::ImageRef(s[2], s[3]), CVD::ImageRef(s[4], s[5]))));
  }
  h->ba = new RefBa(h->cams, use_robust != 0, use_tukey != 0);
  return h;
}
void ref_ba_destroy(void* hv) { Handle* h = static_cast<Handle*>(hv); delete h->ba; delete h; }

// same flat problem as mcp_ba_load / ora_ba_load; ids follow the reference: one counter from 1, poses first
int ref_ba_load(void* hv, int n_pose, const double* pose_Rt, const uint8_t* pose_fixed, int n_pt, const double* pt_xyz, const int32_t* pt_chain,
                const uint8_t* pt_fixed, int n_meas, const double* meas_xy, const int32_t* meas_chain, const int32_t* meas_pt, const double* meas_noise,
                const int32_t* meas_cam)
{
  Handle* h = static_cast<Handle*>(hv);
  h->n_pose = n_pose; h->n_pt = n_pt; h->n_meas = n_meas;
  for (int i = 0; i < n_pose; i++) { const int id = h->ba->AddPose(unpack(pose_Rt + 12 * i), pose_fixed[i] != 0); if (id != i + 1) return -1; }
  for (int p = 0; p < n_pt; p++) {
    std::vector<int> chain;
    for (int k = 0; k < 2; k++) if (pt_chain[2 * p + k] >= 0) chain.push_back(pt_chain[2 * p + k] + 1);
    const int id = h->ba->AddPoint(TooN::makeVector(pt_xyz[3 * p], pt_xyz[3 * p + 1], pt_xyz[3 * p + 2]), chain, pt_fixed[p] != 0);
    if (id != n_pose + 1 + p) return -1;
  }
  h->meas_pt.assign(meas_pt, meas_pt + n_meas);
  for (int m = 0; m < n_meas; m++) {
    std::vector<int> chain;
    for (int k = 0; k < 2; k++) if (meas_chain[2 * m + k] >= 0) chain.push_back(meas_chain[2 * m + k] + 1);
    h->ba->AddMeas(chain, n_pose + 1 + meas_pt[m], TooN::makeVector(meas_xy[2 * m], meas_xy[2 * m + 1]), meas_noise[m], cam_name(meas_cam[m]));
  }
  return 0;
}

// errors (2 per measurement) and EdgeChainMeas::chi2 (signed) at the current estimate
void ref_ba_eval(void* hv, double* err_xy, double* chi2)
{
  Handle* h = static_cast<Handle*>(hv);
  h->ba->Initialize();
  h->ba->opt()->computeActiveErrors();
  const g2o::OptimizableGraph::EdgeContainer& E = h->ba->opt()->activeEdges();
  for (size_t m = 0; m < E.size(); m++) {
    const double* e = E[m]->errorData();
    err_xy[2 * m] = e[0]; err_xy[2 * m + 1] = e[1];
    chi2[m] = E[m]->chi2();
  }
}
// EdgeChainMeas::linearizeOplus of measurement m: one 2 x 6 block per chain link (observer links, then source links: up to 2 each,
// zero if absent or fixed) and the 2 x 3 point block
void ref_ba_jacobians(void* hv, int m, double* J_obs24, double* J_src24, double* J_pt6)
{
  Handle* h = static_cast<Handle*>(hv);
  h->ba->Initialize();
  h->ba->opt()->computeActiveErrors();
  g2o::OptimizableGraph::Edge* e = h->ba->opt()->activeEdges()[m];
  EdgeChainMeas* em = dynamic_cast<EdgeChainMeas*>(e);
  e->linearizeOplus();
  memset(J_obs24, 0, sizeof(double) * 24); memset(J_src24, 0, sizeof(double) * 24); memset(J_pt6, 0, sizeof(double) * 6);
  const int N = (int)em->_pPoseChainHelper->_vpVertices.size();
  const int total = (int)e->vertices().size(), pointN = total - 1 - N;
  for (int i = 0; i < total; i++) {
    g2o::OptimizableGraph::Vertex* v = static_cast<g2o::OptimizableGraph::Vertex*>(e->vertices()[i]);
    if (v->fixed()) continue;
    const double* J = e->jacobianData(i);
    if (i < N) memcpy(J_obs24 + 12 * i, J, sizeof(double) * 12);
    else if (i < N + pointN) memcpy(J_src24 + 12 * (i - N), J, sizeof(double) * 12);
    else memcpy(J_pt6, J, sizeof(double) * 6);
  }
}
// VertexPoseSE3::oplusImpl / VertexRelPoint::oplusImpl on a copy of the estimate
void ref_ba_oplus_pose(void* hv, int i, const double* d6, double* out12)
{
  Handle* h = static_cast<Handle*>(hv);
  VertexPoseSE3* v = dynamic_cast<VertexPoseSE3*>(h->ba->opt()->vertex(i + 1));
  v->push(); v->oplus(d6); pack(v->estimate(), out12); v->pop();
}
void ref_ba_oplus_point(void* hv, int p, const double* d3, double* out3)
{
  Handle* h = static_cast<Handle*>(hv);
  VertexRelPoint* v = dynamic_cast<VertexRelPoint*>(h->ba->opt()->vertex(h->n_pose + 1 + p));
  v->push(); v->oplus(d3);
  for (int k = 0; k < 3; k++) out3[k] = v->estimate()[k];
  v->pop();
}
// RobustKernelAdaptive::robustify with sigma^2 recomputed from the current errors (RobustKernelData::RecomputeNow)
void ref_ba_robustify(void* hv, double* rho3_per_meas, double* sigma_sq)
{
  Handle* h = static_cast<Handle*>(hv);
  h->ba->Initialize();
  h->ba->opt()->computeActiveErrors();
  const g2o::OptimizableGraph::EdgeContainer& E = h->ba->opt()->activeEdges();
  h->ba->robust()->RecomputeNextTime();                  // what UpdateSigmaSquaredAction does before every iteration
  h->ba->opt()->activeRobustChi2();
  *sigma_sq = h->ba->GetSigmaSquared();
  Eigen::Vector3d rho;
  for (size_t m = 0; m < E.size(); m++) {
    if (E[m]->robustKernel()) E[m]->robustKernel()->robustify(E[m]->chi2(), rho);
    else { rho[0] = E[m]->chi2(); rho[1] = 1; rho[2] = 0; }
    for (int k = 0; k < 3; k++) rho3_per_meas[3 * m + k] = rho[k];
  }
}
// ChainBundle::Compute.  stats: [total iterations (trials), converged, sigma^2, mean chi2, lambda, max cov]
int ref_ba_compute(void* hv, int n_iter, double user_lambda, double* stats6)
{
  Handle* h = static_cast<Handle*>(hv);
  h->abort_flag = false;
  const int rc = h->ba->Compute(&h->abort_flag, n_iter, user_lambda);
  stats6[0] = h->ba->TotalIterations(); stats6[1] = h->ba->Converged() ? 1 : 0; stats6[2] = h->ba->GetSigmaSquared();
  stats6[3] = h->ba->GetMeanChiSquared(); stats6[4] = h->ba->GetLambda(); stats6[5] = h->ba->GetMaxCov();
  return rc;
}
void ref_ba_get_state(void* hv, double* pose_Rt, double* pt_xyz)
{
  Handle* h = static_cast<Handle*>(hv);
  for (int i = 0; i < h->n_pose; i++) pack(h->ba->GetPose(i + 1), pose_Rt + 12 * i);
  for (int p = 0; p < h->n_pt; p++) { const TooN::Vector<3> v = h->ba->GetPoint(h->n_pose + 1 + p); for (int k = 0; k < 3; k++) pt_xyz[3 * p + k] = v[k]; }
}
// outliers as measurement indices: (point id, first chain vertex id, camera name) -> the measurement that carries them
int ref_ba_get_outliers(void* hv, const int32_t* meas_chain, const int32_t* meas_cam, int32_t* idx, int cap)
{
  Handle* h = static_cast<Handle*>(hv);
  std::vector<std::tuple<int, int, std::string> > out = h->ba->GetOutlierMeasurements();
  int n = 0;
  for (size_t k = 0; k < out.size(); k++) {
    const int pid = std::get<0>(out[k]) - h->n_pose - 1, base = std::get<1>(out[k]) - 1;
    for (int m = 0; m < h->n_meas; m++)
      if (h->meas_pt[m] == pid && meas_chain[2 * m] == base && cam_name(meas_cam[m]) == std::get<2>(out[k])) { if (n < cap) idx[n] = m; n++; break; }
  }
  return n;
}

}  // extern "C"
