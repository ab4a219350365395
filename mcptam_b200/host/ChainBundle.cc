#include "ChainBundle.h"

#include <algorithm>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace mcp_host {

int ChainBundle::snMaxIterations = 100;
int ChainBundle::snMaxTrialsAfterFailure = 100;
double ChainBundle::sdUpdatePercentConvergenceLimit = 1e-10;
double ChainBundle::sdUpdateRMSConvergenceLimit = 1e-10;
double ChainBundle::sdMinMEstimatorSigma = 0.5;

// one cached device handle per host thread (the MapMaker thread): device buffers, pinned staging and the marshalling
// threads survive across BundleAdjust calls -- the reference builds a new ChainBundle on the stack for every call
// (src/BundleAdjusterMulti.cc:75).  The handle is reused when the next object asks for the same configuration.
struct PooledHandle {
  McpBa* h = nullptr;
  McpBaConfig cfg;
  ~PooledHandle() { if (h) mcp_ba_destroy(h); }            // released when the thread ends
};
static thread_local PooledHandle tl_pool;

// ... and so do the flat host arrays: a ChainBundle hands its vectors (cleared, capacity kept) to the next one built on the
// thread.  Freshly allocated megabyte vectors cost their first-touch page faults on every BundleAdjust call otherwise
// (~2 ms of the ~7 ms marshalling of a 200 KF map).
struct PooledArrays {
  std::vector<int> idKind, idIndex, ptId;
  std::vector<double> poseRt, ptXyz, measXy, measNoise;
  std::vector<uint8_t> poseFixed, ptFixed;
  std::vector<int32_t> ptChain, measChain, measPt, measCam, measFirstId;
};
static thread_local std::vector<PooledArrays> tl_arrays;      // at most two sets are kept
#define tl_pooled tl_pool.h
#define tl_pooled_cfg tl_pool.cfg

ChainBundle::ChainBundle(TaylorCameraMap& cams, bool bUseRobust, bool bUseTukey, bool bVerbose)
    : mmCameraModels(cams), mbUseRobust(bUseRobust), mbUseTukey(bUseTukey), mbVerbose(bVerbose)
{
  if (!tl_arrays.empty()) {
    PooledArrays& a = tl_arrays.back();
    mvIdKind.swap(a.idKind); mvIdIndex.swap(a.idIndex); mvPtId.swap(a.ptId);
    mvPoseRt.swap(a.poseRt); mvPtXyz.swap(a.ptXyz); mvMeasXy.swap(a.measXy); mvMeasNoise.swap(a.measNoise);
    mvPoseFixed.swap(a.poseFixed); mvPtFixed.swap(a.ptFixed);
    mvPtChain.swap(a.ptChain); mvMeasChain.swap(a.measChain); mvMeasPt.swap(a.measPt); mvMeasCam.swap(a.measCam); mvMeasFirstId.swap(a.measFirstId);
    tl_arrays.pop_back();
  }
  mvIdKind.push_back(-1); mvIdIndex.push_back(-1);     // id 0 unused
}
ChainBundle::~ChainBundle()
{
  if (tl_arrays.size() < 2) {
    tl_arrays.emplace_back();
    PooledArrays& a = tl_arrays.back();
    mvIdKind.clear(); mvIdIndex.clear(); mvPtId.clear(); mvPoseRt.clear(); mvPtXyz.clear(); mvMeasXy.clear(); mvMeasNoise.clear();
    mvPoseFixed.clear(); mvPtFixed.clear(); mvPtChain.clear(); mvMeasChain.clear(); mvMeasPt.clear(); mvMeasCam.clear(); mvMeasFirstId.clear();
    mvIdKind.swap(a.idKind); mvIdIndex.swap(a.idIndex); mvPtId.swap(a.ptId);
    mvPoseRt.swap(a.poseRt); mvPtXyz.swap(a.ptXyz); mvMeasXy.swap(a.measXy); mvMeasNoise.swap(a.measNoise);
    mvPoseFixed.swap(a.poseFixed); mvPtFixed.swap(a.ptFixed);
    mvPtChain.swap(a.ptChain); mvMeasChain.swap(a.measChain); mvMeasPt.swap(a.measPt); mvMeasCam.swap(a.measCam); mvMeasFirstId.swap(a.measFirstId);
  }
  if (mpHandle) {
    // a handle whose last call failed (CUDA error state) or that joined a communicator is not worth keeping
    if (!tl_pooled && !mbHandleFailed) { tl_pooled = mpHandle; tl_pooled_cfg = mConfig; } else mcp_ba_destroy(mpHandle);
  }
}

int ChainBundle::AddPose(SE3 pose, bool bFixed)
{
  double rt[12];
  pose.pack(rt);
  mvPoseRt.insert(mvPoseRt.end(), rt, rt + 12);
  mvPoseFixed.push_back(bFixed ? 1 : 0);
  mvIdKind.push_back(0); mvIdIndex.push_back((int)mvPoseFixed.size() - 1);
  return mnCurrId++;
}
int ChainBundle::AddPoint(Vector<3> p, std::vector<int> vCams, bool bFixed)
{
  if (vCams.empty() || vCams.size() > 2) throw std::invalid_argument("ChainBundle::AddPoint: chains of 1 or 2 poses are supported");
  for (int k = 0; k < 3; k++) mvPtXyz.push_back(p[k]);
  for (int k = 0; k < 2; k++) mvPtChain.push_back(k < (int)vCams.size() ? mvIdIndex.at(vCams[k]) : -1);
  mvPtFixed.push_back(bFixed ? 1 : 0);
  mvIdKind.push_back(1); mvIdIndex.push_back((int)mvPtFixed.size() - 1);
  mvPtId.push_back(mnCurrId);
  return mnCurrId++;
}
int ChainBundle::AddPoint(const Vector<3>& p, int nPose0, int nPose1, bool bFixed)
{
  for (int k = 0; k < 3; k++) mvPtXyz.push_back(p[k]);
  mvPtChain.push_back(mvIdIndex.at(nPose0));
  mvPtChain.push_back(nPose1 >= 0 ? mvIdIndex.at(nPose1) : -1);
  mvPtFixed.push_back(bFixed ? 1 : 0);
  mvIdKind.push_back(1); mvIdIndex.push_back((int)mvPtFixed.size() - 1);
  mvPtId.push_back(mnCurrId);
  return mnCurrId++;
}
void ChainBundle::AddMeas(std::vector<int> vCams, int nPointIdx, Vector<2> v2Pos, double dNoiseSigmaSquared, std::string cameraName)
{
  if (vCams.empty() || vCams.size() > 2) throw std::invalid_argument("ChainBundle::AddMeas: chains of 1 or 2 poses are supported");
  mvMeasXy.push_back(v2Pos[0]); mvMeasXy.push_back(v2Pos[1]);
  for (int k = 0; k < 2; k++) mvMeasChain.push_back(k < (int)vCams.size() ? mvIdIndex.at(vCams[k]) : -1);
  mvMeasPt.push_back(mvIdIndex.at(nPointIdx));
  mvMeasNoise.push_back(dNoiseSigmaSquared);
  mvMeasFirstId.push_back(vCams[0]);
  mvMeasCam.push_back(CameraIndex(cameraName));
}

int ChainBundle::CameraIndex(const std::string& cameraName)
{
  for (size_t i = 0; i < mvCamNames.size(); i++) if (mvCamNames[i] == cameraName) return (int)i;
  if (!mmCameraModels.count(cameraName)) throw std::invalid_argument("ChainBundle::AddMeas: unknown camera " + cameraName);
  mvCamNames.push_back(cameraName);
  return (int)mvCamNames.size() - 1;
}

void ChainBundle::AddMeas(int nBasePoseId, int nCamPoseId, int nPointIdx, const Vector<2>& v2Pos, double dNoiseSigmaSquared, int nCameraIndex)
{
  mvMeasXy.push_back(v2Pos[0]); mvMeasXy.push_back(v2Pos[1]);
  mvMeasChain.push_back(mvIdIndex.at(nBasePoseId));
  mvMeasChain.push_back(nCamPoseId >= 0 ? mvIdIndex.at(nCamPoseId) : -1);
  mvMeasPt.push_back(mvIdIndex.at(nPointIdx));
  mvMeasNoise.push_back(dNoiseSigmaSquared);
  mvMeasFirstId.push_back(nBasePoseId);
  mvMeasCam.push_back(nCameraIndex);
}

void ChainBundle::AddMeasBlock(int nBasePoseId, int nCamPoseId, int nCameraIndex, size_t n, const int* pPointIds, const double* pXy, const double* pNoise)
{
  const int32_t nBase = mvIdIndex.at(nBasePoseId), nCam = nCamPoseId >= 0 ? mvIdIndex.at(nCamPoseId) : -1;
  mvMeasXy.insert(mvMeasXy.end(), pXy, pXy + 2 * n);
  mvMeasNoise.insert(mvMeasNoise.end(), pNoise, pNoise + n);
  for (size_t i = 0; i < n; i++) {
    mvMeasChain.push_back(nBase); mvMeasChain.push_back(nCam);
    mvMeasPt.push_back(mvIdIndex.at(pPointIds[i]));
  }
  mvMeasFirstId.insert(mvMeasFirstId.end(), n, nBasePoseId);
  mvMeasCam.insert(mvMeasCam.end(), n, nCameraIndex);
}

size_t ChainBundle::GrowMeas(size_t n)
{
  const size_t at = mvMeasPt.size();
  mvMeasXy.resize(2 * (at + n)); mvMeasNoise.resize(at + n); mvMeasChain.resize(2 * (at + n)); mvMeasPt.resize(at + n);
  mvMeasFirstId.resize(at + n); mvMeasCam.resize(at + n);
  return at;
}
void ChainBundle::FillMeasBlock(size_t nAt, int nBasePoseId, int nCamPoseId, int nCameraIndex, size_t n, const int* pPointIds, const double* pXy, const double* pNoise)
{
  const int32_t nBase = mvIdIndex.at(nBasePoseId), nCam = nCamPoseId >= 0 ? mvIdIndex.at(nCamPoseId) : -1;
  std::copy(pXy, pXy + 2 * n, mvMeasXy.begin() + 2 * nAt);
  std::copy(pNoise, pNoise + n, mvMeasNoise.begin() + nAt);
  for (size_t i = 0; i < n; i++) {
    mvMeasChain[2 * (nAt + i)] = nBase; mvMeasChain[2 * (nAt + i) + 1] = nCam;
    mvMeasPt[nAt + i] = mvIdIndex.at(pPointIds[i]);
    mvMeasFirstId[nAt + i] = nBasePoseId;
    mvMeasCam[nAt + i] = nCameraIndex;
  }
}

void ChainBundle::Reserve(size_t nPoses, size_t nPoints, size_t nMeas)
{
  mvPoseRt.reserve(12 * nPoses); mvPoseFixed.reserve(nPoses);
  mvPtXyz.reserve(3 * nPoints); mvPtChain.reserve(2 * nPoints); mvPtFixed.reserve(nPoints); mvPtId.reserve(nPoints);
  mvIdKind.reserve(nPoses + nPoints + 1); mvIdIndex.reserve(nPoses + nPoints + 1);
  mvMeasXy.reserve(2 * nMeas); mvMeasChain.reserve(2 * nMeas); mvMeasPt.reserve(nMeas); mvMeasNoise.reserve(nMeas);
  mvMeasFirstId.reserve(nMeas); mvMeasCam.reserve(nMeas);
}

int ChainBundle::Upload()
{
  if (!mpHandle) {
    McpBaConfig cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    mcp_ba_default_config(&cfg);
    cfg.use_robust = mbUseRobust; cfg.use_tukey = mbUseTukey; cfg.verbose = mbVerbose;
    cfg.max_trials_after_failure = snMaxTrialsAfterFailure;
    cfg.update_pct_limit = sdUpdatePercentConvergenceLimit; cfg.update_rms_limit = sdUpdateRMSConvergenceLimit;
    cfg.min_sigma = sdMinMEstimatorSigma;
    if (tl_pooled && std::memcmp(&tl_pooled_cfg, &cfg, sizeof(cfg)) == 0) {
      mpHandle = tl_pooled;                      // same configuration: keep the pooled buffers
      tl_pooled = nullptr;
    } else {
      if (tl_pooled) { mcp_ba_destroy(tl_pooled); tl_pooled = nullptr; }
      int rc = mcp_ba_create(&cfg, &mpHandle);
      if (rc) { mbHandleFailed = true; return rc; }
    }
    mConfig = cfg;
  }
  std::vector<McpTaylorCam> cams;
  for (auto& n : mvCamNames) cams.push_back(mmCameraModels[n].ToAbi());
  if (cams.empty()) return MCP_ERR_STATE;
  int rc = mcp_ba_set_cameras(mpHandle, (int)cams.size(), cams.data());
  if (rc) { mbHandleFailed = (rc == MCP_ERR_CUDA); return rc; }
  rc = mcp_ba_load(mpHandle, (int)mvPoseFixed.size(), mvPoseRt.data(), mvPoseFixed.data(), (int)mvPtFixed.size(), mvPtXyz.data(),
                   mvPtChain.data(), mvPtFixed.data(), (int)mvMeasPt.size(), mvMeasXy.data(), mvMeasChain.data(), mvMeasPt.data(),
                   mvMeasNoise.data(), mvMeasCam.data());
  if (rc) { mbHandleFailed = (rc == MCP_ERR_CUDA || rc == MCP_ERR_NCCL); return rc; }
  mbUploaded = true;
  return MCP_OK;
}

void ChainBundle::Fetch()
{
  mcp_ba_get_poses(mpHandle, mvPoseRt.data());
  mcp_ba_get_points(mpHandle, mvPtXyz.data());
}

int ChainBundle::Compute(bool* pAbortSignal, int nNumIter, double dUserLambda)
{
  mvOutlierMeasurementIdx.clear();
  if (!mbUploaded) {
    const int rc = Upload();
    if (rc) { std::fprintf(stderr, "ChainBundle: %s\n", mcp_last_error()); return -1; }
  }
  McpBaStats st;
  static_assert(sizeof(bool) == 1, "abort flag is polled as a byte");
  const int n = mcp_ba_compute(mpHandle, reinterpret_cast<volatile const uint8_t*>(pAbortSignal), nNumIter, dUserLambda, &st);
  if (n < -1) { mbHandleFailed = true; std::fprintf(stderr, "ChainBundle: %s\n", mcp_last_error()); return -1; }
  mbConverged = st.converged != 0;
  // the reference's convergence actions raise the shared abort flag (src/ChainBundle.cc:1028,1106)
  if (mbConverged && pAbortSignal) *pAbortSignal = true;
  mnTotalIterations = st.total_trials;
  mdSigmaSquared = st.sigma_sq; mdMeanChiSquared = st.mean_chi2; mdLambda = st.lambda; mdLastMaxCov = st.max_cov; mdGpuMs = st.gpu_ms;
  Fetch();
  if (n > 0 && mbUseTukey) {
    std::vector<int32_t> idx(st.n_outliers > 0 ? st.n_outliers : 1);
    const int no = mcp_ba_get_outliers(mpHandle, idx.data(), (int)idx.size());
    for (int i = 0; i < no && i < (int)idx.size(); i++) {
      const int m = idx[i];
      const int pid = mvPtId[mvMeasPt[m]];                    // bundle id of the point
      mvOutlierMeasurementIdx.push_back(std::make_tuple(pid, mvMeasFirstId[m], mvCamNames[mvMeasCam[m]]));
    }
  }
  return n;
}

Vector<3> ChainBundle::GetPoint(int n)
{
  const int i = mvIdIndex.at(n);
  return makeVector(mvPtXyz[3 * i], mvPtXyz[3 * i + 1], mvPtXyz[3 * i + 2]);
}
SE3 ChainBundle::GetPose(int n) { return SE3::unpack(&mvPoseRt[12 * (size_t)mvIdIndex.at(n)]); }
std::vector<std::tuple<int, int, std::string> > ChainBundle::GetOutlierMeasurements() { return mvOutlierMeasurementIdx; }

}  // namespace mcp_host
