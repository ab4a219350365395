#include "BundleAdjusterCuda.h"
#include "Workers.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <thread>

namespace mcp_host {

int BundleAdjusterBase::snMinMapPoints = 10;
static inline int LevelScale(int l) { return 1 << l; }     // include/mcptam/LevelHelpers.h:55-58

// Follows src/BundleAdjusterMulti.cc:55-264 step by step.
int BundleAdjusterCuda::BundleAdjust(std::set<MultiKeyFrame*> spAdjustSet, std::set<MultiKeyFrame*> spFixedSet, std::set<MapPoint*> spMapPoints,
                                     std::vector<std::pair<KeyFrame*, MapPoint*> >& vOutliers, bool bRecent)
{
  mmPoint_BundleID.clear(); mmBundleID_Point.clear(); mmBase_BundleID.clear(); mmBundleID_Base.clear(); mmCamName_BundleID.clear();
  if ((int)spMapPoints.size() < snMinMapPoints) return 0;
  ChainBundle multiBundle(mmCameraModels, mbUseRobust, mbUseTukey, mbVerbose);
  mbBundleRunning = true;
  mbBundleRunningIsRecent = bRecent;
  Marshal(multiBundle, spAdjustSet, spFixedSet, spMapPoints);
  int nAccepted = 0;
  mnTotalIterations = 0;
  mdGpuMs = 0;
  if (mbUseTwoStep) {
    nAccepted = AdjustAndUpdate(multiBundle, spAdjustSet, spMapPoints, 10);
    mnTotalIterations = multiBundle.TotalIterations();
    if (nAccepted < 0) return nAccepted;
    if (!multiBundle.Converged()) {
      // the first pass raised the abort flag only on convergence; an external abort request stays set
      nAccepted += AdjustAndUpdate(multiBundle, spAdjustSet, spMapPoints);
      mnTotalIterations += multiBundle.TotalIterations();
    }
  } else {
    nAccepted = AdjustAndUpdate(multiBundle, spAdjustSet, spMapPoints);
    mnTotalIterations = multiBundle.TotalIterations();
  }
  if (nAccepted < 0) return nAccepted;
  if (multiBundle.Converged()) {
    mbBundleConverged_Recent = true;
    if (!mbBundleRunningIsRecent) mbBundleConverged_Full = true;
  }
  for (auto& o : multiBundle.GetOutlierMeasurements()) {
    MapPoint* pPoint = mmBundleID_Point[std::get<0>(o)];
    MultiKeyFrame* pMKF = mmBundleID_Base[std::get<1>(o)];
    vOutliers.push_back(std::make_pair(pMKF->mmpKeyFrames[std::get<2>(o)], pPoint));
  }
  mbBundleRunning = false;
  mbBundleAbortRequested = false;
  return nAccepted;
}

// src/BundleAdjusterMulti.cc:83-203
void BundleAdjusterCuda::Marshal(ChainBundle& multiBundle, std::set<MultiKeyFrame*>& spAdjustSet, std::set<MultiKeyFrame*>& spFixedSet,
                                 std::set<MapPoint*>& spMapPoints)
{
  static const bool bTrace = getenv("MCP_HOST_TRACE") != nullptr;
  auto tTick = std::chrono::steady_clock::now();
  auto HOST_TICK = [&](const char* what) { if (bTrace) { auto n = std::chrono::steady_clock::now(); fprintf(stderr, "MARSHAL %-10s %7.2f ms\n", what, std::chrono::duration<double, std::milli>(n - tTick).count()); tTick = n; } };
  size_t nMeasUpper = 0;
  for (MapPoint* pp : spMapPoints) nMeasUpper += pp->mMMData.spMeasurementKFs.size();
  multiBundle.Reserve(spAdjustSet.size() + spFixedSet.size() + 16, spMapPoints.size(), nMeasUpper);
  mmPoint_BundleID.reserve(spMapPoints.size() * 2);
  for (int pass = 0; pass < 2; pass++) {
    std::set<MultiKeyFrame*>& s = pass == 0 ? spAdjustSet : spFixedSet;
    for (MultiKeyFrame* pm : s) {
      MultiKeyFrame& mkf = *pm;
      if (mkf.mbBad) continue;
      const int id = multiBundle.AddPose(mkf.mse3BaseFromWorld, pass == 0 ? mkf.mbFixed : true);
      mmBase_BundleID[&mkf] = id;
      if ((int)mmBundleID_Base.size() <= id) mmBundleID_Base.resize((size_t)id + 1, nullptr);
      mmBundleID_Base[id] = &mkf;
      for (auto& kv : mkf.mmpKeyFrames)
        if (!mmCamName_BundleID.count(kv.first)) mmCamName_BundleID[kv.first] = multiBundle.AddPose(kv.second->mse3CamFromBase, true);
    }
  }
  HOST_TICK("poses");
  int nWorldID = -1;
  KeyFrame* pLastSrc = nullptr;
  int nLastBase = -1, nLastCam = -1;
  for (MapPoint* pp : spMapPoints) {
    MapPoint& point = *pp;
    Vector<3> v3Pos;
    int nPose0, nPose1 = -1;
    if (point.mbFixed) {
      if (nWorldID == -1) nWorldID = multiBundle.AddPose(SE3(), true);
      v3Pos = point.mv3WorldPos;
      nPose0 = nWorldID;
    } else {
      KeyFrame& src = *point.mpPatchSourceKF;
      v3Pos = src.mse3CamFromWorld * point.mv3WorldPos;
      // (the source keyframe's two ids are looked up once per keyframe, not once per point)
      if (&src != pLastSrc) { pLastSrc = &src; nLastBase = mmBase_BundleID[src.mpParent]; nLastCam = mmCamName_BundleID[src.mCamName]; }
      nPose0 = nLastBase; nPose1 = nLastCam;
    }
    const int id = multiBundle.AddPoint(v3Pos, nPose0, nPose1, point.mbFixed);
    mmPoint_BundleID[&point] = id;
    if ((int)mmBundleID_Point.size() <= id) mmBundleID_Point.resize((size_t)id + 1 + spMapPoints.size(), nullptr);
    mmBundleID_Point[id] = &point;
  }
  // The measurement loop (src/BundleAdjusterMulti.cc:168-199) is a walk over every keyframe's std::map of measurements: pointer
  // chasing, ~0.2 us per measurement on one thread.  The keyframes are independent, so a few threads each gather whole
  // keyframes into private buffers (point ids, positions, noise); the buffers are then appended in the reference's order --
  // the arrays handed to mcp_ba_load are exactly those of the sequential loop (host/test_marshal_cpu.cc).
  HOST_TICK("points");
  std::vector<KfJob>& jobs = mvJobs;
  size_t nJobs = 0;
  for (auto& mb : mmBase_BundleID)
    for (auto& kv : mb.first->mmpKeyFrames) {
      if (jobs.size() <= nJobs) jobs.emplace_back();
      KfJob& j = jobs[nJobs++];
      j.kf = kv.second; j.name = &kv.first; j.nBaseID = mb.second; j.nCamID = mmCamName_BundleID[kv.first];
      j.ids.clear(); j.xy.clear(); j.noise.clear();
    }
  jobs.resize(nJobs);
  std::atomic<size_t> next(0);
  auto gather = [&]() {
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= jobs.size()) break;
      KfJob& j = jobs[i];
      const size_t cap = j.kf->mmpMeasurements.size();
      j.ids.reserve(cap); j.xy.reserve(2 * cap); j.noise.reserve(cap);
      for (auto& mm : j.kf->mmpMeasurements) {
        const auto itPoint = mmPoint_BundleID.find(mm.first);
        if (itPoint == mmPoint_BundleID.end()) continue;
        const Measurement& meas = *mm.second;
        j.ids.push_back(itPoint->second);
        j.xy.push_back(meas.v2RootPos[0]); j.xy.push_back(meas.v2RootPos[1]);
        j.noise.push_back((double)(LevelScale(meas.nLevel) * LevelScale(meas.nLevel)));
      }
    }
  };
  const unsigned hw = std::thread::hardware_concurrency();
  int nWant = (int)std::min<unsigned>(hw ? hw / 2 : 1, 8);
  if (const char* e = getenv("MCP_HOST_THREADS")) nWant = std::max(1, atoi(e));
  const int nThreads = nMeasUpper < 4096 ? 1 : (int)std::min<size_t>((size_t)nWant, jobs.size());
  auto run_parallel = [&](const std::function<void()>& f) { Workers::Get().Run(nThreads, f); };
  run_parallel(gather);
  HOST_TICK("gather");
  // camera indices are handed out at a keyframe's first measurement, in the reference's keyframe order; then every block gets
  // its place and the blocks are written in parallel
  std::vector<size_t> at(jobs.size(), 0);
  std::vector<int> camIndex(jobs.size(), -1);
  size_t nTotal = 0;
  for (size_t i = 0; i < jobs.size(); i++) {
    if (jobs[i].ids.empty()) continue;
    camIndex[i] = multiBundle.CameraIndex(*jobs[i].name);
    at[i] = nTotal; nTotal += jobs[i].ids.size();
  }
  const size_t nBase0 = multiBundle.GrowMeas(nTotal);
  next.store(0);
  run_parallel([&]() {
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= jobs.size()) break;
      const KfJob& j = jobs[i];
      if (j.ids.empty()) continue;
      multiBundle.FillMeasBlock(nBase0 + at[i], j.nBaseID, j.nCamID, camIndex[i], j.ids.size(), j.ids.data(), j.xy.data(), j.noise.data());
    }
  });
  HOST_TICK("append");
}

// src/BundleAdjusterMulti.cc:267-337
int BundleAdjusterCuda::AdjustAndUpdate(ChainBundle& multiBundle, std::set<MultiKeyFrame*> spAdjustSet, std::set<MapPoint*> spMapPoints, int nIterations)
{
  int nAccepted = nIterations <= 0 ? multiBundle.Compute(&mbBundleAbortRequested) : multiBundle.Compute(&mbBundleAbortRequested, nIterations);
  mdGpuMs += multiBundle.LastGpuMs();
  if (nAccepted < 0) return -1;
  if (nAccepted > 0 && mbApplyUpdates) {
    if (mbBundleRunningIsRecent) mbBundleConverged_Recent = false;
    mbBundleConverged_Full = false;
    for (auto& mb : mmBase_BundleID) {
      MultiKeyFrame& mkf = *mb.first;
      mkf.mse3BaseFromWorld = multiBundle.GetPose(mb.second);
      for (auto& kv : mkf.mmpKeyFrames) kv.second->mse3CamFromWorld = kv.second->mse3CamFromBase * mkf.mse3BaseFromWorld;
    }
    for (auto& pb : mmPoint_BundleID) {
      MapPoint& point = *pb.first;
      const Vector<3> v3Pos = multiBundle.GetPoint(pb.second);
      point.mv3WorldPos = point.mbFixed ? v3Pos : point.mpPatchSourceKF->mse3CamFromWorld.inverse() * v3Pos;
      point.RefreshPixelVectors();
      point.mbOptimized = true;
    }
    for (auto& mb : mmBase_BundleID) mb.first->RefreshSceneDepthRobust();
    mdSigmaSquared = multiBundle.GetSigmaSquared();
    mdMeanChiSquared = multiBundle.GetMeanChiSquared();
    mdMaxCov = multiBundle.GetMaxCov();
    if (mUpdateCallback) mUpdateCallback(spAdjustSet, spMapPoints);
  }
  return nAccepted;
}

}  // namespace mcp_host
