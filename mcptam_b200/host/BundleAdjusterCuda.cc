#include "BundleAdjusterCuda.h"

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
      mmBase_BundleID[&mkf] = id; mmBundleID_Base[id] = &mkf;
      for (auto& kv : mkf.mmpKeyFrames)
        if (!mmCamName_BundleID.count(kv.first)) mmCamName_BundleID[kv.first] = multiBundle.AddPose(kv.second->mse3CamFromBase, true);
    }
  }
  int nWorldID = -1;
  for (MapPoint* pp : spMapPoints) {
    MapPoint& point = *pp;
    Vector<3> v3Pos;
    std::vector<int> vPoses;
    if (point.mbFixed) {
      if (nWorldID == -1) nWorldID = multiBundle.AddPose(SE3(), true);
      v3Pos = point.mv3WorldPos;
      vPoses.push_back(nWorldID);
    } else {
      v3Pos = point.mpPatchSourceKF->mse3CamFromWorld * point.mv3WorldPos;
      vPoses.push_back(mmBase_BundleID[point.mpPatchSourceKF->mpParent]);
      vPoses.push_back(mmCamName_BundleID[point.mpPatchSourceKF->mCamName]);
    }
    const int id = multiBundle.AddPoint(v3Pos, vPoses, point.mbFixed);
    mmPoint_BundleID[&point] = id; mmBundleID_Point[id] = &point;
  }
  for (auto& mb : mmBase_BundleID) {
    MultiKeyFrame& mkf = *mb.first;
    for (auto& kv : mkf.mmpKeyFrames) {
      KeyFrame& kf = *kv.second;
      const int nBaseID = mb.second, nCamID = mmCamName_BundleID[kv.first];
      int nCamIndex = -1;                                  // resolved at the keyframe's first measurement
      for (auto& mm : kf.mmpMeasurements) {
        const auto itPoint = mmPoint_BundleID.find(mm.first);
        if (itPoint == mmPoint_BundleID.end()) continue;
        if (nCamIndex < 0) nCamIndex = multiBundle.CameraIndex(kv.first);
        const Measurement& meas = *mm.second;
        multiBundle.AddMeas(nBaseID, nCamID, itPoint->second, meas.v2RootPos, LevelScale(meas.nLevel) * LevelScale(meas.nLevel), nCamIndex);
      }
    }
  }
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
