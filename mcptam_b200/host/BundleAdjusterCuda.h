// BundleAdjusterCuda.h — drop-in for BundleAdjusterMulti (include/mcptam/BundleAdjusterMulti.h:62-113):
// subclasses the BundleAdjusterBase call surface (include/mcptam/BundleAdjusterBase.h:110-128) and marshals the
// map exactly like src/BundleAdjusterMulti.cc:55-337, but the ChainBundle it drives runs on the B200.
#pragma once

#include <functional>
#include <set>
#include <unordered_map>
#include <utility>
#include <vector>

#include "ChainBundle.h"
#include "shim/MapTypes.h"

namespace mcp_host {

class BundleAdjusterBase {       // the members MapMaker relies on (include/mcptam/BundleAdjusterBase.h)
 public:
  virtual ~BundleAdjusterBase() {}
  void RequestAbort() { mbBundleAbortRequested = true; }
  bool Running() const { return mbBundleRunning; }
  bool ConvergedFull() const { return mbBundleConverged_Full; }
  bool ConvergedRecent() const { return mbBundleConverged_Recent; }
  void UseTukey(bool b) { mbUseTukey = b; }
  void UseTwoStep(bool b) { mbUseTwoStep = b; }
  void UseRobust(bool b) { mbUseRobust = b; }
  double GetSigmaSquared() const { return mdSigmaSquared; }
  double GetMeanChiSquared() const { return mdMeanChiSquared; }
  double GetMaxCov() const { return mdMaxCov; }
  int TotalIterations() const { return mnTotalIterations; }
  static int snMinMapPoints;       // src/BundleAdjusterBase.cc:49
  virtual int BundleAdjust(std::set<MultiKeyFrame*> spAdjustSet, std::set<MultiKeyFrame*> spFixedSet, std::set<MapPoint*> spMapPoints,
                           std::vector<std::pair<KeyFrame*, MapPoint*> >& vOutliers, bool bRecent) = 0;

 protected:
  bool mbBundleRunning = false, mbBundleRunningIsRecent = false, mbBundleConverged_Full = false, mbBundleConverged_Recent = false;
  bool mbBundleAbortRequested = false, mbUseTukey = true, mbUseTwoStep = true, mbUseRobust = true, mbApplyUpdates = true, mbVerbose = false;
  double mdSigmaSquared = 0, mdMeanChiSquared = 0, mdMaxCov = 0;
  int mnTotalIterations = 0;
};

class BundleAdjusterCuda : public BundleAdjusterBase {
 public:
  explicit BundleAdjusterCuda(TaylorCameraMap& cameras) : mmCameraModels(cameras) {}
  int BundleAdjust(std::set<MultiKeyFrame*> spAdjustSet, std::set<MultiKeyFrame*> spFixedSet, std::set<MapPoint*> spMapPoints,
                   std::vector<std::pair<KeyFrame*, MapPoint*> >& vOutliers, bool bRecent) override;
  double LastGpuMs() const { return mdGpuMs; }
  // include/mcptam/BundleAdjusterMulti.h:69,86: called after every successful update of the map (src/BundleAdjusterMulti.cc:332-333),
  // i.e. after each of the two Compute passes of the two-step adjuster -- the network MapMaker uses it to push updates
  typedef std::function<void(std::set<MultiKeyFrame*>, std::set<MapPoint*>)> UpdateCallbackType;
  void SetUpdateCallback(UpdateCallbackType up) { mUpdateCallback = up; }

 protected:
  // The marshalling half of BundleAdjust (src/BundleAdjusterMulti.cc:83-203): poses, points and measurements of the map
  // into the bundle, filling the id maps.  Same AddPose / AddPoint order as the reference and the same measurement arrays as
  // its loop produces; the keyframes' measurement maps are walked by a few pooled threads (Workers.h) into per-keyframe
  // buffers that are then placed in the reference's order, lookups go through a hash map, and the flat arrays / buffers are
  // reused across calls (at 80 k measurements the reference-style loop costs ~45 ms on one core -- ten times the bundle
  // adjustment itself on the device; this path 7 ms).
  void Marshal(ChainBundle& multiBundle, std::set<MultiKeyFrame*>& spAdjustSet, std::set<MultiKeyFrame*>& spFixedSet, std::set<MapPoint*>& spMapPoints);
  int AdjustAndUpdate(ChainBundle& multiBundle, std::set<MultiKeyFrame*> spAdjustSet, std::set<MapPoint*> spMapPoints, int nIterations = -1);
  TaylorCameraMap& mmCameraModels;
  std::unordered_map<MapPoint*, int> mmPoint_BundleID;
  std::vector<MapPoint*> mmBundleID_Point;              // indexed by bundle id (ids are handed out consecutively)
  std::map<MultiKeyFrame*, int> mmBase_BundleID;
  std::vector<MultiKeyFrame*> mmBundleID_Base;
  std::map<std::string, int> mmCamName_BundleID;
  // per-keyframe gather buffers of Marshal, kept across calls (capacity reuse)
  struct KfJob { KeyFrame* kf; const std::string* name; int nBaseID, nCamID; std::vector<int> ids; std::vector<double> xy, noise; };
  std::vector<KfJob> mvJobs;
  double mdGpuMs = 0;
  UpdateCallbackType mUpdateCallback;
};

}  // namespace mcp_host
