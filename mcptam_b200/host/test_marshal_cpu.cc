// test_marshal_cpu.cc — CPU-only check and timing of BundleAdjusterCuda::Marshal (no device call is made): the flat
// arrays handed to mcp_ba_load must be exactly those of the reference-style marshalling loop
// (src/BundleAdjusterMulti.cc:83-203: std::map lookups, by-value AddMeas), at a 200 KF / 10 k-point map.
#include <chrono>
#include <cstdio>
#include <map>
#include <random>
#include <thread>
#include <cstdlib>

#include "BundleAdjusterCuda.h"

using namespace mcp_host;

struct ProbeBundle : ChainBundle {
  using ChainBundle::ChainBundle;
  using ChainBundle::mvPoseRt; using ChainBundle::mvPtXyz; using ChainBundle::mvMeasXy; using ChainBundle::mvMeasNoise;
  using ChainBundle::mvPoseFixed; using ChainBundle::mvPtFixed; using ChainBundle::mvPtChain; using ChainBundle::mvMeasChain;
  using ChainBundle::mvMeasPt; using ChainBundle::mvMeasCam; using ChainBundle::mvMeasFirstId; using ChainBundle::mvCamNames;
};
struct ProbeAdjuster : BundleAdjusterCuda {
  using BundleAdjusterCuda::BundleAdjusterCuda;
  using BundleAdjusterCuda::Marshal;
  void Reset() { mmPoint_BundleID.clear(); mmBundleID_Point.clear(); mmBase_BundleID.clear(); mmBundleID_Base.clear(); mmCamName_BundleID.clear(); }   // as BundleAdjust() does
};

static inline int LevelScale(int l) { return 1 << l; }

// the loop as the reference writes it
static void MarshalReferenceStyle(ChainBundle& multiBundle, std::set<MultiKeyFrame*>& spAdjustSet, std::set<MultiKeyFrame*>& spFixedSet,
                                  std::set<MapPoint*>& spMapPoints)
{
  std::map<MapPoint*, int> mmPoint_BundleID;
  std::map<MultiKeyFrame*, int> mmBase_BundleID;
  std::map<std::string, int> mmCamName_BundleID;
  for (int pass = 0; pass < 2; pass++) {
    std::set<MultiKeyFrame*>& s = pass == 0 ? spAdjustSet : spFixedSet;
    for (MultiKeyFrame* pm : s) {
      if (pm->mbBad) continue;
      mmBase_BundleID[pm] = multiBundle.AddPose(pm->mse3BaseFromWorld, pass == 0 ? pm->mbFixed : true);
      for (auto& kv : pm->mmpKeyFrames)
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
    mmPoint_BundleID[&point] = multiBundle.AddPoint(v3Pos, vPoses, point.mbFixed);
  }
  for (auto& mb : mmBase_BundleID)
    for (auto& kv : mb.first->mmpKeyFrames) {
      std::vector<int> vCams(2);
      vCams[0] = mb.second; vCams[1] = mmCamName_BundleID[kv.first];
      for (auto& mm : kv.second->mmpMeasurements) {
        if (!mmPoint_BundleID.count(mm.first)) continue;
        multiBundle.AddMeas(vCams, mmPoint_BundleID[mm.first], mm.second->v2RootPos, LevelScale(mm.second->nLevel) * LevelScale(mm.second->nLevel), kv.first);
      }
    }
}

int main()
{
  TaylorCameraMap cams;
  Vector<9> p;
  const double params[9] = { 250.0, -1.2e-3, 6.0e-7, -1.0e-9, 320.0, 240.0, 1.0, 0.0, 0.0 };
  for (int i = 0; i < 9; i++) p[i] = params[i];
  const char* names[] = { "camera1", "camera2", "camera3", "camera4" };
  for (const char* n : names) cams[n] = TaylorCamera(p, ImageRef(640, 480), ImageRef(640, 480), ImageRef(640, 480));
  std::mt19937 rng(3);
  const int nMKF = 50, nPts = 10000, nTrack = 8;
  std::vector<MultiKeyFrame*> mkfs;
  for (int m = 0; m < nMKF; m++) {
    MultiKeyFrame* mkf = new MultiKeyFrame;
    mkf->mbFixed = (m == 0);
    Vector<6> mu;
    mu[0] = 0.1 * m; mu[1] = 0.01 * m; mu[2] = 0; mu[3] = 0.01 * m; mu[4] = 0.02; mu[5] = 0;
    mkf->mse3BaseFromWorld = SE3::exp(mu);
    for (int c = 0; c < 4; c++) {
      KeyFrame* kf = new KeyFrame;
      kf->mCamName = names[c]; kf->mpParent = mkf;
      Vector<6> ce;
      ce[0] = 0.1 * c; ce[1] = 0; ce[2] = 0; ce[3] = 0; ce[4] = 1.57 * c; ce[5] = 0;
      kf->mse3CamFromBase = SE3::exp(ce);
      kf->mse3CamFromWorld = kf->mse3CamFromBase * mkf->mse3BaseFromWorld;
      mkf->mmpKeyFrames[kf->mCamName] = kf;
    }
    mkfs.push_back(mkf);
  }
  std::vector<MapPoint*> pts;
  size_t nMeas = 0;
  for (int i = 0; i < nPts; i++) {
    MapPoint* pt = new MapPoint;
    pt->mv3WorldPos = makeVector((rng() % 2000) * 0.01 - 10, (rng() % 2000) * 0.01 - 10, 3 + (rng() % 900) * 0.01);
    pt->mbFixed = (i % 997 == 0);
    pt->mpPatchSourceKF = mkfs[rng() % nMKF]->mmpKeyFrames[names[rng() % 4]];
    for (int k = 0; k < nTrack; k++) {
      KeyFrame* kf = mkfs[rng() % nMKF]->mmpKeyFrames[names[rng() % 4]];
      if (kf->mmpMeasurements.count(pt)) continue;
      Measurement* m = new Measurement;
      m->nLevel = rng() % 4;
      m->v2RootPos = makeVector((rng() % 64000) * 0.01, (rng() % 48000) * 0.01);
      kf->mmpMeasurements[pt] = m;
      pt->mMMData.spMeasurementKFs.insert(kf);
      nMeas++;
    }
    pts.push_back(pt);
  }
  std::set<MultiKeyFrame*> adjust(mkfs.begin() + 2, mkfs.end()), fixed(mkfs.begin(), mkfs.begin() + 2);
  std::set<MapPoint*> points(pts.begin(), pts.end() - 50);         // some measured points are not part of this adjustment

  auto ms = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count(); };
  double tRef = 1e30, tNew = 1e30;
  int fail = 0;
  for (int rep = 0; rep < 5; rep++) {
    ProbeBundle a(cams, true, true, false), b(cams, true, true, false);
    auto t0 = std::chrono::steady_clock::now();
    MarshalReferenceStyle(a, adjust, fixed, points);
    tRef = std::min(tRef, ms(t0));
    ProbeAdjuster adj(cams);
    t0 = std::chrono::steady_clock::now();
    adj.Marshal(b, adjust, fixed, points);
    tNew = std::min(tNew, ms(t0));
    fail += !(a.mvPoseRt == b.mvPoseRt && a.mvPoseFixed == b.mvPoseFixed && a.mvPtXyz == b.mvPtXyz && a.mvPtChain == b.mvPtChain &&
              a.mvPtFixed == b.mvPtFixed && a.mvMeasXy == b.mvMeasXy && a.mvMeasChain == b.mvMeasChain && a.mvMeasPt == b.mvMeasPt &&
              a.mvMeasNoise == b.mvMeasNoise && a.mvMeasCam == b.mvMeasCam && a.mvMeasFirstId == b.mvMeasFirstId && a.mvCamNames == b.mvCamNames);
    if (rep == 0) std::printf("marshalled %zu poses, %zu points, %zu measurements (map holds %zu)\n", b.mvPoseFixed.size(), b.mvPtFixed.size(), b.mvMeasPt.size(), nMeas);
  }
  {
    // the worker pool under churn: 150 more calls with the thread count changing from call to call (pool growth, threads that
    // sit out a pass, wake-ups from sleep), one adapter object as in production
    ProbeBundle ref(cams, true, true, false);
    MarshalReferenceStyle(ref, adjust, fixed, points);
    ProbeAdjuster adj(cams);
    for (int rep = 0; rep < 150; rep++) {
      char buf[8];
      std::snprintf(buf, sizeof(buf), "%d", 1 + (rep * 5) % 8);
      setenv("MCP_HOST_THREADS", buf, 1);
      ProbeBundle b(cams, true, true, false);
      adj.Reset();
      adj.Marshal(b, adjust, fixed, points);
      fail += !(ref.mvMeasXy == b.mvMeasXy && ref.mvMeasChain == b.mvMeasChain && ref.mvMeasPt == b.mvMeasPt && ref.mvMeasNoise == b.mvMeasNoise &&
                ref.mvMeasCam == b.mvMeasCam && ref.mvMeasFirstId == b.mvMeasFirstId && ref.mvPtXyz == b.mvPtXyz && ref.mvPtChain == b.mvPtChain);
      if (rep % 37 == 0) std::this_thread::sleep_for(std::chrono::milliseconds(2));      // lets the workers fall asleep
    }
    unsetenv("MCP_HOST_THREADS");
  }
  std::printf("reference-style %.2f ms, Marshal %.2f ms\n", tRef, tNew);
  std::printf(fail ? "MARSHAL_TEST FAILED\n" : "MARSHAL_TEST OK\n");
  return fail ? 1 : 0;
}
