// test_host.cc — exercises the C++ host mirror end to end on a small synthetic map (needs a GPU):
//   TaylorCamera fit, BundleAdjusterCuda::BundleAdjust over MultiKeyFrame/KeyFrame/MapPoint shim types,
//   MakeKeyFrame_Lite + SearchForPoints over the front-end device.  Prints "HOST_TEST OK" on success.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <random>

#include "BundleAdjusterCuda.h"
#include "FrontEnd.h"

using namespace mcp_host;

static SE3 RotZ(double a, Vector<3> t) { SE3 r = SE3::exp(Vector<6>()); Vector<3> w = makeVector(0, 0, a); r.rot = SO3::exp(w); r.trans = t; return r; }

int main()
{
  std::mt19937 rng(3);
  std::normal_distribution<double> N(0, 1);
  std::uniform_real_distribution<double> U(0, 1);
  // ---- camera fit ---------------------------------------------------------------------------------
  Vector<9> p;
  p[0] = 250; p[1] = -1.2e-3; p[2] = 6e-7; p[3] = -1e-9; p[4] = 321; p[5] = 239; p[6] = 1.0; p[7] = 0; p[8] = 0;
  TaylorCameraMap cams;
  const char* names[2] = { "camera1", "camera2" };
  for (auto n : names) cams[n] = TaylorCamera(p, ImageRef(640, 480), ImageRef(640, 480), ImageRef(640, 480));
  if (!cams["camera1"].Good()) { std::printf("camera fit failed\n"); return 1; }
  double worst = 0;
  for (int i = 0; i < 200; i++) {
    Vector<2> px = makeVector(40 + 560 * U(rng), 40 + 400 * U(rng));
    Vector<3> ray = cams["camera1"].UnProject(px);
    Vector<2> back = cams["camera1"].Project(ray * 4.2);
    worst = std::max(worst, std::hypot(back[0] - px[0], back[1] - px[1]));
  }
  std::printf("camera: inverse poly degree %d, project(unproject) max err %.2e px\n", (int)cams["camera1"].InvPoly().size() - 1, worst);
  if (worst > 2e-4) return 1;

  // ---- map: 2 cameras looking +x / -x, 6 MKFs on a line, 400 points ----------------------------------
  SE3 camFromBase[2];
  for (int c = 0; c < 2; c++) {
    // base: x forward, z up.  camera z = optical axis
    const double yaw = c * M_PI;
    Matrix<3> Rb;   // base_from_cam columns: right, -up, fwd
    const Vector<3> fwd = makeVector(std::cos(yaw), std::sin(yaw), 0), up = makeVector(0, 0, 1), right = fwd ^ up;
    for (int i = 0; i < 3; i++) { Rb[i][0] = right[i]; Rb[i][1] = -up[i]; Rb[i][2] = fwd[i]; }
    SE3 baseFromCam; baseFromCam.rot.R = Rb; baseFromCam.trans = fwd * 0.1;
    camFromBase[c] = baseFromCam.inverse();
  }
  const int M = 6, P = 400;
  std::vector<std::unique_ptr<MultiKeyFrame> > mkfs;
  std::vector<std::unique_ptr<KeyFrame> > kfs;
  std::vector<SE3> truth;
  for (int m = 0; m < M; m++) {
    mkfs.emplace_back(new MultiKeyFrame);
    SE3 worldFromBase = RotZ(0.05 * m, makeVector(0.1 * m, 0.5 * m, 0.05 * m));   // sideways motion: parallax for both cameras
    truth.push_back(worldFromBase.inverse());
    mkfs[m]->mse3BaseFromWorld = truth[m];
    mkfs[m]->mbFixed = (m == 0 || m == M - 1);      // two fixed MKFs pin the scale gauge (0.2 m rig baseline alone is weak)
    for (int c = 0; c < 2; c++) {
      kfs.emplace_back(new KeyFrame);
      KeyFrame* kf = kfs.back().get();
      kf->mCamName = names[c]; kf->mpParent = mkfs[m].get(); kf->mse3CamFromBase = camFromBase[c];
      kf->mse3CamFromWorld = camFromBase[c] * truth[m];
      mkfs[m]->mmpKeyFrames[names[c]] = kf;
    }
  }
  std::vector<std::unique_ptr<MapPoint> > pts;
  std::vector<std::unique_ptr<Measurement> > meas;
  std::vector<Vector<3> > ptTruth;
  for (int i = 0; i < P; i++) {
    const double side = (i & 1) ? 1.0 : -1.0;
    Vector<3> w = makeVector(side * (4 + 5 * U(rng)), -3 + 8 * U(rng), -2 + 4 * U(rng));
    std::unique_ptr<MapPoint> mp(new MapPoint);
    mp->mv3WorldPos = w;
    int n_obs = 0;
    for (int m = 0; m < M; m++)
      for (int c = 0; c < 2; c++) {
        KeyFrame* kf = mkfs[m]->mmpKeyFrames[names[c]];
        Vector<2> px = cams[names[c]].Project(kf->mse3CamFromWorld * w);
        if (cams[names[c]].Invalid() || px[0] < 10 || px[1] < 10 || px[0] > 630 || px[1] > 470) continue;
        if (!mp->mpPatchSourceKF) mp->mpPatchSourceKF = kf;
        meas.emplace_back(new Measurement);
        meas.back()->nLevel = (int)(4 * U(rng)) & 3;
        meas.back()->v2RootPos = px + makeVector(0.3 * N(rng), 0.3 * N(rng)) * (1 << meas.back()->nLevel);
        kf->mmpMeasurements[mp.get()] = meas.back().get();
        mp->mMMData.spMeasurementKFs.insert(kf);
        n_obs++;
      }
    if (n_obs < 2) { for (auto& k : kfs) k->mmpMeasurements.erase(mp.get()); continue; }
    ptTruth.push_back(w);
    pts.push_back(std::move(mp));
  }
  // perturb the estimate
  for (int m = 1; m < M - 1; m++) {
    Vector<6> d;
    for (int k = 0; k < 3; k++) { d[k] = 0.02 * N(rng); d[3 + k] = 0.008 * N(rng); }
    mkfs[m]->mse3BaseFromWorld = SE3::exp(d) * mkfs[m]->mse3BaseFromWorld;
    for (auto& kv : mkfs[m]->mmpKeyFrames) kv.second->mse3CamFromWorld = kv.second->mse3CamFromBase * mkfs[m]->mse3BaseFromWorld;
  }
  for (auto& mp : pts) mp->mv3WorldPos = mp->mv3WorldPos + makeVector(0.05 * N(rng), 0.05 * N(rng), 0.05 * N(rng));

  BundleAdjusterCuda ba(cams);
  std::set<MultiKeyFrame*> adj, fixed;
  std::set<MapPoint*> sp;
  for (auto& m : mkfs) adj.insert(m.get());
  for (auto& q : pts) sp.insert(q.get());
  std::vector<std::pair<KeyFrame*, MapPoint*> > outliers;
  int nCallbacks = 0;
  size_t nCallbackPoints = 0;
  ba.SetUpdateCallback([&](std::set<MultiKeyFrame*> a, std::set<MapPoint*> p) { nCallbacks++; nCallbackPoints = p.size(); (void)a; });
  const int n = ba.BundleAdjust(adj, fixed, sp, outliers, false);
  double pose_err = 0, pt_err = 0;
  for (int m = 0; m < M; m++) for (int k = 0; k < 3; k++) pose_err = std::max(pose_err, std::fabs(mkfs[m]->mse3BaseFromWorld.trans[k] - truth[m].trans[k]));
  std::vector<double> errs;
  for (size_t i = 0; i < pts.size(); i++) { Vector<3> d = pts[i]->mv3WorldPos - ptTruth[i]; errs.push_back(std::sqrt(d * d)); }
  std::sort(errs.begin(), errs.end());
  pt_err = errs[errs.size() / 2];                      // median: single low-parallax points may legitimately drift
  std::printf("BundleAdjusterCuda: %zu points, %zu meas, accepted %d, total trials %d, converged %d, sigma^2 %.3f, outliers %zu, pose err %.4f m, median point err %.3f m, gpu %.2f ms\n",
              pts.size(), meas.size(), n, ba.TotalIterations(), (int)ba.ConvergedFull(), ba.GetSigmaSquared(), outliers.size(), pose_err, pt_err, ba.LastGpuMs());
  if (n <= 0 || pose_err > 0.02 || pt_err > 0.05) return 1;
  // the update callback fires after every successful update: once per pass of the two-step adjuster (src/BundleAdjusterMulti.cc:332-333)
  std::printf("update callback: %d calls, %zu points\n", nCallbacks, nCallbackPoints);
  if (nCallbacks < 1 || nCallbacks > 2 || nCallbackPoints != sp.size()) return 1;

  // ---- front end --------------------------------------------------------------------------------------
  FrontEndDevice dev(640, 480);
  Image<byte> imA(ImageRef(640, 480)), imB(ImageRef(640, 480));
  for (int y = 0; y < 480; y++)
    for (int x = 0; x < 640; x++) {
      auto tex = [](int xx, int yy) { return (byte)(((xx / 16 + yy / 12) & 1) ? 200 - (xx * 7 + yy * 3) % 40 : 40 + (xx * 5 + yy * 11) % 50); };
      imA[ImageRef(x, y)] = tex(x, y);
      imB[ImageRef(x, y)] = tex(x + 2, y + 1);          // scene shifted by (-2, -1)
    }
  KeyFrame kfA, kfB;
  MakeKeyFrame_Lite(dev, kfA, 0, imA);
  MakeKeyFrame_Lite(dev, kfB, 1, imB);
  std::printf("MakeKeyFrame_Lite: corners per level %zu %zu %zu %zu, thresholds %d %d %d %d\n", kfA.maLevels[0].vCorners.size(), kfA.maLevels[1].vCorners.size(),
              kfA.maLevels[2].vCorners.size(), kfA.maLevels[3].vCorners.size(), kfA.maLevels[0].nFastThresh, kfA.maLevels[1].nFastThresh,
              kfA.maLevels[2].nFastThresh, kfA.maLevels[3].nFastThresh);
  if (kfA.maLevels[0].vCorners.size() < 100 || (int)kfA.maLevels[0].vCornerRowLUT.size() != 480) return 1;
  std::vector<std::unique_ptr<MapPoint> > fpts;
  std::vector<std::unique_ptr<TrackerData> > tds;
  std::vector<TrackerData*> vTD;
  for (size_t i = 0; i < kfA.maLevels[0].vCorners.size() && vTD.size() < 300; i += 3) {
    const ImageRef c = kfA.maLevels[0].vCorners[i];
    if (c.x < 20 || c.y < 20 || c.x > 620 || c.y > 460) continue;
    fpts.emplace_back(new MapPoint);
    fpts.back()->mpPatchSourceKF = &kfA; fpts.back()->mnSourceLevel = 0; fpts.back()->mirCenter = c;
    tds.emplace_back(new TrackerData);
    TrackerData& td = *tds.back();
    td.mpPoint = fpts.back().get();
    td.mm2WarpInverse[0][0] = 1; td.mm2WarpInverse[1][1] = 1; td.mnSearchLevel = 0; td.mbTemplateBad = false;
    td.mv2Image = makeVector(c.x - 2 + 1, c.y - 1 - 1);
    vTD.push_back(&td);
  }
  int att[LEVELS] = { 0, 0, 0, 0 }, fnd[LEVELS] = { 0, 0, 0, 0 };
  const int nf = SearchForPoints(dev, vTD, kfB, 10, 8, false, att, fnd);
  double med = 0;
  int cnt = 0;
  for (auto* td : vTD) if (td->mbFound) { med += std::hypot(td->mv2Found[0] - (td->mpPoint->mirCenter.x - 2), td->mv2Found[1] - (td->mpPoint->mirCenter.y - 1)); cnt++; }
  std::printf("SearchForPoints: %d of %zu found (attempted L0 %d), mean position error %.3f px\n", nf, vTD.size(), att[0], cnt ? med / cnt : -1.0);
  if (nf < (int)vTD.size() / 2 || med / std::max(cnt, 1) > 0.5) return 1;
  // ---- KeyFrame::MakeKeyFrame_Lite with the reference's member signature: rings, lastMask, glare masking -------------
  {
    KeyFrame kf;
    kf.AttachDevice(&dev, 2);                           // slots 2..4 of the device (0, 1 are used above)
    Image<byte> imC = imA;
    for (int y = 200; y < 206; y++) for (int x = 300; x < 330; x++) imC[ImageRef(x, y)] = 255;     // a glare blob
    kf.MakeKeyFrame_Lite(imA);
    const std::vector<ImageRef> cornersA = kf.maLevels[0].vCorners;
    if (kf.maLevels[0].imagePrev.size() != 0 || kf.nDeviceSlot != 2 || kf.PrevDeviceSlot(1) != -1) { std::printf("ring: first frame\n"); return 1; }
    if (cornersA.size() != kfA.maLevels[0].vCorners.size()) { std::printf("member and free function disagree\n"); return 1; }
    kf.MakeKeyFrame_Lite(imB);
    const std::vector<ImageRef> cornersB = kf.maLevels[0].vCorners;
    kf.MakeKeyFrame_Lite(imC, false, true);
    bool ok = kf.maLevels[0].imagePrev.size() == 2 && kf.maLevels[3].imagePrev.size() == 2 && kf.maLevels[0].vCornersPrev.size() == 2;
    ok = ok && kf.maLevels[0].vCornersPrev[0].size() == cornersA.size() && kf.maLevels[0].vCornersPrev[1].size() == cornersB.size();
    ok = ok && std::memcmp(kf.maLevels[0].imagePrev[0].data(), imA.data(), 640 * 480) == 0 && std::memcmp(kf.maLevels[0].imagePrev[1].data(), imB.data(), 640 * 480) == 0;
    ok = ok && kf.nDeviceSlot == 4 && kf.PrevDeviceSlot(1) == 3 && kf.PrevDeviceSlot(2) == 2;
    int nMasked = 0;
    for (int i = 0; i < 640 * 480; i++) nMasked += kf.maLevels[0].lastMask.data()[i] == 0;
    bool cornerInMask = false;
    for (const ImageRef& c : kf.maLevels[0].vCorners) cornerInMask = cornerInMask || kf.maLevels[0].lastMask[c] != 255;
    std::printf("KeyFrame::MakeKeyFrame_Lite: rings %zu/%zu, glare-masked pixels %d, corners %zu (no corner inside the mask: %d)\n",
                kf.maLevels[0].imagePrev.size(), kf.maLevels[0].vCornersPrev.size(), nMasked, kf.maLevels[0].vCorners.size(), (int)!cornerInMask);
    ok = ok && nMasked > 30 * 6 && nMasked < 80 * 40 && !cornerInMask;
    kf.MakeKeyFrame_Lite(imA);                            // a fourth frame: the rings stay at snNumPrev, the oldest slot is reused
    ok = ok && kf.maLevels[0].imagePrev.size() == 2 && kf.nDeviceSlot == 2 && kf.PrevDeviceSlot(1) == 4 && kf.PrevDeviceSlot(2) == 3;
    ok = ok && std::memcmp(kf.maLevels[0].imagePrev[0].data(), imB.data(), 640 * 480) == 0;
    for (int i = 0; i < 640 * 480 && ok; i++) ok = kf.maLevels[0].lastMask.data()[i] == 255;
    if (!ok) { std::printf("KeyFrame::MakeKeyFrame_Lite boundary state wrong\n"); return 1; }
  }
  std::printf("HOST_TEST OK\n");
  return 0;
}
