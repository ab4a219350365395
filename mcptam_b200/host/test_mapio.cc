// test_mapio.cc — CPU-only self test of MapIO (no CUDA): builds a synthetic multi-camera map, writes the reference's
// dump format, reads it back, and checks the round trip.  usage: test_mapio <output dump path>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <string>

#include "MapIO.h"

using namespace mcp_host;

static int g_fail = 0;
#define CHECK(c) do { if (!(c)) { printf("CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); g_fail++; } } while (0)

static double rel(double a, double b) { return std::fabs(a - b) / std::max(1.0, std::fabs(b)); }

static SE3 MakePose(double ax, double ay, double az, double tx, double ty, double tz)
{
  Vector<6> mu;
  mu[0] = tx; mu[1] = ty; mu[2] = tz; mu[3] = ax; mu[4] = ay; mu[5] = az;
  return SE3::exp(mu);
}

static void BuildMap(Map& map, int n_cam, int n_mkf, int n_pt)
{
  const char* names[] = { "camera1", "camera2", "camera3", "camera4" };
  for (int m = 0; m < n_mkf; m++) {
    MultiKeyFrame* mkf = new MultiKeyFrame;
    mkf->mbFixed = (m == 0);
    mkf->mse3BaseFromWorld = MakePose(0.1 * m, -0.05 * m, 0.3 * m, 0.5 * m, 0.1 * m * m, -0.2 * m);
    for (int c = 0; c < n_cam; c++) {
      KeyFrame* kf = new KeyFrame;
      kf->mCamName = names[c];
      kf->mpParent = mkf;
      kf->mse3CamFromBase = MakePose(0, 2.0 * M_PI * c / n_cam * 0.999, 0.01 * c, 0.1 * std::cos(c), 0.0, 0.1 * std::sin(c));   // includes rotations with trace < 0
      kf->mse3CamFromWorld = kf->mse3CamFromBase * mkf->mse3BaseFromWorld;
      mkf->mmpKeyFrames[kf->mCamName] = kf;
    }
    map.mlpMultiKeyFrames.push_back(mkf);
  }
  std::vector<MultiKeyFrame*> mk(map.mlpMultiKeyFrames.begin(), map.mlpMultiKeyFrames.end());
  for (int p = 0; p < n_pt; p++) {
    MapPoint* pt = new MapPoint;
    pt->mv3WorldPos = makeVector(3.0 * std::sin(0.7 * p), 2.0 * std::cos(1.3 * p), 4.0 + 0.01 * p);
    pt->mpPatchSourceKF = mk[p % n_mkf]->mmpKeyFrames[names[p % n_cam]];
    map.mlpPoints.push_back(pt);
    for (int m = 0; m < n_mkf; m++) {
      if ((p + m) % 3 == 2) continue;
      KeyFrame* kf = mk[m]->mmpKeyFrames[names[(p + m) % n_cam]];
      Measurement* meas = new Measurement;
      meas->nLevel = (p + 2 * m) % LEVELS;
      meas->v2RootPos = makeVector(320.0 + 1.37 * p - 3.0 * m, 240.0 - 0.77 * p + 2.5 * m);
      kf->mmpMeasurements[pt] = meas;
      pt->mMMData.spMeasurementKFs.insert(kf);
    }
  }
}

static void Compare(Map& a, Map& b, double tol)
{
  CHECK(a.mlpMultiKeyFrames.size() == b.mlpMultiKeyFrames.size());
  CHECK(a.mlpPoints.size() == b.mlpPoints.size());
  auto ia = a.mlpMultiKeyFrames.begin();
  auto ib = b.mlpMultiKeyFrames.begin();
  size_t n_meas_a = 0, n_meas_b = 0;
  for (; ia != a.mlpMultiKeyFrames.end() && ib != b.mlpMultiKeyFrames.end(); ++ia, ++ib) {
    double pa[12], pb[12];
    (*ia)->mse3BaseFromWorld.pack(pa); (*ib)->mse3BaseFromWorld.pack(pb);
    for (int k = 0; k < 12; k++) CHECK(rel(pa[k], pb[k]) < tol);
    CHECK((*ia)->mmpKeyFrames.size() == (*ib)->mmpKeyFrames.size());
    for (auto& kv : (*ia)->mmpKeyFrames) {
      CHECK((*ib)->mmpKeyFrames.count(kv.first) == 1);
      KeyFrame* ka = kv.second;
      KeyFrame* kb = (*ib)->mmpKeyFrames[kv.first];
      ka->mse3CamFromBase.pack(pa); kb->mse3CamFromBase.pack(pb);
      for (int k = 0; k < 12; k++) CHECK(rel(pa[k], pb[k]) < tol);
      ka->mse3CamFromWorld.pack(pa); kb->mse3CamFromWorld.pack(pb);
      for (int k = 0; k < 12; k++) CHECK(rel(pa[k], pb[k]) < 10 * tol);
      CHECK(ka->mmpMeasurements.size() == kb->mmpMeasurements.size());
      n_meas_a += ka->mmpMeasurements.size(); n_meas_b += kb->mmpMeasurements.size();
      // measurements are keyed by point pointer: match them through the point ids
      std::map<int, Measurement*> byid;
      for (auto& mm : kb->mmpMeasurements) byid[mm.first->mnID] = mm.second;
      for (auto& mm : ka->mmpMeasurements) {
        CHECK(byid.count(mm.first->mnID) == 1);
        if (!byid.count(mm.first->mnID)) continue;
        Measurement* mb = byid[mm.first->mnID];
        CHECK(mb->nLevel == mm.second->nLevel);
        CHECK(rel(mb->v2RootPos[0], mm.second->v2RootPos[0]) < tol && rel(mb->v2RootPos[1], mm.second->v2RootPos[1]) < tol);
      }
    }
  }
  CHECK(n_meas_a == n_meas_b && n_meas_a > 0);
  auto pa = a.mlpPoints.begin();
  auto pb = b.mlpPoints.begin();
  for (; pa != a.mlpPoints.end() && pb != b.mlpPoints.end(); ++pa, ++pb) {
    for (int k = 0; k < 3; k++) CHECK(rel((*pa)->mv3WorldPos[k], (*pb)->mv3WorldPos[k]) < tol);
    CHECK((*pa)->mpPatchSourceKF->mCamName == (*pb)->mpPatchSourceKF->mCamName);
    CHECK((*pa)->mpPatchSourceKF->mpParent->mnID == (*pb)->mpPatchSourceKF->mpParent->mnID);
    CHECK((*pa)->mMMData.spMeasurementKFs.size() == (*pb)->mMMData.spMeasurementKFs.size());
  }
}

int main(int argc, char** argv)
{
  const std::string path = argc > 1 ? argv[1] : "/tmp/mcptam_b200_mapio_test.txt";
  // quaternion conversions: every branch of getRotation, round trip to 1e-14
  for (int k = 0; k < 200; k++) {
    const SE3 T = MakePose(3.1 * std::sin(1.1 * k), 3.1 * std::cos(0.7 * k), 3.1 * std::sin(0.3 * k + 1), 0, 0, 0);
    double q[4];
    RotationToQuaternion(T.get_rotation().get_matrix(), q);
    CHECK(std::fabs(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3] - 1.0) < 1e-12);
    Matrix<3> R;
    QuaternionToRotation(q, R);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) CHECK(std::fabs(R[i][j] - T.get_rotation().get_matrix()[i][j]) < 1e-13);
  }
  Map map, loaded, again;
  BuildMap(map, 3, 6, 40);
  CHECK(DumpToFile(map, path));
  std::string err;
  CHECK(LoadFromFile(path, loaded, &err));
  if (!err.empty()) printf("load error: %s\n", err.c_str());
  Compare(map, loaded, 2e-5);                 // the dump keeps 6 significant digits
  CHECK((*loaded.mlpMultiKeyFrames.begin())->mbFixed);
  // dump of the loaded map == the same records again (up to the last printed digit)
  CHECK(DumpToFile(loaded, path + ".2"));
  CHECK(LoadFromFile(path + ".2", again, &err));
  Compare(loaded, again, 2e-5);
  // malformed input is rejected with a line number
  {
    std::ofstream bad((path + ".bad").c_str());
    bad << "% header\n2\ncamera1, 0, 0, 0, 0, 0, 0, 1\ncamera2, 0, 0, 0, 0, 0, 0\n";
    bad.close();
    Map m;
    std::string e;
    CHECK(!LoadFromFile(path + ".bad", m, &e));
    CHECK(e.find("line 4") != std::string::npos);
    CHECK(m.mlpMultiKeyFrames.empty());
    CHECK(!LoadFromFile(path + ".does_not_exist", m, &e));
  }
  FreeMap(map); FreeMap(loaded); FreeMap(again);
  printf(g_fail ? "MAPIO_TEST FAILED (%d)\n" : "MAPIO_TEST OK\n", g_fail);
  return g_fail ? 1 : 0;
}
