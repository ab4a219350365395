#include "MapIO.h"

#include <cmath>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <vector>

namespace mcp_host {

// [3P] Bullet / tf Matrix3x3::getRotation
void RotationToQuaternion(const Matrix<3>& R, double q[4])
{
  const double trace = R[0][0] + R[1][1] + R[2][2];
  if (trace > 0.0) {
    double s = std::sqrt(trace + 1.0);
    q[3] = s * 0.5;
    s = 0.5 / s;
    q[0] = (R[2][1] - R[1][2]) * s;
    q[1] = (R[0][2] - R[2][0]) * s;
    q[2] = (R[1][0] - R[0][1]) * s;
  } else {
    const int i = R[0][0] < R[1][1] ? (R[1][1] < R[2][2] ? 2 : 1) : (R[0][0] < R[2][2] ? 2 : 0);
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    double s = std::sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0);
    q[i] = s * 0.5;
    s = 0.5 / s;
    q[3] = (R[k][j] - R[j][k]) * s;
    q[j] = (R[j][i] + R[i][j]) * s;
    q[k] = (R[k][i] + R[i][k]) * s;
  }
}

// [3P] Bullet / tf Matrix3x3::setRotation
void QuaternionToRotation(const double q[4], Matrix<3>& R)
{
  const double d = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  const double s = 2.0 / d;
  const double xs = q[0] * s, ys = q[1] * s, zs = q[2] * s;
  const double wx = q[3] * xs, wy = q[3] * ys, wz = q[3] * zs;
  const double xx = q[0] * xs, xy = q[0] * ys, xz = q[0] * zs;
  const double yy = q[1] * ys, yz = q[1] * zs, zz = q[2] * zs;
  R[0][0] = 1.0 - (yy + zz); R[0][1] = xy - wz; R[0][2] = xz + wy;
  R[1][0] = xy + wz; R[1][1] = 1.0 - (xx + zz); R[1][2] = yz - wx;
  R[2][0] = xz - wy; R[2][1] = yz + wx; R[2][2] = 1.0 - (xx + yy);
}

static inline int LevelScale(int l) { return 1 << l; }

namespace {

// one dump record tail: ", px, py, pz, qx, qy, qz, qw" of the INVERSE of a stored transform (the file keeps the
// conventional "pose of the frame" while PTAM stores frame-from-parent)
void PutInversePose(std::ostream& os, const SE3& stored)
{
  const SE3 pose = stored.inverse();
  double q[4];
  RotationToQuaternion(pose.get_rotation().get_matrix(), q);
  for (int k = 0; k < 3; k++) os << ", " << pose.get_translation()[k];
  for (int k = 0; k < 4; k++) os << ", " << q[k];
  os << std::endl;
}

// the three '%' lines that open every section of the file
void PutHeader(std::ostream& os, const char* what, const char* count, const char* record)
{
  os << "% " << what << std::endl << "% " << count << std::endl << "% " << record << std::endl;
}

}  // namespace

bool DumpToFile(Map& map, const std::string& filename)
{
  if (map.mlpMultiKeyFrames.empty()) return false;
  std::ofstream os(filename.c_str());
  if (!os.good()) return false;

  // section 1: the rig, taken from the first MKF (every MKF carries the same cameras)
  const MultiKeyFrame& first = *map.mlpMultiKeyFrames.front();
  PutHeader(os, "Camera poses in MKF frame, format:", "Total number of cameras",
            "Camera Name, Position (3 vector), Orientation (quaternion, 4 vector)");
  os << first.mmpKeyFrames.size() << std::endl;
  for (const auto& name_kf : first.mmpKeyFrames) {
    os << name_kf.second->mCamName;
    PutInversePose(os, name_kf.second->mse3CamFromBase);
  }

  // section 2: MKFs, numbered in list order (the numbering is stored in mnID, the later sections refer to it)
  PutHeader(os, "MKFs in world frame, format:", "Total number of MKFs", "MKF number, Position (3 vector), Orientation (quaternion, 4 vector)");
  os << map.mlpMultiKeyFrames.size() << std::endl;
  int id = 0;
  for (MultiKeyFrame* mkf : map.mlpMultiKeyFrames) {
    mkf->mnID = id++;
    os << mkf->mnID;
    PutInversePose(os, mkf->mse3BaseFromWorld);
  }

  // section 3: points, numbered in list order, with the keyframe their patch comes from
  PutHeader(os, "Points in world frame, format:", "Total number of points", "Point number, Position (3 vector), Parent MKF number, Parent camera name");
  os << map.mlpPoints.size() << std::endl;
  size_t n_meas = 0;
  id = 0;
  for (MapPoint* pt : map.mlpPoints) {
    pt->mnID = id++;
    os << pt->mnID;
    for (int k = 0; k < 3; k++) os << ", " << pt->mv3WorldPos[k];
    os << ", " << pt->mpPatchSourceKF->mpParent->mnID << ", " << pt->mpPatchSourceKF->mCamName << std::endl;
    n_meas += pt->mMMData.spMeasurementKFs.size();
  }

  // section 4: measurements, grouped by MKF and camera; the noise column is LevelScale(level)^2
  PutHeader(os, "Measurements of points from KeyFrames, format: ", "Total number of measurements",
            "MKF number, camera name, point number, image position (2 vector) at level 0, measurement noise");
  os << n_meas << std::endl;
  for (MultiKeyFrame* mkf : map.mlpMultiKeyFrames)
    for (const auto& name_kf : mkf->mmpKeyFrames)
      for (const auto& pt_meas : name_kf.second->mmpMeasurements) {
        const Measurement& m = *pt_meas.second;
        os << mkf->mnID << ", " << name_kf.second->mCamName << ", " << pt_meas.first->mnID << ", " << m.v2RootPos[0] << ", " << m.v2RootPos[1]
           << ", " << LevelScale(m.nLevel) * LevelScale(m.nLevel) << std::endl;
      }
  os << "% The end";
  return true;
}

namespace {

bool Fail(std::string* error, const std::string& msg, int line)
{
  if (error) { std::ostringstream o; o << "line " << line << ": " << msg; *error = o.str(); }
  return false;
}

// next non-comment, non-empty line split at commas, fields trimmed
bool NextRecord(std::ifstream& ifs, std::vector<std::string>& fields, int& line)
{
  std::string s;
  while (std::getline(ifs, s)) {
    line++;
    size_t b = s.find_first_not_of(" \t\r");
    if (b == std::string::npos || s[b] == '%') continue;
    fields.clear();
    std::stringstream ss(s);
    std::string f;
    while (std::getline(ss, f, ',')) {
      const size_t f0 = f.find_first_not_of(" \t\r"), f1 = f.find_last_not_of(" \t\r");
      fields.push_back(f0 == std::string::npos ? std::string() : f.substr(f0, f1 - f0 + 1));
    }
    return true;
  }
  return false;
}

bool ToDouble(const std::string& s, double& v)
{
  char* end = nullptr;
  v = std::strtod(s.c_str(), &end);
  return end != s.c_str() && *end == 0;
}
bool ToInt(const std::string& s, int& v)
{
  char* end = nullptr;
  const long t = std::strtol(s.c_str(), &end, 10);
  v = (int)t;
  return end != s.c_str() && *end == 0;
}

// pose record fields [first .. first+6]: position, quaternion of the INVERSE of the stored transform
bool ReadInversePose(const std::vector<std::string>& f, int first, SE3& stored)
{
  double p[3], q[4];
  for (int i = 0; i < 3; i++) if (!ToDouble(f[first + i], p[i])) return false;
  for (int i = 0; i < 4; i++) if (!ToDouble(f[first + 3 + i], q[i])) return false;
  if (q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3] < 1e-12) return false;
  SE3 pose;
  QuaternionToRotation(q, pose.get_rotation().get_matrix());
  pose.get_translation() = makeVector(p[0], p[1], p[2]);
  stored = pose.inverse();
  return true;
}

}  // namespace

bool LoadFromFile(const std::string& filename, Map& map, std::string* error)
{
  std::ifstream ifs(filename.c_str());
  int line = 0;
  if (!ifs.good()) return Fail(error, "cannot open " + filename, 0);
  std::vector<std::string> f;
  int n = 0;

  // cameras
  if (!NextRecord(ifs, f, line) || f.size() != 1 || !ToInt(f[0], n) || n <= 0) return Fail(error, "expected the number of cameras", line);
  std::vector<std::pair<std::string, SE3> > cams;
  for (int i = 0; i < n; i++) {
    SE3 camFromBase;
    if (!NextRecord(ifs, f, line) || f.size() != 8 || f[0].empty() || !ReadInversePose(f, 1, camFromBase)) return Fail(error, "bad camera record", line);
    cams.push_back(std::make_pair(f[0], camFromBase));
  }
  // MKFs
  if (!NextRecord(ifs, f, line) || f.size() != 1 || !ToInt(f[0], n) || n < 0) return Fail(error, "expected the number of MKFs", line);
  std::vector<MultiKeyFrame*> mkfs;
  for (int i = 0; i < n; i++) {
    int id = -1;
    SE3 baseFromWorld;
    if (!NextRecord(ifs, f, line) || f.size() != 8 || !ToInt(f[0], id) || id != i || !ReadInversePose(f, 1, baseFromWorld)) {
      FreeMap(map);
      return Fail(error, "bad MKF record", line);
    }
    MultiKeyFrame* mkf = new MultiKeyFrame;
    mkf->mnID = i;
    mkf->mbFixed = (i == 0);
    mkf->mse3BaseFromWorld = baseFromWorld;
    for (size_t c = 0; c < cams.size(); c++) {
      KeyFrame* kf = new KeyFrame;
      kf->mCamName = cams[c].first;
      kf->mpParent = mkf;
      kf->mse3CamFromBase = cams[c].second;
      kf->mse3CamFromWorld = kf->mse3CamFromBase * mkf->mse3BaseFromWorld;
      mkf->mmpKeyFrames[kf->mCamName] = kf;
    }
    mkfs.push_back(mkf);
    map.mlpMultiKeyFrames.push_back(mkf);
  }
  // points
  if (!NextRecord(ifs, f, line) || f.size() != 1 || !ToInt(f[0], n) || n < 0) { FreeMap(map); return Fail(error, "expected the number of points", line); }
  std::vector<MapPoint*> points;
  for (int i = 0; i < n; i++) {
    int id = -1, parent = -1;
    double x[3];
    bool ok = NextRecord(ifs, f, line) && f.size() == 6 && ToInt(f[0], id) && id == i && ToDouble(f[1], x[0]) && ToDouble(f[2], x[1]) &&
              ToDouble(f[3], x[2]) && ToInt(f[4], parent) && parent >= 0 && parent < (int)mkfs.size() && mkfs[parent]->mmpKeyFrames.count(f[5]);
    if (!ok) { FreeMap(map); return Fail(error, "bad point record", line); }
    MapPoint* p = new MapPoint;
    p->mnID = i;
    p->mv3WorldPos = makeVector(x[0], x[1], x[2]);
    p->mpPatchSourceKF = mkfs[parent]->mmpKeyFrames[f[5]];
    points.push_back(p);
    map.mlpPoints.push_back(p);
  }
  // measurements
  if (!NextRecord(ifs, f, line) || f.size() != 1 || !ToInt(f[0], n) || n < 0) { FreeMap(map); return Fail(error, "expected the number of measurements", line); }
  for (int i = 0; i < n; i++) {
    int mkf = -1, pt = -1;
    double u = 0, v = 0, noise = 0;
    bool ok = NextRecord(ifs, f, line) && f.size() == 6 && ToInt(f[0], mkf) && mkf >= 0 && mkf < (int)mkfs.size() && mkfs[mkf]->mmpKeyFrames.count(f[1]) &&
              ToInt(f[2], pt) && pt >= 0 && pt < (int)points.size() && ToDouble(f[3], u) && ToDouble(f[4], v) && ToDouble(f[5], noise) && noise >= 1;
    int level = 0;
    if (ok) {
      level = (int)std::lround(0.5 * std::log2(noise));
      ok = level >= 0 && level < LEVELS && LevelScale(level) * LevelScale(level) == noise;
    }
    KeyFrame* kf = ok ? mkfs[mkf]->mmpKeyFrames[f[1]] : nullptr;
    if (ok && kf->mmpMeasurements.count(points[pt])) ok = false;       // a keyframe measures a point once
    if (!ok) { FreeMap(map); return Fail(error, "bad measurement record", line); }
    Measurement* m = new Measurement;
    m->nLevel = level;
    m->v2RootPos = makeVector(u, v);
    kf->mmpMeasurements[points[pt]] = m;
    points[pt]->mMMData.spMeasurementKFs.insert(kf);
  }
  return true;
}

void FreeMap(Map& map)
{
  for (MultiKeyFrame* mkf : map.mlpMultiKeyFrames) {
    for (auto& kv : mkf->mmpKeyFrames) {
      for (auto& mm : kv.second->mmpMeasurements) delete mm.second;
      delete kv.second;
    }
    delete mkf;
  }
  for (MapPoint* p : map.mlpPoints) delete p;
  map.mlpMultiKeyFrames.clear();
  map.mlpPoints.clear();
}

}  // namespace mcp_host
