// MapIO.h — the reference's on-disk map dump (MapMakerBase::DumpToFile, src/MapMakerBase.cc:475-579) written from
// and read back into the map types the bundle adjuster marshals (the data format on the map side
// of the BA path).  Pure host code.
//
// Format (text, one record per line, '%' lines are comments the writer always emits in the same places):
//   number of cameras;   "name, px, py, pz, qx, qy, qz, qw"      pose of the camera IN the MKF frame (= CamFromBase^-1)
//   number of MKFs;      "id, px, py, pz, qx, qy, qz, qw"        pose of the MKF IN the world     (= BaseFromWorld^-1)
//   number of points;    "id, x, y, z, parent MKF id, parent camera name"
//   number of measurements; "MKF id, camera name, point id, u, v, noise"   level-0 pixels, noise = LevelScale(level)^2
//   "% The end"
// Numbers are written with the stream's default precision (6 significant digits), exactly like the reference, so a
// dump is a lossy snapshot; LoadFromFile() accepts any precision.
#pragma once

#include <string>

#include "shim/MapTypes.h"

namespace mcp_host {

using namespace mcp_shim;

// tf::Matrix3x3::getRotation / tf::Quaternion -> matrix [3P: Bullet], the conversions util::SE3ToPoseMsg /
// util::PoseMsgToSE3 go through (include/mcptam/Utility.h:128-186).  q = (x, y, z, w).
void RotationToQuaternion(const Matrix<3>& R, double q[4]);
void QuaternionToRotation(const double q[4], Matrix<3>& R);

// MapMakerBase::DumpToFile: assigns mnID to MKFs and points in list order, as the reference does.
bool DumpToFile(Map& map, const std::string& filename);

// Rebuilds a map from a dump: one KeyFrame per (MKF, camera), mse3CamFromWorld = CamFromBase * BaseFromWorld,
// measurements with nLevel = log2(sqrt(noise)), source keyframe of every point from (parent MKF, parent camera).
// The first MKF is marked fixed (src/MapMakerServerBase.cc:150: the first MKF of a map never moves).  The caller owns
// the objects (FreeMap).  Returns false and sets *error on a malformed file.
bool LoadFromFile(const std::string& filename, Map& map, std::string* error = nullptr);
void FreeMap(Map& map);

}  // namespace mcp_host
