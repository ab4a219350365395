// shim/MapTypes.h — the members of the reference's map types that the hot-path callers touch
// (include/mcptam/{KeyFrame,MapPoint,Map}.h).  In a real MCPTAM tree include those headers instead.
#pragma once

#include <list>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "Types.h"

namespace mcp_shim {

#define LEVELS 4
#define MAX_FAST_THRESH 30

struct MapPoint;
struct KeyFrame;
struct MultiKeyFrame;

struct Measurement {              // include/mcptam/KeyFrame.h:100-117
  int nLevel = 0;
  bool bSubPix = false;
  Vector<2> v2RootPos;
};
struct Level {                    // include/mcptam/KeyFrame.h:120-150
  Image<byte> image;
  std::vector<ImageRef> vCorners;
  std::vector<int> vCornerRowLUT;
  int vFastFrequency[MAX_FAST_THRESH + 1] = { 0 };
  int nFastThresh = 0;
};
typedef std::map<MapPoint*, Measurement*> MeasPtrMap;
struct KeyFrame {
  Level maLevels[LEVELS];
  MeasPtrMap mmpMeasurements;
  SE3 mse3CamFromBase, mse3CamFromWorld;
  std::string mCamName;
  MultiKeyFrame* mpParent = nullptr;
  int nDeviceSlot = -1;           // resident pyramid slot on the B200 (new)
};
typedef std::map<std::string, KeyFrame*> KeyFramePtrMap;
struct MultiKeyFrame {
  SE3 mse3BaseFromWorld;
  bool mbFixed = false, mbBad = false;
  int mnID = -1;
  KeyFramePtrMap mmpKeyFrames;
  void RefreshSceneDepthRobust() {}
};
struct MapPoint {
  Vector<3> mv3WorldPos;
  bool mbFixed = false, mbBad = false, mbOptimized = false;
  int mnID = -1;
  KeyFrame* mpPatchSourceKF = nullptr;
  int mnSourceLevel = 0;
  ImageRef mirCenter;
  Vector<3> mv3PixelRight_W, mv3PixelDown_W;
  struct { std::set<KeyFrame*> spMeasurementKFs; int GoodMeasCount() const { return (int)spMeasurementKFs.size(); } } mMMData;
  void RefreshPixelVectors() {}
};

typedef std::list<MultiKeyFrame*> MultiKeyFramePtrList;
typedef std::list<MapPoint*> MapPointPtrList;
struct Map {                      // include/mcptam/Map.h: the two lists the dump / load code walks
  MultiKeyFramePtrList mlpMultiKeyFrames;
  MapPointPtrList mlpPoints;
};

}  // namespace mcp_shim
