// shim/MapTypes.h — the members of the reference's map types that the hot-path callers touch
// (include/mcptam/{KeyFrame,MapPoint,Map}.h).  In a real MCPTAM tree include those headers instead.
#pragma once

#include <deque>
#include <list>
#include <map>
#include <tuple>
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
// boost::circular_buffer<T>(capacity) as Level uses it: push_back drops the oldest element once full
template <class T> struct RingBuffer {
  explicit RingBuffer(size_t cap = 2) : mnCap(cap) {}
  void push_back(const T& v) { if (mq.size() == mnCap) mq.pop_front(); mq.push_back(v); }
  T& back() { return mq.back(); }
  T& operator[](size_t i) { return mq[i]; }
  const T& operator[](size_t i) const { return mq[i]; }
  size_t size() const { return mq.size(); }
  size_t capacity() const { return mnCap; }
  void clear() { mq.clear(); }
  std::deque<T> mq;
  size_t mnCap;
};
struct Level {                    // include/mcptam/KeyFrame.h:120-150
  static const int snNumPrev = 2; // include/mcptam/KeyFrame.h:152, src/KeyFrame.cc:59
  Level() : imagePrev(snNumPrev), vCornersPrev(snNumPrev) {}
  Image<byte> image;
  Image<byte> mask;               // internal mask (KeyFrame::SetMask); empty: none
  Image<byte> lastMask;           // the mask the last MakeKeyFrame_Lite filtered the corners with
  std::vector<ImageRef> vCorners;
  std::vector<int> vCornerRowLUT;
  int vFastFrequency[MAX_FAST_THRESH + 1] = { 0 };
  int nFastThresh = 0;
  RingBuffer<Image<byte> > imagePrev;                   // previous images / corners, newest last
  RingBuffer<std::vector<ImageRef> > vCornersPrev;
};
typedef std::map<MapPoint*, Measurement*> MeasPtrMap;
}  // namespace mcp_shim
namespace mcp_host { class FrontEndDevice; }
namespace mcp_shim {
struct KeyFrame {
  // include/mcptam/KeyFrame.h:186 -- same signature; the work runs on the camera's FrontEndDevice (AttachDevice), which keeps
  // the pyramids of this keyframe and of its snNumPrev predecessors resident (slots nFirstSlot .. nFirstSlot + snNumPrev)
  std::tuple<double, double, double> MakeKeyFrame_Lite(Image<byte>& im, bool bDeepCopy = false, bool bGlareMasking = false);
  void SetMask(Image<byte>& m);                          // src/KeyFrame.cc:116-126
  void AttachDevice(mcp_host::FrontEndDevice* dev, int nFirstSlot) { mpDevice = dev; mnFirstSlot = nFirstSlot; mnSlotTurn = 0; nDeviceSlot = -1; }
  int PrevDeviceSlot(int nBack) const;                   // device slot of the image nBack frames ago (1 = imagePrev.back()), -1 if none
  mcp_host::FrontEndDevice* mpDevice = nullptr;
  int mnFirstSlot = 0, mnSlotTurn = 0;
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
