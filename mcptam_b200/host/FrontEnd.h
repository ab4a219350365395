// FrontEnd.h — host mirror of the per-frame front end on the B200 path:
//   KeyFrameCuda::MakeKeyFrame_Lite   <- KeyFrame::MakeKeyFrame_Lite            (src/KeyFrame.cc:145-361)
//   PatchFinderBatch                  <- PatchFinder 5-step API as used by Tracker::SearchForPoints
//                                        (src/Tracker.cc:1299-1377, src/PatchFinder.cc:69-122 on the host)
// One FrontEndDevice (= one mcp_fe handle, one CUDA stream) per camera.
#pragma once

#include <string>
#include <tuple>
#include <vector>

#include "TaylorCamera.h"
#include "shim/MapTypes.h"

namespace mcp_host {

class FrontEndDevice {
 public:
  FrontEndDevice(int width, int height, int max_keyframes = 8, int max_patches = 4096, int device = -1);
  ~FrontEndDevice();
  McpFe* handle() { return mpFe; }
  int width() const { return mnW; }
  int height() const { return mnH; }

 private:
  McpFe* mpFe = nullptr;
  int mnW, mnH;
};

// Fills kf.maLevels[*] (image, vCorners, vCornerRowLUT, nFastThresh, vFastFrequency) from one camera image and
// leaves the pyramid resident in device slot `slot` (kf.nDeviceSlot).  Returns (downsample, mask, feature) seconds.
std::tuple<double, double, double> MakeKeyFrame_Lite(FrontEndDevice& dev, KeyFrame& kf, int slot, BasicImage<byte>& im,
                                                     bool bCopyLevelImages = true);

// Per-(point, camera) scratch of the tracker (include/mcptam/TrackerData.h): the fields SearchForPoints touches.
struct TrackerData {
  MapPoint* mpPoint = nullptr;
  Vector<2> mv2Image;                 // predicted position (level 0)
  Matrix<2> mm2CamDerivs;             // TaylorCamera::GetProjectionDerivs at the prediction
  Matrix<2> mm2WarpInverse;
  int mnSearchLevel = 0;
  bool mbTemplateBad = true, mbSearched = false, mbFound = false, mbDidSubPix = false;
  Vector<2> mv2Found;
  double mdSqrtInvNoise = 0;
};

// PatchFinder::CalcSearchLevelAndWarpMatrix (src/PatchFinder.cc:69-122); returns the level or -1
int CalcSearchLevelAndWarpMatrix(TrackerData& td, const SE3& se3CFromW);

// Tracker::SearchForPoints for one camera: one device call for the whole vTD.  Returns the number found and
// increments the per-level attempted / found counters exactly like src/Tracker.cc:1322,1347,1361.
int SearchForPoints(FrontEndDevice& dev, std::vector<TrackerData*>& vTD, KeyFrame& kfTarget, int nRange, int nSubPixIts,
                    bool bExhaustive, int anAttempted[LEVELS], int anFound[LEVELS]);

}  // namespace mcp_host
