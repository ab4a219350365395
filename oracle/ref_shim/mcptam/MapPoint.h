// Shadows include/mcptam/MapPoint.h for oracle/_ref: the data members src/PatchFinder.cc reads
// (include/mcptam/MapPoint.h:125-148).
#pragma once
#include <TooN/TooN.h>
#include <cvd/image_ref.h>
class KeyFrame;
class MapPoint {
public:
  TooN::Vector<3> mv3WorldPos;
  KeyFrame* mpPatchSourceKF;
  int mnSourceLevel;
  CVD::ImageRef mirCenter;
  TooN::Vector<3> mv3PixelDown_W;
  TooN::Vector<3> mv3PixelRight_W;
};
