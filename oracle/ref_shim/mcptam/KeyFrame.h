// Shadows include/mcptam/KeyFrame.h for oracle/_ref: only the DATA members src/PatchFinder.cc reads (the real header pulls
// in boost, GVars3 and the ROS message types).  Member names/types as in include/mcptam/KeyFrame.h:85,120-136,252.
#pragma once
#include <cvd/byte.h>
#include <cvd/image.h>
#include <vector>
#define LEVELS 4
struct Level {
  CVD::Image<CVD::byte> image;
  std::vector<CVD::ImageRef> vCorners;
  std::vector<int> vCornerRowLUT;
};
class KeyFrame {
public:
  Level maLevels[LEVELS];
};
