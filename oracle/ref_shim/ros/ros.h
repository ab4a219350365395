#pragma once
#include <ros/assert.h>
namespace ros {
struct WallDuration {};
struct WallTime {
  static WallTime now() { return WallTime(); }
  WallDuration operator-(const WallTime&) const { return WallDuration(); }
};
}  // namespace ros
