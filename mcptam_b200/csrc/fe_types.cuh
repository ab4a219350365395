// fe_types.cuh — device-side layout of the front end: resident keyframe pyramids, corner lists, row LUTs.
#pragma once

#include "mcp_common.cuh"

namespace mcp {

struct FeLevel {
  uint8_t* img;          // [h][pitch] level image
  uint8_t* score;        // [h][pitch] FAST score map: 0 = no corner at b=5, else fast_corner_score_10 (5..254)
  const uint8_t* mask;   // [h][pitch] fixed mask pyramid (255 = usable) or nullptr
  int2* corners;         // [corner_cap] raster-ordered corners after threshold + mask
  int* row_lut;          // [h] Level::vCornerRowLUT
  int* rowcount;         // [h]   (four-kernel path; aliases rowhist)
  int* rowhist;          // [h][32] per-row histogram of capped scores of the corners that pass the mask, + ticket at [h*32]; zero between frames
  unsigned* hist;        // [32] capped-score histogram
  int w, h, pitch, pad_;
};

struct FeMetaLevel {
  int width, height, n_corners, fast_thresh;
  int fast_freq[31];
  int pad_;
};
struct FeMeta { FeMetaLevel lv[MCP_LEVELS]; };

struct FeKf {
  FeLevel lv[MCP_LEVELS];
  FeMeta* meta;
  int tile_off[MCP_LEVELS + 1];   // prefix sums of FAST tiles per level
  int row_off[MCP_LEVELS + 1];    // prefix sums of image rows per level
  int fixed_thresh[MCP_LEVELS];
  int corner_cap, pad_;
  long long host_delta;           // != 0: byte offset from this slot's device output block to its pinned host mirror (the two-launch FAST writes both)
};

struct FeMasks { unsigned char* m[MCP_LEVELS]; };   // per-level effective masks (pitch = the level's image pitch)

struct FeDev {
  const FeKf* kf;        // [n_slots] in device memory
  int n_slots;
  int transform_round;
};

}  // namespace mcp
