#include "FrontEnd.h"

#include <cmath>
#include <cstdio>
#include <stdexcept>

namespace mcp_host {

FrontEndDevice::FrontEndDevice(int width, int height, int max_keyframes, int max_patches, int device) : mnW(width), mnH(height)
{
  McpFeConfig cfg;
  mcp_fe_default_config(&cfg);
  cfg.width = width; cfg.height = height; cfg.max_keyframes = max_keyframes; cfg.max_patches = max_patches; cfg.device = device;
  if (mcp_fe_create(&cfg, &mpFe) != MCP_OK) throw std::runtime_error(std::string("FrontEndDevice: ") + mcp_last_error());
}
FrontEndDevice::~FrontEndDevice() { mcp_fe_destroy(mpFe); }

std::tuple<double, double, double> MakeKeyFrame_Lite(FrontEndDevice& dev, KeyFrame& kf, int slot, BasicImage<byte>& im, bool bCopyLevelImages)
{
  McpLevelOut out[MCP_LEVELS];
  std::vector<std::vector<int32_t> > cor(MCP_LEVELS);
  int w = dev.width(), h = dev.height();
  const int cap = 8192;
  for (int l = 0; l < MCP_LEVELS; l++) {
    Level& lev = kf.maLevels[l];
    std::memset(&out[l], 0, sizeof(out[l]));
    cor[l].resize(2 * (size_t)cap);
    lev.vCornerRowLUT.assign(h, 0);
    if (bCopyLevelImages) { lev.image.resize(ImageRef(w, h)); out[l].image = lev.image.data(); }
    out[l].corners_xy = cor[l].data(); out[l].corners_cap = cap; out[l].row_lut = lev.vCornerRowLUT.data();
    w /= 2; h /= 2;
  }
  if (mcp_fe_make_keyframe(dev.handle(), slot, im.data(), im.row_stride(), out) != MCP_OK)
    throw std::runtime_error(std::string("MakeKeyFrame_Lite: ") + mcp_last_error());
  for (int l = 0; l < MCP_LEVELS; l++) {
    Level& lev = kf.maLevels[l];
    lev.vCorners.resize(out[l].n_corners);
    for (int i = 0; i < out[l].n_corners; i++) lev.vCorners[i] = ImageRef(cor[l][2 * i], cor[l][2 * i + 1]);
    lev.nFastThresh = out[l].fast_thresh;
    for (int t = 0; t <= MAX_FAST_THRESH; t++) lev.vFastFrequency[t] = out[l].fast_freq[t];
  }
  kf.nDeviceSlot = slot;
  McpFeTiming tm;
  mcp_fe_get_timing(dev.handle(), &tm);
  return std::make_tuple(tm.ms_pyramid * 1e-3, 0.0, tm.ms_fast * 1e-3);
}

}  // namespace mcp_host

// KeyFrame::MakeKeyFrame_Lite (src/KeyFrame.cc:145-361) with the reference's member signature.  The previous image and corner
// list of every level go into the imagePrev / vCornersPrev rings before they are overwritten (:152-157,:184-198); the device
// keeps the matching pyramids: the keyframe owns snNumPrev + 1 slots and the new frame takes the oldest one.
namespace mcp_shim {
std::tuple<double, double, double> KeyFrame::MakeKeyFrame_Lite(Image<byte>& im, bool bDeepCopy, bool bGlareMasking)
{
  using namespace mcp_host;
  if (!mpDevice) throw std::runtime_error("KeyFrame::MakeKeyFrame_Lite: no FrontEndDevice attached (KeyFrame::AttachDevice)");
  (void)bDeepCopy;                                       // the level images are always copies here (no reference-counted CVD::Image)
  const bool bPushBack = maLevels[0].image.totalsize() > 0;
  for (int l = 0; l < LEVELS; l++) {
    Level& lev = maLevels[l];
    if (bPushBack) {
      lev.imagePrev.push_back(lev.image);
      lev.vCornersPrev.push_back(std::vector<ImageRef>());
      lev.vCornersPrev.back().swap(lev.vCorners);
    }
    lev.vCorners.clear();
    for (int t = 0; t <= MAX_FAST_THRESH; t++) lev.vFastFrequency[t] = 0;
    lev.nFastThresh = 0;
  }
  const int nSlots = Level::snNumPrev + 1;
  const int slot = mnFirstSlot + (mnSlotTurn % nSlots);
  mnSlotTurn++;
  if (mcp_fe_set_glare_masking(mpDevice->handle(), bGlareMasking ? 1 : 0) != MCP_OK) throw std::runtime_error(mcp_last_error());
  McpLevelOut out[MCP_LEVELS];
  std::vector<std::vector<int32_t> > cor(MCP_LEVELS);
  int w = mpDevice->width(), h = mpDevice->height();
  const int cap = 16384;
  for (int l = 0; l < MCP_LEVELS; l++) {
    Level& lev = maLevels[l];
    std::memset(&out[l], 0, sizeof(out[l]));
    cor[l].resize(2 * (size_t)cap);
    lev.vCornerRowLUT.assign(h, 0);
    lev.image.resize(ImageRef(w, h));
    lev.lastMask.resize(ImageRef(w, h));
    out[l].image = lev.image.data(); out[l].last_mask = lev.lastMask.data();
    out[l].corners_xy = cor[l].data(); out[l].corners_cap = cap; out[l].row_lut = lev.vCornerRowLUT.data();
    w /= 2; h /= 2;
  }
  if (mcp_fe_make_keyframe(mpDevice->handle(), slot, im.data(), im.row_stride(), out) != MCP_OK)
    throw std::runtime_error(std::string("KeyFrame::MakeKeyFrame_Lite: ") + mcp_last_error());
  for (int l = 0; l < MCP_LEVELS; l++) {
    Level& lev = maLevels[l];
    lev.vCorners.resize(out[l].n_corners);
    for (int i = 0; i < out[l].n_corners; i++) lev.vCorners[i] = ImageRef(cor[l][2 * i], cor[l][2 * i + 1]);
    lev.nFastThresh = out[l].fast_thresh;
    for (int t = 0; t <= MAX_FAST_THRESH; t++) lev.vFastFrequency[t] = out[l].fast_freq[t];
  }
  nDeviceSlot = slot;
  McpFeTiming tm;
  mcp_fe_get_timing(mpDevice->handle(), &tm);
  return std::make_tuple(tm.ms_pyramid * 1e-3, 0.0, tm.ms_fast * 1e-3);      // (downsample, mask, feature) seconds; the mask is part of the feature pass
}
void KeyFrame::SetMask(Image<byte>& m)
{
  if (!mpDevice) throw std::runtime_error("KeyFrame::SetMask: no FrontEndDevice attached");
  maLevels[0].mask = m;
  if (mcp_fe_set_mask(mpDevice->handle(), m.totalsize() ? m.data() : nullptr, m.row_stride()) != MCP_OK) throw std::runtime_error(mcp_last_error());
}
int KeyFrame::PrevDeviceSlot(int nBack) const
{
  const int nSlots = Level::snNumPrev + 1;
  if (nBack < 1 || nBack > Level::snNumPrev || nBack > (int)maLevels[0].imagePrev.size() || mnSlotTurn - 1 - nBack < 0) return -1;
  return mnFirstSlot + ((mnSlotTurn - 1 - nBack) % nSlots);
}
}  // namespace mcp_shim

namespace mcp_host {

int CalcSearchLevelAndWarpMatrix(TrackerData& td, const SE3& se3CFromW)
{
  MapPoint& point = *td.mpPoint;
  const Vector<3> v3Cam = se3CFromW * point.mv3WorldPos;
  const Vector<3> v3MotionRight = se3CFromW.get_rotation() * point.mv3PixelRight_W;
  const Vector<3> v3MotionDown = se3CFromW.get_rotation() * point.mv3PixelDown_W;
  Vector<3> dTh, dPh;
  TaylorCamera::GetCamSphereDeriv(v3Cam, dTh, dPh);
  const Vector<2> r = makeVector(dTh * v3MotionRight, dPh * v3MotionRight), d = makeVector(dTh * v3MotionDown, dPh * v3MotionDown);
  const Vector<2> c0 = td.mm2CamDerivs * r, c1 = td.mm2CamDerivs * d;
  td.mm2WarpInverse[0][0] = c0[0]; td.mm2WarpInverse[1][0] = c0[1]; td.mm2WarpInverse[0][1] = c1[0]; td.mm2WarpInverse[1][1] = c1[1];
  double dDet = td.mm2WarpInverse[0][0] * td.mm2WarpInverse[1][1] - td.mm2WarpInverse[0][1] * td.mm2WarpInverse[1][0];
  td.mnSearchLevel = 0;
  while (dDet > 3 && td.mnSearchLevel < LEVELS - 1) { td.mnSearchLevel++; dDet *= 0.25; }
  if (dDet > 3 || dDet < 0.5 || !std::isfinite(dDet)) { td.mbTemplateBad = true; return -1; }
  td.mbTemplateBad = false;
  return td.mnSearchLevel;
}

int SearchForPoints(FrontEndDevice& dev, std::vector<TrackerData*>& vTD, KeyFrame& kfTarget, int nRange, int nSubPixIts, bool bExhaustive,
                    int anAttempted[LEVELS], int anFound[LEVELS])
{
  std::vector<McpPatchReq> req;
  std::vector<int> idx;
  for (size_t i = 0; i < vTD.size(); i++) {
    TrackerData& td = *vTD[i];
    if (td.mbTemplateBad || !td.mpPoint->mpPatchSourceKF || td.mpPoint->mpPatchSourceKF->nDeviceSlot < 0) { td.mbFound = false; continue; }
    McpPatchReq r;
    std::memset(&r, 0, sizeof(r));
    r.src_kf = td.mpPoint->mpPatchSourceKF->nDeviceSlot; r.src_level = td.mpPoint->mnSourceLevel;
    r.src_cx = td.mpPoint->mirCenter.x; r.src_cy = td.mpPoint->mirCenter.y;
    r.warp_inv[0] = td.mm2WarpInverse[0][0]; r.warp_inv[1] = td.mm2WarpInverse[0][1];
    r.warp_inv[2] = td.mm2WarpInverse[1][0]; r.warp_inv[3] = td.mm2WarpInverse[1][1];
    r.search_level = td.mnSearchLevel;
    r.pred_x = (int)td.mv2Image[0]; r.pred_y = (int)td.mv2Image[1];          // CVD::ir()
    const bool ex = td.mpPoint->mbFixed || bExhaustive;                       // src/Tracker.cc:1326-1331
    r.range = nRange; r.subpix_its = ex ? 10 : nSubPixIts; r.exhaustive = ex ? 1 : 0;
    req.push_back(r); idx.push_back((int)i);
  }
  std::vector<McpPatchRes> res(req.size());
  if (!req.empty() && mcp_fe_search_patches(dev.handle(), kfTarget.nDeviceSlot, (int)req.size(), req.data(), res.data()) != MCP_OK)
    throw std::runtime_error(std::string("SearchForPoints: ") + mcp_last_error());
  int nFound = 0;
  for (size_t k = 0; k < req.size(); k++) {
    TrackerData& td = *vTD[idx[k]];
    const McpPatchRes& r = res[k];
    if (r.template_bad) { td.mbFound = false; continue; }
    anAttempted[req[k].search_level]++;
    td.mbSearched = true;
    td.mbFound = r.found != 0;
    td.mbDidSubPix = r.did_subpix != 0;
    if (!td.mbFound) continue;
    td.mdSqrtInvNoise = 1.0 / (1 << req[k].search_level);
    td.mv2Found = makeVector(r.found_x, r.found_y);
    nFound++;
    anFound[req[k].search_level]++;
  }
  return nFound;
}

}  // namespace mcp_host
