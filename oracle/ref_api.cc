// oracle/_ref: C entry points around the REFERENCE'S OWN translation units, compiled where they lie under /root/reference
// (build_ref.py).  TEST INFRASTRUCTURE ONLY: tests/test_oracle_vs_ref.py checks the restatement (oracle/*.c) against it.
// Nothing in this file restates the reference; it only moves arrays in and out of the reference's classes.
#include <cmath>
#include <ros/ros.h>
#include <mcptam/MEstimator.h>   // (the reference includes it after <ros/ros.h>, src/ChainBundle.cc:42-44)
#include <mcptam/MiniPatch.h>
#include <mcptam/ShiTomasi.h>
#include <mcptam/LevelHelpers.h>
#include <mcptam/SmallMatrixOpts.h>
#include <mcptam/TaylorCamera.h>
#include <mcptam/PatchFinder.h>
#include <mcptam/KeyFrame.h>
#include <mcptam/MapPoint.h>

#include <cstdint>
#include <cstring>
#include <vector>

TooN::Vector<3> gavLevelColors[LEVELS];          // defined in src/KeyFrame.cc in the reference (LevelHelpers.h declares it)

namespace {
CVD::BasicImage<CVD::byte> wrap(const uint8_t* im, int w, int h, int stride) { return CVD::BasicImage<CVD::byte>(const_cast<uint8_t*>(im), CVD::ImageRef(w, h), stride); }
void fill_level(Level& L, const uint8_t* im, int w, int h, int stride, const int32_t* corners_xy, int n_corners, const int32_t* row_lut)
{
  L.image.resize(CVD::ImageRef(w, h));
  for (int y = 0; y < h; y++) memcpy(L.image[y], im + (size_t)y * stride, (size_t)w);
  L.vCorners.clear();
  for (int i = 0; i < n_corners; i++) L.vCorners.push_back(CVD::ImageRef(corners_xy[2 * i], corners_xy[2 * i + 1]));
  L.vCornerRowLUT.assign(row_lut, row_lut + (row_lut ? h : 0));
}
// PatchFinder keeps its template protected: a subclass hands it out
struct PF : PatchFinder {
  const CVD::Image<CVD::byte>& tmpl() const { return mimTemplate; }
  void set_level(int l) { mnSearchLevel = l; }
  void set_warp(const double* w) { mm2WarpInverse(0, 0) = w[0]; mm2WarpInverse(0, 1) = w[1]; mm2WarpInverse(1, 0) = w[2]; mm2WarpInverse(1, 1) = w[3]; }
  void get_warp(double* w) const { w[0] = mm2WarpInverse(0, 0); w[1] = mm2WarpInverse(0, 1); w[2] = mm2WarpInverse(1, 0); w[3] = mm2WarpInverse(1, 1); }
  int tsum() const { return mnTemplateSum; }
  int tsumsq() const { return mnTemplateSumSq; }
};
}  // namespace

extern "C" {

// ---- include/mcptam/MEstimator.h (compiled verbatim) --------------------------------------------------------------
double ref_huber_sigma_sq(const double* e2, int n) { std::vector<double> v(e2, e2 + n); return Huber::FindSigmaSquared(v); }
double ref_tukey_sigma_sq(const double* e2, int n) { std::vector<double> v(e2, e2 + n); return Tukey::FindSigmaSquared(v); }
double ref_cauchy_sigma_sq(const double* e2, int n) { std::vector<double> v(e2, e2 + n); return Cauchy::FindSigmaSquared(v); }
void ref_mestimator_weights(int which, const double* e2, int n, double sigma_sq, double* sqrt_w, double* w, double* obj)
{
  for (int i = 0; i < n; i++) {
    if (which == 0) { sqrt_w[i] = Tukey::SquareRootWeight(e2[i], sigma_sq); w[i] = Tukey::Weight(e2[i], sigma_sq); obj[i] = Tukey::ObjectiveScore(e2[i], sigma_sq); }
    else if (which == 1) { sqrt_w[i] = Cauchy::SquareRootWeight(e2[i], sigma_sq); w[i] = Cauchy::Weight(e2[i], sigma_sq); obj[i] = Cauchy::ObjectiveScore(e2[i], sigma_sq); }
    else { sqrt_w[i] = Huber::SquareRootWeight(e2[i], sigma_sq); w[i] = Huber::Weight(e2[i], sigma_sq); obj[i] = Huber::ObjectiveScore(e2[i], sigma_sq); }
  }
}

// ---- src/ShiTomasi.cc -----------------------------------------------------------------------------------------------
double ref_shitomasi(const uint8_t* im, int w, int h, int stride, int half_box, int x, int y)
{
  CVD::BasicImage<CVD::byte> I = wrap(im, w, h, stride);
  return FindShiTomasiScoreAtPoint(I, half_box, CVD::ImageRef(x, y));
}

// ---- src/MiniPatch.cc -----------------------------------------------------------------------------------------------
// returns found; pos_xy in: start position, out: best corner
int ref_minipatch_find(const uint8_t* src, const uint8_t* dst, int w, int h, int stride, int src_x, int src_y, int32_t* pos_xy, int range,
                       const int32_t* corners_xy, int n_corners, const int32_t* row_lut)
{
  CVD::BasicImage<CVD::byte> S = wrap(src, w, h, stride), D = wrap(dst, w, h, stride);
  MiniPatch mp;
  mp.SampleFromImage(CVD::ImageRef(src_x, src_y), S);
  std::vector<CVD::ImageRef> vc;
  for (int i = 0; i < n_corners; i++) vc.push_back(CVD::ImageRef(corners_xy[2 * i], corners_xy[2 * i + 1]));
  std::vector<int> lut;
  if (row_lut) lut.assign(row_lut, row_lut + h);
  CVD::ImageRef pos(pos_xy[0], pos_xy[1]);
  const bool ok = mp.FindPatch(pos, D, range, vc, row_lut ? &lut : NULL);
  pos_xy[0] = pos.x; pos_xy[1] = pos.y;
  return ok ? 1 : 0;
}

// ---- include/mcptam/LevelHelpers.h ------------------------------------------------------------------------------------
double ref_level_zero_pos(double p, int level) { return LevelZeroPos(p, level); }
double ref_level_n_pos(double p, int level) { return LevelNPos(p, level); }

// ---- src/TaylorCamera.cc ------------------------------------------------------------------------------------------------
void* ref_cam_create(const double* params9, int calib_w, int calib_h, int full_w, int full_h, int img_w, int img_h)
{
  TooN::Vector<9> p;
  for (int i = 0; i < 9; i++) p[i] = params9[i];
  return new TaylorCamera(p, CVD::ImageRef(calib_w, calib_h), CVD::ImageRef(full_w, full_h), CVD::ImageRef(img_w, img_h));
}
void ref_cam_destroy(void* c) { delete static_cast<TaylorCamera*>(c); }
// the derived quantities RefreshParams / FindInvPolyUsingRoots leave in the object (through a layout-compatible accessor)
struct CamPeek : TaylorCamera {
  CamPeek() : TaylorCamera(CVD::ImageRef(1, 1), CVD::ImageRef(1, 1), CVD::ImageRef(1, 1)) {}
  static void get(TaylorCamera* c, double* center2, double* affine4, double* min_theta, double* mean, double* std_, int* n_inv, double* inv32)
  {
    CamPeek* p = static_cast<CamPeek*>(c);
    center2[0] = p->mv2Center[0]; center2[1] = p->mv2Center[1];
    affine4[0] = p->mm2Affine(0, 0); affine4[1] = p->mm2Affine(0, 1); affine4[2] = p->mm2Affine(1, 0); affine4[3] = p->mm2Affine(1, 1);
    *min_theta = p->mdMinTheta; *mean = p->mdThetaMean; *std_ = p->mdThetaStd;
    *n_inv = p->mvxPolyInvCoeffs.size();
    for (int i = 0; i < p->mvxPolyInvCoeffs.size() && i < 32; i++) inv32[i] = p->mvxPolyInvCoeffs[i];
  }
};
void ref_cam_derived(void* c, double* center2, double* affine4, double* min_theta, double* mean, double* std_, int* n_inv, double* inv32)
{ CamPeek::get(static_cast<TaylorCamera*>(c), center2, affine4, min_theta, mean, std_, n_inv, inv32); }
// Project + GetProjectionDerivs + GetCamSphereDeriv for n camera-frame points
void ref_cam_project(void* c, int n, const double* xyz, double* px2, int32_t* invalid, double* derivs4, double* dtheta3, double* dphi3)
{
  TaylorCamera* cam = static_cast<TaylorCamera*>(c);
  for (int i = 0; i < n; i++) {
    const TooN::Vector<3> v = TooN::makeVector(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    const TooN::Vector<2> u = cam->Project(v);
    px2[2 * i] = u[0]; px2[2 * i + 1] = u[1];
    invalid[i] = cam->Invalid() ? 1 : 0;
    const TooN::Matrix<2> D = cam->GetProjectionDerivs();
    derivs4[4 * i] = D(0, 0); derivs4[4 * i + 1] = D(0, 1); derivs4[4 * i + 2] = D(1, 0); derivs4[4 * i + 3] = D(1, 1);
    TooN::Vector<3> dt, dp;
    TaylorCamera::GetCamSphereDeriv(v, dt, dp);
    for (int k = 0; k < 3; k++) { dtheta3[3 * i + k] = dt[k]; dphi3[3 * i + k] = dp[k]; }
  }
}
void ref_cam_unproject(void* c, int n, const double* px2, double* ray3)
{
  TaylorCamera* cam = static_cast<TaylorCamera*>(c);
  for (int i = 0; i < n; i++) {
    const TooN::Vector<3> r = cam->UnProject(TooN::makeVector(px2[2 * i], px2[2 * i + 1]));
    for (int k = 0; k < 3; k++) ray3[3 * i + k] = r[k];
  }
}

// ---- src/PatchFinder.cc ---------------------------------------------------------------------------------------------------
// ZMSSDAtPoint of an 8x8 template (both the SSE and the scalar branch exist in the reference; which one is compiled is
// chosen by CVD_HAVE_XMMINTRIN inside the reference source: build_ref.py builds the file twice)
int REF_SYM(ref_zmssd)(const uint8_t* im, int w, int h, int stride, const uint8_t* templ64, int x, int y)
{
  PF pf;
  KeyFrame kf;
  fill_level(kf.maLevels[0], templ64, 8, 8, 8, NULL, 0, NULL);
  pf.MakeTemplateCoarseNoWarp(kf, 0, CVD::ImageRef(4, 4));     // border test fails for an 8x8 image: fill the template by hand
  CVD::Image<CVD::byte>& T = const_cast<CVD::Image<CVD::byte>&>(pf.tmpl());
  memcpy(T.data(), templ64, 64);
  struct Sums : PF { static void make(PF& p) { static_cast<Sums&>(p).run(); } void run() { int s = 0, q = 0; for (int i = 0; i < 64; i++) { const int b = mimTemplate.data()[i]; s += b; q += b * b; } mnTemplateSum = s; mnTemplateSumSq = q; } };
  Sums::make(pf);
  CVD::BasicImage<CVD::byte> I = wrap(im, w, h, stride);
  return pf.ZMSSDAtPoint(I, CVD::ImageRef(x, y));
}

// CalcSearchLevelAndWarpMatrix: returns the level (-1: bad), warp_inv4 out
int REF_SYM(ref_calc_search_level)(const double* cfw12, const double* world3, const double* right3, const double* down3, const double* cam_derivs4, double* warp_inv4)
{
  PF pf;
  MapPoint pt;
  TooN::Matrix<3> R;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R(i, j) = cfw12[3 * i + j];
  TooN::SE3<> T(TooN::SO3<>(R), TooN::makeVector(cfw12[9], cfw12[10], cfw12[11]));
  pt.mv3WorldPos = TooN::makeVector(world3[0], world3[1], world3[2]);
  pt.mv3PixelRight_W = TooN::makeVector(right3[0], right3[1], right3[2]);
  pt.mv3PixelDown_W = TooN::makeVector(down3[0], down3[1], down3[2]);
  TooN::Matrix<2> D;
  D(0, 0) = cam_derivs4[0]; D(0, 1) = cam_derivs4[1]; D(1, 0) = cam_derivs4[2]; D(1, 1) = cam_derivs4[3];
  const int lvl = pf.CalcSearchLevelAndWarpMatrix(pt, T, D);
  pf.get_warp(warp_inv4);
  return lvl;
}

// The Tracker::SearchForPoints sequence for one point (src/Tracker.cc:1312-1371) on the reference's PatchFinder:
// MakeTemplateCoarseCont -> FindPatchCoarse -> MakeSubPixTemplate/IterateSubPixToConvergence.
// out: [template_bad, found, did_subpix, score, coarse_x, coarse_y], found_xy, template bytes
void REF_SYM(ref_patch_search)(const uint8_t* src_im, int sw, int sh, int sstride, int src_cx, int src_cy, const double* warp_inv4, int search_level,
                      const uint8_t* tgt_im, int tw, int th, int tstride, const int32_t* corners_xy, int n_corners, const int32_t* row_lut,
                      int pred_x, int pred_y, int range, int subpix_its, int exhaustive, int32_t* out6, double* found_xy, uint8_t* templ64)
{
  PF pf;
  KeyFrame src, tgt;
  fill_level(src.maLevels[0], src_im, sw, sh, sstride, NULL, 0, NULL);
  fill_level(tgt.maLevels[search_level], tgt_im, tw, th, tstride, corners_xy, n_corners, row_lut);
  MapPoint pt;
  pt.mpPatchSourceKF = &src;
  pt.mnSourceLevel = 0;
  pt.mirCenter = CVD::ImageRef(src_cx, src_cy);
  pf.set_warp(warp_inv4);
  pf.set_level(search_level);
  memset(out6, 0, sizeof(int32_t) * 6);
  found_xy[0] = found_xy[1] = 0;
  pf.MakeTemplateCoarseCont(pt);
  memcpy(templ64, pf.tmpl().data(), 64);
  if (pf.TemplateBad()) { out6[0] = 1; return; }
  int score = 0;
  const bool found = pf.FindPatchCoarse(CVD::ImageRef(pred_x, pred_y), tgt, (unsigned)range, score, exhaustive != 0);
  out6[3] = score;
  if (!found) return;
  out6[1] = 1;
  const TooN::Vector<2> coarse = pf.GetCoarsePosAsVector();
  out6[4] = pf.GetCoarsePos().x; out6[5] = pf.GetCoarsePos().y;
  found_xy[0] = coarse[0]; found_xy[1] = coarse[1];
  if (subpix_its > 0) {
    pf.MakeSubPixTemplate();
    pf.SetSubPixPos(coarse);                               // src/Tracker.cc:1351
    const bool ok = pf.IterateSubPixToConvergence(tgt, subpix_its);
    if (!ok) { out6[1] = 0; return; }                      // src/Tracker.cc:1354-1358: rejected
    out6[2] = 1;
    const TooN::Vector<2> sp = pf.GetSubPixPos();
    found_xy[0] = sp[0]; found_xy[1] = sp[1];
  }
}

}  // extern "C"
