#include "Epipolar.h"

#include <algorithm>
#include <cmath>
#include <tuple>

namespace mcp_host {

static inline int LevelScale(int l) { return 1 << l; }
static inline Vector<2> LevelZeroPos(const ImageRef& ir, int nLevel)        // include/mcptam/LevelHelpers.h:61-64
{
  const double s = LevelScale(nLevel);
  return makeVector((ir.x + 0.5) * s - 0.5, (ir.y + 0.5) * s - 0.5);
}
static inline double Norm(const Vector<3>& v) { return std::sqrt(v * v); }
static inline Vector<3> Normalized(const Vector<3>& v) { return v * (1.0 / Norm(v)); }

Vector<3> UnProject(const McpTaylorCam& cam, const Vector<2>& v2Im)
{
  const double det = cam.affine[0] * cam.affine[3] - cam.affine[1] * cam.affine[2];
  const double dx = v2Im[0] - cam.center[0], dy = v2Im[1] - cam.center[1];
  const double u = (cam.affine[3] * dx - cam.affine[1] * dy) / det;          // mm2AffineInv * (v2ImFrame - mv2Center)
  const double v = (-cam.affine[2] * dx + cam.affine[0] * dy) / det;
  const double rho = std::sqrt(u * u + v * v);
  double z = 0;
  for (int i = 4; i > 0; i--) z = (z + cam.poly[i]) * rho;                   // PolyVal(mv5PolyCoeffs, rho)
  z += cam.poly[0];
  return Normalized(makeVector(u, v, z));
}

double OnePixelAngle(const McpTaylorCam& cam)
{
  const Vector<2> c = makeVector(cam.image_size[0] / 2, cam.image_size[1] / 2);
  const Vector<3> a = UnProject(cam, c), b = UnProject(cam, c + makeVector(1, 1));
  return std::acos(a * b) / std::sqrt(2.0);
}

// The depth hypotheses of AddPointEpipolar (:620-724).  The triangle source centre / target centre / scene point fixes
// the depth range on the source ray through the admissible epipolar angle [0.05 rad, 60 deg] (law of sines); the
// hypotheses are spaced EVENLY IN ANGLE as seen from the target camera -- three target pixels at the candidate's level --
// by walking the unit circle of the epipolar plane and intersecting each direction with the source ray.
bool EpipolarHypotheses(const SE3& src, const SE3& tgt, const Vector<3>& ray_s, double one_pixel_angle, int level,
                        std::vector<std::pair<Vector<3>, Vector<3> > >& hypotheses, double* depth_near, double* depth_far)
{
  hypotheses.clear();
  const SE3 world_from_tgt = tgt.inverse(), world_from_src = src.inverse();
  // the source ray in the target frame: origin (source camera centre) and direction
  const Vector<3> origin_t = tgt * world_from_src.get_translation();
  const Vector<3> dir_t = tgt.get_rotation() * (world_from_src.get_rotation() * ray_s);
  // the target centre seen from the source: baseline length and the angle between baseline and ray at the source
  const Vector<3> tgt_centre_s = src * world_from_tgt.get_translation();
  const double baseline = Norm(tgt_centre_s);
  const double angle_at_source = std::acos((tgt_centre_s * ray_s) / baseline);
  const double widest = M_PI / 3, narrowest = 0.05;                          // admissible angles at the scene point
  auto depth_for = [&](double angle_at_point) { return baseline * std::sin(M_PI - angle_at_source - angle_at_point) / std::sin(angle_at_point); };
  double near = depth_for(widest);
  const double far = depth_for(narrowest);
  if (near < 0.2) near = 0.2;                                                // "don't bother looking too close"
  if (depth_near) *depth_near = near;
  if (depth_far) *depth_far = far;
  const Vector<3> p_near = origin_t + dir_t * near, p_far = origin_t + dir_t * far;
  const Vector<3> u_near = Normalized(p_near), u_far = Normalized(p_far);
  const Vector<3> gap = u_near - u_far;
  if (gap * gap < 0.00000001) return false;                                  // the arc is too short to search
  // orthonormal frame of the epipolar plane: e1 towards the near end of the arc, e2 in the plane, towards the far end
  const Vector<3> normal = Normalized(u_near ^ u_far);
  const Vector<3> e1 = u_near, e2 = normal ^ e1;
  auto in_plane = [&](const Vector<3>& v) { return makeVector(e1 * v, e2 * v); };
  const double arc = std::acos(in_plane(u_far)[0]);
  const int n_steps = (int)std::ceil(arc / (one_pixel_angle * LevelScale(level) * 3));
  const double step = arc / n_steps;
  const Vector<2> a2 = in_plane(p_near), b2 = in_plane(p_far);
  Vector<2> along = b2 - a2;
  along = along * (1.0 / std::sqrt(along * along));
  for (int i = 0; i <= n_steps; ++i) {
    // direction (c, s) on the circle; the ray point a2 + t * along is parallel to it when their 2-D cross product vanishes
    const double c = std::cos(i * step), s = std::sin(i * step);
    const double t = (a2[0] * s - a2[1] * c) / (along[1] * c - along[0] * s);
    const Vector<3> p_t = p_near + dir_t * t;
    hypotheses.push_back(std::make_pair(world_from_tgt * p_t, p_t));
  }
  return true;
}

void PixelVectors(const SE3& src, const Vector<3>& v3WorldPos, const Vector<3>& c, const Vector<3>& r, const Vector<3>& d, Vector<3>& v3Right_W,
                  Vector<3>& v3Down_W)
{
  const Vector<3> n = makeVector(0, 0, -1);
  const Vector<3> v3PlanePoint_C = src * v3WorldPos;
  const double dCamHeight = std::fabs(v3PlanePoint_C * n);
  const Vector<3> cp = c * dCamHeight * (1.0 / std::fabs(c * n));
  const Vector<3> rp = r * dCamHeight * (1.0 / std::fabs(r * n));
  const Vector<3> dp = d * dCamHeight * (1.0 / std::fabs(d * n));
  v3Right_W = src.get_rotation().inverse() * (rp - cp);
  v3Down_W = src.get_rotation().inverse() * (dp - cp);
}

Vector<3> ReprojectPoint(const SE3& se3AfromB, const Vector<3>& v3A, const Vector<3>& v3B)
{
  double P[3][4];
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) P[i][j] = se3AfromB.get_rotation().get_matrix()[i][j];
    P[i][3] = se3AfromB.get_translation()[i];
  }
  double A[4][4] = { { -v3B[2], 0.0, v3B[0], 0.0 }, { 0.0, -v3B[2], v3B[1], 0.0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 } };
  for (int j = 0; j < 4; j++) {
    A[2][j] = v3A[0] * P[2][j] - v3A[2] * P[0][j];
    A[3][j] = v3A[1] * P[2][j] - v3A[2] * P[1][j];
  }
  // right singular vector of the smallest singular value = eigenvector of the smallest eigenvalue of A^T A (cyclic Jacobi)
  double S[4][4], V[4][4];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      S[i][j] = 0;
      for (int k = 0; k < 4; k++) S[i][j] += A[k][i] * A[k][j];
      V[i][j] = (i == j) ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0, diag = 0;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) (i == j ? diag : off) += S[i][j] * S[i][j];
    if (off <= 1e-60 * diag || off == 0) break;
    for (int p = 0; p < 3; p++)
      for (int q = p + 1; q < 4; q++) {
        if (S[p][q] == 0) continue;
        const double theta = (S[q][q] - S[p][p]) / (2 * S[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        const double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < 4; k++) { const double a = S[k][p], b = S[k][q]; S[k][p] = c * a - s * b; S[k][q] = s * a + c * b; }
        for (int k = 0; k < 4; k++) { const double a = S[p][k], b = S[q][k]; S[p][k] = c * a - s * b; S[q][k] = s * a + c * b; }
        for (int k = 0; k < 4; k++) { const double a = V[k][p], b = V[k][q]; V[k][p] = c * a - s * b; V[k][q] = s * a + c * b; }
      }
  }
  int m = 0;
  for (int i = 1; i < 4; i++) if (S[i][i] < S[m][m]) m = i;
  double v4[4] = { V[0][m], V[1][m], V[2][m], V[3][m] };
  if (v4[3] == 0.0) v4[3] = 0.00001;
  return makeVector(v4[0] / v4[3], v4[1] / v4[3], v4[2] / v4[3]);
}

namespace {

// the template cache of the ONE PatchFinder the reference function reuses (PatchFinder::MakeTemplateCoarseCont,
// src/PatchFinder.cc:135-182): the template is regenerated only when the warp matrix moved by more than 0.07 since the
// last GENERATED template.  Returns the index of the hypothesis whose warp the template in use was generated from.
struct TemplateCache {
  bool bHave = false;
  double m2Last[4] = { 0, 0, 0, 0 };
  int nFrom = -1;
  int Use(int i, const double* warp_inv, int nSearchLevel)
  {
    const double det = warp_inv[0] * warp_inv[3] - warp_inv[1] * warp_inv[2], idet = 1.0 / det, ls = LevelScale(nSearchLevel);
    const double m2[4] = { warp_inv[3] * idet * ls, -warp_inv[1] * idet * ls, -warp_inv[2] * idet * ls, warp_inv[0] * idet * ls };
    bool bRefresh = !bHave;
    for (int c = 0; !bRefresh && c < 2; c++) {
      const double d0 = m2[c] - m2Last[c], d1 = m2[2 + c] - m2Last[2 + c];      // column c of m2 (m2.T()[c])
      if (d0 * d0 + d1 * d1 > 0.07 * 0.07) bRefresh = true;
    }
    if (bRefresh) { bHave = true; for (int k = 0; k < 4; k++) m2Last[k] = m2[k]; nFrom = i; }
    return nFrom;
  }
};

// mcp_fe_search_patches in chunks of the handle's default request capacity (McpFeConfig::max_patches = 4096)
int SearchChunked(McpFe* fe, int nTargetSlot, const std::vector<McpPatchReq>& vReq, std::vector<McpPatchRes>& vRes)
{
  const size_t nChunk = 4096;
  vRes.resize(vReq.size());
  for (size_t b = 0; b < vReq.size(); b += nChunk) {
    const int n = (int)std::min(nChunk, vReq.size() - b);
    if (mcp_fe_search_patches(fe, nTargetSlot, n, vReq.data() + b, vRes.data() + b) != MCP_OK) return -1;
  }
  return 0;
}

struct Work {                 // per candidate
  std::vector<std::pair<Vector<3>, Vector<3> > > vPos;
  int nFirst = 0;             // index of its first hypothesis in the flat arrays
  TemplateCache cache;
  std::vector<int> vReqHyp;   // hypotheses that reached the coarse search, in order
  int nFirstReq = 0;
  std::vector<std::tuple<int, int, int, int> > vMatches;   // score, hypothesis, coarse x, coarse y (search level)
  int nFirstReq2 = 0, nReq2 = 0;
};

}  // namespace

int AddPointsEpipolar(McpFe* fe, int nSrcSlot, int nTargetSlot, const McpTaylorCam& camSrc, const McpTaylorCam& camTarget, const SE3& src,
                      const SE3& tgt, const unsigned char* pTargetMask, int nMaskStride, const std::vector<EpipolarCandidate>& vCandidates,
                      std::vector<EpipolarResult>& vResults)
{
  const int nCand = (int)vCandidates.size();
  vResults.assign(nCand, EpipolarResult());
  std::vector<Work> vWork(nCand);
  const double dOnePixelAngle = OnePixelAngle(camTarget);
  const int nW = (int)camTarget.image_size[0], nH = (int)camTarget.image_size[1];
  std::vector<double> vWorld, vRight, vDown;
  // ---- hypotheses of every candidate -------------------------------------------------------------------------
  for (int c = 0; c < nCand; c++) {
    const EpipolarCandidate& cand = vCandidates[c];
    EpipolarResult& r = vResults[c];
    Work& w = vWork[c];
    const int nLevelScale = LevelScale(cand.nLevel);
    r.v2RootPos = LevelZeroPos(cand.irLevelPos, cand.nLevel);
    const Vector<3> v3Ray_SC = UnProject(camSrc, r.v2RootPos);
    r.v3Center_NC = Normalized(v3Ray_SC);
    r.v3OneRightFromCenter_NC = Normalized(UnProject(camSrc, r.v2RootPos + makeVector(nLevelScale, 0)));
    r.v3OneDownFromCenter_NC = Normalized(UnProject(camSrc, r.v2RootPos + makeVector(0, nLevelScale)));
    w.nFirst = (int)(vWorld.size() / 3);
    if (!EpipolarHypotheses(src, tgt, v3Ray_SC, dOnePixelAngle, cand.nLevel, w.vPos)) { r.nReason = EPI_ENDPOINTS; continue; }
    r.nSteps = (int)w.vPos.size() - 1;
    for (size_t i = 0; i < w.vPos.size(); i++) {
      Vector<3> rw, dw;
      PixelVectors(src, w.vPos[i].first, r.v3Center_NC, r.v3OneRightFromCenter_NC, r.v3OneDownFromCenter_NC, rw, dw);
      for (int k = 0; k < 3; k++) { vWorld.push_back(w.vPos[i].first[k]); vRight.push_back(rw[k]); vDown.push_back(dw[k]); }
    }
  }
  const int nHyp = (int)(vWorld.size() / 3);
  if (nHyp == 0) return 0;
  // ---- Project + GetProjectionDerivs + CalcSearchLevelAndWarpMatrix for all of them ------------------------------
  double tgtRt[12];
  tgt.pack(tgtRt);
  std::vector<McpProjRes> vProj(nHyp);
  if (mcp_fe_project_points(fe, tgtRt, nHyp, vWorld.data(), vRight.data(), vDown.data(), vProj.data()) != MCP_OK) return -1;
  // ---- first pass: which hypotheses search, and with the template generated from which hypothesis -----------------
  std::vector<McpPatchReq> vReq;
  for (int c = 0; c < nCand; c++) {
    Work& w = vWork[c];
    w.nFirstReq = (int)vReq.size();
    for (size_t i = 0; i < w.vPos.size(); i++) {
      const McpProjRes& pr = vProj[w.nFirst + i];
      if (!pr.in_image) continue;                                          // cameraTarget.Invalid() or outside
      const int ix = (int)pr.px[0], iy = (int)pr.px[1];                    // CVD::ir
      if (!(ix >= 0 && iy >= 0 && ix < nW && iy < nH)) continue;           // maLevels[0].image.in_image
      if (pTargetMask && pTargetMask[(size_t)iy * nMaskStride + ix] == 0) continue;
      if (pr.search_level == -1) continue;
      const int nFrom = w.cache.Use((int)i, pr.warp_inv, pr.search_level);
      McpPatchReq rq;
      rq.src_kf = nSrcSlot; rq.src_level = vCandidates[c].nLevel; rq.src_cx = vCandidates[c].irLevelPos.x; rq.src_cy = vCandidates[c].irLevelPos.y;
      for (int k = 0; k < 4; k++) rq.warp_inv[k] = vProj[w.nFirst + nFrom].warp_inv[k];
      rq.search_level = pr.search_level; rq.pred_x = ix; rq.pred_y = iy; rq.range = 3; rq.subpix_its = 0; rq.exhaustive = 0;
      vReq.push_back(rq);
      w.vReqHyp.push_back((int)i);
    }
  }
  std::vector<McpPatchRes> vRes;
  if (SearchChunked(fe, nTargetSlot, vReq, vRes) != 0) return -1;
  // ---- ambiguity rules (:783-818) and the requests of the second pass ----------------------------------------------
  const int nMaxZMSSD = 8 * 8 * 250 + 1;                                   // finder.mnMaxSSD + 1
  std::vector<McpPatchReq> vReq2;
  for (int c = 0; c < nCand; c++) {
    Work& w = vWork[c];
    EpipolarResult& r = vResults[c];
    if (r.nReason == EPI_ENDPOINTS) continue;
    int nBestZMSSD = nMaxZMSSD, nBest = -1;
    for (size_t k = 0; k < w.vReqHyp.size(); k++) {
      const McpPatchRes& ps = vRes[w.nFirstReq + k];
      if (ps.template_bad || !ps.found) continue;
      w.vMatches.push_back(std::make_tuple(ps.score, w.vReqHyp[k], ps.coarse_x, ps.coarse_y));
      if (ps.score < nBestZMSSD) { nBestZMSSD = ps.score; nBest = w.vReqHyp[k]; }
    }
    r.nMatches = (int)w.vMatches.size();
    r.nBest = nBest; r.nBestScore = nBestZMSSD;
    if (nBest == -1) { r.nReason = EPI_NO_MATCH; continue; }
    std::stable_sort(w.vMatches.begin(), w.vMatches.end(),
                     [](const std::tuple<int, int, int, int>& a, const std::tuple<int, int, int, int>& b) { return std::get<0>(a) < std::get<0>(b); });
    int nResizeTo = 1;
    for (size_t i = 1; i < w.vMatches.size(); ++i)
      if (std::get<0>(w.vMatches[i]) > nBestZMSSD * 0.9) nResizeTo++;       // as written in the reference (:803)
    if (nResizeTo > 3) { r.nReason = EPI_AMBIGUOUS_COUNT; w.vMatches.clear(); continue; }
    w.vMatches.resize(nResizeTo);
    bool bFar = false;
    for (size_t i = 1; i < w.vMatches.size(); ++i)
      if (std::abs(std::get<1>(w.vMatches[i]) - nBest) > 1) bFar = true;
    if (bFar) { r.nReason = EPI_AMBIGUOUS_INDEX; w.vMatches.clear(); continue; }
    // second pass: the same PatchFinder goes on (its template cache continues from the first pass)
    w.nFirstReq2 = (int)vReq2.size();
    for (size_t i = 0; i < w.vMatches.size(); ++i) {
      const int h = std::get<1>(w.vMatches[i]);
      const McpProjRes& pr = vProj[w.nFirst + h];
      const int nFrom = w.cache.Use(h, pr.warp_inv, pr.search_level);
      McpPatchReq rq;
      rq.src_kf = nSrcSlot; rq.src_level = vCandidates[c].nLevel; rq.src_cx = vCandidates[c].irLevelPos.x; rq.src_cy = vCandidates[c].irLevelPos.y;
      for (int k = 0; k < 4; k++) rq.warp_inv[k] = vProj[w.nFirst + nFrom].warp_inv[k];
      rq.search_level = pr.search_level; rq.pred_x = std::get<2>(w.vMatches[i]); rq.pred_y = std::get<3>(w.vMatches[i]);
      rq.range = 0; rq.subpix_its = 10; rq.exhaustive = 2;                 // SetSubPixPos(coarse match) + IterateSubPixToConvergence(kf, 10)
      vReq2.push_back(rq);
    }
    w.nReq2 = (int)w.vMatches.size();
  }
  std::vector<McpPatchRes> vRes2;
  if (SearchChunked(fe, nTargetSlot, vReq2, vRes2) != 0) return -1;
  // ---- first converged match wins; triangulate --------------------------------------------------------------------
  int nFound = 0;
  for (int c = 0; c < nCand; c++) {
    Work& w = vWork[c];
    EpipolarResult& r = vResults[c];
    if (w.nReq2 == 0) continue;
    r.nReason = EPI_SUBPIX;
    for (int i = 0; i < w.nReq2; i++) {
      const McpPatchRes& ps = vRes2[w.nFirstReq2 + i];
      if (ps.template_bad || !ps.found) continue;                          // sub-pixel iteration did not converge
      r.v2SubPixPos = makeVector(ps.found_x, ps.found_y);
      r.nSubPixFrom = std::get<1>(w.vMatches[i]);
      const Vector<3> v3New_TC = ReprojectPoint(src * tgt.inverse(), UnProject(camSrc, r.v2RootPos), UnProject(camTarget, r.v2SubPixPos));
      r.v3WorldPos = tgt.inverse() * v3New_TC;
      r.bOK = true;
      r.nReason = EPI_OK;
      nFound++;
      break;
    }
  }
  return nFound;
}

}  // namespace mcp_host

// ---- C entry points for the tests (ctypes) ------------------------------------------------------------------------
extern "C" {

using namespace mcp_host;

struct McpHostEpiRes {
  int32_t ok, reason, n_steps, n_matches, best, best_score, subpix_from, pad_;
  double world[3], root[2], subpix[2];
};

int mcp_host_epi_hypotheses(const double* src12, const double* tgt12, const double* ray3, double one_pixel_angle, int level, int cap,
                            double* world3n, double* tc3n, double* start_end2)
{
  std::vector<std::pair<Vector<3>, Vector<3> > > v;
  const bool ok = EpipolarHypotheses(SE3::unpack(src12), SE3::unpack(tgt12), makeVector(ray3[0], ray3[1], ray3[2]), one_pixel_angle, level, v,
                                     start_end2, start_end2 ? start_end2 + 1 : nullptr);
  if (!ok) return -1;
  for (int i = 0; i < (int)v.size() && i < cap; i++)
    for (int k = 0; k < 3; k++) { world3n[3 * i + k] = v[i].first[k]; tc3n[3 * i + k] = v[i].second[k]; }
  return (int)v.size();
}
void mcp_host_reproject_point(const double* a_from_b12, const double* va3, const double* vb3, double* out3)
{
  const Vector<3> p = ReprojectPoint(SE3::unpack(a_from_b12), makeVector(va3[0], va3[1], va3[2]), makeVector(vb3[0], vb3[1], vb3[2]));
  for (int k = 0; k < 3; k++) out3[k] = p[k];
}
void mcp_host_unproject(const McpTaylorCam* cam, const double* px2, double* out3)
{
  const Vector<3> r = UnProject(*cam, makeVector(px2[0], px2[1]));
  for (int k = 0; k < 3; k++) out3[k] = r[k];
}
double mcp_host_one_pixel_angle(const McpTaylorCam* cam) { return OnePixelAngle(*cam); }
void mcp_host_pixel_vectors(const double* src12, const double* world3, const double* c3, const double* r3, const double* d3, double* right3, double* down3)
{
  Vector<3> rw, dw;
  PixelVectors(SE3::unpack(src12), makeVector(world3[0], world3[1], world3[2]), makeVector(c3[0], c3[1], c3[2]), makeVector(r3[0], r3[1], r3[2]),
               makeVector(d3[0], d3[1], d3[2]), rw, dw);
  for (int k = 0; k < 3; k++) { right3[k] = rw[k]; down3[k] = dw[k]; }
}
// level_xy: 3 ints per candidate {level, x, y}
int mcp_host_add_points_epipolar(McpFe* fe, int src_slot, int tgt_slot, const McpTaylorCam* cam_src, const McpTaylorCam* cam_tgt, const double* src12,
                                 const double* tgt12, const unsigned char* mask, int mask_stride, int n, const int32_t* level_xy, McpHostEpiRes* out)
{
  std::vector<EpipolarCandidate> vc(n);
  for (int i = 0; i < n; i++) { vc[i].nLevel = level_xy[3 * i]; vc[i].irLevelPos = ImageRef(level_xy[3 * i + 1], level_xy[3 * i + 2]); }
  std::vector<EpipolarResult> vr;
  const int rc = AddPointsEpipolar(fe, src_slot, tgt_slot, *cam_src, *cam_tgt, SE3::unpack(src12), SE3::unpack(tgt12), mask, mask_stride, vc, vr);
  if (rc < 0) return rc;
  for (int i = 0; i < n; i++) {
    McpHostEpiRes& o = out[i];
    o.ok = vr[i].bOK; o.reason = vr[i].nReason; o.n_steps = vr[i].nSteps; o.n_matches = vr[i].nMatches; o.best = vr[i].nBest;
    o.best_score = vr[i].nBestScore; o.subpix_from = vr[i].nSubPixFrom; o.pad_ = 0;
    for (int k = 0; k < 3; k++) o.world[k] = vr[i].v3WorldPos[k];
    for (int k = 0; k < 2; k++) { o.root[k] = vr[i].v2RootPos[k]; o.subpix[k] = vr[i].v2SubPixPos[k]; }
  }
  return rc;
}

}  // extern "C"
