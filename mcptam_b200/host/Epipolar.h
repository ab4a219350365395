// Epipolar.h — MapMakerServerBase::AddPointEpipolar (src/MapMakerServerBase.cc:604-914) on the B200 path, batched:
// ALL candidates of one (source keyframe, target keyframe) pair go through two device round trips
//   1. mcp_fe_project_points  -- every depth hypothesis of every candidate (Project + GetProjectionDerivs +
//                                CalcSearchLevelAndWarpMatrix), then mcp_fe_search_patches with range 3 (coarse only);
//   2. mcp_fe_search_patches  -- the sub-pixel refinement of the <= 3 surviving matches per candidate
// with the reference's host logic in between (epipolar-arc hypotheses, the template cache of the one PatchFinder the
// reference reuses, the ambiguity rules, Hartley-Zisserman triangulation).  SURVEY.md §8 f-3, second half.
#pragma once

#include <vector>

#include "TaylorCamera.h"
#include "shim/MapTypes.h"

namespace mcp_host {

struct EpipolarCandidate {
  int nLevel = 0;            // pyramid level of the candidate (Level::vCandidates[n].irLevelPos lives at this level)
  ImageRef irLevelPos;
};

enum EpipolarReason { EPI_OK = 0, EPI_ENDPOINTS = 1, EPI_NO_MATCH = 2, EPI_AMBIGUOUS_COUNT = 3, EPI_AMBIGUOUS_INDEX = 4, EPI_SUBPIX = 5 };

struct EpipolarResult {
  bool bOK = false;          // the reference's return value
  int nReason = EPI_NO_MATCH;
  Vector<3> v3WorldPos;      // pPointNew->mv3WorldPos
  Vector<2> v2RootPos;       // measurement in the source keyframe (SRC_ROOT), level-0 pixels
  Vector<2> v2SubPixPos;     // measurement in the target keyframe (SRC_EPIPOLAR)
  Vector<3> v3Center_NC, v3OneRightFromCenter_NC, v3OneDownFromCenter_NC;   // patch-source fields of the new point
  int nSteps = 0, nMatches = 0, nBest = -1, nBestScore = 0, nSubPixFrom = -1;
};

// ---- pure host pieces (no device), exposed for the tests ------------------------------------------------------
// TaylorCamera::UnProject / OnePixelAngle from the ABI camera record (src/TaylorCamera.cc:319-346, :192-196)
Vector<3> UnProject(const McpTaylorCam& cam, const Vector<2>& v2Im);
double OnePixelAngle(const McpTaylorCam& cam);
// :620-724.  Positions: first in the world frame, second in the target camera frame.  false = "return false".
bool EpipolarHypotheses(const SE3& se3SrcCamFromWorld, const SE3& se3TgtCamFromWorld, const Vector<3>& v3Ray_SC, double dOnePixelAngle,
                        int nLevel, std::vector<std::pair<Vector<3>, Vector<3> > >& vPositions, double* pdStartDepth = nullptr,
                        double* pdEndDepth = nullptr);
// MapPoint::RefreshPixelVectors (src/MapPoint.cc:62-87) with mv3Normal_NC = (0,0,-1)
void PixelVectors(const SE3& se3SrcCamFromWorld, const Vector<3>& v3WorldPos, const Vector<3>& v3Center_NC, const Vector<3>& v3Right_NC,
                  const Vector<3>& v3Down_NC, Vector<3>& v3PixelRight_W, Vector<3>& v3PixelDown_W);
// MapMakerServerBase::ReprojectPoint (:123-143); [3P] TooN SVD<4,4> -> Jacobi eigenvectors of A^T A
Vector<3> ReprojectPoint(const SE3& se3AfromB, const Vector<3>& v3A, const Vector<3>& v3B);

// ---- the batched search ----------------------------------------------------------------------------------------
// fe: the TARGET camera's front-end handle (its camera model set with mcp_fe_set_camera); the source keyframe's pyramid
// is resident in slot nSrcSlot of the same handle, the target keyframe's in nTargetSlot.  pTargetMask: level-0 mask of
// the target keyframe (0 = masked) or NULL.  Returns the number of points found, < 0 on a device error.
int AddPointsEpipolar(McpFe* fe, int nSrcSlot, int nTargetSlot, const McpTaylorCam& camSrc, const McpTaylorCam& camTarget,
                      const SE3& se3SrcCamFromWorld, const SE3& se3TgtCamFromWorld, const unsigned char* pTargetMask, int nMaskStride,
                      const std::vector<EpipolarCandidate>& vCandidates, std::vector<EpipolarResult>& vResults);

}  // namespace mcp_host
