// TaylorCamera.h — host mirror of the reference camera model (include/mcptam/TaylorCamera.h:90-338,
// src/TaylorCamera.cc): same public surface for the members the hot path uses.  RefreshParams() derives the
// inverse polynomial exactly like FindInvPolyUsingRoots (theta grid, unique real root in [0, maxRho], centred /
// scaled least-squares fit of increasing degree until the fit error is <= 1e-4) and ToAbi() hands the derived
// quantities to the device through McpTaylorCam.
#pragma once

#include <map>
#include <string>
#include <vector>

#include "../../include/mcptam_b200.h"
#include "shim/Types.h"

#define MAX_INV_DEGREE 30

namespace mcp_host {

using namespace mcp_shim;

class TaylorCamera {
 public:
  TaylorCamera() {}
  TaylorCamera(Vector<9> v9Params, ImageRef irCalibSize, ImageRef irFullScaleSize, ImageRef irImageSize);
  void SetImageSize(ImageRef irImageSize);
  ImageRef GetImageSize() const { return ImageRef((int)mv2ImageSize[0], (int)mv2ImageSize[1]); }
  Vector<2> Project(const Vector<3>& v3CamFrame);              // src/TaylorCamera.cc:202-287
  Vector<3> UnProject(const Vector<2>& v2ImFrame);             // :319-346
  Matrix<2> GetProjectionDerivs();                             // :353-383
  static void GetCamSphereDeriv(const Vector<3>& v3Cam, Vector<3>& v3_dTheta, Vector<3>& v3_dPhi);   // :617-669
  bool Invalid() const { return mbInvalid; }
  double OnePixelAngle() const { return mdOnePixelAngle; }
  const std::vector<double>& InvPoly() const { return mvxPolyInvCoeffs; }
  McpTaylorCam ToAbi() const;
  bool Good() const { return mbUsingInversePoly; }

 protected:
  void RefreshParams();                                        // :84-198
  std::vector<double> FindInvPolyUsingRoots(int nSpecifiedDegree, double dErrorLimit);   // :489-604
  static double PolyVal(const double* c, int n, double x);     // :472-485

  Vector<9> mv9CameraParams;
  double mv5PolyCoeffs[5], mv5PolyDerivModCoeffs[5];
  Vector<2> mv2CalibSize, mv2FullScaleSize, mv2ImageSize, mv2Center;
  Matrix<2> mm2Affine, mm2AffineInv;
  double mdLargestRadius = 0, mdMaxRho = 0, mdMinTheta = 0, mdThetaMean = 0, mdThetaStd = 1, mdOnePixelAngle = 0;
  std::vector<double> mvxPolyInvCoeffs;
  bool mbUsingInversePoly = false, mbInvalid = false;
  // cached last projection (as in the reference; one camera object per thread)
  Vector<3> mv3LastCam;
  Vector<2> mv2LastIm, mv2LastDistCam;
  double mdLastRho = 0, mdLastCosPhi = 0, mdLastSinPhi = 0;
};

typedef std::map<std::string, TaylorCamera> TaylorCameraMap;

}  // namespace mcp_host
