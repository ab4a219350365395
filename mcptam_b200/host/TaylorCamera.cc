// TaylorCamera.cc — host mirror of src/TaylorCamera.cc (construction-time fit + scalar project/unproject).
#include "TaylorCamera.h"

#include <algorithm>
#include <cmath>
#include <cstdio>

namespace mcp_host {

static const double kPi = 3.14159265358979323846;

double TaylorCamera::PolyVal(const double* c, int n, double x)
{
  double val = 0;
  for (int i = n - 1; i > 0; i--) { val += c[i]; val *= x; }
  return val + c[0];
}

TaylorCamera::TaylorCamera(Vector<9> v9Params, ImageRef irCalibSize, ImageRef irFullScaleSize, ImageRef irImageSize)
{
  mv9CameraParams = v9Params;
  mv2CalibSize = makeVector(irCalibSize.x, irCalibSize.y);
  mv2FullScaleSize = makeVector(irFullScaleSize.x, irFullScaleSize.y);
  mv2ImageSize = makeVector(irImageSize.x, irImageSize.y);
  RefreshParams();
}
void TaylorCamera::SetImageSize(ImageRef s) { mv2ImageSize = makeVector(s.x, s.y); RefreshParams(); }

void TaylorCamera::RefreshParams()
{
  const Vector<9>& p = mv9CameraParams;
  mv5PolyCoeffs[0] = p[0]; mv5PolyCoeffs[1] = 0; mv5PolyCoeffs[2] = p[1]; mv5PolyCoeffs[3] = p[2]; mv5PolyCoeffs[4] = p[3];
  for (int i = 0; i < 5; i++) mv5PolyDerivModCoeffs[i] = mv5PolyCoeffs[i];
  mv5PolyDerivModCoeffs[0] *= -1; mv5PolyDerivModCoeffs[3] *= 2; mv5PolyDerivModCoeffs[4] *= 3;
  const Vector<2> scale = makeVector(mv2ImageSize[0] / mv2FullScaleSize[0], mv2ImageSize[1] / mv2FullScaleSize[1]);
  const Vector<2> fsc = makeVector(p[4] - (mv2CalibSize[0] - mv2FullScaleSize[0]) / 2, p[5] - (mv2CalibSize[1] - mv2FullScaleSize[1]) / 2);
  mv2Center = makeVector(fsc[0] * scale[0], fsc[1] * scale[1]);
  const Vector<2> corner = makeVector(std::max(fsc[0], mv2FullScaleSize[0] - fsc[0] - 1), std::max(fsc[1], mv2FullScaleSize[1] - fsc[1] - 1));
  mdLargestRadius = std::sqrt(corner * corner);
  mdMaxRho = 1.0 * mdLargestRadius;
  mdMinTheta = std::atan(PolyVal(mv5PolyCoeffs, 5, mdMaxRho) / mdMaxRho);
  mvxPolyInvCoeffs = FindInvPolyUsingRoots(-1, 0.0001);
  mbUsingInversePoly = !mvxPolyInvCoeffs.empty();
  if (!mbUsingInversePoly)
    // The reference falls back to a linear inverse + Newton iterations here (src/TaylorCamera.cc:159-175, 262-267: 'slow').  The
    // device path evaluates the inverse polynomial only, so such a calibration is refused HERE, by name, instead of surfacing
    // later as a bundle adjustment that returns -1 ('map corrupt').
    std::fprintf(stderr, "TaylorCamera: no inverse polynomial of degree <= %d fits this calibration to 1e-4 px; the B200 path does not implement "
                         "the reference's Newton fallback -- Good() is false, projections are invalid, and ChainBundle refuses the camera\n", MAX_INV_DEGREE);
  mm2Affine[0][0] = scale[0] * p[6]; mm2Affine[0][1] = scale[1] * p[7];
  mm2Affine[1][0] = scale[0] * p[8]; mm2Affine[1][1] = scale[1] * 1;
  const double det = mm2Affine[0][0] * mm2Affine[1][1] - mm2Affine[0][1] * mm2Affine[1][0], id = 1.0 / det;   // opts::M2Inverse
  mm2AffineInv[0][0] = mm2Affine[1][1] * id; mm2AffineInv[1][1] = mm2Affine[0][0] * id;
  mm2AffineInv[1][0] = -mm2Affine[1][0] * id; mm2AffineInv[0][1] = -mm2Affine[0][1] * id;
  if (mbUsingInversePoly) {
    const Vector<3> c = UnProject(mv2ImageSize * 0.5), r = UnProject(mv2ImageSize * 0.5 + makeVector(1, 1));
    mdOnePixelAngle = std::acos(c * r) / std::sqrt(2.0);
  }
}

// real roots of c3 x^3 + c2 x^2 + c1 x + c0 (degenerates handled), ascending
static int cubic_roots(double c3, double c2, double c1, double c0, double* r)
{
  int n = 0;
  if (std::fabs(c3) < 1e-300) {
    if (std::fabs(c2) < 1e-300) { if (std::fabs(c1) > 1e-300) r[n++] = -c0 / c1; return n; }
    const double disc = c1 * c1 - 4 * c2 * c0;
    if (disc >= 0) { const double s = std::sqrt(disc), q = -0.5 * (c1 + (c1 >= 0 ? s : -s)); r[n++] = q / c2; if (q != 0) r[n++] = c0 / q; }
    std::sort(r, r + n);
    return n;
  }
  const double a = c2 / c3, b = c1 / c3, c = c0 / c3;
  const double Q = (a * a - 3 * b) / 9, R = (2 * a * a * a - 9 * a * b + 27 * c) / 54;
  if (R * R < Q * Q * Q) {
    const double th = std::acos(R / std::sqrt(Q * Q * Q)), sq = -2 * std::sqrt(Q);
    r[0] = sq * std::cos(th / 3) - a / 3; r[1] = sq * std::cos((th + 2 * kPi) / 3) - a / 3; r[2] = sq * std::cos((th - 2 * kPi) / 3) - a / 3;
    n = 3;
  } else {
    const double A = -(R >= 0 ? 1.0 : -1.0) * std::cbrt(std::fabs(R) + std::sqrt(R * R - Q * Q * Q));
    const double B = (A != 0) ? Q / A : 0;
    r[0] = (A + B) - a / 3;
    n = 1;
  }
  std::sort(r, r + n);
  return n;
}

// number of real roots of the quartic in [0, max_rho] and the root itself when unique
static int quartic_roots_in_range(const double* q /*x^0 first*/, double max_rho, double* root)
{
  auto f = [&](double x) { return (((q[4] * x + q[3]) * x + q[2]) * x + q[1]) * x + q[0]; };
  double crit[3];
  const int nc = cubic_roots(4 * q[4], 3 * q[3], 2 * q[2], q[1], crit);
  double br[6];
  int nb = 0;
  br[nb++] = 0.0;
  for (int i = 0; i < nc; i++) if (crit[i] > 0.0 && crit[i] < max_rho) br[nb++] = crit[i];
  br[nb++] = max_rho;
  int count = 0;
  for (int i = 0; i + 1 < nb; i++) {
    double lo = br[i], hi = br[i + 1];
    double flo = f(lo), fhi = f(hi);
    if (flo == 0.0) { if (i == 0) { *root = lo; count++; } continue; }
    if (fhi == 0.0) { *root = hi; count++; continue; }
    if ((flo < 0) == (fhi < 0)) continue;
    for (int it = 0; it < 200 && hi - lo > 1e-14 * std::max(1.0, std::fabs(hi)); it++) {
      const double mid = 0.5 * (lo + hi), fm = f(mid);
      if ((fm < 0) == (flo < 0)) { lo = mid; flo = fm; } else { hi = mid; }
    }
    *root = 0.5 * (lo + hi);
    count++;
  }
  return count;
}

// least-squares polynomial fit by Householder QR on the Vandermonde matrix (the reference uses TooN SVD backsub)
static std::vector<double> polyfit(const std::vector<double>& x, const std::vector<double>& y, int deg)
{
  const int m = (int)x.size(), n = deg + 1;
  std::vector<double> A((size_t)m * n), b(y);
  for (int i = 0; i < m; i++) { double p = 1; for (int j = 0; j < n; j++) { A[(size_t)i * n + j] = p; p *= x[i]; } }
  for (int k = 0; k < n; k++) {
    double nrm = 0;
    for (int i = k; i < m; i++) nrm += A[(size_t)i * n + k] * A[(size_t)i * n + k];
    nrm = std::sqrt(nrm);
    if (nrm == 0) continue;
    const double alpha = A[(size_t)k * n + k] > 0 ? -nrm : nrm;
    std::vector<double> v(m - k);
    for (int i = k; i < m; i++) v[i - k] = A[(size_t)i * n + k];
    v[0] -= alpha;
    double vn = 0;
    for (double t : v) vn += t * t;
    if (vn == 0) continue;
    for (int j = k; j < n; j++) {
      double s = 0;
      for (int i = k; i < m; i++) s += v[i - k] * A[(size_t)i * n + j];
      s = 2 * s / vn;
      for (int i = k; i < m; i++) A[(size_t)i * n + j] -= s * v[i - k];
    }
    double s = 0;
    for (int i = k; i < m; i++) s += v[i - k] * b[i];
    s = 2 * s / vn;
    for (int i = k; i < m; i++) b[i] -= s * v[i - k];
  }
  std::vector<double> c(n);
  for (int k = n - 1; k >= 0; k--) {
    double s = b[k];
    for (int j = k + 1; j < n; j++) s -= A[(size_t)k * n + j] * c[j];
    c[k] = s / A[(size_t)k * n + k];
  }
  return c;
}

std::vector<double> TaylorCamera::FindInvPolyUsingRoots(int nSpecifiedDegree, double dErrorLimit)
{
  const double dThetaStart = -kPi / 2 + 0.001, dThetaEnd = kPi / 2 - 0.001, dThetaStep = 0.01;
  const int nThetaNum = (int)std::ceil((dThetaEnd - dThetaStart) / dThetaStep) + 1;
  std::vector<double> th, rho;
  double t = dThetaStart;
  for (int i = 0; i < nThetaNum; i++, t += dThetaStep) {
    double q[5] = { mv5PolyCoeffs[0], mv5PolyCoeffs[1] - std::tan(t), mv5PolyCoeffs[2], mv5PolyCoeffs[3], mv5PolyCoeffs[4] };
    double root = 0;
    if (quartic_roots_in_range(q, mdMaxRho, &root) == 1) { th.push_back(t); rho.push_back(root); }
  }
  if (th.size() < 3) return std::vector<double>();
  mdThetaMean = 0;
  for (double v : th) mdThetaMean += v;
  mdThetaMean /= th.size();
  double ss = 0;
  for (double v : th) ss += (v - mdThetaMean) * (v - mdThetaMean);
  mdThetaStd = std::sqrt(ss / th.size());
  std::vector<double> x(th.size());
  for (size_t i = 0; i < th.size(); i++) x[i] = (th[i] - mdThetaMean) / mdThetaStd;
  if (nSpecifiedDegree >= 0) return polyfit(x, rho, nSpecifiedDegree);
  for (int deg = 2; deg <= MAX_INV_DEGREE; deg++) {
    std::vector<double> c = polyfit(x, rho, deg);
    double mx = 0;
    for (size_t i = 0; i < x.size(); i++) mx = std::max(mx, std::fabs(rho[i] - PolyVal(c.data(), (int)c.size(), x[i])));
    if (mx <= dErrorLimit) return c;
  }
  return std::vector<double>();
}

Vector<2> TaylorCamera::Project(const Vector<3>& v)
{
  mv3LastCam = v;
  const double dNorm = std::sqrt(v[0] * v[0] + v[1] * v[1]);
  double dTheta;
  if (dNorm == 0) dTheta = kPi / 2; else dTheta = std::atan(v[2] / dNorm);
  mbInvalid = (dTheta < mdMinTheta);
  if (dNorm == 0) { mdLastRho = 0; mdLastCosPhi = 0; mdLastSinPhi = 0; }
  else {
    mdLastRho = PolyVal(mvxPolyInvCoeffs.data(), (int)mvxPolyInvCoeffs.size(), (dTheta - mdThetaMean) / mdThetaStd);
    mdLastCosPhi = v[0] / dNorm; mdLastSinPhi = v[1] / dNorm;
  }
  mv2LastDistCam = makeVector(mdLastCosPhi * mdLastRho, mdLastSinPhi * mdLastRho);
  mv2LastIm = mm2Affine * mv2LastDistCam + mv2Center;
  if (!(mv2LastIm[0] >= 0 && mv2LastIm[0] < mv2ImageSize[0] && mv2LastIm[1] >= 0 && mv2LastIm[1] < mv2ImageSize[1])) mbInvalid = true;
  return mv2LastIm;
}

Vector<3> TaylorCamera::UnProject(const Vector<2>& im)
{
  mv2LastIm = im;
  mv2LastDistCam = mm2AffineInv * (im - mv2Center);
  mdLastRho = std::sqrt(mv2LastDistCam * mv2LastDistCam);
  mv3LastCam = makeVector(mv2LastDistCam[0], mv2LastDistCam[1], PolyVal(mv5PolyCoeffs, 5, mdLastRho));
  if (mdLastRho == 0) { mdLastCosPhi = 0; mdLastSinPhi = 0; }
  else { mdLastCosPhi = mv3LastCam[0] / mdLastRho; mdLastSinPhi = mv3LastCam[1] / mdLastRho; }
  const double n = std::sqrt(mv3LastCam * mv3LastCam);
  for (int k = 0; k < 3; k++) mv3LastCam[k] /= n;          // TooN::normalize: v /= sqrt(v * v), element-wise division
  return mv3LastCam;
}

Matrix<2> TaylorCamera::GetProjectionDerivs()
{
  const double w = PolyVal(mv5PolyCoeffs, 5, mdLastRho);
  const double dRho_dTheta = (mdLastRho * mdLastRho + w * w) / PolyVal(mv5PolyDerivModCoeffs, 5, mdLastRho);
  const Vector<2> dTh = makeVector(mdLastCosPhi * dRho_dTheta, mdLastSinPhi * dRho_dTheta);
  const Vector<2> dPh = makeVector(-mdLastSinPhi * mdLastRho, mdLastCosPhi * mdLastRho);
  const Vector<2> a = mm2Affine * dTh, b = mm2Affine * dPh;
  Matrix<2> m;
  m[0][0] = a[0]; m[1][0] = a[1]; m[0][1] = b[0]; m[1][1] = b[1];
  return m;
}

void TaylorCamera::GetCamSphereDeriv(const Vector<3>& v, Vector<3>& dth, Vector<3>& dph)
{
  const double x = v[0], y = v[1], z = v[2], x2 = x * x, y2 = y * y, z2 = z * z;
  const double n = std::sqrt(x * x + y * y), n2 = n * n, n3 = n2 * n;
  if (n == 0) dth = makeVector(0, 0, 0);
  else dth = makeVector(-z * x / (n3 + n * z2), -z * y / (n3 + n * z2), n / (n2 + z2));
  if (x == 0 && y == 0) dph = makeVector(0, 0, 0);
  else dph = makeVector(-y / (x2 + y2), x / (x2 + y2), 0);
}

McpTaylorCam TaylorCamera::ToAbi() const
{
  McpTaylorCam c;
  std::memset(&c, 0, sizeof(c));
  for (int i = 0; i < 5; i++) c.poly[i] = mv5PolyCoeffs[i];
  c.center[0] = mv2Center[0]; c.center[1] = mv2Center[1];
  c.affine[0] = mm2Affine[0][0]; c.affine[1] = mm2Affine[0][1]; c.affine[2] = mm2Affine[1][0]; c.affine[3] = mm2Affine[1][1];
  c.image_size[0] = mv2ImageSize[0]; c.image_size[1] = mv2ImageSize[1];
  c.min_theta = mdMinTheta; c.theta_mean = mdThetaMean; c.theta_std = mdThetaStd;
  c.n_inv = (int32_t)mvxPolyInvCoeffs.size();
  for (size_t i = 0; i < mvxPolyInvCoeffs.size() && i < 32; i++) c.inv_poly[i] = mvxPolyInvCoeffs[i];
  return c;
}

}  // namespace mcp_host

// ---- C entry points for the CPU tests (ctypes): the camera mirror against the Python / C restatements -------------
extern "C" {

using namespace mcp_host;

// Builds a TaylorCamera from the 9 calibration parameters and returns its ABI record; 0 on success
int mcp_host_camera_abi(const double* params9, int w, int h, McpTaylorCam* out)
{
  Vector<9> p;
  for (int i = 0; i < 9; i++) p[i] = params9[i];
  TaylorCamera cam(p, ImageRef(w, h), ImageRef(w, h), ImageRef(w, h));
  if (!cam.Good()) return -1;
  *out = cam.ToAbi();
  return 0;
}
// Project + GetProjectionDerivs of n camera-frame points; invalid[i] = TaylorCamera::Invalid()
int mcp_host_camera_project(const double* params9, int w, int h, int n, const double* p3, double* px2, double* derivs4, int* invalid,
                            double* one_pixel_angle)
{
  Vector<9> p;
  for (int i = 0; i < 9; i++) p[i] = params9[i];
  TaylorCamera cam(p, ImageRef(w, h), ImageRef(w, h), ImageRef(w, h));
  if (!cam.Good()) return -1;
  for (int i = 0; i < n; i++) {
    const Vector<2> v = cam.Project(makeVector(p3[3 * i], p3[3 * i + 1], p3[3 * i + 2]));
    invalid[i] = cam.Invalid() ? 1 : 0;
    const Matrix<2> d = cam.GetProjectionDerivs();
    px2[2 * i] = v[0]; px2[2 * i + 1] = v[1];
    derivs4[4 * i] = d[0][0]; derivs4[4 * i + 1] = d[0][1]; derivs4[4 * i + 2] = d[1][0]; derivs4[4 * i + 3] = d[1][1];
  }
  if (one_pixel_angle) *one_pixel_angle = cam.OnePixelAngle();
  return 0;
}
int mcp_host_camera_unproject(const double* params9, int w, int h, int n, const double* px2, double* ray3)
{
  Vector<9> p;
  for (int i = 0; i < 9; i++) p[i] = params9[i];
  TaylorCamera cam(p, ImageRef(w, h), ImageRef(w, h), ImageRef(w, h));
  if (!cam.Good()) return -1;
  for (int i = 0; i < n; i++) {
    const Vector<3> r = cam.UnProject(makeVector(px2[2 * i], px2[2 * i + 1]));
    for (int k = 0; k < 3; k++) ray3[3 * i + k] = r[k];
  }
  return 0;
}

}  // extern "C"
