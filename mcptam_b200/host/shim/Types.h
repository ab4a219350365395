// shim/Types.h — minimal stand-ins for the TooN / libCVD types that appear at the hot-path boundary.
// In a real MCPTAM tree define MCPTAM_HAVE_TOON_CVD and include <TooN/se3.h>, <cvd/image.h> instead; the
// host mirror only uses the members below (operator[], get_rotation().get_matrix(), get_translation(), size(), ...).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace mcp_shim {

template <int N> struct Vector {
  double v[N];
  Vector() { for (int i = 0; i < N; i++) v[i] = 0; }
  double& operator[](int i) { return v[i]; }
  const double& operator[](int i) const { return v[i]; }
  Vector operator+(const Vector& o) const { Vector r; for (int i = 0; i < N; i++) r.v[i] = v[i] + o.v[i]; return r; }
  Vector operator-(const Vector& o) const { Vector r; for (int i = 0; i < N; i++) r.v[i] = v[i] - o.v[i]; return r; }
  Vector operator*(double s) const { Vector r; for (int i = 0; i < N; i++) r.v[i] = v[i] * s; return r; }
  double operator*(const Vector& o) const { double s = 0; for (int i = 0; i < N; i++) s += v[i] * o.v[i]; return s; }
};
inline Vector<2> makeVector(double a, double b) { Vector<2> r; r[0] = a; r[1] = b; return r; }
inline Vector<3> makeVector(double a, double b, double c) { Vector<3> r; r[0] = a; r[1] = b; r[2] = c; return r; }
inline Vector<3> operator^(const Vector<3>& a, const Vector<3>& b)
{
  return makeVector(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}

template <int R, int C = R> struct Matrix {
  double m[R][C];
  Matrix() { std::memset(m, 0, sizeof(m)); }
  double* operator[](int r) { return m[r]; }
  const double* operator[](int r) const { return m[r]; }
};
inline Vector<3> operator*(const Matrix<3>& A, const Vector<3>& x)
{
  Vector<3> r;
  for (int i = 0; i < 3; i++) r[i] = A[i][0] * x[0] + A[i][1] * x[1] + A[i][2] * x[2];
  return r;
}
inline Vector<2> operator*(const Matrix<2>& A, const Vector<2>& x) { return makeVector(A[0][0] * x[0] + A[0][1] * x[1], A[1][0] * x[0] + A[1][1] * x[1]); }
inline Matrix<3> operator*(const Matrix<3>& A, const Matrix<3>& B)
{
  Matrix<3> r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r[i][j] = A[i][0] * B[0][j] + A[i][1] * B[1][j] + A[i][2] * B[2][j];
  return r;
}

// TooN::SO3 / SE3 subset: left-multiplicative composition, exp with translation first (TooN se3.h)
struct SO3 {
  Matrix<3> R;
  SO3() { R[0][0] = R[1][1] = R[2][2] = 1; }
  const Matrix<3>& get_matrix() const { return R; }
  Matrix<3>& get_matrix() { return R; }
  SO3 inverse() const { SO3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.R[i][j] = R[j][i]; return r; }
  Vector<3> operator*(const Vector<3>& x) const { return R * x; }
  SO3 operator*(const SO3& o) const { SO3 r; r.R = R * o.R; return r; }
  static SO3 exp(const Vector<3>& w)
  {
    const double one_6th = 1.0 / 6.0, one_20th = 1.0 / 20.0;
    const double tsq = w * w, th = std::sqrt(tsq);
    double A, B;
    if (tsq < 1e-8) { A = 1.0 - one_6th * tsq; B = 0.5; }
    else if (tsq < 1e-6) { B = 0.5 - 0.25 * one_6th * tsq; A = 1.0 - tsq * one_6th * (1.0 - one_20th * tsq); }
    else { const double it = 1.0 / th; A = std::sin(th) * it; B = (1 - std::cos(th)) * (it * it); }
    SO3 r;
    rodrigues(w, A, B, r.R);
    return r;
  }
  static void rodrigues(const Vector<3>& w, double A, double B, Matrix<3>& R)
  {
    const double wx2 = w[0] * w[0], wy2 = w[1] * w[1], wz2 = w[2] * w[2];
    R[0][0] = 1.0 - B * (wy2 + wz2); R[1][1] = 1.0 - B * (wx2 + wz2); R[2][2] = 1.0 - B * (wx2 + wy2);
    { const double a = A * w[2], b = B * (w[0] * w[1]); R[0][1] = b - a; R[1][0] = b + a; }
    { const double a = A * w[1], b = B * (w[0] * w[2]); R[0][2] = b + a; R[2][0] = b - a; }
    { const double a = A * w[0], b = B * (w[1] * w[2]); R[1][2] = b - a; R[2][1] = b + a; }
  }
};
struct SE3 {
  SO3 rot;
  Vector<3> trans;
  const SO3& get_rotation() const { return rot; }
  SO3& get_rotation() { return rot; }
  const Vector<3>& get_translation() const { return trans; }
  Vector<3>& get_translation() { return trans; }
  Vector<3> operator*(const Vector<3>& x) const { return rot * x + trans; }
  SE3 operator*(const SE3& o) const { SE3 r; r.rot = rot * o.rot; r.trans = rot * o.trans + trans; return r; }
  SE3 inverse() const { SE3 r; r.rot = rot.inverse(); r.trans = (r.rot * trans) * -1.0; return r; }
  static SE3 exp(const Vector<6>& mu)
  {
    const double one_6th = 1.0 / 6.0, one_20th = 1.0 / 20.0;
    const Vector<3> t = makeVector(mu[0], mu[1], mu[2]), w = makeVector(mu[3], mu[4], mu[5]);
    const double tsq = w * w, th = std::sqrt(tsq);
    double A, B;
    SE3 r;
    const Vector<3> cr = w ^ t;
    if (tsq < 1e-8) { A = 1.0 - one_6th * tsq; B = 0.5; r.trans = t + cr * 0.5; }
    else {
      double Cc;
      if (tsq < 1e-6) { Cc = one_6th * (1.0 - one_20th * tsq); A = 1.0 - tsq * Cc; B = 0.5 - 0.25 * one_6th * tsq; }
      else { const double it = 1.0 / th; A = std::sin(th) * it; B = (1 - std::cos(th)) * (it * it); Cc = (1 - A) * (it * it); }
      r.trans = t + cr * B + (w ^ cr) * Cc;
    }
    SO3::rodrigues(w, A, B, r.rot.R);
    return r;
  }
  // 12 doubles: row-major rotation then translation (the layout the C ABI takes)
  void pack(double* o) const { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o[i * 3 + j] = rot.R[i][j]; for (int i = 0; i < 3; i++) o[9 + i] = trans[i]; }
  static SE3 unpack(const double* o) { SE3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.rot.R[i][j] = o[i * 3 + j]; for (int i = 0; i < 3; i++) r.trans[i] = o[9 + i]; return r; }
};

// CVD::ImageRef / BasicImage<byte> subset
struct ImageRef {
  int x = 0, y = 0;
  ImageRef() {}
  ImageRef(int x_, int y_) : x(x_), y(y_) {}
  bool operator==(const ImageRef& o) const { return x == o.x && y == o.y; }
};
typedef unsigned char byte;
template <class T> struct BasicImage {
  T* my_data = nullptr;
  ImageRef my_size;
  int my_stride = 0;
  BasicImage() {}
  BasicImage(T* d, ImageRef s, int stride) : my_data(d), my_size(s), my_stride(stride) {}
  ImageRef size() const { return my_size; }
  int row_stride() const { return my_stride; }
  T* data() { return my_data; }
  const T* data() const { return my_data; }
  T& operator[](const ImageRef& p) { return my_data[(size_t)p.y * my_stride + p.x]; }
  const T& operator[](const ImageRef& p) const { return my_data[(size_t)p.y * my_stride + p.x]; }
  int totalsize() const { return my_size.x * my_size.y; }
};
template <class T> struct Image : BasicImage<T> {
  std::vector<T> store;
  Image() {}
  explicit Image(ImageRef s) { resize(s); }
  Image(const Image& o) : BasicImage<T>() { *this = o; }          // deep copy (CVD::Image shares; nothing here relies on sharing)
  Image& operator=(const Image& o)
  {
    if (this != &o) { store = o.store; this->my_size = o.my_size; this->my_stride = o.my_stride; this->my_data = store.empty() ? nullptr : store.data(); }
    return *this;
  }
  void resize(ImageRef s) { store.assign((size_t)s.x * s.y, T()); this->my_data = store.data(); this->my_size = s; this->my_stride = s.x; }
};

}  // namespace mcp_shim
