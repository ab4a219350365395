// [3P] TooN::SO3 / SE3 subset (left-multiplicative, exp with the translation first) -- see TooN/TooN.h
#pragma once
#include <TooN/TooN.h>
namespace TooN {
template <class P = double> class SO3 {
public:
  SO3() { m = Identity; }
  SO3(const Matrix<3>& r) { m = r; }
  SO3& operator=(const Matrix<3>& r) { m = r; return *this; }
  const Matrix<3>& get_matrix() const { return m; }
  SO3 inverse() const { SO3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m(i, j) = m(j, i); return r; }
  Vector<3> operator*(const Vector<3>& x) const { return m * x; }
  SO3 operator*(const SO3& o) const { SO3 r; r.m = m * o.m; return r; }
  static Vector<3> generator_field(int i, const Vector<3>& pos)
  { Vector<3> r; r[i] = 0; r[(i + 1) % 3] = -pos[(i + 2) % 3]; r[(i + 2) % 3] = pos[(i + 1) % 3]; return r; }
  static void rodrigues(const Vector<3>& w, double A, double B, Matrix<3>& R)
  {
    const double wx2 = w[0] * w[0], wy2 = w[1] * w[1], wz2 = w[2] * w[2];
    R(0, 0) = 1.0 - B * (wy2 + wz2); R(1, 1) = 1.0 - B * (wx2 + wz2); R(2, 2) = 1.0 - B * (wx2 + wy2);
    { const double a = A * w[2], b = B * (w[0] * w[1]); R(0, 1) = b - a; R(1, 0) = b + a; }
    { const double a = A * w[1], b = B * (w[0] * w[2]); R(0, 2) = b + a; R(2, 0) = b - a; }
    { const double a = A * w[0], b = B * (w[1] * w[2]); R(1, 2) = b - a; R(2, 1) = b + a; }
  }
  static SO3 exp(const Vector<3>& w)
  {
    const double one_6th = 1.0 / 6.0, one_20th = 1.0 / 20.0;
    const double tsq = w * w, th = std::sqrt(tsq);
    double A, B;
    if (tsq < 1e-8) { A = 1.0 - one_6th * tsq; B = 0.5; }
    else if (tsq < 1e-6) { B = 0.5 - 0.25 * one_6th * tsq; A = 1.0 - tsq * one_6th * (1.0 - one_20th * tsq); }
    else { const double it = 1.0 / th; A = std::sin(th) * it; B = (1 - std::cos(th)) * (it * it); }
    SO3 r;
    rodrigues(w, A, B, r.m);
    return r;
  }
  Matrix<3> m;
};
}  // namespace TooN
