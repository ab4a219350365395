#pragma once
#include <TooN/so3.h>
namespace TooN {
template <class P = double> class SE3 {
public:
  SE3() {}
  SE3(const SO3<P>& r, const Vector<3>& t) : rot(r), trans(t) {}
  const SO3<P>& get_rotation() const { return rot; }
  SO3<P>& get_rotation() { return rot; }
  const Vector<3>& get_translation() const { return trans; }
  Vector<3>& get_translation() { return trans; }
  Vector<3> operator*(const Vector<3>& x) const { return rot * x + trans; }
  SE3 operator*(const SE3& o) const { return SE3(rot * o.rot, rot * o.trans + trans); }
  Vector<4> operator*(const Vector<4>& x) const { const Vector<3> r = rot * x.template slice<0, 3>() + trans * x[3]; return makeVector(r[0], r[1], r[2], x[3]); }
  static Vector<4> generator_field(int i, const Vector<4>& pos)
  { Vector<4> r; if (i < 3) { r[i] = pos[3]; return r; } r[(i + 1) % 3] = -pos[(i + 2) % 3]; r[(i + 2) % 3] = pos[(i + 1) % 3]; return r; }
  SE3 inverse() const { const SO3<P> ri = rot.inverse(); return SE3(ri, -(ri * trans)); }
  static SE3 exp(const Vector<6>& mu)
  {
    const double one_6th = 1.0 / 6.0, one_20th = 1.0 / 20.0;
    const Vector<3> t = mu.template slice<0, 3>(), w = mu.template slice<3, 3>();
    const double tsq = w * w, th = std::sqrt(tsq);
    double A, B;
    SE3 r;
    const Vector<3> cr = w ^ t;
    if (tsq < 1e-8) { A = 1.0 - one_6th * tsq; B = 0.5; r.trans = t + 0.5 * cr; }
    else {
      double Cc;
      if (tsq < 1e-6) { Cc = one_6th * (1.0 - one_20th * tsq); A = 1.0 - tsq * Cc; B = 0.5 - 0.25 * one_6th * tsq; }
      else { const double it = 1.0 / th; A = std::sin(th) * it; B = (1 - std::cos(th)) * (it * it); Cc = (1 - A) * (it * it); }
      r.trans = t + B * cr + Cc * (w ^ cr);
    }
    SO3<P>::rodrigues(w, A, B, r.rot.m);
    return r;
  }
private:
  SO3<P> rot;
  Vector<3> trans;
};
template <class P> std::ostream& operator<<(std::ostream& os, const SE3<P>& T)
{ for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) os << T.get_rotation().get_matrix()(i, j) << " "; os << T.get_translation()[i] << "\n"; } return os; }
template <class P> std::istream& operator>>(std::istream& is, SE3<P>& T)
{ Matrix<3> R; for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) is >> R(i, j); is >> T.get_translation()[i]; } T.get_rotation() = R; return is; }
}  // namespace TooN
