// Minimal stand-in for TooN 2.x (TEST INFRASTRUCTURE: lets oracle/_ref compile the reference's own translation units
// from /root/reference unmodified).  Only what src/TaylorCamera.cc, src/PatchFinder.cc, LevelHelpers.h and
// SmallMatrixOpts.h use.  Row-major storage, value semantics, no expression templates.  [3P] TooN semantics restated.
#pragma once
#include <cassert>
#include <cmath>
#include <algorithm>
#include <type_traits>
#include <cstddef>
#include <istream>
#include <ostream>
#include <vector>

namespace TooN {

static const int Dynamic = -1;
static const int Resizable = -0x7fffffff;

struct ZerosT { int n; ZerosT operator()(int k) const { return ZerosT{ k }; } };
struct OnesT { int n; OnesT operator()(int k) const { return OnesT{ k }; } };
struct IdentityT { double s; };
static const ZerosT Zeros = { -1 };
static const OnesT Ones = { -1 };
static const IdentityT Identity = { 1.0 };
inline IdentityT operator*(double s, const IdentityT& i) { return IdentityT{ s * i.s }; }
inline IdentityT operator*(const IdentityT& i, double s) { return IdentityT{ s * i.s }; }

namespace detail {
template <int N> struct Store {
  double v[N];
  int size() const { return N; }
  void resize(int n) { (void)n; assert(n == N); }
};
template <> struct Store<Dynamic> {
  std::vector<double> v;
  int size() const { return (int)v.size(); }
  void resize(int n) { v.assign((size_t)n, 0.0); }
};
template <> struct Store<Resizable> : Store<Dynamic> {};
struct DefaultBase {};
}  // namespace detail

template <int Size, class Precision, class Base> struct Vector;
template <int N> struct DiagView { const double* p; int n; };
template <int N> struct ColView { const double* p; int n; };
template <int N> struct RowView { const double* p; int n; };

template <int Size = Dynamic, class Precision = double, class Base = detail::DefaultBase> struct Vector {
  detail::Store<Size> s;
  Vector() { for (int i = 0; i < s.size(); i++) s.v[i] = 0.0; }
  explicit Vector(int n) { s.resize(n); }
  Vector(const ZerosT&) { for (int i = 0; i < s.size(); i++) s.v[i] = 0.0; }
  // static <-> dynamic conversions only (a static vector never converts to one of another size)
  template <int S2, class B2, class = typename std::enable_if<(S2 < 0) || (Size < 0)>::type> Vector(const Vector<S2, Precision, B2>& o) { s.resize(o.size()); for (int i = 0; i < size(); i++) s.v[i] = o[i]; }
  Vector(const OnesT& o) { if (o.n >= 0) s.resize(o.n); for (int i = 0; i < size(); i++) s.v[i] = 1.0; }
  template <int S2, class B2> Vector& operator=(const Vector<S2, Precision, B2>& o) { s.resize(o.size()); for (int i = 0; i < size(); i++) s.v[i] = o[i]; return *this; }
  Vector& operator=(const ZerosT&) { for (int i = 0; i < size(); i++) s.v[i] = 0.0; return *this; }
  Vector& operator=(const OnesT& o) { if (o.n >= 0 && o.n != size()) s.resize(o.n); for (int i = 0; i < size(); i++) s.v[i] = 1.0; return *this; }
  int size() const { return s.size(); }
  double& operator[](int i) { return s.v[i]; }
  const double& operator[](int i) const { return s.v[i]; }
  template <int Start, int Len> Vector<Len> slice() const { Vector<Len> r; for (int i = 0; i < Len; i++) r[i] = s.v[Start + i]; return r; }
  Vector<Dynamic> slice(int start, int len) const { Vector<Dynamic> r(len); for (int i = 0; i < len; i++) r[i] = s.v[start + i]; return r; }
  DiagView<Size> as_diagonal() const { return DiagView<Size>{ &s.v[0], size() }; }
  ColView<Size> as_col() const { return ColView<Size>{ &s.v[0], size() }; }
  RowView<Size> as_row() const { return RowView<Size>{ &s.v[0], size() }; }
  Vector& operator+=(const Vector& o) { for (int i = 0; i < size(); i++) s.v[i] += o[i]; return *this; }
  Vector& operator-=(const Vector& o) { for (int i = 0; i < size(); i++) s.v[i] -= o[i]; return *this; }
  Vector& operator*=(double k) { for (int i = 0; i < size(); i++) s.v[i] *= k; return *this; }
  Vector& operator/=(double k) { for (int i = 0; i < size(); i++) s.v[i] /= k; return *this; }
};

namespace detail {
template <int A, int B> struct Pick { static const int value = (A >= 0) ? A : B; };
}
template <int A, class BA, int B, class BB> Vector<detail::Pick<A, B>::value> operator+(const Vector<A, double, BA>& a, const Vector<B, double, BB>& b)
{ Vector<detail::Pick<A, B>::value> r; r.s.resize(a.size()); for (int i = 0; i < a.size(); i++) r[i] = a[i] + b[i]; return r; }
template <int A, class BA, int B, class BB> Vector<detail::Pick<A, B>::value> operator-(const Vector<A, double, BA>& a, const Vector<B, double, BB>& b)
{ Vector<detail::Pick<A, B>::value> r; r.s.resize(a.size()); for (int i = 0; i < a.size(); i++) r[i] = a[i] - b[i]; return r; }
template <int A, class BA, int B, class BB> double operator*(const Vector<A, double, BA>& a, const Vector<B, double, BB>& b)
{ double t = 0; for (int i = 0; i < a.size(); i++) t += a[i] * b[i]; return t; }
template <int A, class BA> Vector<A> operator*(const Vector<A, double, BA>& a, double k) { Vector<A> r; r.s.resize(a.size()); for (int i = 0; i < a.size(); i++) r[i] = a[i] * k; return r; }
template <int A, class BA> Vector<A> operator*(double k, const Vector<A, double, BA>& a) { return a * k; }
template <int A, class BA> Vector<A> operator/(const Vector<A, double, BA>& a, double k) { Vector<A> r; r.s.resize(a.size()); for (int i = 0; i < a.size(); i++) r[i] = a[i] / k; return r; }
template <int A, class BA> Vector<A> operator-(const Vector<A, double, BA>& a) { return a * -1.0; }
// Ones(n) * scalar (TaylorCamera::CenterAndScale)
inline Vector<Dynamic> operator*(const OnesT& o, double k) { Vector<Dynamic> r(o.n); for (int i = 0; i < o.n; i++) r[i] = k; return r; }
// row vector times diagonal matrix (TaylorCamera::PolyFit)
template <int A, class BA, int B> Vector<A> operator*(const Vector<A, double, BA>& a, const DiagView<B>& d) { Vector<A> r; r.s.resize(a.size()); for (int i = 0; i < a.size(); i++) r[i] = a[i] * d.p[i]; return r; }

template <int A, class BA> std::ostream& operator<<(std::ostream& os, const Vector<A, double, BA>& v) { for (int i = 0; i < v.size(); i++) os << v[i] << " "; return os; }
template <int A, class BA> std::istream& operator>>(std::istream& is, Vector<A, double, BA>& v) { for (int i = 0; i < v.size(); i++) is >> v[i]; return is; }
inline Vector<2> makeVector(double a, double b) { Vector<2> r; r[0] = a; r[1] = b; return r; }
inline Vector<3> makeVector(double a, double b, double c) { Vector<3> r; r[0] = a; r[1] = b; r[2] = c; return r; }
inline Vector<4> makeVector(double a, double b, double c, double d) { Vector<4> r; r[0] = a; r[1] = b; r[2] = c; r[3] = d; return r; }
inline Vector<6> makeVector(double a, double b, double c, double d, double e, double f) { Vector<6> r; r[0] = a; r[1] = b; r[2] = c; r[3] = d; r[4] = e; r[5] = f; return r; }
inline Vector<9> makeVector(double a, double b, double c, double d, double e, double f, double g, double h, double i)
{ Vector<9> r; r[0] = a; r[1] = b; r[2] = c; r[3] = d; r[4] = e; r[5] = f; r[6] = g; r[7] = h; r[8] = i; return r; }
inline Vector<3> operator^(const Vector<3>& a, const Vector<3>& b) { return makeVector(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]); }
template <int A, class BA> double norm(const Vector<A, double, BA>& a) { return std::sqrt(a * a); }
template <int A, class BA> double norm_sq(const Vector<A, double, BA>& a) { return a * a; }
template <int A, class BA> void normalize(Vector<A, double, BA>& a) { const double n = std::sqrt(a * a); for (int i = 0; i < a.size(); i++) a[i] /= n; }
template <int A, class BA> Vector<A> unit(const Vector<A, double, BA>& a) { Vector<A> r = a; normalize(r); return r; }
template <int N> Vector<N + 1> unproject(const Vector<N>& v) { Vector<N + 1> r; for (int i = 0; i < N; i++) r[i] = v[i]; r[N] = 1.0; return r; }
inline Vector<2> project(const Vector<3>& v) { return makeVector(v[0] / v[2], v[1] / v[2]); }

// ---------------------------------------------------------------------------------------------------------------
template <int R, int C, class Precision> struct Matrix;
template <int R, int C> struct MatStore {
  double m[R * C];
  int rows() const { return R; }
  int cols() const { return C; }
  void resize(int r, int c) { (void)r; (void)c; assert(r == R && c == C); }
};
template <> struct MatStore<Dynamic, Dynamic> {
  std::vector<double> m; int nr = 0, nc = 0;
  int rows() const { return nr; }
  int cols() const { return nc; }
  void resize(int r, int c) { nr = r; nc = c; m.assign((size_t)r * c, 0.0); }
};
// a row (or, through T(), a column) of a matrix that can be read, indexed and assigned from a vector
struct StridedRef {
  double* p; int n, stride;
  double& operator[](int i) { return p[i * stride]; }
  const double& operator[](int i) const { return p[i * stride]; }
  int size() const { return n; }
  template <int S, class B> StridedRef& operator=(const Vector<S, double, B>& v) { for (int i = 0; i < n; i++) p[i * stride] = v[i]; return *this; }
  StridedRef& operator=(const OnesT&) { for (int i = 0; i < n; i++) p[i * stride] = 1.0; return *this; }
  StridedRef& operator=(const ZerosT&) { for (int i = 0; i < n; i++) p[i * stride] = 0.0; return *this; }
  StridedRef& operator=(const StridedRef& o) { for (int i = 0; i < n; i++) p[i * stride] = o[i]; return *this; }
  operator Vector<Dynamic>() const { Vector<Dynamic> r(n); for (int i = 0; i < n; i++) r[i] = p[i * stride]; return r; }
  template <int S> operator Vector<S>() const { Vector<S> r; for (int i = 0; i < n; i++) r[i] = p[i * stride]; return r; }
};
template <int S, class B> Vector<Dynamic> operator-(const StridedRef& a, const Vector<S, double, B>& b) { Vector<Dynamic> r(a.n); for (int i = 0; i < a.n; i++) r[i] = a[i] - b[i]; return r; }
inline Vector<Dynamic> operator-(const StridedRef& a, const StridedRef& b) { Vector<Dynamic> r(a.n); for (int i = 0; i < a.n; i++) r[i] = a[i] - b[i]; return r; }
template <int B> Vector<Dynamic> operator*(const StridedRef& a, const DiagView<B>& d) { Vector<Dynamic> r(a.n); for (int i = 0; i < a.n; i++) r[i] = a[i] * d.p[i]; return r; }
struct TransposeRef {
  double* p; int rows, cols;                       // of the ORIGINAL matrix (row-major, stride = cols)
  StridedRef operator[](int c) const { return StridedRef{ p + c, rows, cols }; }
};

template <int R = Dynamic, int C = R, class Precision = double> struct Matrix {
  MatStore<R, C> s;
  Matrix() { for (int i = 0; i < s.rows() * s.cols(); i++) s.m[i] = 0.0; }
  Matrix(int r, int c) { s.resize(r, c); }
  Matrix(const ZerosT&) { for (int i = 0; i < s.rows() * s.cols(); i++) s.m[i] = 0.0; }
  Matrix(const IdentityT& id) { *this = id; }
  Matrix& operator=(const ZerosT&) { for (int i = 0; i < s.rows() * s.cols(); i++) s.m[i] = 0.0; return *this; }
  Matrix& operator=(const IdentityT& id) { for (int i = 0; i < num_rows(); i++) for (int j = 0; j < num_cols(); j++) s.m[i * num_cols() + j] = (i == j) ? id.s : 0.0; return *this; }
  int num_rows() const { return s.rows(); }
  int num_cols() const { return s.cols(); }
  StridedRef operator[](int r) { return StridedRef{ &s.m[0] + (size_t)r * num_cols(), num_cols(), 1 }; }
  const StridedRef operator[](int r) const { return StridedRef{ const_cast<double*>(&s.m[0]) + (size_t)r * num_cols(), num_cols(), 1 }; }
  double& operator()(int r, int c) { return s.m[(size_t)r * num_cols() + c]; }
  const double& operator()(int r, int c) const { return s.m[(size_t)r * num_cols() + c]; }
  TransposeRef T() { return TransposeRef{ &s.m[0], num_rows(), num_cols() }; }
  const TransposeRef T() const { return TransposeRef{ const_cast<double*>(&s.m[0]), num_rows(), num_cols() }; }
  Matrix& operator+=(const Matrix& o) { for (int i = 0; i < num_rows() * num_cols(); i++) s.m[i] += o.s.m[i]; return *this; }
};
template <int R, int C> Vector<R> operator*(const Matrix<R, C>& A, const Vector<C>& x)
{ Vector<R> r; for (int i = 0; i < R; i++) { double t = 0; for (int j = 0; j < C; j++) t += A(i, j) * x[j]; r[i] = t; } return r; }
template <int R, int K, int C> Matrix<R, C> operator*(const Matrix<R, K>& A, const Matrix<K, C>& B)
{ Matrix<R, C> r; for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) { double t = 0; for (int k = 0; k < K; k++) t += A(i, k) * B(k, j); r(i, j) = t; } return r; }
template <int R, int C> Matrix<R, C> operator*(const Matrix<R, C>& A, double k) { Matrix<R, C> r; for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) r(i, j) = A(i, j) * k; return r; }
template <int R, int C> Matrix<R, C> operator*(double k, const Matrix<R, C>& A) { return A * k; }
template <int N> Matrix<N, N> operator*(const ColView<N>& a, const RowView<N>& b)
{ Matrix<N, N> r; for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) r(i, j) = a.p[i] * b.p[j]; return r; }

}  // namespace TooN
