// [3P] TooN::SVD<> stand-in: only backsub() (least-squares solution through the pseudo-inverse) is used, by
// TaylorCamera::PolyFit.  One-sided Jacobi SVD in long double; singular values below 1e20 * eps relative are dropped like
// TooN's default condition number.
#pragma once
#include <TooN/TooN.h>
#include <vector>
namespace TooN {
template <int R = Dynamic, int C = R> class SVD {
public:
  SVD(const TransposeRef& t)            // matrix = transpose of a row-major (t.rows x t.cols) array
  {
    m = t.cols; n = t.rows;             // m x n
    A.assign((size_t)m * n, 0.0L);
    for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) A[(size_t)i * n + j] = t.p[(size_t)j * t.cols + i];
    compute();
  }
  Vector<Dynamic> backsub(const Vector<Dynamic>& rhs, double condition = 1e20) const
  {
    long double smax = 0;
    for (int j = 0; j < n; j++) if (sv[j] > smax) smax = sv[j];
    Vector<Dynamic> x(n);
    std::vector<long double> acc((size_t)n, 0.0L);
    for (int j = 0; j < n; j++) {
      if (!(sv[j] * condition > smax) || sv[j] == 0) continue;
      long double ub = 0;
      for (int i = 0; i < m; i++) ub += A[(size_t)i * n + j] * rhs[i];      // U_j = A_j / sv_j
      ub /= sv[j] * sv[j];
      for (int k = 0; k < n; k++) acc[k] += V[(size_t)k * n + j] * ub;
    }
    for (int k = 0; k < n; k++) x[k] = (double)acc[k];
    return x;
  }
private:
  void compute()
  {
    V.assign((size_t)n * n, 0.0L);
    for (int i = 0; i < n; i++) V[(size_t)i * n + i] = 1;
    for (int sweep = 0; sweep < 60; sweep++) {
      long double off = 0;
      for (int p = 0; p < n - 1; p++)
        for (int q = p + 1; q < n; q++) {
          long double a = 0, b = 0, c = 0;
          for (int i = 0; i < m; i++) { const long double x = A[(size_t)i * n + p], y = A[(size_t)i * n + q]; a += x * x; b += y * y; c += x * y; }
          if (c == 0 || fabsl(c) <= 1e-19L * sqrtl(a * b)) continue;
          off += fabsl(c) / sqrtl(a * b);
          const long double zeta = (b - a) / (2 * c);
          const long double tt = (zeta >= 0 ? 1.0L : -1.0L) / (fabsl(zeta) + sqrtl(1 + zeta * zeta));
          const long double cs = 1 / sqrtl(1 + tt * tt), sn = cs * tt;
          for (int i = 0; i < m; i++) { const long double x = A[(size_t)i * n + p], y = A[(size_t)i * n + q]; A[(size_t)i * n + p] = cs * x - sn * y; A[(size_t)i * n + q] = sn * x + cs * y; }
          for (int i = 0; i < n; i++) { const long double x = V[(size_t)i * n + p], y = V[(size_t)i * n + q]; V[(size_t)i * n + p] = cs * x - sn * y; V[(size_t)i * n + q] = sn * x + cs * y; }
        }
      if (off < 1e-18L) break;
    }
    sv.assign((size_t)n, 0.0L);
    for (int j = 0; j < n; j++) { long double a = 0; for (int i = 0; i < m; i++) a += A[(size_t)i * n + j] * A[(size_t)i * n + j]; sv[j] = sqrtl(a); }
  }
  int m, n;
  std::vector<long double> A, V, sv;     // A: columns = U_j * sv_j after the sweeps
};
}  // namespace TooN
