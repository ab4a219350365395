// [3P] TooN::Cholesky<N>: L D L^T as in TooN 2.x Cholesky.h (do_compute keeps the undivided values in the upper half;
// get_inverse() = backsub(Identity)).
#pragma once
#include <TooN/TooN.h>
namespace TooN {
template <int N> class Cholesky {
public:
  Cholesky(const Matrix<N, N>& m) : c(m) { compute(); }
  Vector<N> backsub(const Vector<N>& v) const
  {
    Vector<N> y, x;
    for (int i = 0; i < N; i++) { double val = v[i]; for (int j = 0; j < i; j++) val -= c(i, j) * y[j]; y[i] = val; }
    for (int i = 0; i < N; i++) y[i] /= c(i, i);
    for (int i = N - 1; i >= 0; i--) { double val = y[i]; for (int j = i + 1; j < N; j++) val -= c(j, i) * x[j]; x[i] = val; }
    return x;
  }
  Matrix<N, N> get_inverse() const
  {
    Matrix<N, N> inv;
    for (int col = 0; col < N; col++) {
      Vector<N> e; e[col] = 1.0;
      const Vector<N> x = backsub(e);
      for (int r = 0; r < N; r++) inv(r, col) = x[r];
    }
    return inv;
  }
private:
  void compute()
  {
    for (int col = 0; col < N; col++) {
      double inv_diag = 1;
      for (int row = col; row < N; row++) {
        double val = c(row, col);
        for (int col2 = 0; col2 < col; col2++) val -= c(col2, col) * c(row, col2);
        if (row == col) { c(row, col) = val; if (val == 0) return; inv_diag = 1 / val; }
        else { c(col, row) = val; c(row, col) = val * inv_diag; }
      }
    }
  }
  Matrix<N, N> c;
};
}  // namespace TooN
