// [3P] CVD::sample / CVD::transform restated from libCVD release 20121025 (vision.h): bilinear sample in double with the
// implicit double -> byte truncation; transform walks the output raster with running sums p += across / carriage return.
#pragma once
#include <TooN/TooN.h>
#include <cvd/image.h>
#include <cvd/utility.h>
namespace CVD {
template <class T, class S> inline void sample(const BasicImage<S>& im, double x, double y, T& result)
{
  const int lx = (int)x, ly = (int)y;
  x -= lx; y -= ly;
  result = (T)((1 - y) * ((1 - x) * im[ly][lx] + x * im[ly][lx + 1]) + y * ((1 - x) * im[ly + 1][lx] + x * im[ly + 1][lx + 1]));
}
template <class T, class S>
int transform(const BasicImage<S>& in, BasicImage<T>& out, const TooN::Matrix<2>& M, const TooN::Vector<2>& inOrig, const TooN::Vector<2>& outOrig,
              const T defaultValue = T())
{
  const int w = out.size().x, h = out.size().y, iw = in.size().x, ih = in.size().y;
  const TooN::Vector<2> across = M.T()[0];
  const TooN::Vector<2> down = M.T()[1];
  const TooN::Vector<2> p0 = inOrig - M * outOrig;
  double min_x = p0[0], min_y = p0[1];
  double max_x = min_x, max_y = min_y;
  if (across[0] < 0) min_x += w * across[0]; else max_x += w * across[0];
  if (down[0] < 0) min_x += h * down[0]; else max_x += h * down[0];
  if (across[1] < 0) min_y += w * across[1]; else max_y += w * across[1];
  if (down[1] < 0) min_y += h * down[1]; else max_y += h * down[1];
  const TooN::Vector<2> carriage_return = down - w * across;
  if (min_x >= 0 && min_y >= 0 && max_x < iw - 1 && max_y < ih - 1) {
    TooN::Vector<2> p = p0;
    for (int i = 0; i < h; ++i, p += carriage_return)
      for (int j = 0; j < w; ++j, p += across) sample(in, p[0], p[1], out[i][j]);
    return 0;
  }
  const double x_bound = iw - 1, y_bound = ih - 1;
  int count = 0;
  TooN::Vector<2> p = p0;
  for (int i = 0; i < h; ++i, p += carriage_return)
    for (int j = 0; j < w; ++j, p += across) {
      if (0 <= p[0] && 0 <= p[1] && p[0] < x_bound && p[1] < y_bound) sample(in, p[0], p[1], out[i][j]);
      else { out[i][j] = defaultValue; count++; }
    }
  return count;
}
}  // namespace CVD
