// [3P] CVD::copy
#pragma once
#include <cvd/image.h>
namespace CVD {
template <class S, class T> void copy(const BasicImage<S>& in, BasicImage<T>& out, ImageRef size = ImageRef(-1, -1), ImageRef begin = ImageRef(), ImageRef dst = ImageRef())
{
  if (size.x == -1 && size.y == -1) size = in.size();
  for (int y = 0; y < size.y; y++)
    for (int x = 0; x < size.x; x++) out[ImageRef(dst.x + x, dst.y + y)] = (T)in[ImageRef(begin.x + x, begin.y + y)];
}
}  // namespace CVD
