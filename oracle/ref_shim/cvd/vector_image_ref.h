// [3P] CVD vec / ir / ir_rounded
#pragma once
#include <TooN/TooN.h>
#include <cvd/image_ref.h>
namespace CVD {
inline TooN::Vector<2> vec(const ImageRef& ir) { return TooN::makeVector(ir.x, ir.y); }
inline ImageRef ir(const TooN::Vector<2>& v) { return ImageRef((int)v[0], (int)v[1]); }
inline ImageRef ir_rounded(const TooN::Vector<2>& v)
{ return ImageRef((int)(v[0] > 0.0 ? v[0] + 0.5 : v[0] - 0.5), (int)(v[1] > 0.0 ? v[1] + 0.5 : v[1] - 0.5)); }
}  // namespace CVD
