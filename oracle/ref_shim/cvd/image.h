// [3P] CVD::BasicImage / Image subset (deep-copying value type; the translation units compiled here never rely on
// libCVD's reference counting)
#pragma once
#include <cvd/image_ref.h>
#include <cassert>
#include <cstddef>
#include <cstring>
#include <cmath>
#include <vector>
namespace CVD {
template <class T> class BasicImage {
public:
  BasicImage() : my_data(0), my_stride(0) {}
  BasicImage(T* d, const ImageRef& s, int stride) : my_data(d), my_size(s), my_stride(stride) {}
  virtual ~BasicImage() {}
  ImageRef size() const { return my_size; }
  int row_stride() const { return my_stride; }
  int totalsize() const { return my_size.x * my_size.y; }
  T* data() { return my_data; }
  const T* data() const { return my_data; }
  T& operator[](const ImageRef& p) { return my_data[(size_t)p.y * my_stride + p.x]; }
  const T& operator[](const ImageRef& p) const { return my_data[(size_t)p.y * my_stride + p.x]; }
  T* operator[](int row) { return my_data + (size_t)row * my_stride; }
  const T* operator[](int row) const { return my_data + (size_t)row * my_stride; }
  bool in_image(const ImageRef& ir) const { return ir.x >= 0 && ir.y >= 0 && ir.x < my_size.x && ir.y < my_size.y; }
  bool in_image_with_border(const ImageRef& ir, int border) const
  { return ir.x >= border && ir.y >= border && ir.x < my_size.x - border && ir.y < my_size.y - border; }
protected:
  T* my_data;
  ImageRef my_size;
  int my_stride;
};
template <class T> class SubImage : public BasicImage<T> {
public:
  SubImage() {}
  SubImage(T* d, const ImageRef& s, int stride) : BasicImage<T>(d, s, stride) {}
};
template <class T> class Image : public BasicImage<T> {
public:
  Image() {}
  explicit Image(const ImageRef& s) { resize(s); }
  Image(const Image& o) : BasicImage<T>() { *this = o; }
  Image& operator=(const Image& o) { store = o.store; this->my_size = o.my_size; this->my_stride = o.my_stride; this->my_data = store.empty() ? 0 : &store[0]; return *this; }
  void resize(const ImageRef& s) { store.assign((size_t)s.x * s.y, T()); this->my_size = s; this->my_stride = s.x; this->my_data = store.empty() ? 0 : &store[0]; }
private:
  std::vector<T> store;
};
}  // namespace CVD
