// [3P] CVD::ImageRef subset
#pragma once
namespace CVD {
struct ImageRef {
  int x, y;
  ImageRef() : x(0), y(0) {}
  ImageRef(int xx, int yy) : x(xx), y(yy) {}
  bool next(const ImageRef& max) { x++; if (x >= max.x) { x = 0; y++; if (y >= max.y) { y = 0; return false; } } return true; }
  ImageRef operator+(const ImageRef& o) const { return ImageRef(x + o.x, y + o.y); }
  ImageRef operator-(const ImageRef& o) const { return ImageRef(x - o.x, y - o.y); }
  ImageRef operator/(int k) const { return ImageRef(x / k, y / k); }
  ImageRef operator*(int k) const { return ImageRef(x * k, y * k); }
  bool operator==(const ImageRef& o) const { return x == o.x && y == o.y; }
  bool operator!=(const ImageRef& o) const { return !(*this == o); }
  unsigned int mag_squared() const { return (unsigned int)(x * x + y * y); }
  int area() const { return x * y; }
  int& operator[](int i) { return i == 0 ? x : y; }
  int operator[](int i) const { return i == 0 ? x : y; }
};
}  // namespace CVD
