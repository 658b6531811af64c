// Minimal stand-ins for the few OpenCV types the hot-path class interfaces mention, used only
// when <opencv2/opencv.hpp> is not installed (it is not in the build image).  With OpenCV present
// the real headers are included instead and these definitions vanish, so the classes keep the
// exact reference signatures (reference include/irmv_detection/yolo_engine.hpp:28-35,
// pnp_solver.hpp:15-23).
#pragma once
#if __has_include(<opencv2/opencv.hpp>)
#include <opencv2/opencv.hpp>
#define IRMV_HAVE_OPENCV 1
#else
#define IRMV_HAVE_OPENCV 0
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

// type codes are global macros in OpenCV
#define CV_8UC3 16
#define CV_64F 6

namespace cv
{

struct Size
{
  int width = 0, height = 0;
  Size() = default;
  Size(int w, int h) : width(w), height(h) {}
};

template <typename T>
struct Point_
{
  T x{}, y{};
  Point_() = default;
  Point_(T x_, T y_) : x(x_), y(y_) {}
  Point_ operator+(const Point_ & o) const { return {T(x + o.x), T(y + o.y)}; }
  Point_ operator-(const Point_ & o) const { return {T(x - o.x), T(y - o.y)}; }
  Point_ operator/(T d) const { return {T(x / d), T(y / d)}; }
};
using Point2f = Point_<float>;
using Point = Point_<int>;

struct Point3d
{
  double x = 0, y = 0, z = 0;
  Point3d() = default;
  Point3d(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
};

template <typename T>
inline T norm(const Point_<T> & p)
{
  return std::sqrt(p.x * p.x + p.y * p.y);
}

struct RotatedRect
{
  Point2f center;
  Size size;
  float angle = 0.f;
};

// Dense row-major matrix: owning (create) or a view over external memory.
class Mat
{
public:
  Mat() = default;
  Mat(Size s, int type, void * external) : rows(s.height), cols(s.width), type_(type), data(static_cast<uint8_t *>(external)) {}
  Mat(int r, int c, int type) { create(r, c, type); }
  void create(int r, int c, int type)
  {
    rows = r; cols = c; type_ = type;
    own_.assign(static_cast<size_t>(r) * c * elemSize(), 0);
    data = own_.data();
  }
  size_t elemSize() const { return type_ == CV_64F ? 8 : (type_ == CV_8UC3 ? 3 : 1); }
  size_t total() const { return static_cast<size_t>(rows) * cols; }
  template <typename T> T & at(int i) { return reinterpret_cast<T *>(data)[i]; }
  template <typename T> const T & at(int i) const { return reinterpret_cast<const T *>(data)[i]; }
  template <typename T> T & at(int r, int c) { return reinterpret_cast<T *>(data)[r * cols + c]; }
  template <typename T> const T & at(int r, int c) const { return reinterpret_cast<const T *>(data)[r * cols + c]; }
  bool empty() const { return data == nullptr; }
  int rows = 0, cols = 0;
  uint8_t * data = nullptr;

private:
  int type_ = 0;
  std::vector<uint8_t> own_;
};
}  // namespace cv
#endif
