// Minimal stand-in for <opencv2/core/core.hpp>: just enough of cv::Mat / Mat_<T> / Rect / Size / RNG for the reference's
// own src/jda/cascador.cpp and src/jda/cart.cpp (and the detect-path pieces of data.cpp / btcart.cpp / common.cpp) to
// compile unmodified into oracle/_ref_cpp/ (oracle/Makefile).  TEST INFRASTRUCTURE ONLY -- it exists so that the
// restatement oracle/jda_oracle_cpp.c can be pinned against the reference's real C++ detector in an image without
// OpenCV.  Semantics follow OpenCV where the detect path depends on them:
//   * Mat is a reference-counted header over row-major data; copies share data, clone() copies, operator()(Rect) is a
//     view with the parent's row step;
//   * Mat_<T>(std::vector<T>) is an n x 1 column;
//   * nothing here does arithmetic the detect path's RESULTS depend on, except Mat_ += (element-wise double adds),
//     norm / mean (training and the similarity transform only);
//   * cv::resize is NOT OpenCV's (oracle/cvshim/opencv2/imgproc/imgproc.hpp): models whose nodes sample the half /
//     quarter images, and fddb.method = 0, are therefore outside what this build can pin.
#ifndef JDA_CVSHIM_CORE_HPP_
#define JDA_CVSHIM_CORE_HPP_
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <cmath>
#include <memory>
#include <string>
#include <vector>

namespace cv {

typedef unsigned char uchar;
typedef int64_t int64;
typedef uint64_t uint64;

enum { CV_8UC1 = 0, CV_32SC1 = 4, CV_64FC1 = 6 };
inline size_t cvshim_elem_size(int type) { return type == CV_8UC1 ? 1 : type == CV_32SC1 ? 4 : 8; }

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};

template <typename T>
struct Rect_ {
  T x, y, width, height;
  Rect_() : x(0), y(0), width(0), height(0) {}
  Rect_(T x_, T y_, T w_, T h_) : x(x_), y(y_), width(w_), height(h_) {}
};
typedef Rect_<int> Rect;

struct Scalar {
  double val[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
  double operator[](int i) const { return val[i]; }
};

class Mat {
 public:
  int rows, cols;
  size_t step;  // bytes between rows
  uchar *data;
  int mtype;
  std::shared_ptr<std::vector<uchar> > buf;  // owner (empty for user data)

  Mat() : rows(0), cols(0), step(0), data(NULL), mtype(CV_8UC1) {}
  Mat(int r, int c, int type) : rows(0), cols(0), step(0), data(NULL), mtype(type) { create(r, c, type); }
  Mat(int r, int c, int type, void *user, size_t step_ = 0)
      : rows(r), cols(c), step(step_ ? step_ : c * cvshim_elem_size(type)), data((uchar *)user), mtype(type) {}

  void create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == mtype && buf) return;  // OpenCV keeps a matching allocation
    mtype = type; rows = r; cols = c; step = c * cvshim_elem_size(type);
    buf.reset(new std::vector<uchar>((size_t)r * step + 8));
    data = buf->data();
  }
  void release() { buf.reset(); data = NULL; rows = cols = 0; step = 0; }
  bool empty() const { return data == NULL || rows == 0 || cols == 0; }
  size_t elemSize() const { return cvshim_elem_size(mtype); }
  Mat clone() const {
    Mat m(rows, cols, mtype);
    for (int i = 0; i < rows; i++) memcpy(m.data + i * m.step, data + i * step, cols * elemSize());
    return m;
  }
  void copyTo(Mat &m) const { m = clone(); }
  Mat operator()(const Rect &r) const {
    Mat m(*this);
    m.data = data + (size_t)r.y * step + (size_t)r.x * elemSize();
    m.rows = r.height; m.cols = r.width;
    return m;
  }
  Mat row(int i) const { return (*this)(Rect(0, i, cols, 1)); }
  template <typename T> T &at(int y, int x) { return *(T *)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  template <typename T> const T &at(int y, int x) const { return *(const T *)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  template <typename T> T *ptr(int y = 0) { return (T *)(data + (size_t)y * step); }
  template <typename T> const T *ptr(int y = 0) const { return (const T *)(data + (size_t)y * step); }
  Size size() const { return Size(cols, rows); }
};

template <typename T> struct cvshim_type;
template <> struct cvshim_type<uchar> { enum { value = CV_8UC1 }; };
template <> struct cvshim_type<int> { enum { value = CV_32SC1 }; };
template <> struct cvshim_type<double> { enum { value = CV_64FC1 }; };

template <typename T>
class Mat_ : public Mat {
 public:
  Mat_() { mtype = cvshim_type<T>::value; }
  Mat_(int r, int c) : Mat(r, c, cvshim_type<T>::value) {}
  Mat_(const Mat &m) : Mat(m) {}
  explicit Mat_(const std::vector<T> &v) : Mat((int)v.size(), 1, cvshim_type<T>::value) {
    for (size_t i = 0; i < v.size(); i++) (*this)((int)i, 0) = v[i];
  }
  void create(int r, int c) { Mat::create(r, c, cvshim_type<T>::value); }
  static Mat_ zeros(int r, int c) {
    Mat_ m(r, c);
    for (int i = 0; i < r; i++) memset(m.data + i * m.step, 0, c * sizeof(T));
    return m;
  }
  T &operator()(int y, int x) { return *(T *)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  const T &operator()(int y, int x) const { return *(const T *)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  Mat_ clone() const { return Mat_(Mat::clone()); }
  Mat_ row(int i) const { return Mat_(Mat::row(i)); }
  Mat_ mul(const Mat_ &o) const {
    Mat_ m(rows, cols);
    for (int i = 0; i < rows; i++)
      for (int j = 0; j < cols; j++) m(i, j) = (*this)(i, j) * o(i, j);
    return m;
  }
  Mat_ &operator+=(const Mat_ &o) {
    for (int i = 0; i < rows; i++)
      for (int j = 0; j < cols; j++) (*this)(i, j) += o(i, j);
    return *this;
  }
  Mat_ &operator/=(double d) {
    for (int i = 0; i < rows; i++)
      for (int j = 0; j < cols; j++) (*this)(i, j) /= d;
    return *this;
  }
};

inline double norm(const Mat_<double> &m) {
  double s = 0;
  for (int i = 0; i < m.rows; i++)
    for (int j = 0; j < m.cols; j++) s += m(i, j) * m(i, j);
  return std::sqrt(s);
}
inline Scalar mean(const Mat_<double> &m) {
  double s = 0;
  for (int i = 0; i < m.rows; i++)
    for (int j = 0; j < m.cols; j++) s += m(i, j);
  return Scalar(m.rows * m.cols ? s / (m.rows * m.cols) : 0.);
}

// DataSet::RandomShape seeds an RNG with getTickCount() for every window (data.cpp:227): the reference's initial shift
// is different on every run.  A test can fix the tick (and with it the seed and the shift) to pin that path too.
inline int64 &shim_fixed_tick() { static int64 t = 0; return t; }
inline int64 getTickCount() {
  if (shim_fixed_tick()) return shim_fixed_tick();
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (int64)ts.tv_sec * 1000000000LL + ts.tv_nsec;
}
inline double getTickFrequency() { return 1e9; }

// OpenCV's multiply-with-carry generator (public algorithm); the detect path only draws from it with
// shift_size = 0 (src/test.cpp:17,75), where every draw is multiplied by zero
class RNG {
 public:
  uint64 state;
  RNG() : state(0xffffffff) {}
  RNG(uint64 s) : state(s ? s : 0xffffffff) {}
  unsigned next() {
    state = (uint64)(unsigned)state * 4164903690U + (unsigned)(state >> 32);
    return (unsigned)state;
  }
  int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
  double uniform(double a, double b) {
    unsigned t = next();
    double r = ((uint64)t << 32 | next()) * 5.4210108624275221700372640043497e-20;
    return r * (b - a) + a;
  }
};

}  // namespace cv
#endif
