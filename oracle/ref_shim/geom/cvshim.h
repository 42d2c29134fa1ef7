/* cvshim.h -- TEST INFRASTRUCTURE ONLY (oracle/README_ref.md).
 *
 * The few cv:: names the reference's host-side GEOMETRY code touches, so that it compiles here
 * unmodified and can pin the product's restatement:
 *   Thirdparty/Localization/image.cpp            whole file  (warp matrices :76-108, :236-268)
 *   Thirdparty/Localization/ioData.cpp:177-197   setupCam
 *   Thirdparty/Localization/rendering.hpp:642-694 calculateTranslation / calculateTranslationCV
 *   src/Tracking.cc:2374-2419                    Tracking::CalculateNMIRelocalization
 * OpenCV 3.4.0 itself is not in the reference tree (build/ORB_SLAM2.vcxproj:41 links a prebuilt
 * opencv-3.4.0_CUDA).  What is restated here is its PUBLISHED behaviour for the operations used:
 *   Mat * Mat   cv::gemm's fixed-size path for inner dimension <= 4: every element is the products
 *               summed left to right in the element type (no FMA contraction: compile this shim
 *               with -ffp-contract=off);
 *   Mat::inv()  cv::invert(DECOMP_LU) 3x3 special case: cofactors over det3, d = 1/det, each
 *               cofactor times d -- in double for CV_64F and for CV_32F alike (result cast back);
 *   eye, clone, copyTo, at<T>, ROI by Rect, Mat_<T>(r, c) << a, b, ... (row-major fill).
 * The same two arithmetic rules are cross-checked against the real OpenCV through Python cv2 in
 * tests/test_cv2_crosschecks.py (shim-compiled reference == cv2-driven restatement, bit for bit).
 * cv::cuda::GpuMat / createContinuous / warpPerspective are inert: nothing here touches a GPU. */
#ifndef NMI_CVSHIM_H_
#define NMI_CVSHIM_H_
#include <cmath>
#include <cstddef>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_64F 6

namespace cv {

struct Size {
  int width = 0, height = 0;
  Size() {}
  Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
  int x, y, width, height;
  Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

class Mat {
 public:
  int rows = 0, cols = 0;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  static Mat eye(int r, int c, int type) {
    Mat m(r, c, type);
    std::memset(m.buf_->data(), 0, m.buf_->size());
    for (int i = 0; i < (r < c ? r : c); i++) m.set(i, i, 1.0);
    return m;
  }
  int type() const { return type_; }
  bool empty() const { return rows == 0 || cols == 0; }
  Size size() const { return Size(cols, rows); }
  template <typename T> T& at(int i, int j) { return *reinterpret_cast<T*>(ptr(i, j)); }
  template <typename T> const T& at(int i, int j) const { return *reinterpret_cast<const T*>(ptr(i, j)); }
  template <typename T> T& at(int i) { return cols == 1 ? at<T>(i, 0) : at<T>(i / cols, i % cols); }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int i = 0; i < rows; i++) std::memcpy(m.ptr(i, 0), ptr(i, 0), (size_t)cols * esz());
    return m;
  }
  void copyTo(Mat& dst) const {  // cv::Mat::copyTo: (re)allocates unless size and type already match
    if (dst.rows != rows || dst.cols != cols || dst.type_ != type_) dst.create(rows, cols, type_);
    for (int i = 0; i < rows; i++) std::memcpy(dst.ptr(i, 0), ptr(i, 0), (size_t)cols * esz());
  }
  void copyTo(Mat&& roi) const { copyTo(roi); }  // a ROI header: same size, copied in place
  Mat operator()(const Rect& r) const {          // ROI sharing the data
    Mat m;
    m.rows = r.height; m.cols = r.width; m.type_ = type_; m.buf_ = buf_; m.stride_ = stride_;
    m.off_ = off_ + (size_t)r.y * stride_ + (size_t)r.x * esz();
    return m;
  }
  Mat inv() const {  // cv::invert, DECOMP_LU, the n == 3 special case (double arithmetic for both depths)
    Mat r(rows, cols, type_);
    double s[3][3];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) s[i][j] = get(i, j);
    double d = s[0][0] * (s[1][1] * s[2][2] - s[1][2] * s[2][1]) - s[0][1] * (s[1][0] * s[2][2] - s[1][2] * s[2][0]) +
               s[0][2] * (s[1][0] * s[2][1] - s[1][1] * s[2][0]);
    if (d != 0.0) {
      d = 1.0 / d;
      double t[9];
      t[0] = (s[1][1] * s[2][2] - s[1][2] * s[2][1]) * d;
      t[1] = (s[0][2] * s[2][1] - s[0][1] * s[2][2]) * d;
      t[2] = (s[0][1] * s[1][2] - s[0][2] * s[1][1]) * d;
      t[3] = (s[1][2] * s[2][0] - s[1][0] * s[2][2]) * d;
      t[4] = (s[0][0] * s[2][2] - s[0][2] * s[2][0]) * d;
      t[5] = (s[0][2] * s[1][0] - s[0][0] * s[1][2]) * d;
      t[6] = (s[1][0] * s[2][1] - s[1][1] * s[2][0]) * d;
      t[7] = (s[0][1] * s[2][0] - s[0][0] * s[2][1]) * d;
      t[8] = (s[0][0] * s[1][1] - s[0][1] * s[1][0]) * d;
      for (int i = 0; i < 9; i++) r.set(i / 3, i % 3, t[i]);
    }
    return r;
  }
  double get(int i, int j) const { return type_ == CV_64F ? at<double>(i, j) : (type_ == CV_32F ? (double)at<float>(i, j) : (double)at<unsigned char>(i, j)); }
  void set(int i, int j, double v) {
    if (type_ == CV_64F) at<double>(i, j) = v;
    else if (type_ == CV_32F) at<float>(i, j) = (float)v;
    else at<unsigned char>(i, j) = (unsigned char)v;
  }
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type; off_ = 0; stride_ = (size_t)c * esz();
    buf_ = std::make_shared<std::vector<unsigned char>>((size_t)r * stride_);
  }
  size_t esz() const { return type_ == CV_64F ? 8 : (type_ == CV_32F ? 4 : 1); }
  unsigned char* data_ptr() const { return buf_ ? buf_->data() + off_ : nullptr; }

 protected:
  unsigned char* ptr(int i, int j) const { return buf_->data() + off_ + (size_t)i * stride_ + (size_t)j * esz(); }
  int type_ = CV_8U;
  size_t off_ = 0, stride_ = 0;
  std::shared_ptr<std::vector<unsigned char>> buf_;
};

// cv::gemm for small matrices: each element is the sum of products, left to right, in the depth
inline Mat operator*(const Mat& a, const Mat& b) {
  const int type = a.type();
  Mat d(a.rows, b.cols, type);
  for (int i = 0; i < a.rows; i++)
    for (int j = 0; j < b.cols; j++) {
      if (type == CV_64F) {
        double s = a.at<double>(i, 0) * b.at<double>(0, j);
        for (int k = 1; k < a.cols; k++) s = s + a.at<double>(i, k) * b.at<double>(k, j);
        d.at<double>(i, j) = s;
      } else {
        float s = a.at<float>(i, 0) * b.at<float>(0, j);
        for (int k = 1; k < a.cols; k++) s = s + a.at<float>(i, k) * b.at<float>(k, j);
        d.at<float>(i, j) = s;
      }
    }
  return d;
}

template <typename T> struct DepthOf;
template <> struct DepthOf<double> { enum { value = CV_64F }; };
template <> struct DepthOf<float> { enum { value = CV_32F }; };

template <typename T> class Mat_;
template <typename T>
struct MatCommaInit {  // (Mat_<T>(r, c) << a, b, c ...): row-major fill, every value converted to T
  Mat_<T>* m;
  int n;
  template <typename V> MatCommaInit& operator,(V v) {
    m->template at<T>(n / m->cols, n % m->cols) = (T)v;
    n++;
    return *this;
  }
  operator Mat_<T>() const { return *m; }
};

template <typename T>
class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int r, int c) : Mat(r, c, DepthOf<T>::value) {}
  Mat_(const Mat& m) : Mat(m) {}
  Mat_& operator=(const Mat& m) {
    Mat::operator=(m);
    return *this;
  }
  template <typename V> MatCommaInit<T> operator<<(V v) {
    MatCommaInit<T> ci{this, 0};
    ci, v;
    return ci;
  }
};

namespace cuda {
template <typename T> struct PtrStep { T* data; size_t step; };
struct GpuMat {
  unsigned char* data = nullptr;
  void release() {}
  void upload(const Mat&) {}
};
inline void createContinuous(int, int, int, GpuMat&) {}
inline void warpPerspective(const GpuMat&, GpuMat&, const Mat&, Size) {}
}  // namespace cuda
}  // namespace cv
#endif
