// nmi_compat_types.hpp -- the handful of OpenCV / GLM types that appear in the
// signatures of the reference's Thirdparty/Localization + CUDA_Functions interfaces.
//
// When the integrator builds against real OpenCV / GLM (as the reference does,
// build/ORB_SLAM2.vcxproj:41), define NMI_COMPAT_HAVE_OPENCV / NMI_COMPAT_HAVE_GLM and
// the real headers are used.  This image has neither, so minimal
// stand-ins with the same names and member layout used on this path are provided:
//   cv::Mat (2-D, CV_8U / CV_32F / CV_64F, at<T>, clone, eye, inv, operator*),
//   cv::cuda::GpuMat {data, rows, cols, step}, cv::cuda::PtrStep<T>, glm::vec3.
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <vector>

#ifdef NMI_COMPAT_HAVE_OPENCV
#include <opencv2/core/core.hpp>
#include <opencv2/core/cuda.hpp>
#else
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_64F 6
namespace cv {
class Mat {
 public:
  int rows = 0, cols = 0;
  unsigned char* data = nullptr;
  Mat() = default;
  Mat(int r, int c, int type) { create(r, c, type); }
  void create(int r, int c, int type) {
    rows = r;
    cols = c;
    type_ = type;
    buf_ = std::make_shared<std::vector<unsigned char>>((size_t)r * c * elemSize(), 0);
    data = buf_->data();
  }
  int type() const { return type_; }
  size_t elemSize() const { return type_ == CV_64F ? 8 : type_ == CV_32F ? 4 : 1; }
  bool empty() const { return rows == 0 || cols == 0; }
  bool isContinuous() const { return true; }
  template <typename T>
  T& at(int r, int c) { return reinterpret_cast<T*>(data)[(size_t)r * cols + c]; }
  template <typename T>
  const T& at(int r, int c) const { return reinterpret_cast<const T*>(data)[(size_t)r * cols + c]; }
  template <typename T>
  T& at(int i) { return reinterpret_cast<T*>(data)[i]; }
  Mat clone() const {
    Mat m(rows, cols, type_);
    if (!empty()) std::memcpy(m.data, data, buf_->size());
    return m;
  }
  void copyTo(Mat& o) const { o = clone(); }
  static Mat eye(int r, int c, int type) {
    Mat m(r, c, type);
    for (int i = 0; i < (r < c ? r : c); i++) {
      if (type == CV_64F) m.at<double>(i, i) = 1.0;
      else if (type == CV_32F) m.at<float>(i, i) = 1.0f;
      else m.at<unsigned char>(i, i) = 1;
    }
    return m;
  }
  // rigid-transform inverse is all this path needs (Tcw <-> Twc, 4x4 CV_32F)
  Mat inv() const {
    if (rows != 4 || cols != 4 || type_ != CV_32F) throw std::runtime_error("cv::Mat shim: inv() is 4x4 CV_32F only");
    Mat o = eye(4, 4, CV_32F);
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) o.at<float>(i, j) = at<float>(j, i);
      float t = 0;
      for (int j = 0; j < 3; j++) t -= at<float>(j, i) * at<float>(j, 3);
      o.at<float>(i, 3) = t;
    }
    return o;
  }
  friend Mat operator*(const Mat& a, const Mat& b) {
    if (a.type_ != CV_32F || b.type_ != CV_32F || a.cols != b.rows) throw std::runtime_error("cv::Mat shim: operator* is CV_32F only");
    Mat c(a.rows, b.cols, CV_32F);
    for (int i = 0; i < a.rows; i++)
      for (int j = 0; j < b.cols; j++) {
        float s = 0;
        for (int k = 0; k < a.cols; k++) s += a.at<float>(i, k) * b.at<float>(k, j);
        c.at<float>(i, j) = s;
      }
    return c;
  }

 private:
  int type_ = CV_8U;
  std::shared_ptr<std::vector<unsigned char>> buf_;
};
struct Size {
  int width = 0, height = 0;
};
namespace cuda {
// device image header; memory is owned by the nmi context (Image), never by the GpuMat
struct GpuMat {
  unsigned char* data = nullptr;
  int rows = 0, cols = 0;
  size_t step = 0;
  bool isContinuous() const { return step == (size_t)cols; }
  void release() { data = nullptr; rows = cols = 0; step = 0; }
};
template <typename T>
struct PtrStep {
  T* data;
  size_t step;
};
}  // namespace cuda
}  // namespace cv
#endif

#ifdef NMI_COMPAT_HAVE_GLM
#include <glm/glm.hpp>
#else
namespace glm {
struct vec3 {
  float x = 0, y = 0, z = 0;
  vec3() = default;
  vec3(float a, float b, float c) : x(a), y(b), z(c) {}
  float& operator[](int i) { return i == 0 ? x : i == 1 ? y : z; }
  const float& operator[](int i) const { return i == 0 ? x : i == 1 ? y : z; }
};
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
}  // namespace glm
#endif
