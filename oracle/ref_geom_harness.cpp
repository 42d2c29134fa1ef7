// ref_geom_harness.cpp -- TEST INFRASTRUCTURE ONLY.  C ABI over the reference's own host-side
// GEOMETRY code, compiled unmodified from /root/reference by oracle/Makefile.ref (g++, no GPU):
//   image.cpp (whole file)                     Image::Image warp matrices :76-108, resizeKernel :236-268   (a5)
//   ioData.cpp:177-197                         setupCam(Twc, K) -> (pos, dir, up)                          (a2)
//   rendering.hpp:642-694                      Rendering::calculateTranslation / calculateTranslationCV     (a3)
//   src/Tracking.cc:2374-2419                  Tracking::CalculateNMIRelocalization                         (a14)
// The last three are functions inside files that need all of OpenGL / ORB-SLAM2 to compile, so the
// Makefile cuts exactly those line ranges out of the reference where it lies (sed -n, into
// oracle/_ref/gen/*.inc, never committed) and this file supplies the declarations around them: a
// Rendering<> / Tracking class with just the members those functions read.  OpenCV and GLM are not in
// the reference tree; oracle/ref_shim/geom/{cvshim,glmshim}.h restate the few operations used.
#include <cstring>

#include "cvshim.h"  // (standard headers come in here, before the access hack below)
#include "glmshim.h"
// Image::warpingMatrixes has no getter and sits in the class's default-private part: read the header as
// `struct Image` (layout and mangled member names are the same; the shim headers are already in)
#define class struct
#include "image.hpp"
#undef class
#include "cameraSettings.hpp"
#include "nmiSearchKernel.hpp"
#include "glmshim.h"

#undef nmi_prop_RENDER
#define nmi_prop_RENDER 4
typedef float GLfloat;

// rendering.hpp:64-163 declares many more members; these are the ones :642-690 reads (same types:
// rendering.hpp:71-76, :111-113)
template <unsigned char RenderingMode>
class Rendering {
 public:
  int numSynthX, numSynthY, numSynthZ;
  float stepX, stepY, stepZ;
  glm::vec3 Camera_pos, Camera_direction, Camera_up;
  glm::vec3 calculateTranslation(int synthx, int synthy, int synthz);
  cv::Mat calculateTranslationCV(int synthx, int synthy, int synthz);
};
#include "gen/rendering_translation.inc"  // rendering.hpp:642-694, verbatim

#include "gen/setupCam.inc"  // ioData.cpp:177-197, verbatim

class Tracking {  // include/Tracking.h declares it inside ORB_SLAM2::Tracking; only this member is compiled
 public:
  cv::Mat CalculateNMIRelocalization(cv::Mat Twc, NmiSearchKernel& nmiKernel, Rendering<nmi_prop_RENDER>& renderer);
};
#include "gen/calc_reloc.inc"  // src/Tracking.cc:2374-2419, verbatim

namespace {
cv::Mat mat4f(const float* p) {
  cv::Mat m(4, 4, CV_32F);
  for (int i = 0; i < 16; i++) m.at<float>(i / 4, i % 4) = p[i];
  return m;
}
Rendering<nmi_prop_RENDER> renderer_for(const float* Twc, const int* nS, const float* stepT) {
  cv::Mat T = mat4f(Twc), K = cv::Mat::eye(3, 3, CV_64F);
  CameraSettings s = setupCam(T, K);  // what Tracking.cc:1873-1876 hands to Rendering::setCamera
  Rendering<nmi_prop_RENDER> r;
  r.numSynthX = nS[0]; r.numSynthY = nS[1]; r.numSynthZ = nS[2];
  r.stepX = stepT[0]; r.stepY = stepT[1]; r.stepZ = stepT[2];
  r.Camera_pos = s.getPosition(); r.Camera_direction = s.getDirection(); r.Camera_up = s.getUp();
  return r;
}
}  // namespace

extern "C" {

// Image::Image (image.cpp:33-111): forward matrices K R K^-1 (double) of every rotation cell,
// out[((z * nWy + y) * nWx + x) * 9 ...], optionally after Image::resizeKernel(new counts / steps)
int nmirefg_warp_matrices(const int nW[3], const float stepR[3], int W, int H, double fx, double fy, double cx, double cy,
                          const int* resize_nW, const float* resize_stepR, double* out) {
  cv::Mat K = cv::Mat::eye(3, 3, CV_64F);  // localization.cpp:165-169
  K.at<double>(0, 0) = fx; K.at<double>(1, 1) = fy; K.at<double>(0, 2) = cx; K.at<double>(1, 2) = cy;
  Image img(nW[2], nW[1], nW[0], stepR[2], stepR[1], stepR[0], W, H, K);
  int n[3] = {nW[0], nW[1], nW[2]};
  if (resize_nW && resize_stepR) {
    img.resizeKernel(resize_nW[0], resize_nW[1], resize_nW[2], resize_stepR[0], resize_stepR[1], resize_stepR[2]);
    n[0] = resize_nW[0]; n[1] = resize_nW[1]; n[2] = resize_nW[2];
  }
  size_t k = 0;
  for (int z = 0; z < n[2]; z++)
    for (int y = 0; y < n[1]; y++)
      for (int x = 0; x < n[0]; x++) {
        const cv::Mat& m = img.warpingMatrixes[z][y][x];
        if (m.rows != 3 || m.cols != 3 || m.type() != CV_64F) return -1;
        for (int i = 0; i < 9; i++) out[k++] = m.at<double>(i / 3, i % 3);
      }
  return (int)k / 9;
}

// setupCam (ioData.cpp:177-197): pos, dir, up
void nmirefg_setup_cam(const float Twc[16], float pos[3], float dir[3], float up[3]) {
  cv::Mat T = mat4f(Twc), K = cv::Mat::eye(3, 3, CV_64F);
  CameraSettings s = setupCam(T, K);
  const glm::vec3 p = s.getPosition(), d = s.getDirection(), u = s.getUp();
  pos[0] = p.x; pos[1] = p.y; pos[2] = p.z;
  dir[0] = d.x; dir[1] = d.y; dir[2] = d.z;
  up[0] = u.x; up[1] = u.y; up[2] = u.z;
}

// Rendering::calculateTranslation (rendering.hpp:644-665) after setupCam + setCamera, as
// Tracking.cc:1873-1882 calls it
void nmirefg_cell_translation(const float Twc[16], const int nS[3], const float stepT[3], int sx, int sy, int sz,
                              float t[3]) {
  Rendering<nmi_prop_RENDER> r = renderer_for(Twc, nS, stepT);
  const glm::vec3 v = r.calculateTranslation(sx, sy, sz);
  t[0] = v.x; t[1] = v.y; t[2] = v.z;
}

// Tracking::CalculateNMIRelocalization (Tracking.cc:2374-2419): winner -> new Twc
void nmirefg_apply_winner(const float Twc[16], const int nS[3], const int nW[3], const float stepT[3],
                          const float stepR[3], const int s[3], const int w[3], float out[16]) {
  Rendering<nmi_prop_RENDER> r = renderer_for(Twc, nS, stepT);
  NmiSearchKernel k(nS[0], nS[1], nS[2], nW[0], nW[1], nW[2], stepT[0], stepT[1], stepT[2], stepR[0], stepR[1], stepR[2]);
  k.setBest(s[0], s[1], s[2], w[0], w[1], w[2], 0.0f);
  Tracking trk;
  cv::Mat n = trk.CalculateNMIRelocalization(mat4f(Twc), k, r);
  for (int i = 0; i < 16; i++) out[i] = n.at<float>(i / 4, i % 4);
}

}  // extern "C"
