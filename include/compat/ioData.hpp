// ioData.hpp -- drop-in for the part of Thirdparty/Localization/ioData.hpp the hot path
// uses: setupCam(Twc, K) (ioData.cpp:177-197).  The save*() debug-overlay writers are out
// of scope.
#pragma once
#include "cameraSettings.hpp"
#include "nmi_compat.hpp"

// pos = Twc[:3,3]; dir = pos + Twc[:3,2] (a point one unit ahead); up = Twc[:3,1] (CV +y,
// i.e. image-down: the render is flipped back by NMI.cu:82 in the reference).
inline CameraSettings setupCam(cv::Mat& PosInverse, cv::Mat& K_Mat) {
  glm::vec3 pos(PosInverse.at<float>(0, 3), PosInverse.at<float>(1, 3), PosInverse.at<float>(2, 3));
  glm::vec3 dir(PosInverse.at<float>(0, 2) + pos.x, PosInverse.at<float>(1, 2) + pos.y,
                PosInverse.at<float>(2, 2) + pos.z);
  glm::vec3 up(PosInverse.at<float>(0, 1), PosInverse.at<float>(1, 1), PosInverse.at<float>(2, 1));
  if (PosInverse.rows == 4 && PosInverse.cols == 4) {  // so that Rendering::setCamera can get the exact pose back
    float T[16];
    for (int i = 0; i < 16; i++) T[i] = PosInverse.at<float>(i / 4, i % 4);
    const float p[3] = {pos.x, pos.y, pos.z}, d[3] = {dir.x, dir.y, dir.z}, u[3] = {up.x, up.y, up.z};
    nmi_compat::remember_setup_cam(p, d, u, T);
  }
  return CameraSettings("", K_Mat, pos, dir, up);
}
