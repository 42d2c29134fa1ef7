// ioData.hpp -- drop-in for the part of Thirdparty/Localization/ioData.hpp the hot path
// uses: setupCam(Twc, K) (ioData.cpp:177-197).  The save*() debug-overlay writers are out
// of scope.
#pragma once
#include "cameraSettings.hpp"

// pos = Twc[:3,3]; dir = pos + Twc[:3,2] (a point one unit ahead); up = Twc[:3,1] (CV +y,
// i.e. image-down: the render is flipped back by NMI.cu:82 in the reference).
inline CameraSettings setupCam(cv::Mat& PosInverse, cv::Mat& K_Mat) {
  glm::vec3 pos(PosInverse.at<float>(0, 3), PosInverse.at<float>(1, 3), PosInverse.at<float>(2, 3));
  glm::vec3 dir(PosInverse.at<float>(0, 2) + pos.x, PosInverse.at<float>(1, 2) + pos.y,
                PosInverse.at<float>(2, 2) + pos.z);
  glm::vec3 up(PosInverse.at<float>(0, 1), PosInverse.at<float>(1, 1), PosInverse.at<float>(2, 1));
  return CameraSettings("", K_Mat, pos, dir, up);
}
