// cameraSettings.hpp -- drop-in for Thirdparty/Localization/cameraSettings.hpp:32-48.
#pragma once
#include <string>

#include "nmi_compat_types.hpp"

class CameraSettings {
  std::string fileName;
  cv::Mat K_Matrix;
  glm::vec3 Camera_position, Camera_direction, Camera_up;

 public:
  CameraSettings(glm::vec3 Pos, glm::vec3 Dir, glm::vec3 Up)
      : Camera_position(Pos), Camera_direction(Dir), Camera_up(Up) {}
  CameraSettings(std::string file, cv::Mat K, glm::vec3 Pos, glm::vec3 Dir, glm::vec3 Up)
      : fileName(file), K_Matrix(K), Camera_position(Pos), Camera_direction(Dir), Camera_up(Up) {}
  glm::vec3 getPosition() { return Camera_position; }
  glm::vec3 getDirection() { return Camera_direction; }
  glm::vec3 getUp() { return Camera_up; }
  cv::Mat getK() { return K_Matrix; }
  std::string getFileName() { return fileName; }
};
