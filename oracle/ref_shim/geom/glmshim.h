/* glmshim.h -- TEST INFRASTRUCTURE ONLY.  The handful of GLM 0.9.7.1 names
 * (Thirdparty/Localization/Localization.vcxproj:50 pins the version by include path; GLM itself is
 * not in the reference tree) used by ioData.cpp:177-197, cameraSettings.hpp and
 * rendering.hpp:642-694: vec3 (x/y/z, operator[], scalar * vec, vec + vec), radians, and
 * gtx/rotate_vector's rotate(v, angle, normal) = mat3(rotate(mat4(1), angle, normal)) * v with
 * gtc/matrix_transform's axis-angle matrix (axis normalised, temp = (1 - cos) * axis, columns as
 * published) -- restated from GLM's published source, float arithmetic, no FMA contraction.  */
#ifndef NMI_GLMSHIM_H_
#define NMI_GLMSHIM_H_
#include <cmath>
namespace glm {
struct vec3 {
  float x, y, z;
  vec3() : x(0), y(0), z(0) {}
  vec3(float a, float b, float c) : x(a), y(b), z(c) {}
  float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  const float& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
struct vec2 { float x, y; };
inline vec3 operator*(float s, const vec3& v) { return vec3(s * v.x, s * v.y, s * v.z); }
inline vec3 operator*(const vec3& v, float s) { return vec3(v.x * s, v.y * s, v.z * s); }
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 normalize(const vec3& v) { return v * (1.0f / std::sqrt(dot(v, v))); }  // v * inversesqrt(dot(v, v))
inline vec3 rotate(const vec3& v, float angle, const vec3& normal) {
  const float c = std::cos(angle), s = std::sin(angle);
  const vec3 axis = normalize(normal);
  const vec3 temp = (1.0f - c) * axis;
  float R[3][3];  // R[col][row], GLM is column-major
  R[0][0] = c + temp[0] * axis[0];
  R[0][1] = 0 + temp[0] * axis[1] + s * axis[2];
  R[0][2] = 0 + temp[0] * axis[2] - s * axis[1];
  R[1][0] = 0 + temp[1] * axis[0] - s * axis[2];
  R[1][1] = c + temp[1] * axis[1];
  R[1][2] = 0 + temp[1] * axis[2] + s * axis[0];
  R[2][0] = 0 + temp[2] * axis[0] + s * axis[1];
  R[2][1] = 0 + temp[2] * axis[1] - s * axis[0];
  R[2][2] = c + temp[2] * axis[2];
  // rotate(mat4(1), ...) multiplies the identity by R (exact), mat3(...) keeps the upper 3x3;
  // mat3 * vec3: result[row] = m[0][row] * v.x + m[1][row] * v.y + m[2][row] * v.z
  return vec3(R[0][0] * v.x + R[1][0] * v.y + R[2][0] * v.z, R[0][1] * v.x + R[1][1] * v.y + R[2][1] * v.z,
              R[0][2] * v.x + R[1][2] * v.y + R[2][2] * v.z);
}
}  // namespace glm
#endif
