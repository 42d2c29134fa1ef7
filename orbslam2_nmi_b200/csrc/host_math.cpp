// host_math.cpp -- host-side geometry of the pose grid (C++, no CUDA).
//
// Mirrors the scalar host code the reference runs around its kernels:
//   Rendering::calculateTranslation      Thirdparty/Localization/rendering.hpp:644-665
//   Rendering::initVBO projection        rendering.hpp:196-202
//   setupCam                             Thirdparty/Localization/ioData.cpp:177-197
//   Image::Image warp matrices           Thirdparty/Localization/image.cpp:76-108
//   Tracking::CalculateNMIRelocalization src/Tracking.cc:2374-2419
//   NmiSearchKernel::isMiddle/resizeKernel  nmiSearchKernel.cpp:99-141
// plus the multi-GPU pose partitioner.  Compiled with
// -ffp-contract=off: the fp32/fp64 operation order written here is the contract
// the parity tests check against the oracle.
#include <cmath>
#include <cstring>

#include "nmi_internal.h"

namespace {

struct V3 {
  float x, y, z;
};

inline V3 column(const float* T, int c) { return V3{T[c], T[4 + c], T[8 + c]}; }


struct M3 {
  double m[9];
};

inline M3 mul(const M3& a, const M3& b) {
  M3 c;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      c.m[3 * i + j] = (a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j]) + a.m[3 * i + 2] * b.m[6 + j];
  return c;
}

// adjugate / determinant inverse (what cv::invert does for 3x3)
inline M3 inverse(const M3& a) {
  const double* m = a.m;
  const double c00 = m[4] * m[8] - m[5] * m[7];
  const double c01 = m[5] * m[6] - m[3] * m[8];
  const double c02 = m[3] * m[7] - m[4] * m[6];
  const double det = (m[0] * c00 + m[1] * c01) + m[2] * c02;
  const double id = 1.0 / det;
  M3 o;
  o.m[0] = c00 * id;
  o.m[1] = (m[2] * m[7] - m[1] * m[8]) * id;
  o.m[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o.m[3] = c01 * id;
  o.m[4] = (m[0] * m[8] - m[2] * m[6]) * id;
  o.m[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o.m[6] = c02 * id;
  o.m[7] = (m[1] * m[6] - m[0] * m[7]) * id;
  o.m[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  return o;
}

// image.cpp:77,86,95: start = -(n-1)/2*step with integer division, then += step
inline double cell_angle(int n, float step, int i) {
  const float start = static_cast<float>(-(n - 1) / 2) * step;
  double th = static_cast<double>(start);
  for (int k = 0; k < i; ++k) th += static_cast<double>(step);
  return th;
}

}  // namespace

namespace nmi {

void make_view_const(const nmi_camera& cam, const float Twc[16], ViewConst* vc) {
  for (int i = 0; i < 3; ++i) {
    vc->r0[i] = Twc[4 * i + 0];
    vc->r1[i] = Twc[4 * i + 1];
    vc->r2[i] = Twc[4 * i + 2];
  }
  vc->kx = static_cast<float>(cam.fx / cam.cx);  // rendering.hpp:196
  vc->ky = static_cast<float>(cam.fy / cam.cy);  // rendering.hpp:197
  vc->hw = 0.5f * static_cast<float>(cam.W);
  vc->hh = 0.5f * static_cast<float>(cam.H);
  vc->zn = static_cast<float>(cam.zn);
  vc->zf = static_cast<float>(cam.zf);
  vc->W = cam.W;
  vc->H = cam.H;
  const int s = static_cast<int>(std::lround(cam.point_size));  // rendering.hpp:307
  vc->s = s < 1 ? 1 : s;
}

uint64_t pack_key(float max_score, int64_t index) {
  uint32_t bits;
  std::memcpy(&bits, &max_score, 4);
  const uint32_t low = index < 0 ? 0u : 0xFFFFFFFFu - static_cast<uint32_t>(index);
  return (static_cast<uint64_t>(bits) << 32) | low;
}

}  // namespace nmi

extern "C" {

// Statement for statement what the reference computes between the prior pose and the translation of a
// grid cell -- setupCam (ioData.cpp:177-197), Rendering::setCamera, calculateTranslation
// (rendering.hpp:644-665) with glm::rotate of GLM 0.9.7.1 (gtx/rotate_vector over
// gtc/matrix_transform's axis-angle matrix) -- in the float / double mix a C++11 compiler gives that
// source (pow(float, int) is a double).  Bit-identical to the reference's own lines compiled here
// (the geometry target of oracle/Makefile.ref; tests/test_reference_geometry.py).  Note what the reference does:
// `dir` is stored as pos + z and subtracted again, so its camera axes carry the rounding of a sum at the
// magnitude of the camera POSITION; the third axis comes from a -90 degree rotation whose cosine is
// cosf(-pi/2) = -4.4e-8, not 0.
void nmi_cell_translation(const float Twc[16], const nmi_grid* g, int sx, int sy, int sz,
                          float t[3]) {
  // setupCam
  const float pos[3] = {Twc[3], Twc[7], Twc[11]};
  const float dir[3] = {Twc[2] + pos[0], Twc[6] + pos[1], Twc[10] + pos[2]};
  const float up[3] = {Twc[1], Twc[5], Twc[9]};
  // calculateTranslation
  const float x_offset = (static_cast<float>(g->nS[0]) - 1.0f) / 2.0f;
  const float y_offset = (static_cast<float>(g->nS[1]) - 1.0f) / 2.0f;
  const float z_offset = (static_cast<float>(g->nS[2]) - 1.0f) / 2.0f;
  float length = static_cast<float>(std::sqrt(static_cast<double>(up[0]) * up[0] + static_cast<double>(up[1]) * up[1] +
                                              static_cast<double>(up[2]) * up[2]));
  const float dy[3] = {up[0] / length, up[1] / length, up[2] / length};
  const float d0 = dir[0] - pos[0], d1 = dir[1] - pos[1], d2 = dir[2] - pos[2];
  length = static_cast<float>(std::sqrt(static_cast<double>(d0) * d0 + static_cast<double>(d1) * d1 +
                                        static_cast<double>(d2) * d2));
  const float dz[3] = {-(d0 / length), -(d1 / length), -(d2 / length)};  // -1.0 * (float)(..): exact negation
  // dir_x = glm::rotate(dir_y, glm::radians(-90.0f), dir_z)
  const float angle = -90.0f * 0.01745329251994329576923690768489f;
  const float c = std::cos(angle), sn = std::sin(angle);  // float overloads: cosf / sinf
  const float inv = 1.0f / std::sqrt((dz[0] * dz[0] + dz[1] * dz[1]) + dz[2] * dz[2]);  // normalize = v * inversesqrt(dot)
  const float ax[3] = {dz[0] * inv, dz[1] * inv, dz[2] * inv};
  const float tmp[3] = {(1.0f - c) * ax[0], (1.0f - c) * ax[1], (1.0f - c) * ax[2]};
  float R[3][3];  // R[column][row]
  R[0][0] = c + tmp[0] * ax[0];
  R[0][1] = 0 + tmp[0] * ax[1] + sn * ax[2];
  R[0][2] = 0 + tmp[0] * ax[2] - sn * ax[1];
  R[1][0] = 0 + tmp[1] * ax[0] - sn * ax[2];
  R[1][1] = c + tmp[1] * ax[1];
  R[1][2] = 0 + tmp[1] * ax[2] + sn * ax[0];
  R[2][0] = 0 + tmp[2] * ax[0] + sn * ax[1];
  R[2][1] = 0 + tmp[2] * ax[1] - sn * ax[0];
  R[2][2] = c + tmp[2] * ax[2];
  const float dx[3] = {R[0][0] * dy[0] + R[1][0] * dy[1] + R[2][0] * dy[2], R[0][1] * dy[0] + R[1][1] * dy[1] + R[2][1] * dy[2],
                       R[0][2] * dy[0] + R[1][2] * dy[1] + R[2][2] * dy[2]};
  const float cx = (static_cast<float>(sx) - x_offset) * g->stepT[0];
  const float cy = (static_cast<float>(sy) - y_offset) * g->stepT[1];
  const float cz = (static_cast<float>(sz) - z_offset) * g->stepT[2];
  for (int k = 0; k < 3; ++k) t[k] = (cx * dx[k] + cy * dy[k]) + cz * dz[k];
}

void nmi_cell_homography_inv(const nmi_camera* cam, const nmi_grid* g, int wx, int wy, int wz,
                             float minv[9]) {
  const double tx = cell_angle(g->nW[0], g->stepR[0], wx);
  const double ty = cell_angle(g->nW[1], g->stepR[1], wy);
  const double tz = cell_angle(g->nW[2], g->stepR[2], wz);
  const double cX = std::cos(tx), sX = std::sin(tx);
  const double cY = std::cos(ty), sY = std::sin(ty);
  const double cZ = std::cos(tz), sZ = std::sin(tz);
  const M3 Rx{{1, 0, 0, 0, cX, -sX, 0, sX, cX}};   // image.cpp:98-101
  const M3 Ry{{cY, 0, sY, 0, 1, 0, -sY, 0, cY}};   // image.cpp:89-92
  const M3 Rz{{cZ, -sZ, 0, sZ, cZ, 0, 0, 0, 1}};   // image.cpp:80-83
  const M3 K{{cam->fx, 0, cam->cx, 0, cam->fy, cam->cy, 0, 0, 1}};
  const M3 R = mul(mul(Rz, Ry), Rx);               // image.cpp:103
  const M3 M = mul(mul(K, R), inverse(K));         // image.cpp:104
  const M3 Mi = inverse(M);                        // warpPerspective maps dst -> src
  for (int i = 0; i < 9; ++i) minv[i] = static_cast<float>(Mi.m[i]);
}

void nmi_apply_winner(const float Twc[16], const nmi_grid* g, const int32_t s[3],
                      const int32_t w[3], float out[16]) {
  // Tracking.cc:2383-2385: integer n/2
  const float rx = static_cast<float>(w[0] - g->nW[0] / 2) * g->stepR[0];
  const float ry = static_cast<float>(w[1] - g->nW[1] / 2) * g->stepR[1];
  const float rz = static_cast<float>(w[2] - g->nW[2] / 2) * g->stepR[2];
  const float Rx[9] = {1, 0, 0, 0, std::cos(rx), -std::sin(rx), 0, std::sin(rx), std::cos(rx)};
  const float Ry[9] = {std::cos(ry), 0, std::sin(ry), 0, 1, 0, -std::sin(ry), 0, std::cos(ry)};
  const float Rz[9] = {std::cos(rz), -std::sin(rz), 0, std::sin(rz), std::cos(rz), 0, 0, 0, 1};
  auto mul3 = [](const float* a, const float* b, float* c) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        c[3 * i + j] = (a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j]) + a[3 * i + 2] * b[6 + j];
  };
  float RzRy[9], R[9], t[3];
  mul3(Rz, Ry, RzRy);
  mul3(RzRy, Rx, R);  // Tracking.cc:2407
  nmi_cell_translation(Twc, g, s[0], s[1], s[2], t);  // calculateTranslationCV
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      out[4 * i + j] = (Twc[4 * i] * R[j] + Twc[4 * i + 1] * R[3 + j]) + Twc[4 * i + 2] * R[6 + j];
    out[4 * i + 3] = Twc[4 * i + 3] + t[i];  // Tracking.cc:2414-2416
  }
  for (int j = 0; j < 4; ++j) out[12 + j] = Twc[12 + j];
}

int nmi_grid_is_middle(const nmi_grid* g, const int32_t s[3], const int32_t w[3]) {
  for (int k = 0; k < 3; ++k)
    if (s[k] != g->nS[k] / 2 || w[k] != g->nW[k] / 2) return 0;
  return 1;
}

void nmi_grid_resize(nmi_grid* g, const int32_t s[3], const int32_t w[3]) {
  const float factor = 0.5f;         // nmi_prop_STEPFACTOR, allProperties.hpp:33
  const double min_rot = 0.001;      // nmi_prop_MIN_KERNEL_ROTATION, :48
  const double min_trans = 0.005;    // nmi_prop_MIN_KERNEL_TRANSLATION, :49
  for (int k = 0; k < 3; ++k) {
    const bool s_edge = (s[k] == g->nS[k] - 1 || s[k] == 0) && g->nS[k] > 1;
    const bool w_edge = (w[k] == g->nW[k] - 1 || w[k] == 0) && g->nW[k] > 1;
    if (!s_edge) g->stepT[k] *= factor;
    if (!w_edge) g->stepR[k] *= factor;
  }
  for (int k = 0; k < 3; ++k) {
    if (g->stepT[k] < min_trans) g->nS[k] = 1;
    if (g->stepR[k] < min_rot) g->nW[k] = 1;
  }
}

int nmi_partition(const nmi_grid* g, int rank, int world, int* axis, int* begin, int* end) {
  if (!g || world < 1 || rank < 0 || rank >= world) return NMI_ERR_INVALID;
  const int nS = g->nS[0] * g->nS[1] * g->nS[2];
  const int nW = g->nW[0] * g->nW[1] * g->nW[2];
  // shard the synthetic views (keeps render-once / score-nW-times per GPU); when
  // there are fewer views than GPUs shard the rotation cells instead.
  const int ax = (nS >= world || nS >= nW) ? 0 : 1;
  const int n = ax == 0 ? nS : nW;
  const int base = n / world, rem = n % world;
  const int b = rank * base + (rank < rem ? rank : rem);
  const int e = b + base + (rank < rem ? 1 : 0);
  if (axis) *axis = ax;
  if (begin) *begin = b;
  if (end) *end = e;
  return NMI_OK;
}

int nmi_decode_key(const nmi_grid* g, uint64_t key, nmi_result* out) {
  if (!g || !out) return NMI_ERR_INVALID;
  const uint32_t hi = static_cast<uint32_t>(key >> 32), lo = static_cast<uint32_t>(key);
  std::memcpy(&out->best_score, &hi, 4);
  out->key = key;
  if (key == NMI_KEY_RETRY) {
    out->best_index = -1;
    for (int k = 0; k < 3; ++k) out->best_s[k] = out->best_w[k] = -1;
    return NMI_ERR_RETRY;
  }
  if (lo == 0) {
    out->best_index = -1;
    for (int k = 0; k < 3; ++k) out->best_s[k] = out->best_w[k] = -1;
    return NMI_ERR_NO_WINNER;
  }
  uint64_t l = 0xFFFFFFFFu - lo;
  out->best_index = static_cast<int64_t>(l);
  out->best_s[0] = static_cast<int32_t>(l % g->nS[0]); l /= g->nS[0];
  out->best_s[1] = static_cast<int32_t>(l % g->nS[1]); l /= g->nS[1];
  out->best_s[2] = static_cast<int32_t>(l % g->nS[2]); l /= g->nS[2];
  out->best_w[0] = static_cast<int32_t>(l % g->nW[0]); l /= g->nW[0];
  out->best_w[1] = static_cast<int32_t>(l % g->nW[1]); l /= g->nW[1];
  out->best_w[2] = static_cast<int32_t>(l);
  return NMI_OK;
}

}  // extern "C"
