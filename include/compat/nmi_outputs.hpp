// nmi_outputs.hpp -- the result files the reference writes after the NMI path ran
// (SURVEY.md 8(f) row 4), host-only, so existing evaluation scripts keep working:
//   * FrameTrajectory<...>.txt / <...>_twc.txt  -- System::SaveFullTrajectory
//     (src/System.cc:514-599): "<frame> <time>[ KF[, NMI][, FAILED]] tx ty tz qx qy qz qw"
//   * red/green overlay of a warped camera frame and a render -- saveImage / saveBMP
//     (Thirdparty/Localization/ioData.cpp:199-347).  The reference hands the overlay to
//     cv::imwrite (JPEG); there is no JPEG encoder here, so the container is 24-bit BMP.
// The ORB-SLAM2 keyframe bookkeeping that feeds SaveFullTrajectory (spanning-tree walk,
// relative frame poses) is outside the hot path; TrajectoryRecorder takes the final
// camera-to-world pose of each frame directly.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <string>
#include <vector>

#include "../nmi_b200.h"

namespace nmi_compat {

// Converter::toQuaternion (src/Converter.cc:138-150): Eigen::Quaterniond(R) in double,
// returned as x, y, z, w.  Eigen's constructor is Shepperd's method: trace branch when
// trace > 0, otherwise the largest diagonal element picks the pivot.
inline void rotation_to_quaternion(const double R[9], float q_xyzw[4]) {
  double q[4];  // x y z w
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0.0) {
    double t = std::sqrt(tr + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t;
    q[1] = (R[2] - R[6]) * t;
    q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double t = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * t;
    q[j] = (R[3 * j + i] + R[3 * i + j]) * t;
    q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
  }
  for (int n = 0; n < 4; n++) q_xyzw[n] = (float)q[n];
}

// cv::Mat's default stream format for a 4x4 CV_32F: "[a, b, c, d;\n e, ...]", %.8g.
inline std::string format_mat4(const float T[16]) {
  std::string s = "[";
  char buf[32];
  for (int r = 0; r < 4; r++) {
    for (int c = 0; c < 4; c++) {
      std::snprintf(buf, sizeof buf, "%.8g", (double)T[4 * r + c]);
      s += buf;
      if (c < 3) s += ", ";
    }
    s += r < 3 ? ";\n " : "]";
  }
  return s;
}

class TrajectoryRecorder {
 public:
  struct Entry {
    int frame = 0;
    double time = 0.0;
    float Twc[16] = {0};
    bool keyframe = false, nmi = false, failed = false, lost = false;
    std::vector<std::vector<float>> previous;  // KeyFrame::mvPreviousPoses (4x4 each)
  };

  void add(int frame, double time, const float Twc[16], bool keyframe = true, bool lost = false) {
    Entry e;
    e.frame = frame;
    e.time = time;
    for (int i = 0; i < 16; i++) e.Twc[i] = Twc[i];
    e.keyframe = keyframe;
    e.lost = lost;
    entries_.push_back(e);
  }
  // Records the outcome of nmi_relocalize() for a keyframe: pose, NMI / FAILED tags and the
  // poses it went through (Tracking.cc:2094-2097, System.cc:578-591).
  void add(int frame, double time, const nmi_reloc_result& r, bool keyframe = true) {
    add(frame, time, r.Twc, keyframe, false);
    Entry& e = entries_.back();
    e.nmi = r.relocalized != 0;
    e.failed = r.failed != 0;
    for (int p = 0; p < r.n_prev; p++) e.previous.emplace_back(r.prev_Twc[p], r.prev_Twc[p] + 16);
  }
  size_t size() const { return entries_.size(); }
  const Entry& operator[](size_t i) const { return entries_[i]; }

  // System::SaveFullTrajectory (src/System.cc:514-599): writes <filename>.txt and
  // <filename>_twc.txt; frames whose tracking was lost are skipped.
  bool SaveFullTrajectory(const std::string& filename) const {
    std::ofstream f1((filename + ".txt").c_str()), f2((filename + "_twc.txt").c_str());
    if (!f1 || !f2) return false;
    f1 << std::fixed;
    f2 << std::fixed;
    for (const Entry& e : entries_) {
      if (e.lost) continue;
      double R[9];
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) R[3 * r + c] = e.Twc[4 * r + c];
      float q[4];
      rotation_to_quaternion(R, q);
      f1 << e.frame << " " << std::setprecision(6) << e.time;
      f2 << e.frame << " " << std::setprecision(6) << e.time;
      if (e.keyframe) {
        f1 << " KF";
        f2 << " KF";
        if (e.nmi) {
          f1 << ", NMI";
          f2 << ", NMI";
          f2 << std::endl << "//////////Previous Poses\\\\\\\\\\\\\\\\\\\\" << std::endl;
          for (const auto& p : e.previous) f2 << format_mat4(p.data()) << std::endl;
          f2 << "//////////Previous Poses End\\\\\\\\\\\\\\\\\\\\";
        }
        if (e.failed) {
          f1 << ", FAILED";
          f2 << ", FAILED";
        }
      }
      f1 << " " << std::setprecision(9) << e.Twc[3] << " " << e.Twc[7] << " " << e.Twc[11] << " " << q[0]
         << " " << q[1] << " " << q[2] << " " << q[3] << std::endl;
      f2 << std::endl << format_mat4(e.Twc) << std::endl;
    }
    return true;
  }

 private:
  std::vector<Entry> entries_;
};

// Overlay of ioData.cpp:262-285 (getImageNoHalf): blue 0, red = camera image, green =
// render.  Both inputs are top-down W*H u8 here (the reference's render is bottom-up and is
// flipped while it is copied in; ours is already top-down).  Written as a 24-bit BMP.
inline bool saveOverlayBMP(const char* fileName, const uint8_t* image, const uint8_t* synthetic, int W,
                           int H) {
  FILE* f = std::fopen(fileName, "wb");
  if (!f) return false;
  const int pad = (4 - (W * 3) % 4) % 4;
  const uint32_t size = 54u + (uint32_t)(3 * W + pad) * (uint32_t)H;
  uint8_t hdr[54] = {'B', 'M'};
  auto put32 = [&](int at, uint32_t v) {
    for (int i = 0; i < 4; i++) hdr[at + i] = (uint8_t)(v >> (8 * i));
  };
  put32(2, size);
  put32(10, 54);
  put32(14, 40);
  put32(18, (uint32_t)W);
  put32(22, (uint32_t)H);
  hdr[26] = 1;
  hdr[28] = 24;
  put32(34, size - 54u);
  std::fwrite(hdr, 1, 54, f);
  std::vector<uint8_t> row((size_t)3 * W + pad, 0);
  for (int y = H - 1; y >= 0; y--) {  // BMP rows are bottom-up
    for (int x = 0; x < W; x++) {
      row[3 * x] = 0;
      row[3 * x + 1] = synthetic[(size_t)y * W + x];
      row[3 * x + 2] = image[(size_t)y * W + x];
    }
    std::fwrite(row.data(), 1, row.size(), f);
  }
  std::fclose(f);
  return true;
}

// File name of the per-search overlay the reference writes (Tracking.cc:1915-1928); kept so
// result folders sort the same way.
inline std::string overlay_name(const std::string& resultsPath, int n, const nmi_grid& g, float nmi,
                                const int32_t s[3], const int32_t w[3]) {
  std::ostringstream ss;
  ss << resultsPath << "/" << std::setw(4) << std::setfill('0') << n << std::setfill(' ') << "_NMI_[" << nmi
     << "]_WzyxSzyx_[" << w[2] << "," << w[1] << "," << w[0] << "," << s[2] << "," << s[1] << "," << s[0]
     << "]_grid_[" << g.nS[0] << "x" << g.nS[1] << "x" << g.nS[2] << "_" << g.nW[0] << "x" << g.nW[1] << "x"
     << g.nW[2] << "].bmp";
  return ss.str();
}

}  // namespace nmi_compat
