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

// ---- the same overlay as a JPEG (the container cv::imwrite gives the reference's ".jpg" names, ioData.cpp:262-285)
// Baseline sequential DCT, 8-bit, three components Y / Cb / Cr (JFIF conversion) without chroma subsampling, the
// Annex K luminance quantisation table scaled libjpeg-style by `quality` (cv::imwrite's default: 95) for all three
// components, one DC and one AC Huffman table (the Annex K luminance tables) shared by them.  Not OpenCV's libjpeg bit
// for bit -- a JPEG any decoder reads, within rounding of the same picture.
namespace jpeg_detail {
struct BitWriter {
  std::vector<uint8_t>& out;
  uint32_t acc = 0;
  int nbits = 0;
  explicit BitWriter(std::vector<uint8_t>& o) : out(o) {}
  void put(uint32_t code, int len) {
    for (int i = len - 1; i >= 0; i--) {
      acc = (acc << 1) | ((code >> i) & 1u);
      if (++nbits == 8) {
        out.push_back((uint8_t)acc);
        if ((uint8_t)acc == 0xFF) out.push_back(0);  // byte stuffing
        acc = 0;
        nbits = 0;
      }
    }
  }
  void flush() {
    while (nbits != 0) put(1, 1);  // pad with ones
  }
};
struct Huff {
  uint16_t code[256];
  uint8_t len[256];
};
inline void build_huff(const uint8_t bits[16], const uint8_t* vals, Huff* h) {
  for (int i = 0; i < 256; i++) h->len[i] = 0;
  uint32_t code = 0;
  int k = 0;
  for (int l = 1; l <= 16; l++) {
    for (int i = 0; i < bits[l - 1]; i++, k++) {
      h->code[vals[k]] = (uint16_t)code++;
      h->len[vals[k]] = (uint8_t)l;
    }
    code <<= 1;
  }
}
inline int category(int v) {
  int a = v < 0 ? -v : v, n = 0;
  while (a) {
    n++;
    a >>= 1;
  }
  return n;
}
}  // namespace jpeg_detail

inline bool saveOverlayJPG(const char* fileName, const uint8_t* image, const uint8_t* synthetic, int W, int H,
                           int quality = 95) {
  using namespace jpeg_detail;
  if (W <= 0 || H <= 0 || W > 65535 || H > 65535) return false;
  static const uint8_t zz[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                 41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
  static const uint8_t qbase[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                                    14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                                    18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                                    49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
  static const uint8_t dc_bits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
  static const uint8_t dc_vals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
  static const uint8_t ac_bits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
  static const uint8_t ac_vals[162] = {
      0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
      0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72,
      0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
      0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
      0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83,
      0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
      0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
      0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
      0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
  if (quality < 1) quality = 1;
  if (quality > 100) quality = 100;
  const int scale = quality < 50 ? 5000 / quality : 200 - 2 * quality;
  uint8_t q[64];
  for (int i = 0; i < 64; i++) {
    int v = (qbase[i] * scale + 50) / 100;
    q[i] = (uint8_t)(v < 1 ? 1 : (v > 255 ? 255 : v));
  }
  Huff hdc, hac;
  build_huff(dc_bits, dc_vals, &hdc);
  build_huff(ac_bits, ac_vals, &hac);
  float c[8][8];  // c[u][x] = C(u)/2 * cos((2x+1) u pi / 16)
  for (int u = 0; u < 8; u++)
    for (int x = 0; x < 8; x++)
      c[u][x] = (u == 0 ? 0.35355339059327373f : 0.5f) * (float)std::cos((2 * x + 1) * u * 3.14159265358979323846 / 16.0);

  std::vector<uint8_t> out;
  auto put16 = [&](int v) {
    out.push_back((uint8_t)(v >> 8));
    out.push_back((uint8_t)v);
  };
  out.push_back(0xFF); out.push_back(0xD8);  // SOI
  out.push_back(0xFF); out.push_back(0xE0);  // APP0 JFIF 1.01, no density, no thumbnail
  put16(16);
  for (char ch : {'J', 'F', 'I', 'F', '\0'}) out.push_back((uint8_t)ch);
  out.push_back(1); out.push_back(1); out.push_back(0);
  put16(1); put16(1);
  out.push_back(0); out.push_back(0);
  out.push_back(0xFF); out.push_back(0xDB);  // DQT, table 0, 8-bit, zigzag order
  put16(67);
  out.push_back(0);
  for (int i = 0; i < 64; i++) out.push_back(q[zz[i]]);
  out.push_back(0xFF); out.push_back(0xC0);  // SOF0
  put16(17);
  out.push_back(8);
  put16(H); put16(W);
  out.push_back(3);
  for (int comp = 1; comp <= 3; comp++) {
    out.push_back((uint8_t)comp);
    out.push_back(0x11);  // 1 x 1 sampling
    out.push_back(0);     // quantisation table 0
  }
  out.push_back(0xFF); out.push_back(0xC4);  // DHT: DC table 0 and AC table 0
  put16(2 + (1 + 16 + 12) + (1 + 16 + 162));
  out.push_back(0x00);
  for (int i = 0; i < 16; i++) out.push_back(dc_bits[i]);
  for (int i = 0; i < 12; i++) out.push_back(dc_vals[i]);
  out.push_back(0x10);
  for (int i = 0; i < 16; i++) out.push_back(ac_bits[i]);
  for (int i = 0; i < 162; i++) out.push_back(ac_vals[i]);
  out.push_back(0xFF); out.push_back(0xDA);  // SOS
  put16(12);
  out.push_back(3);
  for (int comp = 1; comp <= 3; comp++) {
    out.push_back((uint8_t)comp);
    out.push_back(0x00);  // DC table 0, AC table 0
  }
  out.push_back(0); out.push_back(63); out.push_back(0);

  BitWriter bw(out);
  int pred[3] = {0, 0, 0};
  for (int by = 0; by < H; by += 8)
    for (int bx = 0; bx < W; bx += 8) {
      float px[3][64];
      for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) {
          const int yy = by + y < H ? by + y : H - 1, xx = bx + x < W ? bx + x : W - 1;  // edge replication
          const float R = image[(size_t)yy * W + xx], G = synthetic[(size_t)yy * W + xx];  // B = 0
          px[0][8 * y + x] = 0.299f * R + 0.587f * G - 128.0f;
          px[1][8 * y + x] = -0.168736f * R - 0.331264f * G;
          px[2][8 * y + x] = 0.5f * R - 0.418688f * G;
        }
      for (int comp = 0; comp < 3; comp++) {
        float tmp[64], F[64];
        for (int y = 0; y < 8; y++)      // rows
          for (int u = 0; u < 8; u++) {
            float a = 0;
            for (int x = 0; x < 8; x++) a += c[u][x] * px[comp][8 * y + x];
            tmp[8 * y + u] = a;
          }
        for (int u = 0; u < 8; u++)      // columns
          for (int v = 0; v < 8; v++) {
            float a = 0;
            for (int y = 0; y < 8; y++) a += c[v][y] * tmp[8 * y + u];
            F[8 * v + u] = a;
          }
        int coef[64];
        for (int i = 0; i < 64; i++) coef[i] = (int)std::lround(F[zz[i]] / (float)q[zz[i]]);
        const int diff = coef[0] - pred[comp];
        pred[comp] = coef[0];
        int cat = category(diff);
        bw.put(hdc.code[cat], hdc.len[cat]);
        if (cat) bw.put((uint32_t)(diff < 0 ? diff + (1 << cat) - 1 : diff), cat);
        int run = 0;
        for (int i = 1; i < 64; i++) {
          if (coef[i] == 0) {
            run++;
            continue;
          }
          while (run > 15) {
            bw.put(hac.code[0xF0], hac.len[0xF0]);  // ZRL
            run -= 16;
          }
          cat = category(coef[i]);
          const int sym = (run << 4) | cat;
          bw.put(hac.code[sym], hac.len[sym]);
          bw.put((uint32_t)(coef[i] < 0 ? coef[i] + (1 << cat) - 1 : coef[i]), cat);
          run = 0;
        }
        if (run) bw.put(hac.code[0x00], hac.len[0x00]);  // EOB
      }
    }
  bw.flush();
  out.push_back(0xFF); out.push_back(0xD9);  // EOI
  FILE* f = std::fopen(fileName, "wb");
  if (!f) return false;
  const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
  std::fclose(f);
  return ok;
}

// File name of the per-search overlay the reference writes (Tracking.cc:1915-1928); kept so
// result folders sort the same way.
inline std::string overlay_name(const std::string& resultsPath, int n, const nmi_grid& g, float nmi,
                                const int32_t s[3], const int32_t w[3], const char* ext = ".bmp") {
  std::ostringstream ss;
  ss << resultsPath << "/" << std::setw(4) << std::setfill('0') << n << std::setfill(' ') << "_NMI_[" << nmi
     << "]_WzyxSzyx_[" << w[2] << "," << w[1] << "," << w[0] << "," << s[2] << "," << s[1] << "," << s[0]
     << "]_grid_[" << g.nS[0] << "x" << g.nS[1] << "x" << g.nS[2] << "_" << g.nW[0] << "x" << g.nW[1] << "x"
     << g.nW[2] << "]" << ext;
  return ss.str();
}

}  // namespace nmi_compat
