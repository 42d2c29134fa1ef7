// compat.cpp -- out-of-line parts of the reference-named C++ drop-ins (include/compat/):
// CUDAF::NMIWithCuda_noMask, NmiObjects, the process-wide context, the YAML reader and the
// .xyz loader.  Host C++ only; everything numerical goes through the C ABI.
#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <ctime>
#include <filesystem>
#include <fstream>
#include <iomanip>
#include <sstream>

#include "../../include/compat/helperFunctions.hpp"
#include "../../include/compat/kernel.cuh"
#include "../../include/compat/localization.hpp"

namespace nmi_compat {

static nmi_ctx* g_ctx = nullptr;
static bool g_owned = false;
static nmi_flags g_flags = {256, NMI_SCORE_SUC, 1, 0};
static nmi_camera g_cam = {0, 0, 0, 0, 0, 0, 5.0, 30.0, 3.0f};

[[noreturn]] void die(const char* where) {
  // checkCudaErrors semantics (kernel.cu:53): report and terminate
  std::fprintf(stderr, "%s: %s\n", where, nmi_last_error());
  std::exit(EXIT_FAILURE);
}

nmi_ctx* context() {
  if (!g_ctx) {
    const char* dev = std::getenv("NMI_DEVICE");
    if (nmi_ctx_create(dev ? std::atoi(dev) : 0, &g_ctx) != NMI_OK) die("nmi_ctx_create");
    g_owned = true;
  }
  return g_ctx;
}
void set_context(nmi_ctx* ctx) {
  if (g_ctx && g_owned && g_ctx != ctx) nmi_ctx_destroy(g_ctx);
  g_ctx = ctx;
  g_owned = false;
}
void shutdown() {
  if (g_ctx && g_owned) nmi_ctx_destroy(g_ctx);
  g_ctx = nullptr;
  g_owned = false;
}
nmi_flags& flags() { return g_flags; }
nmi_camera& camera() { return g_cam; }
void apply_camera() {
  if (g_cam.W > 0 && g_cam.H > 0 && g_cam.fx != 0 && g_cam.cx != 0)
    check(nmi_set_camera(context(), &g_cam), "nmi_set_camera");
}

// ---- YAML (OpenCV FileStorage subset) -----------------------------------------------
static std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}

Yaml::Yaml(const std::string& path) {
  std::ifstream in(path);
  if (!in.is_open()) return;
  ok_ = true;
  std::string line, mat_key;
  M cur;
  bool in_data = false;
  std::string data_acc;
  auto finish_mat = [&]() {
    if (!mat_key.empty()) mats_[mat_key] = cur;
    mat_key.clear();
    cur = M();
  };
  auto parse_data = [&](const std::string& txt) {
    std::string t = txt;
    for (char& ch : t)
      if (ch == '[' || ch == ']' || ch == ',') ch = ' ';
    std::istringstream ss(t);
    double v;
    while (ss >> v) cur.data.push_back(v);
  };
  while (std::getline(in, line)) {
    const size_t hash = line.find('#');
    if (hash != std::string::npos && line.find('"') == std::string::npos) line = line.substr(0, hash);
    if (trim(line).empty() || line[0] == '%' || trim(line) == "---") continue;
    if (in_data) {
      data_acc += " " + line;
      if (line.find(']') != std::string::npos) {
        parse_data(data_acc);
        in_data = false;
        finish_mat();
      }
      continue;
    }
    const size_t colon = line.find(':');
    if (colon == std::string::npos) continue;
    const bool indented = line[0] == ' ' || line[0] == '\t';
    const std::string key = trim(line.substr(0, colon));
    std::string val = trim(line.substr(colon + 1));
    if (!indented) {
      finish_mat();
      if (val.find("!!opencv-matrix") != std::string::npos) {
        mat_key = key;
        continue;
      }
      if (val.size() >= 2 && val.front() == '"' && val.back() == '"') val = val.substr(1, val.size() - 2);
      scalars_[key] = val;
    } else if (!mat_key.empty()) {
      if (key == "rows") cur.rows = std::atoi(val.c_str());
      else if (key == "cols") cur.cols = std::atoi(val.c_str());
      else if (key == "data") {
        data_acc = val;
        if (val.find(']') != std::string::npos) {
          parse_data(data_acc);
          finish_mat();
        } else {
          in_data = true;
        }
      }
    }
  }
  finish_mat();
}

double Yaml::num(const std::string& key, double dflt) const {
  auto it = scalars_.find(key);
  if (it == scalars_.end()) return dflt;
  char* end = nullptr;
  const double v = std::strtod(it->second.c_str(), &end);
  return end == it->second.c_str() ? dflt : v;
}
std::string Yaml::str(const std::string& key, const std::string& dflt) const {
  auto it = scalars_.find(key);
  return it == scalars_.end() ? dflt : it->second;
}
std::vector<double> Yaml::mat(const std::string& key, int* rows, int* cols) const {
  auto it = mats_.find(key);
  if (it == mats_.end()) return {};
  if (rows) *rows = it->second.rows;
  if (cols) *cols = it->second.cols;
  return it->second.data;
}

// ---- .xyz loader (objloader.cpp:225-264) ----------------------------------------------
bool loadXYZ(const char* path, const char* offset_path, std::vector<float>& xyzi) {
  double ox = 0, oy = 0, oz = 0;
  {
    std::ifstream off(offset_path);
    if (!off.is_open()) return false;
    off >> ox >> oy >> oz;
  }
  std::ifstream in(path);
  if (!in.is_open()) return false;
  xyzi.clear();
  double x = 0, y = 0, z = 0;
  float r = 0, g = 0, b = 0;
  // same loop shape as the reference: the stream is only known to be exhausted AFTER a
  // failed read, so a trailing newline pushes the last point twice
  while (!in.eof()) {
    in >> x >> y >> z >> r >> g >> b;
    xyzi.push_back((float)(x - ox));
    xyzi.push_back((float)(y - oy));
    xyzi.push_back((float)(z - oz));
    xyzi.push_back((1.0f / 256.0f) * r);
    if (in.fail() && !in.eof()) return false;  // malformed line: the reference would spin forever
  }
  return !xyzi.empty();
}

// ---- .obj / .bmp loaders (objloader.cpp:140-223, texture.cpp:31-107) ---------------------
bool loadOBJ(const char* path, std::vector<float>& xyz, std::vector<float>& uv) {
  std::ifstream in(path);
  if (!in.is_open()) return false;
  std::vector<float> v, vt;
  std::vector<unsigned> vi, ti;
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    std::string head;
    if (!(ss >> head)) continue;
    if (head == "v") {
      float x, y, z;
      if (!(ss >> x >> y >> z)) return false;
      v.insert(v.end(), {x, y, z});
    } else if (head == "vt") {
      float a, b;
      if (!(ss >> a >> b)) return false;
      vt.insert(vt.end(), {a, b});
    } else if (head == "f") {
      for (int k = 0; k < 3; k++) {
        std::string tok;
        if (!(ss >> tok)) return false;
        unsigned a = 0, b = 0;
        if (std::sscanf(tok.c_str(), "%u/%u", &a, &b) != 2 || tok.find('/') != tok.rfind('/')) return false;
        vi.push_back(a);
        ti.push_back(b);
      }
      std::string extra;
      if (ss >> extra) return false;  // quads etc.: "can't be read by this simple parser"
    }  // anything else: comment / unsupported statement, skipped like the reference
  }
  xyz.clear();
  uv.clear();
  for (size_t i = 0; i < vi.size(); i++) {
    if (vi[i] == 0 || ti[i] == 0 || 3 * (size_t)vi[i] > v.size() || 2 * (size_t)ti[i] > vt.size()) return false;
    xyz.insert(xyz.end(), v.begin() + 3 * (vi[i] - 1), v.begin() + 3 * (vi[i] - 1) + 3);
    uv.insert(uv.end(), vt.begin() + 2 * (ti[i] - 1), vt.begin() + 2 * (ti[i] - 1) + 2);
  }
  return !xyz.empty();
}

bool loadBMP24(const char* path, int& width, int& height, std::vector<unsigned char>& rgb) {
  std::ifstream f(path, std::ios::binary);
  unsigned char h[54];
  if (!f.read(reinterpret_cast<char*>(h), 54) || h[0] != 'B' || h[1] != 'M') return false;
  auto u32 = [&](int o) { return (unsigned)h[o] | ((unsigned)h[o + 1] << 8) | ((unsigned)h[o + 2] << 16) | ((unsigned)h[o + 3] << 24); };
  if (u32(0x1E) != 0 || (u32(0x1C) & 0xFFFF) != 24) return false;  // texture.cpp:58-59
  unsigned dataPos = u32(0x0A), imageSize = u32(0x22);
  width = (int)u32(0x12);
  height = (int)u32(0x16);
  if (width <= 0 || height <= 0) return false;
  if (imageSize == 0) imageSize = (unsigned)width * height * 3;  // texture.cpp:68
  if (dataPos == 0) dataPos = 54;
  rgb.assign(imageSize, 0);
  f.seekg(dataPos);
  f.read(reinterpret_cast<char*>(rgb.data()), imageSize);
  return (size_t)f.gcount() >= (size_t)width * height * 3;
}

namespace {
struct SetupCamMemo {
  float pos[3], dir[3], up[3], Twc[16];
  bool valid = false;
};
thread_local SetupCamMemo g_memo;
}  // namespace

void remember_setup_cam(const float pos[3], const float dir[3], const float up[3], const float Twc[16]) {
  std::memcpy(g_memo.pos, pos, sizeof g_memo.pos);
  std::memcpy(g_memo.dir, dir, sizeof g_memo.dir);
  std::memcpy(g_memo.up, up, sizeof g_memo.up);
  std::memcpy(g_memo.Twc, Twc, sizeof g_memo.Twc);
  g_memo.valid = true;
}

bool recall_setup_cam(const float pos[3], const float dir[3], const float up[3], float Twc[16]) {
  if (!g_memo.valid || std::memcmp(g_memo.pos, pos, sizeof g_memo.pos) || std::memcmp(g_memo.dir, dir, sizeof g_memo.dir) ||
      std::memcmp(g_memo.up, up, sizeof g_memo.up))
    return false;
  std::memcpy(Twc, g_memo.Twc, sizeof g_memo.Twc);
  return true;
}

// OBJ + BMP -> the arguments of nmi_set_mesh_textured: un-indexed corners like the reference's VBOs
// (objloader.cpp:206-220: one position and one UV per face corner), the BMP payload untouched
// (texture.cpp:90 hands it to glTexImage2D(GL_RGB) as it lies in the file: B,G,R, bottom row first;
// tightly packed rows, GL_UNPACK_ALIGNMENT 1, rendering.hpp:313).  verts[4i+3] additionally carries
// the luma of the corner's nearest texel -- only used by the flat-shaded fallback nmi_set_mesh.
bool meshFromObjBmp(const char* obj_path, const char* bmp_path, std::vector<float>& verts,
                    std::vector<uint32_t>& tris, std::vector<float>* corner_uv, std::vector<unsigned char>* texture,
                    int* tex_w, int* tex_h) {
  std::vector<float> xyz, uv;
  std::vector<unsigned char> tex;
  int tw = 0, th = 0;
  if (!loadOBJ(obj_path, xyz, uv) || !loadBMP24(bmp_path, tw, th, tex)) return false;
  const size_t nv = xyz.size() / 3;
  verts.resize(4 * nv);
  tris.resize(nv);
  for (size_t i = 0; i < nv; i++) {
    const float u = uv[2 * i] - std::floor(uv[2 * i]), v = uv[2 * i + 1] - std::floor(uv[2 * i + 1]);
    int tx = (int)std::floor(u * tw), ty = (int)std::floor(v * th);
    tx = tx < 0 ? 0 : (tx >= tw ? tw - 1 : tx);
    ty = ty < 0 ? 0 : (ty >= th ? th - 1 : ty);
    const unsigned char* t = &tex[3 * ((size_t)ty * tw + tx)];
    const float r = t[0] / 255.0f, g = t[1] / 255.0f, b = t[2] / 255.0f;  // file B,G,R read as R,G,B
    verts[4 * i] = xyz[3 * i];
    verts[4 * i + 1] = xyz[3 * i + 1];
    verts[4 * i + 2] = xyz[3 * i + 2];
    verts[4 * i + 3] = 0.299f * r + 0.587f * g + 0.114f * b;
    tris[i] = (uint32_t)i;
  }
  if (corner_uv) *corner_uv = uv;  // 2 floats per corner, 3 corners per triangle, in face order
  if (texture) {
    tex.resize((size_t)tw * th * 3);
    *texture = tex;
  }
  if (tex_w) *tex_w = tw;
  if (tex_h) *tex_h = th;
  return nv >= 3 && nv % 3 == 0;
}

bool meshFromObjBmp(const char* obj_path, const char* bmp_path, std::vector<float>& verts,
                    std::vector<uint32_t>& tris) {
  return meshFromObjBmp(obj_path, bmp_path, verts, tris, nullptr, nullptr, nullptr, nullptr);
}

}  // namespace nmi_compat

// ---- CUDAF::NMIWithCuda_noMask (kernel.cuh:37, kernel.cu:49-114) -------------------------
namespace CUDAF {
void NMIWithCuda_noMask(cv::cuda::PtrStep<unsigned char>* d_Warped, int NMI_mode, int MatchingMode,
                        int width, int height, float* NMI, unsigned int syntGL) {
  (void)NMI_mode;      // accepted and never read, like the reference
  (void)MatchingMode;
  nmi_compat::check(nmi_eval_pair(nmi_compat::context(), reinterpret_cast<const void*>(d_Warped), syntGL,
                                  width, height, &nmi_compat::flags(), NMI),
                    "CUDAF::NMIWithCuda_noMask");
}
}  // namespace CUDAF

// ---- NmiObjects (localization.cpp) ---------------------------------------------------------
static std::string make_results_dir(std::string* log_path) {
  namespace fs = std::filesystem;
  const std::time_t t = std::time(nullptr);
  std::tm tmv{};
  localtime_r(&t, &tmv);
  std::ostringstream name;
  // allProperties.hpp's output location, overridable at run time ($NMI_OUTPUT_LOC)
  const char* base = std::getenv("NMI_OUTPUT_LOC");
  name << (base && *base ? base : nmi_prop_OUTPUT_LOC) << "/" << std::put_time(&tmv, "%d-%m-%Y_%Hh%Mm%Ss");
  fs::path dir(name.str());
  std::error_code ec;
  // localization.cpp:94-100: a relative location is taken under the working directory
  fs::path rel = fs::current_path(ec) / fs::path(name.str()).relative_path();
  if (!fs::create_directories(dir, ec) && !fs::exists(dir)) {
    dir = rel;
    if (!fs::create_directories(dir, ec) && !fs::exists(dir)) {
      dir = fs::temp_directory_path(ec) / fs::path(name.str()).relative_path();
      fs::create_directories(dir, ec);
    }
  }
  *log_path = (dir / "_log.txt").string();
  return dir.string();
}

NmiObjects::NmiObjects(const std::string& strSettingsFile, int render_mode) : rating(nullptr), N(0) {
  std::stringstream ss_log;
  resultsPath = make_results_dir(&logPath);
  ss_log << resultsPath << std::endl << "logPath: " << logPath << std::endl;
  helperFunctions::log(ss_log, logPath);

  nmi_compat::Yaml y(strSettingsFile);
  if (!y.ok()) {
    std::fprintf(stderr, "NmiObjects: cannot open settings file %s\n", strSettingsFile.c_str());
    std::exit(EXIT_FAILURE);
  }
  const glm::vec3 zero(0.0f, 0.0f, 0.0f);
  const int W = (int)y.num("Camera.Width"), H = (int)y.num("Camera.Height");
  const int nsx = (int)y.num("NMI.SynthNumX", 1), nsy = (int)y.num("NMI.SynthNumY", 1),
            nsz = (int)y.num("NMI.SynthNumZ", 1);
  const int nwx = (int)y.num("NMI.WarpNumX", 1), nwy = (int)y.num("NMI.WarpNumY", 1),
            nwz = (int)y.num("NMI.WarpNumZ", 1);
  const float sx = (float)y.num("NMI.SynthStepX"), sy = (float)y.num("NMI.SynthStepY"),
              sz = (float)y.num("NMI.SynthStepZ");
  const float rx = (float)y.num("NMI.WarpStepX"), ry = (float)y.num("NMI.WarpStepY"),
              rz = (float)y.num("NMI.WarpStepZ");
  const double fx = y.num("Camera.fx"), fy = y.num("Camera.fy"), cx = y.num("Camera.cx"),
               cy = y.num("Camera.cy");
  threshold_ = (float)y.num("NMI.Treshold");

  // localization.cpp:133-158 (1280x720 is the hidden GL window of the reference)
  // nmi_prop_RENDER is the CALLER's compile-time choice (allProperties.hpp:42, handed over by the inline
  // constructor in localization.hpp); Rendering<1> and Rendering<4> differ only in which loader their
  // constructor runs (rendering.hpp:172-189), their layout is the same
#define NMI_NEW_RENDERER(MODE)                                                                                    \
  reinterpret_cast<Rendering<nmi_prop_RENDER>*>(new Rendering<MODE>(                                              \
      (float)y.num("NMI.Render.PointSize", 3.0), 1280, 720, W, H, nsx, nsy, nsz, sx, sy, sz, zero, zero, zero,     \
      y.num("NMI.Render.NearPlane", 5.0), y.num("NMI.Render.FarPlane", 30.0), fx, fy, cx, cy,                      \
      y.str("NMI.Render.Object"), y.str("NMI.Render.Texture"), y.str("NMI.Render.Cloud"),                          \
      y.str("NMI.Render.Offset"), logPath))
  myRenderer = render_mode == RENDER_TEXTURE ? NMI_NEW_RENDERER(RENDER_TEXTURE) : NMI_NEW_RENDERER(RENDER_POINT_CLOUD);
#undef NMI_NEW_RENDERER

  cv::Mat K = cv::Mat::eye(3, 3, CV_64F);  // localization.cpp:165-169
  K.at<double>(0, 0) = fx;
  K.at<double>(1, 1) = fy;
  K.at<double>(0, 2) = cx;
  K.at<double>(1, 2) = cy;
  myImage = new Image(nwz, nwy, nwx, rz, ry, rx, W, H, K);  // localization.cpp:172-181

  NmiKernel = new NmiSearchKernel(nsx, nsy, nsz, nwx, nwy, nwz, sx, sy, sz, rx, ry, rz);
  LastNmiKernel = new NmiSearchKernel(nsx, nsy, nsz, nwx, nwy, nwz, sx, sy, sz, rx, ry, rz);
  InitialNmiKernel = new NmiSearchKernel(nsx, nsy, nsz, nwx, nwy, nwz, sx, sy, sz, rx, ry, rz);
  allocRating(NmiKernel->grid());
}

NmiObjects::~NmiObjects() {
  deleteRating();
  delete myRenderer;
  delete myImage;
  delete NmiKernel;
  delete LastNmiKernel;
  delete InitialNmiKernel;
}

void NmiObjects::allocRating(const nmi_grid& g) {
  for (int k = 0; k < 3; k++) {
    rw_[k] = g.nW[k];
    rs_[k] = g.nS[k];
  }
  rating = new float*****[rw_[2]];
  for (int wz = 0; wz < rw_[2]; wz++) {
    rating[wz] = new float****[rw_[1]];
    for (int wy = 0; wy < rw_[1]; wy++) {
      rating[wz][wy] = new float***[rw_[0]];
      for (int wx = 0; wx < rw_[0]; wx++) {
        rating[wz][wy][wx] = new float**[rs_[2]];
        for (int z = 0; z < rs_[2]; z++) {
          rating[wz][wy][wx][z] = new float*[rs_[1]];
          for (int yy = 0; yy < rs_[1]; yy++) rating[wz][wy][wx][z][yy] = new float[rs_[0]]();
        }
      }
    }
  }
}

void NmiObjects::deleteRating() {
  if (!rating) return;
  for (int wz = 0; wz < rw_[2]; wz++) {
    for (int wy = 0; wy < rw_[1]; wy++) {
      for (int wx = 0; wx < rw_[0]; wx++) {
        for (int z = 0; z < rs_[2]; z++) {
          for (int yy = 0; yy < rs_[1]; yy++) delete[] rating[wz][wy][wx][z][yy];
          delete[] rating[wz][wy][wx][z];
        }
        delete[] rating[wz][wy][wx];
      }
      delete[] rating[wz][wy];
    }
    delete[] rating[wz];
  }
  delete[] rating;
  rating = nullptr;
}

void NmiObjects::resizeKernel(int nsx, int nsy, int nsz, int nwx, int nwy, int nwz, float, float, float,
                              float, float, float) {
  deleteRating();
  nmi_grid g{};
  g.nS[0] = nsx; g.nS[1] = nsy; g.nS[2] = nsz;
  g.nW[0] = nwx; g.nW[1] = nwy; g.nW[2] = nwz;
  allocRating(g);
}

void NmiObjects::setImageVars(int nwx, int nwy, int nwz, float rx, float ry, float rz) {
  myImage->setStepX(rx); myImage->setStepY(ry); myImage->setStepZ(rz);
  myImage->setNumWarpX(nwx); myImage->setNumWarpY(nwy); myImage->setNumWarpZ(nwz);
}
void NmiObjects::setRendererVars(int nsx, int nsy, int nsz, float sx, float sy, float sz) {
  myRenderer->setStep_x(sx); myRenderer->setStep_y(sy); myRenderer->setStep_z(sz);
  myRenderer->setSynthetic_count_x(nsx); myRenderer->setSynthetic_count_y(nsy);
  myRenderer->setSynthetic_count_z(nsz);
}

// localization.cpp:390-400
void NmiObjects::setNmiObjectsKernel(int nsx, int nsy, int nsz, int nwx, int nwy, int nwz, float sx, float sy,
                                     float sz, float rx, float ry, float rz) {
  resizeKernel(nsx, nsy, nsz, nwx, nwy, nwz, sx, sy, sz, rx, ry, rz);
  myImage->resizeKernel(nwx, nwy, nwz, rx, ry, rz);
  myRenderer->resizeKernel(nsx, nsy, nsz, sx, sy, sz);
}
void NmiObjects::setNmiObjectsKernel(NmiSearchKernel* k) {
  setNmiObjectsKernel(k->numSynthX, k->numSynthY, k->numSynthZ, k->numWarpX, k->numWarpY, k->numWarpZ,
                      k->stepX, k->stepY, k->stepZ, k->stepRadX, k->stepRadY, k->stepRadZ);
}
// localization.cpp:410-420
void NmiObjects::NMIobjectsReInitialization() {
  LastNmiKernel->setTo(NmiKernel);
  NmiKernel->resizeKernel();
  setNmiObjectsKernel(NmiKernel);
}

static void mat_to_twc(const cv::Mat& m, float T[16]) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) T[4 * i + j] = m.at<float>(i, j);
}
static cv::Mat twc_to_mat(const float T[16]) {
  cv::Mat m(4, 4, CV_32F);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) m.at<float>(i, j) = T[4 * i + j];
  return m;
}

cv::Mat NmiObjects::searchGrid(cv::Mat Twc, cv::Mat gray) {
  myImage->loadOriginal(gray);
  const nmi_grid g = NmiKernel->grid();
  if (g.nS[0] != rs_[0] || g.nS[1] != rs_[1] || g.nS[2] != rs_[2] || g.nW[0] != rw_[0] ||
      g.nW[1] != rw_[1] || g.nW[2] != rw_[2])
    setNmiObjectsKernel(NmiKernel);
  float T[16], Tn[16];
  mat_to_twc(Twc, T);
  const size_t nP = (size_t)g.nS[0] * g.nS[1] * g.nS[2] * g.nW[0] * g.nW[1] * g.nW[2];
  std::vector<float> scores(nP);
  nmi_result r{};
  nmi_compat::check(nmi_search(nmi_compat::context(), T, &g, &nmi_compat::flags(), &r, scores.data()),
                    "NmiObjects::searchGrid");
  size_t l = 0;
  for (int wz = 0; wz < g.nW[2]; wz++)
    for (int wy = 0; wy < g.nW[1]; wy++)
      for (int wx = 0; wx < g.nW[0]; wx++)
        for (int z = 0; z < g.nS[2]; z++)
          for (int yy = 0; yy < g.nS[1]; yy++)
            for (int x = 0; x < g.nS[0]; x++) rating[wz][wy][wx][z][yy][x] = scores[l++];
  NmiKernel->setBest(r.best_s[0], r.best_s[1], r.best_s[2], r.best_w[0], r.best_w[1], r.best_w[2],
                     r.best_score);
  nmi_apply_winner(T, &g, r.best_s, r.best_w, Tn);
  return twc_to_mat(Tn);
}

nmi_reloc_result NmiObjects::relocalize(cv::Mat Twc, cv::Mat gray, const nmi_reloc_params& params,
                                        bool not_initialized) {
  return relocalizeSharded(Twc, gray, params, 0, 1, nullptr, nullptr, nullptr, not_initialized);
}

nmi_reloc_result NmiObjects::relocalizeSharded(cv::Mat Twc, cv::Mat gray, const nmi_reloc_params& params,
                                               int rank, int world, void* key_dev, nmi_exchange_fn exchange,
                                               void* user, bool not_initialized) {
  myImage->loadOriginal(gray);
  NmiKernel->reset();      // Tracking.cc:1997
  LastNmiKernel->reset();  // Tracking.cc:1998
  nmi_grid init = InitialNmiKernel->grid(), start{};
  nmi_grid_from_motion(&init, params.distance_since_last, params.rotation_since_last,
                       not_initialized ? 1 : 0, &start);
  float T[16];
  mat_to_twc(Twc, T);
  nmi_reloc_result out{};
  if (exchange == nullptr)
    nmi_compat::check(nmi_relocalize(nmi_compat::context(), T, &start, &nmi_compat::flags(), &params, &out),
                      "NmiObjects::relocalize");
  else
    nmi_compat::check(nmi_relocalize_sharded(nmi_compat::context(), T, &start, &nmi_compat::flags(), &params,
                                             rank, world, key_dev, exchange, user, &out),
                      "NmiObjects::relocalizeSharded");
  // the lines Tracking.cc:2103-2106 appends to _log.txt after every search
  if (!logPath.empty()) {
    for (int l = 0; l < out.n_levels; l++) {
      NmiSearchKernel cur, last;
      cur.setGrid(out.levels[l].grid);
      cur.setBest(out.levels[l].best_s[0], out.levels[l].best_s[1], out.levels[l].best_s[2],
                  out.levels[l].best_w[0], out.levels[l].best_w[1], out.levels[l].best_w[2], out.levels[l].nmi);
      if (l == 0) {  // LastNmiKernel->reset() (Tracking.cc:1998): no winner yet, NMI 0
        last.setGrid(out.levels[0].grid);
        last.reset();
      } else {
        last.setGrid(out.levels[l - 1].grid);
        last.setBest(out.levels[l - 1].best_s[0], out.levels[l - 1].best_s[1], out.levels[l - 1].best_s[2],
                     out.levels[l - 1].best_w[0], out.levels[l - 1].best_w[1], out.levels[l - 1].best_w[2],
                     out.levels[l - 1].nmi);
      }
      std::stringstream ss_log;
      ss_log << "NmiKernel:\t" << cur;
      ss_log << "\nLastNmiKernel:\t" << last;
      ss_log << "\nKernel rate:\t" << (out.levels[l].nmi / out.levels[l].last_nmi) << "\n";
      helperFunctions::log(ss_log, logPath);
    }
  }
  // leave the objects in the state the reference's loop would (kernel + winner + NMI)
  NmiKernel->setGrid(out.final_grid);
  NmiKernel->setBest(out.best_s[0], out.best_s[1], out.best_s[2], out.best_w[0], out.best_w[1],
                     out.best_w[2], out.nmi);
  LastNmiKernel->NMI = out.last_nmi;
  setNmiObjectsKernel(NmiKernel);
  return out;
}
