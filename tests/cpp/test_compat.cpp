// test_compat.cpp -- exercises the reference-named C++ drop-ins (include/compat/) the way
// src/Tracking.cc uses them.  Built and driven by tests/test_compat_cpp.py.
//   test_compat host <yaml> <xyz> <offset>          host-only checks (no GPU)
//   test_compat gpu  <yaml> <frame.raw> <twc.txt>    the reference's loop vs the batched call
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>

#include "compat/helperFunctions.hpp"
#include "compat/ioData.hpp"
#include "compat/kernel.cuh"
#include "compat/localization.hpp"
#include "compat/nmi_outputs.hpp"

#define EXPECT(c)                                                     \
  do {                                                                \
    if (!(c)) {                                                       \
      std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c);         \
      return 1;                                                       \
    }                                                                 \
  } while (0)

// SaveFullTrajectory / overlay writer (System.cc:514-599, ioData.cpp:262-285)
static int output_checks(const char* dir) {
  using namespace nmi_compat;
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  float q[4];
  rotation_to_quaternion(I, q);
  EXPECT(q[0] == 0 && q[1] == 0 && q[2] == 0 && q[3] == 1);
  const double Rz180[9] = {-1, 0, 0, 0, -1, 0, 0, 0, 1};  // trace < 0: pivot branch
  rotation_to_quaternion(Rz180, q);
  EXPECT(q[0] == 0 && q[1] == 0 && q[2] == 1 && q[3] == 0);
  const double c = std::cos(0.5), sn = std::sin(0.5);
  const double Rx[9] = {1, 0, 0, 0, c, -sn, 0, sn, c};
  rotation_to_quaternion(Rx, q);
  EXPECT(std::fabs(q[0] - (float)std::sin(0.25)) < 1e-7f && std::fabs(q[3] - (float)std::cos(0.25)) < 1e-7f);

  TrajectoryRecorder tr;
  float T[16] = {1, 0, 0, 1.5f, 0, 1, 0, -2.25f, 0, 0, 1, 3, 0, 0, 0, 1};
  tr.add(7, 12.5, T, true);
  nmi_reloc_result r{};
  for (int i = 0; i < 16; i++) r.Twc[i] = T[i];
  r.Twc[3] = 1.75f;
  r.relocalized = 1;
  r.n_prev = 1;
  for (int i = 0; i < 16; i++) r.prev_Twc[0][i] = T[i];
  tr.add(8, 12.6, r, true);
  r.relocalized = 0;
  r.failed = 1;
  r.n_prev = 0;
  tr.add(9, 12.7, r, true);
  tr.add(10, 12.8, T, false);        // ordinary frame: no tag
  tr.add(11, 12.9, T, false, true);  // lost: skipped
  EXPECT(tr.size() == 5);
  EXPECT(tr.SaveFullTrajectory(std::string(dir) + "/FrameTrajectory"));

  const int W = 5, H = 3;  // 15-byte rows -> 1 pad byte
  std::vector<uint8_t> img(W * H), syn(W * H);
  for (int i = 0; i < W * H; i++) {
    img[i] = (uint8_t)(10 + i);
    syn[i] = (uint8_t)(100 + i);
  }
  EXPECT(saveOverlayBMP((std::string(dir) + "/overlay.bmp").c_str(), img.data(), syn.data(), W, H));
  nmi_grid g{{3, 3, 3}, {3, 3, 3}, {0.2f, 0.2f, 0.5f}, {0.02f, 0.02f, 0.05f}};
  const int32_t bs[3] = {0, 1, 2}, bw[3] = {2, 1, 0};
  std::printf("NAME %s\n", overlay_name("res", 12, g, 0.25f, bs, bw).c_str());
  std::printf("JNAME %s\n", overlay_name("res", 12, g, 0.25f, bs, bw, ".jpg").c_str());
  {  // the same overlay as a JPEG (what cv::imwrite gives the reference's .jpg names): 83 x 37, not a multiple of 8
    const int JW = 83, JH = 37;
    std::vector<uint8_t> ji(JW * JH), js(JW * JH);
    for (int y = 0; y < JH; y++)
      for (int x = 0; x < JW; x++) {
        ji[y * JW + x] = (uint8_t)(128 + 100 * std::sin(x / 9.0) * std::cos(y / 7.0));
        js[y * JW + x] = (uint8_t)((x * 3 + y * 5) % 256 > 200 ? 255 : 40 + (x + 2 * y) % 150);
      }
    EXPECT(saveOverlayJPG((std::string(dir) + "/overlay.jpg").c_str(), ji.data(), js.data(), JW, JH));
    EXPECT(saveOverlayBMP((std::string(dir) + "/overlay_ref.bmp").c_str(), ji.data(), js.data(), JW, JH));
  }
  std::printf("OUTPUTS OK\n");
  return 0;
}

static int loader_checks(const char* obj, const char* bmp) {
  // loadOBJ (objloader.cpp:140-223): 2 triangles, un-indexed output in face order
  std::vector<float> xyz, uv;
  EXPECT(nmi_compat::loadOBJ(obj, xyz, uv));
  EXPECT(xyz.size() == 18 && uv.size() == 12);
  EXPECT(xyz[0] == 0.0f && xyz[3] == 1.0f && xyz[7] == 1.0f);     // v1, v2, v3 of face 1
  EXPECT(uv[0] == 0.25f && uv[1] == 0.25f && uv[2] == 0.75f);
  // loadBMP24 (texture.cpp:31-107): 2x2, bytes as stored (B,G,R read as R,G,B), rows bottom-up
  int w = 0, h = 0;
  std::vector<unsigned char> rgb;
  EXPECT(nmi_compat::loadBMP24(bmp, w, h, rgb));
  EXPECT(w == 2 && h == 2 && rgb.size() >= 12);
  EXPECT(rgb[0] == 10 && rgb[1] == 20 && rgb[2] == 30);  // bottom-left texel as written in the file
  // meshFromObjBmp: luma of the nearest texel on the swapped channels
  std::vector<float> verts;
  std::vector<uint32_t> tris;
  EXPECT(nmi_compat::meshFromObjBmp(obj, bmp, verts, tris));
  EXPECT(verts.size() == 24 && tris.size() == 6 && tris[5] == 5);
  const float want = 0.299f * (10 / 255.0f) + 0.587f * (20 / 255.0f) + 0.114f * (30 / 255.0f);
  EXPECT(std::fabs(verts[3] - want) < 1e-6f);             // uv (0.25, 0.25) -> texel (0, 0)
  const float want2 = 0.299f * (40 / 255.0f) + 0.587f * (50 / 255.0f) + 0.114f * (60 / 255.0f);
  EXPECT(std::fabs(verts[7] - want2) < 1e-6f);            // uv (0.75, 0.25) -> texel (1, 0)
  std::printf("LOADERS OK\n");
  return 0;
}

static int host_checks(const char* yaml, const char* xyz, const char* off) {
  // NmiSearchKernel (nmiSearchKernel.cpp)
  NmiSearchKernel k(3, 3, 3, 3, 3, 3, 0.2f, 0.2f, 0.5f, 0.02f, 0.02f, 0.05f);
  EXPECT(k.getNumSynthX() == 3 && k.getBestWarpZ() == -1 && k.getNmi() == 0);
  k.setBest(1, 1, 1, 1, 1, 1, 0.5f);
  EXPECT(k.isMiddle());
  k.setBest(0, 1, 1, 1, 1, 2, 0.5f);
  EXPECT(!k.isMiddle());
  k.resizeKernel();  // sX and wZ on the periphery keep their step, the rest halve
  EXPECT(k.stepX == 0.2f && k.stepY == 0.1f && k.stepZ == 0.25f);
  EXPECT(k.stepRadX == 0.01f && k.stepRadY == 0.01f && k.stepRadZ == 0.05f);
  NmiSearchKernel tiny(3, 3, 3, 3, 3, 3, 0.009f, 0.2f, 0.5f, 0.0019f, 0.02f, 0.05f);
  tiny.setBest(1, 1, 1, 1, 1, 1, 0.1f);
  tiny.resizeKernel();
  EXPECT(tiny.numSynthX == 1 && tiny.numWarpX == 1 && tiny.numSynthY == 3);
  NmiSearchKernel copy;
  copy.setTo(&k);
  EXPECT(copy.stepY == k.stepY && copy.bestWarpZ == 2 && copy.NMI == 0.5f);
  copy.reset();
  EXPECT(copy.numSynthX == -1 && copy.bestSynthX == -1 && copy.NMI == 0);

  // find_max_elements (helperFunctions.cpp:50-103) on a 2x1x1 x 1x1x2 rating
  NmiSearchKernel g(2, 1, 1, 1, 1, 2, 0, 0, 0, 0, 0, 0);
  float v[4] = {0.25f, 0.75f, 0.75f, 0.5f};
  float* r0[1] = {&v[0]};
  float* r1[1] = {&v[2]};
  float** z0[1] = {r0};
  float** z1[1] = {r1};
  float*** x0[1] = {z0};
  float*** x1[1] = {z1};
  float**** y0[1] = {x0};
  float**** y1[1] = {x1};
  float***** wz[2] = {y0, y1};
  auto m = helperFunctions::find_max_elements(wz, g);
  EXPECT(m.size() == 2 && m[0].getBestSynthX() == 1 && m[0].getBestWarpZ() == 0 && m[0].getNmi() == 0.75f);
  EXPECT(m[1].getBestSynthX() == 0 && m[1].getBestWarpZ() == 1);
  v[0] = v[1] = v[2] = v[3] = -1.0f;  // nothing reaches 0: the reference's vector is empty
  EXPECT(helperFunctions::find_max_elements(wz, g).empty());

  // setupCam (ioData.cpp:177-197)
  cv::Mat T = cv::Mat::eye(4, 4, CV_32F), K = cv::Mat::eye(3, 3, CV_64F);
  T.at<float>(0, 3) = 1; T.at<float>(1, 3) = 2; T.at<float>(2, 3) = 3;
  CameraSettings cs = setupCam(T, K);
  EXPECT(cs.getPosition().x == 1 && cs.getDirection().z == 4 && cs.getUp().y == 1 && cs.getUp().x == 0);

  // YAML reader on the reference-style settings file
  nmi_compat::Yaml y(yaml);
  EXPECT(y.ok());
  EXPECT(y.num("NMI.SynthNumX") == 2 && y.num("NMI.WarpNumY") == 2);
  EXPECT(std::fabs(y.num("NMI.SynthStepZ") - 0.5) < 1e-12 && std::fabs(y.num("Camera.fx") - 95.0) < 1e-12);
  EXPECT(y.str("NMI.Render.Cloud") == xyz);
  int rows = 0, cols = 0;
  auto init1 = y.mat("NMI.Init1", &rows, &cols);
  EXPECT(rows == 4 && cols == 4 && init1.size() == 16 && init1[15] == 1.0);

  // loadXYZ (objloader.cpp:225-264): offset subtraction, /256, trailing-newline duplicate
  std::vector<float> pts;
  EXPECT(nmi_compat::loadXYZ(xyz, off, pts));
  std::ifstream f(xyz);
  size_t lines = 0;
  std::string line;
  while (std::getline(f, line))
    if (!line.empty()) lines++;
  EXPECT(pts.size() / 4 == lines + 1);  // last point pushed twice
  const size_t n = pts.size() / 4;
  EXPECT(pts[4 * (n - 1)] == pts[4 * (n - 2)] && pts[4 * (n - 1) + 3] == pts[4 * (n - 2) + 3]);
  std::printf("HOST OK %zu points\n", n);
  return 0;
}

static int gpu_checks(const char* yaml, const char* frame_raw, const char* twc_txt) {
  NmiObjects objs(yaml);
  Image* img = objs.myImage;
  Rendering<nmi_prop_RENDER>* ren = objs.myRenderer;
  const int W = ren->getImageWidth(), H = ren->getImageHeight();
  cv::Mat gray(H, W, CV_8U);
  {
    std::ifstream f(frame_raw, std::ios::binary);
    f.read(reinterpret_cast<char*>(gray.data), (std::streamsize)W * H);
    if (!f) { std::printf("FAIL cannot read frame\n"); return 1; }
  }
  cv::Mat Twc(4, 4, CV_32F);
  {
    std::ifstream f(twc_txt);
    for (int i = 0; i < 16; i++) f >> Twc.at<float>(i);
  }
  // ---- the reference's loop, verbatim in structure (src/Tracking.cc:1871-1905) ----
  img->loadOriginal(gray.clone());
  img->calculateWarping();
  cv::Mat K = img->getK();
  CameraSettings settings = setupCam(Twc, K);
  ren->setCamera(settings.getPosition(), settings.getDirection(), settings.getUp());
  for (int sX = 0; sX < ren->getNumSynthX(); sX++)
    for (int sY = 0; sY < ren->getNumSynthY(); sY++)
      for (int sZ = 0; sZ < ren->getNumSynthZ(); sZ++) {
        ren->renderToTextureOnGPU(ren->calculateTranslation(sX, sY, sZ));
        for (int wX = 0; wX < img->getNumWarpX(); wX++)
          for (int wY = 0; wY < img->getNumWarpY(); wY++)
            for (int wZ = 0; wZ < img->getNumWarpZ(); wZ++)
              CUDAF::NMIWithCuda_noMask(
                  (cv::cuda::PtrStep<unsigned char>*)img->getImageGPU(wZ, wY, wX).data, SUC, MATCHING_NMI,
                  ren->getImageWidth(), ren->getImageHeight(), &(objs.rating[wZ][wY][wX][sZ][sY][sX]),
                  ren->getrenderedTexture());
      }
  std::vector<NmiSearchKernel> ext = helperFunctions::find_max_elements(objs.rating, *objs.NmiKernel);
  if (ext.empty()) { std::printf("FAIL no extreme element\n"); return 1; }
  std::printf("LOOP");
  for (int wz = 0; wz < img->getNumWarpZ(); wz++)
    for (int wy = 0; wy < img->getNumWarpY(); wy++)
      for (int wx = 0; wx < img->getNumWarpX(); wx++)
        for (int sz = 0; sz < ren->getNumSynthZ(); sz++)
          for (int sy = 0; sy < ren->getNumSynthY(); sy++)
            for (int sx = 0; sx < ren->getNumSynthX(); sx++) std::printf(" %.9g", objs.rating[wz][wy][wx][sz][sy][sx]);
  std::printf("\nLOOPBEST %d %d %d %d %d %d %.9g\n", ext[0].getBestSynthX(), ext[0].getBestSynthY(),
              ext[0].getBestSynthZ(), ext[0].getBestWarpX(), ext[0].getBestWarpY(), ext[0].getBestWarpZ(),
              ext[0].getNmi());
  // ---- the batched drop-in for the same loop ----
  cv::Mat newTwc = objs.searchGrid(Twc, gray);
  std::printf("BATCH");
  for (int wz = 0; wz < img->getNumWarpZ(); wz++)
    for (int wy = 0; wy < img->getNumWarpY(); wy++)
      for (int wx = 0; wx < img->getNumWarpX(); wx++)
        for (int sz = 0; sz < ren->getNumSynthZ(); sz++)
          for (int sy = 0; sy < ren->getNumSynthY(); sy++)
            for (int sx = 0; sx < ren->getNumSynthX(); sx++) std::printf(" %.9g", objs.rating[wz][wy][wx][sz][sy][sx]);
  NmiSearchKernel* nk = objs.NmiKernel;
  std::printf("\nBATCHBEST %d %d %d %d %d %d %.9g\n", nk->bestSynthX, nk->bestSynthY, nk->bestSynthZ,
              nk->bestWarpX, nk->bestWarpY, nk->bestWarpZ, nk->NMI);
  std::printf("NEWTWC");
  for (int i = 0; i < 16; i++) std::printf(" %.9g", newTwc.at<float>(i));
  std::printf("\n");
  std::fflush(stdout);  // relocalize() echoes its log lines through std::cout
  // ---- the multi-level driver ----
  nmi_reloc_params prm{};
  prm.threshold = objs.threshold();
  prm.max_iterations = nmi_prop_MAX_ITERATION_COUNT;
  nmi_reloc_result rr = objs.relocalize(Twc, gray, prm);
  std::printf("\nRELOC %d %d %d %.9g %.9g", rr.relocalized, rr.failed, rr.iterations, rr.nmi, rr.last_nmi);
  EXPECT(rr.n_levels == rr.iterations && rr.levels[rr.n_levels - 1].nmi == rr.nmi);
  {
    // the sharded driver as rank 0 of 1: the exchange step has nothing to combine, and the
    // result must be the single-GPU driver's, decision for decision
    int calls = 0;
    const nmi_exchange_fn same = [](void* user, void* key_dev, void* stream) -> int {
      ++*static_cast<int*>(user);
      return key_dev != nullptr && stream != nullptr ? 0 : 1;
    };
    nmi_reloc_result rs = objs.relocalizeSharded(Twc, gray, prm, 0, 1, nullptr, same, &calls);
    EXPECT(calls == rs.iterations && rs.iterations == rr.iterations);
    EXPECT(rs.relocalized == rr.relocalized && rs.failed == rr.failed && rs.nmi == rr.nmi);
    EXPECT(std::memcmp(rs.Twc, rr.Twc, sizeof rs.Twc) == 0);
  }
  std::printf("\nLOGPATH %s", objs.logPath.c_str());
  for (int i = 0; i < 16; i++) std::printf(" %.9g", rr.Twc[i]);
  std::stringstream ss;
  ss << *objs.NmiKernel;
  std::printf("\nKERNEL %s\nGPU OK\n", ss.str().c_str());
  nmi_compat::shutdown();
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 3 && std::string(argv[1]) == "outputs") return output_checks(argv[2]);
  if (argc >= 4 && std::string(argv[1]) == "loaders") return loader_checks(argv[2], argv[3]);
  if (argc >= 5 && std::string(argv[1]) == "host") return host_checks(argv[2], argv[3], argv[4]);
  if (argc >= 5 && std::string(argv[1]) == "gpu") return gpu_checks(argv[2], argv[3], argv[4]);
  std::printf("usage: test_compat host|gpu ...\n");
  return 2;
}
