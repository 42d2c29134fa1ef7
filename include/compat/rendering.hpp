// rendering.hpp -- drop-in for Thirdparty/Localization/rendering.hpp:64-163.
// Same template, constructor and methods; the OpenGL pipeline (shaders, VBOs, FBO, depth
// renderbuffer) is replaced by the CUDA projection / z-buffer kernels of libnmi_b200.so.
// getrenderedTexture() returns an opaque render handle instead of a GL texture name; it is
// only ever passed back to CUDAF::NMIWithCuda_noMask (Tracking.cc:1893).
#pragma once
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "allProperties.hpp"
#include "nmi_compat.hpp"

template <unsigned char RenderingMode>
class Rendering {
  int windowWidth, windowHeight, imageWidth, imageHeight;
  int numSynthX, numSynthY, numSynthZ;
  float stepX, stepY, stepZ;
  float PointSize;
  glm::vec3 Camera_pos, Camera_direction, Camera_up;
  unsigned int renderedTexture = 0;
  std::string logPath;

  nmi_grid grid() const {
    nmi_grid g{};
    g.nS[0] = numSynthX; g.nS[1] = numSynthY; g.nS[2] = numSynthZ;
    g.nW[0] = g.nW[1] = g.nW[2] = 1;
    g.stepT[0] = stepX; g.stepT[1] = stepY; g.stepT[2] = stepZ;
    return g;
  }
  // setupCam convention (ioData.cpp:177-197) back to a camera->world matrix:
  // z_cam = dir - pos, y_cam = up, x_cam = y_cam x z_cam.
  void twc(float T[16]) const {
    const float p[3] = {Camera_pos.x, Camera_pos.y, Camera_pos.z}, d[3] = {Camera_direction.x, Camera_direction.y, Camera_direction.z},
                u[3] = {Camera_up.x, Camera_up.y, Camera_up.z};
    if (nmi_compat::recall_setup_cam(p, d, u, T)) return;  // the triple came from setupCam: the exact pose
    const glm::vec3 z = Camera_direction - Camera_pos, y = Camera_up;
    const glm::vec3 x(y.y * z.z - y.z * z.y, y.z * z.x - y.x * z.z, y.x * z.y - y.y * z.x);
    const float m[16] = {x.x, y.x, z.x, Camera_pos.x, x.y, y.y, z.y, Camera_pos.y,
                         x.z, y.z, z.z, Camera_pos.z, 0, 0, 0, 1};
    for (int i = 0; i < 16; i++) T[i] = m[i];
  }

 public:
  // rendering.hpp:483: window size is accepted and ignored (no window exists)
  Rendering(const float& pointSize, const int& windowwidth, const int& windowheight, const int& imagewidth,
            const int& imageheight, const int& num_of_views_x, const int& num_of_views_y,
            const int& num_of_views_z, const float& stepx, const float& stepy, const float& stepz,
            const glm::vec3& Cam_pos, const glm::vec3& Cam_dir, const glm::vec3& Cam_up,
            double near_clipping_plane, double far_clipping_plane, double fx, double fy, double cx, double cy,
            std::string object_path, std::string texture_path, std::string cloud_path,
            std::string offset_path, std::string log_Path)
      : windowWidth(windowwidth), windowHeight(windowheight), imageWidth(imagewidth),
        imageHeight(imageheight), numSynthX(num_of_views_x), numSynthY(num_of_views_y),
        numSynthZ(num_of_views_z), stepX(stepx), stepY(stepy), stepZ(stepz), PointSize(pointSize),
        Camera_pos(Cam_pos), Camera_direction(Cam_dir), Camera_up(Cam_up), logPath(log_Path) {
    nmi_camera& cam = nmi_compat::camera();
    cam.W = imagewidth; cam.H = imageheight;
    cam.fx = fx; cam.fy = fy; cam.cx = cx; cam.cy = cy;     // rendering.hpp:196-197
    cam.zn = near_clipping_plane; cam.zf = far_clipping_plane;  // rendering.hpp:198-199
    cam.point_size = pointSize;                              // rendering.hpp:307
    nmi_compat::apply_camera();
    if (RenderingMode == RENDER_POINT_CLOUD) {
      if (!cloud_path.empty()) {
        std::vector<float> xyzi;
        if (!nmi_compat::loadXYZ(cloud_path.c_str(), offset_path.c_str(), xyzi)) {
          std::fprintf(stderr, "Rendering: cannot load %s\n", cloud_path.c_str());
          std::exit(EXIT_FAILURE);
        }
        setCloud(xyzi.data(), xyzi.size() / 4);
      }
    } else if (!object_path.empty()) {
      // RENDER_TEXTURE (rendering.hpp:172-178, 219-229): OBJ + 24-bpp BMP
      // per-fragment texture shading like ShadingWithTexture.fragmentshader (nmi_set_mesh_textured)
      std::vector<float> verts, uv;
      std::vector<uint32_t> tris;
      std::vector<unsigned char> tex;
      int tw = 0, th = 0;
      if (!nmi_compat::meshFromObjBmp(object_path.c_str(), texture_path.c_str(), verts, tris, &uv, &tex, &tw, &th)) {
        std::fprintf(stderr, "Rendering: cannot load %s / %s\n", object_path.c_str(), texture_path.c_str());
        std::exit(EXIT_FAILURE);
      }
      nmi_compat::check(nmi_set_mesh_textured(nmi_compat::context(), verts.data(), verts.size() / 4, tris.data(),
                                              tris.size() / 3, uv.data(), tex.data(), tw, th),
                        "Rendering: textured mesh");
    }
  }

  // extension: hand over an in-memory mesh (nv x {x,y,z,grey}, nt x 3 indices)
  void setMesh(const float* verts, size_t nv, const uint32_t* tris, size_t nt) {
    nmi_compat::check(nmi_set_mesh(nmi_compat::context(), verts, nv, tris, nt), "Rendering::setMesh");
  }

  // extension: hand over an in-memory cloud (n x {x,y,z,I}) instead of a file
  void setCloud(const float* xyzi, size_t n) {
    nmi_compat::check(nmi_set_points(nmi_compat::context(), xyzi, n), "Rendering::setCloud");
  }

  // rendering.hpp:634
  bool setCamera(const glm::vec3& Cam_pos, const glm::vec3& Cam_dir, const glm::vec3& Cam_up) {
    Camera_direction = Cam_dir;
    Camera_pos = Cam_pos;
    Camera_up = Cam_up;
    return true;
  }

  // rendering.hpp:530-630: render the model from Camera_pos + translation; result stays on
  // the device behind getrenderedTexture().
  bool renderToTextureOnGPU(const glm::vec3& translation) {
    float T[16];
    twc(T);
    const float t[3] = {translation.x, translation.y, translation.z};
    nmi_compat::check(nmi_render_at(nmi_compat::context(), T, t, &renderedTexture),
                      "Rendering::renderToTextureOnGPU");
    return true;
  }
  // rendering.hpp:415-527: same, then read back.  Rows are returned bottom-up, as
  // glReadPixels delivers them (rendering.hpp:520).
  bool renderToTexture(const glm::vec3& translation, std::vector<unsigned char>& synth) {
    renderToTextureOnGPU(translation);
    std::vector<unsigned char> top_down((size_t)imageWidth * imageHeight);
    nmi_compat::check(nmi_get_render(nmi_compat::context(), 0, top_down.data()),
                      "Rendering::renderToTexture");
    synth.resize(top_down.size());
    for (int r = 0; r < imageHeight; r++)
      std::copy(top_down.begin() + (size_t)r * imageWidth, top_down.begin() + (size_t)(r + 1) * imageWidth,
                synth.begin() + (size_t)(imageHeight - 1 - r) * imageWidth);
    return true;
  }

  // rendering.hpp:644-665
  glm::vec3 calculateTranslation(int synthx, int synthy, int synthz) {
    float T[16], t[3];
    twc(T);
    nmi_grid g = grid();
    nmi_cell_translation(T, &g, synthx, synthy, synthz, t);
    return glm::vec3(t[0], t[1], t[2]);
  }
  // rendering.hpp:669-696
  cv::Mat calculateTranslationCV(int synthx, int synthy, int synthz) {
    const glm::vec3 t = calculateTranslation(synthx, synthy, synthz);
    cv::Mat m(3, 1, CV_32F);
    m.at<float>(0, 0) = t.x;
    m.at<float>(1, 0) = t.y;
    m.at<float>(2, 0) = t.z;
    return m;
  }

  // rendering.hpp:712-723
  void resizeKernel(const int& numsynthx, const int& numsynthy, const int& numsynthz, const float& stepx,
                    const float& stepy, const float& stepz) {
    stepX = stepx; stepY = stepy; stepZ = stepz;
    numSynthX = numsynthx; numSynthY = numsynthy; numSynthZ = numsynthz;
  }

  int getImageWidth() { return imageWidth; }
  int getImageHeight() { return imageHeight; }
  int getNumSynthX() { return numSynthX; }
  int getNumSynthY() { return numSynthY; }
  int getNumSynthZ() { return numSynthZ; }
  float getStepX() { return stepX; }
  float getStepY() { return stepY; }
  float getStepZ() { return stepZ; }
  unsigned int getFramebufferName() { return 0; }
  unsigned int getrenderedTexture() { return renderedTexture ? renderedTexture : 1u; }
  void setSynthetic_count_x(int v) { numSynthX = v; }
  void setSynthetic_count_y(int v) { numSynthY = v; }
  void setSynthetic_count_z(int v) { numSynthZ = v; }
  void setStep_x(float v) { stepX = v; }
  void setStep_y(float v) { stepY = v; }
  void setStep_z(float v) { stepZ = v; }
};
