// image.hpp -- drop-in for Thirdparty/Localization/image.hpp:26-74 / image.cpp.
// Same constructor and methods; the warps live in the nmi context (one batched kernel for
// all rotation cells instead of nW cv::cuda::warpPerspective calls, image.cpp:115-128).
#pragma once
#include "nmi_compat.hpp"

class Image {
  cv::Mat imgOriginal;
  int numWarpZ, numWarpY, numWarpX;
  float stepRadZ, stepRadY, stepRadX;  // radian
  cv::Mat K_mat;
  int width, height;

  nmi_grid grid() const {
    nmi_grid g{};
    g.nS[0] = g.nS[1] = g.nS[2] = 1;
    g.nW[0] = numWarpX; g.nW[1] = numWarpY; g.nW[2] = numWarpZ;
    g.stepR[0] = stepRadX; g.stepR[1] = stepRadY; g.stepR[2] = stepRadZ;
    return g;
  }

 public:
  // image.cpp:33: note the (z, y, x) argument order of the reference
  Image(const int& numWarpz, const int& numWarpy, const int& numWarpx, const float& stepWarpz,
        const float& stepWarpy, const float& stepWarpx, const int& width, const int& height, cv::Mat K)
      : numWarpZ(numWarpz), numWarpY(numWarpy), numWarpX(numWarpx), stepRadZ(stepWarpz),
        stepRadY(stepWarpy), stepRadX(stepWarpx), width(width), height(height) {
    K_mat = K.clone();
    nmi_camera& cam = nmi_compat::camera();
    cam.W = width;
    cam.H = height;
    cam.fx = K.at<double>(0, 0);
    cam.fy = K.at<double>(1, 1);
    cam.cx = K.at<double>(0, 2);
    cam.cy = K.at<double>(1, 2);
    nmi_compat::apply_camera();
  }

  // image.cpp:130-135: host frame -> device (continuous u8, width*height)
  void loadOriginal(cv::Mat Original) {
    Original.copyTo(imgOriginal);
    nmi_compat::check(nmi_set_frame(nmi_compat::context(), imgOriginal.data, width, height),
                      "Image::loadOriginal");
  }
  // image.cpp:115-128: all numWarpZ*numWarpY*numWarpX homography warps of the frame
  void calculateWarping() {
    nmi_grid g = grid();
    nmi_compat::check(nmi_warp_cells(nmi_compat::context(), &g), "Image::calculateWarping");
  }
  // image.cpp:142: header over the device-resident warp (continuous, step == width)
  cv::cuda::GpuMat getImageGPU(const int& indexZ, const int& indexY, const int& indexX) {
    nmi_grid g = grid();
    void* p = nullptr;
    nmi_compat::check(nmi_warp_ptr(nmi_compat::context(), &g, indexX, indexY, indexZ, &p),
                      "Image::getImageGPU");
    cv::cuda::GpuMat m;
    m.data = static_cast<unsigned char*>(p);
    m.rows = height;
    m.cols = width;
    m.step = (size_t)width;
    return m;
  }
  // image.cpp:186: new rotation grid (note the (x, y, z) order here, as in the reference)
  void resizeKernel(const int& numwarpx, const int& numwarpy, const int& numwarpz, const float& stepradx,
                    const float& steprady, const float& stepradz) {
    numWarpX = numwarpx; numWarpY = numwarpy; numWarpZ = numwarpz;
    stepRadX = stepradx; stepRadY = steprady; stepRadZ = stepradz;
  }

  cv::Mat getOriginal() { return imgOriginal; }
  cv::Mat getK() { return K_mat; }
  float getStepZ() { return stepRadZ; }
  float getStepY() { return stepRadY; }
  float getStepX() { return stepRadX; }
  void setStepZ(float s) { stepRadZ = s; }
  void setStepY(float s) { stepRadY = s; }
  void setStepX(float s) { stepRadX = s; }
  int getNumWarpZ() { return numWarpZ; }
  int getNumWarpY() { return numWarpY; }
  int getNumWarpX() { return numWarpX; }
  void setNumWarpZ(int n) { numWarpZ = n; }
  void setNumWarpY(int n) { numWarpY = n; }
  void setNumWarpX(int n) { numWarpX = n; }
  int getWidth() const { return width; }
  int getHeight() const { return height; }
};
