// npp_check.cu -- TEST INFRASTRUCTURE (a checker, never linked or loaded by the product).
//
// The reference warps the camera frame with cv::cuda::warpPerspective (Thirdparty/Localization/
// image.cpp:123; OpenCV 3.4.0 + CUDA 9.2, un-vendored).  For 8UC1 / INTER_LINEAR /
// BORDER_CONSTANT that call ends in NPP's nppiWarpPerspective_8u_C1R (when OpenCV's useNpp table
// allows it) or in OpenCV's own fp32 kernel -- neither is in the reference tree.  What IS in this
// image is libnppig (CUDA 12.x), so the GPU tests run THAT routine next to warp_kernel on the same
// frame and forward matrices K R K^-1 and report how far the two are apart
// (tests/test_gpu_npp_warp.py, profiles/r02_npp_warp.json).  Caveat stated there too: NPP 12.4 is
// not NPP 9.2; this pins the warp against the library family the reference reaches, not its build.
//
// Build: oracle/Makefile (target npp) -> oracle/_build/libnmi_nppcheck.so
#include <cuda_runtime.h>
#include <nppi_geometry_transforms.h>
#include <nppcore.h>
#include <stdint.h>

extern "C" {

int nppchk_version(int* major, int* minor, int* build) {
  const NppLibraryVersion* v = nppGetLibVersion();
  if (!v) return -1;
  *major = v->major;
  *minor = v->minor;
  *build = v->build;
  return 0;
}

// dst = warpPerspective(src, M) the way cv::cuda::warpPerspective's NPP branch issues it:
// forward coefficients, full-image source and destination ROI, NPPI_INTER_LINEAR; pixels NPP does
// not write keep the 0 the destination was cleared to (BORDER_CONSTANT, value 0).
// Returns the NppStatus (0 = success, > 0 warnings) or -1000 - cudaError.
int nppchk_warp_perspective_8u(const uint8_t* src_host, int W, int H, const double M[9], int interp_linear,
                               uint8_t* dst_host) {
  uint8_t *d_src = nullptr, *d_dst = nullptr;
  size_t ps = 0, pd = 0;
  cudaError_t e = cudaMallocPitch(&d_src, &ps, W, H);
  if (e != cudaSuccess) return -1000 - (int)e;
  e = cudaMallocPitch(&d_dst, &pd, W, H);
  if (e != cudaSuccess) { cudaFree(d_src); return -1000 - (int)e; }
  cudaMemcpy2D(d_src, ps, src_host, W, W, H, cudaMemcpyHostToDevice);
  cudaMemset2D(d_dst, pd, 0, W, H);
  const NppiSize ssz = {W, H};
  const NppiRect sroi = {0, 0, W, H}, droi = {0, 0, W, H};
  const double c[3][3] = {{M[0], M[1], M[2]}, {M[3], M[4], M[5]}, {M[6], M[7], M[8]}};
  const NppStatus st = nppiWarpPerspective_8u_C1R(d_src, ssz, (int)ps, sroi, d_dst, (int)pd, droi, c,
                                                 interp_linear ? NPPI_INTER_LINEAR : NPPI_INTER_NN);
  e = cudaDeviceSynchronize();
  cudaMemcpy2D(dst_host, W, d_dst, pd, W, H, cudaMemcpyDeviceToHost);
  cudaFree(d_src);
  cudaFree(d_dst);
  if (e != cudaSuccess) return -1000 - (int)e;
  return (int)st;
}

}  // extern "C"
