// compat_nmi.cu -- the extern "C" histogram entry points of Thirdparty/CUDA_Functions/NMI.cuh:60-71
// (bodies NMI.cu:171-226), kept for callers that still run the reference's own
// NMIWithCuda_noMask (kernel.cu:49-114).  The three __global__ functions of NMI.cuh:74-78 are
// header-only (include/compat/NMI.cuh).  Everything numerical goes through the C ABI.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "../../include/compat/nmi_compat.hpp"

namespace {
unsigned char* g_linear = nullptr;  // cudaArray contents as rows (bottom-up, like the GL texture)
size_t g_linear_cap = 0;

void cuda_or_die(cudaError_t e, const char* what) {  // checkCudaErrors (kernel.cu:53): report and exit
  if (e != cudaSuccess) {
    std::fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e));
    std::exit(EXIT_FAILURE);
  }
}
}  // namespace

// NMI.cu:171-186 allocate / free the 16 partial histograms of the reference's kernel; the
// fused kernel keeps its histogram in shared memory and its scratch in the context.
extern "C" void initHistogram256all(void) {}
extern "C" void closeHistogram256all(void) {}

extern "C" void histogram256all(unsigned int* d_JointHistogram, unsigned int* d_Histogram1,
                                unsigned int* d_Histogram2, unsigned char* d_Warped, unsigned int width,
                                unsigned int height, cudaArray* synthCUDA) {
  nmi_ctx* ctx = nmi_compat::context();
  nmi_camera& cam = nmi_compat::camera();
  if (cam.W != (int)width || cam.H != (int)height || cam.fx == 0) {
    // called without Rendering / Image having set a camera: only the image size matters here
    cam.W = (int)width;
    cam.H = (int)height;
    if (cam.fx == 0) {
      cam.fx = cam.fy = (double)width;
      cam.cx = width / 2.0;
      cam.cy = height / 2.0;
    }
    nmi_compat::check(nmi_set_camera(ctx, &cam), "histogram256all: nmi_set_camera");
  }
  cudaStream_t st = static_cast<cudaStream_t>(nmi_ctx_stream(ctx));
  const size_t bytes = (size_t)width * height;
  if (bytes > g_linear_cap) {
    if (g_linear) cudaFree(g_linear);
    cuda_or_die(cudaMalloc(&g_linear, bytes), "histogram256all: cudaMalloc");
    g_linear_cap = bytes;
  }
  cuda_or_die(cudaMemcpy2DFromArrayAsync(g_linear, width, synthCUDA, 0, 0, width, height,
                                         cudaMemcpyDeviceToDevice, st),
              "histogram256all: cudaMemcpy2DFromArray");
  unsigned int handle = 0;
  nmi_compat::check(nmi_import_render(ctx, g_linear, width, (int)width, (int)height, /*bottom_up=*/1, &handle),
                    "histogram256all: nmi_import_render");
  nmi_flags f = nmi_compat::flags();
  f.bins = 256;  // HISTOGRAM256_BIN_COUNT
  nmi_compat::check(nmi_eval_pair_dev(ctx, d_Warped, handle, (int)width, (int)height, &f, d_JointHistogram,
                                      d_Histogram1, d_Histogram2, nullptr),
                    "histogram256all: nmi_eval_pair_dev");  // returns after a stream sync
}
