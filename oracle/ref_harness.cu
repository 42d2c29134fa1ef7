// ref_harness.cu -- TEST INFRASTRUCTURE ONLY. Links with the reference's own NMI.cu + kernel.cu
// (compiled unmodified from /root/reference by oracle/Makefile.ref into oracle/_ref/) and gives
// the tests a plain C ABI to run them on the GPU box:
//   nmiref_score   CUDAF::NMIWithCuda_noMask itself (kernel.cuh:35-38), called the way
//                  Tracking does (src/Tracking.cc:1886-1894): device pointer cast to PtrStep*,
//                  SUC, render given as a texture name
//   nmiref_stages  the same sequence as kernel.cu:57-100, stage by stage, with every
//                  intermediate copied back (histograms, entropy terms, row sums, the three
//                  totals), plus the first total from a one-block launch so that the score can
//                  be recomputed without the inter-block race of NMI.cu:340-362
//   nmiref_time    wall time per NMIWithCuda_noMask call (the reference's per-evaluation cost:
//                  interop map, 10 cudaMalloc/cudaFree, 4 MiB memset, 6 launches, blocking D2H)
// Nothing under orbslam2_nmi_b200/ loads this library.
#include "ref_shim/refshim.h"
#undef texture
#include "NMI.cuh"      // reference header (found through -I/root/reference/Thirdparty/CUDA_Functions)
#include "kernel.cuh"

#include <cstdint>
#include <map>
#include <chrono>

// ------------------------------------------------------------------ texture reference shim
namespace refshim {
static std::map<const void*, cudaTextureObject_t> g_bound;

cudaError_t bind_array(const void* symbol, cudaArray_const_t array) {
  cudaResourceDesc rd;
  memset(&rd, 0, sizeof(rd));
  rd.resType = cudaResourceTypeArray;
  rd.res.array.array = const_cast<cudaArray_t>(array);
  cudaTextureDesc td;                         // a texture reference's default state
  memset(&td, 0, sizeof(td));
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModePoint;
  td.readMode = cudaReadModeElementType;
  td.normalizedCoords = 0;
  cudaTextureObject_t obj = 0;
  cudaError_t e = cudaCreateTextureObject(&obj, &rd, &td, nullptr);
  if (e != cudaSuccess) return e;
  unbind(symbol);
  g_bound[symbol] = obj;
  return cudaMemcpyToSymbol(symbol, &obj, sizeof(obj));
}

cudaError_t unbind(const void* symbol) {
  auto it = g_bound.find(symbol);
  if (it == g_bound.end()) return cudaSuccess;
  cudaError_t e = cudaDeviceSynchronize();    // kernels reading through the handle have finished
  cudaDestroyTextureObject(it->second);
  g_bound.erase(it);
  return e;
}
}  // namespace refshim

// ------------------------------------------------------------------ GL interop registry
struct RefshimResource { unsigned name; cudaArray_t array; bool mapped; };
static std::map<unsigned, cudaArray_t> g_textures;

extern "C" int refshim_register_gl_texture(unsigned int name, cudaArray_t array) {
  if (array) g_textures[name] = array; else g_textures.erase(name);
  return 0;
}
cudaError_t refshim_GLRegisterImage(cudaGraphicsResource_t* res, GLuint name, GLenum target, unsigned int) {
  auto it = g_textures.find(name);
  if (it == g_textures.end() || target != GL_TEXTURE_2D) return cudaErrorInvalidValue;
  *res = reinterpret_cast<cudaGraphicsResource_t>(new RefshimResource{name, it->second, false});
  return cudaSuccess;
}
cudaError_t refshim_MapResources(int n, cudaGraphicsResource_t* res, cudaStream_t) {
  for (int i = 0; i < n; i++) reinterpret_cast<RefshimResource*>(res[i])->mapped = true;
  return cudaSuccess;
}
cudaError_t refshim_UnmapResources(int n, cudaGraphicsResource_t* res, cudaStream_t) {
  for (int i = 0; i < n; i++) reinterpret_cast<RefshimResource*>(res[i])->mapped = false;
  return cudaSuccess;
}
cudaError_t refshim_GetMappedArray(cudaArray_t* array, cudaGraphicsResource_t res, unsigned int, unsigned int) {
  RefshimResource* r = reinterpret_cast<RefshimResource*>(res);
  if (!r->mapped) return cudaErrorNotMapped;
  *array = r->array;
  return cudaSuccess;
}
cudaError_t refshim_UnregisterResource(cudaGraphicsResource_t res) {
  delete reinterpret_cast<RefshimResource*>(res);
  return cudaSuccess;
}

// ------------------------------------------------------------------ C ABI for the tests
#define HCHECK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
  fprintf(stderr, "ref_harness: %s -> %s\n", #call, cudaGetErrorString(e_)); return (int)e_; } } while (0)

namespace {
struct Pair {
  cudaArray_t render = nullptr;   // what the GL colour texture holds: bottom-up rows, R8
  uint8_t* warped = nullptr;      // continuous W*H u8 (image.cpp:67)
  int W = 0, H = 0;
  unsigned name = 7;              // the "GL texture name"
};

int make_pair(Pair& p, const uint8_t* render_gl_rows, const uint8_t* warped, int W, int H) {
  p.W = W; p.H = H;
  cudaChannelFormatDesc cd = cudaCreateChannelDesc<unsigned char>();
  HCHECK(cudaMallocArray(&p.render, &cd, (size_t)W, (size_t)H));
  HCHECK(cudaMemcpy2DToArray(p.render, 0, 0, render_gl_rows, (size_t)W, (size_t)W, (size_t)H, cudaMemcpyHostToDevice));
  HCHECK(cudaMalloc(&p.warped, (size_t)W * H));
  HCHECK(cudaMemcpy(p.warped, warped, (size_t)W * H, cudaMemcpyHostToDevice));
  refshim_register_gl_texture(p.name, p.render);
  return 0;
}
void free_pair(Pair& p) {
  refshim_register_gl_texture(p.name, nullptr);
  if (p.render) cudaFreeArray(p.render);
  if (p.warped) cudaFree(p.warped);
  p = Pair();
}
}  // namespace

// render_gl_rows: H rows of W bytes, row 0 = BOTTOM of the image (GL texture order).
extern "C" int nmiref_score(const uint8_t* render_gl_rows, const uint8_t* warped, int W, int H, float* score) {
  Pair p;
  int rc = make_pair(p, render_gl_rows, warped, W, H);
  if (rc == 0) {
    *score = -2.f;
    CUDAF::NMIWithCuda_noMask((cv::cuda::PtrStep<unsigned char>*)p.warped, SUC, 0, W, H, score, p.name);
    rc = (int)cudaDeviceSynchronize();
  }
  free_pair(p);
  return rc;
}

// joint[65536], h1[256], h2[256], e1[256], e2[256], ej[65536], mid[256],
// sums[3] = {sum e1 (one-block launch, race-free), sum e2, sum mid}, raw_score = d_Entropy1[0]
// after the reference's three-block launch (what NMIWithCuda_noMask copies back).
extern "C" int nmiref_stages(const uint8_t* render_gl_rows, const uint8_t* warped, int W, int H,
                             uint32_t* joint, uint32_t* h1, uint32_t* h2, float* e1, float* e2, float* ej,
                             float* mid, float* sums, float* raw_score) {
  Pair p;
  int rc = make_pair(p, render_gl_rows, warped, W, H);
  if (rc) { free_pair(p); return rc; }
  uint *dH1, *dH2, *dJ;
  float *dE1, *dE2, *dEJ, *dMid, *dE1copy, *dDummy;
  HCHECK(cudaMalloc(&dH1, 256 * 4)); HCHECK(cudaMalloc(&dH2, 256 * 4)); HCHECK(cudaMalloc(&dJ, 65536 * 4));
  HCHECK(cudaMalloc(&dE1, 256 * 4)); HCHECK(cudaMalloc(&dE2, 256 * 4)); HCHECK(cudaMalloc(&dEJ, 65536 * 4));
  HCHECK(cudaMalloc(&dMid, 256 * 4)); HCHECK(cudaMalloc(&dE1copy, 256 * 4)); HCHECK(cudaMalloc(&dDummy, 256 * 4));
  initHistogram256all();
  histogram256all(dJ, dH1, dH2, p.warped, (uint)W, (uint)H, p.render);                   // kernel.cu:79
  HCHECK(cudaMemcpy(joint, dJ, 65536 * 4, cudaMemcpyDeviceToHost));
  HCHECK(cudaMemcpy(h1, dH1, 256 * 4, cudaMemcpyDeviceToHost));
  HCHECK(cudaMemcpy(h2, dH2, 256 * 4, cudaMemcpyDeviceToHost));
  ComputeEntropyKernel<<<258, 256>>>(dH1, dH2, dJ, W * H, dE1, dE2, dEJ);                // kernel.cu:83-85
  HCHECK(cudaMemcpy(e1, dE1, 256 * 4, cudaMemcpyDeviceToHost));
  HCHECK(cudaMemcpy(e2, dE2, 256 * 4, cudaMemcpyDeviceToHost));
  HCHECK(cudaMemcpy(ej, dEJ, 65536 * 4, cudaMemcpyDeviceToHost));
  AddvectorParwiseMidKernel<<<256, 128>>>(dEJ, dMid);                                    // kernel.cu:88-90
  HCHECK(cudaMemcpy(mid, dMid, 256 * 4, cudaMemcpyDeviceToHost));
  // first total alone: block 0 only, so nothing reads it while it is being summed
  HCHECK(cudaMemcpy(dE1copy, dE1, 256 * 4, cudaMemcpyDeviceToDevice));
  HCHECK(cudaMemset(dDummy, 0, 256 * 4));
  AddVectorPairwiseKernel<<<1, 128>>>(dE1copy, dDummy, dDummy);
  HCHECK(cudaMemcpy(&sums[0], dE1copy, 4, cudaMemcpyDeviceToHost));
  AddVectorPairwiseKernel<<<3, 128>>>(dE1, dE2, dMid);                                   // kernel.cu:93-95
  HCHECK(cudaMemcpy(raw_score, dE1, 4, cudaMemcpyDeviceToHost));
  HCHECK(cudaMemcpy(&sums[1], dE2, 4, cudaMemcpyDeviceToHost));
  HCHECK(cudaMemcpy(&sums[2], dMid, 4, cudaMemcpyDeviceToHost));
  closeHistogram256all();
  cudaFree(dH1); cudaFree(dH2); cudaFree(dJ); cudaFree(dE1); cudaFree(dE2); cudaFree(dEJ);
  cudaFree(dMid); cudaFree(dE1copy); cudaFree(dDummy);
  rc = (int)cudaDeviceSynchronize();
  free_pair(p);
  return rc;
}

// ms per call of the reference's entry point, `iters` calls on one resident pair.
extern "C" int nmiref_time(const uint8_t* render_gl_rows, const uint8_t* warped, int W, int H, int iters,
                           double* ms_per_eval, float* last_score) {
  Pair p;
  int rc = make_pair(p, render_gl_rows, warped, W, H);
  if (rc) { free_pair(p); return rc; }
  float s = 0.f;
  CUDAF::NMIWithCuda_noMask((cv::cuda::PtrStep<unsigned char>*)p.warped, SUC, 0, W, H, &s, p.name);  // warm-up
  cudaDeviceSynchronize();
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < iters; i++)
    CUDAF::NMIWithCuda_noMask((cv::cuda::PtrStep<unsigned char>*)p.warped, SUC, 0, W, H, &s, p.name);
  cudaDeviceSynchronize();
  auto t1 = std::chrono::steady_clock::now();
  *ms_per_eval = std::chrono::duration<double, std::milli>(t1 - t0).count() / (iters > 0 ? iters : 1);
  *last_score = s;
  free_pair(p);
  return 0;
}

extern "C" const char* nmiref_describe(void) {
  return "orbslam2_NMI Thirdparty/CUDA_Functions/{NMI.cu,kernel.cu}, unmodified, nvcc " 
#define REFSHIM_STR2(x) #x
#define REFSHIM_STR(x) REFSHIM_STR2(x)
         REFSHIM_STR(__CUDACC_VER_MAJOR__) "." REFSHIM_STR(__CUDACC_VER_MINOR__)
         " sm_100a, texture-reference/GL-interop/helper_cuda/PtrStep supplied by oracle/ref_shim";
}
