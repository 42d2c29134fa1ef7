/* refshim.h -- TEST INFRASTRUCTURE ONLY (see oracle/README_ref.md).
 *
 * Lets the reference's Thirdparty/CUDA_Functions/{NMI.cu,kernel.cu} compile UNMODIFIED with
 * CUDA 12.9 for sm_100a, from where they lie under /root/reference, into oracle/_ref/.
 * It supplies only the APIs those two files use that no longer exist or are absent here:
 *   - CUDA 9 texture REFERENCES (`texture<T,dim,mode>`, `tex2D(ref,x,y)`,
 *     `cudaBindTextureToArray`, `cudaUnbindTexture`; removed in CUDA 12) on top of texture
 *     objects, with a texture reference's default state: point filter, clamp addressing,
 *     unnormalised coordinates, element read mode;
 *   - `checkCudaErrors` of the CUDA samples' helper_cuda.h (print + exit);
 *   - `cv::cuda::PtrStep<T>` (only ever used as a pointer type that kernel.cu casts away,
 *     kernel.cu:79) and the GL typedefs / GL_TEXTURE_2D of glew.h;
 *   - the four CUDA<->GL interop calls of kernel.cu:52-56,110-111: a "GL texture name" is
 *     resolved through a small registry of cudaArrays the test harness fills
 *     (refshim_register_gl_texture) -- there is no GL on this machine.
 * No arithmetic of the reference is replaced: every kernel, launch shape, merge, entropy term,
 * tree and score formula that runs is the reference's own source.
 */
#ifndef NMI_REFSHIM_H_
#define NMI_REFSHIM_H_

#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <chrono>
#include <vector>

/* ---- helper_cuda.h ---- */
#ifndef checkCudaErrors
#define checkCudaErrors(call)                                                                   \
  do {                                                                                          \
    cudaError_t refshim_e_ = (call);                                                            \
    if (refshim_e_ != cudaSuccess) {                                                            \
      fprintf(stderr, "CUDA error at %s:%d code=%d(%s) \"%s\"\n", __FILE__, __LINE__,          \
              (int)refshim_e_, cudaGetErrorName(refshim_e_), #call);                            \
      exit(EXIT_FAILURE);                                                                       \
    }                                                                                           \
  } while (0)
#endif

/* ---- GL names used by kernel.cu ---- */
typedef unsigned int GLuint;
typedef unsigned int GLenum;
#ifndef GL_TEXTURE_2D
#define GL_TEXTURE_2D 0x0DE1
#endif

/* ---- OpenCV type that appears in the entry point's signature ---- */
namespace cv { namespace cuda {
template <typename T> struct PtrStep { T* data; size_t step; };
} }

/* ---- texture references ---- */
namespace refshim {
template <typename T, int Dim, cudaTextureReadMode Mode>
struct texref { cudaTextureObject_t obj; };

/* host side of the one bound reference (NMI.cu binds, launches, unbinds) */
cudaError_t bind_array(const void* symbol, cudaArray_const_t array);
cudaError_t unbind(const void* symbol);
}  // namespace refshim

template <typename T, int Dim, cudaTextureReadMode Mode>
static __device__ __forceinline__ T tex2D(const refshim::texref<T, Dim, Mode>& ref, float x, float y) {
  return tex2D<T>(ref.obj, x, y);
}
template <typename T, int Dim, cudaTextureReadMode Mode>
static inline cudaError_t cudaBindTextureToArray(const refshim::texref<T, Dim, Mode>& ref, cudaArray_const_t array) {
  return refshim::bind_array((const void*)&ref, array);
}
template <typename T, int Dim, cudaTextureReadMode Mode>
static inline cudaError_t cudaUnbindTexture(const refshim::texref<T, Dim, Mode>& ref) {
  return refshim::unbind((const void*)&ref);
}
/* `texture<uchar, cudaTextureType2D, cudaReadModeElementType> texCUDA;` at namespace scope
 * (NMI.cu:39) becomes a __device__ variable holding a texture object handle. */
#define texture __device__ ::refshim::texref

/* ---- CUDA <-> GL interop (kernel.cu:52-56, 110-111) ---- */
extern "C" int refshim_register_gl_texture(unsigned int name, cudaArray_t array); /* array == 0 removes it */
cudaError_t refshim_GLRegisterImage(cudaGraphicsResource_t* res, GLuint name, GLenum target, unsigned int flags);
cudaError_t refshim_MapResources(int n, cudaGraphicsResource_t* res, cudaStream_t s = 0);
cudaError_t refshim_UnmapResources(int n, cudaGraphicsResource_t* res, cudaStream_t s = 0);
cudaError_t refshim_GetMappedArray(cudaArray_t* array, cudaGraphicsResource_t res, unsigned int idx, unsigned int mip);
cudaError_t refshim_UnregisterResource(cudaGraphicsResource_t res);
#define cudaGraphicsGLRegisterImage refshim_GLRegisterImage
#define cudaGraphicsMapResources refshim_MapResources
#define cudaGraphicsUnmapResources refshim_UnmapResources
#define cudaGraphicsSubResourceGetMappedArray refshim_GetMappedArray
#define cudaGraphicsUnregisterResource refshim_UnregisterResource

#endif /* NMI_REFSHIM_H_ */
