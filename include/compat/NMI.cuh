// NMI.cuh -- drop-in for Thirdparty/CUDA_Functions/NMI.cuh:60-78, the secondary exports only
// kernel.cu uses.  A caller that keeps the reference's own
// NMIWithCuda_noMask body (kernel.cu:49-114: map the GL texture, histogram256all, three
// launches, blocking copy) links against these; the batched path never goes through them.
//
//   initHistogram256all / closeHistogram256all   no-ops: scratch lives in the nmi_ctx
//   histogram256all                               joint + marginal histograms of a device
//                                                 image and a cudaArray render, computed by the
//                                                 fused shared-memory kernel (csrc/hist.cu)
//   ComputeEntropyKernel, AddvectorParwiseMidKernel, AddVectorPairwiseKernel
//                                                 __global__ functions cannot be exported from a
//                                                 shared library without relocatable device
//                                                 code, so they are defined here (nvcc only),
//                                                 for the reference's launch shapes
//                                                 <<<258,256>>>, <<<256,128>>>, <<<3,128>>>.
// Same arithmetic and summation order as the fused kernel; AddVectorPairwiseKernel
// orders its three blocks with a counter instead of racing like NMI.cu:340-362.
#pragma once
#include <cuda_runtime.h>

typedef unsigned int uint;   // NMI.cuh:25-26
typedef unsigned char uchar;

#define HISTOGRAM256_BIN_COUNT 256  // NMI.cuh:39
#define JOINT_HISTOGRAM256_BIN_COUNT (HISTOGRAM256_BIN_COUNT * HISTOGRAM256_BIN_COUNT)

#ifndef ENMI
#define ENMI 0  // kernel.cuh:22
#endif
#ifndef SUC
#define SUC 1  // kernel.cuh:23
#endif

extern "C" void initHistogram256all(void);
extern "C" void closeHistogram256all(void);
// d_JointHistogram[render * 256 + camera], d_Histogram1 = render, d_Histogram2 = camera
// (NMI.cu:79-87); synthCUDA rows are bottom-up like the GL texture and are flipped (NMI.cu:82).
extern "C" void histogram256all(uint* d_JointHistogram, uint* d_Histogram1, uint* d_Histogram2,
                                uchar* d_Warped, uint width, uint height, cudaArray* synthCUDA);

#ifdef __CUDACC__
namespace nmi_compat_detail {
__device__ __forceinline__ float entropy_term(uint count, int length) {  // NMI.cu:240-266
  if (count == 0) return 0.0f;
  const float p = __fdiv_rn((float)count, (float)length);
  return __fmul_rn(p, log2f(p));  // libdevice log2f, what NMI.cu:248 calls (see csrc/hist.cu)
}
// sum of 256 floats in the order of NMI.cu:270-287: strides 128, 64, ..., 1.  128 threads.
__device__ __forceinline__ float tree256(const float* in, float* scratch) {
  const int t = threadIdx.x;
  scratch[t] = __fadd_rn(in[t], in[t + 128]);
  __syncthreads();
  for (int n = 64; n >= 1; n /= 2) {
    if (t < n) scratch[t] = __fadd_rn(scratch[t], scratch[t + n]);
    __syncthreads();
  }
  return scratch[0];
}
}  // namespace nmi_compat_detail

// <<<258, 256>>>: block 0 -> histogram 1, block 1 -> histogram 2, blocks 2.. -> joint rows
static __global__ void ComputeEntropyKernel(uint* d_Histogram1, uint* d_Histogram2, uint* d_JointHistogram,
                                            int length, float* d_EntropyArray1, float* d_EntropyArray2,
                                            float* d_JointEntropyArray) {
  using nmi_compat_detail::entropy_term;
  const int t = threadIdx.x;
  if (blockIdx.x == 0) {
    d_EntropyArray1[t] = entropy_term(d_Histogram1[t], length);
  } else if (blockIdx.x == 1) {
    d_EntropyArray2[t] = entropy_term(d_Histogram2[t], length);
  } else {
    const int i = blockDim.x * (blockIdx.x - 2) + t;
    d_JointEntropyArray[i] = entropy_term(d_JointHistogram[i], length);
  }
}

// <<<256, 128>>>: d_out[row] = sum of the row's 256 joint-entropy terms
static __global__ void AddvectorParwiseMidKernel(float* d_Array, float* d_out) {
  __shared__ float s[128];
  const float v = nmi_compat_detail::tree256(d_Array + 256 * blockIdx.x, s);
  if (threadIdx.x == 0) d_out[blockIdx.x] = v;
}

// <<<3, 128>>>: sums of the three 256-arrays, then the score into d_Array1[0]
// (kernel.cuh:22-23 picks the variant at compile time: -DNMI_COMPAT_SCORE=ENMI, default SUC)
#ifndef NMI_COMPAT_SCORE
#define NMI_COMPAT_SCORE SUC
#endif
static __global__ void AddVectorPairwiseKernel(float* d_Array1, float* d_Array2, float* d_Array3) {
  // The reference's three blocks each reduce one array and block 1 then reads the other two blocks'
  // results without any grid-wide synchronisation (NMI.cu:340-342: a race that does fire on B200).
  // Here block 0 walks the three trees one after the other -- same order of additions inside each tree,
  // no inter-block communication, no global state (two launches on different streams cannot disturb
  // each other); blocks 1 and 2 of the reference's <<<3, 128>>> launch shape simply return.
  __shared__ float s[128];
  if (blockIdx.x != 0) return;
  const float sa = nmi_compat_detail::tree256(d_Array1, s);
  __syncthreads();
  const float sb = nmi_compat_detail::tree256(d_Array2, s);
  __syncthreads();
  const float sab = nmi_compat_detail::tree256(d_Array3, s);
  if (threadIdx.x == 0) {
    d_Array2[0] = sb;   // the reference leaves each array's total in its element 0
    d_Array3[0] = sab;
    float score;
    if (sa == 0.0f && sb == 0.0f && sab == 0.0f)
      score = 0.0f;  // NMI.cu:344,353
    else if (NMI_COMPAT_SCORE == ENMI)
      score = __fdiv_rn(__fadd_rn(-sa, -sb), -sab);  // NMI.cu:348
    else
      score = __fmul_rn(2.0f, __fsub_rn(1.0f, __fdiv_rn(-sab, __fadd_rn(-sa, -sb))));  // NMI.cu:357
    d_Array1[0] = score;
  }
}
#endif  // __CUDACC__
