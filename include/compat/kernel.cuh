// kernel.cuh -- drop-in for Thirdparty/CUDA_Functions/kernel.cuh:22-38.
// Same namespace, name and signature as the reference's only device entry point, so
// src/Tracking.cc:1886-1894 compiles and links unchanged against libnmi_b200.so.
#pragma once
#include "nmi_compat_types.hpp"

#define ENMI 0   // kernel.cuh:22
#define SUC 1    // kernel.cuh:23

#define MATCHING_NMI 0
#define MATCHING_HOG 1
#define MATCHING_CANNY 2
#define MATCHING_HOUGH 3

namespace CUDAF {
// d_Warped : borrowed device pointer to a continuous width*height u8 image (the reference
//            passes GpuMat::data cast to PtrStep*, Tracking.cc:1887 / kernel.cu:79);
// NMI_mode, MatchingMode : accepted and ignored, exactly like kernel.cu:49-114 (the score
//            variant is nmi_compat::score_mode(), default SUC);
// NMI      : caller-owned host float, receives the score (blocking, kernel.cu:100);
// syntGL   : render handle from Rendering::getrenderedTexture().
// Errors print the message and exit(EXIT_FAILURE) like checkCudaErrors (kernel.cu:53).
void NMIWithCuda_noMask(cv::cuda::PtrStep<unsigned char>* d_Warped, int NMI_mode, int MatchingMode,
                        int width, int height, float* NMI, unsigned int syntGL);
}  // namespace CUDAF
