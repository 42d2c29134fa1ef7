// argmax.cu -- on-device winner selection over the rating array (sm_100a).
//
// Replaces helperFunctions::find_max_elements + "take element [0]"
// (Thirdparty/Localization/helperFunctions.cpp:50-103, src/Tracking.cc:1952):
//   max starts at 0 and only a strictly greater score replaces it; the winner is
//   the lowest linear index (order wz,wy,wx,sz,sy,sx) whose score == max.
// The result is one packed 64-bit key
//   (bits(max) << 32) | (0xFFFFFFFF - index)          (low word 0: no winner)
// whose unsigned max over ranks reproduces the same rule across GPUs, so the
// multi-GPU combine is a single 8-byte max-allreduce.
#include "nmi_internal.h"

namespace nmi {
namespace {

constexpr int kArgThreads = 1024;

__global__ void __launch_bounds__(kArgThreads)
argmax_kernel(const float* __restrict__ scores, const uint32_t* __restrict__ list, int n_list,
              uint32_t n_total, unsigned long long* __restrict__ key,
              const uint32_t* __restrict__ retry_flag) {
  __shared__ float s_max[32];
  __shared__ uint32_t s_idx[32];
  __shared__ float s_m;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = list ? n_list : (int)n_total;

  float m = 0.0f;
  for (int i = tid; i < n; i += kArgThreads) {
    const float v = scores[list ? list[i] : (uint32_t)i];
    if (v > m) m = v;  // NaN never wins, exactly like the host loop
  }
#pragma unroll
  for (int d = 16; d >= 1; d /= 2) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  if (lane == 0) s_max[warp] = m;
  __syncthreads();
  if (warp == 0) {
    float x = s_max[lane];
#pragma unroll
    for (int d = 16; d >= 1; d /= 2) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, d));
    if (lane == 0) s_m = x;
  }
  __syncthreads();
  m = s_m;

  uint32_t best = 0xFFFFFFFFu;
  for (int i = tid; i < n; i += kArgThreads) {
    const uint32_t l = list ? list[i] : (uint32_t)i;
    if (scores[l] == m) best = min(best, l);
  }
#pragma unroll
  for (int d = 16; d >= 1; d /= 2) best = min(best, __shfl_xor_sync(0xffffffffu, best, d));
  if (lane == 0) s_idx[warp] = best;
  __syncthreads();
  if (warp == 0) {
    uint32_t x = s_idx[lane];
#pragma unroll
    for (int d = 16; d >= 1; d /= 2) x = min(x, __shfl_xor_sync(0xffffffffu, x, d));
    if (lane == 0) {
      // incomplete renders (record bins overflowed): publish the retry key, never a winner
      const bool retry = retry_flag != nullptr && *retry_flag != 0;
      *key = retry ? NMI_KEY_RETRY
                   : ((unsigned long long)__float_as_uint(m) << 32) | (unsigned long long)(0xFFFFFFFFu - x);
    }
  }
}

}  // namespace

void launch_argmax(const float* scores, const uint32_t* index_list, int n_list, uint32_t n_total,
                   unsigned long long* key, const uint32_t* retry_flag, cudaStream_t st) {
  argmax_kernel<<<1, kArgThreads, 0, st>>>(scores, index_list, n_list, n_total, key, retry_flag);
}

}  // namespace nmi
