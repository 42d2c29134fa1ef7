// ubench_atoms.cu -- shared-memory atomic throughput on B200 (design input for hist.cu).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_atoms tools/ubench_atoms.cu
// Prints atomics/s (whole chip) and lanes/clk/SM at the SM clock it measures itself.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

constexpr int kThreads = 512;
constexpr int kWords = 32768;  // 128 KiB
constexpr int kIters = 4096;

enum Mode { RANDOM = 0, DISTINCT_BANKS, SAME_ADDR, RANDOM_RET, U16_RET, MATCH_RANDOM, MATCH_16VALS,
            SMOOTH, RANDOM_1K, SAME_BANK_DIFF_ADDR, RANDOM_256T, RANDOM_1024T, LOWENT16, NMODES };
const char* kNames[] = {"random idx (32K words), no return", "distinct banks, random rows", "same address (all lanes)",
                        "random idx, return value used", "u16-packed + return + wrap check (hist.cu P_U16G)",
                        "match_any + leader add, random idx", "match_any + leader add, 16 distinct values",
                        "smooth image-like idx (neighbouring lanes +-2)", "random idx in 1K words",
                        "same bank, different addresses", "random idx, 256 thr/CTA x2 CTAs/SM?", "random idx 1024 thr",
                        "plain atomics, 16 distinct values/warp"};

__device__ __forceinline__ uint32_t lcg(uint32_t& x) { x = x * 1664525u + 1013904223u; return x; }

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_atoms(uint32_t* out, long long* cycles) {
  extern __shared__ uint32_t h[];
  for (int i = threadIdx.x; i < kWords; i += blockDim.x) h[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  uint32_t x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
  uint32_t acc = 0;
  long long t0 = clock64();
#pragma unroll 8
  for (int it = 0; it < kIters; it++) {
    uint32_t r = lcg(x);
    if (MODE == RANDOM) {
      atomicAdd(&h[r >> 17], 1u);
    } else if (MODE == DISTINCT_BANKS) {
      uint32_t rr = __shfl_sync(0xffffffffu, r, 0);  // same row for the warp
      atomicAdd(&h[((rr >> 22) << 5) + lane], 1u);
    } else if (MODE == SAME_ADDR) {
      uint32_t rr = __shfl_sync(0xffffffffu, r, 0);
      atomicAdd(&h[rr >> 17], 1u);
    } else if (MODE == RANDOM_RET) {
      acc ^= atomicAdd(&h[r >> 17], 1u);
    } else if (MODE == U16_RET) {
      uint32_t t = r >> 16;
      uint32_t sh = (t & 1u) << 4;
      uint32_t old = atomicAdd(&h[t >> 1], 1u << sh);
      if (((old >> sh) & 0x3FFFu) == 0x3FFFu) { atomicSub(&h[t >> 1], 0x4000u << sh); acc++; }
    } else if (MODE == MATCH_RANDOM || MODE == MATCH_16VALS) {
      uint32_t idx = MODE == MATCH_RANDOM ? (r >> 17) : ((r >> 28) * 37u);
      unsigned m = __match_any_sync(0xffffffffu, idx);
      if (lane == (__ffs(m) - 1)) atomicAdd(&h[idx], (uint32_t)__popc(m));
    } else if (MODE == SMOOTH) {
      uint32_t rr = __shfl_sync(0xffffffffu, r, 0);
      uint32_t a = ((rr >> 24) + ((r >> 5) & 3u)) & 255u, b = ((rr >> 16) + ((r >> 9) & 3u)) & 255u;
      atomicAdd(&h[((a << 8) | b) >> 1], 1u);
    } else if (MODE == RANDOM_1K) {
      atomicAdd(&h[r >> 22], 1u);
    } else if (MODE == SAME_BANK_DIFF_ADDR) {
      atomicAdd(&h[(r >> 22) << 5], 1u);
    } else if (MODE == LOWENT16) {
      atomicAdd(&h[(r >> 28) * 37u], 1u);
    } else {
      atomicAdd(&h[r >> 17], 1u);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  uint32_t s = acc;
  for (int i = threadIdx.x; i < kWords; i += blockDim.x) s += h[i];
  if (s == 0xdeadbeef) out[0] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(int threads, int ctas_per_sm, int sms, uint32_t* d_out, long long* d_cyc) {
  auto k = k_atoms<MODE>;
  size_t smem = kWords * 4 / ctas_per_sm;  // split the 128 KiB when two CTAs share an SM
  if (ctas_per_sm == 1) smem = kWords * 4;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kWords * 4));
  int grid = sms * ctas_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<grid, threads, kWords * 4 / (ctas_per_sm > 1 ? 1 : 1), 0>>>(d_out, d_cyc);  // warm
  cudaEventRecord(e0);
  k<<<grid, threads, kWords * 4, 0>>>(d_out, d_cyc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long* h = (long long*)malloc(sizeof(long long) * grid);
  cudaMemcpy(h, d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double cyc = 0; for (int i = 0; i < grid; i++) cyc += (double)h[i]; cyc /= grid;
  free(h);
  double total = (double)grid * threads * kIters;
  printf("%-58s thr=%4d  %8.1f Gatom/s  %6.2f lanes/clk/SM (in-kernel)  %7.3f ms  err=%s\n", kNames[MODE], threads,
         total / ms / 1e6, (double)threads * kIters / cyc, ms, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s  SMs=%d  smem/SM=%zu\n", p.name, p.multiProcessorCount, p.sharedMemPerMultiprocessor);
  uint32_t* d_out; long long* d_cyc;
  cudaMalloc(&d_out, 4); cudaMalloc(&d_cyc, 8 * 4096);
  int sms = p.multiProcessorCount;
  run<RANDOM>(512, 1, sms, d_out, d_cyc);
  run<RANDOM>(256, 1, sms, d_out, d_cyc);
  run<RANDOM>(1024, 1, sms, d_out, d_cyc);
  run<DISTINCT_BANKS>(512, 1, sms, d_out, d_cyc);
  run<SAME_ADDR>(512, 1, sms, d_out, d_cyc);
  run<SAME_BANK_DIFF_ADDR>(512, 1, sms, d_out, d_cyc);
  run<RANDOM_RET>(512, 1, sms, d_out, d_cyc);
  run<U16_RET>(512, 1, sms, d_out, d_cyc);
  run<U16_RET>(1024, 1, sms, d_out, d_cyc);
  run<MATCH_RANDOM>(512, 1, sms, d_out, d_cyc);
  run<MATCH_16VALS>(512, 1, sms, d_out, d_cyc);
  run<LOWENT16>(512, 1, sms, d_out, d_cyc);
  run<SMOOTH>(512, 1, sms, d_out, d_cyc);
  run<RANDOM_1K>(512, 1, sms, d_out, d_cyc);
  return 0;
}
