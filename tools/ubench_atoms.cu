// ubench_atoms.cu -- shared-memory atomic throughput on B200 (design input for csrc/hist.cu and the
// denominator of bench.py's "smem_atomic" roofline).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_atoms tools/ubench_atoms.cu
// Prints atomics/s (whole chip) and lanes/clk/SM at the SM clock the kernel measures itself.
//
// Round-2 rewrite: round 1 seeded an LCG with threadIdx.x * constant, so the 32 lanes of a warp
// walked an arithmetic progression -- its "random" case was close to conflict-free (12.3 lanes/clk
// vs 12.8 for distinct banks).  Here every lane owns an independently seeded stream, address
// generation costs one IMAD per index so that the ATOMS pipe and not the issue port bounds every
// case, and the cases that matter for the histogram kernel are separate: conflict-free, 2-way /
// 4-way conflicts, uniformly random words, the kernel's own packed-u16 + swizzle + return-value
// pattern, and that pattern with the LDS.128 staging reads.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

constexpr int kWords = 32768;  // 128 KiB, the joint histogram of hist.cu
constexpr int kIters = 4096;

enum Mode { HASH_ONLY = 0, DISTINCT_BANKS, RANDOM, RANDOM_RET, SAME_ADDR, SAME_BANK_DIFF_ADDR, TWO_WAY, HALF_WARP_RANDOM,
            HIST_PATTERN, HIST_PATTERN_LDS, RANDOM_RED, FOUR_WAY, NMODES };
const char* kNames[] = {"control: hash only, no atomics",
                        "distinct banks (lane = bank), random rows",
                        "uniformly random words (independent per lane)",
                        "uniformly random words, return value used",
                        "same address in all 32 lanes",
                        "same bank, 32 different addresses",
                        "2-way: lanes l and l+16 same bank, different rows",
                        "16 active lanes, random words",
                        "hist.cu pattern: packed u16 + XOR swizzle + return + OR-check",
                        "hist.cu pattern + 2 LDS.128 of staged pixels per 16 atomics",
                        "uniformly random words, red.shared (no return)",
                        "4-way: lanes l, l+8, l+16, l+24 same bank, different rows"};

__device__ __forceinline__ uint32_t hash32(uint32_t x) {  // lowbias32
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_atoms(uint32_t* out, long long* cycles) {
  extern __shared__ __align__(16) uint32_t h[];  // kWords of histogram + 16 KiB of "staged pixels"
  uint32_t* stage = h + kWords;
  for (int i = threadIdx.x; i < kWords + 4096; i += blockDim.x) h[i] = i * 2654435761u;
  for (int i = threadIdx.x; i < kWords; i += blockDim.x) h[i] = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(h);
  uint32_t acc = 0;
  const long long t0 = clock64();
  if (MODE == HIST_PATTERN || MODE == HIST_PATTERN_LDS) {
    // 16 pixels per thread and step, as accum_fast<P_U16G, SWZ>: t = a << 8 | b
    // pixels from a per-thread multiplicative-congruential stream (independently hashed seeds, one IMAD
    // per word of four pixels) so that generating them costs 0.5 issue slots per pixel, not 4.5
    uint32_t x = hash32(gtid) | 1u;
    for (int it = 0; it < kIters / 16; it++) {
      uint32_t r[4], w[4];
      if (MODE == HIST_PATTERN_LDS) {
        const uint4 rv = *reinterpret_cast<const uint4*>(stage + ((threadIdx.x * 4 + it * 64) & 2047));
        const uint4 wv = *reinterpret_cast<const uint4*>(stage + 2048 + ((threadIdx.x * 4 + it * 64) & 2047));
        r[0] = rv.x; r[1] = rv.y; r[2] = rv.z; r[3] = rv.w;
        w[0] = wv.x; w[1] = wv.y; w[2] = wv.z; w[3] = wv.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++) r[j] = w[j] = 0;
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
        x = x * 0x9E3779B1u + 0x7F4A7C15u;
        r[j] ^= x;
        x = x * 0x9E3779B1u + 0x7F4A7C15u;
        w[j] ^= __byte_perm(x, 0u, 0x0123);  // the well-mixed high bytes into every byte position
      }
      uint32_t orv = 0, ws[4];
#pragma unroll
      for (int j = 0; j < 4; j++) ws[j] = w[j] ^ ((r[j] << 1) & 0xFEFEFEFEu);  // swizzle on four pixels at once
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const uint32_t t = __byte_perm(ws[i >> 2], r[i >> 2], 0x4440 + (i & 3) * 0x11);
        uint32_t addr, inc;
        asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(addr) : "r"(t & 0xFFFEu), "r"(base));
        asm("mad.lo.u32 %0, %1, 0xFFFF, 1;" : "=r"(inc) : "r"(t & 1u));
        uint32_t old;
        asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(inc) : "memory");
        orv |= old + inc;
      }
      if (orv & 0xF000F000u) acc++;
    }
  } else {
    // One IMAD per index (a multiplicative-congruential stream per thread, seeds hashed independently per
    // lane; the index is taken from the well-mixed high bits), so that the loop is bound by the ATOMS
    // pipe and not by generating addresses: with a full hash per index (first version of this file)
    // every mode but the serialised one sat at the issue limit the HASH_ONLY control shows.
    uint32_t x = hash32(gtid) | 1u;
#pragma unroll 8
    for (int it = 0; it < kIters; it++) {
      x = x * 0x9E3779B1u + 0x7F4A7C15u;
      const uint32_t r = MODE == HASH_ONLY ? hash32(x) : x;
      if (MODE == HASH_ONLY) {
        acc ^= r;
      } else if (MODE == DISTINCT_BANKS) {
        atomicAdd(&h[((r >> 22) << 5) + lane], 1u);
      } else if (MODE == RANDOM) {
        atomicAdd(&h[r >> 17], 1u);
      } else if (MODE == RANDOM_RED) {
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(base + ((r >> 17) << 2)) : "memory");
      } else if (MODE == RANDOM_RET) {
        acc |= atomicAdd(&h[r >> 17], 1u);
      } else if (MODE == SAME_ADDR) {
        atomicAdd(&h[(it * 37) & (kWords - 1)], 1u);   // warp-uniform address, no shuffle needed
      } else if (MODE == SAME_BANK_DIFF_ADDR) {
        atomicAdd(&h[(r >> 22) << 5], 1u);
      } else if (MODE == TWO_WAY) {
        atomicAdd(&h[((r >> 22) << 5) + (lane & 15u)], 1u);
      } else if (MODE == FOUR_WAY) {
        atomicAdd(&h[((r >> 22) << 5) + (lane & 7u)], 1u);
      } else if (MODE == HALF_WARP_RANDOM) {
        if (lane < 16) atomicAdd(&h[r >> 17], 1u);
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  uint32_t s = acc;
  for (int i = threadIdx.x; i < kWords; i += blockDim.x) s += h[i];
  if (s == 0xdeadbeef) out[0] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static double g_lanes[NMODES];

template <int MODE>
void run(int threads, int sms, uint32_t* d_out, long long* d_cyc) {
  auto k = k_atoms<MODE>;
  const size_t smem = (size_t)(kWords + 4096) * 4;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<sms, threads, smem>>>(d_out, d_cyc);  // warm
  cudaEventRecord(e0);
  k<<<sms, threads, smem>>>(d_out, d_cyc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long* hc = (long long*)malloc(sizeof(long long) * sms);
  cudaMemcpy(hc, d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double cyc = 0; for (int i = 0; i < sms; i++) cyc += (double)hc[i]; cyc /= sms;
  free(hc);
  const double active = MODE == HALF_WARP_RANDOM ? 0.5 : 1.0;
  const double total = (double)sms * threads * kIters * active;
  const double lanes = (double)threads * kIters * active / cyc;
  g_lanes[MODE] = lanes;
  printf("%-66s thr=%4d %8.1f Gop/s %6.2f lanes/clk/SM %6.2f clk per 32-lane instr  %7.3f ms  %s\n", kNames[MODE], threads,
         total / ms / 1e6, lanes, 32.0 * active / lanes, ms, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s  SMs=%d  smem/SM=%zu\n", p.name, p.multiProcessorCount, p.sharedMemPerMultiprocessor);
  uint32_t* d_out; long long* d_cyc;
  cudaMalloc(&d_out, 4); cudaMalloc(&d_cyc, 8 * 4096);
  const int sms = p.multiProcessorCount;
  run<HASH_ONLY>(512, sms, d_out, d_cyc);
  run<DISTINCT_BANKS>(512, sms, d_out, d_cyc);
  run<DISTINCT_BANKS>(1024, sms, d_out, d_cyc);
  run<RANDOM>(512, sms, d_out, d_cyc);
  run<RANDOM>(1024, sms, d_out, d_cyc);
  run<RANDOM_RED>(512, sms, d_out, d_cyc);
  run<RANDOM_RET>(512, sms, d_out, d_cyc);
  run<SAME_ADDR>(512, sms, d_out, d_cyc);
  run<SAME_BANK_DIFF_ADDR>(512, sms, d_out, d_cyc);
  run<TWO_WAY>(512, sms, d_out, d_cyc);
  run<FOUR_WAY>(512, sms, d_out, d_cyc);
  run<HALF_WARP_RANDOM>(512, sms, d_out, d_cyc);
  run<HIST_PATTERN>(512, sms, d_out, d_cyc);
  run<HIST_PATTERN>(1024, sms, d_out, d_cyc);
  run<HIST_PATTERN_LDS>(512, sms, d_out, d_cyc);
  // one machine-readable line for bench.py / profiles
  printf("JSON {\"sms\": %d, \"conflict_free_lanes_per_clk_sm\": %.3f, \"random_lanes_per_clk_sm\": %.3f, "
         "\"hist_pattern_lanes_per_clk_sm\": %.3f, \"hist_pattern_lds_lanes_per_clk_sm\": %.3f}\n",
         sms, g_lanes[DISTINCT_BANKS], g_lanes[RANDOM], g_lanes[HIST_PATTERN], g_lanes[HIST_PATTERN_LDS]);
  return 0;
}
