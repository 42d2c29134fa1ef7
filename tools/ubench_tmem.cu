// ubench_tmem.cu -- discovery run for staging pixel data through tensor memory:
// shared memory -> tcgen05.cp.128x256b -> TMEM -> tcgen05.ld.32x32b -> registers, no LSU wavefronts for
// the read.  Fills 4 KiB of shared memory with word index i at word i, copies it with two descriptor
// settings, reads it back per warp and prints which shared-memory words every thread received, plus
// a timing loop (cp + ld of 4 KiB per iteration on one CTA).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_tmem tools/ubench_tmem.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
  return d;                // layout type 0: no swizzle
}

__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done && spins < (1u << 22); spins++)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done != 0;
}

__global__ void __launch_bounds__(128, 1) k_tmem(uint32_t* out, long long* cycles, int iters) {
  __shared__ __align__(1024) uint32_t buf[1024];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tbase;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 1024; i += 128) buf[i] = i;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tbase)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes of buf -> async-proxy reads
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = tbase;
  uint32_t parity = 0;
  bool ok = true;
  for (int cfg = 0; cfg < 2; cfg++) {
    if (tid == 0) {
      const uint64_t desc = cfg == 0 ? make_desc(smem_u32(buf), 128, 256) : make_desc(smem_u32(buf), 2048, 128);
      asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr + 8u * cfg), "l"(desc) : "memory");
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    ok = ok && mbar_wait_bounded(&bar, parity);
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr + 8u * cfg + ((uint32_t)(32 * warp) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; i++) out[(cfg * 128 + tid) * 8 + i] = ok ? r[i] : 0xDEADBEEFu;
  }
  // timing: cp (4 KiB) + ld per iteration
  __syncthreads();
  long long t0 = clock64();
  uint32_t acc = 0;
  for (int it = 0; it < iters && ok; it++) {
    if (tid == 0) {
      const uint64_t desc = make_desc(smem_u32(buf), 128, 256);
      asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    ok = ok && mbar_wait_bounded(&bar, parity);
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr + ((uint32_t)(32 * warp) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; i++) acc += r[i];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // everyone has read before the next copy overwrites
  }
  long long t1 = clock64();
  if (tid == 0) { cycles[0] = t1 - t0; cycles[1] = ok ? 1 : 0; }
  if (acc == 0xFFFFFFFFu) out[0] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(taddr) : "memory");
}

int main() {
  uint32_t* d_out; long long* d_cyc;
  cudaMalloc(&d_out, 2 * 128 * 8 * 4); cudaMalloc(&d_cyc, 16);
  cudaMemset(d_out, 0xFF, 2 * 128 * 8 * 4);
  const int iters = 2000;
  k_tmem<<<1, 128>>>(d_out, d_cyc, iters);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  static uint32_t h[2 * 128 * 8]; long long c[2];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(c, d_cyc, 16, cudaMemcpyDeviceToHost);
  for (int cfg = 0; cfg < 2; cfg++) {
    printf("config %d (%s): smem WORD index received by thread t, register i\n", cfg, cfg == 0 ? "LBO 128 SBO 256" : "LBO 2048 SBO 128");
    for (int t : {0, 1, 2, 7, 8, 9, 31, 32, 33, 64, 127}) {
      printf("  t=%3d:", t);
      for (int i = 0; i < 8; i++) printf(" %4u", h[(cfg * 128 + t) * 8 + i]);
      printf("\n");
    }
    // is it a bijection over the 1024 words?
    static int seen[1024]; for (int i = 0; i < 1024; i++) seen[i] = 0; int bad = 0;
    for (int j = 0; j < 1024; j++) { uint32_t v = h[cfg * 1024 + j]; if (v < 1024) seen[v]++; else bad++; }
    int missing = 0; for (int i = 0; i < 1024; i++) if (seen[i] != 1) missing++;
    printf("  out-of-range values %d, words not received exactly once %d\n", bad, missing);
  }
  printf("timing: %d iterations of cp(4 KiB)+commit+wait+ld(x8)+barrier: %.1f cycles each, barrier waits %s\n", iters,
         (double)c[0] / iters, c[1] ? "all completed" : "TIMED OUT");
  return 0;
}
