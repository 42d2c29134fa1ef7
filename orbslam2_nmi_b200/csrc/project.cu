// project.cu -- batched point-cloud projection + z-buffer + resolve (sm_100a).
//
// Replaces the reference's OpenGL point render, one glDrawArrays(GL_POINTS) per
// synthetic view (Thirdparty/Localization/rendering.hpp:530-587, projection
// matrix :196-202, depth test / point size :294-307, clear colour :533).
// Arithmetic follows SURVEY.md Appendix A.2/A.3 and is written with explicit
// round-to-nearest intrinsics so it is bit-identical to oracle/nmi_oracle.c.
//
// Three kernels:
//   cull_compact   one pass over the float4 cloud (coalesced 16 B loads): drops
//                  points outside the union of all view frusta (conservative),
//                  appends survivors {xyz, original index};
//   project_splat  survivors x all views: fp32 projection, s x s splat, packed
//                  (depth bits << 32 | point index) atomicMin into the per-view
//                  z-buffer, with a plain-load early-z test in front;
//   resolve        z-buffer -> u8 render (background 255) and reset to ~0.
#include "nmi_internal.h"

namespace nmi {

namespace {

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__global__ void fill_u64_kernel(unsigned long long* p, size_t n, unsigned long long v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

// value = floor(255 I + 0.5) clamped (objloader.cpp:261 stores I = red/256;
// GL float->unorm8 conversion, tie rule ours)
__global__ void intensity_u8_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ orig,
                                    uint8_t* __restrict__ val, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float f = floorf(__fadd_rn(__fmul_rn(255.0f, pts[i].w), 0.5f));
  if (!(f >= 0.0f)) f = 0.0f;
  if (f > 255.0f) f = 255.0f;
  val[orig[i]] = (uint8_t)f;  // looked up by the original index the z-buffer key carries
}

struct CullConst {
  float c0[3];       // central camera centre (Twc translation)
  float mx, my, mz;  // max |camera-frame offset| of any view in the batch
};

// Conservative frustum-union test, then warp-aggregated append.
__global__ void __launch_bounds__(256)
cull_compact_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ orig, uint32_t n,
                    ViewConst vc, CullConst cc, float4* __restrict__ out_pts,
                    uint32_t* __restrict__ out_idx, uint32_t* __restrict__ counter) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool keep = false;
  float4 p = make_float4(0, 0, 0, 0);
  if (i < n) {
    p = ldg_stream(pts + i);
    float dx = p.x - cc.c0[0], dy = p.y - cc.c0[1], dz = p.z - cc.c0[2];
    float X = vc.r0[0] * dx + vc.r0[1] * dy + vc.r0[2] * dz;
    float Y = vc.r1[0] * dx + vc.r1[1] * dy + vc.r1[2] * dz;
    float Z = vc.r2[0] * dx + vc.r2[1] * dy + vc.r2[2] * dz;
    // slack covers fp32 rounding of the exact per-view test (1 % + 5 cm)
    float sl = 0.05f + 0.01f * (fabsf(X) + fabsf(Y) + fabsf(Z));
    float zmax = Z + cc.mz + sl;
    keep = (Z >= vc.zn - cc.mz - sl) && (Z <= vc.zf + cc.mz + sl) &&
           (vc.kx * (fabsf(X) - cc.mx - sl) <= zmax) &&
           (vc.ky * (fabsf(Y) - cc.my - sl) <= zmax);
  }
  unsigned m = __ballot_sync(0xffffffffu, keep);
  if (m == 0) return;
  int lane = threadIdx.x & 31;
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(counter, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (keep) {
    uint32_t o = base + __popc(m & ((1u << lane) - 1u));
    out_pts[o] = p;
    out_idx[o] = orig[i];
  }
}

// One thread per surviving point, looping over the views of the batch.
__global__ void __launch_bounds__(256)
project_splat_kernel(const float4* __restrict__ cpts, const uint32_t* __restrict__ cidx,
                     const uint32_t* __restrict__ counter, const float4* __restrict__ centres,
                     int nviews, ViewConst vc, unsigned long long* __restrict__ zbuf, size_t P) {
  extern __shared__ float4 s_c[];
  for (int i = threadIdx.x; i < nviews; i += blockDim.x) s_c[i] = centres[i];
  __syncthreads();
  const uint32_t count = *counter;
  const float half = 0.5f * (float)(vc.s - 1);
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < count;
       t += gridDim.x * blockDim.x) {
    const float4 p = cpts[t];
    const unsigned long long lo = cidx[t];
    for (int v = 0; v < nviews; v++) {
      const float4 c = s_c[v];
      const float dx = __fsub_rn(p.x, c.x), dy = __fsub_rn(p.y, c.y), dz = __fsub_rn(p.z, c.z);
      const float Zc =
          __fmaf_rn(vc.r2[2], dz, __fmaf_rn(vc.r2[1], dy, __fmul_rn(vc.r2[0], dx)));
      if (!(Zc >= vc.zn && Zc <= vc.zf)) continue;
      const float Xc =
          __fmaf_rn(vc.r0[2], dz, __fmaf_rn(vc.r0[1], dy, __fmul_rn(vc.r0[0], dx)));
      const float Yc =
          __fmaf_rn(vc.r1[2], dz, __fmaf_rn(vc.r1[1], dy, __fmul_rn(vc.r1[0], dx)));
      const float nx = __fdiv_rn(__fmul_rn(vc.kx, Xc), Zc);
      const float ny = __fdiv_rn(__fmul_rn(vc.ky, Yc), Zc);
      if (!(fabsf(nx) <= 1.0f && fabsf(ny) <= 1.0f)) continue;
      const float xw = __fmaf_rn(nx, vc.hw, vc.hw);
      const float yr = __fmaf_rn(ny, vc.hh, vc.hh);
      const int i0 = (int)floorf(__fsub_rn(xw, half));
      const int j0 = (int)floorf(__fsub_rn(yr, half));
      const unsigned long long key = ((unsigned long long)__float_as_uint(Zc) << 32) | lo;
      unsigned long long* zb = zbuf + (size_t)v * P;
      for (int j = j0; j < j0 + vc.s; j++) {
        if (j < 0 || j >= vc.H) continue;
        for (int ii = i0; ii < i0 + vc.s; ii++) {
          if (ii < 0 || ii >= vc.W) continue;
          unsigned long long* cell = zb + (size_t)j * vc.W + ii;
          // early-z: the cell only ever decreases, so a stale read is conservative
          if (key < *cell) atomicMin(cell, key);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256)
resolve_kernel(unsigned long long* __restrict__ zbuf, const uint8_t* __restrict__ val, size_t P,
               uint8_t* __restrict__ images, size_t pitch, uint32_t* __restrict__ winners) {
  const int v = blockIdx.y;
  unsigned long long* zb = zbuf + (size_t)v * P;
  uint8_t* img = images + (size_t)v * pitch;
  // 4 pixels per thread -> one 32-bit store of the render
  size_t q = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (q >= P) return;
  uint32_t packed = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    size_t p = q + k;
    uint32_t pix = 0;
    if (p < P) {
      unsigned long long key = zb[p];
      uint32_t w = key == ~0ull ? NMI_EMPTY : (uint32_t)(key & 0xFFFFFFFFull);
      pix = w == NMI_EMPTY ? 255u : (uint32_t)__ldg(val + w);
      if (winners) winners[(size_t)v * P + p] = w;
      zb[p] = ~0ull;
    }
    packed |= pix << (8 * k);
  }
  if (q + 3 < P && (pitch % 4 == 0)) {
    *reinterpret_cast<uint32_t*>(img + q) = packed;
  } else {
    for (int k = 0; k < 4 && q + k < P; k++) img[q + k] = (uint8_t)(packed >> (8 * k));
  }
}

}  // namespace

void launch_fill_u64(unsigned long long* p, size_t n, unsigned long long v, cudaStream_t st) {
  if (n == 0) return;
  fill_u64_kernel<<<148 * 8, 256, 0, st>>>(p, n, v);
}

void launch_intensity_u8(const float4* pts, const uint32_t* orig, uint8_t* val, size_t n,
                         cudaStream_t st) {
  if (n == 0) return;
  intensity_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pts, orig, val, n);
}

void launch_cull_compact(const float4* pts, const uint32_t* orig, uint32_t n, const ViewConst& vc,
                         const float c0[3], const float margin[3], float4* out_pts,
                         uint32_t* out_idx, uint32_t* counter, cudaStream_t st) {
  if (n == 0) return;
  CullConst cc;
  for (int i = 0; i < 3; i++) cc.c0[i] = c0[i];
  cc.mx = margin[0];
  cc.my = margin[1];
  cc.mz = margin[2];
  cull_compact_kernel<<<(n + 255) / 256, 256, 0, st>>>(pts, orig, n, vc, cc, out_pts, out_idx,
                                                       counter);
}

void launch_project_splat(const float4* cpts, const uint32_t* cidx, const uint32_t* counter,
                          const float4* centres, int nviews, const ViewConst& vc,
                          unsigned long long* zbuf, size_t P, cudaStream_t st) {
  if (nviews == 0) return;
  project_splat_kernel<<<148 * 16, 256, sizeof(float4) * nviews, st>>>(cpts, cidx, counter,
                                                                      centres, nviews, vc, zbuf, P);
}

void launch_resolve(unsigned long long* zbuf, const uint8_t* val, int nviews, size_t P,
                    uint8_t* images, size_t pitch, uint32_t* winners, cudaStream_t st) {
  if (nviews == 0 || P == 0) return;
  dim3 grid((unsigned)((P + 1023) / 1024), (unsigned)nviews);
  resolve_kernel<<<grid, 256, 0, st>>>(zbuf, val, P, images, pitch, winners);
}

}  // namespace nmi
