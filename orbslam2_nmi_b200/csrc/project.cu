// project.cu -- batched point-cloud projection + z-buffer + resolve (sm_100a).
//
// Replaces the reference's OpenGL point render, one glDrawArrays(GL_POINTS) per
// synthetic view (Thirdparty/Localization/rendering.hpp:530-587, projection
// matrix :196-202, depth test / point size :294-307, clear colour :533).
// Arithmetic follows SURVEY.md Appendix A.2/A.3 and is written with explicit
// round-to-nearest intrinsics so it is bit-identical to oracle/nmi_oracle.c.
//
// Three kernels:
//   cull_*         two passes over the float4 cloud (coalesced 16 B loads): drop the
//                  points outside the union of all view frusta (conservative) and
//                  compact the survivors {xyz, original index} IN ORDER (count, scan,
//                  scatter), so the Morton order of the cloud survives;
//   project_splat  runs of Morton-neighbouring survivors x the views of a group: fp32
//                  projection, s x s splat resolved in a shared-memory tile, then one
//                  packed (depth bits << 32 | point index) atomicMin per touched cell
//                  into the per-view z-buffer, with a plain-load early-z test in front;
//   resolve        z-buffer -> u8 render (background 255) and reset to ~0.
#include <climits>

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

// Conservative frustum-union test (slack covers the fp32 rounding of the exact per-view test).
__device__ __forceinline__ bool cull_keep(const float4& p, const ViewConst& vc, const CullConst& cc) {
  const float dx = p.x - cc.c0[0], dy = p.y - cc.c0[1], dz = p.z - cc.c0[2];
  const float X = vc.r0[0] * dx + vc.r0[1] * dy + vc.r0[2] * dz;
  const float Y = vc.r1[0] * dx + vc.r1[1] * dy + vc.r1[2] * dz;
  const float Z = vc.r2[0] * dx + vc.r2[1] * dy + vc.r2[2] * dz;
  const float sl = 0.05f + 0.01f * (fabsf(X) + fabsf(Y) + fabsf(Z));  // 1 % + 5 cm
  const float zmax = Z + cc.mz + sl;
  return (Z >= vc.zn - cc.mz - sl) && (Z <= vc.zf + cc.mz + sl) &&
         (vc.kx * (fabsf(X) - cc.mx - sl) <= zmax) && (vc.ky * (fabsf(Y) - cc.my - sl) <= zmax);
}

// Order-preserving (stable) compaction in three small kernels, so that the survivors keep
// the Morton order of the cloud and a run of consecutive survivors stays a compact image
// patch:  count per CTA -> exclusive scan of the CTA counts -> scatter.
constexpr int kCullThreads = 256;

__global__ void __launch_bounds__(kCullThreads)
cull_count_kernel(const float4* __restrict__ pts, uint32_t n, ViewConst vc, CullConst cc,
                  uint32_t* __restrict__ block_counts) {
  const uint32_t i = blockIdx.x * kCullThreads + threadIdx.x;
  const bool keep = i < n && cull_keep(ldg_stream(pts + i), vc, cc);
  const int cnt = __syncthreads_count(keep);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = (uint32_t)cnt;
}

// single CTA: in-place exclusive scan of nblocks counts, total -> *counter
__global__ void __launch_bounds__(1024)
cull_scan_kernel(uint32_t* __restrict__ block_counts, uint32_t nblocks, uint32_t* __restrict__ counter) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nblocks; base += 1024) {
    const uint32_t i = base + tid;
    const uint32_t v = i < nblocks ? block_counts[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_warp[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += y;
      }
      s_warp[lane] = w;  // inclusive scan of the warp totals
    }
    __syncthreads();
    const uint32_t carry = s_carry;
    const uint32_t incl = x + (warp ? s_warp[warp - 1] : 0u);
    if (i < nblocks) block_counts[i] = carry + incl - v;  // exclusive
    __syncthreads();
    if (tid == 1023) s_carry = carry + incl;
    __syncthreads();
  }
  if (tid == 0) *counter = s_carry;
}

__global__ void __launch_bounds__(kCullThreads)
cull_scatter_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ orig, uint32_t n,
                    ViewConst vc, CullConst cc, const uint32_t* __restrict__ block_offsets,
                    float4* __restrict__ out_pts, uint32_t* __restrict__ out_idx) {
  __shared__ uint32_t s_warp[kCullThreads / 32];
  const uint32_t i = blockIdx.x * kCullThreads + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 p = make_float4(0, 0, 0, 0);
  bool keep = false;
  if (i < n) {
    p = ldg_stream(pts + i);
    keep = cull_keep(p, vc, cc);
  }
  const unsigned m = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) s_warp[warp] = (uint32_t)__popc(m);
  __syncthreads();
  uint32_t before = block_offsets[blockIdx.x];
  for (int w = 0; w < warp; w++) before += s_warp[w];
  if (keep) {
    const uint32_t o = before + __popc(m & ((1u << lane) - 1u));
    out_pts[o] = p;
    out_idx[o] = orig[i];
  }
}

// ---- project_splat: tile-binned splat -------------------------------------------------
// One CTA = a run of kRun consecutive (Morton-ordered) survivors x the views of the group.
// Neighbouring points land in a small image patch, so for each view the CTA
//   1. projects its points (fp32, bit-identical to the oracle) and reduces their pixel
//      bounding box,
//   2. resolves all of the run's fragments inside a SHARED-MEMORY tile of that box with
//      native 32-bit atomics: pass 1 atomicMin on the depth bits, pass 2 atomicMin on the
//      original index among the fragments that hold the minimum depth (== the 64-bit
//      packed-key minimum, split in two),
//   3. flushes the touched cells to the global z-buffer with one early-z load + one packed
//      64-bit atomicMin per CELL (coalesced row segments) instead of one per FRAGMENT.
// A box larger than the tile is processed in horizontal strips; only a box needing more
// than kMaxStrips strips falls back to per-fragment global atomics.
constexpr int kSplatThreads = 256;
constexpr int kPtsPerThread = 4;
constexpr int kRun = kSplatThreads * kPtsPerThread;  // 1024 survivors per CTA
constexpr int kTileCap = 8192;                       // cells of the shared-memory tile
constexpr int kMaxStrips = 16;                       // box height / strip height before falling back

struct Frag {  // one projected point of one view
  int i0, j0;
  uint32_t zbits;  // 0xFFFFFFFF = clipped
};

__device__ __forceinline__ Frag project_point(const float4& p, const float4& c, const ViewConst& vc,
                                              float half) {
  Frag f;
  f.i0 = f.j0 = 0;
  f.zbits = 0xFFFFFFFFu;
  const float dx = __fsub_rn(p.x, c.x), dy = __fsub_rn(p.y, c.y), dz = __fsub_rn(p.z, c.z);
  const float Zc = __fmaf_rn(vc.r2[2], dz, __fmaf_rn(vc.r2[1], dy, __fmul_rn(vc.r2[0], dx)));
  if (!(Zc >= vc.zn && Zc <= vc.zf)) return f;
  const float Xc = __fmaf_rn(vc.r0[2], dz, __fmaf_rn(vc.r0[1], dy, __fmul_rn(vc.r0[0], dx)));
  const float Yc = __fmaf_rn(vc.r1[2], dz, __fmaf_rn(vc.r1[1], dy, __fmul_rn(vc.r1[0], dx)));
  const float nx = __fdiv_rn(__fmul_rn(vc.kx, Xc), Zc);
  const float ny = __fdiv_rn(__fmul_rn(vc.ky, Yc), Zc);
  if (!(fabsf(nx) <= 1.0f && fabsf(ny) <= 1.0f)) return f;
  const float xw = __fmaf_rn(nx, vc.hw, vc.hw);
  const float yr = __fmaf_rn(ny, vc.hh, vc.hh);
  f.i0 = (int)floorf(__fsub_rn(xw, half));
  f.j0 = (int)floorf(__fsub_rn(yr, half));
  f.zbits = __float_as_uint(Zc);
  return f;
}

__global__ void __launch_bounds__(kSplatThreads)
project_splat_kernel(const float4* __restrict__ cpts, const uint32_t* __restrict__ cidx,
                     const uint32_t* __restrict__ counter, const float4* __restrict__ centres,
                     int nviews, ViewConst vc, unsigned long long* __restrict__ zbuf, size_t P) {
  extern __shared__ uint32_t s_tile[];  // depth[kTileCap] | index[kTileCap]
  uint32_t* s_depth = s_tile;
  uint32_t* s_index = s_tile + kTileCap;
  __shared__ int s_bb[4];  // xmin, xmax, ymin, ymax of the run's valid points (i0 / j0)

  const uint32_t count = *counter;
  const uint32_t base = blockIdx.x * (uint32_t)kRun;
  if (base >= count) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float half = 0.5f * (float)(vc.s - 1);
  const int S = vc.s;

  float4 pt[kPtsPerThread];
  uint32_t oi[kPtsPerThread];
#pragma unroll
  for (int k = 0; k < kPtsPerThread; k++) {
    const uint32_t t = base + k * kSplatThreads + tid;
    const bool in = t < count;
    pt[k] = in ? cpts[t] : make_float4(0.f, 0.f, 0.f, 0.f);
    oi[k] = in ? cidx[t] : 0xFFFFFFFFu;  // 0xFFFFFFFF marks a padding slot
  }
  if (tid == 0) { s_bb[0] = INT_MAX; s_bb[1] = INT_MIN; s_bb[2] = INT_MAX; s_bb[3] = INT_MIN; }
  __syncthreads();

  for (int v = 0; v < nviews; v++) {
    const float4 c = centres[v];
    unsigned long long* zb = zbuf + (size_t)v * P;
    Frag f[kPtsPerThread];
    int xmin = INT_MAX, xmax = INT_MIN, ymin = INT_MAX, ymax = INT_MIN;
#pragma unroll
    for (int k = 0; k < kPtsPerThread; k++) {
      f[k] = project_point(pt[k], c, vc, half);
      if (oi[k] == 0xFFFFFFFFu) f[k].zbits = 0xFFFFFFFFu;
      if (f[k].zbits != 0xFFFFFFFFu) {
        xmin = min(xmin, f[k].i0); xmax = max(xmax, f[k].i0);
        ymin = min(ymin, f[k].j0); ymax = max(ymax, f[k].j0);
      }
    }
    xmin = __reduce_min_sync(0xffffffffu, xmin); xmax = __reduce_max_sync(0xffffffffu, xmax);
    ymin = __reduce_min_sync(0xffffffffu, ymin); ymax = __reduce_max_sync(0xffffffffu, ymax);
    if (lane == 0 && xmin <= xmax) {
      atomicMin(&s_bb[0], xmin); atomicMax(&s_bb[1], xmax);
      atomicMin(&s_bb[2], ymin); atomicMax(&s_bb[3], ymax);
    }
    __syncthreads();  // (1) box complete
    const int bx0 = s_bb[0], bx1 = s_bb[1], by0 = s_bb[2], by1 = s_bb[3];
    if (bx0 > bx1) {  // nothing of this run is visible in this view (uniform)
      __syncthreads();
      continue;
    }
    const int x0 = max(bx0, 0), x1 = min(bx1 + S, vc.W);  // clipped cell box [x0,x1) x [Y0,Y1)
    const int Y0 = max(by0, 0), Y1 = min(by1 + S, vc.H);
    const int tw = x1 - x0, TH = Y1 - Y0;
    // the box is processed in horizontal strips of at most kTileCap cells
    const int strip_rows = tw > 0 ? kTileCap / tw : 0;
    const int nstrips = (tw > 0 && TH > 0 && strip_rows > 0) ? (TH + strip_rows - 1) / strip_rows : 0;
    const bool tiled = nstrips > 0 && nstrips <= kMaxStrips;
    __syncthreads();  // (2) everybody has read the box
    if (tid == 0) { s_bb[0] = INT_MAX; s_bb[1] = INT_MIN; s_bb[2] = INT_MAX; s_bb[3] = INT_MIN; }
    if (tiled) {
      for (int sidx = 0; sidx < nstrips; sidx++) {
        const int y0 = Y0 + sidx * strip_rows, y1 = min(y0 + strip_rows, Y1);
        const int th = y1 - y0, ncell = tw * th;
        for (int q = tid; q < ncell; q += kSplatThreads) { s_depth[q] = 0xFFFFFFFFu; s_index[q] = 0xFFFFFFFFu; }
        __syncthreads();
        // pass 1: minimum depth per cell
#pragma unroll
        for (int k = 0; k < kPtsPerThread; k++) {
          if (f[k].zbits == 0xFFFFFFFFu) continue;
          for (int j = max(f[k].j0, y0); j < min(f[k].j0 + S, y1); j++)
            for (int i = max(f[k].i0, x0); i < min(f[k].i0 + S, x1); i++)
              atomicMin(&s_depth[(j - y0) * tw + (i - x0)], f[k].zbits);
        }
        __syncthreads();
        // pass 2: lowest original index among the fragments at the minimum depth
#pragma unroll
        for (int k = 0; k < kPtsPerThread; k++) {
          if (f[k].zbits == 0xFFFFFFFFu) continue;
          for (int j = max(f[k].j0, y0); j < min(f[k].j0 + S, y1); j++)
            for (int i = max(f[k].i0, x0); i < min(f[k].i0 + S, x1); i++) {
              const int q = (j - y0) * tw + (i - x0);
              if (s_depth[q] == f[k].zbits) atomicMin(&s_index[q], oi[k]);
            }
        }
        __syncthreads();
        // flush: one early-z load + at most one 64-bit atomicMin per touched cell
        for (int r = warp; r < th; r += kSplatThreads / 32) {
          unsigned long long* row = zb + (size_t)(y0 + r) * vc.W + x0;
          for (int x = lane; x < tw; x += 32) {
            const uint32_t d = s_depth[r * tw + x];
            if (d != 0xFFFFFFFFu) {
              const unsigned long long key = ((unsigned long long)d << 32) | s_index[r * tw + x];
              if (key < row[x]) atomicMin(row + x, key);
            }
          }
        }
        __syncthreads();  // strip done: the tile may be cleared again
      }
    } else {
      // box far larger than the tile (degenerate view): per-fragment global path
#pragma unroll
      for (int k = 0; k < kPtsPerThread; k++) {
        if (f[k].zbits == 0xFFFFFFFFu) continue;
        const unsigned long long key = ((unsigned long long)f[k].zbits << 32) | oi[k];
        for (int j = f[k].j0; j < f[k].j0 + S; j++) {
          if (j < 0 || j >= vc.H) continue;
          for (int i = f[k].i0; i < f[k].i0 + S; i++) {
            if (i < 0 || i >= vc.W) continue;
            unsigned long long* cell = zb + (size_t)j * vc.W + i;
            if (key < *cell) atomicMin(cell, key);  // stale reads are conservative
          }
        }
      }
    }
    __syncthreads();  // (5) tile free again, s_bb reset visible
  }
}

__global__ void __launch_bounds__(256)
resolve_kernel(unsigned long long* __restrict__ zbuf, const uint8_t* __restrict__ val, size_t P,
               uint8_t* __restrict__ images, size_t pitch, uint32_t* __restrict__ winners) {
  const int v = blockIdx.y;
  unsigned long long* zb = zbuf + (size_t)v * P;
  uint8_t* img = images + (size_t)v * pitch;
  // 4 pixels per thread: two 16-byte key loads, two 16-byte resets, one 4-byte render store
  const size_t q = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (q >= P) return;
  unsigned long long key[4];
  const bool full = q + 3 < P && (((size_t)v * P + q) % 2 == 0);  // 16 B aligned quad
  if (full) {
    const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(zb + q);
    const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(zb + q + 2);
    key[0] = a.x; key[1] = a.y; key[2] = b.x; key[3] = b.y;
    const ulonglong2 ones = make_ulonglong2(~0ull, ~0ull);
    *reinterpret_cast<ulonglong2*>(zb + q) = ones;
    *reinterpret_cast<ulonglong2*>(zb + q + 2) = ones;
  } else {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      key[k] = ~0ull;
      if (q + k < P) { key[k] = zb[q + k]; zb[q + k] = ~0ull; }
    }
  }
  uint32_t packed = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint32_t w = key[k] == ~0ull ? NMI_EMPTY : (uint32_t)(key[k] & 0xFFFFFFFFull);
    const uint32_t pix = w == NMI_EMPTY ? 255u : (uint32_t)__ldg(val + w);
    if (winners && q + k < P) winners[(size_t)v * P + q + k] = w;
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
                         uint32_t* out_idx, uint32_t* counter, uint32_t* block_counts,
                         cudaStream_t st) {
  if (n == 0) return;
  CullConst cc;
  for (int i = 0; i < 3; i++) cc.c0[i] = c0[i];
  cc.mx = margin[0];
  cc.my = margin[1];
  cc.mz = margin[2];
  const uint32_t nblocks = (n + kCullThreads - 1) / kCullThreads;
  cull_count_kernel<<<nblocks, kCullThreads, 0, st>>>(pts, n, vc, cc, block_counts);
  cull_scan_kernel<<<1, 1024, 0, st>>>(block_counts, nblocks, counter);
  cull_scatter_kernel<<<nblocks, kCullThreads, 0, st>>>(pts, orig, n, vc, cc, block_counts, out_pts,
                                                        out_idx);
}

void launch_project_splat(const float4* cpts, const uint32_t* cidx, const uint32_t* counter,
                          const float4* centres, int nviews, const ViewConst& vc,
                          unsigned long long* zbuf, size_t P, uint32_t max_points, cudaStream_t st) {
  if (nviews == 0 || max_points == 0) return;
  static bool configured = false;
  const size_t smem = sizeof(uint32_t) * 2 * kTileCap;
  if (!configured) {
    cudaFuncSetAttribute(project_splat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = true;
  }
  // one CTA per run of kRun survivors; the survivor count lives on the device, so launch for
  // the upper bound (all points) and let the surplus CTAs exit at once
  const unsigned grid = (max_points + kRun - 1) / kRun;
  project_splat_kernel<<<grid, kSplatThreads, smem, st>>>(cpts, cidx, counter, centres, nviews, vc,
                                                          zbuf, P);
}

void launch_resolve(unsigned long long* zbuf, const uint8_t* val, int nviews, size_t P,
                    uint8_t* images, size_t pitch, uint32_t* winners, cudaStream_t st) {
  if (nviews == 0 || P == 0) return;
  dim3 grid((unsigned)((P + 1023) / 1024), (unsigned)nviews);
  resolve_kernel<<<grid, 256, 0, st>>>(zbuf, val, P, images, pitch, winners);
}

}  // namespace nmi
