// project.cu -- batched point-cloud projection + z-buffer + resolve (sm_100a).
//
// Replaces the reference's OpenGL point render, one glDrawArrays(GL_POINTS) per
// synthetic view (Thirdparty/Localization/rendering.hpp:530-587, projection
// matrix :196-202, depth test / point size :294-307, clear colour :533).
// Arithmetic follows SURVEY.md Appendix A.2/A.3 and is written with explicit
// round-to-nearest intrinsics so it is bit-identical to oracle/nmi_oracle.c.
//
// Kernels:
//   cull_compact   one pass over the float4 cloud (coalesced 16 B loads; whole 1 024-point blocks
//                  are rejected by their load-time box without a load): drop the points outside
//                  the union of all view frusta (conservative) and compact the survivors
//                  {xyz, tie-break word}; blocks keep their Morton order inside, their order among
//                  each other is that of one atomic per block (cull_count / cull_scan /
//                  cull_scatter: the stable three-pass variant, $NMI_CULL_PASSES=3, and the scan
//                  the mesh cull and the two-pass binning share);
//   project_splat  survivors x the views of a group: fp32 projection, s x s splat, packed
//                  (depth bits << 32 | tie-break word) atomicMin into the per-view
//                  z-buffer, with a plain-load early-z test in front;
//   resolve        z-buffer -> u8 render (background 255) and reset to ~0.
// Point clouds go through the binned TILE renderer further down (bin_kernel / tile_resolve);
// the global z-buffer path above serves the mesh model's resolve and splats wider than a tile.
#include <climits>
#include <cstdlib>

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

// W x H u8 image with an arbitrary row pitch -> packed rows, optionally bottom-up -> top-down
// (an externally produced render: a GL texture's first row is the bottom one)
__global__ void copy_rows_kernel(const uint8_t* __restrict__ src, size_t src_pitch, uint8_t* __restrict__ dst,
                                 int W, int H, bool flip) {
  const int y = blockIdx.y;
  const uint8_t* s = src + (size_t)(flip ? H - 1 - y : y) * src_pitch;
  uint8_t* d = dst + (size_t)y * W;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < W; x += gridDim.x * blockDim.x) d[x] = s[x];
}

__global__ void fill_u64_kernel(unsigned long long* p, size_t n, unsigned long long v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

// value = floor(255 I + 0.5) clamped (objloader.cpp:261 stores I = red/256;
// GL float->unorm8 conversion, tie rule ours)
__global__ void intensity_u8_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ orig,
                                    uint8_t* __restrict__ val, uint32_t* __restrict__ tag, bool packed,
                                    size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float f = floorf(__fadd_rn(__fmul_rn(255.0f, pts[i].w), 0.5f));
  if (!(f >= 0.0f)) f = 0.0f;
  if (f > 255.0f) f = 255.0f;
  const uint32_t o = orig[i];
  val[o] = (uint8_t)f;  // looked up by the original index when the key cannot carry the value
  tag[i] = packed ? (o << 8) | (uint32_t)f : o;
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
// patch:  count per CTA -> exclusive scan of the CTA counts -> scatter.  (The default is the one-pass
// cull_compact_kernel below; this variant remains as $NMI_CULL_PASSES=3.)
constexpr int kCullThreads = 256;
constexpr int kCullPer = 4;                              // points per thread
static_assert(kCullThreads * kCullPer == kCullBlock, "one CTA per AABB block");

// Block cull.  The cloud is Morton-ordered, so the kCullBlock consecutive points of a CTA form a
// compact patch whose axis-aligned box was computed at load time (nmi_set_points).  A box that lies
// entirely beyond one of the (already conservatively enlarged) planes of the frustum union is
// skipped without reading a single point: at C2 ~85 % of the cloud.  The test uses the eight
// corners: every plane function is linear, the slack `sl` is convex, so the extreme over the
// box is attained at a corner; `eps` covers the fp32 rounding of the two evaluations, and both
// are far inside the 5 cm + 1 % slack that already separates cull_keep from the exact per-view
// clip test, so no visible point can be lost.  Boxes holding NaN / inf never reject.
__device__ __forceinline__ bool block_may_survive(const float* __restrict__ aabb, uint32_t b, const ViewConst& vc,
                                                  const CullConst& cc) {
  if (aabb == nullptr || !(vc.kx > 0.0f) || !(vc.ky > 0.0f)) return true;
  const float lo[3] = {__ldg(aabb + 6 * (size_t)b), __ldg(aabb + 6 * (size_t)b + 1), __ldg(aabb + 6 * (size_t)b + 2)};
  const float hi[3] = {__ldg(aabb + 6 * (size_t)b + 3), __ldg(aabb + 6 * (size_t)b + 4), __ldg(aabb + 6 * (size_t)b + 5)};
  float zmin = INFINITY, zmax = -INFINITY, l1 = 0.0f;
  float fxp = INFINITY, fxn = INFINITY, fyp = INFINITY, fyn = INFINITY;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const float dx = ((k & 1) ? hi[0] : lo[0]) - cc.c0[0];
    const float dy = ((k & 2) ? hi[1] : lo[1]) - cc.c0[1];
    const float dz = ((k & 4) ? hi[2] : lo[2]) - cc.c0[2];
    const float X = vc.r0[0] * dx + vc.r0[1] * dy + vc.r0[2] * dz;
    const float Y = vc.r1[0] * dx + vc.r1[1] * dy + vc.r1[2] * dz;
    const float Z = vc.r2[0] * dx + vc.r2[1] * dy + vc.r2[2] * dz;
    zmin = fminf(zmin, Z);
    zmax = fmaxf(zmax, Z);
    l1 = fmaxf(l1, fabsf(X) + fabsf(Y) + fabsf(Z));
    fxp = fminf(fxp, vc.kx * X - Z);
    fxn = fminf(fxn, -vc.kx * X - Z);
    fyp = fminf(fyp, vc.ky * Y - Z);
    fyn = fminf(fyn, -vc.ky * Y - Z);
  }
  const float sl = 0.05f + 0.01f * l1;
  const float eps = 1e-3f * (1.0f + l1);
  const float bx = vc.kx * (cc.mx + sl) + cc.mz + sl + eps * (1.0f + vc.kx);
  const float by = vc.ky * (cc.my + sl) + cc.mz + sl + eps * (1.0f + vc.ky);
  const bool out = (zmax < vc.zn - cc.mz - sl - eps) || (zmin > vc.zf + cc.mz + sl + eps) || (fxp > bx) ||
                   (fxn > bx) || (fyp > by) || (fyn > by);
  return !out;  // NaN anywhere: every comparison is false -> kept
}

__global__ void __launch_bounds__(kCullThreads)
cull_count_kernel(const float4* __restrict__ pts, uint32_t n, ViewConst vc, CullConst cc,
                  const float* __restrict__ aabb, uint32_t* __restrict__ block_counts) {
  __shared__ uint32_t s_cnt[kCullThreads / 32];
  const uint32_t base = blockIdx.x * kCullBlock;
  if (!block_may_survive(aabb, blockIdx.x, vc, cc)) {  // CTA-uniform
    if (threadIdx.x == 0) block_counts[blockIdx.x] = 0;
    return;
  }
  uint32_t mine = 0;
#pragma unroll
  for (int k = 0; k < kCullPer; k++) {
    const uint32_t i = base + k * kCullThreads + threadIdx.x;
    mine += (i < n && cull_keep(ldg_stream(pts + i), vc, cc)) ? 1u : 0u;
  }
#pragma unroll
  for (int d = 16; d >= 1; d /= 2) mine += __shfl_xor_sync(0xffffffffu, mine, d);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < kCullThreads / 32; w++) t += s_cnt[w];
    block_counts[blockIdx.x] = t;
  }
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
cull_scatter_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ tag, uint32_t n,
                    ViewConst vc, CullConst cc, const float* __restrict__ aabb,
                    const uint32_t* __restrict__ block_offsets,
                    float4* __restrict__ out_pts, uint32_t* __restrict__ out_idx) {
  __shared__ uint32_t s_warp[kCullPer][kCullThreads / 32];
  const uint32_t base = blockIdx.x * kCullBlock;
  if (!block_may_survive(aabb, blockIdx.x, vc, cc)) return;  // same decision as the count pass
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 p[kCullPer];
  unsigned m[kCullPer];
  bool keep[kCullPer];
#pragma unroll
  for (int k = 0; k < kCullPer; k++) {  // sub-block k holds points base + k*256 .. +255 (in order)
    const uint32_t i = base + k * kCullThreads + threadIdx.x;
    keep[k] = false;
    p[k] = make_float4(0, 0, 0, 0);
    if (i < n) {
      p[k] = ldg_stream(pts + i);
      keep[k] = cull_keep(p[k], vc, cc);
    }
    m[k] = __ballot_sync(0xffffffffu, keep[k]);
    if (lane == 0) s_warp[k][warp] = (uint32_t)__popc(m[k]);
  }
  __syncthreads();
  uint32_t before = block_offsets[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kCullPer; k++) {
    uint32_t off = before;
    for (int w = 0; w < warp; w++) off += s_warp[k][w];
    if (keep[k]) {
      const uint32_t o = off + __popc(m[k] & ((1u << lane) - 1u));
      out_pts[o] = p[k];
      out_idx[o] = tag[base + k * kCullThreads + threadIdx.x];  // the z-buffer key's tie-break word
    }
    for (int w = 0; w < kCullThreads / 32; w++) before += s_warp[k][w];
  }
}

// One-pass compaction: a CTA that survives the block test counts its survivors, takes its output range
// with ONE atomic on the running total and writes them in order.  The survivors of a block stay in Morton
// order and the blocks arrive roughly in launch order, which is all the locality the binning pass needs
// (the z-buffer minimum does not depend on the order of its candidates); one launch instead of
// count -> scan -> scatter, and the cloud is read once.  `counter` must be zero at launch.
__global__ void __launch_bounds__(kCullThreads)
cull_compact_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ tag, uint32_t n,
                    ViewConst vc, CullConst cc, const float* __restrict__ aabb,
                    uint32_t* __restrict__ counter, float4* __restrict__ out_pts, uint32_t* __restrict__ out_idx) {
  __shared__ uint32_t s_warp[kCullPer][kCullThreads / 32];
  __shared__ uint32_t s_base;
  const uint32_t base = blockIdx.x * kCullBlock;
  if (!block_may_survive(aabb, blockIdx.x, vc, cc)) return;  // CTA-uniform
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 p[kCullPer];
  unsigned m[kCullPer];
  bool keep[kCullPer];
#pragma unroll
  for (int k = 0; k < kCullPer; k++) {  // sub-block k holds points base + k*256 .. +255 (in order)
    const uint32_t i = base + k * kCullThreads + threadIdx.x;
    keep[k] = false;
    p[k] = make_float4(0, 0, 0, 0);
    if (i < n) {
      p[k] = ldg_stream(pts + i);
      keep[k] = cull_keep(p[k], vc, cc);
    }
    m[k] = __ballot_sync(0xffffffffu, keep[k]);
    if (lane == 0) s_warp[k][warp] = (uint32_t)__popc(m[k]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int k = 0; k < kCullPer; k++)
      for (int w = 0; w < kCullThreads / 32; w++) t += s_warp[k][w];
    s_base = t ? atomicAdd(counter, t) : 0u;
  }
  __syncthreads();
  uint32_t before = s_base;
#pragma unroll
  for (int k = 0; k < kCullPer; k++) {
    uint32_t off = before;
    for (int w = 0; w < warp; w++) off += s_warp[k][w];
    if (keep[k]) {
      const uint32_t o = off + __popc(m[k] & ((1u << lane) - 1u));
      out_pts[o] = p[k];
      out_idx[o] = tag[base + k * kCullThreads + threadIdx.x];  // the z-buffer key's tie-break word
    }
    for (int w = 0; w < kCullThreads / 32; w++) before += s_warp[k][w];
  }
}

// ---- project_splat -----------------------------------------------------------------------
// One thread per surviving point, looping over the views of the group (whose z-buffers are
// L2 resident).  Key = bits(Zc) << 32 | tag, tag = the survivor's 32-bit tie-break word
// (original index, or original index << 8 | value when the model has < 2^24 primitives).
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

// ---- binned tile renderer (point clouds) ---------------------------------------------------
// The classic tiled rasteriser, exact by construction:
//   bin_kernel<2>  steady state, ONE pass: every (survivor, view) pair is projected (fp32,
//                  bit-identical to the oracle) and one 16-byte record {i0|j0, depth bits, tag}
//                  per (splat, 32x32 tile) is written into that [view][tile] bin.  Bins have a
//                  fixed capacity taken from the previous search's fullest bin; a full bin
//                  raises the overflow flag and the search is redone with the exact two-pass
//                  counting sort: bin_kernel<0> counts, a scan turns the counts into offsets,
//                  bin_kernel<1> projects again and writes the records at their slots;
//   tile_resolve   one CTA per (view, tile): the tile's z-buffer lives in SHARED memory
//                  (depth + tag words), every record's fragments are resolved there with
//                  native 32-bit atomics -- pass 1 min depth, pass 2 min tag among the
//                  fragments at the minimum depth (== the packed 64-bit key minimum) -- and
//                  the CTA, sole owner of the tile, stores the finished u8 pixels directly.
// No global z-buffer, no global atomics on pixels, no separate resolve pass.
// bin_kernel waits on the L2 round trip of its slot atomics (ncu: 42 % of the stall cycles).  Measured at
// C2 (round 2, bin + tile_resolve ms): 8 views in flight per thread / 3 CTAs per SM 1.24, 6 / 4 1.14,
// 4 / 5 (round 1, 48 registers with spills) 1.04, 2 / 6 1.01, 1 / 8 and 2 / 8 0.98 -- full occupancy
// (32 registers, 64 warps per SM, no spills) hides the latency better than unrolling does.
#ifndef NMI_BIN_U
#define NMI_BIN_U 1      // views projected together per thread (independent slot atomics in flight)
#endif
#ifndef NMI_BIN_CTAS
#define NMI_BIN_CTAS 8   // resident CTAs per SM the register budget is cut for
#endif
#ifndef NMI_BIN_THREADS
#define NMI_BIN_THREADS 256  // threads per CTA (128 / 512 measured: no difference)
#endif
constexpr int kTile = 32;  // (bin_kernel shifts by 5)
constexpr int kTileThreads = 128;

struct Splat {
  int i0, j0;
  uint32_t zbits;
  bool ok;
};

__device__ __forceinline__ Splat project_splat_point(const float4& p, const float4& c, const ViewConst& vc,
                                                     float half) {
  Splat f;
  f.i0 = f.j0 = 0;
  f.zbits = 0;
  f.ok = false;
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
  // the splat must touch the image at all
  f.ok = f.i0 < vc.W && f.j0 < vc.H && f.i0 + vc.s > 0 && f.j0 + vc.s > 0;
  return f;
}

// MODE 0: count records per (view, tile).  MODE 1: write them at offsets[bin] + slot (second
// pass of the counting sort).  MODE 2: single pass into fixed-capacity bins (capacity known
// from the previous search; a bin that fills up raises the overflow flag).
template <int MODE>
__global__ void __launch_bounds__(NMI_BIN_THREADS, NMI_BIN_CTAS * 256 / NMI_BIN_THREADS)
bin_kernel(const float4* __restrict__ cpts, const uint32_t* __restrict__ ctag,
           const uint32_t* __restrict__ counter, const float4* __restrict__ centres, int nviews,
           ViewConst vc, int ntx, int nt, uint32_t* __restrict__ counts,
           const uint32_t* __restrict__ offsets, uint4* __restrict__ rec, uint32_t rec_cap,
           uint32_t bin_cap, uint32_t* __restrict__ overflow) {
  constexpr bool SCATTER = MODE != 0;
  extern __shared__ float4 s_c[];
  for (int i = threadIdx.x; i < nviews; i += blockDim.x) s_c[i] = centres[i];
  __syncthreads();
  const uint32_t count = *counter;
  const float half = 0.5f * (float)(vc.s - 1);
  const int W = vc.W, H = vc.H, S = vc.s;
  auto emit1 = [&](uint32_t bin, uint32_t ij, uint32_t zbits, uint32_t tag) {
    const uint32_t slot = atomicAdd(&counts[bin], 1u);
    if (MODE == 1) {
      const uint32_t pos = offsets[bin] + slot;
      if (pos < rec_cap)
        rec[pos] = make_uint4(ij, zbits, tag, 0u);
      else
        *overflow = 1u;
    } else if (MODE == 2) {
      if (slot < bin_cap)
        rec[(size_t)bin * bin_cap + slot] = make_uint4(ij, zbits, tag, 0u);
      else
        *overflow = 1u;
    }
  };
  // a splat of s <= 32 pixels touches at most 2 x 2 tiles (usually one)
  auto emit = [&](const Splat& f, int v, uint32_t tag) {
    const int xa = max(f.i0, 0) >> 5, xb = min(f.i0 + S - 1, W - 1) >> 5;
    const int ya = max(f.j0, 0) >> 5, yb = min(f.j0 + S - 1, H - 1) >> 5;
    const uint32_t ij = (uint32_t)(f.i0 + 32768) | ((uint32_t)(f.j0 + 32768) << 16);
    const uint32_t row_a = (uint32_t)v * nt + ya * ntx;
    emit1(row_a + xa, ij, f.zbits, tag);
    if (xb != xa) emit1(row_a + xb, ij, f.zbits, tag);
    if (yb != ya) {
      const uint32_t row_b = (uint32_t)v * nt + yb * ntx;
      emit1(row_b + xa, ij, f.zbits, tag);
      if (xb != xa) emit1(row_b + xb, ij, f.zbits, tag);
    }
  };
  // Slot allocation.  The cloud is Morton-ordered, so neighbouring lanes usually fall into the
  // same (view, tile) bin: each run of consecutive lanes with the same bin takes its slots with
  // ONE atomic (issued by the run's first lane, which adds the run length), an order of
  // magnitude fewer L2 atomics than one per record.  U views are projected together and their
  // atomics issued back to back (U round trips of ~700 cycles in flight per thread; see NMI_BIN_U).
  constexpr int U = NMI_BIN_U;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t t0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); t0 < count; t0 += stride) {
    const uint32_t t = t0 + lane;
    const bool act = t < count;  // whole warps iterate together (shuffles below)
    const float4 p = act ? cpts[t] : make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t tag = (SCATTER && act) ? ctag[t] : 0u;
    int v = 0;
    for (; v + U <= nviews; v += U) {
      Splat f[U];
      uint32_t bin0[U], slot[U], leader[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        f[u] = project_splat_point(p, s_c[v + u], vc, half);
        f[u].ok = f[u].ok && act;
        bin0[u] = (uint32_t)(v + u) * nt + (max(f[u].j0, 0) >> 5) * ntx + (max(f[u].i0, 0) >> 5);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const uint32_t key = f[u].ok ? bin0[u] : 0xFFFFFFFFu;
        const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, key, 1);
        const uint32_t heads = __ballot_sync(0xFFFFFFFFu, lane == 0 || key != prev);
        leader[u] = 31u - (uint32_t)__clz(heads & (0xFFFFFFFFu >> (31u - lane)));
        const uint32_t above = heads & ~((2u << leader[u]) - 1u);
        const uint32_t len = (above ? (uint32_t)__ffs(above) - 1u : 32u) - leader[u];
        slot[u] = 0;
        if (f[u].ok && lane == leader[u]) slot[u] = atomicAdd(&counts[bin0[u]], len);
      }
#pragma unroll
      for (int u = 0; u < U; u++)
        slot[u] = __shfl_sync(0xFFFFFFFFu, slot[u], leader[u]) + (lane - leader[u]);
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (!f[u].ok) continue;
        const uint32_t ij = (uint32_t)(f[u].i0 + 32768) | ((uint32_t)(f[u].j0 + 32768) << 16);
        if (MODE == 1) {
          const uint32_t pos = offsets[bin0[u]] + slot[u];
          if (pos < rec_cap)
            rec[pos] = make_uint4(ij, f[u].zbits, tag, 0u);
          else
            *overflow = 1u;
        } else if (MODE == 2) {
          if (slot[u] < bin_cap)
            rec[(size_t)bin0[u] * bin_cap + slot[u]] = make_uint4(ij, f[u].zbits, tag, 0u);
          else
            *overflow = 1u;
        }
        // the other (up to three) tiles of a splat that straddles a tile edge: ~12 % of splats
        const int xa = max(f[u].i0, 0) >> 5, xb = min(f[u].i0 + S - 1, W - 1) >> 5;
        const int ya = max(f[u].j0, 0) >> 5, yb = min(f[u].j0 + S - 1, H - 1) >> 5;
        if (xb != xa) emit1(bin0[u] + 1, ij, f[u].zbits, tag);
        if (yb != ya) {
          const uint32_t row_b = (uint32_t)(v + u) * nt + yb * ntx;
          emit1(row_b + xa, ij, f[u].zbits, tag);
          if (xb != xa) emit1(row_b + xb, ij, f[u].zbits, tag);
        }
      }
    }
    if (act)
      for (; v < nviews; v++) {
        const Splat f0 = project_splat_point(p, s_c[v], vc, half);
        if (f0.ok) emit(f0, v, tag);
      }
  }
}

// total[2] = fullest bin: the number of records the binning kernel WANTED to file in this CTA's bin, also when that
// exceeded the capacity (the retry sizes its bins from it).  Not inlined: written inside tile_resolve_kernel the same
// statement costs eight registers (40 instead of 32) and a quarter of the kernel's occupancy.
__device__ __noinline__ void report_wanted_fill(uint32_t* __restrict__ total, const uint32_t* __restrict__ offsets) {
  atomicMax(total + 2, offsets[blockIdx.x]);
}

// A splat covers the S x S pixels whose top-left corner is its ANCHOR (i0, j0), so
//   zbuf(x, y) = min over anchors in [x-S+1, x] x [y-S+1, y] of  min key of the splats anchored there
// (min is associative): ONE shared-memory atomic per record and pass into an anchor buffer of
// (32+S-1)^2 cells, then an S x S min filter over the finished cells -- instead of S*S atomics per
// record and pass (9 for the reference's glPointSize(3): 18 atomics per record, 3 wavefronts each,
// was 97 % of the L1 pipe).  Keys are compared as (depth bits, tag), the packed 64-bit order.
template <bool PACKED, int ST>  // ST = 3: the reference's point size, filter in registers; 0: any S <= 32
__global__ void __launch_bounds__(kTileThreads)
tile_resolve_kernel(const uint4* __restrict__ rec, uint32_t rec_cap, uint32_t bin_cap,
                    const uint32_t* __restrict__ offsets, uint32_t* __restrict__ total, int ntx, int nt,
                    int W, int H, int S,
                    const uint8_t* __restrict__ val, uint8_t* __restrict__ images, size_t pitch,
                    uint32_t* __restrict__ winners, size_t P) {
  extern __shared__ __align__(16) uint32_t s_dyn[];
  if (ST) S = ST;
  const int E = kTile + S - 1;      // anchors li, lj in [-(S-1), 31] -> u, v in [0, E)
  uint32_t* s_depth = s_dyn;        // [E * E]
  uint32_t* s_tag = s_dyn + E * E;  // [E * E]
  const uint32_t bin = blockIdx.x;
  const int v = bin / nt, tile = bin - v * nt;
  const int ty = tile / ntx, tx = tile - ty * ntx;
  const int x0 = tx * kTile, y0 = ty * kTile;
  const int tid = threadIdx.x;
  // clamp to the record buffer: after an overflow (flagged by bin_scatter, the search is then
  // redone) the offsets may point past it
  // bin_cap > 0: fixed-capacity bins, `offsets` holds the fill counts; else counting-sort offsets
  // feedback (total[2]): fullest bin -- the number of records the binning kernel WANTED to file here, also when
  // that exceeded the capacity (the retry then sizes its bins from it)
  // (fixed-capacity bins report at the end of the kernel, report_wanted_fill)
  size_t start, end;
  if (bin_cap) {
    start = (size_t)bin * bin_cap;
    end = start + min(offsets[bin], bin_cap);
  } else {
    start = min(offsets[bin], rec_cap);
    end = min(bin + 1 < gridDim.x ? offsets[bin + 1] : *total, rec_cap);
    if (tid == 0) atomicMax(total + 2, (uint32_t)(end - start));
  }
  {
    uint4* s4 = reinterpret_cast<uint4*>(s_dyn);  // 2 * E * E words, rounded up to whole uint4s (the buffer is padded)
    for (int q = tid; q < (2 * E * E + 3) / 4; q += kTileThreads) s4[q] = make_uint4(~0u, ~0u, ~0u, ~0u);
  }
  __syncthreads();
  const int ox = x0 - (S - 1) + 32768, oy = y0 - (S - 1) + 32768;
  // pass 1: minimum depth per anchor cell.  The first kKeep records of a thread stay in registers
  // for pass 2 (the average bin holds ~4 per thread), the rest are read again (L2).
  constexpr int kKeep = 4;
  uint4 keep[kKeep];
  int ckeep[kKeep];
#pragma unroll
  for (int k = 0; k < kKeep; k++) {
    const size_t r = start + tid + (size_t)k * kTileThreads;
    ckeep[k] = -1;
    if (r < end) {
      keep[k] = rec[r];
      const int u = (int)(keep[k].x & 0xFFFFu) - ox, w = (int)(keep[k].x >> 16) - oy;
      if ((unsigned)u < (unsigned)E && (unsigned)w < (unsigned)E) {
        ckeep[k] = w * E + u;
        atomicMin(&s_depth[ckeep[k]], keep[k].y);
      }
    }
  }
  for (size_t r = start + tid + (size_t)kKeep * kTileThreads; r < end; r += kTileThreads) {
    const uint4 e = rec[r];
    const int u = (int)(e.x & 0xFFFFu) - ox, w = (int)(e.x >> 16) - oy;
    if ((unsigned)u < (unsigned)E && (unsigned)w < (unsigned)E) atomicMin(&s_depth[w * E + u], e.y);
  }
  __syncthreads();
  // pass 2: lowest tie-break word among the records at the minimum depth of their cell
#pragma unroll
  for (int k = 0; k < kKeep; k++)
    if (ckeep[k] >= 0 && s_depth[ckeep[k]] == keep[k].y) atomicMin(&s_tag[ckeep[k]], keep[k].z);
  for (size_t r = start + tid + (size_t)kKeep * kTileThreads; r < end; r += kTileThreads) {
    const uint4 e = rec[r];
    const int u = (int)(e.x & 0xFFFFu) - ox, w = (int)(e.x >> 16) - oy;
    if ((unsigned)u < (unsigned)E && (unsigned)w < (unsigned)E && s_depth[w * E + u] == e.y)
      atomicMin(&s_tag[w * E + u], e.z);
  }
  __syncthreads();
  // S x S min filter; the CTA owns the tile: plain stores of the finished pixels (8 per thread, one
  // row segment).  Pixel (i, j) of the tile <- cells u in [i, i+S), v in [j, j+S).
  const int row = tid >> 2, col = (tid & 3) * 8;
  // The minimum of the packed (depth, tag) keys = the minimum depth, then the minimum tag among the cells
  // that hold it: two 32-bit stages (one VIMNMX3 per three values) instead of 64-bit compares.
  uint32_t bd[8], bt[8];
  if (ST == 3) {
    uint32_t vd[10], vt[10];  // per column: min over the three anchor rows
#pragma unroll
    for (int c = 0; c < 10; c++) {
      const int q = row * E + col + c;
      const uint32_t d0 = s_depth[q], d1 = s_depth[q + E], d2 = s_depth[q + 2 * E];
      const uint32_t m = min(d0, min(d1, d2));
      const uint32_t t0 = d0 == m ? s_tag[q] : 0xFFFFFFFFu, t1 = d1 == m ? s_tag[q + E] : 0xFFFFFFFFu,
                     t2 = d2 == m ? s_tag[q + 2 * E] : 0xFFFFFFFFu;
      vd[c] = m;
      vt[c] = min(t0, min(t1, t2));
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const uint32_t m = min(vd[k], min(vd[k + 1], vd[k + 2]));
      const uint32_t t0 = vd[k] == m ? vt[k] : 0xFFFFFFFFu, t1 = vd[k + 1] == m ? vt[k + 1] : 0xFFFFFFFFu,
                     t2 = vd[k + 2] == m ? vt[k + 2] : 0xFFFFFFFFu;
      bd[k] = m;
      bt[k] = min(t0, min(t1, t2));
    }
  } else {
    // generic point size: horizontal minima of every anchor row into shared memory, then vertical
    auto cell_key = [&](int q) { return ((unsigned long long)s_depth[q] << 32) | s_tag[q]; };
    unsigned long long* s_h = reinterpret_cast<unsigned long long*>(s_dyn + ((2 * E * E + 3) & ~3));  // [E][32]
    for (int q = tid; q < E * kTile; q += kTileThreads) {
      const int vv = q >> 5, i = q & 31;
      unsigned long long m = ~0ull;
      for (int du = 0; du < S; du++) m = min(m, cell_key(vv * E + i + du));
      s_h[q] = m;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; k++) {
      unsigned long long m = ~0ull;
      for (int dv = 0; dv < S; dv++) m = min(m, s_h[(row + dv) * kTile + col + k]);
      bd[k] = (uint32_t)(m >> 32);
      bt[k] = (uint32_t)m;
    }
  }
  uint8_t* img = images + (size_t)v * pitch;
  const int y = y0 + row;
  if (y < H) {
    unsigned long long packed = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const uint32_t tg = bt[k];
      const bool empty = bd[k] == 0xFFFFFFFFu;  // depth bits of a real splat are < 0x7F800000
      const int x = x0 + col + k;
      uint32_t pix = 255u;
      if (!empty) pix = PACKED ? (tg & 0xFFu) : (x < W ? (uint32_t)__ldg(val + tg) : 0u);
      packed |= (unsigned long long)pix << (8 * k);
      if (winners && x < W) winners[(size_t)v * P + (size_t)y * W + x] = empty ? NMI_EMPTY : (PACKED ? tg >> 8 : tg);
    }
    const size_t o = (size_t)y * W + x0 + col;
    if (x0 + col + 7 < W && ((reinterpret_cast<size_t>(img) + o) & 7) == 0) {
      *reinterpret_cast<unsigned long long*>(img + o) = packed;
    } else {
      for (int k = 0; k < 8 && x0 + col + k < W; k++) img[o + k] = (uint8_t)(packed >> (8 * k));
    }
  }
  if (bin_cap && tid == 0) {
    if (ST == 3) report_wanted_fill(total, offsets);
    else atomicMax(total + 2, offsets[blockIdx.x]);
  }
}

// z-buffer -> u8 render (background 255, rendering.hpp:533) + reset to ~0.
// PACKED: the key's low word is (original index << 8 | value): no gather is needed.
template <bool PACKED>
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
      const unsigned long long key = zb[p];
      const uint32_t tag = (uint32_t)(key & 0xFFFFFFFFull);
      const bool empty = key == ~0ull;
      const uint32_t w = empty ? NMI_EMPTY : (PACKED ? tag >> 8 : tag);
      pix = empty ? 255u : (PACKED ? (tag & 0xFFu) : (uint32_t)__ldg(val + tag));
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
  fill_u64_kernel<<<sm_count() * 8, 256, 0, st>>>(p, n, v);
}

void launch_intensity_u8(const float4* pts, const uint32_t* orig, uint8_t* val, uint32_t* tag,
                         bool packed, size_t n, cudaStream_t st) {
  if (n == 0) return;
  intensity_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pts, orig, val, tag, packed, n);
}

int launch_cull_compact(const float4* pts, const uint32_t* orig, uint32_t n, const float* aabb, const ViewConst& vc,
                         const float c0[3], const float margin[3], float4* out_pts,
                         uint32_t* out_idx, uint32_t* counter, uint32_t* block_counts,
                         cudaStream_t st) {
  if (n == 0) return 0;
  CullConst cc;
  for (int i = 0; i < 3; i++) cc.c0[i] = c0[i];
  cc.mx = margin[0];
  cc.my = margin[1];
  cc.mz = margin[2];
  const uint32_t nblocks = (n + kCullBlock - 1) / kCullBlock;
  // $NMI_CULL_PASSES=3: the stable three-kernel compaction (A/B switch)
  static const bool one_pass = [] {
    const char* e = getenv("NMI_CULL_PASSES");
    return !(e && atoi(e) == 3);
  }();
  if (one_pass) {
    prefer_max_shared((const void*)cull_compact_kernel);
    cudaMemsetAsync(counter, 0, sizeof(uint32_t), st);
    cull_compact_kernel<<<nblocks, kCullThreads, 0, st>>>(pts, orig, n, vc, cc, aabb, counter, out_pts, out_idx);
    return 1;
  }
  prefer_max_shared((const void*)cull_count_kernel);
  prefer_max_shared((const void*)cull_scan_kernel);
  prefer_max_shared((const void*)cull_scatter_kernel);
  cull_count_kernel<<<nblocks, kCullThreads, 0, st>>>(pts, n, vc, cc, aabb, block_counts);
  cull_scan_kernel<<<1, 1024, 0, st>>>(block_counts, nblocks, counter);
  cull_scatter_kernel<<<nblocks, kCullThreads, 0, st>>>(pts, orig, n, vc, cc, aabb, block_counts, out_pts,
                                                        out_idx);
  return 3;
}

void launch_scan_counts(uint32_t* block_counts, uint32_t nblocks, uint32_t* counter, cudaStream_t st) {
  cull_scan_kernel<<<1, 1024, 0, st>>>(block_counts, nblocks, counter);
}

void launch_project_splat(const float4* cpts, const uint32_t* cidx, const uint32_t* counter,
                          const float4* centres, int nviews, const ViewConst& vc,
                          unsigned long long* zbuf, size_t P, uint32_t max_points, cudaStream_t st) {
  if (nviews == 0 || max_points == 0) return;
  project_splat_kernel<<<sm_count() * 16, 256, sizeof(float4) * nviews, st>>>(cpts, cidx, counter,
                                                                      centres, nviews, vc, zbuf, P);
}

// mode 0 = count, 1 = scatter at offsets (two-pass counting sort), 2 = single pass into
// fixed-capacity bins.  counts / offsets: [nviews * tiles]; rec: rec_cap records.
void launch_bin_points(int mode, const float4* cpts, const uint32_t* ctag, const uint32_t* counter,
                       const float4* centres, int nviews, const ViewConst& vc, uint32_t* counts,
                       const uint32_t* offsets, uint4* rec, uint32_t rec_cap, uint32_t bin_cap,
                       uint32_t* overflow, cudaStream_t st) {
  if (nviews == 0) return;
  const int ntx = (vc.W + kTile - 1) / kTile, nty = (vc.H + kTile - 1) / kTile;
  const dim3 grid(sm_count() * 16 * 256 / NMI_BIN_THREADS), block(NMI_BIN_THREADS);
  const size_t smem = sizeof(float4) * nviews;
  prefer_max_shared((const void*)bin_kernel<0>);
  prefer_max_shared((const void*)bin_kernel<1>);
  prefer_max_shared((const void*)bin_kernel<2>);
  if (mode == 0)
    bin_kernel<0><<<grid, block, smem, st>>>(cpts, ctag, counter, centres, nviews, vc, ntx, ntx * nty, counts,
                                             offsets, rec, rec_cap, bin_cap, overflow);
  else if (mode == 1)
    bin_kernel<1><<<grid, block, smem, st>>>(cpts, ctag, counter, centres, nviews, vc, ntx, ntx * nty, counts,
                                             offsets, rec, rec_cap, bin_cap, overflow);
  else
    bin_kernel<2><<<grid, block, smem, st>>>(cpts, ctag, counter, centres, nviews, vc, ntx, ntx * nty, counts,
                                             offsets, rec, rec_cap, bin_cap, overflow);
}

void launch_copy_rows(const uint8_t* src, size_t src_pitch, uint8_t* dst, int W, int H, bool flip, cudaStream_t st) {
  if (W <= 0 || H <= 0) return;
  copy_rows_kernel<<<dim3((unsigned)((W + 255) / 256), (unsigned)H), 256, 0, st>>>(src, src_pitch, dst, W, H, flip);
}

int tiles_per_view(int W, int H) { return ((W + kTile - 1) / kTile) * ((H + kTile - 1) / kTile); }

void launch_tile_resolve(const uint4* rec, uint32_t rec_cap, uint32_t bin_cap, const uint32_t* offsets, uint32_t* total, int nviews,
                         const ViewConst& vc, const uint8_t* val, bool packed, uint8_t* images,
                         size_t pitch, uint32_t* winners, size_t P, cudaStream_t st) {
  if (nviews == 0) return;
  const int ntx = (vc.W + kTile - 1) / kTile, nty = (vc.H + kTile - 1) / kTile;
  const unsigned grid = (unsigned)nviews * ntx * nty;
  const int E = kTile + vc.s - 1;
  // anchor cells (depth + tag words); generic point sizes add the [E][32] u64 row minima
  size_t smem = (size_t)((2 * E * E + 3) & ~3) * sizeof(uint32_t);  // whole uint4s (the init stores 16 bytes at a time)
  if (vc.s != 3) smem += (size_t)E * kTile * sizeof(unsigned long long);
#define NMI_TR(PK, ST3)                                                                                         \
  prefer_max_shared((const void*)tile_resolve_kernel<PK, ST3>);                                                   \
  tile_resolve_kernel<PK, ST3><<<grid, kTileThreads, smem, st>>>(rec, rec_cap, bin_cap, offsets, total, ntx,     \
                                                                  ntx * nty, vc.W, vc.H, vc.s, val, images, pitch, \
                                                                  winners, P)
  if (vc.s == 3) {
    if (packed) { NMI_TR(true, 3); } else { NMI_TR(false, 3); }
  } else {
    if (packed) { NMI_TR(true, 0); } else { NMI_TR(false, 0); }
  }
#undef NMI_TR
}

void launch_resolve(unsigned long long* zbuf, const uint8_t* val, int nviews, size_t P,
                    uint8_t* images, size_t pitch, uint32_t* winners, bool packed_value,
                    cudaStream_t st) {
  if (nviews == 0 || P == 0) return;
  dim3 grid((unsigned)((P + 1023) / 1024), (unsigned)nviews);
  if (packed_value)
    resolve_kernel<true><<<grid, 256, 0, st>>>(zbuf, val, P, images, pitch, winners);
  else
    resolve_kernel<false><<<grid, 256, 0, st>>>(zbuf, val, P, images, pitch, winners);
}

}  // namespace nmi
