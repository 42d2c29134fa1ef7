// nmi_internal.h -- shared declarations of the sm_100a NMI pose-search library.
// Host-side context + the launch wrappers each .cu file exports.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/nmi_b200.h"

namespace nmi {

// Per synthetic-view camera, consumed by the projection kernel.
// Layout of A.2: Rwc columns + view centre + projection scales.
struct ViewConst {
  float r0[3], r1[3], r2[3];  // columns of Rwc (Twc[:3,0], [:3,1], [:3,2])
  float kx, ky, hw, hh, zn, zf;
  int W, H, s;  // image size, integer point size
};

constexpr int kMaxViewsPerLaunch = 512;
constexpr int kCullBlock = 1024;  // points per CTA of the cull kernels = points per load-time AABB
constexpr size_t kImgAlign = 128;  // every render / warp image starts 128 B aligned

inline size_t img_pitch(size_t P) { return (P + kImgAlign - 1) / kImgAlign * kImgAlign; }

// --- project.cu ------------------------------------------------------------
void launch_fill_u64(unsigned long long* p, size_t n, unsigned long long v, cudaStream_t st);
// val[orig[i]] = u8 intensity; tag[i] = orig[i] << 8 | value when `packed`, else orig[i]
void launch_intensity_u8(const float4* pts, const uint32_t* orig, uint8_t* val, uint32_t* tag,
                         bool packed, size_t n, cudaStream_t st);
// z-buffer -> u8 image (+ optional winner indices); resets the z-buffer to ~0
void launch_resolve(unsigned long long* zbuf, const uint8_t* val, int nviews, size_t P,
                    uint8_t* images, size_t pitch, uint32_t* winners, bool packed_value,
                    cudaStream_t st);

// --- warp.cu ---------------------------------------------------------------
// tex != 0: the frame as a point-sampled, border-0 u8 texture over a gather-capable array
// (same pixels as src); the kernel then fetches the four bilinear taps with one gather
void launch_warp(const uint8_t* src, cudaTextureObject_t tex, int W, int H,
                 const float* minv /*nW x 9, device*/, int nW, uint8_t* dst, size_t pitch, cudaStream_t st);

// --- hist.cu ---------------------------------------------------------------
struct HistArgs {
  const uint8_t* renders;  // render images, pitch bytes apart
  const uint8_t* warps;    // warped images, pitch bytes apart
  size_t render_pitch, warp_pitch;
  const int2* pairs;  // (render slot, warp slot) per evaluation, schedule order
  const uint32_t* out_index;  // score slot per evaluation (rating linear index)
  int npairs;
  uint32_t npix;    // W*H
  uint32_t length;  // kernel.cu:85: always W*H
  int bins, bg, mode, variant;
  float* scores;
  // optional: e(c) = (c/length) log2(c/length) for c = 0..length, from launch_term_table
  const float* term_tab;
  // hot-bin skipping (hist.cu): per image, sampled count << 8 | level of its most frequent
  // grey level (renders [0, nrenders), then the warps); sample_total = pixels sampled per
  // image; skip_mode 0 never, 1 when a pair's two levels cover >= 1/6 of the samples,
  // 2 always; skipcap = launch the build that carries the side tables.
  const uint32_t* img_mode;
  // optional (SKIPCAP builds): full 256-bin histogram of every render, then of every warp
  // (launch_image_hists).  Then a skipped pixel is not counted at all: row a* / column b* of the joint
  // histogram are reconstructed from these marginals (hist.cu, marginal_side_counts).
  const uint32_t* img_hist;
  uint32_t sample_total;
  int nrenders;
  int skip_mode;
  bool skipcap;
  // optional dumps (parity): when non-null, pair 0 of the launch writes them (the DUMP build of the
  // persistent kernel: pair `dump_pair`)
  uint32_t* dumpJ;
  uint32_t* dumpHA;
  uint32_t* dumpHB;
  int dump_pair;
  // parity read-back of the batched build: do not route a single evaluation to the cluster kernel
  bool force_batched;
  // img_mode == nullptr: let the (cluster) kernel sample the two images' dominant grey levels itself
  bool sample_in_kernel;
};
// true when launch_joint_hist_score will route `a` to the single-evaluation cluster kernel (which can do its
// own mode sampling: no image_mode launch needed in front of it)
bool hist_uses_cluster(const HistArgs& a);
int launch_joint_hist_score(const HistArgs& a, cudaStream_t st);  // returns launches, <0 on error
// sampled per-image modes for HistArgs::img_mode; hot[0] / hot[1] = largest sampled count over
// the renders / the warps.  Returns launches.
// clear_hot = false: hot[] was zeroed by the caller (a search samples its renders and its warps in two launches,
// each on the stream that produced them)
int launch_image_modes(const uint8_t* renders, size_t rpitch, int nr, const uint8_t* warps, size_t wpitch,
                       int nw, uint32_t npix, uint32_t* img_mode, uint32_t* hot, cudaStream_t st, bool clear_hot = true);
uint32_t image_mode_sample_total(uint32_t npix);
// img_hist[(nr + nw) * 256]: exact histograms of all npix pixels of every image.  Returns launches.
int launch_image_hists(const uint8_t* renders, size_t rpitch, int nr, const uint8_t* warps, size_t wpitch,
                       int nw, uint32_t npix, uint32_t* img_hist, cudaStream_t st);
void launch_term_table(float* tab, uint32_t length, cudaStream_t st);

// --- argmax.cu -------------------------------------------------------------
// key = (bits(max(0, scores)) << 32) | (0xFFFFFFFF - lowest index with score == max)
void launch_argmax(const float* scores, const uint32_t* index_list, int n_list, uint32_t n_total,
                   unsigned long long* key, const uint32_t* retry_flag, cudaStream_t st);

// --- host_math.cpp ----------------------------------------------------------
void make_view_const(const nmi_camera& cam, const float Twc[16], ViewConst* vc);
uint64_t pack_key(float max_score, int64_t index);

void set_error(const std::string& msg);

// An SM's L1 / shared-memory split is a per-SM state: CTAs of two kernels that ask for different
// carve-outs cannot be resident on one SM together.  With $NMI_CARVEOUT=1 the render-stage and warp
// kernels prefer the maximum shared-memory split the histogram kernel needs, so that their CTAs can
// run next to a histogram CTA when two searches are in flight on one GPU (tools/exp_pipeline.py).
// Measured in round 2: no gain (6.17 vs 6.06 ms per search at depth 2, every kernel a little slower
// with the small L1) -- off by default, kept as a switch.
void prefer_max_shared(const void* kernel);
// SM count of the current device (cached per device; 148 on B200): grid-stride kernels size their grids
// as a multiple of it.
int sm_count();

}  // namespace nmi
