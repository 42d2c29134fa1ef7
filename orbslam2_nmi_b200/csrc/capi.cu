// capi.cu -- the C ABI (include/nmi_b200.h) over the sm_100a kernels.
//
// One nmi_ctx owns a device, a stream and every scratch buffer; nothing is
// allocated per evaluation (the reference allocates/frees 10 buffers and
// registers a GL texture per evaluation, kernel.cu:52-113).  A grid search is
// params-upload + a handful of kernels on one stream:
//   cull_compact -> (project_splat -> resolve) per L2-sized view group -> warp ->
//   joint_hist_score -> argmax
// No CPU fallback exists: without a usable CUDA device every entry point fails.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "nmi_internal.h"

namespace nmi {

// project.cu (not in the header: only capi uses them)
int launch_cull_compact(const float4* pts, const uint32_t* orig, uint32_t n, const float* aabb, const ViewConst& vc,
                         const float c0[3], const float margin[3], float4* out_pts,
                         uint32_t* out_idx, uint32_t* counter, uint32_t* block_counts,
                         cudaStream_t st);
void launch_project_splat(const float4* cpts, const uint32_t* cidx, const uint32_t* counter,
                          const float4* centres, int nviews, const ViewConst& vc,
                          unsigned long long* zbuf, size_t P, uint32_t max_points, cudaStream_t st);

void launch_scan_counts(uint32_t* block_counts, uint32_t nblocks, uint32_t* counter, cudaStream_t st);
void launch_copy_rows(const uint8_t* src, size_t src_pitch, uint8_t* dst, int W, int H, bool flip, cudaStream_t st);
void launch_bin_points(int mode, const float4* cpts, const uint32_t* ctag, const uint32_t* counter,
                       const float4* centres, int nviews, const ViewConst& vc, uint32_t* counts,
                       const uint32_t* offsets, uint4* rec, uint32_t rec_cap, uint32_t bin_cap,
                       uint32_t* overflow, cudaStream_t st);
int tiles_per_view(int W, int H);
void launch_tile_resolve(const uint4* rec, uint32_t rec_cap, uint32_t bin_cap, const uint32_t* offsets, uint32_t* total, int nviews,
                         const ViewConst& vc, const uint8_t* val, bool packed, uint8_t* images,
                         size_t pitch, uint32_t* winners, size_t P, cudaStream_t st);

// mesh.cu
void launch_mesh_values(const float4* verts, const uint3* tris, const uint32_t* tri_orig, uint8_t* val,
                        uint32_t nt, cudaStream_t st);
int launch_mesh_cull(const float4* verts, const uint3* tris, uint32_t nt, const ViewConst& vc,
                      const float c0[3], const float margin[3], uint32_t* slots, uint32_t* counter,
                      uint32_t* block_counts, uint8_t* vflag, uint32_t nv, cudaStream_t st);
void launch_mesh_vertices(const float4* verts, const uint8_t* vflag, uint32_t nv, const float4* centres, int nviews,
                          const ViewConst& vc, int4* tv, cudaStream_t st);
void launch_mesh_raster(const int4* tv, uint32_t nv, const uint3* tris, const uint32_t* tri_orig,
                        const uint32_t* slots, const uint32_t* counter, int nviews,
                        const ViewConst& vc, unsigned long long* zbuf, size_t P, cudaStream_t st);
void launch_mesh_luma(const uint8_t* tex, float* luma, size_t n, cudaStream_t st);
void launch_mesh_shade(unsigned long long* zbuf, const int4* tv, uint32_t nv, const uint4* tris_orig, const float4* corner_uv,
                       const float* luma, int tw, int th, int nviews, const ViewConst& vc,
                       size_t P, uint8_t* images, size_t pitch, uint32_t* winners, cudaStream_t st);

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

int sm_count() {
  static thread_local int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

void prefer_max_shared(const void* kernel) {
  static const bool on = [] {
    const char* e = getenv("NMI_CARVEOUT");
    return e && atoi(e) != 0;
  }();
  if (!on) return;
  struct Seen { const void* k; int dev; };
  static thread_local std::vector<Seen> seen;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return;
  for (const Seen& s : seen)
    if (s.k == kernel && s.dev == dev) return;
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaGetLastError();
  seen.push_back({kernel, dev});
}

}  // namespace nmi

using namespace nmi;

#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      char buf__[512];                                                                   \
      snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
               __FILE__, __LINE__);                                                      \
      set_error(buf__);                                                                  \
      return NMI_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define REQUIRE(cond, code, msg) \
  do {                           \
    if (!(cond)) {               \
      set_error(msg);            \
      return code;               \
    }                            \
  } while (0)

namespace {
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};
}  // namespace

struct nmi_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // the frame warps run here, concurrently with the render stage
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // mesh models: the view groups of a search take turns on `stream` and up to three more streams (each with
  // its own slot of the z-buffer and of the vertex table), so that the issue-bound raster kernel of one group
  // runs under the latency-bound shading kernel of another
  static constexpr int kMeshStreams = 4;  // slot 0 = `stream`, slots 1.. = mesh_stream[]
  cudaStream_t mesh_stream[kMeshStreams - 1] = {};
  cudaEvent_t ev_fork3 = nullptr, ev_join3[kMeshStreams - 1] = {};
  cudaEvent_t ev[8] = {};
  cudaEvent_t ev_params = nullptr;  // completion of the last H2D from h_params
  bool params_in_flight = false;
  bool timed = false;

  bool has_cam = false, has_frame = false;
  nmi_camera cam{};
  size_t P = 0, pitch = 0;

  DevBuf<float4> pts;
  DevBuf<uint32_t> orig;  // original index of every (Morton-ordered) point
  DevBuf<uint32_t> tag;   // tie-break word of the z-buffer key: orig, or orig << 8 | value
  bool packed_value = false;  // model has < 2^24 primitives: the key carries the value
  DevBuf<uint8_t> val;    // u8 intensity, indexed by ORIGINAL index
  DevBuf<float> aabb;     // {lo xyz, hi xyz} of every kCullBlock consecutive (Morton-ordered) points
  bool use_block_cull = true;  // $NMI_BLOCK_CULL=0: per-point cull only (A/B switch)
  size_t n_pts = 0;
  DevBuf<float4> cpts;    // compacted survivors of the cull
  DevBuf<uint32_t> cidx;  // their original indices
  DevBuf<uint32_t> counter;
  DevBuf<uint32_t> block_counts;  // per-CTA survivor counts / offsets of the stable compaction

  // mesh model (Rendering<1>): Morton-ordered triangles + their original (draw-order) index
  DevBuf<float4> mverts;
  DevBuf<uint3> mtris;
  DevBuf<uint32_t> mtri_orig, mslots;
  size_t n_tris = 0;  // > 0: the model is a mesh, else a point cloud
  // textured mesh (Rendering<1> as the reference runs it): triangles in ORIGINAL order (what the
  // z-buffer key names), un-indexed corner UVs, the texture's luma in 0..255 units
  DevBuf<uint4> mtris_o;   // {v0, v1, v2, -}
  DevBuf<float4> muv;      // two per triangle: {u0, v0, u1, v1}, {u2, v2, -, -}
  DevBuf<float> mluma;
  // per-view vertex table of the current view group ([vertex][view], mesh.cu) + the vertices the
  // cull flagged (corners of surviving triangles)
  int mesh_slots = 1;  // view groups in flight in the current mesh search (stream slots of z-buffer / vertex table)
  DevBuf<int4> mtv;
  DevBuf<uint8_t> mvflag;
  size_t n_verts = 0;
  int tex_w = 0, tex_h = 0;  // > 0: per-fragment texture shading

  // feedback from the previous search, copied to pinned host memory asynchronously:
  // [0] survivors, [1] records of the last view group, [2] overflow flag, [3] fullest bin,
  // [4] views of that group (host-written)
  uint32_t* h_feedback = nullptr;
  uint32_t bin_cap = 0;  // > 0: this search bins in a single pass into bins of that capacity
  cudaEvent_t ev_feedback = nullptr;
  bool feedback_pending = false;
  // fullest bins of the last searches: the single-pass bin capacity covers the fullest of them,
  // so a driver that alternates between coarse and fine levels (or a tracker whose pose jumps
  // back) does not overflow the bins sized by the one search before
  static constexpr int kFullestHist = 16;
  uint32_t fullest_hist[kFullestHist] = {};
  int fullest_pos = 0;
  bool force_conservative = false;
  uint32_t retry_fullest = 0;  // > 0: redo of a search whose bins overflowed; the fullest bin it wanted
  bool conservative_once = false;  // an enqueued search overflowed: size the next one exactly
  // an enqueued (sharded) search overflowed its fixed-capacity bins: the fullest bin it wanted, for ONE redo in a
  // single pass with bins of that size (what nmi_search does through retry_fullest); was_sized_retry: the search
  // that just ran was such a redo, so another overflow falls back to the pose-independent sizing
  uint32_t retry_fullest_once = 0;
  bool was_sized_retry = false;
  // binned tile renderer scratch (point clouds)
  DevBuf<uint32_t> img_hist;  // exact 256-bin histogram of every render / warp of the current search (hot-bin skipping)
  DevBuf<uint32_t> bin_offsets, bin_cursor;  // [views of a group * tiles]
  DevBuf<uint32_t> bin_total;                // [0] records of the group, [1] overflow flag
  DevBuf<uint4> records;

  DevBuf<uint8_t> frame;
  // the same frame as a gather-capable array + point-sampled border-0 texture for the warp
  // kernel ($NMI_WARP_TEX=0 keeps the plain-load kernel); refreshed lazily before a warp launch
  cudaArray_t frame_arr = nullptr;
  cudaTextureObject_t frame_tex = 0;
  int frame_arr_w = 0, frame_arr_h = 0;
  bool frame_tex_dirty = true;
  bool use_warp_tex = true;
  // two pinned staging buffers for pageable host frames: frame k+1 is copied in while the
  // H2D of frame k (and whatever follows it on the stream) is still in flight
  unsigned char* h_frame[2] = {nullptr, nullptr};
  size_t h_frame_cap[2] = {0, 0};
  cudaEvent_t ev_frame[2] = {nullptr, nullptr};  // "the H2D out of staging buffer i is done"
  // The frame is uploaded on stream2 -- the only stream that reads it during a search (texture refresh + warp
  // kernel) -- so the cull / render stage on the main stream does not queue behind 2 MB crossing PCIe.
  // ev_frame_order: main-stream work enqueued before the upload (it may still read the old frame);
  // ev_frame_up: the upload; main-stream readers of the frame (nmi_warp_cells) and nmi_ctx_sync wait for it.
  cudaEvent_t ev_frame_order = nullptr, ev_frame_up = nullptr;
  bool frame_up_pending = false;
  int frame_slot = 0;

  DevBuf<unsigned long long> zbuf;
  size_t zbuf_clean = 0;  // elements known to hold ~0
  DevBuf<uint8_t> renders, warps;
  DevBuf<float> scores;
  // hot-bin skipping of the histogram kernel (hist.cu): sampled mode per image; hot[0..1] =
  // largest mode count over the renders / warps, copied to h_feedback[5..6] after every
  // search so the NEXT search knows whether to launch the build with the side tables
  DevBuf<float> term_tab;      // entropy term of every count 0..term_tab_len (4 B per pixel)
  uint32_t term_tab_len = 0;
  DevBuf<uint32_t> img_mode, hot;
  int hist_skip = 1;  // 0 never, 1 automatic, 2 always ($NMI_HIST_SKIP, nmi_ctx_set_hist_skip)
  bool last_skipcap = false;  // the last batched launch was the build with the side tables
  cudaEvent_t ev_hot = nullptr;
  bool hot_pending = false;
  uint32_t hot_total = 0;  // pixels sampled per image when h_feedback[5..6] were written
  DevBuf<unsigned long long> key;
  DevBuf<unsigned long long> xkey;  // nmi_ctx_key_buffer(): exchange buffer of the sharded driver
  DevBuf<unsigned char> params;  // device copy of the per-search parameter block
  unsigned char* h_params = nullptr;
  size_t h_params_cap = 0;

  // single-evaluation (stage API) + parity scratch
  DevBuf<uint8_t> one_render, one_warp;
  DevBuf<uint32_t> winners, dumpJ, dumpH;
  DevBuf<float> one_score;
  float* h_score = nullptr;  // pinned and device-visible: a single evaluation's kernel stores its score here, no copy call
  DevBuf<int2> zero_pair;  // device int2 {0,0}: the pair list of a single evaluation

  // last search
  bool has_search = false;
  nmi_grid grid{};
  float Twc[16] = {};
  int nvl = 0, nwl = 0;  // local views / warps held in renders / warps
  int v_begin = 0, w_begin = 0;
  size_t off_centres = 0;
  int launches = 0;
};

namespace {

// h_params is reused by every call: wait until the previous upload has been consumed
int params_acquire(nmi_ctx* c) {
  if (c->params_in_flight) {
    CK(cudaEventSynchronize(c->ev_params));
    c->params_in_flight = false;
  }
  return NMI_OK;
}
int params_uploaded(nmi_ctx* c) {
  CK(cudaEventRecord(c->ev_params, c->stream));
  c->params_in_flight = true;
  return NMI_OK;
}

int ensure_pinned(unsigned char** p, size_t* cap, size_t n) {
  if (n <= *cap) return NMI_OK;
  if (*p) cudaFreeHost(*p);
  *p = nullptr;
  *cap = 0;
  CK(cudaMallocHost(p, n));
  *cap = n;
  return NMI_OK;
}

int ensure_zbuf(nmi_ctx* c, size_t elems) {
  if (elems > c->zbuf.cap) {
    CK(c->zbuf.reserve(elems));
    c->zbuf_clean = 0;
  }
  if (c->zbuf_clean < elems) {
    launch_fill_u64(c->zbuf.p, c->zbuf.cap, ~0ull, c->stream);
    CK(cudaGetLastError());
    c->zbuf_clean = c->zbuf.cap;
  }
  return NMI_OK;
}

int vc_point_size(const nmi_camera& cam) {
  const int s = (int)lround(cam.point_size);
  return s < 1 ? 1 : s;
}

// Record buffer of the tile renderer: one 16-byte record per (splat, tile).  A splat of
// s <= 32 px touches at most 4 tiles but ~1.13 on average; the buffer is sized for
// 1.5 records per (point, view) of a group and the group shrinks until that fits 2 GiB.
// bin_scatter never writes past the buffer; an overflow raises a flag that every
// synchronous entry point turns into an error (never a silently wrong render).
// Growing the record buffer is a cudaFree + cudaMalloc of gigabytes (~0.5 s, seen as one slow frame in a
// sequence): once a request passes 1 GiB the buffer goes straight to the 4 GiB that the single-pass layout
// may ever ask for, so a sequence pays for it once, in its first search.
int grow_records(nmi_ctx* c, size_t want) {
  if (want <= c->records.cap) return NMI_OK;
  const size_t limit = (4ull << 30) / sizeof(uint4);
  if (want * sizeof(uint4) > (1ull << 30) && want < limit) want = limit;
  CK(cudaStreamSynchronize(c->stream));
  CK(c->records.reserve(want));
  return NMI_OK;
}

int ensure_tile_buffers(nmi_ctx* c, int nviews, int* group) {
  const size_t tiles = (size_t)tiles_per_view(c->cam.W, c->cam.H);
  const bool fb = c->h_feedback && !c->force_conservative && !c->conservative_once && c->feedback_pending &&
                  cudaEventQuery(c->ev_feedback) == cudaSuccess && c->h_feedback[4] > 0 &&
                  c->h_feedback[2] == 0;
  CK(c->bin_total.reserve(4));
  c->bin_cap = 0;
  // (a) single pass: every bin gets the capacity of the previous search's fullest bin (+25 %)
  if (!c->feedback_pending)  // new model / camera: the history belongs to the old one
    for (uint32_t& v : c->fullest_hist) v = 0;
  const uint32_t retry_size = c->retry_fullest ? c->retry_fullest : c->retry_fullest_once;
  const bool retry = retry_size != 0 && !c->force_conservative && !c->conservative_once;
  if (((fb && c->h_feedback[3] > 0) || retry) && nviews <= kMaxViewsPerLaunch) {
    c->fullest_hist[c->fullest_pos] = retry ? retry_size : c->h_feedback[3];
    c->fullest_pos = (c->fullest_pos + 1) % nmi_ctx::kFullestHist;
    uint32_t fullest = 0;
    for (uint32_t v : c->fullest_hist) fullest = v > fullest ? v : fullest;
    size_t cap = (size_t)fullest + fullest / 4 + 32;
    cap = (cap + 7) / 8 * 8;
    const size_t want = tiles * (size_t)nviews * cap;
    if (want <= (4ull << 30) / sizeof(uint4) && want < 0xFFFFFFFFull) {
      if (int rc = grow_records(c, want)) return rc;
      c->bin_cap = (uint32_t)cap;
      *group = nviews;
      CK(c->bin_cursor.reserve(tiles * (size_t)nviews));
      CK(cudaMemsetAsync(c->bin_total.p, 0, 4 * sizeof(uint32_t), c->stream));
      return NMI_OK;
    }
  }
  // (b) two-pass counting sort; record buffer sized for any pose (1.5 records per point and
  // view) or, in steady state, from what the previous search really produced (x1.5)
  const size_t cap_records = (2ull << 30) / sizeof(uint4);
  // records per (point, view): a splat of s px touches (1 + (s-1)/32)^2 tiles on average, 4 at most
  const double sps = (double)(vc_point_size(c->cam) - 1) / 32.0;
  const double tiles_per_splat = std::min(4.0, (1.0 + sps) * (1.0 + sps)) * 1.35 + 0.15;  // 1.5 for s = 3
  size_t per_view = (size_t)((double)c->n_pts * tiles_per_splat) + 65536;
  if (fb && c->h_feedback[1] > 0) {  // ([1] is only filled by a two-pass search)
    const size_t seen = (size_t)c->h_feedback[1] / c->h_feedback[4];
    const size_t guess = seen + seen / 2 + 65536;
    if (guess < per_view) per_view = guess;
  }
  int g = (int)(cap_records / per_view);
  if (g < 1) g = 1;
  if (g > nviews) g = nviews;
  if (g > kMaxViewsPerLaunch) g = kMaxViewsPerLaunch;
  *group = g;
  const size_t nbins = (size_t)g * tiles;
  CK(c->bin_offsets.reserve(nbins));
  CK(c->bin_cursor.reserve(nbins));
  size_t want = per_view * (size_t)g;
  if (want > cap_records) want = cap_records;
  if (int rc = grow_records(c, want)) return rc;
  CK(cudaMemsetAsync(c->bin_total.p, 0, 4 * sizeof(uint32_t), c->stream));
  return NMI_OK;
}

// Frame -> texture for the warp kernel; enqueued on `st` right before the warp launch.
int ensure_frame_texture(nmi_ctx* c, cudaStream_t st, cudaTextureObject_t* tex) {
  *tex = 0;
  if (!c->use_warp_tex) return NMI_OK;
  if (!c->frame_arr || c->frame_arr_w != c->cam.W || c->frame_arr_h != c->cam.H) {
    if (c->frame_tex) cudaDestroyTextureObject(c->frame_tex);
    if (c->frame_arr) cudaFreeArray(c->frame_arr);
    c->frame_tex = 0;
    c->frame_arr = nullptr;
    const cudaChannelFormatDesc desc = cudaCreateChannelDesc<unsigned char>();
    CK(cudaMallocArray(&c->frame_arr, &desc, (size_t)c->cam.W, (size_t)c->cam.H, cudaArrayTextureGather));
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = c->frame_arr;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;  // BORDER_CONSTANT 0 (image.cpp:123)
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    CK(cudaCreateTextureObject(&c->frame_tex, &rd, &td, nullptr));
    c->frame_arr_w = c->cam.W;
    c->frame_arr_h = c->cam.H;
    c->frame_tex_dirty = true;
  }
  if (c->frame_tex_dirty) {
    CK(cudaMemcpy2DToArrayAsync(c->frame_arr, 0, 0, c->frame.p, (size_t)c->cam.W, (size_t)c->cam.W,
                                (size_t)c->cam.H, cudaMemcpyDeviceToDevice, st));
    c->frame_tex_dirty = false;
  }
  *tex = c->frame_tex;
  return NMI_OK;
}

// entropy-term table of the histogram kernel: depends on the image size only
int ensure_term_table(nmi_ctx* c, uint32_t length) {
  if (c->term_tab.p && c->term_tab_len == length) return NMI_OK;
  CK(c->term_tab.reserve((size_t)length + 1));
  launch_term_table(c->term_tab.p, length, c->stream);
  CK(cudaGetLastError());
  c->term_tab_len = length;
  return NMI_OK;
}

bool valid_grid(const nmi_grid* g) {
  if (!g) return false;
  long np = 1;
  for (int k = 0; k < 3; k++) {
    if (g->nS[k] < 1 || g->nW[k] < 1 || g->nS[k] > 4096 || g->nW[k] > 4096) return false;
    np *= (long)g->nS[k] * g->nW[k];
    if (np > (1l << 26)) return false;
  }
  return true;
}

bool valid_flags(const nmi_flags* f) {
  return f && (f->bins == 256 || f->bins == 64) &&
         (f->score_mode == NMI_SCORE_SUC || f->score_mode == NMI_SCORE_ENMI) && f->variant >= 0 &&
         f->variant <= 11;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Cull the model against the union of the view frusta (order-preserving compaction).
int cull_model(nmi_ctx* c, const ViewConst& vc, const float Twc[16], const float margin[3]) {
  const float c0[3] = {Twc[3], Twc[7], Twc[11]};
  if (c->n_tris) {
    c->launches += launch_mesh_cull(c->mverts.p, c->mtris.p, (uint32_t)c->n_tris, vc, c0, margin, c->mslots.p,
                                    c->counter.p, c->block_counts.p, c->mvflag.p, (uint32_t)c->n_verts, c->stream);
  } else {
    c->launches += launch_cull_compact(c->pts.p, c->tag.p, (uint32_t)c->n_pts, c->use_block_cull ? c->aabb.p : nullptr, vc, c0, margin,
                        c->cpts.p, c->cidx.p,
                        c->counter.p, c->block_counts.p, c->stream);
  }
  CK(cudaGetLastError());
  return NMI_OK;
}

// z-buffer the culled model into `nviews` views (z-buffer views [0, nviews)) and resolve them.
int draw_views(nmi_ctx* c, const ViewConst& vc, const float4* d_centres, int nviews, uint8_t* images,
               uint32_t* winners, int half = 0, size_t half_views = 0) {
  if (!c->n_tris && vc.s <= 32) {
    // point cloud: binned tile renderer, no global z-buffer
    const size_t nbins = (size_t)nviews * tiles_per_view(vc.W, vc.H);
    CK(cudaMemsetAsync(c->bin_cursor.p, 0, nbins * sizeof(uint32_t), c->stream));
    if (c->bin_cap) {
      // steady state: the fullest bin of the previous search bounds every bin -> one pass
      launch_bin_points(2, c->cpts.p, c->cidx.p, c->counter.p, d_centres, nviews, vc, c->bin_cursor.p,
                        nullptr, c->records.p, (uint32_t)c->records.cap, c->bin_cap, c->bin_total.p + 1,
                        c->stream);
      launch_tile_resolve(c->records.p, (uint32_t)c->records.cap, c->bin_cap, c->bin_cursor.p,
                          c->bin_total.p, nviews, vc, c->val.p, c->packed_value, images, c->pitch, winners,
                          c->P, c->stream);
      c->launches += 2;
    } else {
      CK(cudaMemsetAsync(c->bin_offsets.p, 0, nbins * sizeof(uint32_t), c->stream));
      launch_bin_points(0, c->cpts.p, c->cidx.p, c->counter.p, d_centres, nviews, vc, c->bin_offsets.p,
                        nullptr, nullptr, 0, 0, nullptr, c->stream);
      launch_scan_counts(c->bin_offsets.p, (uint32_t)nbins, c->bin_total.p, c->stream);
      launch_bin_points(1, c->cpts.p, c->cidx.p, c->counter.p, d_centres, nviews, vc, c->bin_cursor.p,
                        c->bin_offsets.p, c->records.p, (uint32_t)c->records.cap, 0, c->bin_total.p + 1,
                        c->stream);
      launch_tile_resolve(c->records.p, (uint32_t)c->records.cap, 0, c->bin_offsets.p, c->bin_total.p,
                          nviews, vc, c->val.p, c->packed_value, images, c->pitch, winners, c->P, c->stream);
      c->launches += 4;
    }
    if (c->h_feedback) {  // survivors / records / overflow of this group -> pinned host words
      CK(cudaMemcpyAsync(c->h_feedback, c->counter.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaMemcpyAsync(c->h_feedback + 1, c->bin_total.p, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                         c->stream));
      c->h_feedback[4] = (uint32_t)nviews;
      CK(cudaEventRecord(c->ev_feedback, c->stream));
      c->feedback_pending = true;
    }
    CK(cudaGetLastError());
    return NMI_OK;
  }
  // half 1 (mesh view groups only): second stream, second half of the z-buffer and of the vertex table
  cudaStream_t st = half ? c->mesh_stream[half - 1] : c->stream;
  unsigned long long* zb = c->zbuf.p + (size_t)half * half_views * c->P;
  if (c->n_tris) {
    CK(c->mtv.reserve(c->n_verts * (half_views ? (size_t)c->mesh_slots * half_views : (size_t)nviews)));
    int4* tv = c->mtv.p + (size_t)half * half_views * c->n_verts;
    launch_mesh_vertices(c->mverts.p, c->mvflag.p, (uint32_t)c->n_verts, d_centres, nviews, vc, tv, st);
    launch_mesh_raster(tv, (uint32_t)c->n_verts, c->mtris.p, c->mtri_orig.p, c->mslots.p, c->counter.p, nviews, vc, zb, c->P, st);
    c->launches++;
    if (c->tex_w > 0)
      launch_mesh_shade(zb, tv, (uint32_t)c->n_verts, c->mtris_o.p, c->muv.p, c->mluma.p, c->tex_w, c->tex_h, nviews, vc, c->P, images,
                        c->pitch, winners, st);
    else
      launch_resolve(zb, c->val.p, nviews, c->P, images, c->pitch, winners, c->packed_value, st);
    c->launches += 2;
    CK(cudaGetLastError());
    return NMI_OK;
  } else {
    launch_project_splat(c->cpts.p, c->cidx.p, c->counter.p, d_centres, nviews, vc, c->zbuf.p, c->P,
                         (uint32_t)c->n_pts, c->stream);
  }
  launch_resolve(c->zbuf.p, c->val.p, nviews, c->P, images, c->pitch, winners, c->packed_value, c->stream);
  c->launches += 2;
  CK(cudaGetLastError());
  return NMI_OK;
}

// Render `nviews` cells whose centres sit at d_centres into images[0..nviews).
int render_views(nmi_ctx* c, const ViewConst& vc, const float4* d_centres, int nviews,
                 const float margin[3], bool recull, uint8_t* images, uint32_t* winners) {
  if (recull)
    if (int rc = cull_model(c, vc, c->Twc, margin)) return rc;
  return draw_views(c, vc, d_centres, nviews, images, winners);
}

// schedule order: tiles of 8 renders x 16 warps, so the ~148 CTAs in flight share
// ~24 images (48 MB at 1080p) and the histogram inputs are served from L2.
// hp[k] = (local render, local warp) of evaluation k, hi[k] = its slot in the rating array
// (warp-major, synthetic view fastest: rating[wz][wy][wx][sz][sy][sx], localization.cpp:185-210).
void fill_pair_schedule(int2* hp, uint32_t* hi, int nvl, int nwl, int nS_total, int vb, int wb) {
  const int TS = 8, TW = 16;
  size_t k = 0;
  for (int s0 = 0; s0 < nvl; s0 += TS)
    for (int w0 = 0; w0 < nwl; w0 += TW)
      for (int w = w0; w < w0 + TW && w < nwl; w++)
        for (int s = s0; s < s0 + TS && s < nvl; s++) {
          hp[k] = make_int2(s, w);
          hi[k] = (uint32_t)((size_t)(w + wb) * nS_total + (s + vb));
          k++;
        }
}

// Histogram + score of every scheduled (render, warp) pair of c->renders x c->warps, then the
// on-device argmax -- the nS*nW calls of CUDAF::NMIWithCuda_noMask and find_max_elements of one
// grid search (Tracking.cc:1886-1905).  Records ev[4..6].
int score_pairs_launch(nmi_ctx* c, const nmi_flags* f, int nvl, int nwl, const int2* d_pairs,
                       const uint32_t* d_index, size_t npl, size_t nP, unsigned long long* key_dev,
                       float* scores_dev, bool tiled, bool modes_sampled = false) {
  if (c->timed) CK(cudaEventRecord(c->ev[4], c->stream));
  // hot-bin skipping: sample every image's dominant grey level (decides per pair inside the
  // kernel); the build with the side tables is launched when the previous search saw levels
  // that can reach the 1/6 threshold
  const bool use_skip = c->hist_skip != 0 && f->bins == 256 && f->bg;
  bool skipcap = c->hist_skip == 2;
  if (use_skip) {
    if (c->hist_skip == 1 && c->hot_pending && cudaEventQuery(c->ev_hot) == cudaSuccess)
      skipcap = ((unsigned long long)c->h_feedback[5] + c->h_feedback[6]) * 6ull >= c->hot_total;
    if (!modes_sampled) {  // (a search has done it already: renders and warps each behind their own kernels)
      CK(c->img_mode.reserve((size_t)(nvl + nwl)));
      CK(c->hot.reserve(2));
      c->launches += launch_image_modes(c->renders.p, c->pitch, nvl, c->warps.p, c->pitch, nwl, (uint32_t)c->P,
                                        c->img_mode.p, c->hot.p, c->stream);
    }
    CK(cudaMemcpyAsync(c->h_feedback + 5, c->hot.p, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaEventRecord(c->ev_hot, c->stream));
    c->hot_pending = true;
    c->hot_total = image_mode_sample_total((uint32_t)c->P);
  }
  HistArgs a{};
  a.renders = c->renders.p;
  a.warps = c->warps.p;
  a.render_pitch = a.warp_pitch = c->pitch;
  a.pairs = d_pairs;
  a.out_index = d_index;
  a.npairs = (int)npl;
  a.npix = (uint32_t)c->P;
  a.length = (uint32_t)c->P;
  a.bins = f->bins;
  a.bg = f->bg;
  a.mode = f->score_mode;
  a.variant = f->variant;
  a.scores = scores_dev ? scores_dev : c->scores.p;
  if (int rc = ensure_term_table(c, (uint32_t)c->P)) return rc;
  a.term_tab = c->term_tab.p;
  if (use_skip) {
    a.img_mode = c->img_mode.p;
    a.sample_total = image_mode_sample_total((uint32_t)c->P);
    a.nrenders = nvl;
    a.skip_mode = c->hist_skip;
    a.skipcap = skipcap;
    // the build with the skip code is about to run: give it the images' exact marginals, so that the
    // skipped pixels cost nothing at all ($NMI_HIST_MARGINALS=0: count them into side tables instead)
    static const bool marginals = [] {
      const char* e = getenv("NMI_HIST_MARGINALS");
      return !(e && atoi(e) == 0);
    }();
    if (skipcap && marginals && (f->variant == 0 || f->variant == 8 || f->variant == 9 || f->variant == 4 || f->variant == 5)) {
      CK(c->img_hist.reserve((size_t)(nvl + nwl) * 256));
      c->launches += launch_image_hists(c->renders.p, c->pitch, nvl, c->warps.p, c->pitch, nwl, (uint32_t)c->P,
                                        c->img_hist.p, c->stream);
      a.img_hist = c->img_hist.p;
    }
  }
  c->last_skipcap = use_skip && skipcap;
  const int nl = launch_joint_hist_score(a, c->stream);
  REQUIRE(nl >= 0, NMI_ERR_CUDA, "histogram kernel configuration failed");
  c->launches += nl;
  if (c->timed) CK(cudaEventRecord(c->ev[5], c->stream));
  // an enqueued (multi-GPU) search whose record bins filled up publishes NMI_KEY_RETRY instead
  // of a winner taken from incomplete renders
  launch_argmax(a.scores, d_index, (int)npl, (uint32_t)nP, key_dev ? key_dev : c->key.p,
                key_dev && tiled ? c->bin_total.p + 1 : nullptr, c->stream);
  c->launches++;
  if (c->timed) CK(cudaEventRecord(c->ev[6], c->stream));
  CK(cudaGetLastError());
  return NMI_OK;
}

int search_impl(nmi_ctx* c, const float Twc[16], const nmi_grid* g, const nmi_flags* f, int rank,
                int world, unsigned long long* key_dev, float* scores_dev) {
  REQUIRE(c && Twc, NMI_ERR_INVALID, "null ctx / Twc");
  REQUIRE(valid_grid(g), NMI_ERR_INVALID, "invalid grid (counts must be 1..4096, nP <= 2^26)");
  REQUIRE(valid_flags(f), NMI_ERR_INVALID, "invalid flags (bins 256|64, score 0|1, variant 0..11)");
  REQUIRE(c->has_cam && (c->n_pts > 0 || c->n_tris > 0) && c->has_frame, NMI_ERR_STATE,
          "camera, model and frame must be set before a search");
  CK(cudaSetDevice(c->device));

  const int nS = g->nS[0] * g->nS[1] * g->nS[2], nW = g->nW[0] * g->nW[1] * g->nW[2];
  const size_t nP = (size_t)nS * nW;
  int axis = 0, b = 0, e = 0;
  if (nmi_partition(g, rank, world, &axis, &b, &e) != NMI_OK) {
    set_error("invalid rank / world");
    return NMI_ERR_INVALID;
  }
  const int vb = axis == 0 ? b : 0, ve = axis == 0 ? e : nS;
  const int wb = axis == 1 ? b : 0, we = axis == 1 ? e : nW;
  const int nvl = ve - vb, nwl = we - wb;
  const size_t npl = (size_t)nvl * nwl;
  REQUIRE(nvl <= kMaxViewsPerLaunch * 64, NMI_ERR_INVALID, "too many synthetic views");
  REQUIRE(c->P / 4096 + 2 < 2048, NMI_ERR_INVALID, "image too large for the histogram kernel");

  // ---- host parameter block: centres | minv | pairs | out_index ----
  const size_t off_c = 0;
  const size_t off_m = align_up(off_c + sizeof(float4) * (size_t)nvl, 256);
  const size_t off_p = align_up(off_m + sizeof(float) * 9 * (size_t)nwl, 256);
  const size_t off_i = align_up(off_p + sizeof(int2) * npl, 256);
  const size_t bytes = align_up(off_i + sizeof(uint32_t) * npl, 256);
  if (int rc = params_acquire(c)) return rc;
  if (int rc = ensure_pinned(&c->h_params, &c->h_params_cap, bytes)) return rc;
  if (bytes > c->params.cap) CK(cudaStreamSynchronize(c->stream));  // old block may be in use
  CK(c->params.reserve(bytes));

  ViewConst vc;
  make_view_const(c->cam, Twc, &vc);
  float margin[3] = {0, 0, 0};
  float4* hc = reinterpret_cast<float4*>(c->h_params + off_c);
  for (int s = vb; s < ve; s++) {
    const int sx = s % g->nS[0], sy = (s / g->nS[0]) % g->nS[1], sz = s / (g->nS[0] * g->nS[1]);
    float t[3];
    nmi_cell_translation(Twc, g, sx, sy, sz, t);
    hc[s - vb] = make_float4(Twc[3] + t[0], Twc[7] + t[1], Twc[11] + t[2], 0.0f);
    // camera-frame offset of this view (exact per-view shift of Pc), for the cull margin
    const double ox = (double)vc.r0[0] * t[0] + (double)vc.r0[1] * t[1] + (double)vc.r0[2] * t[2];
    const double oy = (double)vc.r1[0] * t[0] + (double)vc.r1[1] * t[1] + (double)vc.r1[2] * t[2];
    const double oz = (double)vc.r2[0] * t[0] + (double)vc.r2[1] * t[1] + (double)vc.r2[2] * t[2];
    margin[0] = fmaxf(margin[0], (float)fabs(ox));
    margin[1] = fmaxf(margin[1], (float)fabs(oy));
    margin[2] = fmaxf(margin[2], (float)fabs(oz));
  }
  float* hm = reinterpret_cast<float*>(c->h_params + off_m);
  for (int w = wb; w < we; w++) {
    const int wx = w % g->nW[0], wy = (w / g->nW[0]) % g->nW[1], wz = w / (g->nW[0] * g->nW[1]);
    nmi_cell_homography_inv(&c->cam, g, wx, wy, wz, hm + 9 * (size_t)(w - wb));
  }
  fill_pair_schedule(reinterpret_cast<int2*>(c->h_params + off_p), reinterpret_cast<uint32_t*>(c->h_params + off_i),
                     nvl, nwl, nS, vb, wb);

  // ---- buffers ----
  CK(c->renders.reserve((size_t)nvl * c->pitch));
  CK(c->warps.reserve((size_t)nwl * c->pitch));
  CK(c->scores.reserve(nP));
  CK(c->key.reserve(1));
  // Views are rendered in groups whose z-buffers (8 B per pixel and view) fit in L2
  // together (~64 MB of the 126 MB): the z-buffer is written, resolved and reset without
  // ever being streamed through HBM, and only `group` views of it exist.
  const size_t zb_view = c->P * sizeof(unsigned long long);
  const size_t zb_env = [] {  // bytes of z-buffer kept in flight (L2 is 126 MB); read per search (tests shrink it)
    const char* e = getenv("NMI_ZBUF_MB");
    return (size_t)(e && atoi(e) > 0 ? atoi(e) : 64) << 20;
  }();
  // A mesh search keeps, per view in flight, its z-buffer AND its slice of the vertex table in L2, next to the
  // triangle / UV / texture data the shading pass gathers from: half the budget (C3, render stage: groups of
  // 2 / 4 / 8 / 16 views per stream 1.42 / 1.27 / 1.47 / 1.79 ms)
  const size_t zb_budget = c->n_tris ? zb_env / 2 : zb_env;
  static const int mesh_streams = [] {  // view groups of a mesh search in flight, one stream each (1..4)
    const char* e = getenv("NMI_MESH_STREAMS");
    const int k = e ? atoi(e) : 4;  // C3 render stage with 1 / 2 / 3 / 4 streams: 1.44 / 1.28 / 1.26 / 1.25 ms
    return k < 1 ? 1 : (k > nmi_ctx::kMeshStreams ? nmi_ctx::kMeshStreams : k);
  }();
  // mesh: two groups are in flight (one per stream), each with half of the budget
  const bool pingpong = c->n_tris && mesh_streams > 1 && (size_t)nvl * zb_view > zb_budget / 2 && nvl >= 4;
  const int slots = pingpong ? mesh_streams : 1;
  c->mesh_slots = slots;
  int group = (int)(zb_budget / (size_t)slots / (zb_view ? zb_view : 1));
  if (group < 1) group = 1;
  if (group > nvl) group = nvl;
  if (group > kMaxViewsPerLaunch) group = kMaxViewsPerLaunch;
  if (c->n_tris) {  // the mesh rasteriser gives every view of a group a lane: power-of-two groups
    int p2 = 1;
    while (p2 * 2 <= group && p2 < 32) p2 *= 2;
    // ... and a 16-byte entry of the vertex table per (vertex, view of the group): at most 4 GiB
    while (p2 > 1 && c->n_verts * (size_t)p2 * sizeof(int4) * (size_t)slots > (4ull << 30)) p2 /= 2;
    group = p2;
  }
  const bool tiled = !c->n_tris && vc_point_size(c->cam) <= 32;
  if (tiled) {
    if (int rc = ensure_tile_buffers(c, nvl, &group)) return rc;
  } else {
    if (int rc = ensure_zbuf(c, (size_t)group * c->P * (size_t)slots)) return rc;
  }

  c->launches = 0;
  if (c->timed) CK(cudaEventRecord(c->ev[0], c->stream));
  CK(cudaMemcpyAsync(c->params.p, c->h_params, bytes, cudaMemcpyHostToDevice, c->stream));
  if (int rc = params_uploaded(c)) return rc;
  const float4* d_centres = reinterpret_cast<const float4*>(c->params.p + off_c);
  const float* d_minv = reinterpret_cast<const float*>(c->params.p + off_m);
  const int2* d_pairs = reinterpret_cast<const int2*>(c->params.p + off_p);
  const uint32_t* d_index = reinterpret_cast<const uint32_t*>(c->params.p + off_i);

  memcpy(c->Twc, Twc, sizeof(float) * 16);
  // fork: the warps of the camera frame (issue-bound) overlap the render stage (atomic- and
  // latency-bound) on a second stream; both only need the uploaded parameters
  // hot-bin skipping samples every image's dominant grey level (score_pairs_launch): the warps right behind the
  // warp kernel on stream2, the renders behind the render stage -- neither sits between the join and the histogram
  const bool sample_modes = c->hist_skip != 0 && f->bins == 256 && f->bg;
  if (sample_modes) {
    CK(c->img_mode.reserve((size_t)(nvl + nwl)));
    CK(c->hot.reserve(2));
    CK(cudaMemsetAsync(c->hot.p, 0, 2 * sizeof(uint32_t), c->stream));
  }
  CK(cudaEventRecord(c->ev_fork, c->stream));
  CK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
  cudaTextureObject_t frame_tex = 0;
  if (int rc = ensure_frame_texture(c, c->stream2, &frame_tex)) return rc;
  launch_warp(c->frame.p, frame_tex, c->cam.W, c->cam.H, d_minv, nwl, c->warps.p, c->pitch, c->stream2);
  c->launches++;
  if (sample_modes)
    c->launches += launch_image_modes(nullptr, 0, 0, c->warps.p, c->pitch, nwl, (uint32_t)c->P, c->img_mode.p + nvl,
                                      c->hot.p, c->stream2, false);
  CK(cudaEventRecord(c->ev_join, c->stream2));
  if (int rc = cull_model(c, vc, Twc, margin)) return rc;
  if (c->timed) CK(cudaEventRecord(c->ev[1], c->stream));
  if (pingpong) {  // all streams start behind the cull
    CK(cudaEventRecord(c->ev_fork3, c->stream));
    for (int k = 1; k < slots; k++) CK(cudaStreamWaitEvent(c->mesh_stream[k - 1], c->ev_fork3, 0));
  }
  for (int v0 = 0, gi = 0; v0 < nvl; v0 += group, gi++) {
    const int nv = nvl - v0 < group ? nvl - v0 : group;
    if (int rc = draw_views(c, vc, d_centres + v0, nv, c->renders.p + (size_t)v0 * c->pitch, nullptr,
                            pingpong ? gi % slots : 0, pingpong ? (size_t)group : 0))
      return rc;
  }
  if (pingpong)
    for (int k = 1; k < slots; k++) {
      CK(cudaEventRecord(c->ev_join3[k - 1], c->mesh_stream[k - 1]));
      CK(cudaStreamWaitEvent(c->stream, c->ev_join3[k - 1], 0));
    }
  if (sample_modes)
    c->launches += launch_image_modes(c->renders.p, c->pitch, nvl, nullptr, 0, 0, (uint32_t)c->P, c->img_mode.p, c->hot.p,
                                      c->stream, false);
  // stage events: [1] = project + resolve of all view groups (interleaved), [2] = 0
  if (c->timed) CK(cudaEventRecord(c->ev[2], c->stream));
  if (c->timed) CK(cudaEventRecord(c->ev[3], c->stream));
  CK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));  // join: the warps are done
  if (int rc = score_pairs_launch(c, f, nvl, nwl, d_pairs, d_index, npl, nP, key_dev, scores_dev, tiled, sample_modes))
    return rc;

  c->conservative_once = false;
  c->was_sized_retry = c->retry_fullest_once != 0;
  c->retry_fullest_once = 0;
  c->has_search = true;
  c->grid = *g;
  c->nvl = nvl;
  c->nwl = nwl;
  c->v_begin = vb;
  c->w_begin = wb;
  c->off_centres = off_c;
  return NMI_OK;
}

}  // namespace

extern "C" {

const char* nmi_last_error(void) { return g_err.c_str(); }

int nmi_ctx_create(int device, nmi_ctx** out) {
  REQUIRE(out, NMI_ERR_INVALID, "null out");
  *out = nullptr;
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  REQUIRE(device >= 0 && device < ndev, NMI_ERR_CUDA, "no such CUDA device");
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  REQUIRE(prop.major == 10, NMI_ERR_CUDA,
          "this library is built for sm_100a (B200) only; no fallback path exists");
  CK(cudaSetDevice(device));
  nmi_ctx* c = new nmi_ctx();
  c->device = device;
  // the main stream outranks the warp stream: the small, latency-bound cull launches get their CTAs
  // placed as soon as SM resources free up instead of queueing behind the (issue-bound) warp kernel
  // that was launched first ($NMI_STREAM_PRIO=0: equal priorities)
  {
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    const char* e = getenv("NMI_STREAM_PRIO");
    const bool prio = !(e && atoi(e) == 0);
    CK(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio ? hi : lo));
    CK(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, lo));
  }
  for (auto& st : c->mesh_stream) CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&c->ev_fork3, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_frame_order, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_frame_up, cudaEventDisableTiming));
  for (auto& ev : c->ev_join3) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  for (auto& e : c->ev) CK(cudaEventCreate(&e));
  CK(cudaEventCreateWithFlags(&c->ev_params, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_feedback, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_hot, cudaEventDisableTiming));
  CK(cudaMallocHost(&c->h_feedback, 8 * sizeof(uint32_t)));
  memset(c->h_feedback, 0, 8 * sizeof(uint32_t));
  CK(cudaMallocHost(&c->h_score, 4 * sizeof(float)));
  CK(c->counter.reserve(1));
  CK(c->key.reserve(1));
  CK(c->one_score.reserve(1));
  CK(c->zero_pair.reserve(1));
  CK(cudaMemset(c->zero_pair.p, 0, sizeof(int2)));
  c->timed = true;
  if (const char* e = getenv("NMI_WARP_TEX")) c->use_warp_tex = atoi(e) != 0;
  if (const char* e = getenv("NMI_BLOCK_CULL")) c->use_block_cull = atoi(e) != 0;
  if (const char* e = getenv("NMI_HIST_SKIP")) {
    const int m = atoi(e);
    if (m >= 0 && m <= 2) c->hist_skip = m;
  }
  *out = c;
  return NMI_OK;
}

void nmi_ctx_destroy(nmi_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  c->pts.release(); c->aabb.release(); c->orig.release(); c->tag.release(); c->val.release(); c->mverts.release(); c->mtris.release();
  c->mtri_orig.release(); c->mslots.release(); c->mtris_o.release(); c->muv.release(); c->mtv.release(); c->mvflag.release(); c->mluma.release(); c->bin_offsets.release(); c->bin_cursor.release();
  c->bin_total.release(); c->records.release(); c->cpts.release(); c->cidx.release(); c->counter.release(); c->block_counts.release();
  c->frame.release(); c->zbuf.release(); c->renders.release(); c->warps.release();
  c->scores.release(); c->key.release(); c->params.release(); c->one_render.release();
  c->img_mode.release(); c->img_hist.release(); c->hot.release(); c->term_tab.release(); c->one_warp.release(); c->winners.release(); c->dumpJ.release(); c->dumpH.release();
  c->one_score.release(); c->zero_pair.release();
  for (int i = 0; i < 2; i++) {
    if (c->h_frame[i]) cudaFreeHost(c->h_frame[i]);
    if (c->ev_frame[i]) cudaEventDestroy(c->ev_frame[i]);
  }
  if (c->h_params) cudaFreeHost(c->h_params);
  if (c->h_score) cudaFreeHost(c->h_score);
  for (auto& e : c->ev) if (e) cudaEventDestroy(e);
  if (c->ev_params) cudaEventDestroy(c->ev_params);
  if (c->ev_feedback) cudaEventDestroy(c->ev_feedback);
  if (c->ev_hot) cudaEventDestroy(c->ev_hot);
  if (c->frame_tex) cudaDestroyTextureObject(c->frame_tex);
  if (c->frame_arr) cudaFreeArray(c->frame_arr);
  if (c->h_feedback) cudaFreeHost(c->h_feedback);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  if (c->ev_fork3) cudaEventDestroy(c->ev_fork3);
  if (c->ev_frame_order) cudaEventDestroy(c->ev_frame_order);
  if (c->ev_frame_up) cudaEventDestroy(c->ev_frame_up);
  for (auto ev : c->ev_join3)
    if (ev) cudaEventDestroy(ev);
  for (auto st : c->mesh_stream)
    if (st) cudaStreamDestroy(st);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

// An enqueued search overflowed its bins (pinned feedback words of that search are complete: the stream has
// been synchronised).  The resolve kernel has reported the fullest bin the search WANTED: the redo runs in a
// single pass with bins of that size; only a redo that overflows again gets the two-pass sizing.
static void arm_enqueued_retry(nmi_ctx* c) {
  if (c->bin_cap != 0 && c->h_feedback[3] > 0 && !c->was_sized_retry)
    c->retry_fullest_once = c->h_feedback[3];
  else
    c->conservative_once = true;
  c->h_feedback[2] = 0;
}

void* nmi_ctx_stream(nmi_ctx* c) { return c ? (void*)c->stream : nullptr; }

int nmi_ctx_set_hist_skip(nmi_ctx* c, int mode) {
  REQUIRE(c && mode >= 0 && mode <= 2, NMI_ERR_INVALID, "mode must be 0 (never), 1 (automatic) or 2 (always)");
  c->hist_skip = mode;
  return NMI_OK;
}

int nmi_ctx_sync(nmi_ctx* c) {
  REQUIRE(c, NMI_ERR_INVALID, "null ctx");
  CK(cudaStreamSynchronize(c->stream));
  if (c->frame_up_pending) {  // a frame upload without a search behind it: the caller's pinned buffer is free after this
    CK(cudaEventSynchronize(c->ev_frame_up));
    c->frame_up_pending = false;
  }
  if (c->feedback_pending && c->h_feedback && c->h_feedback[2] != 0) {
    arm_enqueued_retry(c);  // the next search gets bins of the size this one wanted
    set_error("tile renderer record buffer overflow in an enqueued search: its renders are incomplete; "
              "re-enqueue it");
    return NMI_ERR_CUDA;
  }
  return NMI_OK;
}

int nmi_set_camera(nmi_ctx* c, const nmi_camera* cam) {
  REQUIRE(c && cam, NMI_ERR_INVALID, "null ctx / camera");
  REQUIRE(cam->W > 0 && cam->H > 0 && (size_t)cam->W * cam->H < (1ull << 31), NMI_ERR_INVALID,
          "bad image size");
  REQUIRE(cam->cx != 0 && cam->cy != 0 && cam->fx != 0 && cam->fy != 0, NMI_ERR_INVALID,
          "bad intrinsics");
  REQUIRE(cam->zn > 0 && cam->zf > cam->zn, NMI_ERR_INVALID, "bad near/far planes");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  if (c->has_cam && (cam->W != c->cam.W || cam->H != c->cam.H)) {
    c->has_frame = false;
    c->has_search = false;
  }
  c->cam = *cam;
  c->feedback_pending = false;
  c->P = (size_t)cam->W * cam->H;
  c->pitch = img_pitch(c->P);
  c->has_cam = true;
  return NMI_OK;
}

// Model upload.  The cloud is stored in 3-D Morton order (a load-time permutation) so that
// neighbouring threads of the projection kernel splat neighbouring pixels: z-buffer lines
// are then fetched from HBM about once per search instead of once per fragment.  The
// ORIGINAL index travels with every point -- it is what the z-buffer key carries, so the
// GL "earlier primitive wins a depth tie" rule (rendering.hpp:297) is unaffected.
static inline uint32_t spread3(uint32_t v) {  // 10 bits -> every third bit
  v &= 0x3FFu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

int nmi_set_points(nmi_ctx* c, const float* xyzi, size_t n) {
  REQUIRE(c && xyzi && n > 0 && n < 0xFFFFFFFFull, NMI_ERR_INVALID, "bad point cloud");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (size_t i = 0; i < n; i++)
    for (int k = 0; k < 3; k++) {
      const float v = xyzi[4 * i + k];
      if (v < lo[k]) lo[k] = v;
      if (v > hi[k]) hi[k] = v;
    }
  // cubic cells (one scale for all axes): a thin axis only contributes its coarse bits
  float ext = 0.0f;
  for (int k = 0; k < 3; k++) ext = fmaxf(ext, hi[k] - lo[k]);
  float scale[3];
  for (int k = 0; k < 3; k++) scale[k] = ext > 0.0f ? 1023.0f / ext : 0.0f;
  std::vector<uint64_t> order(n);
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)n; i++) {
    uint32_t q[3];
    for (int k = 0; k < 3; k++) {
      const float f = (xyzi[4 * i + k] - lo[k]) * scale[k];
      q[k] = f >= 0.0f ? (f < 1023.0f ? (uint32_t)f : 1023u) : 0u;  // NaN -> 0
    }
    const uint64_t m = spread3(q[0]) | (spread3(q[1]) << 1) | (spread3(q[2]) << 2);
    order[i] = (m << 32) | (uint64_t)i;
  }
  std::sort(order.begin(), order.end());
  std::vector<float> sorted(4 * n);
  std::vector<uint32_t> orig(n);
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)n; i++) {
    const uint32_t src = (uint32_t)(order[i] & 0xFFFFFFFFu);
    orig[i] = src;
    memcpy(&sorted[4 * i], &xyzi[4 * (size_t)src], 4 * sizeof(float));
  }
  // axis-aligned box of every kCullBlock consecutive points (the cull kernels skip whole blocks)
  const size_t nblk = (n + kCullBlock - 1) / kCullBlock;
  std::vector<float> boxes(6 * nblk);
#pragma omp parallel for schedule(static)
  for (long long b = 0; b < (long long)nblk; b++) {
    float blo[3] = {INFINITY, INFINITY, INFINITY}, bhi[3] = {-INFINITY, -INFINITY, -INFINITY};
    const size_t e = std::min(n, (size_t)(b + 1) * kCullBlock);
    for (size_t i = (size_t)b * kCullBlock; i < e; i++)
      for (int k = 0; k < 3; k++) {
        blo[k] = fminf(blo[k], sorted[4 * i + k]);  // fminf / fmaxf ignore NaN
        bhi[k] = fmaxf(bhi[k], sorted[4 * i + k]);
      }
    for (int k = 0; k < 3; k++) {
      boxes[6 * b + k] = blo[k];
      boxes[6 * b + 3 + k] = bhi[k];
    }
  }
  CK(c->aabb.reserve(6 * nblk));
  CK(cudaMemcpyAsync(c->aabb.p, boxes.data(), 6 * nblk * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  CK(c->pts.reserve(n));
  CK(c->orig.reserve(n));
  CK(c->val.reserve(n));
  CK(c->cpts.reserve(n));
  CK(c->cidx.reserve(n));
  CK(c->block_counts.reserve((n + 255) / 256 + 1));  // >= CTAs of either cull kernel
  CK(cudaMemcpyAsync(c->pts.p, sorted.data(), n * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->orig.p, orig.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  CK(c->tag.reserve(n));
  c->packed_value = n < (1u << 24);
  launch_intensity_u8(c->pts.p, c->orig.p, c->val.p, c->tag.p, c->packed_value, n, c->stream);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  c->n_pts = n;
  c->n_tris = 0;
  c->feedback_pending = false;  // record-count feedback belongs to the previous model
  c->has_search = false;
  return NMI_OK;
}

int nmi_set_mesh(nmi_ctx* c, const float* verts, size_t nv, const uint32_t* tris, size_t nt) {
  REQUIRE(c && verts && tris && nv > 0 && nt > 0 && nv < 0xFFFFFFFFull && nt < 0xFFFFFFFFull,
          NMI_ERR_INVALID, "bad mesh");
  for (size_t i = 0; i < 3 * nt; i++) REQUIRE(tris[i] < nv, NMI_ERR_INVALID, "triangle index out of range");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  // Morton order of the triangle centroids (load-time): neighbouring threads raster neighbouring
  // pixels.  The ORIGINAL triangle index is the z-buffer key's tie-break (GL draw order).
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (size_t i = 0; i < nv; i++)
    for (int k = 0; k < 3; k++) {
      const float v = verts[4 * i + k];
      if (v < lo[k]) lo[k] = v;
      if (v > hi[k]) hi[k] = v;
    }
  float ext = 0.0f;
  for (int k = 0; k < 3; k++) ext = fmaxf(ext, hi[k] - lo[k]);
  const float scale = ext > 0.0f ? 1023.0f / ext : 0.0f;
  std::vector<uint64_t> order(nt);
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)nt; i++) {
    uint32_t q[3];
    for (int k = 0; k < 3; k++) {
      const float g = (verts[4 * (size_t)tris[3 * i] + k] + verts[4 * (size_t)tris[3 * i + 1] + k] +
                       verts[4 * (size_t)tris[3 * i + 2] + k]) * (1.0f / 3.0f);
      const float f = (g - lo[k]) * scale;
      q[k] = f >= 0.0f ? (f < 1023.0f ? (uint32_t)f : 1023u) : 0u;
    }
    const uint64_t m = spread3(q[0]) | (spread3(q[1]) << 1) | (spread3(q[2]) << 2);
    order[i] = (m << 32) | (uint64_t)i;
  }
  std::sort(order.begin(), order.end());
  std::vector<uint32_t> sorted(3 * nt), orig(nt);
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)nt; i++) {
    const uint32_t src = (uint32_t)(order[i] & 0xFFFFFFFFu);
    orig[i] = src;
    for (int k = 0; k < 3; k++) sorted[3 * i + k] = tris[3 * (size_t)src + k];
  }
  CK(c->mverts.reserve(nv));
  CK(c->mtris.reserve(nt));
  CK(c->mtri_orig.reserve(nt));
  CK(c->mslots.reserve(nt));
  CK(c->mvflag.reserve(nv));
  CK(c->val.reserve(nt));
  CK(c->block_counts.reserve((nt + 255) / 256 + 1));
  CK(cudaMemcpyAsync(c->mverts.p, verts, nv * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->mtris.p, sorted.data(), nt * sizeof(uint3), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->mtri_orig.p, orig.data(), nt * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  launch_mesh_values(c->mverts.p, c->mtris.p, c->mtri_orig.p, c->val.p, (uint32_t)nt, c->stream);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  c->n_tris = nt;
  c->n_verts = nv;
  c->packed_value = false;  // mesh keys carry the plain triangle index
  c->n_pts = 0;
  c->tex_w = c->tex_h = 0;  // flat grey of the first vertex until nmi_set_mesh_textured says otherwise
  c->has_search = false;
  return NMI_OK;
}

int nmi_set_mesh_textured(nmi_ctx* c, const float* verts, size_t nv, const uint32_t* tris, size_t nt,
                          const float* corner_uv, const uint8_t* texture, int tex_w, int tex_h) {
  REQUIRE(c && corner_uv && texture && tex_w > 0 && tex_h > 0 && (size_t)tex_w * tex_h < (1ull << 31), NMI_ERR_INVALID,
          "bad texture / UVs");
  if (int rc = nmi_set_mesh(c, verts, nv, tris, nt)) return rc;
  const size_t ntex = (size_t)tex_w * tex_h;
  CK(c->mtris_o.reserve(nt));
  CK(c->muv.reserve(2 * nt));
  CK(c->mluma.reserve(ntex));
  DevBuf<uint8_t> raw;
  CK(raw.reserve(3 * ntex));
  // padded to 16-byte entries: the shading kernel fetches a triangle's corners / UVs with one / two loads
  std::vector<uint4> tri4(nt);
  std::vector<float4> uv4(2 * nt);
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)nt; i++) {
    tri4[i] = make_uint4(tris[3 * i], tris[3 * i + 1], tris[3 * i + 2], 0u);
    const float* q = corner_uv + 6 * (size_t)i;
    uv4[2 * i] = make_float4(q[0], q[1], q[2], q[3]);
    uv4[2 * i + 1] = make_float4(q[4], q[5], 0.0f, 0.0f);
  }
  CK(cudaMemcpyAsync(c->mtris_o.p, tri4.data(), nt * sizeof(uint4), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->muv.p, uv4.data(), 2 * nt * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(raw.p, texture, 3 * ntex, cudaMemcpyHostToDevice, c->stream));
  launch_mesh_luma(raw.p, c->mluma.p, ntex, c->stream);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  raw.release();
  c->tex_w = tex_w;
  c->tex_h = tex_h;
  return NMI_OK;
}

int nmi_set_points_device(nmi_ctx* c, const void* xyzi_dev, size_t n) {
  REQUIRE(c && xyzi_dev && n > 0 && n < 0xFFFFFFFFull, NMI_ERR_INVALID, "bad point cloud");
  CK(cudaSetDevice(c->device));
  // load-time path: bring the cloud to the host once for the Morton ordering
  std::vector<float> host(4 * n);
  CK(cudaMemcpy(host.data(), xyzi_dev, n * sizeof(float4), cudaMemcpyDeviceToHost));
  return nmi_set_points(c, host.data(), n);
}

int nmi_set_frame(nmi_ctx* c, const uint8_t* gray, int W, int H) {
  REQUIRE(c && gray, NMI_ERR_INVALID, "null ctx / frame");
  REQUIRE(c->has_cam, NMI_ERR_STATE, "set the camera first");
  REQUIRE(W == c->cam.W && H == c->cam.H, NMI_ERR_INVALID, "frame size != camera size");
  CK(cudaSetDevice(c->device));
  CK(c->frame.reserve(c->pitch));
  cudaPointerAttributes attr{};
  const bool pinned = cudaPointerGetAttributes(&attr, gray) == cudaSuccess &&
                      attr.type == cudaMemoryTypeHost;
  cudaGetLastError();  // pageable pointers may leave a sticky-free error behind
  // the upload runs on stream2, behind everything already enqueued on the main stream
  CK(cudaEventRecord(c->ev_frame_order, c->stream));
  CK(cudaStreamWaitEvent(c->stream2, c->ev_frame_order, 0));
  if (pinned) {
    // caller's buffer is page-locked: DMA straight from it (caller keeps it alive until sync)
    CK(cudaMemcpyAsync(c->frame.p, gray, c->P, cudaMemcpyHostToDevice, c->stream2));
  } else {
    // pageable -> pinned staging so the H2D copy is a true async DMA on our stream
    const int slot = c->frame_slot;
    c->frame_slot ^= 1;
    if (!c->ev_frame[slot]) CK(cudaEventCreateWithFlags(&c->ev_frame[slot], cudaEventDisableTiming));
    CK(cudaEventSynchronize(c->ev_frame[slot]));  // the upload that last used this buffer is done
    if (int rc = ensure_pinned(&c->h_frame[slot], &c->h_frame_cap[slot], c->P)) return rc;
    memcpy(c->h_frame[slot], gray, c->P);
    CK(cudaMemcpyAsync(c->frame.p, c->h_frame[slot], c->P, cudaMemcpyHostToDevice, c->stream2));
    CK(cudaEventRecord(c->ev_frame[slot], c->stream2));
  }
  CK(cudaEventRecord(c->ev_frame_up, c->stream2));
  c->frame_up_pending = true;
  c->frame_tex_dirty = true;
  c->has_frame = true;
  return NMI_OK;
}

int nmi_set_frame_device(nmi_ctx* c, const void* gray_dev, int W, int H) {
  REQUIRE(c && gray_dev, NMI_ERR_INVALID, "null ctx / frame");
  REQUIRE(c->has_cam, NMI_ERR_STATE, "set the camera first");
  REQUIRE(W == c->cam.W && H == c->cam.H, NMI_ERR_INVALID, "frame size != camera size");
  CK(cudaSetDevice(c->device));
  CK(c->frame.reserve(c->pitch));
  if (c->frame_up_pending) CK(cudaStreamWaitEvent(c->stream, c->ev_frame_up, 0));  // an earlier host upload lands first
  CK(cudaMemcpyAsync(c->frame.p, gray_dev, c->P, cudaMemcpyDeviceToDevice, c->stream));
  c->frame_tex_dirty = true;
  c->has_frame = true;
  return NMI_OK;
}

int nmi_search(nmi_ctx* c, const float Twc[16], const nmi_grid* g, const nmi_flags* f,
               nmi_result* out, float* scores_host) {
  REQUIRE(out, NMI_ERR_INVALID, "null result");
  if (int rc = search_impl(c, Twc, g, f, 0, 1, nullptr, nullptr)) return rc;
  unsigned long long key = 0;
  uint32_t tile_state[3] = {0, 0, 0};
  CK(cudaMemcpyAsync(&key, c->key.p, sizeof key, cudaMemcpyDeviceToHost, c->stream));
  if (c->bin_total.p)
    CK(cudaMemcpyAsync(tile_state, c->bin_total.p, sizeof tile_state, cudaMemcpyDeviceToHost, c->stream));
  if (scores_host) {
    const size_t nP = (size_t)c->nvl * c->nwl;
    CK(cudaMemcpyAsync(scores_host, c->scores.p, nP * sizeof(float), cudaMemcpyDeviceToHost,
                       c->stream));
  }
  CK(cudaStreamSynchronize(c->stream));
  if (tile_state[1] != 0 && !c->force_conservative) {
    // the bin capacity guessed from the previous searches was too small for this pose.  The resolve kernel
    // has reported the fullest bin this search wanted: redo it in a single pass with bins of that size ...
    if (c->retry_fullest == 0 && tile_state[2] > 0 && c->bin_cap != 0) {  // (the two-pass layout reports clamped counts)
      c->retry_fullest = tile_state[2];
      const int rc2 = nmi_search(c, Twc, g, f, out, scores_host);
      c->retry_fullest = 0;
      return rc2;
    }
    // ... and only if that fails too (it cannot, unless the buffer limit is hit) with the pose-independent sizing
    c->force_conservative = true;
    const int rc2 = nmi_search(c, Twc, g, f, out, scores_host);
    c->force_conservative = false;
    return rc2;
  }
  REQUIRE(tile_state[1] == 0, NMI_ERR_CUDA, "tile renderer record buffer overflow (model too dense for 2 GiB of records)");
  float ms = 0;
  if (c->timed) cudaEventElapsedTime(&ms, c->ev[0], c->ev[6]);
  const int rc = nmi_decode_key(g, key, out);
  out->gpu_ms = ms;
  if (rc == NMI_ERR_NO_WINNER) set_error("every score is negative: no winner");
  return rc;
}

void* nmi_ctx_key_buffer(nmi_ctx* c) {
  if (!c || cudaSetDevice(c->device) != cudaSuccess || c->xkey.reserve(1) != cudaSuccess) return nullptr;
  return c->xkey.p;
}

int nmi_read_key(nmi_ctx* c, const void* key_dev, uint64_t* key) {
  REQUIRE(c && key_dev && key, NMI_ERR_INVALID, "null argument");
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(key, key_dev, sizeof *key, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  // a local overflow is also visible in the pinned feedback words: size the next search exactly
  if (c->feedback_pending && c->h_feedback && c->h_feedback[2] != 0) arm_enqueued_retry(c);
  return NMI_OK;
}

int nmi_search_enqueue(nmi_ctx* c, const float Twc[16], const nmi_grid* g, const nmi_flags* f,
                       int rank, int world, void* key_dev, void* scores_dev) {
  REQUIRE(key_dev, NMI_ERR_INVALID, "null key_dev");
  return search_impl(c, Twc, g, f, rank, world, static_cast<unsigned long long*>(key_dev),
                     static_cast<float*>(scores_dev));
}

// ---- stage-level API ------------------------------------------------------

int nmi_render_at(nmi_ctx* c, const float Twc[16], const float t[3], unsigned int* handle) {
  REQUIRE(c && Twc && t && handle, NMI_ERR_INVALID, "null argument");
  REQUIRE(c->has_cam && (c->n_pts > 0 || c->n_tris > 0), NMI_ERR_STATE, "camera and model must be set");
  CK(cudaSetDevice(c->device));
  ViewConst vc;
  make_view_const(c->cam, Twc, &vc);
  if (int rc = ensure_pinned(&c->h_params, &c->h_params_cap, 4096)) return rc;
  CK(c->params.reserve(4096));
  CK(c->one_render.reserve(c->pitch));
  if (!c->n_tris && vc.s <= 32) {
    int g1 = 1;
    if (int rc = ensure_tile_buffers(c, 1, &g1)) return rc;
  } else {
    if (int rc = ensure_zbuf(c, c->P)) return rc;
  }
  CK(cudaStreamSynchronize(c->stream));
  float4 centre = make_float4(Twc[3] + t[0], Twc[7] + t[1], Twc[11] + t[2], 0.0f);
  memcpy(c->h_params, &centre, sizeof centre);
  CK(cudaMemcpyAsync(c->params.p, c->h_params, sizeof centre, cudaMemcpyHostToDevice, c->stream));
  float margin[3];
  margin[0] = (float)fabs((double)vc.r0[0] * t[0] + (double)vc.r0[1] * t[1] + (double)vc.r0[2] * t[2]);
  margin[1] = (float)fabs((double)vc.r1[0] * t[0] + (double)vc.r1[1] * t[1] + (double)vc.r1[2] * t[2]);
  margin[2] = (float)fabs((double)vc.r2[0] * t[0] + (double)vc.r2[1] * t[1] + (double)vc.r2[2] * t[2]);
  memcpy(c->Twc, Twc, sizeof(float) * 16);
  c->has_search = false;  // the cull list no longer matches the last search
  const bool timed = c->timed;
  c->timed = false;
  const int rc = render_views(c, vc, reinterpret_cast<const float4*>(c->params.p), 1, margin, true,
                              c->one_render.p, nullptr);
  c->timed = timed;
  if (rc) return rc;
  *handle = 1;  // the one "rendered texture" (rendering.hpp:341 creates exactly one)
  return NMI_OK;
}

int nmi_render_cell(nmi_ctx* c, const float Twc[16], const nmi_grid* g, int sx, int sy, int sz,
                    unsigned int* handle) {
  REQUIRE(c && Twc && handle, NMI_ERR_INVALID, "null argument");
  REQUIRE(valid_grid(g), NMI_ERR_INVALID, "invalid grid");
  REQUIRE(sx >= 0 && sx < g->nS[0] && sy >= 0 && sy < g->nS[1] && sz >= 0 && sz < g->nS[2],
          NMI_ERR_INVALID, "cell out of range");
  float t[3];
  nmi_cell_translation(Twc, g, sx, sy, sz, t);
  return nmi_render_at(c, Twc, t, handle);
}

int nmi_warp_cells(nmi_ctx* c, const nmi_grid* g) {
  REQUIRE(c, NMI_ERR_INVALID, "null ctx");
  REQUIRE(valid_grid(g), NMI_ERR_INVALID, "invalid grid");
  REQUIRE(c->has_cam && c->has_frame, NMI_ERR_STATE, "camera and frame must be set");
  CK(cudaSetDevice(c->device));
  const int nW = g->nW[0] * g->nW[1] * g->nW[2];
  const size_t bytes = sizeof(float) * 9 * (size_t)nW;
  if (int rc = ensure_pinned(&c->h_params, &c->h_params_cap, bytes)) return rc;
  CK(c->params.reserve(bytes));
  CK(c->warps.reserve((size_t)nW * c->pitch));
  CK(cudaStreamSynchronize(c->stream));
  float* hm = reinterpret_cast<float*>(c->h_params);
  for (int w = 0; w < nW; w++) {
    const int wx = w % g->nW[0], wy = (w / g->nW[0]) % g->nW[1], wz = w / (g->nW[0] * g->nW[1]);
    nmi_cell_homography_inv(&c->cam, g, wx, wy, wz, hm + 9 * (size_t)w);
  }
  CK(cudaMemcpyAsync(c->params.p, c->h_params, bytes, cudaMemcpyHostToDevice, c->stream));
  if (c->frame_up_pending) CK(cudaStreamWaitEvent(c->stream, c->ev_frame_up, 0));  // the frame is uploaded on stream2
  cudaTextureObject_t frame_tex = 0;
  if (int rc = ensure_frame_texture(c, c->stream, &frame_tex)) return rc;
  launch_warp(c->frame.p, frame_tex, c->cam.W, c->cam.H, reinterpret_cast<const float*>(c->params.p), nW,
              c->warps.p, c->pitch, c->stream);
  CK(cudaGetLastError());
  c->has_search = false;
  c->nwl = nW;
  c->w_begin = 0;
  c->grid = *g;
  return NMI_OK;
}

int nmi_warp_ptr(nmi_ctx* c, const nmi_grid* g, int wx, int wy, int wz, void** dev_ptr) {
  REQUIRE(c && dev_ptr && valid_grid(g), NMI_ERR_INVALID, "bad argument");
  REQUIRE(wx >= 0 && wx < g->nW[0] && wy >= 0 && wy < g->nW[1] && wz >= 0 && wz < g->nW[2],
          NMI_ERR_INVALID, "cell out of range");
  const int w = (wz * g->nW[1] + wy) * g->nW[0] + wx;
  REQUIRE(c->warps.p && w - c->w_begin >= 0 && w - c->w_begin < c->nwl, NMI_ERR_STATE,
          "warps not computed for this cell");
  *dev_ptr = c->warps.p + (size_t)(w - c->w_begin) * c->pitch;
  return NMI_OK;
}

// One evaluation.  J / HA / HB: optional DEVICE destinations of the integer histograms
// (bins*bins, bins, bins u32); when only some are wanted the rest land in context scratch.
static int eval_images(nmi_ctx* c, const uint8_t* render, const uint8_t* warped, uint32_t npix,
                       const nmi_flags* f, uint32_t* J, uint32_t* HA, uint32_t* HB, float* score_host,
                       int path = 0) {
  HistArgs a{};
  a.renders = render;
  a.warps = warped;
  a.render_pitch = a.warp_pitch = 0;
  a.pairs = c->zero_pair.p;
  a.out_index = nullptr;
  a.npairs = 1;
  a.npix = npix;
  a.length = npix;
  a.bins = f->bins;
  a.bg = f->bg;
  a.mode = f->score_mode;
  a.variant = f->variant;
  // with a host destination the kernel writes the score straight into pinned host memory (one store over PCIe
  // instead of a copy call after the kernel: several microseconds of a 40 us call)
  a.scores = score_host ? c->h_score : c->one_score.p;
  a.force_batched = path == 1;  // path 1 = the persistent build a search launches, not the single-evaluation cluster kernel
  if (int rc = ensure_term_table(c, npix)) return rc;
  a.term_tab = c->term_tab.p;
  if (J || HA || HB) {
    if (!J) CK(c->dumpJ.reserve(65536));
    if (!HA || !HB) CK(c->dumpH.reserve(512));
    a.dumpJ = J ? J : c->dumpJ.p;
    a.dumpHA = HA ? HA : c->dumpH.p;
    a.dumpHB = HB ? HB : c->dumpH.p + 256;
  }
  if (path != 1 && (c->hist_skip != 0 || path == 2) && f->bins == 256 && f->bg && ((uintptr_t)render % 16) == 0 &&
      ((uintptr_t)warped % 16) == 0) {
    a.skip_mode = path == 2 ? 2 : c->hist_skip;
    a.skipcap = true;
    a.nrenders = 1;
    if (hist_uses_cluster(a)) {
      a.sample_in_kernel = true;  // the cluster kernel samples the two dominant grey levels itself: one launch per call
    } else {
      // single evaluation: always the build with the side tables, the kernel decides
      CK(c->img_mode.reserve(2));
      CK(c->hot.reserve(2));
      launch_image_modes(render, 0, 1, warped, 0, 1, npix, c->img_mode.p, c->hot.p, c->stream);
      a.img_mode = c->img_mode.p;
      a.sample_total = image_mode_sample_total(npix);
      const char* e = getenv("NMI_HIST_MARGINALS");
      if (!(e && atoi(e) == 0)) {  // as in a batched search: skipped pixels come back through the marginals
        CK(c->img_hist.reserve(512));
        launch_image_hists(render, 0, 1, warped, 0, 1, npix, c->img_hist.p, c->stream);
        a.img_hist = c->img_hist.p;
      }
    }
  }
  REQUIRE(launch_joint_hist_score(a, c->stream) >= 0, NMI_ERR_CUDA,
          "histogram kernel configuration failed");
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  if (score_host) *score_host = c->h_score[0];
  return NMI_OK;
}

int nmi_eval_pair(nmi_ctx* c, const void* warped_dev, unsigned int handle, int W, int H,
                  const nmi_flags* f, float* score_host) {
  REQUIRE(c && warped_dev && score_host, NMI_ERR_INVALID, "null argument");
  REQUIRE(valid_flags(f), NMI_ERR_INVALID, "invalid flags");
  REQUIRE(c->has_cam && W == c->cam.W && H == c->cam.H, NMI_ERR_INVALID, "size != camera size");
  REQUIRE(handle == 1 && c->one_render.p, NMI_ERR_STATE, "unknown render handle");
  REQUIRE(c->P / 4096 + 2 < 2048, NMI_ERR_INVALID, "image too large for the histogram kernel");
  CK(cudaSetDevice(c->device));
  const uint8_t* wp = static_cast<const uint8_t*>(warped_dev);
  const bool ours = c->warps.p && wp >= c->warps.p && wp < c->warps.p + c->warps.cap &&
                    ((size_t)(wp - c->warps.p) % c->pitch) == 0;
  if (!ours) {
    // borrowed buffer (e.g. a GpuMat): copy into a padded, 128 B aligned slot so the
    // TMA loads of the last chunk stay inside memory we own
    CK(c->one_warp.reserve(c->pitch));
    CK(cudaMemcpyAsync(c->one_warp.p, wp, c->P, cudaMemcpyDeviceToDevice, c->stream));
    wp = c->one_warp.p;
  }
  return eval_images(c, c->one_render.p, wp, (uint32_t)c->P, f, nullptr, nullptr, nullptr, score_host);
}

int nmi_eval_pair_dev(nmi_ctx* c, const void* warped_dev, unsigned int handle, int W, int H,
                      const nmi_flags* f, uint32_t* J_dev, uint32_t* HA_dev, uint32_t* HB_dev,
                      float* score_host) {
  REQUIRE(c && warped_dev, NMI_ERR_INVALID, "null argument");
  REQUIRE(valid_flags(f), NMI_ERR_INVALID, "invalid flags");
  REQUIRE(c->has_cam && W == c->cam.W && H == c->cam.H, NMI_ERR_INVALID, "size != camera size");
  REQUIRE(handle == 1 && c->one_render.p, NMI_ERR_STATE, "unknown render handle");
  REQUIRE(c->P / 4096 + 2 < 2048, NMI_ERR_INVALID, "image too large for the histogram kernel");
  CK(cudaSetDevice(c->device));
  CK(c->one_warp.reserve(c->pitch));
  CK(cudaMemcpyAsync(c->one_warp.p, warped_dev, c->P, cudaMemcpyDeviceToDevice, c->stream));
  return eval_images(c, c->one_render.p, c->one_warp.p, (uint32_t)c->P, f, J_dev, HA_dev, HB_dev, score_host);
}

int nmi_import_render(nmi_ctx* c, const void* render_dev, size_t pitch_bytes, int W, int H, int bottom_up,
                      unsigned int* handle) {
  REQUIRE(c && render_dev && handle, NMI_ERR_INVALID, "null argument");
  REQUIRE(c->has_cam && W == c->cam.W && H == c->cam.H, NMI_ERR_INVALID, "size != camera size");
  REQUIRE(pitch_bytes >= (size_t)W, NMI_ERR_INVALID, "pitch smaller than a row");
  CK(cudaSetDevice(c->device));
  CK(c->one_render.reserve(c->pitch));
  launch_copy_rows(static_cast<const uint8_t*>(render_dev), pitch_bytes, c->one_render.p, W, H,
                   bottom_up != 0, c->stream);
  CK(cudaGetLastError());
  c->has_search = false;  // nmi_get_render(0) now returns the imported image
  *handle = 1;
  return NMI_OK;
}

// ---- parity read-backs ------------------------------------------------------

int nmi_get_render(nmi_ctx* c, int s, uint8_t* host) {
  REQUIRE(c && host, NMI_ERR_INVALID, "null argument");
  CK(cudaSetDevice(c->device));
  const uint8_t* src = nullptr;
  if (c->has_search) {
    REQUIRE(s - c->v_begin >= 0 && s - c->v_begin < c->nvl, NMI_ERR_INVALID, "view not on this rank");
    src = c->renders.p + (size_t)(s - c->v_begin) * c->pitch;
  } else {
    REQUIRE(c->one_render.p && s == 0, NMI_ERR_STATE, "no render available");
    src = c->one_render.p;
  }
  CK(cudaMemcpyAsync(host, src, c->P, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return NMI_OK;
}

int nmi_get_winners(nmi_ctx* c, int s, uint32_t* host) {
  REQUIRE(c && host, NMI_ERR_INVALID, "null argument");
  REQUIRE(c->has_search, NMI_ERR_STATE, "no search has run");
  REQUIRE(s - c->v_begin >= 0 && s - c->v_begin < c->nvl, NMI_ERR_INVALID, "view not on this rank");
  CK(cudaSetDevice(c->device));
  // re-project this one view from the survivors of the last cull, keeping the indices
  ViewConst vc;
  make_view_const(c->cam, c->Twc, &vc);
  CK(c->winners.reserve(c->P));
  CK(c->one_render.reserve(c->pitch));
  const float4* d_centres = reinterpret_cast<const float4*>(c->params.p + c->off_centres);
  const float margin[3] = {0, 0, 0};
  const bool timed = c->timed;
  c->timed = false;
  const int rc = render_views(c, vc, d_centres + (s - c->v_begin), 1, margin, false,
                              c->one_render.p, c->winners.p);
  c->timed = timed;
  if (rc) return rc;
  CK(cudaMemcpyAsync(host, c->winners.p, c->P * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return NMI_OK;
}

int nmi_get_warp(nmi_ctx* c, int w, uint8_t* host) {
  REQUIRE(c && host, NMI_ERR_INVALID, "null argument");
  REQUIRE(c->warps.p && w - c->w_begin >= 0 && w - c->w_begin < c->nwl, NMI_ERR_STATE,
          "warp not available on this rank");
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(host, c->warps.p + (size_t)(w - c->w_begin) * c->pitch, c->P,
                     cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return NMI_OK;
}

int nmi_get_hist_path(nmi_ctx* c, int s, int w, const nmi_flags* f, int path, uint32_t* J, uint32_t* HA,
                      uint32_t* HB, float* score) {
  REQUIRE(c && J && HA && HB, NMI_ERR_INVALID, "null argument");
  REQUIRE(valid_flags(f), NMI_ERR_INVALID, "invalid flags");
  REQUIRE(path >= 0 && path <= 2, NMI_ERR_INVALID, "path must be 0 (automatic), 1 (plain batched build) or 2 (side tables)");
  REQUIRE(c->has_search, NMI_ERR_STATE, "no search has run");
  REQUIRE(s - c->v_begin >= 0 && s - c->v_begin < c->nvl && w - c->w_begin >= 0 &&
              w - c->w_begin < c->nwl,
          NMI_ERR_INVALID, "pair not on this rank");
  CK(cudaSetDevice(c->device));
  CK(c->dumpJ.reserve(65536));
  CK(c->dumpH.reserve(512));
  const int rc = eval_images(c, c->renders.p + (size_t)(s - c->v_begin) * c->pitch,
                             c->warps.p + (size_t)(w - c->w_begin) * c->pitch, (uint32_t)c->P, f,
                             c->dumpJ.p, c->dumpH.p, c->dumpH.p + 256, score, path);
  if (rc) return rc;
  const size_t nb = (size_t)f->bins;
  CK(cudaMemcpy(J, c->dumpJ.p, nb * nb * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(HA, c->dumpH.p, nb * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(HB, c->dumpH.p + 256, nb * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return NMI_OK;
}

int nmi_get_hist(nmi_ctx* c, int s, int w, const nmi_flags* f, uint32_t* J, uint32_t* HA,
                 uint32_t* HB, float* score) {
  return nmi_get_hist_path(c, s, w, f, 0, J, HA, HB, score);
}

int nmi_last_hist_path(nmi_ctx* c) { return c && c->has_search ? (c->last_skipcap ? 2 : 1) : 0; }

// Score every pair of two caller-supplied image stacks with the batched launch of a grid search.
int nmi_score_pairs(nmi_ctx* c, const void* renders_dev, int n_r, size_t r_stride, const void* warps_dev,
                    int n_w, size_t w_stride, int W, int H, const nmi_flags* f, float* scores_host) {
  REQUIRE(c && renders_dev && warps_dev, NMI_ERR_INVALID, "null argument");
  REQUIRE(valid_flags(f), NMI_ERR_INVALID, "invalid flags");
  REQUIRE(c->has_cam && W == c->cam.W && H == c->cam.H, NMI_ERR_INVALID, "size != camera size");
  REQUIRE(n_r >= 1 && n_w >= 1 && (size_t)n_r * n_w <= (1u << 26), NMI_ERR_INVALID, "bad stack sizes");
  REQUIRE(r_stride >= c->P && w_stride >= c->P, NMI_ERR_INVALID, "stride smaller than an image");
  REQUIRE(c->P / 4096 + 2 < 2048, NMI_ERR_INVALID, "image too large for the histogram kernel");
  CK(cudaSetDevice(c->device));
  const size_t npl = (size_t)n_r * n_w;
  const size_t off_p = 0;
  const size_t off_i = align_up(off_p + sizeof(int2) * npl, 256);
  const size_t bytes = align_up(off_i + sizeof(uint32_t) * npl, 256);
  if (int rc = params_acquire(c)) return rc;
  if (int rc = ensure_pinned(&c->h_params, &c->h_params_cap, bytes)) return rc;
  CK(cudaStreamSynchronize(c->stream));
  CK(c->params.reserve(bytes));
  CK(c->renders.reserve((size_t)n_r * c->pitch));
  CK(c->warps.reserve((size_t)n_w * c->pitch));
  CK(c->scores.reserve(npl));
  CK(c->key.reserve(1));
  fill_pair_schedule(reinterpret_cast<int2*>(c->h_params + off_p), reinterpret_cast<uint32_t*>(c->h_params + off_i),
                     n_r, n_w, n_r, 0, 0);
  c->launches = 0;
  if (c->timed) {
    for (int i = 0; i < 4; i++) CK(cudaEventRecord(c->ev[i], c->stream));
  }
  CK(cudaMemcpyAsync(c->params.p, c->h_params, bytes, cudaMemcpyHostToDevice, c->stream));
  if (int rc = params_uploaded(c)) return rc;
  // into our padded, 128 B aligned slots (the TMA loads of a last partial chunk stay inside memory we own)
  CK(cudaMemcpy2DAsync(c->renders.p, c->pitch, renders_dev, r_stride, c->P, (size_t)n_r, cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaMemcpy2DAsync(c->warps.p, c->pitch, warps_dev, w_stride, c->P, (size_t)n_w, cudaMemcpyDeviceToDevice, c->stream));
  if (int rc = score_pairs_launch(c, f, n_r, n_w, reinterpret_cast<const int2*>(c->params.p + off_p),
                                  reinterpret_cast<const uint32_t*>(c->params.p + off_i), npl, npl, nullptr, nullptr,
                                  false))
    return rc;
  if (scores_host)
    CK(cudaMemcpyAsync(scores_host, c->scores.p, npl * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->has_search = true;  // nmi_get_render / nmi_get_warp / nmi_get_hist* now address these stacks
  c->grid = nmi_grid{};
  c->grid.nS[0] = n_r; c->grid.nS[1] = c->grid.nS[2] = 1;
  c->grid.nW[0] = n_w; c->grid.nW[1] = c->grid.nW[2] = 1;
  c->nvl = n_r;
  c->nwl = n_w;
  c->v_begin = c->w_begin = 0;
  return NMI_OK;
}

int nmi_get_timings(nmi_ctx* c, float ms[8], int* launches) {
  REQUIRE(c && ms, NMI_ERR_INVALID, "null argument");
  REQUIRE(c->has_search && c->timed, NMI_ERR_STATE, "no timed search has run");
  CK(cudaSetDevice(c->device));
  CK(cudaEventSynchronize(c->ev[6]));
  for (int i = 0; i < 6; i++) CK(cudaEventElapsedTime(&ms[i], c->ev[i], c->ev[i + 1]));
  CK(cudaEventElapsedTime(&ms[6], c->ev[0], c->ev[6]));
  ms[7] = 0;
  if (launches) *launches = c->launches;
  return NMI_OK;
}

}  // extern "C"
