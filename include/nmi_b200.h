/*
 * nmi_b200.h -- C ABI of the B200-native NMI pose search.
 *
 * This is the drop-in boundary for orbslam2_NMI's localization hot path
 * (Thirdparty/Localization + Thirdparty/CUDA_Functions).  Plain pointers and
 * sizes only; no torch / OpenCV / GL types.  Every entry point names the
 * reference interface it replaces (paths relative to the reference repo).
 * The reference-named C++ drop-ins (CUDAF::NMIWithCuda_noMask, NmiObjects,
 * Image, Rendering<>, NmiSearchKernel, helperFunctions::find_max_elements,
 * setupCam) live in include/compat/ and are thin wrappers over this ABI.
 *
 * Conventions
 *   - images: row-major u8, top-down rows, stride == W;
 *   - Twc: row-major 4x4 fp32 camera->world, CV axes (x right, y down, z fwd);
 *   - grid arrays are ordered {x, y, z};
 *   - scores / rating order: linear index
 *       l = ((((wz*nWy + wy)*nWx + wx)*nSz + sz)*nSy + sy)*nSx + sx
 *     == the reference's rating[wz][wy][wx][sz][sy][sx] (localization.hpp:36);
 *   - all functions return NMI_OK (0) or an error code; nmi_last_error() gives
 *     the message (the reference prints and exit()s instead, kernel.cu:53).
 *   - there is NO CPU fallback: every compute entry point fails with
 *     NMI_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef NMI_B200_H
#define NMI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NMI_OK 0
#define NMI_ERR_INVALID 1   /* bad argument                                  */
#define NMI_ERR_CUDA 2      /* CUDA runtime error / no usable device         */
#define NMI_ERR_STATE 3     /* camera / model / frame not set yet            */
#define NMI_ERR_NO_WINNER 4 /* every score < 0: the reference's max-vector   */
                            /* is empty (helperFunctions.cpp:50-103)         */
#define NMI_ERR_RETRY 5     /* a rank's enqueued search was incomplete (its   */
                            /* splat-record bins filled up): redo the level  */
/* Key an enqueued search leaves instead of a winner when its renders are
 * incomplete.  Above every real key in signed and unsigned order (its score
 * word is a NaN pattern, which never wins find_max_elements), so the
 * max-allreduce hands it to every rank and all ranks redo the level together. */
#define NMI_KEY_RETRY 0x7FFFFFFFFFFFFFFFull

#define NMI_SCORE_ENMI 0 /* kernel.cuh:22  (H(A)+H(B))/H(A,B)                */
#define NMI_SCORE_SUC 1  /* kernel.cuh:23  2(1-H(A,B)/(H(A)+H(B))) (default) */

#define NMI_EMPTY 0xFFFFFFFFu /* z-buffer winner for "no primitive"          */

typedef struct nmi_ctx nmi_ctx;

/* YAML Camera.* and NMI.Render.* (localization.cpp:131-181) */
typedef struct {
  int W, H;
  double fx, fy, cx, cy;
  double zn, zf;    /* NMI.Render.NearPlane / FarPlane */
  float point_size; /* NMI.Render.PointSize            */
} nmi_camera;

/* NmiSearchKernel's grid part (nmiSearchKernel.hpp:29-30) */
typedef struct {
  int nS[3];      /* numSynth{X,Y,Z} */
  int nW[3];      /* numWarp{X,Y,Z}  */
  float stepT[3]; /* step{X,Y,Z}     metres  */
  float stepR[3]; /* stepRad{X,Y,Z}  radians */
} nmi_grid;

/* compile-time knobs of the reference made run-time */
typedef struct {
  int bins;       /* 256 (NMI.cuh:39) or 64                                   */
  int score_mode; /* NMI_SCORE_SUC / NMI_SCORE_ENMI (kernel.cuh:22-23)        */
  int bg;         /* nmi_prop_BG (allProperties.hpp:39); 0 skips value-0 px   */
  int variant;    /* histogram kernel variant 0..11, 0 = default */
} nmi_flags;

typedef struct {
  int32_t best_s[3]; /* bestSynth{X,Y,Z} */
  int32_t best_w[3]; /* bestWarp{X,Y,Z}  */
  int64_t best_index; /* linear index l, -1 when there is no winner           */
  float best_score;   /* NmiSearchKernel::NMI                                 */
  uint64_t key;       /* (bits(max(score,0)) << 32) | (0xFFFFFFFF - l)        */
  float gpu_ms;       /* device time of the search (CUDA events)              */
} nmi_result;

/* ---- context ----------------------------------------------------------- */
/* Replaces Rendering::initGL / CreateFrameBuffers (rendering.hpp:239,332) and
 * initHistogram256all / closeHistogram256all (NMI.cu:171-186): all scratch is
 * owned by the context and reused; nothing is allocated per evaluation.      */
int nmi_ctx_create(int device, nmi_ctx **out);
void nmi_ctx_destroy(nmi_ctx *ctx);
const char *nmi_last_error(void);
/* CUDA stream (cudaStream_t) every kernel of this context is enqueued on.    */
void *nmi_ctx_stream(nmi_ctx *ctx);
int nmi_ctx_sync(nmi_ctx *ctx);
/* Histogram kernel option, no reference counterpart (the reference's global-atomic
 * histogram, NMI.cu:52-108, degrades the same way on flat images): where all 16
 * pixels of a thread sit at the most frequent grey level of the render (or of the
 * warp) they are kept out of the shared-memory joint histogram and counted in small
 * side tables that the epilogue adds back.  0 never, 1 automatic (default; also
 * $NMI_HIST_SKIP): when those two levels cover >= 1/6 of the pixels, 2 always.
 * Results are bit-identical in every mode.                                     */
int nmi_ctx_set_hist_skip(nmi_ctx *ctx, int mode);

/* ---- model / camera / frame -------------------------------------------- */
/* Rendering ctor + initVBO projection (rendering.hpp:167-236, 196-202).      */
int nmi_set_camera(nmi_ctx *ctx, const nmi_camera *cam);
/* loadXYZ -> VBO upload (objloader.cpp:225-264, rendering.hpp:205-217).
 * xyzi: n x {x, y, z, I}, I = red/256 as the reference stores it.            */
int nmi_set_points(nmi_ctx *ctx, const float *xyzi_host, size_t n);
int nmi_set_points_device(nmi_ctx *ctx, const void *xyzi_dev, size_t n);
/* loadOBJ -> VBO upload of Rendering<1> (objloader.cpp:140-223,
 * rendering.hpp:219-229).  verts: nv x {x, y, z, grey 0..1}; tris: nt x 3 vertex
 * indices in draw order.  Shading is flat with the first vertex's grey (the
 * reference's mip-mapped texture lookup is a documented deviation).
 * Setting a mesh replaces a point cloud and vice versa.                        */
int nmi_set_mesh(nmi_ctx *ctx, const float *verts_host, size_t nv,
                 const uint32_t *tris_host, size_t nt);
/* Rendering<1> as the reference runs it (rendering.hpp:176-189: loadOBJ + loadBMP_custom, draw at
 * :588-620): the mesh with per-fragment texture shading.  corner_uv: 6 floats per triangle, (u, v)
 * of its corners 0, 1, 2 -- un-indexed, the layout loadOBJ's out_uvs has (objloader.cpp:206-220).
 * texture: tex_h rows of tex_w texels, 3 bytes each, in the FILE's byte order and row order
 * (loadBMP_custom hands the BMP payload to glTexImage2D(GL_RGB) untouched, texture.cpp:90: the
 * file's B,G,R land in the shader's r,g,b; row 0 = v 0).  Per fragment: perspective-correct UV,
 * GL_REPEAT, level-0 bilinear fetch of 0.299 r + 0.587 g + 0.114 b
 * (ShadingWithTexture.fragmentshader:16).  The reference's minification filter is trilinear over a
 * driver-built mip chain -- not reproducible, level 0 is used at every scale.    */
int nmi_set_mesh_textured(nmi_ctx *ctx, const float *verts, size_t nv, const uint32_t *tris,
                          size_t nt, const float *corner_uv, const uint8_t *texture, int tex_w,
                          int tex_h);
/* Image::loadOriginal (image.cpp:130-135): H2D upload of the grey frame.     */
int nmi_set_frame(nmi_ctx *ctx, const uint8_t *gray_host, int W, int H);
int nmi_set_frame_device(nmi_ctx *ctx, const void *gray_dev, int W, int H);

/* ---- the batched search (one call per grid) ----------------------------- */
/* Replaces the six-nested loop of Tracking::RelocalizeWithNMI
 * (src/Tracking.cc:1871-1905): Image::calculateWarping, nS x
 * renderToTextureOnGPU, nS*nW x CUDAF::NMIWithCuda_noMask and
 * helperFunctions::find_max_elements.  scores_host (optional, nP floats)
 * receives the rating array.  Synchronous.                                   */
int nmi_search(nmi_ctx *ctx, const float Twc[16], const nmi_grid *grid,
               const nmi_flags *flags, nmi_result *out, float *scores_host);

/* Multi-GPU form.  Scores only this rank's slice of the grid (see
 * nmi_partition) and writes the packed local winner key to key_dev (device
 * u64, e.g. the send buffer of an NCCL max-allreduce).  Asynchronous on
 * nmi_ctx_stream(); scores_dev (optional, nP floats, device) receives this
 * rank's scores, others untouched.                                           */
int nmi_search_enqueue(nmi_ctx *ctx, const float Twc[16], const nmi_grid *grid,
                       const nmi_flags *flags, int rank, int world,
                       void *key_dev, void *scores_dev);
/* Pose-grid partitioner: axis 0 = synthetic views, 1 = warps.    */
int nmi_partition(const nmi_grid *grid, int rank, int world, int *axis,
                  int *begin, int *end);
/* find_max_elements' answer from a (reduced) key.  NMI_ERR_RETRY for
 * NMI_KEY_RETRY, NMI_ERR_NO_WINNER when the low word is 0.                    */
int nmi_decode_key(const nmi_grid *grid, uint64_t key, nmi_result *out);
/* 8 bytes of device memory owned by the context, for hosts that have no allocator
 * of their own at hand (the exchange buffer of nmi_relocalize_sharded).          */
void *nmi_ctx_key_buffer(nmi_ctx *ctx);
/* Stream-synchronise, then copy the (reduced) 8-byte key at key_dev to *key.  */
int nmi_read_key(nmi_ctx *ctx, const void *key_dev, uint64_t *key);

/* ---- the coarse-to-fine driver ------------------------------------------ */
#define NMI_MAX_PREV_POSES 8
#define NMI_MAX_LEVELS 8
/* inputs of Tracking::RelocalizeWithNMIStrategy that live in Tracking's state  */
typedef struct {
  float threshold;              /* NMI.Treshold -> mfNmiInitTresholf (Tracking.cc:157)   */
  int max_iterations;           /* nmi_prop_MAX_ITERATION_COUNT (allProperties.hpp:27), 0 = 4 */
  float distance_since_last[3]; /* mDistanceSinceLastNMI (Tracking.cc:648-652)           */
  float rotation_since_last[3]; /* mRotationSinceLastNMI (Tracking.cc:654-660)           */
} nmi_reloc_params;

typedef struct {
  float Twc[16];     /* pose after accept / restore / reject                             */
  int relocalized;   /* Frame::SetNMIRelocalized                                          */
  int failed;        /* Frame::SetNMIFailed                                               */
  int iterations;    /* grid searches run (<= max_iterations)                             */
  float nmi;         /* NmiKernel->NMI     of the last search                             */
  float last_nmi;    /* LastNmiKernel->NMI                                                */
  float threshold_used;
  nmi_grid final_grid;       /* NmiKernel after the last resizeKernel                     */
  nmi_grid last_search_grid; /* grid of the last search that ran                          */
  int32_t best_s[3], best_w[3];
  int n_prev;                /* mvPreviousPoses (Tracking.cc:2094-2097)                    */
  float prev_Twc[NMI_MAX_PREV_POSES][16];
  int n_evals;               /* total (render, warp) pairs scored                         */
  float gpu_ms;
  /* per search that ran (first NMI_MAX_LEVELS): what the reference logs after each
   * RelocalizeWithNMI (Tracking.cc:2103-2106) -- NmiKernel with its winner, LastNmiKernel->NMI */
  int n_levels;
  struct {
    nmi_grid grid;
    int32_t best_s[3], best_w[3];
    float nmi, last_nmi;
  } levels[NMI_MAX_LEVELS];
} nmi_reloc_result;

/* Grid choice at the top of RelocalizeWithNMIStrategy (Tracking.cc:2001-2069):
 * 2 % of the motion since the last fix, axes under 5 mm / 1 mrad collapse to 1;
 * 5x5x5 translations while NOT_INITIALIZED; else the YAML "Initial" kernel.     */
void nmi_grid_from_motion(const nmi_grid *initial, const float dist[3],
                          const float rot[3], int not_initialized,
                          nmi_grid *out);
/* Tracking::RelocalizeWithNMIStrategy (src/Tracking.cc:1987-2179): up to
 * max_iterations batched searches, step halving, stop rules, accept/reject.     */
int nmi_relocalize(nmi_ctx *ctx, const float Twc[16], const nmi_grid *start_grid,
                   const nmi_flags *flags, const nmi_reloc_params *params,
                   nmi_reloc_result *out);

/* The same driver with the per-level grid search supplied by the caller: `search`
 * stands where Tracking::RelocalizeWithNMI (src/Tracking.cc:1851-1985) stands in the
 * reference -- score the grid around `Twc`, fill *out with the winner (best_s, best_w,
 * best_score; gpu_ms optional) and return NMI_OK, or an error code that aborts the
 * driver.  nmi_relocalize and nmi_relocalize_sharded are this driver over nmi_search
 * and over the sharded search below.                                              */
typedef int (*nmi_level_search_fn)(void *user, const float Twc[16],
                                   const nmi_grid *grid, nmi_result *out);
int nmi_relocalize_with(nmi_level_search_fn search, void *user, const float Twc[16],
                        const nmi_grid *start_grid, const nmi_reloc_params *params,
                        nmi_reloc_result *out);

/* Multi-GPU coarse-to-fine search (BASELINE config 4): every rank runs
 * this driver with the same arguments; each level is one nmi_search_enqueue of the
 * rank's slice followed by `exchange`, which must max-reduce the u64 at key_dev over
 * all ranks, in place, on `stream` (one ncclAllReduce(ncclUint64 / ncclInt64, ncclMax,
 * 8 bytes); torch.distributed.all_reduce in the harness) and return 0.  All ranks
 * then decode the same winner, resize the grid identically and go on; no other
 * data moves between GPUs.  A rank whose splat-record bins filled up publishes
 * NMI_KEY_RETRY and every rank redoes that level once with exact sizing.
 * key_dev: 8 bytes of device memory on the context's GPU (the collective's buffer);
 * NULL = nmi_ctx_key_buffer(ctx).                                                   */
typedef int (*nmi_exchange_fn)(void *user, void *key_dev, void *stream);
int nmi_relocalize_sharded(nmi_ctx *ctx, const float Twc[16], const nmi_grid *start_grid,
                           const nmi_flags *flags, const nmi_reloc_params *params,
                           int rank, int world, void *key_dev,
                           nmi_exchange_fn exchange, void *user, nmi_reloc_result *out);
/* Host-side cost of level `level` (searches in launch order, retries included) of the calling
 * thread's last nmi_relocalize_sharded, in microseconds: [0] enqueueing the level's kernels,
 * [1] the exchange callback, [2] waiting for the 8-byte key, [3] this rank's device time.
 * NMI_ERR_INVALID past the last level.  (No reference counterpart: Tracking.cc:2088-2130 runs
 * its levels synchronously on one GPU.)                                               */
int nmi_last_level_trace(int level, float out_us[4]);

/* ---- stage-level entry points (the reference's own call granularity) ---- */
/* Rendering::renderToTextureOnGPU(calculateTranslation(sx,sy,sz))
 * (rendering.hpp:530-630, 644-665).  Returns an opaque render handle in
 * *handle -- the GL-free stand-in for getrenderedTexture() (rendering.hpp:749).*/
int nmi_render_cell(nmi_ctx *ctx, const float Twc[16], const nmi_grid *grid,
                    int sx, int sy, int sz, unsigned int *handle);
/* Same with an explicit world-frame translation, the argument
 * renderToTextureOnGPU itself takes (rendering.hpp:530).                      */
int nmi_render_at(nmi_ctx *ctx, const float Twc[16], const float t[3],
                  unsigned int *handle);
/* Image::calculateWarping (image.cpp:115-128): all nW warps of the frame.    */
int nmi_warp_cells(nmi_ctx *ctx, const nmi_grid *grid);
/* Image::getImageGPU(z,y,x).data (image.cpp:142): device pointer, W*H u8.    */
int nmi_warp_ptr(nmi_ctx *ctx, const nmi_grid *grid, int wx, int wy, int wz,
                 void **dev_ptr);
/* CUDAF::NMIWithCuda_noMask (kernel.cuh:37, kernel.cu:49-114): one evaluation
 * of a device-resident warped image against a render handle.                 */
int nmi_eval_pair(nmi_ctx *ctx, const void *warped_dev, unsigned int handle,
                  int W, int H, const nmi_flags *flags, float *score_host);

/* Same evaluation with the integer histograms left on the device, the layout
 * histogram256all fills (NMI.cuh:63-71): J[render * bins + camera], HA = render,
 * HB = camera.  Any of J_dev / HA_dev / HB_dev / score_host may be NULL.       */
int nmi_eval_pair_dev(nmi_ctx *ctx, const void *warped_dev, unsigned int handle,
                      int W, int H, const nmi_flags *flags, uint32_t *J_dev,
                      uint32_t *HA_dev, uint32_t *HB_dev, float *score_host);
/* Adopt an externally produced render as the current render handle -- e.g. the
 * reference's GL texture mapped through cuda_gl_interop (kernel.cu:53-59) and
 * copied out of its cudaArray.  W*H u8 on the device, `pitch_bytes` between rows;
 * bottom_up != 0: first row is the bottom one (GL), flipped like NMI.cu:82 does. */
int nmi_import_render(nmi_ctx *ctx, const void *render_dev, size_t pitch_bytes,
                      int W, int H, int bottom_up, unsigned int *handle);

/* ---- host-side helpers of the search driver ----------------------------- */
/* Rendering::calculateTranslation (rendering.hpp:644-665)                    */
void nmi_cell_translation(const float Twc[16], const nmi_grid *grid, int sx,
                          int sy, int sz, float t[3]);
/* Image ctor warp matrices (image.cpp:76-108) -> inverse map used by warp.   */
void nmi_cell_homography_inv(const nmi_camera *cam, const nmi_grid *grid,
                             int wx, int wy, int wz, float minv[9]);
/* Tracking::CalculateNMIRelocalization (src/Tracking.cc:2374-2419)           */
void nmi_apply_winner(const float Twc[16], const nmi_grid *grid,
                      const int32_t s[3], const int32_t w[3],
                      float Twc_new[16]);
/* NmiSearchKernel::isMiddle / resizeKernel (nmiSearchKernel.cpp:99-141)      */
int nmi_grid_is_middle(const nmi_grid *grid, const int32_t s[3],
                       const int32_t w[3]);
void nmi_grid_resize(nmi_grid *grid, const int32_t s[3], const int32_t w[3]);

/* ---- parity / debug read-backs (results of the LAST search or stage) ---- */
int nmi_get_render(nmi_ctx *ctx, int s, uint8_t *host);    /* W*H u8         */
int nmi_get_winners(nmi_ctx *ctx, int s, uint32_t *host);  /* W*H u32        */
int nmi_get_warp(nmi_ctx *ctx, int w, uint8_t *host);      /* W*H u8         */
/* integer histograms of pair (s, w): J bins*bins, HA/HB bins (u32)           */
int nmi_get_hist(nmi_ctx *ctx, int s, int w, const nmi_flags *flags,
                 uint32_t *J, uint32_t *HA, uint32_t *HB, float *score);
/* The same read-back through a chosen build of the histogram kernel: path 0 = what
 * nmi_get_hist picks, 1 = the plain batched build a search launches when no hot grey level
 * was seen (variant 0: persistent CTAs; J is copied out before the epilogue, HA / HB / score
 * are the fast epilogue's own -- the code the benchmark times), 2 = the build with the
 * hot-bin side tables.  nmi_last_hist_path: which of 1 / 2 the last batched launch used.     */
int nmi_get_hist_path(nmi_ctx *ctx, int s, int w, const nmi_flags *flags, int path,
                      uint32_t *J, uint32_t *HA, uint32_t *HB, float *score);
int nmi_last_hist_path(nmi_ctx *ctx);
/* nS*nW x CUDAF::NMIWithCuda_noMask (src/Tracking.cc:1879-1894) over image stacks produced
 * elsewhere -- e.g. the reference's own GL renders and cv::cuda warps: every (render r,
 * warp w) pair of n_r renders and n_w warped frames (device memory, W*H u8 each, top-down
 * rows, `*_stride` bytes apart) is scored by the one batched launch a grid search uses.
 * scores_host[w * n_r + r] (may be NULL).  Afterwards nmi_get_render / nmi_get_warp /
 * nmi_get_hist* address these stacks.                                                        */
int nmi_score_pairs(nmi_ctx *ctx, const void *renders_dev, int n_r, size_t r_stride,
                    const void *warps_dev, int n_w, size_t w_stride, int W, int H,
                    const nmi_flags *flags, float *scores_host);
/* per-stage device times of the last search, ms:
 * [0] params+cull [1] render (bin + tile resolve, all view groups) [2] unused
 * [3] warp [4] hist+score [5] argmax
 * [6] total.  launches = kernels launched by the last search.                */
int nmi_get_timings(nmi_ctx *ctx, float ms[8], int *launches);

#ifdef __cplusplus
}
#endif
#endif
