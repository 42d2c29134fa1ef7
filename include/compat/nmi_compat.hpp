// nmi_compat.hpp -- process-wide state behind the reference-named drop-ins.
//
// The reference keeps its device state in process globals too: one GL context
// (rendering.hpp:268), one FBO texture (rendering.hpp:341) and static scratch pointers
// (NMI.cu:165-167).  Here that role is played by one nmi_ctx, created on first use.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "../nmi_b200.h"
#include "nmi_compat_types.hpp"

namespace nmi_compat {

// The context every compat object uses (device = $NMI_DEVICE or 0).  Aborts like
// checkCudaErrors (kernel.cu:53) when no B200 is usable: there is no CPU fallback.
nmi_ctx* context();
void set_context(nmi_ctx* ctx);  // adopt an externally created context (e.g. per-rank)
void shutdown();                 // destroy the process-wide context

// run-time versions of the reference's compile-time knobs (kernel.cuh:22-23, NMI.cuh:39,
// allProperties.hpp:39); defaults: 256 bins, SUC, BG = true.
nmi_flags& flags();

// camera shared by Rendering<> and Image (both read the same YAML in NmiObjects)
nmi_camera& camera();
void apply_camera();  // push camera() into the context

[[noreturn]] void die(const char* where);  // print nmi_last_error() and exit(EXIT_FAILURE)
inline void check(int code, const char* where) {
  if (code != NMI_OK) die(where);
}

// Minimal reader for the OpenCV-FileStorage YAML the reference uses (localization.cpp:131):
// "key: value" scalars, quoted strings and !!opencv-matrix blocks.
class Yaml {
 public:
  explicit Yaml(const std::string& path);
  bool ok() const { return ok_; }
  bool has(const std::string& key) const { return scalars_.count(key) || mats_.count(key); }
  double num(const std::string& key, double dflt = 0.0) const;
  std::string str(const std::string& key, const std::string& dflt = "") const;
  // rows*cols values, row-major; empty when absent
  std::vector<double> mat(const std::string& key, int* rows = nullptr, int* cols = nullptr) const;

 private:
  struct M { int rows = 0, cols = 0; std::vector<double> data; };
  bool ok_ = false;
  std::map<std::string, std::string> scalars_;
  std::map<std::string, M> mats_;
};

// loadXYZ (objloader.cpp:225-264): ASCII "x y z r g b" minus the offset file's "ox oy oz";
// colour scaled by 1/256; only R is kept (the FBO is GL_RED, rendering.hpp:347).
// Like the reference's `while(!in.eof())` loop, a trailing newline duplicates the last point.
bool loadXYZ(const char* path, const char* offset_path, std::vector<float>& xyzi);

// setupCam (ioData.cpp:177-197) reduces the pose to (pos, dir = pos + z, up) and the GL renderer rebuilds
// its view from those three vectors (rendering.hpp:547-553) -- `dir - pos` is then z only up to the
// rounding of a sum at the magnitude of the camera position.  The drop-in keeps the reference's
// signatures, but has no reason to lose those bits: setupCam remembers the Twc it was given and
// Rendering::setCamera, handed the very triple setupCam produced, gets the exact matrix back, so the
// zero-change call sequence (Tracking.cc:1873-1894) scores bit for bit what the batched search scores.
void remember_setup_cam(const float pos[3], const float dir[3], const float up[3], const float Twc[16]);
bool recall_setup_cam(const float pos[3], const float dir[3], const float up[3], float Twc[16]);

// loadOBJ (objloader.cpp:140-223): "v x y z", "vt u v", "f a/b c/d e/f" (1-based); any other
// face syntax is rejected like the reference ("File can't be read by this simple parser").
// Output is un-indexed like the reference's VBOs: 3 positions + 3 uvs per triangle.
bool loadOBJ(const char* path, std::vector<float>& xyz /*3 per vertex*/, std::vector<float>& uv /*2 per vertex*/);
// loadBMP_custom (texture.cpp:31-107): 24-bpp uncompressed BMP; bytes are returned exactly as
// the reference hands them to glTexImage2D(GL_RGB): file order B,G,R is READ AS R,G,B, rows
// bottom-up.
bool loadBMP24(const char* path, int& width, int& height, std::vector<unsigned char>& rgb);
// Mesh model for nmi_set_mesh from the reference's OBJ + BMP: per vertex the fragment
// shader's luma 0.299 r + 0.587 g + 0.114 b (ShadingWithTexture.fragmentshader:16) of the
// NEAREST level-0 texel under GL_REPEAT (texture.cpp:100-101), on the byte-swapped channels
// the reference uploads.  The reference samples a trilinear mip-mapped texture per fragment;
// flat per-triangle shading from the first vertex is this build's documented deviation.
bool meshFromObjBmp(const char* obj_path, const char* bmp_path, std::vector<float>& verts_xyzg,
                    std::vector<uint32_t>& tris);
// The same plus what per-fragment shading needs (nmi_set_mesh_textured): the UV of every face corner
// (un-indexed, objloader.cpp:206-220) and the BMP payload exactly as loadBMP_custom uploads it.
bool meshFromObjBmp(const char* obj_path, const char* bmp_path, std::vector<float>& verts_xyzg,
                    std::vector<uint32_t>& tris, std::vector<float>* corner_uv, std::vector<unsigned char>* texture,
                    int* tex_w, int* tex_h);

}  // namespace nmi_compat
