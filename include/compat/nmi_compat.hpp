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

}  // namespace nmi_compat
