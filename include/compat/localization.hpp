// localization.hpp -- drop-in for Thirdparty/Localization/localization.hpp:31-72.
// NmiObjects owns the renderer, the image/warps, the 6-D rating array and the three
// NmiSearchKernels, constructed from the same YAML keys (localization.cpp:131-253).
// Additions for the batched path are marked "B200 extension".
#pragma once
#include <string>

#include "allProperties.hpp"
#include "cameraSettings.hpp"
#include "image.hpp"
#include "nmiSearchKernel.hpp"
#include "rendering.hpp"

class NmiObjects {
 public:
  Image* myImage;
  Rendering<nmi_prop_RENDER>* myRenderer;
  float****** rating;  // 6DoF, indexed [wz][wy][wx][sz][sy][sx] (localization.cpp:185-210)
  std::string resultsPath;
  std::string logPath;

  NmiSearchKernel* NmiKernel;
  NmiSearchKernel* LastNmiKernel;
  NmiSearchKernel* InitialNmiKernel;

  // localization.cpp:82.  The renderer's mode is the compile-time nmi_prop_RENDER of the code that
  // includes this header (allProperties.hpp:42: 1 = textured OBJ mesh, 4 = .xyz cloud); the library
  // carries both renderers and builds the one named here.
  explicit NmiObjects(const std::string& strSettingsFile) : NmiObjects(strSettingsFile, nmi_prop_RENDER) {}
  NmiObjects(const std::string& strSettingsFile, int render_mode);
  ~NmiObjects();

  void setNmiObjectsKernel(int numsynthx, int numsynthy, int numsynthz, int numwarpx, int numwarpy,
                           int numwarpz, float stepx, float stepy, float stepz, float stepradx,
                           float steprady, float stepradz);
  void setNmiObjectsKernel(NmiSearchKernel* NmiKernel);
  void NMIobjectsReInitialization();  // localization.cpp:410-420
  void setRendererVars(int numsyntx, int numsynty, int numsyntz, float stepx, float stepy, float stepz);
  void setImageVars(int numwarpx, int numwarpy, int numwarpz, float stepradx, float steprady, float stepradz);
  void incN() { N++; }
  int getN() { return N; }

  // ---- B200 extension: the whole grid of Tracking::RelocalizeWithNMI in one call --------
  // Uploads `gray` (Image::loadOriginal), runs nmi_search() for the current NmiKernel grid
  // around Twc (4x4 CV_32F), fills `rating`, and stores the winner in NmiKernel (setBest +
  // NMI), i.e. everything src/Tracking.cc:1871-1953 does.  Returns the new Twc
  // (CalculateNMIRelocalization, Tracking.cc:2374-2419).
  cv::Mat searchGrid(cv::Mat Twc, cv::Mat gray);
  // Tracking::RelocalizeWithNMIStrategy (Tracking.cc:1987-2179) on a pose.
  nmi_reloc_result relocalize(cv::Mat Twc, cv::Mat gray, const nmi_reloc_params& params,
                              bool not_initialized = false);
  // The same on `world` GPUs, one process and one NmiObjects per GPU: every level's
  // grid is sharded by nmi_partition and `exchange` max-reduces the 8-byte winner key at key_dev
  // over the ranks on the given stream (ncclAllReduce(key_dev, key_dev, 1, ncclUint64, ncclMax,
  // comm, stream)).  Every rank returns the same result.
  nmi_reloc_result relocalizeSharded(cv::Mat Twc, cv::Mat gray, const nmi_reloc_params& params, int rank,
                                     int world, void* key_dev, nmi_exchange_fn exchange, void* user,
                                     bool not_initialized = false);
  float threshold() const { return threshold_; }  // NMI.Treshold (Tracking.cc:157)

 private:
  int N;
  float threshold_ = 0.0f;
  int rw_[3] = {0, 0, 0}, rs_[3] = {0, 0, 0};  // dimensions `rating` was allocated with
  void allocRating(const nmi_grid& g);
  void deleteRating();
  void resizeKernel(int numsynthx, int numsynthy, int numsynthz, int numwarpx, int numwarpy, int numwarpz,
                    float stepx, float stepy, float stepz, float stepradx, float steprady, float stepradz);
};
