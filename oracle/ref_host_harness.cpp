// ref_host_harness.cpp -- TEST INFRASTRUCTURE ONLY.  C ABI over the reference's own host-side
// search bookkeeping, compiled unmodified from /root/reference/Thirdparty/Localization by
// oracle/Makefile.ref (g++, no GPU needed):
//   helperFunctions::find_max_elements   helperFunctions.cpp:50-103   
//   NmiSearchKernel::resizeKernel/isMiddle nmiSearchKernel.cpp:99-141 (row a15)
//   operator<<(ostream&, NmiSearchKernel) nmiSearchKernel.cpp:183-197 (the _log.txt line, row f4)
// The rating array is built exactly as NmiObjects does (localization.cpp:184-210): a jagged
// float****** indexed [wz][wy][wx][sz][sy][sx].
#include "helperFunctions.hpp"
#include "nmiSearchKernel.hpp"

#include <cstring>
#include <sstream>
#include <vector>

namespace {
struct Rating {
  int nS[3], nW[3];
  float****** r;
  Rating(const int* s, const int* w, const float* linear) {
    memcpy(nS, s, sizeof(nS)); memcpy(nW, w, sizeof(nW));
    size_t l = 0;  // linear order wz,wy,wx,sz,sy,sx with sx fastest (src/Tracking.cc:1879-1894 fills it so)
    r = new float*****[nW[2]];
    for (int wz = 0; wz < nW[2]; wz++) {
      r[wz] = new float****[nW[1]];
      for (int wy = 0; wy < nW[1]; wy++) {
        r[wz][wy] = new float***[nW[0]];
        for (int wx = 0; wx < nW[0]; wx++) {
          r[wz][wy][wx] = new float**[nS[2]];
          for (int sz = 0; sz < nS[2]; sz++) {
            r[wz][wy][wx][sz] = new float*[nS[1]];
            for (int sy = 0; sy < nS[1]; sy++) {
              r[wz][wy][wx][sz][sy] = new float[nS[0]];
              for (int sx = 0; sx < nS[0]; sx++) r[wz][wy][wx][sz][sy][sx] = linear[l++];
            }
          }
        }
      }
    }
  }
  ~Rating() {
    for (int wz = 0; wz < nW[2]; wz++) {
      for (int wy = 0; wy < nW[1]; wy++) {
        for (int wx = 0; wx < nW[0]; wx++) {
          for (int sz = 0; sz < nS[2]; sz++) {
            for (int sy = 0; sy < nS[1]; sy++) delete[] r[wz][wy][wx][sz][sy];
            delete[] r[wz][wy][wx][sz];
          }
          delete[] r[wz][wy][wx];
        }
        delete[] r[wz][wy];
      }
      delete[] r[wz];
    }
    delete[] r;
  }
};

NmiSearchKernel make_kernel(const int* nS, const int* nW, const float* stepT, const float* stepR) {
  return NmiSearchKernel(nS[0], nS[1], nS[2], nW[0], nW[1], nW[2], stepT[0], stepT[1], stepT[2],
                         stepR[0], stepR[1], stepR[2]);
}
}  // namespace

// -> number of maximal elements the reference collects (0: its vector is empty and [0] would be
// undefined behaviour); best_s/best_w/score = element [0], what Tracking.cc:1952 takes.
extern "C" int nmirefh_find_max(const float* rating_linear, const int* nS, const int* nW,
                                int* best_s, int* best_w, float* score) {
  Rating rt(nS, nW, rating_linear);
  const float one[3] = {1.f, 1.f, 1.f};
  NmiSearchKernel k = make_kernel(nS, nW, one, one);
  std::vector<NmiSearchKernel> m = helperFunctions::find_max_elements(rt.r, k);
  if (!m.empty()) {
    best_s[0] = m[0].getBestSynthX(); best_s[1] = m[0].getBestSynthY(); best_s[2] = m[0].getBestSynthZ();
    best_w[0] = m[0].getBestWarpX(); best_w[1] = m[0].getBestWarpY(); best_w[2] = m[0].getBestWarpZ();
    *score = m[0].getNmi();
  }
  return (int)m.size();
}

// resizeKernel on a kernel with the given best cell; nS/nW/stepT/stepR are updated in place.
// Returns isMiddle() evaluated BEFORE the resize (the order Tracking.cc:2090-2110 uses them in).
extern "C" int nmirefh_resize(int* nS, int* nW, float* stepT, float* stepR, const int* best_s, const int* best_w) {
  NmiSearchKernel k = make_kernel(nS, nW, stepT, stepR);
  k.setBest(best_s[0], best_s[1], best_s[2], best_w[0], best_w[1], best_w[2], 0.5f);
  const int middle = k.isMiddle() ? 1 : 0;
  k.resizeKernel();
  nS[0] = k.getNumSynthX(); nS[1] = k.getNumSynthY(); nS[2] = k.getNumSynthZ();
  nW[0] = k.getNumWarpX(); nW[1] = k.getNumWarpY(); nW[2] = k.getNumWarpZ();
  stepT[0] = k.getStepX(); stepT[1] = k.getStepY(); stepT[2] = k.getStepZ();
  stepR[0] = k.getStepRadX(); stepR[1] = k.getStepRadY(); stepR[2] = k.getStepRadZ();
  return middle;
}

// the log line operator<< prints for a kernel + best cell + score; returns its length.
extern "C" int nmirefh_format(const int* nS, const int* nW, const float* stepT, const float* stepR,
                              const int* best_s, const int* best_w, float nmi, char* out, int cap) {
  NmiSearchKernel k = make_kernel(nS, nW, stepT, stepR);
  k.setBest(best_s[0], best_s[1], best_s[2], best_w[0], best_w[1], best_w[2], nmi);
  std::ostringstream os;
  os << k;
  const std::string s = os.str();
  if (cap > 0) { strncpy(out, s.c_str(), (size_t)cap - 1); out[cap - 1] = 0; }
  return (int)s.size();
}
