// nmiSearchKernel.hpp -- drop-in for Thirdparty/Localization/nmiSearchKernel.hpp:25-86.
// Same public data members and methods; the refinement rules are delegated to the
// library's host code (nmi_grid_resize / nmi_grid_is_middle, csrc/host_math.cpp), which
// restates nmiSearchKernel.cpp:99-141.
#pragma once
#include <iomanip>
#include <iostream>

#include "../nmi_b200.h"

class NmiSearchKernel {
 public:
  int numSynthX, numSynthY, numSynthZ, numWarpX, numWarpY, numWarpZ;
  float stepX, stepY, stepZ, stepRadX, stepRadY, stepRadZ;
  float NMI;
  int bestSynthX, bestSynthY, bestSynthZ, bestWarpX, bestWarpY, bestWarpZ;

  NmiSearchKernel(int nsx, int nsy, int nsz, int nwx, int nwy, int nwz, float sx, float sy, float sz,
                  float rx, float ry, float rz) {
    setKernel(nsx, nsy, nsz, nwx, nwy, nwz, sx, sy, sz, rx, ry, rz);
    resetBest();
  }
  NmiSearchKernel() { reset(); }

  void setKernel(int nsx, int nsy, int nsz, int nwx, int nwy, int nwz, float sx, float sy, float sz,
                 float rx, float ry, float rz) {
    numSynthX = nsx; numSynthY = nsy; numSynthZ = nsz;
    numWarpX = nwx; numWarpY = nwy; numWarpZ = nwz;
    stepX = sx; stepY = sy; stepZ = sz;
    stepRadX = rx; stepRadY = ry; stepRadZ = rz;
  }
  void setKernel(NmiSearchKernel* k) {
    setKernel(k->numSynthX, k->numSynthY, k->numSynthZ, k->numWarpX, k->numWarpY, k->numWarpZ,
              k->stepX, k->stepY, k->stepZ, k->stepRadX, k->stepRadY, k->stepRadZ);
  }
  void setBest(int sx, int sy, int sz, int wx, int wy, int wz, float nmi) {
    bestSynthX = sx; bestSynthY = sy; bestSynthZ = sz;
    bestWarpX = wx; bestWarpY = wy; bestWarpZ = wz;
    NMI = nmi;
  }
  void setBest(NmiSearchKernel* k) {  // copies the indices only (nmiSearchKernel.cpp:82-90)
    bestSynthX = k->bestSynthX; bestSynthY = k->bestSynthY; bestSynthZ = k->bestSynthZ;
    bestWarpX = k->bestWarpX; bestWarpY = k->bestWarpY; bestWarpZ = k->bestWarpZ;
  }
  void setTo(NmiSearchKernel* k) {
    setKernel(k);
    setBest(k);
    NMI = k->NMI;
  }

  bool isMiddle() {  // nmiSearchKernel.cpp:99
    nmi_grid g = grid();
    int32_t s[3], w[3];
    best(s, w);
    return nmi_grid_is_middle(&g, s, w) != 0;
  }
  void resizeKernel() {  // nmiSearchKernel.cpp:104-141
    nmi_grid g = grid();
    int32_t s[3], w[3];
    best(s, w);
    nmi_grid_resize(&g, s, w);
    setGrid(g);
  }

  void resetKernel() { setKernel(-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1); }
  void resetBest() { setBest(-1, -1, -1, -1, -1, -1, 0); }
  void reset() {
    resetKernel();
    resetBest();
    NMI = 0;
  }

  // grid <-> C ABI
  nmi_grid grid() const {
    nmi_grid g;
    g.nS[0] = numSynthX; g.nS[1] = numSynthY; g.nS[2] = numSynthZ;
    g.nW[0] = numWarpX; g.nW[1] = numWarpY; g.nW[2] = numWarpZ;
    g.stepT[0] = stepX; g.stepT[1] = stepY; g.stepT[2] = stepZ;
    g.stepR[0] = stepRadX; g.stepR[1] = stepRadY; g.stepR[2] = stepRadZ;
    return g;
  }
  void setGrid(const nmi_grid& g) {
    setKernel(g.nS[0], g.nS[1], g.nS[2], g.nW[0], g.nW[1], g.nW[2], g.stepT[0], g.stepT[1],
              g.stepT[2], g.stepR[0], g.stepR[1], g.stepR[2]);
  }
  void best(int32_t s[3], int32_t w[3]) const {
    s[0] = bestSynthX; s[1] = bestSynthY; s[2] = bestSynthZ;
    w[0] = bestWarpX; w[1] = bestWarpY; w[2] = bestWarpZ;
  }

  int getNumSynthX() { return numSynthX; }
  int getNumSynthY() { return numSynthY; }
  int getNumSynthZ() { return numSynthZ; }
  int getNumWarpX() { return numWarpX; }
  int getNumWarpY() { return numWarpY; }
  int getNumWarpZ() { return numWarpZ; }
  float getStepX() { return stepX; }
  float getStepY() { return stepY; }
  float getStepZ() { return stepZ; }
  float getStepRadX() { return stepRadX; }
  float getStepRadY() { return stepRadY; }
  float getStepRadZ() { return stepRadZ; }
  int getBestSynthX() { return bestSynthX; }
  int getBestSynthY() { return bestSynthY; }
  int getBestSynthZ() { return bestSynthZ; }
  int getBestWarpX() { return bestWarpX; }
  int getBestWarpY() { return bestWarpY; }
  int getBestWarpZ() { return bestWarpZ; }
  float getNmi() { return NMI; }

  // same line format as the reference's _log.txt entries (nmiSearchKernel.cpp:183-195)
  friend std::ostream& operator<<(std::ostream& os, const NmiSearchKernel& k) {
    os.precision(5);
    os << std::fixed;
    const char* tag[6] = {"sX: ", ";\t sY: ", ";\t sZ: ", ";\t rX: ", ";\t rY: ", ";\t rZ: "};
    const int b[6] = {k.bestSynthX, k.bestSynthY, k.bestSynthZ, k.bestWarpX, k.bestWarpY, k.bestWarpZ};
    const int n[6] = {k.numSynthX, k.numSynthY, k.numSynthZ, k.numWarpX, k.numWarpY, k.numWarpZ};
    const float st[6] = {k.stepX, k.stepY, k.stepZ, k.stepRadX, k.stepRadY, k.stepRadZ};
    for (int i = 0; i < 6; i++)
      os << tag[i] << std::setw(2) << b[i] << "/" << std::setw(1) << n[i] << ": " << std::setw(6) << st[i];
    os << ";\t NMI: " << k.NMI;
    return os;
  }
};
