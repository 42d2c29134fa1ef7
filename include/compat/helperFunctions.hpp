// helperFunctions.hpp -- drop-in for Thirdparty/Localization/helperFunctions.hpp:26-30.
#pragma once
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "nmiSearchKernel.hpp"

namespace helperFunctions {

// helperFunctions.cpp:50-103: max starts at 0 (strict >), then every element == max is
// collected in loop order wz,wy,wx,sz,sy,sx; Tracking.cc:1952 uses element [0].
// The batched path gets the same answer from the device argmax (csrc/argmax.cu); this host
// version serves callers that fill `rating` themselves.
inline std::vector<NmiSearchKernel> find_max_elements(float****** nmi, NmiSearchKernel& k) {
  float best = 0;
  auto each = [&](auto&& fn) {
    for (int wz = 0; wz < k.getNumWarpZ(); wz++)
      for (int wy = 0; wy < k.getNumWarpY(); wy++)
        for (int wx = 0; wx < k.getNumWarpX(); wx++)
          for (int sz = 0; sz < k.getNumSynthZ(); sz++)
            for (int sy = 0; sy < k.getNumSynthY(); sy++)
              for (int sx = 0; sx < k.getNumSynthX(); sx++) fn(wz, wy, wx, sz, sy, sx);
  };
  each([&](int wz, int wy, int wx, int sz, int sy, int sx) {
    if (nmi[wz][wy][wx][sz][sy][sx] > best) best = nmi[wz][wy][wx][sz][sy][sx];
  });
  std::vector<NmiSearchKernel> out;
  each([&](int wz, int wy, int wx, int sz, int sy, int sx) {
    if (nmi[wz][wy][wx][sz][sy][sx] == best) {
      NmiSearchKernel e;
      e.setBest(sx, sy, sz, wx, wy, wz, nmi[wz][wy][wx][sz][sy][sx]);
      out.push_back(e);
    }
  });
  return out;
}

// helperFunctions.cpp:105-113: append to the log file, echo to stdout, clear the stream.
inline void log(std::stringstream& text, std::string Filename) {
  std::ofstream f(Filename, std::ios_base::app);
  std::cout << text.str();
  f << text.str();
  text.str("");
}

}  // namespace helperFunctions
