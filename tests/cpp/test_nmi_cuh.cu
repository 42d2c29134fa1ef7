// test_nmi_cuh.cu -- the secondary exports of NMI.cuh (include/compat/NMI.cuh) driven with the
// call sequence of the reference's NMIWithCuda_noMask (kernel.cu:63-100), minus the GL
// interop: the render arrives in a plain cudaArray.  Built with nvcc by tests/test_compat_cpp.py.
//   test_nmi_cuh <W> <H> <render_bottom_up.raw> <warped.raw> <out_hist.bin>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "compat/NMI.cuh"

#define CHECK(x)                                                                  \
  do {                                                                            \
    cudaError_t e_ = (x);                                                         \
    if (e_ != cudaSuccess) {                                                      \
      std::printf("CUDA FAIL %s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return 1;                                                                   \
    }                                                                             \
  } while (0)

static bool read_all(const char* path, std::vector<unsigned char>& buf) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return false;
  const size_t n = std::fread(buf.data(), 1, buf.size(), f);
  std::fclose(f);
  return n == buf.size();
}

int main(int argc, char** argv) {
  if (argc < 6) return 2;
  const int W = std::atoi(argv[1]), H = std::atoi(argv[2]);
  std::vector<unsigned char> render((size_t)W * H), warped((size_t)W * H);
  if (!read_all(argv[3], render) || !read_all(argv[4], warped)) return 3;

  cudaArray* synthCUDA = nullptr;
  cudaChannelFormatDesc desc = cudaCreateChannelDesc<unsigned char>();
  CHECK(cudaMallocArray(&synthCUDA, &desc, W, H));
  CHECK(cudaMemcpy2DToArray(synthCUDA, 0, 0, render.data(), W, W, H, cudaMemcpyHostToDevice));
  uchar* d_Warped = nullptr;
  CHECK(cudaMalloc(&d_Warped, (size_t)W * H));
  CHECK(cudaMemcpy(d_Warped, warped.data(), (size_t)W * H, cudaMemcpyHostToDevice));

  uint *d_Histogram1, *d_Histogram2, *d_JointHistogram;
  float *d_Entropy1, *d_Entropy2, *d_JointEntropy, *d_JointEntropyShort;
  CHECK(cudaMalloc(&d_Histogram1, 256 * sizeof(uint)));
  CHECK(cudaMalloc(&d_Histogram2, 256 * sizeof(uint)));
  CHECK(cudaMalloc(&d_JointHistogram, 256 * 256 * sizeof(uint)));
  CHECK(cudaMalloc(&d_Entropy1, 256 * sizeof(float)));
  CHECK(cudaMalloc(&d_Entropy2, 256 * sizeof(float)));
  CHECK(cudaMalloc(&d_JointEntropyShort, 256 * sizeof(float)));
  CHECK(cudaMalloc(&d_JointEntropy, 256 * 256 * sizeof(float)));

  float score[2] = {0, 0};
  for (int rep = 0; rep < 2; rep++) {  // twice: the pairwise kernel's block counter must reset
    initHistogram256all();
    histogram256all(d_JointHistogram, d_Histogram1, d_Histogram2, d_Warped, (uint)W, (uint)H, synthCUDA);
    ComputeEntropyKernel<<<258, 256>>>(d_Histogram1, d_Histogram2, d_JointHistogram, W * H, d_Entropy1,
                                       d_Entropy2, d_JointEntropy);
    AddvectorParwiseMidKernel<<<256, 128>>>(d_JointEntropy, d_JointEntropyShort);
    AddVectorPairwiseKernel<<<3, 128>>>(d_Entropy1, d_Entropy2, d_JointEntropyShort);
    closeHistogram256all();
    CHECK(cudaMemcpy(&score[rep], d_Entropy1, sizeof(float), cudaMemcpyDeviceToHost));
  }
  std::vector<uint> out(65536 + 512);
  CHECK(cudaMemcpy(out.data(), d_JointHistogram, 65536 * sizeof(uint), cudaMemcpyDeviceToHost));
  CHECK(cudaMemcpy(out.data() + 65536, d_Histogram1, 256 * sizeof(uint), cudaMemcpyDeviceToHost));
  CHECK(cudaMemcpy(out.data() + 65536 + 256, d_Histogram2, 256 * sizeof(uint), cudaMemcpyDeviceToHost));
  FILE* f = std::fopen(argv[5], "wb");
  if (!f) return 4;
  std::fwrite(out.data(), sizeof(uint), out.size(), f);
  std::fclose(f);
  std::printf("SCORE %.9g %.9g\nNMICUH OK\n", score[0], score[1]);
  return 0;
}
