// test_nccl_host.cpp -- the multi-GPU coarse-to-fine search driven from C++ with NCCL itself
// (VERDICT r1 missing #5: "Host code stays C++"; INTEGRATION.md section 5 shows this callback).
//
// One process, one thread per GPU: ncclCommInitAll, one nmi_ctx per GPU, the same model / camera /
// frame on every rank, and nmi_relocalize_sharded with
//     ncclAllReduce(key, key, 1, ncclUint64, ncclMax, comm, stream)
// as the exchange callback (the only thing that crosses NVLink: 8 bytes per search level).
// Asserts: every rank returns the SAME result, equal to nmi_relocalize on one GPU
// (levels, every level's winner and score bits, final pose, accept / reject flags).
//
//   test_nccl_host [n_gpus | 0 = all visible]      -> prints "NCCL HOST OK <n>"
#include <cuda_runtime.h>
#include <nccl.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "nmi_b200.h"

#define REQ(cond)                                                                     \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      std::fprintf(stderr, "FAILED %s:%d: %s (%s)\n", __FILE__, __LINE__, #cond, nmi_last_error()); \
      std::exit(1);                                                                   \
    }                                                                                 \
  } while (0)

namespace {

// deterministic terrain-like cloud under a camera looking down (the shape of synth.make_cloud)
struct Lcg {
  uint64_t s;
  double next() {  // U[0,1)
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return (double)(s >> 11) * (1.0 / 9007199254740992.0);
  }
};

std::vector<float> make_cloud(size_t n) {
  std::vector<float> p(4 * n);
  Lcg r{1234};
  for (size_t i = 0; i < n; i++) {
    const double x = -24.0 + 48.0 * r.next(), y = -24.0 + 48.0 * r.next();
    const double z = 2.0 * std::sin(0.13 * x) * std::cos(0.11 * y) + 0.2 * (r.next() - 0.5);
    double v = 127.0 + 60.0 * std::sin(0.5 * x) + 50.0 * std::cos(0.37 * y) + 15.0 * (r.next() - 0.5);
    v = std::fmin(254.0, std::fmax(0.0, std::rint(v)));
    p[4 * i] = (float)x; p[4 * i + 1] = (float)y; p[4 * i + 2] = (float)z; p[4 * i + 3] = (float)(v / 256.0);
  }
  return p;
}

std::vector<uint8_t> make_frame(int W, int H) {
  std::vector<uint8_t> f((size_t)W * H);
  Lcg r{5};
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      const double v = 128 + 50 * std::sin(x / 37.0) * std::cos(y / 29.0) + 40 * std::sin((x + y) / 91.0) + 12 * (r.next() - 0.5);
      f[(size_t)y * W + x] = (uint8_t)std::fmin(255.0, std::fmax(0.0, std::rint(v)));
    }
  return f;
}

int exchange(void* user, void* key_dev, void* stream) {
  return ncclAllReduce(key_dev, key_dev, 1, ncclUint64, ncclMax, *static_cast<ncclComm_t*>(user),
                       static_cast<cudaStream_t>(stream)) == ncclSuccess ? 0 : 1;
}

bool same_result(const nmi_reloc_result& a, const nmi_reloc_result& b) {
  if (a.iterations != b.iterations || a.relocalized != b.relocalized || a.failed != b.failed ||
      a.n_levels != b.n_levels || a.n_evals != b.n_evals)
    return false;
  if (std::memcmp(a.Twc, b.Twc, sizeof a.Twc) || std::memcmp(&a.nmi, &b.nmi, sizeof a.nmi) ||
      std::memcmp(&a.last_nmi, &b.last_nmi, sizeof a.nmi))
    return false;
  for (int i = 0; i < a.n_levels; i++)
    if (std::memcmp(a.levels[i].best_s, b.levels[i].best_s, sizeof a.levels[i].best_s) ||
        std::memcmp(a.levels[i].best_w, b.levels[i].best_w, sizeof a.levels[i].best_w) ||
        std::memcmp(&a.levels[i].nmi, &b.levels[i].nmi, sizeof(float)) ||
        std::memcmp(&a.levels[i].grid, &b.levels[i].grid, sizeof(nmi_grid)))
      return false;
  return true;
}

}  // namespace

int main(int argc, char** argv) {
  int ndev = 0;
  REQ(cudaGetDeviceCount(&ndev) == cudaSuccess && ndev >= 1);
  int world = argc > 1 ? std::atoi(argv[1]) : 0;
  if (world <= 0 || world > ndev) world = ndev;

  const int W = 320, H = 200;
  const nmi_camera cam{W, H, 190.0, 190.0, 158.0, 101.0, 5.0, 30.0, 3.0f};
  const std::vector<float> cloud = make_cloud(250000);
  const std::vector<uint8_t> frame = make_frame(W, H);
  // prior: 15 m above the terrain, looking down, tilted 10 degrees about x (synth.prior_pose)
  const double a = 10.0 * M_PI / 180.0;
  const float Twc[16] = {1, 0, 0, 0, 0, (float)-std::cos(a), (float)std::sin(a), 0,
                         0, (float)-std::sin(a), (float)-std::cos(a), 15.0f, 0, 0, 0, 1};
  nmi_grid grid{};
  const int nS[3] = {3, 3, 2}, nW[3] = {3, 2, 3};
  const float sT[3] = {0.2f, 0.2f, 0.5f}, sR[3] = {0.02f, 0.02f, 0.05f};
  for (int k = 0; k < 3; k++) { grid.nS[k] = nS[k]; grid.nW[k] = nW[k]; grid.stepT[k] = sT[k]; grid.stepR[k] = sR[k]; }
  const nmi_flags flags{256, NMI_SCORE_SUC, 1, 0};
  nmi_reloc_params prm{};
  prm.threshold = 0.0f;
  prm.max_iterations = 4;

  std::vector<int> devs(world);
  for (int i = 0; i < world; i++) devs[i] = i;
  std::vector<ncclComm_t> comms(world);
  REQ(ncclCommInitAll(comms.data(), world, devs.data()) == ncclSuccess);

  std::vector<nmi_reloc_result> res(world);
  std::vector<int> rcs(world, -1);
  std::vector<std::thread> th;
  for (int r = 0; r < world; r++)
    th.emplace_back([&, r] {
      nmi_ctx* ctx = nullptr;
      if (nmi_ctx_create(r, &ctx) != NMI_OK) return;
      if (nmi_set_camera(ctx, &cam) == NMI_OK && nmi_set_points(ctx, cloud.data(), cloud.size() / 4) == NMI_OK &&
          nmi_set_frame(ctx, frame.data(), W, H) == NMI_OK)
        rcs[r] = nmi_relocalize_sharded(ctx, Twc, &grid, &flags, &prm, r, world, /*key_dev=*/nullptr, exchange,
                                        &comms[r], &res[r]);
      nmi_ctx_sync(ctx);
      nmi_ctx_destroy(ctx);
    });
  for (auto& t : th) t.join();
  for (int r = 0; r < world; r++) REQ(rcs[r] == NMI_OK);
  for (int r = 1; r < world; r++) REQ(same_result(res[0], res[r]));

  // the single-GPU driver must agree with the sharded one
  nmi_ctx* ctx = nullptr;
  REQ(nmi_ctx_create(0, &ctx) == NMI_OK);
  REQ(nmi_set_camera(ctx, &cam) == NMI_OK);
  REQ(nmi_set_points(ctx, cloud.data(), cloud.size() / 4) == NMI_OK);
  REQ(nmi_set_frame(ctx, frame.data(), W, H) == NMI_OK);
  nmi_reloc_result one{};
  REQ(nmi_relocalize(ctx, Twc, &grid, &flags, &prm, &one) == NMI_OK);
  nmi_ctx_destroy(ctx);
  REQ(same_result(one, res[0]));
  REQ(one.iterations >= 2 && one.n_evals >= 2 * 18 * 18);

  for (auto& c : comms) ncclCommDestroy(c);
  std::printf("levels %d evals %d nmi %.6f winner s=(%d,%d,%d) w=(%d,%d,%d)\n", one.iterations, one.n_evals, one.nmi,
              one.best_s[0], one.best_s[1], one.best_s[2], one.best_w[0], one.best_w[1], one.best_w[2]);
  std::printf("NCCL HOST OK %d\n", world);
  return 0;
}
