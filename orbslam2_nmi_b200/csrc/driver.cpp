// driver.cpp -- the coarse-to-fine NMI relocalisation driver (host C++, no CUDA).
//
// Restates Tracking::RelocalizeWithNMIStrategy (src/Tracking.cc:1987-2179) and the
// per-level body Tracking::RelocalizeWithNMI (src/Tracking.cc:1851-1985) on a plain
// pose instead of a Frame/KeyFrame object.  Every level is ONE nmi_search() call
// (the reference runs nS*nW separate evaluations), re-centred on the previous winner
// with NmiSearchKernel::resizeKernel's step-halving rule (nmiSearchKernel.cpp:104-141).
// The accept / reject logic is kept decision for decision.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "nmi_internal.h"

extern "C" {

void nmi_grid_from_motion(const nmi_grid* initial, const float dist[3], const float rot[3],
                          int not_initialized, nmi_grid* out) {
  // src/Tracking.cc:2001-2069
  const double min_trans = 0.005, min_rot = 0.001;  // allProperties.hpp:48-49
  if (dist[0] > 0.0f) {
    for (int k = 0; k < 3; ++k) {
      // "ORB-SLAM 2 has a drift about 1% so we conservatively search in the 2% proximity"
      const float st = static_cast<float>(dist[k] * 0.02);
      const float sr = static_cast<float>(rot[k] * 0.02);
      out->stepT[k] = st;
      out->stepR[k] = sr;
      out->nS[k] = st < min_trans ? 1 : initial->nS[k];
      out->nW[k] = sr < min_rot ? 1 : initial->nW[k];
    }
  } else if (not_initialized) {
    *out = *initial;
    out->nS[0] = out->nS[1] = out->nS[2] = 5;  // Tracking.cc:2057
  } else {
    *out = *initial;
  }
}

int nmi_relocalize_with(nmi_level_search_fn search, void* user, const float Twc_in[16],
                        const nmi_grid* start_grid, const nmi_reloc_params* prm, nmi_reloc_result* out) {
  if (!search || !Twc_in || !start_grid || !prm || !out) {
    nmi::set_error("nmi_relocalize: null argument");
    return NMI_ERR_INVALID;
  }
  std::memset(out, 0, sizeof *out);
  const int max_it = prm->max_iterations > 0 ? prm->max_iterations : 4;  // allProperties.hpp:27

  nmi_grid kernel = *start_grid;        // MyObjects->NmiKernel (grid part)
  float kernel_nmi = 0.0f;              //   ->NMI, reset() at Tracking.cc:1997
  float last_nmi = 0.0f;                // MyObjects->LastNmiKernel->NMI, reset() at :1998
  float pose[16], save[16], save_last[16];
  std::memcpy(pose, Twc_in, sizeof pose);
  std::memcpy(save, Twc_in, sizeof save);            // TcwSave      (:2082-2085)
  std::memcpy(save_last, Twc_in, sizeof save_last);  // TcwSaveLast  (:2087)
  int under = 0, i = 0;
  nmi_result best{};
  nmi_grid used = kernel;

  for (;;) {
    ++i;
    if (i > max_it) break;  // :2091
    if (out->n_prev < NMI_MAX_PREV_POSES)
      std::memcpy(out->prev_Twc[out->n_prev++], pose, sizeof pose);  // mvPreviousPoses (:2094-2097)

    // ---- RelocalizeWithNMI: one batched grid search around `pose` ----
    nmi_result r{};
    const int rc = search(user, pose, &kernel, &r);
    if (rc != NMI_OK) return rc;  // NMI_ERR_NO_WINNER is the reference's UB case: surface it
    out->gpu_ms += r.gpu_ms;
    out->n_evals += kernel.nS[0] * kernel.nS[1] * kernel.nS[2] * kernel.nW[0] * kernel.nW[1] * kernel.nW[2];
    kernel_nmi = r.best_score;  // NmiKernel->setBest + NMI (:1952-1953)
    best = r;
    used = kernel;
    float moved[16];
    nmi_apply_winner(pose, &kernel, r.best_s, r.best_w, moved);  // :1956, :1976-1983
    std::memcpy(pose, moved, sizeof pose);
    out->relocalized = 1;  // SetNMIRelocalized(true) (:1978,:1982)
    out->iterations = i;
    if (out->n_levels < NMI_MAX_LEVELS) {  // the state Tracking.cc:2103-2106 logs
      auto& lv = out->levels[out->n_levels++];
      lv.grid = kernel;
      for (int k = 0; k < 3; ++k) {
        lv.best_s[k] = r.best_s[k];
        lv.best_w[k] = r.best_w[k];
      }
      lv.nmi = kernel_nmi;
      lv.last_nmi = last_nmi;
    }

    if (i > 1 && nmi_grid_is_middle(&kernel, r.best_s, r.best_w)) break;  // :2108-2110
    if (i > 1) {                                                         // :2112-2121
      if (static_cast<double>(kernel_nmi / last_nmi) < 1.001) {
        if (under > 0) break;
        ++under;
      } else {
        under = 0;
      }
    }
    // NMIobjectsReInitialization (:2124): Last <- current, then halve the steps
    last_nmi = kernel_nmi;
    nmi_grid_resize(&kernel, r.best_s, r.best_w);
    std::memcpy(save_last, pose, sizeof pose);  // :2126-2129
  }

  if (kernel_nmi < last_nmi) std::memcpy(pose, save_last, sizeof pose);  // :2134-2139

  // :2143-2152 threshold relaxed with the distance travelled since the last NMI fix
  const double base = 5.0;
  const double d = std::sqrt(std::pow(prm->distance_since_last[0], 2) + std::pow(prm->distance_since_last[1], 2) +
                             std::pow(prm->distance_since_last[2], 2));
  double thr = prm->threshold;
  if (!(d < base)) {
    thr = prm->threshold * (base / d);
    if (thr < prm->threshold / 2) thr = prm->threshold / 2;
  }
  if (kernel_nmi < thr) {  // :2157-2168
    std::memcpy(pose, save, sizeof pose);
    out->relocalized = 0;
    out->failed = 1;
  }
  std::memcpy(out->Twc, pose, sizeof pose);
  out->nmi = kernel_nmi;
  out->last_nmi = last_nmi;
  out->threshold_used = static_cast<float>(thr);
  out->final_grid = kernel;
  out->last_search_grid = used;
  for (int k = 0; k < 3; ++k) {
    out->best_s[k] = best.best_s[k];
    out->best_w[k] = best.best_w[k];
  }
  return NMI_OK;
}

namespace {

// host-side cost of every level of this thread's last nmi_relocalize_sharded (nmi_last_level_trace)
struct LevelTrace {
  float enqueue_us, exchange_us, wait_us, device_ms;
};
thread_local LevelTrace g_trace[16];
thread_local int g_trace_n = 0;

struct SingleSearch {
  nmi_ctx* ctx;
  const nmi_flags* flags;
};
int single_level(void* user, const float Twc[16], const nmi_grid* grid, nmi_result* out) {
  const SingleSearch* s = static_cast<const SingleSearch*>(user);
  return nmi_search(s->ctx, Twc, grid, s->flags, out, nullptr);
}

// One level of the multi-GPU driver: this rank's slice, then the 8-byte max-allreduce of the
// packed key on the context's stream, then every rank decodes the same winner.
struct ShardedSearch {
  nmi_ctx* ctx;
  const nmi_flags* flags;
  int rank, world;
  void* key_dev;
  nmi_exchange_fn exchange;
  void* user;
};
int sharded_level(void* user, const float Twc[16], const nmi_grid* grid, nmi_result* out) {
  const ShardedSearch* s = static_cast<const ShardedSearch*>(user);
  static const bool trace = getenv("NMI_TRACE_LEVELS") != nullptr;  // host-side cost of a level, to stderr
  using clk = std::chrono::steady_clock;
  auto us = [](clk::time_point a, clk::time_point b) {
    return (long)std::chrono::duration_cast<std::chrono::microseconds>(b - a).count();
  };
  for (int attempt = 0; attempt < 2; ++attempt) {
    const auto t0 = clk::now();
    if (int rc = nmi_search_enqueue(s->ctx, Twc, grid, s->flags, s->rank, s->world, s->key_dev, nullptr)) return rc;
    const auto t1 = clk::now();
    if (s->exchange(s->user, s->key_dev, nmi_ctx_stream(s->ctx)) != 0) {
      nmi::set_error("nmi_relocalize_sharded: the exchange callback failed");
      return NMI_ERR_CUDA;
    }
    const auto t2 = clk::now();
    uint64_t key = 0;
    if (int rc = nmi_read_key(s->ctx, s->key_dev, &key)) return rc;
    const auto t3 = clk::now();
    const int rc = nmi_decode_key(grid, key, out);
    float ms[8];
    if (nmi_get_timings(s->ctx, ms, nullptr) == NMI_OK) out->gpu_ms = ms[6];  // this rank's device time
    if (g_trace_n < 16)
      g_trace[g_trace_n++] = LevelTrace{(float)us(t0, t1), (float)us(t1, t2), (float)us(t2, t3), out->gpu_ms};
    if (trace)
      fprintf(stderr, "[nmi level] rank %d/%d grid %dx%dx%d x %dx%dx%d: enqueue %ld us, exchange call %ld us, "
                      "wait %ld us, device %.3f ms (cull %.3f render %.3f hist %.3f)\n",
              s->rank, s->world, grid->nS[0], grid->nS[1], grid->nS[2], grid->nW[0], grid->nW[1], grid->nW[2],
              us(t0, t1), us(t1, t2), us(t2, t3), ms[6], ms[0], ms[1], ms[4]);
    // some rank's record bins filled up: all ranks saw NMI_KEY_RETRY and redo the level (the
    // rank concerned sizes its bins exactly this time)
    if (rc != NMI_ERR_RETRY) {
      if (rc == NMI_ERR_NO_WINNER) nmi::set_error("every score is negative: no winner");
      return rc;
    }
  }
  nmi::set_error("nmi_relocalize_sharded: a rank's splat-record buffer overflowed twice");
  return NMI_ERR_CUDA;
}

}  // namespace

int nmi_relocalize(nmi_ctx* ctx, const float Twc_in[16], const nmi_grid* start_grid,
                   const nmi_flags* flags, const nmi_reloc_params* prm, nmi_reloc_result* out) {
  if (!ctx || !flags) {
    nmi::set_error("nmi_relocalize: null argument");
    return NMI_ERR_INVALID;
  }
  SingleSearch s{ctx, flags};
  return nmi_relocalize_with(single_level, &s, Twc_in, start_grid, prm, out);
}

int nmi_relocalize_sharded(nmi_ctx* ctx, const float Twc_in[16], const nmi_grid* start_grid,
                           const nmi_flags* flags, const nmi_reloc_params* prm, int rank, int world,
                           void* key_dev, nmi_exchange_fn exchange, void* user, nmi_reloc_result* out) {
  if (!ctx || !flags || !exchange || world < 1 || rank < 0 || rank >= world) {
    nmi::set_error("nmi_relocalize_sharded: bad argument");
    return NMI_ERR_INVALID;
  }
  if (!key_dev) key_dev = nmi_ctx_key_buffer(ctx);  // the context's own 8-byte exchange buffer
  if (!key_dev) {
    nmi::set_error("nmi_relocalize_sharded: no key buffer");
    return NMI_ERR_CUDA;
  }
  ShardedSearch s{ctx, flags, rank, world, key_dev, exchange, user};
  g_trace_n = 0;
  const auto t0 = std::chrono::steady_clock::now();
  const int rc = nmi_relocalize_with(sharded_level, &s, Twc_in, start_grid, prm, out);
  if (getenv("NMI_TRACE_LEVELS"))
    fprintf(stderr, "[nmi driver] rank %d/%d: %d levels, %ld us in nmi_relocalize_sharded\n", rank, world,
            out->iterations,
            (long)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count());
  return rc;
}

int nmi_last_level_trace(int level, float out_us[4]) {
  if (level < 0 || level >= g_trace_n || !out_us) return NMI_ERR_INVALID;
  out_us[0] = g_trace[level].enqueue_us;
  out_us[1] = g_trace[level].exchange_us;
  out_us[2] = g_trace[level].wait_us;
  out_us[3] = g_trace[level].device_ms * 1e3f;
  return NMI_OK;
}

}  // extern "C"
