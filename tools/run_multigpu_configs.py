"""BASELINE.json configs 4 and 5 on N GPUs of one box (strong scaling of ONE search):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/run_multigpu_configs.py [--frames 40] [--only C4,C5]
    python tools/run_multigpu_configs.py            # N = 1 (no process group)

C4  coarse-to-fine 3-level search, level 0 = 8^3 x 4^3 = 32 768 poses, every level sharded over
    the ranks by nmi_partition and combined by ONE 8-byte NCCL max-allreduce
    (csrc/driver.cpp nmi_relocalize_sharded; 3 allreduces per search).
C5  sequence: a synthetic frame per step, prior = true pose + drift, NMI pose correction every
    frame (3^6 grid, up to 4 levels), each correction sharded over the ranks.  The frame is
    uploaded by every rank (replicated, 2 MB); a stub stands in for ORB-SLAM2's tracker.

Times are wall clock around the synchronous driver call, max over ranks (the driver returns
after the last level's key came back, so this is the latency a tracker would see).
Rank 0 prints one JSON line per config.  Synthetic data only.
"""
import argparse
import json
import math
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from orbslam2_nmi_b200 import multigpu, synth  # noqa: E402
from orbslam2_nmi_b200.capi import Grid  # noqa: E402
from orbslam2_nmi_b200.search import NmiSearcher  # noqa: E402


def main():
    import torch
    import torch.distributed as dist

    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=40)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="C4,C5")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def emit(**kw):
        if rank == 0:
            print(json.dumps(kw), flush=True)

    sc = synth.make_scene("C2")  # same seed on every rank: the cloud is replicated
    s = NmiSearcher(local)
    s.set_scene(sc)
    key = torch.zeros(1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    want = a.only.split(",")

    if "C4" in want:
        s.set_frame(synth.frame_textured(sc.W, sc.H))
        g = Grid.make((8, 8, 8), (4, 4, 4), (0.2, 0.2, 0.5), (0.02, 0.02, 0.05))
        for _ in range(2):  # warm-up: buffers, bin-capacity feedback, NCCL channels
            out = multigpu.relocalize_sharded(s, sc.Twc, g, None, key, rank, world, threshold=0.0, max_iterations=3)
        ms = []
        for _ in range(a.reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = multigpu.relocalize_sharded(s, sc.Twc, g, None, key, rank, world, threshold=0.0,
                                              max_iterations=3)
            ms.append(max_over_ranks((time.perf_counter() - t0) * 1e3))
        med = float(np.median(ms))
        emit(config="C4 coarse-to-fine 3-level search, level 0 = 8^3 x 4^3 = 32768 poses, levels sharded over the "
                    "ranks, one 8-byte NCCL max-allreduce per level", n_gpus=world, levels=out.iterations,
             evals=out.n_evals, search_ms=med, rank0_device_ms=out.gpu_ms,
             level_grids=[[list(out.levels[i].grid.nS), list(out.levels[i].grid.nW)] for i in range(out.n_levels)], search_ms_all=ms, evals_per_s=out.n_evals / med * 1e3,
             nmi=out.nmi, best_s=list(out.best_s), best_w=list(out.best_w), pose_t=[out.Twc[3], out.Twc[7], out.Twc[11]],
             timing="wall clock around nmi_relocalize_sharded, max over ranks", scaling="strong")

    if "C5" in want:
        g0 = synth.default_grid()
        rng = np.random.default_rng(0)
        errs_before, errs_after, ms = [], [], []
        for k in range(a.frames):
            T = sc.Twc.copy()
            T[0, 3] += 0.05 * k
            T[1, 3] += 2.0 * math.sin(0.05 * k)
            gt = Grid.make((1, 1, 1), (1, 1, 1), (0.1,) * 3, (0.01,) * 3)
            s.render_cell(T, gt, 0, 0, 0)
            frame = synth.frame_from_render(s.get_render(0), seed=k)  # identical on every rank
            prior = T.copy()
            drift = rng.integers(-1, 2, size=3) * np.array([0.2, 0.2, 0.5], dtype=np.float32)
            ax, ay, az = -T[:3, 0], T[:3, 1], -T[:3, 2]
            prior[:3, 3] += (drift[0] * ax + drift[1] * ay + drift[2] * az).astype(np.float32)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            s.set_frame(frame)  # H2D of the frame is part of the per-frame cost
            out = multigpu.relocalize_sharded(s, prior, g0, None, key, rank, world, threshold=0.02)
            ms.append(max_over_ranks((time.perf_counter() - t0) * 1e3))
            got = np.array(out.Twc[:]).reshape(4, 4)
            errs_before.append(float(np.linalg.norm(prior[:3, 3] - T[:3, 3])))
            errs_after.append(float(np.linalg.norm(got[:3, 3] - T[:3, 3])))
        med = float(np.median(ms[2:] if len(ms) > 4 else ms))
        emit(config="C5 sequence: synthetic frames, NMI pose correction every frame (stub tracker), each correction "
                    "sharded over the ranks", n_gpus=world, frames=a.frames, ms_per_frame_median=med,
             frames_per_s=1e3 / med, seconds_for_1000_frames=med, mean_err_before_m=float(np.mean(errs_before)),
             mean_err_after_m=float(np.mean(errs_after)),
             corrected_fraction=float(np.mean(np.array(errs_after) < 0.5 * np.maximum(np.array(errs_before), 1e-9))),
             timing="wall clock per frame (frame H2D + sharded driver), max over ranks", scaling="strong")

    s.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
