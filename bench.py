#!/usr/bin/env python
"""bench.py -- NMI pose evaluations / s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one full pose-grid search (the hot path of Tracking::RelocalizeWithNMI,
src/Tracking.cc:1851-1985) over one synthetic frame: project + z-buffer nS views of
the cloud, warp the frame nW times, score nS*nW (render, warp) pairs, argmax.

Workload at N=1 is BASELINE.json configs[1] (ZU-MAV-shaped): 1920x1080 frame,
10M-point cloud, 4^3 x 4^3 = 4096 poses.  At N GPUs the synthetic-view axis grows to
64*N views (4096*N poses; N=8 is configs[3]'s 32 768-pose level): each rank scores
64 views x 64 warps, the cloud and frame are replicated, and the only exchange is
one 8-byte max-allreduce of the packed (score, index) key per search -> "weak".

Printed JSON (one line, rank 0): see the task contract.  `value` is device-timed with
the frame already in HBM; `e2e` goes through the public C ABI with the frame in
pinned HOST memory (H2D inside the timed region, key D2H at the end of every step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

_OUT = sys.stdout
METRIC = "nmi_pose_evals_per_s"
UNIT = "evals/s"
WORKLOAD = "C2 ZU-MAV-shaped search: 1920x1080 frame, 10M-point cloud, 4^3 synth x 4^3 warp = 4096 poses per GPU, 256 bins, SUC"


def grid_for(world: int):
    """64*world synthetic views x 64 warps; world=8 -> 8x8x8 views (configs[3] level 0)."""
    from orbslam2_nmi_b200 import synth

    ns = {1: (4, 4, 4), 2: (8, 4, 4), 4: (8, 8, 4), 8: (8, 8, 8)}.get(world)
    if ns is None:
        ns = (4, 4, 4 * world)
    return synth.default_grid(ns, (4, 4, 4))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._halt = threading.Event()

    def run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                     str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_bytes():
    """dram bytes per launch of the histogram kernel from the committed ncu capture, or None."""
    p = ROOT / "profiles" / "hist_kernel_traffic.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["dram_bytes_per_launch"])
        except Exception:
            return None
    return None


# ------------------------------------------------------------------ CPU baseline ----
def cpu_baseline_run(scene, frame, threads: int, sample=((4, 4, 4), (4, 4, 4))):
    """Times the CPU oracle (the `port` of the reference's arithmetic; the reference itself
    has no CPU NMI path and cannot be built here) on a bounded sample of the
    same workload: the whole C2 search once (full-size frame and cloud, all 4^3 x 4^3 = 4096
    poses), a few seconds on a multi-core host."""
    from oracle import oracle_py as oracle  # test infrastructure: cpu_baseline leg only
    from orbslam2_nmi_b200 import synth

    g = synth.default_grid(*sample)
    t0 = time.perf_counter()
    scores, _, _ = oracle.search_points(scene, scene.Twc, g, scene.xyzi, frame, threads=threads)
    dt = time.perf_counter() - t0
    n = g.n_pose
    cpu_baseline_run.last_scores = scores  # the checker's answer, for the parity line of the N=1 run
    return n / dt, dt, f"{n}-pose grid {sample[0]}x{sample[1]} at full 1920x1080 / 10M points ({dt:.1f} s)"


def reference_gpu_kernels(render, warped, iters: int = 20):
    """The reference's OWN CUDA routine -- CUDAF::NMIWithCuda_noMask from oracle/_ref (NMI.cu +
    kernel.cu compiled unmodified for sm_100a, oracle/Makefile.ref) -- timed per evaluation on
    this GPU with the pair resident: what the reference pays per pose for a7-a10 alone (render and
    warp excluded; its GL / NPP stages cannot run here).  Part of the baseline leg; None when the
    library was not built or no GPU is visible."""
    try:
        from oracle import ref_py  # test infrastructure: baseline leg only

        if not ref_py.available():
            return None
        ms, score = ref_py.time_per_eval(render, warped, iters)
        st = ref_py.stages(render, warped)
        sa, sb, sab = (float(x) for x in st["sums"])
        # its three-block last kernel reads the other blocks' totals unsynchronised (NMI.cu:340-362):
        # `score` is what it returned, `score_from_its_totals` the formula over its own three sums
        race_free = float(np.float32(2.0) * (np.float32(1.0) - (np.float32(-sab) / (np.float32(-sa) + np.float32(-sb))))) \
            if (sa or sb or sab) else 0.0
        return {"value": 1e3 / ms, "unit": UNIT, "ms_per_eval": ms, "kind": "reference",
                "score_from_its_totals": race_free,
                "sample": f"{iters} calls of CUDAF::NMIWithCuda_noMask on one resident {render.shape[1]}x{render.shape[0]} pair "
                          "(histogram + entropy + score only; renders and warps supplied)",
                "score": float(score), "build": ref_py.describe()}
    except Exception as e:  # a baseline that cannot run must not take the bench line down
        return {"unavailable": f"{type(e).__name__}: {e}"}


def run_reference(args):
    """--impl reference: the CPU restatement on the box's host cores (bounded sample)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from orbslam2_nmi_b200 import synth

    threads = os.cpu_count() or 1
    scene = synth.make_scene("C2")
    frame = synth.frame_textured(scene.W, scene.H)
    for _ in range(args.warmup and 1):  # one warm-up pass is plenty for a CPU loop
        cpu_baseline_run(scene, frame, threads, sample=((2, 1, 1), (2, 1, 1)))
    vals, total = [], 0.0
    for _ in range(args.steps):
        v, dt, sample = cpu_baseline_run(scene, frame, threads)
        vals.append(v)
        total += dt
    value = float(np.mean(vals))
    ref_gpu = None
    try:
        import torch

        if torch.cuda.is_available():
            from oracle import oracle_py as oracle

            g1 = synth.default_grid((1, 1, 1), (1, 1, 1))
            _, renders, warps = oracle.search_points(scene, scene.Twc, g1, scene.xyzi, frame, keep_images=True,
                                                     threads=threads)
            ref_gpu = reference_gpu_kernels(renders[0], warps[0])
    except Exception as e:
        ref_gpu = {"unavailable": f"{type(e).__name__}: {e}"}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32+f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "note": "each step is one whole 4096-pose C2 search on the host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "reference_gpu_kernels": ref_gpu},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_OUT, flush=True)
    return 0


# ------------------------------------------------------------------------ GPU arm ----
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from orbslam2_nmi_b200 import build, synth
    from orbslam2_nmi_b200.search import NmiSearcher

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the NMI search has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    build.build_cuda()

    scene = synth.make_scene("C2", extent=args.extent)  # same seed on every rank: the cloud is replicated
    frame = {"textured": synth.frame_textured, "uniform": synth.frame_uniform,
             "constant": synth.frame_constant, "sky": synth.frame_sky,
             "smooth": synth.frame_smooth}[args.frame](scene.W, scene.H)
    grid = grid_for(world)
    searcher = NmiSearcher(local)
    searcher.set_scene(scene)
    searcher.set_frame(frame)
    flags = searcher.flags(variant=args.variant)
    stream = torch.cuda.ExternalStream(searcher.stream(), device=local)
    key = torch.zeros(1, dtype=torch.int64, device="cuda")
    h_frame = torch.from_numpy(frame.copy()).pin_memory()
    h_key = torch.zeros(1, dtype=torch.int64).pin_memory()
    torch.cuda.synchronize()
    evals_per_step = grid.n_pose  # whole job, all ranks
    per_rank_pairs = grid.n_pose // world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        with torch.cuda.stream(stream):
            searcher.search_enqueue(scene.Twc, grid, flags, rank, world, key.data_ptr())
            if world > 1:
                dist.all_reduce(key, op=dist.ReduceOp.MAX)
        searcher.sync()  # stream sync + surfaces a record-buffer overflow of the enqueued search

    def step_e2e():
        # public API with HOST buffers: frame H2D + search + winner key D2H, every step
        searcher.set_frame(h_frame.numpy())
        with torch.cuda.stream(stream):
            searcher.search_enqueue(scene.Twc, grid, flags, rank, world, key.data_ptr())
            if world > 1:
                dist.all_reduce(key, op=dist.ReduceOp.MAX)
            h_key.copy_(key, non_blocking=True)
        searcher.sync()
        return int(h_key.item())

    def timed(fn, steps, collect=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
            if collect is not None:
                collect()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step_device()

    stage_ms, launches = [], []

    def collect():
        t, n = searcher.timings()
        stage_ms.append(t)
        launches.append(n)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_dev = timed(step_device, args.steps, collect)
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if sampler else None
    winner = searcher.decode(grid, int(key.item()))

    # CPU baseline on rank 0 at N=1 only (bounded sample, all host cores)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, dt, sample = cpu_baseline_run(scene, frame, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
        # the oracle just scored the whole workload: check the GPU's 4096 scores against it
        # (outside every timed region; the checker is never on the measured path)
        from oracle import oracle_py as oracle

        want = cpu_baseline_run.last_scores
        got = searcher.search(scene.Twc, grid, flags, want_scores=True)
        rel = np.abs(got.scores.astype(np.float64) - want) / np.maximum(np.abs(want.astype(np.float64)), 1e-30)
        cpu["parity"] = {"poses": int(want.size), "max_rel_err": float(rel.max()),
                         "same_winner": bool(got.best_index == oracle.argmax(want)[0]),
                         "tolerance": 1e-5}
        t1 = time.perf_counter()
        oracle.search_points(scene, scene.Twc, synth.default_grid((1, 1, 1), (1, 1, 1)), scene.xyzi, frame,
                             threads=1)
        cpu["single_eval_ms_1core"] = 1e3 * (time.perf_counter() - t1)  # one evaluation, one core
        # the reference's own CUDA kernels on this GPU, on the pair (render 0, warp 0) of that search
        cpu["reference_gpu_kernels"] = reference_gpu_kernels(searcher.get_render(0), searcher.get_warp(0))

    if rank == 0:
        P = scene.W * scene.H
        hist_ms = float(np.mean([t["hist_score"] for t in stage_ms]))
        mean_stage = {k: float(np.mean([t[k] for t in stage_ms])) for k in stage_ms[0]}
        # the histogram stage reads one render + one warped frame per evaluation
        hist_bytes = per_rank_pairs * 2.0 * P + 4.0 * per_rank_pairs
        peak, peak_src = measured_peak_gbs()
        achieved = hist_bytes / (hist_ms * 1e-3) / 1e9
        # SURVEY 8(d) whole-search algorithmic bytes per rank (cloud re-read per view, as the
        # reference's per-view GL draw does)
        nvl, nwl = grid.n_synth // world, grid.n_warp
        search_bytes = nvl * (16.0 * scene.xyzi.shape[0] + P) + nwl * 2.0 * P + per_rank_pairs * 2.0 * P + 8.0 * per_rank_pairs
        value = evals_per_step * args.steps / (ms_dev * 1e-3)
        e2e = evals_per_step * args.steps / (ms_e2e * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/u32+f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "poses_per_step": evals_per_step,
                       "grid": {"nS": list(grid.nS), "nW": list(grid.nW)},
                       "l2": "inputs larger than L2 (160 MB cloud, ~1.2 GB of splat records, 266 MB of renders + warps per step)",
                       "hist_variant": args.variant, "frame": args.frame + " synthetic",
                       "winner": {"index": winner.best_index, "score": winner.best_score}},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(P + 256 * 1024), "d2h_bytes_per_step": 8},
            "gpu_launches": int(sum(launches)) * 2,  # device-timed loop + e2e loop
            "roofline": {"bound": "hbm", "kernel": "joint_hist_score_persistent_kernel",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic_bytes(), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": hist_bytes,
                         "kernel_ms": hist_ms, "kernel_share_of_step": hist_ms / mean_stage["total"],
                         "smem_atomics_per_s": per_rank_pairs * P / (hist_ms * 1e-3),
                         "search_algorithmic_GBps": search_bytes / (mean_stage["total"] * 1e-3) / 1e9,
                         "note": "HBM is the contract's denominator; the kernel's measured limiter is the "
                                 "shared-memory data pipe (ncu l1tex__data_pipe_lsu_wavefronts 91 % of peak, "
                                 "3.75 wavefronts per warp-level ATOMS = random bank collisions; DRAM traffic is "
                                 "8 % of the algorithmic bytes): profiles/r01_hist_ncu_summary.txt"},
            "stage_ms": mean_stage,
            "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _quiet_stdout():
    """Route fd 1 to stderr for the whole run (NCCL prints its version banner to stdout) and
    return a file object on the real stdout for the single JSON line."""
    real = os.dup(1)
    sys.stdout.flush()
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def main():
    global _OUT
    _OUT = _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", type=int, default=0, help="histogram kernel variant (0..10)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--frame", default="textured", choices=["textured", "uniform", "constant", "sky", "smooth"])
    ap.add_argument("--extent", type=float, default=None,
                    help="half-width of the synthetic cloud in metres (default 40; small values leave "
                         "render background, a stress case -- not the BASELINE workload)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
