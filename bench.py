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
`configs` carries the other BASELINE.json configurations, each with its own parity flag
(C1 single evaluation, C2 on the uniform / planted frames of SURVEY 8(d), C3 mesh, C4 the
3-level 32 768-pose search -- sharded over the ranks at N > 1 = strong scaling --, C5 a
sequence of >= 200 frames with one C2-sized 4096-pose search per frame, pageable frames).
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


def core_config(world: int, grid) -> dict:
    """The `config` object both arms print (the reference arm scores the N=1 grid)."""
    return {"workload": WORKLOAD, "poses_per_step": grid.n_pose,
            "grid": {"nS": list(grid.nS), "nW": list(grid.nW)}, "frame": "textured synthetic",
            "l2": "inputs larger than L2 (160 MB cloud, ~1.2 GB of splat records, 266 MB of renders + warps per step)"}


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


def measured_atomics_peak(sm_mhz):
    """Shared-memory atomic peak for the histogram roofline: tools/ubench_atoms (built by
    orbslam2_nmi_b200.build) run live on this GPU -- independent uniformly random words per lane,
    the access pattern of exact per-pixel counting on data-independent bins -- else the number
    committed in profiles/r02_ubench_shared_atomics.txt."""
    exe = ROOT / "orbslam2_nmi_b200" / "_lib" / "ubench_atoms"
    lanes, pattern, src = 9.05, 7.59, "committed (profiles/r02_ubench_shared_atomics.txt)"
    sms = 148
    try:
        out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120).stdout
        for ln in out.splitlines():
            if ln.startswith("JSON "):
                d = json.loads(ln[5:])
                lanes, pattern, sms = d["random_lanes_per_clk_sm"], d["hist_pattern_lanes_per_clk_sm"], d["sms"]
                src = "measured live (tools/ubench_atoms.cu, uniformly random words, in-kernel clock)"
    except Exception:
        pass
    mhz = sm_mhz or 1965.0
    return {"lanes_per_clk_sm": lanes, "hist_pattern_lanes_per_clk_sm": pattern, "sms": sms, "sm_mhz": mhz,
            "atomics_per_s": lanes * sms * mhz * 1e6, "source": src}


def roofline_block(pairs, P, hist_ms, mean_stage, hist_bytes, search_bytes, hbm_peak, hbm_src, clocks):
    """Roofline of the dominant kernel (joint_hist_score_persistent_kernel).  Its bound is the
    shared-memory atomic rate: one ATOMS lane-operation per pixel and evaluation, on bins
    the data picks -- the peak is what independent uniformly random 32-bit words sustain on this GPU
    (measured live).  The HBM view the contract names is carried alongside: the kernel moves 0.07 of
    its algorithmic bytes through DRAM (L2 reuse across the pair tiles), so HBM is not what limits it."""
    t = hist_ms * 1e-3
    atom = measured_atomics_peak((clocks or {}).get("sm_mhz"))
    ach_atoms = pairs * P / t
    ach_gbs = hist_bytes / t / 1e9
    traffic = ncu_traffic_bytes()
    return {"bound": "smem_atomic", "kernel": "joint_hist_score_persistent_kernel",
            "achieved": ach_atoms / 1e9, "peak": atom["atomics_per_s"] / 1e9, "unit": "Gatom/s",
            "frac": ach_atoms / atom["atomics_per_s"], "traffic": traffic,
            "algorithmic_atomics_per_launch": pairs * P, "kernel_ms": hist_ms,
            "kernel_share_of_step": hist_ms / mean_stage["total"],
            "achieved_lanes_per_clk_sm": ach_atoms / (atom["sms"] * atom["sm_mhz"] * 1e6),
            "peak_detail": atom,
            "frac_of_own_pattern_ceiling": ach_atoms / (atom["hist_pattern_lanes_per_clk_sm"] * atom["sms"] * atom["sm_mhz"] * 1e6),
            "hbm": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                    "peak_source": hbm_src, "algorithmic_bytes_per_launch": hist_bytes, "traffic": traffic,
                    "traffic_over_algorithmic": (traffic / hist_bytes) if traffic else None,
                    "note": "algorithmic 2P bytes per evaluation / kernel time; real DRAM traffic is ~7 % of that "
                            "(inputs served from L2), so this fraction is L2 reuse, not an HBM limit"},
            "search_algorithmic_GBps": search_bytes / (mean_stage["total"] * 1e-3) / 1e9,
            "note": "ncu: l1tex throughput 93 % of peak, 3.78 shared-memory wavefronts per warp-level ATOMS "
                    "(32 random lanes over 32 banks: expected maximum multiplicity 3.5), 11.8 instructions per ATOMS, "
                    "issue slots 64 % busy (profiles/r02_hist_ncu_summary.txt, profiles/r02_ubench_shared_atomics.txt)"}


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
        "data": "synthetic", "config": core_config(args.gpus, grid_for(args.gpus)),
        "note": "each step scores one whole 4096-pose C2 search (one GPU's share of the workload) on all host cores",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "reference_gpu_kernels": ref_gpu},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_OUT, flush=True)
    return 0


# ------------------------------------------------- the other BASELINE.json configs ----
def _bits_equal(a, b):
    return bool(np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32))


def config_c1(local):
    """configs[0]: one NMI evaluation, 752x480 frame vs the render of a 1M-point cloud at one pose."""
    from oracle import oracle_py as oracle  # checker
    from orbslam2_nmi_b200 import synth
    from orbslam2_nmi_b200.search import NmiSearcher

    sc = synth.make_scene("C1")
    frame = synth.frame_textured(sc.W, sc.H)
    g = synth.default_grid((1, 1, 1), (1, 1, 1))
    s = NmiSearcher(local)
    try:
        s.set_scene(sc)
        s.set_frame(frame)
        for _ in range(3):
            r = s.search(sc.Twc, g)
        ms = []
        for _ in range(20):
            t0 = time.perf_counter()
            r = s.search(sc.Twc, g)
            ms.append((time.perf_counter() - t0) * 1e3)
        dev = s.timings()[0]
        # the reference's own call granularity: CUDAF::NMIWithCuda_noMask on a resident pair
        s.warp_cells(g)
        h = s.render_cell(sc.Twc, g, 0, 0, 0)
        p = s.warp_ptr(g, 0, 0, 0)
        for _ in range(5):
            s.eval_pair(p, h)
        t0 = time.perf_counter()
        for _ in range(200):
            sc1 = s.eval_pair(p, h)
        per_call_us = (time.perf_counter() - t0) / 200 * 1e6
        want, _, _ = oracle.search_points(sc, sc.Twc, g, sc.xyzi, frame)
        return {"workload": "C1 single NMI evaluation: 752x480 frame vs render of a 1M-point cloud, 1 pose, 256 bins",
                "search_ms_wall": float(np.median(ms)), "search_ms_device": dev["total"],
                "stage_ms": dev, "eval_pair_call_us": per_call_us,
                "eval_pair_note": "CUDAF::NMIWithCuda_noMask drop-in (nmi_eval_pair): wall clock per call incl. the blocking 4-byte D2H",
                "score": float(r.best_score),
                "parity": {"checker": "oracle", "score_bit_identical": _bits_equal(r.best_score, want[0]) and _bits_equal(sc1, want[0]),
                           "rel_err": float(abs(float(r.best_score) - float(want[0])) / max(abs(float(want[0])), 1e-30))}}
    finally:
        s.close()


def config_c2_frames(searcher, scene, grid, flags):
    """configs[1] on the other two frame variants of SURVEY 8(d): `uniform` (iid U{0..255}, the
    throughput floor for the histogram) and `planted` (LUT + noise of the oracle's render at a hidden
    grid cell: the argmax is known by construction)."""
    from oracle import oracle_py as oracle  # checker
    from orbslam2_nmi_b200 import synth

    out = {}

    def time_frame(frame):
        searcher.set_frame(frame)
        for _ in range(3):
            searcher.search(scene.Twc, grid, flags)
        ms = [searcher.search(scene.Twc, grid, flags).gpu_ms for _ in range(8)]
        return float(np.median(ms)), searcher.timings()[0]

    ms, st = time_frame(synth.frame_uniform(scene.W, scene.H))
    out["uniform"] = {"search_ms": ms, "evals_per_s": grid.n_pose / ms * 1e3, "hist_ms": st["hist_score"]}
    # flat image content (hot-bin skipping through the image marginals, hist.cu): a frame whose top 35 % is a
    # saturated sky, and a constant frame; the winner is checked against the side-table build of the same kernel
    for name, fr in (("sky", synth.frame_sky(scene.W, scene.H)), ("constant", synth.frame_constant(scene.W, scene.H))):
        ms, st = time_frame(fr)
        r = searcher.search(scene.Twc, grid, flags)
        out[name] = {"search_ms": ms, "evals_per_s": grid.n_pose / ms * 1e3, "hist_ms": st["hist_score"],
                     "hist_path": searcher.last_hist_path(),
                     "winner": {"index": int(r.best_index), "score": r.best_score}}
    # planted: translation cell (2, 1, 3), rotation cell (1, 1, 1) = the cell whose evaluated angles
    # are 0 for n = 4 (image.cpp:77: start = -trunc((n-1)/2) * step)
    cell_s, cell_w = (2, 1, 3), (1, 1, 1)
    cam = oracle.camera(scene)
    t = oracle.cell_translation(scene.Twc, grid, *cell_s)
    render = oracle.render_points(scene, scene.Twc, t, scene.xyzi)[1]
    frame = synth.frame_from_render(render, seed=7)
    ms, st = time_frame(frame)
    r = searcher.search(scene.Twc, grid, flags)
    out["planted"] = {"search_ms": ms, "evals_per_s": grid.n_pose / ms * 1e3, "hist_ms": st["hist_score"],
                      "planted_cell": {"s": list(cell_s), "w": list(cell_w)},
                      "winner": {"s": list(r.best_s), "w": list(r.best_w), "score": r.best_score},
                      "parity": {"checker": "planted pose (oracle render at the hidden cell)",
                                 "winner_is_planted_cell": bool(tuple(r.best_s) == cell_s and tuple(r.best_w) == cell_w)}}
    del cam
    return out


def config_c3(local):
    """configs[2]: 848x480 frame vs a 2M-triangle mesh, 1024 poses, 64 bins."""
    from oracle import oracle_py as oracle  # checker
    from orbslam2_nmi_b200 import synth
    from orbslam2_nmi_b200.capi import Grid
    from orbslam2_nmi_b200.search import NmiSearcher

    c = synth.CONFIGS["C3"]
    verts, tris = synth.make_mesh(1000, 1000)
    uv = synth.make_mesh_uv(verts, tris, repeats=6.0)
    tex = synth.make_texture(1024, 1024)
    frame = synth.frame_textured(c["W"], c["H"])
    Twc = synth.prior_pose()
    s = NmiSearcher(local)
    try:
        s.set_camera(c["W"], c["H"], c["fx"], c["fy"], c["cx"], c["cy"], synth.ZN, synth.ZF, 3.0)
        s.set_mesh_textured(verts, tris, uv, tex)  # Rendering<1>: per-fragment texture shading
        s.set_frame(frame)
        g = Grid.make((4, 4, 4), (4, 4, 1), (0.2, 0.2, 0.5), (0.02, 0.02, 0.05))
        fl = s.flags(bins=64)
        for _ in range(3):
            s.search(Twc, g, fl)
        ms = [s.search(Twc, g, fl).gpu_ms for _ in range(8)]
        st = s.timings()[0]
        # parity on a bounded sample at full size: a 2 x 1 x 1 x (2 x 2 x 1) sub-grid against the oracle
        gs = Grid.make((2, 1, 1), (2, 2, 1), (0.2, 0.2, 0.5), (0.02, 0.02, 0.05))
        got = s.search(Twc, gs, fl, want_scores=True)
        sc = synth.Scene(c["W"], c["H"], c["fx"], c["fy"], c["cx"], c["cy"], synth.ZN, synth.ZF, 3.0, None, Twc)
        want, renders, _ = oracle.search_mesh_tex(sc, Twc, gs, verts, tris, uv, tex, frame, bins=64, keep_images=True)
        same_render = bool(np.array_equal(s.get_render(0), renders[0]))
        med = float(np.median(ms))
        return {"workload": "C3 Newer-College-shaped: 848x480 frame vs 2M-triangle mesh with a 1024^2 texture shaded per fragment, 4^3 x (4x4x1) = 1024 poses, 64 bins",
                "search_ms": med, "evals_per_s": g.n_pose / med * 1e3, "stage_ms": st,
                "parity": {"checker": "oracle on a 8-pose sub-grid at full size", "renders_bit_exact": same_render,
                           "scores_bit_identical": bool(np.array_equal(got.scores.view(np.uint32), want.view(np.uint32))),
                           "same_winner": bool(got.best_index == oracle.argmax(want)[0])}}
    finally:
        s.close()


def config_c4(searcher, scene, key, rank, world, dist):
    """configs[3]: coarse-to-fine search, level 0 = 8^3 x 4^3 = 32 768 poses, 3 levels
    (Tracking.cc:2088-2130).  At N > 1 every level is sharded over the ranks and combined by one
    8-byte NCCL max-allreduce (nmi_relocalize_sharded): STRONG scaling of one search."""
    import torch
    from orbslam2_nmi_b200 import multigpu, synth
    from orbslam2_nmi_b200.capi import Grid

    searcher.set_frame(synth.frame_textured(scene.W, scene.H))
    g = Grid.make((8, 8, 8), (4, 4, 4), (0.2, 0.2, 0.5), (0.02, 0.02, 0.05))

    def once():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = multigpu.relocalize_sharded(searcher, scene.Twc, g, None, key, rank, world, threshold=0.0,
                                          max_iterations=3)
        t = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return out, float(t.item())

    once()
    runs = [once() for _ in range(3)]
    out = runs[-1][0]
    med = float(np.median([m for _, m in runs]))
    res = {"workload": "C4 coarse-to-fine search: level 0 = 8^3 x 4^3 = 32768 poses, 3 levels, 1920x1080 / 10M points",
           "n_gpus": world, "scaling": "strong", "levels": int(out.iterations), "evals": int(out.n_evals),
           "search_ms": med, "evals_per_s": out.n_evals / med * 1e3,
           "level_grids": [[list(out.levels[i].grid.nS), list(out.levels[i].grid.nW)] for i in range(out.n_levels)],
           "winner": {"s": list(out.best_s), "w": list(out.best_w), "nmi": float(out.nmi)},
           "timing": "wall clock around the synchronous driver call, max over ranks, median of 3"}
    if rank == 0:
        # parity: the level-0 winner's score recomputed by the oracle (one render, one warp, one evaluation)
        from oracle import oracle_py as oracle  # checker

        l0 = out.levels[0]
        g0 = Grid.make(tuple(l0.grid.nS), tuple(l0.grid.nW), tuple(l0.grid.stepT), tuple(l0.grid.stepR))
        bs, bw = tuple(l0.best_s), tuple(l0.best_w)
        t = oracle.cell_translation(scene.Twc, g0, *bs)
        render = oracle.render_points(scene, scene.Twc, t, scene.xyzi)[1]
        warped = oracle.warp(searcher_frame(searcher, scene), oracle.cell_homography_inv(scene, g0, *bw))
        J, HA, HB = oracle.joint_hist(render, warped)
        want = oracle.score_stages_f32(J, HA, HB, scene.W * scene.H)["score"]
        res["parity"] = {"checker": "oracle re-evaluation of the level-0 winner",
                         "level0_winner_score_bit_identical": _bits_equal(l0.nmi, want)}
    return res


def searcher_frame(searcher, scene):
    from orbslam2_nmi_b200 import synth

    return synth.frame_textured(scene.W, scene.H)


def config_c5(searcher, scene, key, rank, world, dist, frames):
    """configs[4] at spec: a sequence of `frames` synthetic frames, ONE C2-sized 4096-pose search per
    frame (Tracking::RelocalizeWithNMI on every frame), the frame arriving as a pageable host buffer
    (what Image::loadOriginal(cv::Mat) passes: staged through pinned memory inside the timed region).
    A stub stands in for ORB-SLAM2's tracker (the vocabulary blob is absent): the
    prior moves along a smooth trajectory and the true pose is the prior displaced to a grid cell
    that changes every frame, so every search has a known answer.  At N > 1 each search is sharded
    over the ranks (frame replicated, 8-byte allreduce): strong scaling."""
    import torch
    from orbslam2_nmi_b200 import multigpu, synth
    from orbslam2_nmi_b200.capi import Grid

    g = synth.default_grid((4, 4, 4), (4, 4, 4))
    g1 = Grid.make((4, 4, 4), (1, 1, 1), (0.2, 0.2, 0.5), (0.02, 0.02, 0.05))
    flags = searcher.flags()
    rng = np.random.default_rng(11)
    lut = np.clip(np.rint(255.0 * (np.arange(256) / 255.0) ** 0.7), 0, 255).astype(np.int16)
    noise = [np.rint(4.0 * np.random.default_rng(100 + i).standard_normal((scene.H, scene.W))).astype(np.int16)
             for i in range(4)]
    priors, cells, frame_list = [], [], []
    for k in range(frames):  # untimed: the frames of the sequence (our renderer; 2 of them re-checked by the oracle)
        T = scene.Twc.copy()
        T[0, 3] += 0.02 * k
        T[1, 3] += 0.5 * np.sin(0.05 * k)
        cell = (int(rng.integers(0, 4)), int(rng.integers(0, 4)), int(rng.integers(0, 4)))
        searcher.render_cell(T, g1, *cell)
        render = searcher.get_render(0)
        frame = np.clip(lut[render] + noise[k % 4], 0, 255).astype(np.uint8)
        priors.append(T)
        cells.append(cell)
        frame_list.append(frame)
    ms, found, winners = [], 0, []
    for k in range(frames):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        searcher.set_frame(frame_list[k])  # pageable numpy buffer -> pinned staging -> H2D
        r = multigpu.sharded_search(searcher, priors[k], g, flags, key, rank, world,
                                    stream=torch.cuda.ExternalStream(searcher.stream()))
        t = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms.append(float(t.item()))
        ok = tuple(r.best_s) == cells[k] and tuple(r.best_w) == (1, 1, 1)
        found += int(ok)
        winners.append(r)
    steady = ms[3:] if len(ms) > 6 else ms
    res = {"workload": f"C5 sequence: {frames} synthetic 1920x1080 frames, one 4096-pose C2-sized search per frame, "
                       "10M-point cloud, pageable frames uploaded every frame",
           "n_gpus": world, "scaling": "strong" if world > 1 else "n/a", "frames": frames,
           "ms_per_frame_mean": float(np.mean(steady)), "ms_per_frame_median": float(np.median(steady)),
           "ms_per_frame_p99": float(np.percentile(steady, 99)), "ms_per_frame_max": float(np.max(steady)),
           "slowest_frames": [[int(i) + (3 if len(ms) > 6 else 0), float(steady[i])] for i in np.argsort(steady)[-3:][::-1]],
           "frames_per_s": 1e3 / float(np.mean(steady)),
           "seconds_for_1000_frames": float(np.mean(steady)), "evals_per_s": g.n_pose / float(np.mean(steady)) * 1e3,
           "h2d_bytes_per_frame": int(scene.W * scene.H), "timing": "wall clock per frame (set_frame + search + key read-back), max over ranks",
           "parity": {"checker": "planted pose per frame", "frames_with_planted_winner": found,
                      "fraction": found / frames}}
    if rank == 0:
        from oracle import oracle_py as oracle  # checker

        same = []
        for k in (0, frames // 2):
            t = oracle.cell_translation(priors[k], g, *winners[k].best_s)
            render = oracle.render_points(scene, priors[k], t, scene.xyzi)[1]
            warped = oracle.warp(frame_list[k], oracle.cell_homography_inv(scene, g, *winners[k].best_w))
            J, HA, HB = oracle.joint_hist(render, warped)
            same.append(_bits_equal(winners[k].best_score, oracle.score_stages_f32(J, HA, HB, scene.W * scene.H)["score"]))
        res["parity"]["winner_score_bit_identical_frames_checked"] = same
    return res


def config_c5_small_grid(searcher, scene, key, rank, world, dist, frames=60):
    """The reference's own per-frame load (Tracking.cc:2088-2130 with the default kernel of
    nmiSearchKernel.cpp:104-141): a 3^3 x 3^3 = 729-pose grid, up to three levels with the steps
    halved, per frame.  Small grids are where fixed per-level costs (cull, warps, launches, the key
    round trip) show; at N > 1 each level is sharded and combined by the 8-byte allreduce.  The
    per-level host costs come from nmi_last_level_trace."""
    import torch
    from orbslam2_nmi_b200 import multigpu, synth
    from orbslam2_nmi_b200.capi import Grid

    searcher.set_frame(synth.frame_textured(scene.W, scene.H))
    g = Grid.make((3, 3, 3), (3, 3, 3), (0.2, 0.2, 0.5), (0.02, 0.02, 0.05))
    ms, traces, out = [], [], None
    for k in range(frames):
        T = scene.Twc.copy()
        T[0, 3] += 0.01 * k
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = multigpu.relocalize_sharded(searcher, T, g, None, key, rank, world, threshold=0.0, max_iterations=3)
        t = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms.append(float(t.item()))
        traces.append(searcher.level_trace())
    steady = ms[5:]
    lv = [t for t in traces[5:] if len(t) == len(traces[-1])]
    mean_trace = [{k2: float(np.mean([t[i][k2] for t in lv])) for k2 in lv[0][i]} for i in range(len(lv[0]))] if lv else []
    return {"workload": f"C5 small grid: {frames} frames, coarse-to-fine 3^3 x 3^3 = 729 poses per level, 3 levels per frame, "
                        "1920x1080 / 10M points",
            "n_gpus": world, "scaling": "strong" if world > 1 else "n/a", "levels": int(out.iterations),
            "evals_per_frame": int(out.n_evals), "ms_per_frame_mean": float(np.mean(steady)),
            "ms_per_frame_median": float(np.median(steady)), "evals_per_s": out.n_evals / float(np.mean(steady)) * 1e3,
            "per_level_rank0_us": mean_trace,
            "timing": "wall clock around the synchronous driver call, max over ranks; per-level figures: this rank's host clock"}


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
        if world > 1:
            with torch.cuda.stream(stream):
                searcher.search_enqueue(scene.Twc, grid, flags, rank, world, key.data_ptr())
                dist.all_reduce(key, op=dist.ReduceOp.MAX)
        else:  # the library enqueues on its own stream; torch's current stream only matters for the collective
            searcher.search_enqueue(scene.Twc, grid, flags, rank, world, key.data_ptr())
        searcher.sync()  # stream sync + surfaces a record-buffer overflow of the enqueued search

    h_frame_np = h_frame.numpy()

    def step_e2e():
        # public API with HOST buffers: frame H2D + search + winner key D2H, every step
        searcher.set_frame(h_frame_np)
        if world == 1:  # the C ABI alone: enqueue, then nmi_read_key (8-byte D2H + stream sync)
            searcher.search_enqueue(scene.Twc, grid, flags, rank, world, key.data_ptr())
            return searcher.read_key(key.data_ptr())
        with torch.cuda.stream(stream):
            searcher.search_enqueue(scene.Twc, grid, flags, rank, world, key.data_ptr())
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
    ms_dev = timed(step_device, args.steps)
    for _ in range(min(args.steps, 8)):  # per-stage event times + launch counts: read outside the timed region
        step_device()
        collect()
    for _ in range(max(args.warmup, 3)):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    # the same end-to-end step with the frame in PAGEABLE host memory (what Image::loadOriginal's
    # cv::Mat is): nmi_set_frame stages it through one of two pinned buffers inside the timed region
    pageable = frame.copy()

    def step_e2e_pageable():
        searcher.set_frame(pageable)
        if world == 1:
            searcher.search_enqueue(scene.Twc, grid, flags, rank, world, key.data_ptr())
            return searcher.read_key(key.data_ptr())
        with torch.cuda.stream(stream):
            searcher.search_enqueue(scene.Twc, grid, flags, rank, world, key.data_ptr())
            dist.all_reduce(key, op=dist.ReduceOp.MAX)
            h_key.copy_(key, non_blocking=True)
        searcher.sync()

    for _ in range(max(args.warmup, 3)):
        step_e2e_pageable()
    ms_e2e_pg = timed(step_e2e_pageable, args.steps)
    clocks = sampler.stop() if sampler else None
    winner = searcher.decode(grid, int(key.item()))

    # the other BASELINE configs (outside the timed regions above; each times itself)
    configs = {}
    if not args.no_configs:
        if world > 1:  # collectives inside: an exception on one rank must end the job, not hang the others
            configs["C4"] = config_c4(searcher, scene, key, rank, world, dist)
            configs["C5"] = config_c5(searcher, scene, key, rank, world, dist, args.frames)
            configs["C5_small_grid"] = config_c5_small_grid(searcher, scene, key, rank, world, dist)
        else:
            for name, fn in (("C2_frames", lambda: config_c2_frames(searcher, scene, grid, flags)),
                             ("C4", lambda: config_c4(searcher, scene, key, rank, world, None)),
                             ("C5", lambda: config_c5(searcher, scene, key, rank, world, None, args.frames)),
                             ("C5_small_grid", lambda: config_c5_small_grid(searcher, scene, key, rank, world, None)),
                             ("C1", lambda: config_c1(local)), ("C3", lambda: config_c3(local))):
                try:
                    configs[name] = fn()
                except Exception as e:  # reported, never hidden; the headline line still prints
                    import traceback

                    configs[name] = {"error": f"{type(e).__name__}: {e}", "where": traceback.format_exc().splitlines()[-3:]}
        searcher.set_frame(frame)

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
            "config": dict(core_config(world, grid), frame=args.frame + " synthetic"),
            "result": {"hist_variant": args.variant, "winner": {"index": winner.best_index, "score": winner.best_score}},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(P + 256 * 1024), "d2h_bytes_per_step": 8,
                    "pageable_frame": {"value": evals_per_step * args.steps / (ms_e2e_pg * 1e-3), "unit": UNIT,
                                       "ms_per_step": ms_e2e_pg / args.steps,
                                       "note": "same step, frame in pageable host memory: memcpy into a pinned staging "
                                               "buffer + H2D, both inside the timed region"}},
            "gpu_launches": int(round(float(np.mean(launches)) * args.steps)) * 3,  # device-timed loop + the two e2e loops
            "roofline": roofline_block(per_rank_pairs, P, hist_ms, mean_stage, hist_bytes, search_bytes, peak, peak_src,
                                       clocks),
            "stage_ms": mean_stage,
            "clocks": clocks,
        }
        if configs:
            line["configs"] = configs
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
    ap.add_argument("--no-configs", action="store_true", help="skip the C1 / C3 / C4 / C5 block")
    ap.add_argument("--frames", type=int, default=200, help="frames of the C5 sequence")
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
