"""Runs BASELINE.json's five configurations on one B200 and prints one JSON line each.

C2 is what bench.py measures; the others are parity-test shapes that are
timed here for the record (profiles/rNN_configs.jsonl).  Synthetic data only.
    python tools/run_configs.py [--frames 40]
"""
import argparse
import json
import math
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from orbslam2_nmi_b200 import capi, synth  # noqa: E402
from orbslam2_nmi_b200.capi import Grid  # noqa: E402
from orbslam2_nmi_b200.search import NmiSearcher  # noqa: E402


def timed_search(s, Twc, g, flags, reps=5):
    s.search(Twc, g, flags)
    s.search(Twc, g, flags)
    ms = []
    for _ in range(reps):
        r = s.search(Twc, g, flags)
        ms.append(r.gpu_ms)
    t, n = s.timings()
    return r, float(np.median(ms)), t, n


def emit(**kw):
    print(json.dumps(kw), flush=True)


def c1(s):
    sc = synth.make_scene("C1")
    s.set_scene(sc)
    s.set_frame(synth.frame_textured(sc.W, sc.H))
    g = synth.default_grid((1, 1, 1), (1, 1, 1))
    r, ms, t, n = timed_search(s, sc.Twc, g, s.flags())
    # NMIWithCuda_noMask alone (the reference's per-evaluation entry point), wall clock incl. D2H
    s.warp_cells(g)
    h = s.render_cell(sc.Twc, g, 0, 0, 0)
    p = s.warp_ptr(g, 0, 0, 0)
    s.eval_pair(p, h)
    t0 = time.perf_counter()
    for _ in range(50):
        s.eval_pair(p, h)
    per_eval = (time.perf_counter() - t0) / 50 * 1e3
    emit(config="C1 single evaluation 752x480 vs 1M-point render, 256 bins", search_ms=ms, stage_ms=t,
         launches=n, score=r.best_score, eval_pair_call_ms=per_eval, evals_per_s_per_call=1e3 / per_eval)


def c3(s):
    c = synth.CONFIGS["C3"]
    verts, tris = synth.make_mesh(1000, 1000)
    s.set_camera(c["W"], c["H"], c["fx"], c["fy"], c["cx"], c["cy"], synth.ZN, synth.ZF, 3.0)
    s.set_mesh(verts, tris)
    s.set_frame(synth.frame_textured(c["W"], c["H"]))
    g = Grid.make((4, 4, 4), (4, 4, 1), (0.2, 0.2, 0.5), (0.02, 0.02, 0.05))
    r, ms, t, n = timed_search(s, synth.prior_pose(), g, s.flags(bins=64))
    emit(config="C3 Newer-College-shaped: 848x480 frame vs 2M-triangle mesh, 1024 poses, 64 bins",
         search_ms=ms, evals_per_s=g.n_pose / ms * 1e3, stage_ms=t, launches=n, winner=r.best_index)


def c4(s, sc):
    s.set_scene(sc)
    s.set_frame(synth.frame_textured(sc.W, sc.H))
    g = Grid.make((8, 8, 8), (4, 4, 4), (0.2, 0.2, 0.5), (0.02, 0.02, 0.05))
    s.relocalize(sc.Twc, g, threshold=0.0, max_iterations=3)
    t0 = time.perf_counter()
    out = s.relocalize(sc.Twc, g, threshold=0.0, max_iterations=3)
    wall = (time.perf_counter() - t0) * 1e3
    emit(config="C4 coarse-to-fine 3-level search, level 0 = 8^3 x 4^3 = 32768 poses (1 GPU here)",
         levels=out.iterations, evals=out.n_evals, gpu_ms=out.gpu_ms, wall_ms=wall,
         evals_per_s=out.n_evals / out.gpu_ms * 1e3, nmi=out.nmi)


def c5(s, sc, frames):
    """Sequence: frames along a smooth trajectory; prior = true pose + drift; NMI correction
    (3^6 grid, up to 4 levels) every frame.  A stub stands in for ORB-SLAM2's tracker (the
    vocabulary blob is not in the reference checkout)."""
    s.set_scene(sc)
    g0 = synth.default_grid()
    rng = np.random.default_rng(0)
    errs_before, errs_after, ms = [], [], []
    for k in range(frames):
        T = sc.Twc.copy()
        T[0, 3] += 0.05 * k
        T[1, 3] += 2.0 * math.sin(0.05 * k)
        gt = Grid.make((1, 1, 1), (1, 1, 1), (0.1,) * 3, (0.01,) * 3)
        s.render_cell(T, gt, 0, 0, 0)
        frame = synth.frame_from_render(s.get_render(0), seed=k)
        s.set_frame(frame)
        prior = T.copy()
        drift = rng.integers(-1, 2, size=3) * np.array([0.2, 0.2, 0.5], dtype=np.float32)
        ax, ay, az = -T[:3, 0], T[:3, 1], -T[:3, 2]
        prior[:3, 3] += (drift[0] * ax + drift[1] * ay + drift[2] * az).astype(np.float32)
        t0 = time.perf_counter()
        out = s.relocalize(prior, g0, threshold=0.02)
        ms.append((time.perf_counter() - t0) * 1e3)
        got = np.array(out.Twc[:]).reshape(4, 4)
        errs_before.append(float(np.linalg.norm(prior[:3, 3] - T[:3, 3])))
        errs_after.append(float(np.linalg.norm(got[:3, 3] - T[:3, 3])))
    emit(config="C5 sequence: synthetic frames, NMI pose correction every frame (stub tracker), 1 GPU",
         frames=frames, ms_per_frame_median=float(np.median(ms)), frames_per_s=1e3 / float(np.median(ms)),
         seconds_for_1000_frames=float(np.median(ms)), mean_err_before_m=float(np.mean(errs_before)),
         mean_err_after_m=float(np.mean(errs_after)),
         corrected_fraction=float(np.mean(np.array(errs_after) < 0.5 * np.maximum(np.array(errs_before), 1e-9))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=40)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    s = NmiSearcher(0)
    want = a.only.split(",") if a.only else ["C1", "C3", "C4", "C5"]
    if "C1" in want:
        c1(s)
    if "C3" in want:
        c3(s)
    if "C4" in want or "C5" in want:
        sc = synth.make_scene("C2")
        if "C4" in want:
            c4(s, sc)
        if "C5" in want:
            c5(s, sc, a.frames)
    s.close()


if __name__ == "__main__":
    main()
