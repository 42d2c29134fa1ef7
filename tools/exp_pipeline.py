"""Experiment: two contexts on ONE GPU, searches alternating between them, so that the render /
warp stage of search k+1 runs under the histogram kernel of search k (throughput mode for a
frame sequence whose priors do not depend on the previous search).  Prints ms per search for
depth 1 (one context, what bench.py times) and depth 2.

    python tools/exp_pipeline.py [steps]
"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from orbslam2_nmi_b200 import synth  # noqa: E402
from orbslam2_nmi_b200.search import NmiSearcher  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
scene = synth.make_scene("C2")
frame = synth.frame_textured(scene.W, scene.H)
grid = synth.default_grid((4, 4, 4), (4, 4, 4))
ctx = [NmiSearcher(0), NmiSearcher(0)]
keys = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in ctx]
for s in ctx:
    s.set_scene(scene)
    s.set_frame(frame)
flags = ctx[0].flags()


def run(depth, n):
    for s in ctx[:depth]:
        s.sync()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(n):
        i = k % depth
        if k >= depth:
            ctx[i].sync()  # the search this context ran `depth` steps ago
        ctx[i].search_enqueue(scene.Twc, grid, flags, 0, 1, keys[i].data_ptr())
    for s in ctx[:depth]:
        s.sync()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n


for depth in (1, 2):
    run(depth, 6)
    ms = run(depth, steps)
    print(f"depth {depth}: {ms:.3f} ms per search, {4096 / ms * 1e3:.0f} evals/s, keys {[int(k.item()) for k in keys[:depth]]}")
