"""Scores of one histogram-kernel variant against variant 0 (bit for bit) on a few scenes, with the stage time.
    python tools/exp_variant_check.py 10 tiny C1 C2"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from orbslam2_nmi_b200 import synth  # noqa: E402
from orbslam2_nmi_b200.capi import Grid  # noqa: E402
from orbslam2_nmi_b200.search import NmiSearcher  # noqa: E402

v = int(sys.argv[1])
s = NmiSearcher(0)
s.set_hist_skip(0)
for name in sys.argv[2:] or ["tiny"]:
    sc = synth.make_scene(name)
    g = Grid.make((4, 4, 4), (4, 4, 4), (0.2, 0.2, 0.5), (0.02, 0.02, 0.05)) if name != "tiny" else \
        Grid.make((3, 2, 2), (2, 3, 1), (0.2, 0.2, 0.5), (0.02, 0.02, 0.05))
    s.set_scene(sc)
    s.set_frame(synth.frame_textured(sc.W, sc.H))
    out = {}
    for var in (0, v):
        fl = s.flags(variant=var)
        for _ in range(3):
            res = s.search(sc.Twc, g, fl, want_scores=True)
        ts = []
        for _ in range(5):
            res = s.search(sc.Twc, g, fl, want_scores=True)
            ts.append(s.timings()[0]["hist_score"])
        out[var] = (res, float(np.mean(ts)))
        gJ, gHA, gHB, gs = s.get_hist(1, 2, fl)
        out[var] += ((gJ, gHA, gHB, gs),)
    a, b = out[0], out[v]
    same = np.array_equal(a[0].scores.view(np.uint32), b[0].scores.view(np.uint32))
    hsame = all(np.array_equal(x, y) for x, y in zip(a[2][:3], b[2][:3])) and a[2][3] == b[2][3]
    print(f"{name} {sc.W}x{sc.H} poses {g.n_pose}: scores identical {same}, winner {a[0].best_index}/{b[0].best_index}, "
          f"hist dump identical {hsame}, hist ms variant 0 {a[1]:.4f} variant {v} {b[1]:.4f}", flush=True)
    if not same:
        d = np.flatnonzero(a[0].scores.view(np.uint32) != b[0].scores.view(np.uint32))
        print("  differing:", d.size, "first", d[:5], a[0].scores[d[:5]], b[0].scores[d[:5]])
s.close()
