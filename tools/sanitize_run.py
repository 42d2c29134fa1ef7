"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck): point-cloud search
(first search = two-pass binning, second = single-pass bins), BG-off + 64-bin flags, a mesh
search and the level driver, all on tiny scenes.
    compute-sanitizer --tool memcheck python tools/sanitize_run.py
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from orbslam2_nmi_b200 import synth  # noqa: E402
from orbslam2_nmi_b200.capi import Grid  # noqa: E402
from orbslam2_nmi_b200.search import NmiSearcher  # noqa: E402

s = NmiSearcher(0)
sc = synth.make_scene("tiny", n_points=20000)
s.set_scene(sc)
s.set_frame(synth.frame_textured(sc.W, sc.H, seed=3))
g = synth.default_grid((3, 2, 2), (2, 3, 1))
for flags in (s.flags(), s.flags(), s.flags(bg=0), s.flags(bins=64), s.flags(variant=1), s.flags(variant=2)):
    r = s.search(sc.Twc, g, flags)
print("points", r.best_index, r.best_score)
out = s.relocalize(sc.Twc, synth.default_grid(), threshold=0.0, max_iterations=2)
print("relocalize", out.iterations, out.nmi)
verts, tris = synth.make_mesh(40, 40)
s.set_mesh(verts, tris)
r = s.search(synth.prior_pose(), Grid.make((2, 2, 1), (2, 1, 1), (0.2, 0.2, 0.5), (0.02, 0.02, 0.05)), s.flags(bins=64))
print("mesh", r.best_index, r.best_score)
s.close()
print("SANITIZE RUN OK")
