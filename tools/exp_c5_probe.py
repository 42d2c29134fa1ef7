"""C5 after C2_frames and C4, as bench.py orders them: which frames of the sequence are slow (buffer regrowth, retries)."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np, torch
import bench
from orbslam2_nmi_b200 import synth
from orbslam2_nmi_b200.search import NmiSearcher
sc = synth.make_scene("C2")
s = NmiSearcher(0); s.set_scene(sc)
key = torch.zeros(1, dtype=torch.int64, device="cuda")
grid = synth.default_grid((4, 4, 4), (4, 4, 4))
s.set_frame(synth.frame_textured(sc.W, sc.H))
bench.config_c2_frames(s, sc, grid, s.flags())
bench.config_c4(s, sc, key, 0, 1, None)
r = bench.config_c5(s, sc, key, 0, 1, None, 200)
print({k: r[k] for k in ("ms_per_frame_mean", "ms_per_frame_median", "ms_per_frame_p99", "ms_per_frame_max", "slowest_frames")})
