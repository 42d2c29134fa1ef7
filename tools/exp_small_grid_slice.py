"""Stage times of one rank's slice of a 3^3 x 3^3 level at world = 1 / 2 / 4 / 8 (run on one GPU: rank 0's slice)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from orbslam2_nmi_b200 import synth
from orbslam2_nmi_b200.capi import Grid
from orbslam2_nmi_b200.search import NmiSearcher

sc = synth.make_scene("C2")
s = NmiSearcher(0)
s.set_scene(sc)
s.set_frame(synth.frame_textured(sc.W, sc.H))
g = Grid.make((3, 3, 3), (3, 3, 3), (0.2, 0.2, 0.5), (0.02, 0.02, 0.05))
key = torch.zeros(1, dtype=torch.int64, device="cuda")
fl = s.flags()
for world in (1, 2, 4, 8):
    acc = []
    for it in range(8):
        s.search_enqueue(sc.Twc, g, fl, 0, world, key.data_ptr())
        s.sync()
        t, n = s.timings()
        if it >= 3:
            acc.append(t)
    m = {k: round(float(np.mean([a[k] for a in acc])), 4) for k in acc[0]}
    print("world", world, "launches", n, m)
