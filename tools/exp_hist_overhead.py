"""Per-CTA fixed cost of joint_hist_score_kernel: a 4096-pose search on images so small that the
pixel loop is a handful of chunks -- what remains is prologue (zero 128 KiB, term table, barriers)
+ epilogue (rows, entropy terms, trees) x 4096 CTAs.  Prints stage times for a few image sizes."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from orbslam2_nmi_b200 import synth  # noqa: E402
from orbslam2_nmi_b200.capi import Grid  # noqa: E402
from orbslam2_nmi_b200.search import NmiSearcher  # noqa: E402

s = NmiSearcher(0)
s.set_hist_skip(0)  # the plain build: small scenes are mostly background and would take the side-table build
FL = s.flags(variant=int(os.environ.get('NMI_EXP_VARIANT', '0')))
g = Grid.make((4, 4, 4), (4, 4, 4), (0.2, 0.2, 0.5), (0.02, 0.02, 0.05))
for name in sys.argv[1:] or ["tiny", "small"]:
    sc = synth.make_scene(name)
    s.set_scene(sc)
    s.set_frame(synth.frame_textured(sc.W, sc.H))
    for _ in range(3):
        s.search(sc.Twc, g, FL)
    ts = []
    for _ in range(5):
        s.search(sc.Twc, g, FL)
        ts.append(s.timings()[0])
    m = {k: round(float(np.mean([t[k] for t in ts])), 4) for k in ts[0]}
    print("variant", FL.variant, name, f"{sc.W}x{sc.H}", "pixels", sc.W * sc.H, "chunks", -(-sc.W * sc.H // 8192), m, flush=True)
s.close()
