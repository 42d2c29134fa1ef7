"""A few 4096-pose searches on a small scene (for ncu: the per-evaluation fixed cost of the histogram kernel)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from orbslam2_nmi_b200 import synth  # noqa: E402
from orbslam2_nmi_b200.capi import Grid  # noqa: E402
from orbslam2_nmi_b200.search import NmiSearcher  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
name = sys.argv[2] if len(sys.argv) > 2 else "tiny"
s = NmiSearcher(0)
s.set_hist_skip(0)  # the plain build: small scenes are mostly background and would take the side-table build
g = Grid.make((4, 4, 4), (4, 4, 4), (0.2, 0.2, 0.5), (0.02, 0.02, 0.05))
sc = synth.make_scene(name)
s.set_scene(sc)
s.set_frame(synth.frame_textured(sc.W, sc.H))
for _ in range(4):
    s.search(sc.Twc, g, s.flags(variant=variant))
print(s.timings())
s.close()
