"""Small driver for ncu: two C2-sized searches with a given histogram variant."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from orbslam2_nmi_b200 import synth  # noqa: E402
from orbslam2_nmi_b200.search import NmiSearcher  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
nsearch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sc = synth.make_scene("C2")
s = NmiSearcher(0)
s.set_scene(sc)
s.set_frame(synth.frame_textured(sc.W, sc.H))
g = synth.default_grid((4, 4, 4), (4, 4, 4))
for _ in range(nsearch):
    r = s.search(sc.Twc, g, s.flags(variant=variant))
t, n = s.timings()
print("variant", variant, "winner", r.best_index, r.best_score, "stage ms", t, "launches", n)
