"""C1 single evaluation: the launches of one nmi_search (1 pose) and of nmi_eval_pair calls, for ncu's launch list."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from orbslam2_nmi_b200 import synth  # noqa: E402
from orbslam2_nmi_b200.search import NmiSearcher  # noqa: E402

sc = synth.make_scene("C1")
s = NmiSearcher(0)
s.set_scene(sc)
s.set_frame(synth.frame_textured(sc.W, sc.H))
g = synth.default_grid((1, 1, 1), (1, 1, 1))
for _ in range(4):
    r = s.search(sc.Twc, g)
s.warp_cells(g)
h = s.render_cell(sc.Twc, g, 0, 0, 0)
p = s.warp_ptr(g, 0, 0, 0)
for _ in range(4):
    v = s.eval_pair(p, h)
print("score", r.best_score, v)
