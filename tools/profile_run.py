"""Small driver for ncu: C2-sized searches with a given histogram variant.  Only the LAST search sits
between cudaProfilerStart/Stop, so `ncu --profile-from-start off` captures one steady-state search
(the first search sizes its tile bins conservatively and launches more kernels)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from orbslam2_nmi_b200 import synth  # noqa: E402
from orbslam2_nmi_b200.search import NmiSearcher  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
nsearch = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sc = synth.make_scene("C2")
s = NmiSearcher(0)
s.set_scene(sc)
s.set_frame(synth.frame_textured(sc.W, sc.H))
g = synth.default_grid((4, 4, 4), (4, 4, 4))
import torch  # noqa: E402

rt = torch.cuda.cudart()
for i in range(nsearch):
    if i == nsearch - 1:
        s.sync()
        rt.cudaProfilerStart()
    r = s.search(sc.Twc, g, s.flags(variant=variant))
s.sync()
rt.cudaProfilerStop()
t, n = s.timings()
print("variant", variant, "winner", r.best_index, r.best_score, "stage ms", t, "launches", n)
