"""C3-sized searches over the textured mesh (ncu driver: only the LAST search sits between
cudaProfilerStart/Stop).  argv[1] = number of searches (default 4)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from orbslam2_nmi_b200 import synth  # noqa: E402
from orbslam2_nmi_b200.capi import Grid  # noqa: E402
from orbslam2_nmi_b200.search import NmiSearcher  # noqa: E402

nsearch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
c = synth.CONFIGS["C3"]
verts, tris = synth.make_mesh(1000, 1000)
uv = synth.make_mesh_uv(verts, tris, repeats=6.0)
tex = synth.make_texture(1024, 1024)
s = NmiSearcher(0)
s.set_camera(c["W"], c["H"], c["fx"], c["fy"], c["cx"], c["cy"], synth.ZN, synth.ZF, 3.0)
s.set_mesh_textured(verts, tris, uv, tex)
s.set_frame(synth.frame_textured(c["W"], c["H"]))
g = Grid.make((4, 4, 4), (4, 4, 1), (0.2, 0.2, 0.5), (0.02, 0.02, 0.05))
fl = s.flags(bins=64)
Twc = synth.prior_pose()
rt = torch.cuda.cudart()
ms = []
for i in range(nsearch):
    if i == nsearch - 1:
        s.sync()
        rt.cudaProfilerStart()
    r = s.search(Twc, g, fl)
    ms.append(r.gpu_ms)
s.sync()
rt.cudaProfilerStop()
t, n = s.timings()
print("C3 winner", r.best_index, r.best_score, "ms", [round(x, 3) for x in ms], "stage ms", {k: round(v, 4) for k, v in t.items()}, "launches", n)
print("render0 crc", int(np.bitwise_xor.reduce(s.get_render(0).astype(np.uint32).ravel() * np.arange(1, c["W"] * c["H"] + 1, dtype=np.uint32))))
