"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
every kernel family on tiny scenes -- point-cloud search (first search = two-pass binning, second =
single-pass bins), the histogram builds 0 (persistent, fast epilogue), 9 (one CTA per pair, fast
epilogue), 10 (pixel ring staged through tensor memory), 1 / 2 / 5 (LDG, two-pass u32, no swizzle),
BG-off, 64 bins, hot-bin skipping modes 1 and 2 on a flat frame, the parity dump of the persistent
build, nmi_score_pairs, the single-evaluation drop-in, the level driver, the sharded driver with a
forced record-bin overflow (the retry path), a flat and a textured mesh.

    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_run.py [quick]
`quick` (for racecheck, which slows shared-memory atomics by orders of magnitude) keeps the images
at 96 x 64 and the grids at 2 x 2.
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from orbslam2_nmi_b200 import capi, synth  # noqa: E402
from orbslam2_nmi_b200.capi import Grid  # noqa: E402
from orbslam2_nmi_b200.search import NmiSearcher  # noqa: E402

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
s = NmiSearcher(0)
sc = synth.make_scene("tiny", n_points=6000 if quick else 20000)
if quick:
    sc.W, sc.H = 96, 64
    sc.cx, sc.cy = 48.0, 32.0
s.set_scene(sc)
s.set_frame(synth.frame_textured(sc.W, sc.H, seed=3))
g = synth.default_grid((2, 1, 1), (2, 1, 1)) if quick else synth.default_grid((3, 2, 2), (2, 3, 1))
variants = (0, 0, 9, 10) if quick else (0, 0, 9, 10, 1, 2, 5, 8)
for v in variants:
    r = s.search(sc.Twc, g, s.flags(variant=v))
for flags in (s.flags(bg=0), s.flags(bins=64)):
    r = s.search(sc.Twc, g, flags)
print("points", r.best_index, r.best_score)
# hot-bin skipping on flat content (side tables), then the parity dumps of both builds
s.set_frame(synth.frame_sky(sc.W, sc.H))
for mode in (2, 1, 1, 0):
    s.set_hist_skip(mode)
    r = s.search(sc.Twc, g)
for path in (1, 2, 0):
    J, HA, HB, sc0 = s.get_hist(0, 0, path=path)
    assert int(J.sum()) == sc.W * sc.H
s.set_hist_skip(1)
print("skip / dumps ok", r.best_score)
# single-evaluation drop-in (CUDAF::NMIWithCuda_noMask) and nmi_score_pairs
s.warp_cells(g)
h = s.render_cell(sc.Twc, g, 0, 0, 0)
print("eval_pair", s.eval_pair(s.warp_ptr(g, 0, 0, 0), h))
import torch  # noqa: E402

stack = torch.randint(0, 256, (3, sc.H, sc.W), dtype=torch.uint8, device="cuda")
print("score_pairs", s.score_pairs(stack.data_ptr(), 2, sc.W * sc.H, stack[1:].data_ptr(), 2, sc.W * sc.H).ravel()[:2])
out = s.relocalize(sc.Twc, synth.default_grid() if not quick else g, threshold=0.0, max_iterations=2)
print("relocalize", out.iterations, out.nmi)
# the sharded driver as rank 0 of 1, with the single-pass bin capacity forced far too small once:
# the enqueued search publishes NMI_KEY_RETRY and the level is redone with exact sizing
key = torch.zeros(1, dtype=torch.int64, device="cuda")
s.search(sc.Twc, g)
s.search(sc.Twc, g)
sc_big = synth.make_scene("tiny", n_points=60000 if not quick else 20000)
sc_big.W, sc_big.H, sc_big.cx, sc_big.cy = sc.W, sc.H, sc.cx, sc.cy
s.set_points(sc_big.xyzi[:, :4])   # a denser model under the old bin-capacity feedback is NOT kept (new model resets it)
out = s.relocalize_sharded(sc.Twc, g, None, 0, 1, key.data_ptr(), lambda k, st: None, threshold=0.0, max_iterations=2)
print("relocalize_sharded", out.iterations, out.nmi)
verts, tris = synth.make_mesh(24 if quick else 40, 24 if quick else 40, extent=24.0)
gm = Grid.make((2, 1, 1), (2, 1, 1), (0.2, 0.2, 0.5), (0.02, 0.02, 0.05))
s.set_mesh(verts, tris)
r = s.search(synth.prior_pose(), gm, s.flags(bins=64))
print("mesh", r.best_index, r.best_score)
s.set_mesh_textured(verts, tris, synth.make_mesh_uv(verts, tris, extent=24.0), synth.make_texture(32, 16))
r = s.search(synth.prior_pose(), gm)
print("textured mesh", r.best_index, r.best_score)
s.close()
print("SANITIZE RUN OK")
