"""Cost of a bin-overflow retry at C2: a search at the usual pose, then one from a pose that fills the bins differently."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from orbslam2_nmi_b200 import synth
from orbslam2_nmi_b200.search import NmiSearcher

sc = synth.make_scene("C2")
s = NmiSearcher(0)
s.set_scene(sc)
s.set_frame(synth.frame_textured(sc.W, sc.H))
g = synth.default_grid((4, 4, 4), (4, 4, 4))
def timed(T):
    s.sync(); t0 = time.perf_counter(); r = s.search(T, g); return (time.perf_counter() - t0) * 1e3, r
for _ in range(4):
    ms, r = timed(sc.Twc)
print("steady %.2f ms" % ms, r.best_index)
far = sc.Twc.copy(); far[:3, 3] += np.array([0.0, 0.0, -30.0], dtype=np.float32)  # step back: the cloud shrinks on screen, bins fill up
for k in range(3):
    ms, r = timed(far)
    print("moved pose, search %d: %.2f ms" % (k, ms), r.best_index, "%.3f device" % r.gpu_ms)
for k in range(2):
    ms, r = timed(sc.Twc)
    print("back, search %d: %.2f ms" % (k, ms), r.best_index)
