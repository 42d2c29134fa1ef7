"""Writes tests/golden/*.json from the CPU oracle (run once, commit the output).

The reference has no golden vectors of its own and cannot run here, so
these fixtures pin the ORACLE against accidental change; the oracle itself is pinned
by the hand-computed cases in tests/test_oracle_kat.py.
    python tests/golden/make_golden.py
"""
import json
import sys
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import oracle_py as oracle  # noqa: E402
from orbslam2_nmi_b200 import synth  # noqa: E402
from orbslam2_nmi_b200.capi import Grid  # noqa: E402


def main():
    cfg = dict(config="tiny", n_points=20000, seed=4321, frame_seed=17,
               nS=[2, 1, 2], nW=[1, 2, 1], stepT=[0.25, 0.2, 0.5], stepR=[0.02, 0.03, 0.05])
    sc = synth.make_scene(cfg["config"], n_points=cfg["n_points"], seed=cfg["seed"])
    g = Grid.make(cfg["nS"], cfg["nW"], cfg["stepT"], cfg["stepR"])
    frame = synth.frame_textured(sc.W, sc.H, seed=cfg["frame_seed"])
    scores, renders, warps = oracle.search_points(sc, sc.Twc, g, sc.xyzi, frame, keep_images=True)
    J, HA, HB = oracle.joint_hist(renders[0], warps[0])
    cfg.update(
        render_crc32=[zlib.crc32(r.tobytes()) for r in renders],
        warp_crc32=[zlib.crc32(w.tobytes()) for w in warps],
        joint_crc32_pair00=zlib.crc32(J.tobytes()),
        scores=[float(s) for s in scores],
        argmax=oracle.argmax(scores)[0],
    )
    out = Path(__file__).parent / "oracle_small.json"
    out.write_text(json.dumps(cfg, indent=1))
    print("wrote", out, "scores", np.round(scores, 5))


if __name__ == "__main__":
    main()
