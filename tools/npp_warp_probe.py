"""GPU probe: NPP's nppiWarpPerspective_8u_C1R vs warp_kernel on the 64 C2 homographies.
Writes gpurun_out/npp_warp_report.json and a few NPP outputs (npz) for offline analysis."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import npp_py, oracle_py as oracle  # noqa: E402  (checkers)
from orbslam2_nmi_b200 import synth  # noqa: E402
from orbslam2_nmi_b200.search import NmiSearcher  # noqa: E402

out = ROOT / "gpurun_out"
out.mkdir(exist_ok=True)
sc = synth.make_scene("C2")
g = synth.default_grid((1, 1, 1), (4, 4, 4))
s = NmiSearcher(0)
s.set_camera(sc.W, sc.H, sc.fx, sc.fy, sc.cx, sc.cy, sc.zn, sc.zf, sc.point_size)
rep = {"npp_version": npp_py.version(), "image": [sc.W, sc.H], "frames": {}}
keep = {}
for fname, fn in (("textured", synth.frame_textured), ("smooth", synth.frame_smooth)):
    frame = fn(sc.W, sc.H)
    s.set_frame(frame)
    s.warp_cells(g)
    hist = np.zeros(511, dtype=np.int64)
    per_cell = []
    for w in range(64):
        wx, wy, wz = w % 4, (w // 4) % 4, w // 16
        M = oracle.cell_homography(sc, g, wx, wy, wz)
        ref = npp_py.warp_perspective(frame, M)
        ours = s.get_warp(w)
        d = ours.astype(np.int32) - ref.astype(np.int32)
        hist += np.bincount((d + 255).ravel(), minlength=511)
        per_cell.append({"cell": [wx, wy, wz], "exact": float((d == 0).mean()), "max_abs": int(np.abs(d).max()),
                         "within1": float((np.abs(d) <= 1).mean()),
                         "border_mismatch": int(((ours == 0) != (ref == 0)).sum())})
        if fname == "textured" and w in (0, 21, 42, 63):
            keep[f"npp_{w}"] = ref
            keep[f"M_{w}"] = M
    tot = hist.sum()
    rep["frames"][fname] = {
        "exact_fraction": float(hist[255] / tot), "within_1": float(hist[254:257].sum() / tot),
        "within_2": float(hist[253:258].sum() / tot), "max_abs_diff": int(max(abs(i - 255) for i in np.nonzero(hist)[0])),
        "diff_histogram": {str(i - 255): int(hist[i]) for i in np.nonzero(hist)[0] if abs(i - 255) <= 8},
        "worst_cell": min(per_cell, key=lambda c: c["exact"]), "best_cell": max(per_cell, key=lambda c: c["exact"]),
        "cells": per_cell}
    print(fname, {k: v for k, v in rep["frames"][fname].items() if k != "cells"})
(out / "npp_warp_report.json").write_text(json.dumps(rep, indent=1))
np.savez_compressed(out / "npp_warp_samples.npz", **keep)
