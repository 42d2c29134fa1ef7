"""Seeded synthetic inputs of SURVEY.md section 8(d): cloud, prior pose, frames, grids.

There is no dataset in the container (the reference's ZU-MAV / Newer College
models are download links, README.md:12-16), so every test and the benchmark use
these generators.  numpy only; deterministic for a given seed.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from .capi import Grid

# BASELINE.json configs (intrinsics: C1 EuRoC-like, C2 ETH.yaml:8-11, C3 NewerCollege_short.yaml)
CONFIGS = {
    "C1": dict(W=752, H=480, fx=458.0, fy=458.0, cx=376.0, cy=240.0, n_points=1_000_000),
    "C2": dict(W=1920, H=1080, fx=870.0918641, fy=870.0918641, cx=951.1156353, cy=549.4975458,
               n_points=10_000_000),
    "C3": dict(W=848, H=480, fx=431.3873, fy=430.2496, cx=427.4407, cy=238.5269, n_points=2_000_000),
    # small case the CPU oracle finishes in seconds
    "tiny": dict(W=160, H=96, fx=95.0, fy=95.0, cx=80.0, cy=48.0, n_points=60_000),
    "small": dict(W=320, H=200, fx=190.0, fy=190.0, cx=158.0, cy=101.0, n_points=250_000),
}
ZN, ZF, POINT_SIZE = 5.0, 30.0, 3.0  # ETH_small.yaml:90-92


@dataclass
class Scene:
    W: int
    H: int
    fx: float
    fy: float
    cx: float
    cy: float
    zn: float
    zf: float
    point_size: float
    xyzi: np.ndarray  # (N, 4) float32: x, y, z, I = v/256
    Twc: np.ndarray   # (4, 4) float32 prior pose, camera -> world


def height(x, y):
    return 2.0 * np.sin(0.13 * x) * np.cos(0.11 * y)


def make_cloud(n: int, seed: int = 1234, extent: float = 40.0) -> np.ndarray:
    """x,y ~ U(-extent, extent); z = heightfield + noise; v = textured grey in [0, 254]."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-extent, extent, n)
    y = rng.uniform(-extent, extent, n)
    z = height(x, y) + 0.2 * rng.standard_normal(n)
    v = 127.0 + 60.0 * np.sin(0.5 * x) + 50.0 * np.cos(0.37 * y) + 15.0 * rng.standard_normal(n)
    v = np.clip(np.rint(v), 0, 254)
    out = np.empty((n, 4), dtype=np.float32)
    out[:, 0], out[:, 1], out[:, 2] = x, y, z
    out[:, 3] = (v / 256.0).astype(np.float32)  # objloader.cpp:261: colour * (1/256)
    return out


def make_mesh(nx: int = 1000, ny: int = 1000, seed: int = 1234, extent: float = 40.0):
    """Regular (nx+1) x (ny+1) vertex grid over the same height field -> 2*nx*ny triangles
    (1000 x 1000 -> 2 M triangles), per-vertex grey in [0, 254]/256.
    Winding is chosen so the triangles face a camera above the terrain looking down."""
    rng = np.random.default_rng(seed)
    xs = np.linspace(-extent, extent, nx + 1)
    ys = np.linspace(-extent, extent, ny + 1)
    X, Y = np.meshgrid(xs, ys)  # (ny+1, nx+1)
    Z = height(X, Y) + 0.05 * rng.standard_normal(X.shape)
    V = 127.0 + 60.0 * np.sin(0.5 * X) + 50.0 * np.cos(0.37 * Y) + 15.0 * rng.standard_normal(X.shape)
    V = np.clip(np.rint(V), 0, 254)
    verts = np.stack([X, Y, Z, V / 256.0], axis=-1).reshape(-1, 4).astype(np.float32)
    j, i = np.mgrid[0:ny, 0:nx]
    v00 = (j * (nx + 1) + i).reshape(-1)
    v10 = v00 + 1
    v01 = v00 + (nx + 1)
    v11 = v01 + 1
    t1 = np.stack([v00, v10, v11], axis=-1)
    t2 = np.stack([v00, v11, v01], axis=-1)
    tris = np.concatenate([t1, t2], axis=0)
    tris = np.stack([t1, t2], axis=1).reshape(-1, 3).astype(np.uint32)
    return verts, tris


def make_mesh_uv(verts: np.ndarray, tris: np.ndarray, extent: float = 40.0, repeats: float = 3.0) -> np.ndarray:
    """Corner UVs (nt, 3, 2) for make_mesh's terrain, un-indexed like loadOBJ's out_uvs
    (objloader.cpp:206-220): the texture is laid over the terrain `repeats` times per axis, so the
    UVs leave [0, 1] and GL_REPEAT (texture.cpp:100-101) is exercised."""
    xy = verts[:, :2].astype(np.float64)
    uv = (xy + extent) / (2.0 * extent) * repeats - 0.25
    return uv[tris.astype(np.int64)].astype(np.float32)


def make_texture(tw: int = 512, th: int = 512, seed: int = 77) -> np.ndarray:
    """(th, tw, 3) u8 texture in BMP payload order (row 0 = bottom row, bytes B, G, R): smooth colour
    fields + texel noise, the three channels different so that the channel swap of
    loadBMP_custom (texture.cpp:90) matters."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:th, 0:tw].astype(np.float64)
    c0 = 128 + 90 * np.sin(xx / 23.0) * np.cos(yy / 31.0)
    c1 = 128 + 80 * np.cos((xx + 2 * yy) / 47.0)
    c2 = 128 + 100 * np.sin((xx - yy) / 19.0)
    t = np.stack([c0, c1, c2], axis=-1) + 10.0 * rng.standard_normal((th, tw, 3))
    return np.clip(np.rint(t), 0, 255).astype(np.uint8)


def prior_pose(height_above: float = 15.0, tilt_deg: float = 10.0) -> np.ndarray:
    """Camera `height_above` m over the terrain (world z up), looking down, tilted about x."""
    # CV camera: x right, y down, z forward.  Looking straight down: z_cam = -z_world.
    R0 = np.array([[1, 0, 0], [0, -1, 0], [0, 0, -1]], dtype=np.float64)  # columns = cam axes in world
    a = math.radians(tilt_deg)
    Rx = np.array([[1, 0, 0], [0, math.cos(a), -math.sin(a)], [0, math.sin(a), math.cos(a)]])
    R = R0 @ Rx
    T = np.eye(4, dtype=np.float64)
    T[:3, :3] = R
    T[:3, 3] = [0.0, 0.0, height_above + float(height(0.0, 0.0))]
    return T.astype(np.float32)


def make_scene(config: str = "tiny", n_points: int | None = None, seed: int = 1234,
               extent: float | None = None) -> Scene:
    c = dict(CONFIGS[config])
    n = n_points if n_points is not None else c["n_points"]
    if extent is None:
        extent = 40.0 if config in ("C1", "C2", "C3") else 24.0
    return Scene(W=c["W"], H=c["H"], fx=c["fx"], fy=c["fy"], cx=c["cx"], cy=c["cy"], zn=ZN, zf=ZF,
                 point_size=POINT_SIZE, xyzi=make_cloud(n, seed, extent), Twc=prior_pose())


def frame_uniform(W: int, H: int, seed: int = 99) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 256, size=(H, W), dtype=np.uint8)


def frame_constant(W: int, H: int, value: int = 128) -> np.ndarray:
    return np.full((H, W), value, dtype=np.uint8)


def frame_from_render(render: np.ndarray, seed: int = 7, gamma: float = 0.7,
                      noise: float = 4.0) -> np.ndarray:
    """'planted' frame: monotone LUT of a render + Gaussian noise."""
    rng = np.random.default_rng(seed)
    v = 255.0 * (render.astype(np.float64) / 255.0) ** gamma + noise * rng.standard_normal(render.shape)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def frame_textured(W: int, H: int, seed: int = 5) -> np.ndarray:
    """Smooth camera-like image (low-frequency texture + mild noise)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    v = (128 + 50 * np.sin(xx / 37.0) * np.cos(yy / 29.0) + 40 * np.sin((xx + yy) / 91.0)
         + 6 * rng.standard_normal((H, W)))
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def frame_sky(W: int, H: int, seed: int = 5, sky: float = 0.35) -> np.ndarray:
    """Textured frame whose top `sky` fraction is saturated (255), like an over-exposed sky:
    together with the render background (255) it piles a large share of all pixels onto one
    joint-histogram bin -- the contention case real outdoor frames produce."""
    f = frame_textured(W, H, seed)
    f[: int(H * sky)] = 255
    return f


def frame_smooth(W: int, H: int, seed: int = 5) -> np.ndarray:
    """Noise-free low-frequency texture: long runs of equal grey levels."""
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    v = 128 + 50 * np.sin(xx / 37.0) * np.cos(yy / 29.0) + 40 * np.sin((xx + yy) / 91.0)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def default_grid(n_synth=(3, 3, 3), n_warp=(3, 3, 3)) -> Grid:
    """ETH_small.yaml:77-88 steps: 0.2/0.2/0.5 m, 0.02/0.02/0.05 rad."""
    return Grid.make(n_synth, n_warp, (0.2, 0.2, 0.5), (0.02, 0.02, 0.05))
