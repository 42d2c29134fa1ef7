"""ctypes loader of the CPU oracle (oracle/_build/libnmi_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by the orbslam2_nmi_b200 package.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_build" / "libnmi_oracle.so"

SUC, ENMI = 1, 0
EMPTY = 0xFFFFFFFF


class OrcCamera(C.Structure):
    _fields_ = [("W", C.c_int), ("H", C.c_int), ("fx", C.c_double), ("fy", C.c_double),
                ("cx", C.c_double), ("cy", C.c_double), ("zn", C.c_double), ("zf", C.c_double),
                ("point_size", C.c_float)]


class OrcGrid(C.Structure):
    _fields_ = [("nS", C.c_int * 3), ("nW", C.c_int * 3), ("stepT", C.c_float * 3),
                ("stepR", C.c_float * 3)]


class OrcRelocParams(C.Structure):
    _fields_ = [("threshold", C.c_float), ("max_iterations", C.c_int), ("dist", C.c_float * 3),
                ("rot", C.c_float * 3)]


class OrcRelocResult(C.Structure):
    _fields_ = [("Twc", C.c_float * 16), ("relocalized", C.c_int), ("failed", C.c_int),
                ("iterations", C.c_int), ("nmi", C.c_float), ("last_nmi", C.c_float),
                ("final_grid", OrcGrid), ("best_s", C.c_int * 3), ("best_w", C.c_int * 3),
                ("n_evals", C.c_int)]


_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if not build_if_missing:
            raise RuntimeError(f"{LIB_PATH} not built")
        subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)
    lib = C.CDLL(str(LIB_PATH))
    P = C.c_void_p
    lib.orc_cell_translation.argtypes = [P, C.POINTER(OrcGrid), C.c_int, C.c_int, C.c_int, P]
    lib.orc_render_points.argtypes = [C.POINTER(OrcCamera), P, P, P, C.c_size_t, P, P]
    lib.orc_render_mesh.argtypes = [C.POINTER(OrcCamera), P, P, P, C.c_size_t, P, C.c_size_t, P, P]
    lib.orc_search_mesh.argtypes = [C.POINTER(OrcCamera), P, C.POINTER(OrcGrid), P, C.c_size_t, P, C.c_size_t,
                                    P, C.c_int, C.c_int, C.c_int, P, P, P, C.c_int]
    lib.orc_search_mesh.restype = C.c_int
    lib.orc_cell_angles.argtypes = [C.POINTER(OrcGrid), C.c_int, C.c_int, C.c_int, P]
    lib.orc_cell_homography_inv.argtypes = [C.POINTER(OrcCamera), C.POINTER(OrcGrid), C.c_int,
                                            C.c_int, C.c_int, P]
    lib.orc_warp.argtypes = [P, C.c_int, C.c_int, P, P]
    lib.orc_joint_hist.argtypes = [P, P, C.c_size_t, C.c_int, C.c_int, P, P, P]
    lib.orc_score_f32.argtypes = [P, P, P, C.c_int, C.c_uint32, C.c_int]
    lib.orc_score_f32.restype = C.c_float
    lib.orc_score_f64.argtypes = [P, P, P, C.c_int, C.c_uint32, C.c_int]
    lib.orc_score_f64.restype = C.c_double
    lib.orc_score_stages_f32.argtypes = [P, P, P, C.c_int, C.c_uint32, C.c_int, P, P, P, P, P]
    lib.orc_score_stages_f32.restype = C.c_float
    lib.orc_log2f.argtypes = [C.c_float]
    lib.orc_log2f.restype = C.c_float
    lib.orc_tree_f32.argtypes = [P, C.c_int]
    lib.orc_tree_f32.restype = C.c_float
    lib.orc_finish_f32.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int]
    lib.orc_finish_f32.restype = C.c_float
    lib.orc_eval_one.argtypes = [P, P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.orc_eval_one.restype = C.c_float
    lib.orc_argmax.argtypes = [P, C.c_size_t, P]
    lib.orc_argmax.restype = C.c_long
    lib.orc_linear_index.argtypes = [C.POINTER(OrcGrid)] + [C.c_int] * 6
    lib.orc_linear_index.restype = C.c_size_t
    lib.orc_unravel_index.argtypes = [C.POINTER(OrcGrid), C.c_size_t, P, P]
    lib.orc_search_points.argtypes = [C.POINTER(OrcCamera), P, C.POINTER(OrcGrid), P, C.c_size_t, P,
                                      C.c_int, C.c_int, C.c_int, P, P, P, C.c_int]
    lib.orc_search_points.restype = C.c_int
    lib.orc_apply_winner.argtypes = [P, C.POINTER(OrcGrid), P, P, P]
    lib.orc_is_middle.argtypes = [C.POINTER(OrcGrid), P, P]
    lib.orc_is_middle.restype = C.c_int
    lib.orc_resize_grid.argtypes = [C.POINTER(OrcGrid), P, P]
    lib.orc_grid_from_motion.argtypes = [C.POINTER(OrcGrid), P, P, C.c_int, C.POINTER(OrcGrid)]
    lib.orc_relocalize_points.argtypes = [C.POINTER(OrcCamera), P, C.POINTER(OrcGrid), P, C.c_size_t, P,
                                          C.c_int, C.c_int, C.c_int, C.POINTER(OrcRelocParams),
                                          C.POINTER(OrcRelocResult), C.c_int]
    lib.orc_relocalize_points.restype = C.c_int
    _lib = lib
    return lib


def _p(a):
    return a.ctypes.data if a is not None else None


def camera(scene_or_cam) -> OrcCamera:
    s = scene_or_cam
    return OrcCamera(int(s.W), int(s.H), s.fx, s.fy, s.cx, s.cy, s.zn, s.zf, s.point_size)


def grid(g) -> OrcGrid:
    o = OrcGrid()
    for k in range(3):
        o.nS[k], o.nW[k], o.stepT[k], o.stepR[k] = g.nS[k], g.nW[k], g.stepT[k], g.stepR[k]
    return o


def _twc(Twc):
    return np.ascontiguousarray(Twc, dtype=np.float32).reshape(16)


def cell_translation(Twc, g, sx, sy, sz):
    t = np.zeros(3, dtype=np.float32)
    T = _twc(Twc)
    load().orc_cell_translation(_p(T), C.byref(grid(g)), sx, sy, sz, _p(t))
    return t


def render_points(cam, Twc, t, xyzi):
    W, H = cam.W, cam.H
    win = np.empty((H, W), dtype=np.uint32)
    img = np.empty((H, W), dtype=np.uint8)
    T = _twc(Twc)
    t = np.ascontiguousarray(t, dtype=np.float32)
    xyzi = np.ascontiguousarray(xyzi, dtype=np.float32)
    load().orc_render_points(C.byref(camera(cam)), _p(T), _p(t), _p(xyzi), xyzi.shape[0], _p(win), _p(img))
    return win, img


def render_mesh(cam, Twc, t, verts, tris):
    W, H = cam.W, cam.H
    win = np.empty((H, W), dtype=np.uint32)
    img = np.empty((H, W), dtype=np.uint8)
    T = _twc(Twc)
    t = np.ascontiguousarray(t, dtype=np.float32)
    verts = np.ascontiguousarray(verts, dtype=np.float32)
    tris = np.ascontiguousarray(tris, dtype=np.uint32)
    load().orc_render_mesh(C.byref(camera(cam)), _p(T), _p(t), _p(verts), verts.shape[0], _p(tris),
                           tris.shape[0], _p(win), _p(img))
    return win, img


def render_mesh_tex(cam, Twc, t, verts, tris, corner_uv, tex):
    """Rendering<1> with the texture: corner_uv (nt, 3, 2) float32, tex (th, tw, 3) u8 in file byte order."""
    W, H = cam.W, cam.H
    win = np.empty((H, W), dtype=np.uint32)
    img = np.empty((H, W), dtype=np.uint8)
    T = _twc(Twc)
    t = np.ascontiguousarray(t, dtype=np.float32)
    verts = np.ascontiguousarray(verts, dtype=np.float32)
    tris = np.ascontiguousarray(tris, dtype=np.uint32)
    uv = np.ascontiguousarray(corner_uv, dtype=np.float32)
    tex = np.ascontiguousarray(tex, dtype=np.uint8)
    lib = load()
    lib.orc_render_mesh_tex.argtypes = [C.POINTER(OrcCamera), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                        C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p]
    lib.orc_render_mesh_tex(C.byref(camera(cam)), _p(T), _p(t), _p(verts), verts.shape[0], _p(tris), tris.shape[0],
                            _p(uv), _p(tex), tex.shape[1], tex.shape[0], _p(win), _p(img))
    return win, img


def search_mesh_tex(cam, Twc, g, verts, tris, corner_uv, tex, frame, bins=256, bg=True, mode=SUC,
                    keep_images=False, threads=0):
    og = grid(g)
    nS = g.nS[0] * g.nS[1] * g.nS[2]
    nW = g.nW[0] * g.nW[1] * g.nW[2]
    scores = np.zeros(nS * nW, dtype=np.float32)
    renders = np.empty((nS, cam.H, cam.W), dtype=np.uint8) if keep_images else None
    warps = np.empty((nW, cam.H, cam.W), dtype=np.uint8) if keep_images else None
    T = _twc(Twc)
    verts = np.ascontiguousarray(verts, dtype=np.float32)
    tris = np.ascontiguousarray(tris, dtype=np.uint32)
    uv = np.ascontiguousarray(corner_uv, dtype=np.float32)
    tex = np.ascontiguousarray(tex, dtype=np.uint8)
    frame = np.ascontiguousarray(frame, dtype=np.uint8)
    lib = load()
    lib.orc_search_mesh_tex.argtypes = [C.POINTER(OrcCamera), C.c_void_p, C.POINTER(OrcGrid), C.c_void_p, C.c_size_t,
                                        C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                        C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.orc_search_mesh_tex(C.byref(camera(cam)), _p(T), C.byref(og), _p(verts), verts.shape[0], _p(tris),
                            tris.shape[0], _p(uv), _p(tex), tex.shape[1], tex.shape[0], _p(frame), bins, int(bg), mode,
                            _p(scores), _p(renders), _p(warps), threads)
    return scores, renders, warps


def search_mesh(cam, Twc, g, verts, tris, frame, bins=256, bg=True, mode=SUC, keep_images=False, threads=0):
    og = grid(g)
    nS = g.nS[0] * g.nS[1] * g.nS[2]
    nW = g.nW[0] * g.nW[1] * g.nW[2]
    scores = np.zeros(nS * nW, dtype=np.float32)
    renders = np.empty((nS, cam.H, cam.W), dtype=np.uint8) if keep_images else None
    warps = np.empty((nW, cam.H, cam.W), dtype=np.uint8) if keep_images else None
    T = _twc(Twc)
    verts = np.ascontiguousarray(verts, dtype=np.float32)
    tris = np.ascontiguousarray(tris, dtype=np.uint32)
    frame = np.ascontiguousarray(frame, dtype=np.uint8)
    load().orc_search_mesh(C.byref(camera(cam)), _p(T), C.byref(og), _p(verts), verts.shape[0], _p(tris),
                           tris.shape[0], _p(frame), bins, int(bg), mode, _p(scores), _p(renders), _p(warps),
                           threads)
    return scores, renders, warps


def cell_homography(cam, g, ix, iy, iz):
    """Forward K R K^-1 (float64 3x3), the matrix image.cpp:104-106 hands to warpPerspective."""
    m = np.zeros(9, dtype=np.float64)
    lib = load()
    lib.orc_cell_homography.argtypes = [C.POINTER(OrcCamera), C.POINTER(OrcGrid), C.c_int, C.c_int, C.c_int,
                                        C.c_void_p]
    lib.orc_cell_homography(C.byref(camera(cam)), C.byref(grid(g)), ix, iy, iz, _p(m))
    return m.reshape(3, 3)


def cell_homography_inv(cam, g, ix, iy, iz):
    m = np.zeros(9, dtype=np.float32)
    load().orc_cell_homography_inv(C.byref(camera(cam)), C.byref(grid(g)), ix, iy, iz, _p(m))
    return m


def cell_angles(g, ix, iy, iz):
    th = np.zeros(3, dtype=np.float64)
    load().orc_cell_angles(C.byref(grid(g)), ix, iy, iz, _p(th))
    return th


def warp(src, minv):
    src = np.ascontiguousarray(src, dtype=np.uint8)
    H, W = src.shape
    dst = np.empty_like(src)
    minv = np.ascontiguousarray(minv, dtype=np.float32)
    load().orc_warp(_p(src), W, H, _p(minv), _p(dst))
    return dst


def joint_hist(render, warped, bins=256, bg=True):
    render = np.ascontiguousarray(render, dtype=np.uint8)
    warped = np.ascontiguousarray(warped, dtype=np.uint8)
    J = np.zeros((bins, bins), dtype=np.uint32)
    HA = np.zeros(bins, dtype=np.uint32)
    HB = np.zeros(bins, dtype=np.uint32)
    load().orc_joint_hist(_p(render), _p(warped), render.size, bins, int(bg), _p(J), _p(HA), _p(HB))
    return J, HA, HB


def score_f32(J, HA, HB, length, mode=SUC):
    return float(load().orc_score_f32(_p(J), _p(HA), _p(HB), J.shape[0], int(length), mode))


def score_stages_f32(J, HA, HB, length, mode=SUC):
    """-> dict(ea, eb, ej, mid, sums[3], score): the score computation stage by stage."""
    bins = J.shape[0]
    ea = np.zeros(bins, np.float32); eb = np.zeros(bins, np.float32)
    ej = np.zeros((bins, bins), np.float32); mid = np.zeros(bins, np.float32)
    sums = np.zeros(3, np.float32)
    J = np.ascontiguousarray(J, np.uint32); HA = np.ascontiguousarray(HA, np.uint32)
    HB = np.ascontiguousarray(HB, np.uint32)
    s = load().orc_score_stages_f32(_p(J), _p(HA), _p(HB), bins, int(length), mode,
                                    _p(ea), _p(eb), _p(ej), _p(mid), _p(sums))
    return dict(ea=ea, eb=eb, ej=ej, mid=mid, sums=sums, score=float(s))


def log2f(x):
    return float(load().orc_log2f(float(x)))


def tree_f32(x):
    x = np.ascontiguousarray(x, np.float32)
    return float(load().orc_tree_f32(_p(x), x.size))


def finish_f32(sa, sb, sab, mode=SUC):
    return float(load().orc_finish_f32(float(sa), float(sb), float(sab), mode))


def score_f64(J, HA, HB, length, mode=SUC):
    return float(load().orc_score_f64(_p(J), _p(HA), _p(HB), J.shape[0], int(length), mode))


def eval_one(render, warped, bins=256, bg=True, mode=SUC):
    render = np.ascontiguousarray(render, dtype=np.uint8)
    warped = np.ascontiguousarray(warped, dtype=np.uint8)
    H, W = render.shape
    return float(load().orc_eval_one(_p(render), _p(warped), W, H, bins, int(bg), mode))


def argmax(scores):
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    m = np.zeros(1, dtype=np.float32)
    i = load().orc_argmax(_p(scores), scores.size, _p(m))
    return int(i), float(m[0])


def unravel(g, l):
    s = np.zeros(3, dtype=np.int32)
    w = np.zeros(3, dtype=np.int32)
    load().orc_unravel_index(C.byref(grid(g)), l, _p(s), _p(w))
    return tuple(int(v) for v in s), tuple(int(v) for v in w)


def search_points(cam, Twc, g, xyzi, frame, bins=256, bg=True, mode=SUC, keep_images=False, threads=0):
    og = grid(g)
    nS = g.nS[0] * g.nS[1] * g.nS[2]
    nW = g.nW[0] * g.nW[1] * g.nW[2]
    scores = np.zeros(nS * nW, dtype=np.float32)
    P = cam.W * cam.H
    renders = np.empty((nS, cam.H, cam.W), dtype=np.uint8) if keep_images else None
    warps = np.empty((nW, cam.H, cam.W), dtype=np.uint8) if keep_images else None
    T = _twc(Twc)
    xyzi = np.ascontiguousarray(xyzi, dtype=np.float32)
    frame = np.ascontiguousarray(frame, dtype=np.uint8)
    load().orc_search_points(C.byref(camera(cam)), _p(T), C.byref(og), _p(xyzi), xyzi.shape[0],
                             _p(frame), bins, int(bg), mode, _p(scores), _p(renders), _p(warps), threads)
    return scores, renders, warps


def apply_winner(Twc, g, s, w):
    out = np.zeros(16, dtype=np.float32)
    T = _twc(Twc)
    s = np.asarray(s, dtype=np.int32)
    w = np.asarray(w, dtype=np.int32)
    load().orc_apply_winner(_p(T), C.byref(grid(g)), _p(s), _p(w), _p(out))
    return out.reshape(4, 4)


def is_middle(g, s, w):
    s = np.asarray(s, dtype=np.int32)
    w = np.asarray(w, dtype=np.int32)
    return bool(load().orc_is_middle(C.byref(grid(g)), _p(s), _p(w)))


def resize_grid(g, s, w):
    og = grid(g)
    s = np.asarray(s, dtype=np.int32)
    w = np.asarray(w, dtype=np.int32)
    load().orc_resize_grid(C.byref(og), _p(s), _p(w))
    return og


def grid_from_motion(initial, dist, rot, not_initialized=False):
    out = OrcGrid()
    d = np.asarray(dist, dtype=np.float32)
    r = np.asarray(rot, dtype=np.float32)
    load().orc_grid_from_motion(C.byref(grid(initial)), _p(d), _p(r), int(not_initialized), C.byref(out))
    return out


def relocalize_points(cam, Twc, g, xyzi, frame, threshold, max_iterations=4, dist=(0, 0, 0), rot=(0, 0, 0),
                      bins=256, bg=True, mode=SUC, threads=0):
    prm = OrcRelocParams(threshold, max_iterations, (C.c_float * 3)(*dist), (C.c_float * 3)(*rot))
    out = OrcRelocResult()
    T = _twc(Twc)
    xyzi = np.ascontiguousarray(xyzi, dtype=np.float32)
    frame = np.ascontiguousarray(frame, dtype=np.uint8)
    rc = load().orc_relocalize_points(C.byref(camera(cam)), _p(T), C.byref(grid(g)), _p(xyzi), xyzi.shape[0],
                                      _p(frame), bins, int(bg), mode, C.byref(prm), C.byref(out), threads)
    return rc, out
