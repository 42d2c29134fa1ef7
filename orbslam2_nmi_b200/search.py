"""Thin Python host over the C ABI -- used by tests/, bench.py and the multi-GPU launcher.

Mirrors the reference's host objects only as far as the harness needs them:
`NmiSearcher` ~ NmiObjects (localization.hpp:31) + the grid loop of
Tracking::RelocalizeWithNMI (src/Tracking.cc:1851-1985).  The production host side is
the C++ drop-in layer in include/compat/ + libnmi_b200.so; this module adds nothing
numerical -- every number comes from the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi
from .capi import Camera, Flags, Grid, Result, check, ptr


@dataclass
class SearchResult:
    best_s: tuple
    best_w: tuple
    best_index: int
    best_score: float
    key: int
    gpu_ms: float
    scores: np.ndarray | None = None


def _result(r: Result, scores=None) -> SearchResult:
    return SearchResult(tuple(r.best_s), tuple(r.best_w), int(r.best_index), float(r.best_score),
                        int(r.key), float(r.gpu_ms), scores)


class NmiSearcher:
    """One context on one GPU."""

    def __init__(self, device: int = 0):
        self.lib = capi.load()
        h = C.c_void_p()
        check(self.lib.nmi_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device
        self.cam: Camera | None = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.nmi_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- model / camera / frame ------------------------------------------------
    def set_camera(self, W, H, fx, fy, cx, cy, zn, zf, point_size=3.0):
        self.cam = Camera(int(W), int(H), fx, fy, cx, cy, zn, zf, point_size)
        check(self.lib.nmi_set_camera(self.h, C.byref(self.cam)))

    def set_scene(self, scene):
        self.set_camera(scene.W, scene.H, scene.fx, scene.fy, scene.cx, scene.cy, scene.zn, scene.zf,
                        scene.point_size)
        self.set_points(scene.xyzi)

    def set_points(self, xyzi: np.ndarray):
        xyzi = np.ascontiguousarray(xyzi, dtype=np.float32)
        assert xyzi.ndim == 2 and xyzi.shape[1] == 4
        check(self.lib.nmi_set_points(self.h, ptr(xyzi), xyzi.shape[0]))

    def set_mesh(self, verts: np.ndarray, tris: np.ndarray):
        verts = np.ascontiguousarray(verts, dtype=np.float32)
        tris = np.ascontiguousarray(tris, dtype=np.uint32)
        assert verts.ndim == 2 and verts.shape[1] == 4 and tris.ndim == 2 and tris.shape[1] == 3
        check(self.lib.nmi_set_mesh(self.h, ptr(verts), verts.shape[0], ptr(tris), tris.shape[0]))

    def set_mesh_textured(self, verts: np.ndarray, tris: np.ndarray, corner_uv: np.ndarray, texture: np.ndarray):
        """Rendering<1> with its texture: corner_uv (nt, 3, 2) float32, texture (th, tw, 3) u8 in file byte order."""
        verts = np.ascontiguousarray(verts, dtype=np.float32)
        tris = np.ascontiguousarray(tris, dtype=np.uint32)
        uv = np.ascontiguousarray(corner_uv, dtype=np.float32)
        tex = np.ascontiguousarray(texture, dtype=np.uint8)
        assert uv.size == 6 * tris.shape[0] and tex.ndim == 3 and tex.shape[2] == 3
        check(self.lib.nmi_set_mesh_textured(self.h, ptr(verts), verts.shape[0], ptr(tris), tris.shape[0], ptr(uv),
                                             ptr(tex), tex.shape[1], tex.shape[0]))

    def set_frame(self, gray: np.ndarray):
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        check(self.lib.nmi_set_frame(self.h, ptr(gray), gray.shape[1], gray.shape[0]))

    def set_frame_device(self, dev_ptr: int, W: int, H: int):
        check(self.lib.nmi_set_frame_device(self.h, dev_ptr, W, H))

    # -- search -----------------------------------------------------------------
    @staticmethod
    def flags(bins=256, score=capi.SCORE_SUC, bg=True, variant=0) -> Flags:
        return Flags(bins, score, 1 if bg else 0, variant)

    def search(self, Twc, grid: Grid, flags: Flags | None = None, want_scores=False) -> SearchResult:
        Twc = np.ascontiguousarray(Twc, dtype=np.float32).reshape(16)
        flags = flags or self.flags()
        r = Result()
        scores = np.empty(grid.n_pose, dtype=np.float32) if want_scores else None
        code = self.lib.nmi_search(self.h, ptr(Twc), C.byref(grid), C.byref(flags), C.byref(r),
                                   ptr(scores) if want_scores else None)
        if code not in (capi.NMI_OK, capi.NMI_ERR_NO_WINNER):
            check(code)
        return _result(r, scores)

    def search_enqueue(self, Twc, grid: Grid, flags: Flags, rank: int, world: int, key_dev: int,
                       scores_dev: int | None = None):
        Twc = np.ascontiguousarray(Twc, dtype=np.float32).reshape(16)
        check(self.lib.nmi_search_enqueue(self.h, ptr(Twc), C.byref(grid), C.byref(flags), rank,
                                          world, key_dev, scores_dev))

    def stream(self) -> int:
        return int(self.lib.nmi_ctx_stream(self.h) or 0)

    def sync(self):
        check(self.lib.nmi_ctx_sync(self.h))

    def set_hist_skip(self, mode: int):
        """Hot-bin skipping of the histogram kernel: 0 never, 1 automatic, 2 always."""
        check(self.lib.nmi_ctx_set_hist_skip(self.h, int(mode)))

    def decode(self, grid: Grid, key: int, strict: bool = True) -> SearchResult:
        """nmi_decode_key.  strict: NMI_ERR_RETRY (a rank's splat bins overflowed: the key carries no
        winner, the level must be redone) raises NmiError instead of returning best_index -1."""
        r = Result()
        code = self.lib.nmi_decode_key(C.byref(grid), C.c_uint64(key), C.byref(r))
        if strict and code == capi.NMI_ERR_RETRY:
            raise capi.NmiError(code, "reduced key is NMI_KEY_RETRY: a rank's record bins overflowed, redo the search")
        return _result(r)

    def read_key(self, key_dev: int) -> int:
        """nmi_read_key: stream-synchronise and fetch the (reduced) key; a local bin overflow arms the
        exact two-pass sizing for this context's next search."""
        k = C.c_uint64(0)
        check(self.lib.nmi_read_key(self.h, key_dev, C.byref(k)))
        return int(k.value)

    def relocalize(self, Twc, grid: Grid, flags: Flags | None = None, threshold=0.1, max_iterations=4,
                   dist=(0.0, 0.0, 0.0), rot=(0.0, 0.0, 0.0)):
        """Tracking::RelocalizeWithNMIStrategy on a pose (csrc/driver.cpp)."""
        Twc = np.ascontiguousarray(Twc, dtype=np.float32).reshape(16)
        flags = flags or self.flags()
        prm = capi.RelocParams(threshold, max_iterations, (C.c_float * 3)(*dist), (C.c_float * 3)(*rot))
        out = capi.RelocResult()
        check(self.lib.nmi_relocalize(self.h, ptr(Twc), C.byref(grid), C.byref(flags), C.byref(prm),
                                      C.byref(out)))
        return out

    def relocalize_sharded(self, Twc, grid: Grid, flags: Flags | None, rank: int, world: int, key_dev: int,
                           exchange, threshold=0.1, max_iterations=4, dist=(0.0, 0.0, 0.0),
                           rot=(0.0, 0.0, 0.0)):
        """nmi_relocalize_sharded: the level driver with every level sharded over `world` ranks;
        `exchange(key_dev, stream)` max-reduces the 8-byte key in place on that stream."""
        Twc = np.ascontiguousarray(Twc, dtype=np.float32).reshape(16)
        flags = flags or self.flags()
        prm = capi.RelocParams(threshold, max_iterations, (C.c_float * 3)(*dist), (C.c_float * 3)(*rot))
        out = capi.RelocResult()
        failure = []

        def _cb(_user, kdev, stream):
            try:
                exchange(kdev, stream)
                return 0
            except BaseException as e:  # never unwind through the C frame
                failure.append(e)
                return 1

        cb = capi.EXCHANGE_FN(_cb)
        rc = self.lib.nmi_relocalize_sharded(self.h, ptr(Twc), C.byref(grid), C.byref(flags), C.byref(prm),
                                             rank, world, C.c_void_p(key_dev), cb, None, C.byref(out))
        if failure:
            raise failure[0]
        check(rc)
        return out

    def level_trace(self):
        """Per level of this thread's last relocalize_sharded: host microseconds spent enqueueing the
        level, in the exchange callback, waiting for the key, and this rank's device time."""
        out, us = [], (C.c_float * 4)()
        while self.lib.nmi_last_level_trace(len(out), us) == 0:
            out.append({"enqueue_us": us[0], "exchange_us": us[1], "wait_us": us[2], "device_us": us[3]})
        return out

    # -- stage-level API (reference call granularity) ------------------------------
    def render_cell(self, Twc, grid: Grid, sx, sy, sz) -> int:
        Twc = np.ascontiguousarray(Twc, dtype=np.float32).reshape(16)
        h = C.c_uint(0)
        check(self.lib.nmi_render_cell(self.h, ptr(Twc), C.byref(grid), sx, sy, sz, C.byref(h)))
        return h.value

    def warp_cells(self, grid: Grid):
        check(self.lib.nmi_warp_cells(self.h, C.byref(grid)))

    def warp_ptr(self, grid: Grid, wx, wy, wz) -> int:
        p = C.c_void_p()
        check(self.lib.nmi_warp_ptr(self.h, C.byref(grid), wx, wy, wz, C.byref(p)))
        return int(p.value)

    def eval_pair(self, warped_dev: int, handle: int, flags: Flags | None = None) -> float:
        if flags is None:
            flags = self._default_flags = getattr(self, "_default_flags", None) or self.flags()
        s = self._one_score = getattr(self, "_one_score", None)
        if s is None:
            s = self._one_score = C.c_float(0.0)
        check(self.lib.nmi_eval_pair(self.h, warped_dev, handle, self.cam.W, self.cam.H,
                                     C.byref(flags), C.byref(s)))
        return s.value

    def import_render(self, render_dev: int, pitch_bytes: int, bottom_up: bool = False) -> int:
        """Adopt a device image as the current render (kernel.cu:53-59's mapped GL texture)."""
        h = C.c_uint(0)
        check(self.lib.nmi_import_render(self.h, render_dev, pitch_bytes, self.cam.W, self.cam.H,
                                         int(bottom_up), C.byref(h)))
        return h.value

    def eval_pair_dev(self, warped_dev: int, handle: int, J_dev: int, HA_dev: int, HB_dev: int,
                      flags: Flags | None = None) -> float:
        """NMIWithCuda_noMask with the integer histograms left on the device (NMI.cuh:63-71 layout)."""
        flags = flags or self.flags()
        s = np.zeros(1, dtype=np.float32)
        check(self.lib.nmi_eval_pair_dev(self.h, warped_dev, handle, self.cam.W, self.cam.H,
                                         C.byref(flags), J_dev, HA_dev, HB_dev, ptr(s)))
        return float(s[0])

    # -- parity read-backs -------------------------------------------------------
    def get_render(self, s: int) -> np.ndarray:
        out = np.empty((self.cam.H, self.cam.W), dtype=np.uint8)
        check(self.lib.nmi_get_render(self.h, s, ptr(out)))
        return out

    def get_winners(self, s: int) -> np.ndarray:
        out = np.empty((self.cam.H, self.cam.W), dtype=np.uint32)
        check(self.lib.nmi_get_winners(self.h, s, ptr(out)))
        return out

    def get_warp(self, w: int) -> np.ndarray:
        out = np.empty((self.cam.H, self.cam.W), dtype=np.uint8)
        check(self.lib.nmi_get_warp(self.h, w, ptr(out)))
        return out

    def get_hist(self, s: int, w: int, flags: Flags | None = None, path: int = 0):
        """Integer histograms + score of pair (s, w).  path 0: automatic; 1: the plain batched build
        (variant 0 = the persistent kernel with the fast epilogue, the code a benchmark step times);
        2: the build with the hot-bin side tables."""
        flags = flags or self.flags()
        b = flags.bins
        J = np.zeros((b, b), dtype=np.uint32)
        HA = np.zeros(b, dtype=np.uint32)
        HB = np.zeros(b, dtype=np.uint32)
        sc = np.zeros(1, dtype=np.float32)
        check(self.lib.nmi_get_hist_path(self.h, s, w, C.byref(flags), path, ptr(J), ptr(HA), ptr(HB), ptr(sc)))
        return J, HA, HB, float(sc[0])

    def last_hist_path(self) -> int:
        return int(self.lib.nmi_last_hist_path(self.h))

    def score_pairs(self, renders_dev: int, n_r: int, r_stride: int, warps_dev: int, n_w: int, w_stride: int,
                    flags: Flags | None = None) -> np.ndarray:
        """nmi_score_pairs: every (render, warp) pair of two device image stacks through the batched
        launch of a grid search; scores[w, r]."""
        flags = flags or self.flags()
        out = np.zeros(n_r * n_w, dtype=np.float32)
        check(self.lib.nmi_score_pairs(self.h, renders_dev, n_r, r_stride, warps_dev, n_w, w_stride,
                                       self.cam.W, self.cam.H, C.byref(flags), ptr(out)))
        return out.reshape(n_w, n_r)

    def timings(self):
        ms = np.zeros(8, dtype=np.float32)
        n = C.c_int(0)
        check(self.lib.nmi_get_timings(self.h, ptr(ms), C.byref(n)))
        names = ["params_cull", "render", "spare", "warp", "hist_score", "argmax", "total"]
        return {k: float(ms[i]) for i, k in enumerate(names)}, n.value


# -- host helpers that need no GPU (bound straight to the library's host code) ------
def partition(grid: Grid, rank: int, world: int):
    lib = capi.load()
    ax, b, e = C.c_int(), C.c_int(), C.c_int()
    check(lib.nmi_partition(C.byref(grid), rank, world, C.byref(ax), C.byref(b), C.byref(e)))
    return ax.value, b.value, e.value


def cell_translation(Twc, grid: Grid, sx, sy, sz) -> np.ndarray:
    Twc = np.ascontiguousarray(Twc, dtype=np.float32).reshape(16)
    t = np.zeros(3, dtype=np.float32)
    capi.load().nmi_cell_translation(ptr(Twc), C.byref(grid), sx, sy, sz, ptr(t))
    return t


def cell_homography_inv(cam: Camera, grid: Grid, wx, wy, wz) -> np.ndarray:
    m = np.zeros(9, dtype=np.float32)
    capi.load().nmi_cell_homography_inv(C.byref(cam), C.byref(grid), wx, wy, wz, ptr(m))
    return m


def apply_winner(Twc, grid: Grid, s, w) -> np.ndarray:
    Twc = np.ascontiguousarray(Twc, dtype=np.float32).reshape(16)
    s = np.asarray(s, dtype=np.int32)
    w = np.asarray(w, dtype=np.int32)
    out = np.zeros(16, dtype=np.float32)
    capi.load().nmi_apply_winner(ptr(Twc), C.byref(grid), ptr(s), ptr(w), ptr(out))
    return out.reshape(4, 4)


def grid_is_middle(grid: Grid, s, w) -> bool:
    s = np.asarray(s, dtype=np.int32)
    w = np.asarray(w, dtype=np.int32)
    return bool(capi.load().nmi_grid_is_middle(C.byref(grid), ptr(s), ptr(w)))


def grid_resize(grid: Grid, s, w) -> Grid:
    g = grid.copy()
    s = np.asarray(s, dtype=np.int32)
    w = np.asarray(w, dtype=np.int32)
    capi.load().nmi_grid_resize(C.byref(g), ptr(s), ptr(w))
    return g


def grid_from_motion(initial: Grid, dist, rot, not_initialized=False) -> Grid:
    out = Grid()
    d = np.asarray(dist, dtype=np.float32)
    r = np.asarray(rot, dtype=np.float32)
    capi.load().nmi_grid_from_motion(C.byref(initial), ptr(d), ptr(r), int(not_initialized), C.byref(out))
    return out


def relocalize_with(level_search, Twc, grid: Grid, threshold=0.1, max_iterations=4, dist=(0.0, 0.0, 0.0),
                    rot=(0.0, 0.0, 0.0)):
    """nmi_relocalize_with: the coarse-to-fine driver (host logic only, no GPU needed) over a
    caller-supplied level search  level_search(Twc[4x4], grid) -> (best_s, best_w, best_score)
    or an integer error code."""
    Twc = np.ascontiguousarray(Twc, dtype=np.float32).reshape(16)
    prm = capi.RelocParams(threshold, max_iterations, (C.c_float * 3)(*dist), (C.c_float * 3)(*rot))
    out = capi.RelocResult()
    failure = []

    def _cb(_user, twc_p, grid_p, res_p):
        try:
            T = np.array([twc_p[i] for i in range(16)], dtype=np.float32).reshape(4, 4)
            r = level_search(T, grid_p.contents.copy())
            if isinstance(r, int):
                return r
            bs, bw, score = r
            res = res_p.contents
            for k in range(3):
                res.best_s[k] = int(bs[k])
                res.best_w[k] = int(bw[k])
            res.best_score = float(score)
            return 0
        except BaseException as e:
            failure.append(e)
            return capi.NMI_ERR_INVALID

    cb = capi.LEVEL_SEARCH_FN(_cb)
    rc = capi.load().nmi_relocalize_with(cb, None, ptr(Twc), C.byref(grid), C.byref(prm), C.byref(out))
    if failure:
        raise failure[0]
    check(rc)
    return out


def decode_key(grid: Grid, key: int) -> SearchResult:
    r = Result()
    capi.load().nmi_decode_key(C.byref(grid), C.c_uint64(key), C.byref(r))
    return _result(r)
