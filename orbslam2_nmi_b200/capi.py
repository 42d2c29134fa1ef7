"""ctypes binding of the C ABI in include/nmi_b200.h (libnmi_b200.so).

This is harness plumbing (tests, bench, multi-GPU launcher): the product is the
shared library.  There is no CPU fallback -- if the library is missing or no B200
is visible, calls raise.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
import os

# NMI_B200_LIB lets an experiment load an alternative build of the same library
LIB_PATH = Path(os.environ.get("NMI_B200_LIB", ROOT / "orbslam2_nmi_b200" / "_lib" / "libnmi_b200.so"))

NMI_OK, NMI_ERR_INVALID, NMI_ERR_CUDA, NMI_ERR_STATE, NMI_ERR_NO_WINNER = range(5)
SCORE_ENMI, SCORE_SUC = 0, 1
EMPTY = 0xFFFFFFFF


class Camera(C.Structure):
    _fields_ = [("W", C.c_int), ("H", C.c_int), ("fx", C.c_double), ("fy", C.c_double),
                ("cx", C.c_double), ("cy", C.c_double), ("zn", C.c_double), ("zf", C.c_double),
                ("point_size", C.c_float)]


class Grid(C.Structure):
    _fields_ = [("nS", C.c_int * 3), ("nW", C.c_int * 3), ("stepT", C.c_float * 3),
                ("stepR", C.c_float * 3)]

    @classmethod
    def make(cls, nS, nW, stepT, stepR) -> "Grid":
        g = cls()
        for k in range(3):
            g.nS[k], g.nW[k] = int(nS[k]), int(nW[k])
            g.stepT[k], g.stepR[k] = float(stepT[k]), float(stepR[k])
        return g

    def copy(self) -> "Grid":
        return Grid.make(list(self.nS), list(self.nW), list(self.stepT), list(self.stepR))

    @property
    def n_synth(self) -> int:
        return self.nS[0] * self.nS[1] * self.nS[2]

    @property
    def n_warp(self) -> int:
        return self.nW[0] * self.nW[1] * self.nW[2]

    @property
    def n_pose(self) -> int:
        return self.n_synth * self.n_warp


class Flags(C.Structure):
    _fields_ = [("bins", C.c_int), ("score_mode", C.c_int), ("bg", C.c_int), ("variant", C.c_int)]


class RelocParams(C.Structure):
    _fields_ = [("threshold", C.c_float), ("max_iterations", C.c_int),
                ("distance_since_last", C.c_float * 3), ("rotation_since_last", C.c_float * 3)]


MAX_PREV_POSES = 8


MAX_LEVELS = 8


class RelocLevel(C.Structure):
    _fields_ = [("grid", Grid), ("best_s", C.c_int32 * 3), ("best_w", C.c_int32 * 3),
                ("nmi", C.c_float), ("last_nmi", C.c_float)]


class RelocResult(C.Structure):
    _fields_ = [("Twc", C.c_float * 16), ("relocalized", C.c_int), ("failed", C.c_int),
                ("iterations", C.c_int), ("nmi", C.c_float), ("last_nmi", C.c_float),
                ("threshold_used", C.c_float), ("final_grid", Grid), ("last_search_grid", Grid),
                ("best_s", C.c_int32 * 3), ("best_w", C.c_int32 * 3), ("n_prev", C.c_int),
                ("prev_Twc", (C.c_float * 16) * MAX_PREV_POSES), ("n_evals", C.c_int),
                ("gpu_ms", C.c_float), ("n_levels", C.c_int), ("levels", RelocLevel * MAX_LEVELS)]


class Result(C.Structure):
    _fields_ = [("best_s", C.c_int32 * 3), ("best_w", C.c_int32 * 3), ("best_index", C.c_int64),
                ("best_score", C.c_float), ("key", C.c_uint64), ("gpu_ms", C.c_float)]


# every symbol include/nmi_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
_F16 = C.POINTER(C.c_float)
# nmi_level_search_fn / nmi_exchange_fn (include/nmi_b200.h)
LEVEL_SEARCH_FN = C.CFUNCTYPE(C.c_int, _P, _F16, C.POINTER(Grid), C.POINTER(Result))
EXCHANGE_FN = C.CFUNCTYPE(C.c_int, _P, _P, _P)
NMI_ERR_NO_WINNER = 4
NMI_ERR_RETRY = 5
NMI_KEY_RETRY = 0x7FFFFFFFFFFFFFFF
SYMBOLS = [
    ("nmi_ctx_create", C.c_int, [C.c_int, C.POINTER(_P)]),
    ("nmi_ctx_destroy", None, [_P]),
    ("nmi_last_error", C.c_char_p, []),
    ("nmi_ctx_stream", _P, [_P]),
    ("nmi_ctx_sync", C.c_int, [_P]),
    ("nmi_ctx_set_hist_skip", C.c_int, [_P, C.c_int]),
    ("nmi_set_camera", C.c_int, [_P, C.POINTER(Camera)]),
    ("nmi_set_points", C.c_int, [_P, _P, C.c_size_t]),
    ("nmi_set_points_device", C.c_int, [_P, _P, C.c_size_t]),
    ("nmi_set_mesh", C.c_int, [_P, _P, C.c_size_t, _P, C.c_size_t]),
    ("nmi_set_mesh_textured", C.c_int, [_P, _P, C.c_size_t, _P, C.c_size_t, _P, _P, C.c_int, C.c_int]),
    ("nmi_set_frame", C.c_int, [_P, _P, C.c_int, C.c_int]),
    ("nmi_set_frame_device", C.c_int, [_P, _P, C.c_int, C.c_int]),
    ("nmi_search", C.c_int, [_P, _P, C.POINTER(Grid), C.POINTER(Flags), C.POINTER(Result), _P]),
    ("nmi_search_enqueue", C.c_int, [_P, _P, C.POINTER(Grid), C.POINTER(Flags), C.c_int, C.c_int, _P, _P]),
    ("nmi_partition", C.c_int, [C.POINTER(Grid), C.c_int, C.c_int, C.POINTER(C.c_int),
                                C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("nmi_decode_key", C.c_int, [C.POINTER(Grid), C.c_uint64, C.POINTER(Result)]),
    ("nmi_grid_from_motion", None, [C.POINTER(Grid), _P, _P, C.c_int, C.POINTER(Grid)]),
    ("nmi_relocalize", C.c_int, [_P, _P, C.POINTER(Grid), C.POINTER(Flags), C.POINTER(RelocParams),
                                 C.POINTER(RelocResult)]),
    ("nmi_relocalize_with", C.c_int, [LEVEL_SEARCH_FN, _P, _P, C.POINTER(Grid), C.POINTER(RelocParams),
                                      C.POINTER(RelocResult)]),
    ("nmi_relocalize_sharded", C.c_int, [_P, _P, C.POINTER(Grid), C.POINTER(Flags), C.POINTER(RelocParams),
                                         C.c_int, C.c_int, _P, EXCHANGE_FN, _P, C.POINTER(RelocResult)]),
    ("nmi_read_key", C.c_int, [_P, _P, C.POINTER(C.c_uint64)]),
    ("nmi_ctx_key_buffer", _P, [_P]),
    ("nmi_render_at", C.c_int, [_P, _P, _P, C.POINTER(C.c_uint)]),
    ("nmi_render_cell", C.c_int, [_P, _P, C.POINTER(Grid), C.c_int, C.c_int, C.c_int,
                                  C.POINTER(C.c_uint)]),
    ("nmi_warp_cells", C.c_int, [_P, C.POINTER(Grid)]),
    ("nmi_warp_ptr", C.c_int, [_P, C.POINTER(Grid), C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    ("nmi_eval_pair", C.c_int, [_P, _P, C.c_uint, C.c_int, C.c_int, C.POINTER(Flags), _P]),
    ("nmi_eval_pair_dev", C.c_int, [_P, _P, C.c_uint, C.c_int, C.c_int, C.POINTER(Flags), _P, _P, _P, _P]),
    ("nmi_import_render", C.c_int, [_P, _P, C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint)]),
    ("nmi_cell_translation", None, [_P, C.POINTER(Grid), C.c_int, C.c_int, C.c_int, _P]),
    ("nmi_cell_homography_inv", None, [C.POINTER(Camera), C.POINTER(Grid), C.c_int, C.c_int, C.c_int, _P]),
    ("nmi_apply_winner", None, [_P, C.POINTER(Grid), _P, _P, _P]),
    ("nmi_grid_is_middle", C.c_int, [C.POINTER(Grid), _P, _P]),
    ("nmi_grid_resize", None, [C.POINTER(Grid), _P, _P]),
    ("nmi_get_render", C.c_int, [_P, C.c_int, _P]),
    ("nmi_get_winners", C.c_int, [_P, C.c_int, _P]),
    ("nmi_get_warp", C.c_int, [_P, C.c_int, _P]),
    ("nmi_get_hist", C.c_int, [_P, C.c_int, C.c_int, C.POINTER(Flags), _P, _P, _P, _P]),
    ("nmi_get_hist_path", C.c_int, [_P, C.c_int, C.c_int, C.POINTER(Flags), C.c_int, _P, _P, _P, _P]),
    ("nmi_last_hist_path", C.c_int, [_P]),
    ("nmi_score_pairs", C.c_int, [_P, _P, C.c_int, C.c_size_t, _P, C.c_int, C.c_size_t, C.c_int, C.c_int,
                                  C.POINTER(Flags), _P]),
    ("nmi_get_timings", C.c_int, [_P, _P, C.POINTER(C.c_int)]),
    ("nmi_last_level_trace", C.c_int, [C.c_int, _P]),
]

_lib = None


def load() -> C.CDLL:
    """Load libnmi_b200.so and bind every declared symbol. Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is not built (run `python -m orbslam2_nmi_b200.build`); "
            "the NMI search has no CPU fallback")
    lib = C.CDLL(str(LIB_PATH))
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class NmiError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"nmi error {code}: {msg}")
        self.code = code


def check(code: int) -> None:
    if code != NMI_OK:
        raise NmiError(code, load().nmi_last_error().decode(errors="replace"))


def ptr(a: np.ndarray) -> int:
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data
