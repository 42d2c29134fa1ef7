"""ctypes loader of oracle/_build/libnmi_nppcheck.so (NPP's nppiWarpPerspective_8u_C1R behind a C ABI).

TEST INFRASTRUCTURE ONLY: a checker for the warp stage, used by the GPU tests and tools/; never
imported by the orbslam2_nmi_b200 package.  Needs a GPU (NPP runs on the device)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_build" / "libnmi_nppcheck.so"
_lib = None


def available() -> bool:
    return LIB_PATH.exists()


def load():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            subprocess.run(["make", "-C", str(HERE), "npp"], check=True, capture_output=True)
        _lib = C.CDLL(str(LIB_PATH))
        _lib.nppchk_warp_perspective_8u.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    return _lib


def version() -> str:
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    load().nppchk_version(C.byref(a), C.byref(b), C.byref(c))
    return f"{a.value}.{b.value}.{c.value}"


def warp_perspective(src: np.ndarray, M: np.ndarray, linear: bool = True) -> np.ndarray:
    """dst = nppiWarpPerspective_8u_C1R(src, forward M), destination cleared to 0 first."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    M = np.ascontiguousarray(M, dtype=np.float64).reshape(9)
    dst = np.empty_like(src)
    st = load().nppchk_warp_perspective_8u(src.ctypes.data, src.shape[1], src.shape[0], M.ctypes.data,
                                           int(linear), dst.ctypes.data)
    if st < 0:
        raise RuntimeError(f"nppiWarpPerspective_8u_C1R failed: status {st}")
    return dst
