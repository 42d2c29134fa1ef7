"""ctypes access to oracle/_ref/libnmi_ref.so -- the reference's OWN CUDA NMI routine
(Thirdparty/CUDA_Functions/NMI.cu + kernel.cu, compiled unmodified for sm_100a by
oracle/Makefile.ref in the build container; see oracle/ref_harness.cu and
oracle/ref_shim/refshim.h for what is supplied around it).

TEST INFRASTRUCTURE ONLY: imported by tests/ and by bench.py's baseline leg, never by the
product.  Needs a GPU to run (the reference has no CPU path); `/root/reference` is needed
only to BUILD the library, never at run time."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_ref" / "libnmi_ref.so"
REF_SOURCES = Path("/root/reference/Thirdparty/CUDA_Functions")

_lib = None


def build(force: bool = False) -> Path | None:
    """Compile the reference sources where they lie; None when /root/reference is absent."""
    if not (REF_SOURCES / "NMI.cu").exists():
        return LIB_PATH if LIB_PATH.exists() else None
    if force and LIB_PATH.exists():
        LIB_PATH.unlink()
    subprocess.run(["make", "-C", str(HERE), "-f", "Makefile.ref"], check=True, capture_output=True)
    return LIB_PATH


def available() -> bool:
    return LIB_PATH.exists()


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} not built (python -m orbslam2_nmi_b200.build, needs /root/reference)")
        lib = C.CDLL(str(LIB_PATH))   # RTLD_LOCAL: its CUDAF::NMIWithCuda_noMask never meets ours
        P = C.c_void_p
        lib.nmiref_score.argtypes = [P, P, C.c_int, C.c_int, P]
        lib.nmiref_stages.argtypes = [P, P, C.c_int, C.c_int] + [P] * 9
        lib.nmiref_time.argtypes = [P, P, C.c_int, C.c_int, C.c_int, P, P]
        lib.nmiref_describe.restype = C.c_char_p
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _gl_rows(render_topdown):
    """Our renders are top-down; the GL colour texture the reference samples is bottom-up
    (it flips the row itself, NMI.cu:82)."""
    r = np.ascontiguousarray(render_topdown, dtype=np.uint8)
    return np.ascontiguousarray(r[::-1])


def describe() -> str:
    return load().nmiref_describe().decode()


def score(render_topdown, warped) -> float:
    """CUDAF::NMIWithCuda_noMask (kernel.cu:49-114) on one pair, as Tracking calls it."""
    r = _gl_rows(render_topdown)
    w = np.ascontiguousarray(warped, dtype=np.uint8)
    H, W = r.shape
    out = np.zeros(1, np.float32)
    rc = load().nmiref_score(_p(r), _p(w), W, H, _p(out))
    if rc:
        raise RuntimeError(f"nmiref_score: CUDA error {rc}")
    return float(out[0])


def stages(render_topdown, warped) -> dict:
    """kernel.cu:57-100 stage by stage with every intermediate read back."""
    r = _gl_rows(render_topdown)
    w = np.ascontiguousarray(warped, dtype=np.uint8)
    H, W = r.shape
    J = np.zeros((256, 256), np.uint32); h1 = np.zeros(256, np.uint32); h2 = np.zeros(256, np.uint32)
    e1 = np.zeros(256, np.float32); e2 = np.zeros(256, np.float32); ej = np.zeros((256, 256), np.float32)
    mid = np.zeros(256, np.float32); sums = np.zeros(3, np.float32); raw = np.zeros(1, np.float32)
    rc = load().nmiref_stages(_p(r), _p(w), W, H, _p(J), _p(h1), _p(h2), _p(e1), _p(e2), _p(ej),
                              _p(mid), _p(sums), _p(raw))
    if rc:
        raise RuntimeError(f"nmiref_stages: CUDA error {rc}")
    return dict(J=J, HA=h1, HB=h2, ea=e1, eb=e2, ej=ej, mid=mid, sums=sums, raw_score=float(raw[0]))


def time_per_eval(render_topdown, warped, iters: int = 20):
    """(ms per NMIWithCuda_noMask call, last score) with the pair resident on the device."""
    r = _gl_rows(render_topdown)
    w = np.ascontiguousarray(warped, dtype=np.uint8)
    H, W = r.shape
    ms = C.c_double(0.0); s = C.c_float(0.0)
    rc = load().nmiref_time(_p(r), _p(w), W, H, int(iters), C.byref(ms), C.byref(s))
    if rc:
        raise RuntimeError(f"nmiref_time: CUDA error {rc}")
    return ms.value, s.value


# ---------------------------------------------------------------- host-side reference (CPU) ----
# oracle/_ref/libnmi_ref_host.so: Thirdparty/Localization/{nmiSearchKernel,helperFunctions}.cpp,
# compiled unmodified with g++ (oracle/Makefile.ref) + oracle/ref_host_harness.cpp.  Runs anywhere.
HOST_LIB_PATH = HERE / "_ref" / "libnmi_ref_host.so"
_hlib = None


def host_available() -> bool:
    return HOST_LIB_PATH.exists()


def load_host() -> C.CDLL:
    global _hlib
    if _hlib is None:
        if not HOST_LIB_PATH.exists():
            raise RuntimeError(f"{HOST_LIB_PATH} not built (needs /root/reference at build time)")
        lib = C.CDLL(str(HOST_LIB_PATH))
        P = C.c_void_p
        lib.nmirefh_find_max.argtypes = [P] * 6
        lib.nmirefh_resize.argtypes = [P] * 6
        lib.nmirefh_format.argtypes = [P] * 6 + [C.c_float, C.c_char_p, C.c_int]
        _hlib = lib
    return _hlib


def _i3(v):
    return np.ascontiguousarray(v, dtype=np.int32).reshape(3).copy()


def _f3(v):
    return np.ascontiguousarray(v, dtype=np.float32).reshape(3).copy()


def find_max(rating_linear, nS, nW):
    """helperFunctions::find_max_elements on the rating array (linear order wz,wy,wx,sz,sy,sx).
    -> (count of maximal elements, best_s, best_w, score) with element [0] of its vector."""
    r = np.ascontiguousarray(rating_linear, dtype=np.float32)
    s, w = _i3(nS), _i3(nW)
    assert r.size == int(np.prod(s)) * int(np.prod(w))
    bs, bw = np.full(3, -1, np.int32), np.full(3, -1, np.int32)
    sc = np.zeros(1, np.float32)
    n = load_host().nmirefh_find_max(_p(r), _p(s), _p(w), _p(bs), _p(bw), _p(sc))
    return n, tuple(int(x) for x in bs), tuple(int(x) for x in bw), float(sc[0])


def resize(nS, nW, stepT, stepR, best_s, best_w):
    """NmiSearchKernel::isMiddle (before) and ::resizeKernel -> (is_middle, nS, nW, stepT, stepR)."""
    s, w, t, r = _i3(nS), _i3(nW), _f3(stepT), _f3(stepR)
    bs, bw = _i3(best_s), _i3(best_w)
    mid = load_host().nmirefh_resize(_p(s), _p(w), _p(t), _p(r), _p(bs), _p(bw))
    return bool(mid), tuple(int(x) for x in s), tuple(int(x) for x in w), t, r


def format_kernel(nS, nW, stepT, stepR, best_s, best_w, nmi) -> str:
    """operator<<(std::ostream&, const NmiSearchKernel&): the _log.txt line."""
    s, w, t, r, bs, bw = _i3(nS), _i3(nW), _f3(stepT), _f3(stepR), _i3(best_s), _i3(best_w)
    buf = C.create_string_buffer(1024)
    load_host().nmirefh_format(_p(s), _p(w), _p(t), _p(r), _p(bs), _p(bw), C.c_float(nmi), buf, 1024)
    return buf.value.decode()


# ---- host-side geometry of the reference (oracle/_ref/libnmi_ref_geom.so; CPU only) ----------------
GEOM_LIB_PATH = HERE / "_ref" / "libnmi_ref_geom.so"
_geom = None


def geom_available() -> bool:
    return GEOM_LIB_PATH.exists()


def _geom_lib():
    global _geom
    if _geom is None:
        if not GEOM_LIB_PATH.exists():
            raise RuntimeError(f"{GEOM_LIB_PATH} not built (make -C oracle -f Makefile.ref, needs /root/reference)")
        _geom = C.CDLL(str(GEOM_LIB_PATH))
        P = C.c_void_p
        _geom.nmirefg_warp_matrices.argtypes = [P, P, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                                P, P, P]
        _geom.nmirefg_setup_cam.argtypes = [P, P, P, P]
        _geom.nmirefg_cell_translation.argtypes = [P, P, P, C.c_int, C.c_int, C.c_int, P]
        _geom.nmirefg_apply_winner.argtypes = [P] * 8
    return _geom


def geom_warp_matrices(nW, stepR, W, H, fx, fy, cx, cy, resize=None):
    """Image::Image (image.cpp:33-111) [+ Image::resizeKernel(nW', stepR')]: forward K R K^-1, float64,
    shape (nWz, nWy, nWx, 3, 3)."""
    nW = np.asarray(nW, np.int32)
    stepR = np.asarray(stepR, np.float32)
    n = np.asarray(resize[0], np.int32) if resize else nW
    out = np.zeros((int(n[2]), int(n[1]), int(n[0]), 3, 3), np.float64)
    rn = np.asarray(resize[0], np.int32) if resize else None
    rs = np.asarray(resize[1], np.float32) if resize else None
    got = _geom_lib().nmirefg_warp_matrices(_p(nW), _p(stepR), W, H, fx, fy, cx, cy, _p(rn) if resize else None,
                                            _p(rs) if resize else None, _p(out))
    assert got == out.size // 9
    return out


def geom_setup_cam(Twc):
    T = np.ascontiguousarray(Twc, np.float32).reshape(16)
    pos, d, up = (np.zeros(3, np.float32) for _ in range(3))
    _geom_lib().nmirefg_setup_cam(_p(T), _p(pos), _p(d), _p(up))
    return pos, d, up


def geom_cell_translation(Twc, nS, stepT, sx, sy, sz):
    """setupCam -> Rendering::setCamera -> calculateTranslation(sx, sy, sz) (Tracking.cc:1873-1882)."""
    T = np.ascontiguousarray(Twc, np.float32).reshape(16)
    nS = np.asarray(nS, np.int32)
    st = np.asarray(stepT, np.float32)
    t = np.zeros(3, np.float32)
    _geom_lib().nmirefg_cell_translation(_p(T), _p(nS), _p(st), sx, sy, sz, _p(t))
    return t


def geom_apply_winner(Twc, nS, nW, stepT, stepR, s, w):
    """Tracking::CalculateNMIRelocalization (Tracking.cc:2374-2419)."""
    T = np.ascontiguousarray(Twc, np.float32).reshape(16)
    a = [np.asarray(nS, np.int32), np.asarray(nW, np.int32), np.asarray(stepT, np.float32), np.asarray(stepR, np.float32),
         np.asarray(s, np.int32), np.asarray(w, np.int32)]
    out = np.zeros(16, np.float32)
    _geom_lib().nmirefg_apply_winner(_p(T), *[_p(x) for x in a], _p(out))
    return out.reshape(4, 4)
