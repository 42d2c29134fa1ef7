"""Host-side geometry pinned on the reference's OWN lines (VERDICT r1 next #5 ii).

oracle/_ref/libnmi_ref_geom.so (oracle/Makefile.ref, built where /root/reference exists; CPU only)
holds, compiled unmodified by g++:
  * Thirdparty/Localization/image.cpp, whole        -> Image::Image / resizeKernel warp matrices (a5)
  * Thirdparty/Localization/ioData.cpp:177-197      -> setupCam                                   (a2)
  * Thirdparty/Localization/rendering.hpp:642-694   -> calculateTranslation / calculateTranslationCV (a3)
  * src/Tracking.cc:2374-2419                       -> Tracking::CalculateNMIRelocalization       (a14)
behind oracle/ref_shim/geom/{cvshim,glmshim}.h, which restate the few OpenCV 3.4 / GLM 0.9.7.1
operations those lines use (neither library is in the reference tree).  Everything is compared BIT FOR
BIT with the product's C ABI (host code of libnmi_b200.so, no GPU) and with the oracle.
The shim's two arithmetic rules (Mat * Mat summed left to right, 3x3 inverse by cofactors / det) are in
turn held against the real OpenCV through Python cv2 below.
"""
import numpy as np
import pytest

from orbslam2_nmi_b200 import search, synth
from orbslam2_nmi_b200.capi import Camera, Grid


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_py

    if not ref_py.geom_available():
        pytest.skip("oracle/_ref/libnmi_ref_geom.so not built (python -m orbslam2_nmi_b200.build where /root/reference exists)")
    return ref_py


def _poses(n, seed=0, spread=200.0):
    rng = np.random.default_rng(seed)
    for _ in range(n):
        T = synth.prior_pose(height_above=5 + 25 * rng.random(), tilt_deg=60 * rng.random() - 30)
        a = rng.random() * 2 * np.pi
        Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
        T[:3, :3] = (Rz @ T[:3, :3].astype(np.float64)).astype(np.float32)
        T[:3, 3] += (rng.random(3) * spread - spread / 2).astype(np.float32)
        yield T, rng


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


@pytest.mark.parametrize("nW,stepR", [((3, 3, 3), (0.02, 0.02, 0.05)), ((4, 4, 4), (0.02, 0.02, 0.05)),
                                      ((5, 2, 1), (0.013, 0.07, 0.001)), ((1, 1, 7), (0.5, 0.5, 0.004))])
def test_warp_matrices_are_the_references(ref, oracle, nW, stepR):
    """image.cpp:76-108: theta start -(n-1)/2*step with INTEGER division, += step in double, Rz*Ry*Rx,
    K*R*K.inv(): the forward matrices of every cell, and the inverse map the product hands to its warp."""
    sc = synth.make_scene("C2", n_points=10)
    g = Grid.make((1, 1, 1), nW, (0.2, 0.2, 0.5), stepR)
    cam = Camera(sc.W, sc.H, sc.fx, sc.fy, sc.cx, sc.cy, sc.zn, sc.zf, 3.0)
    M = ref.geom_warp_matrices(nW, stepR, sc.W, sc.H, sc.fx, sc.fy, sc.cx, sc.cy)
    for z in range(nW[2]):
        for y in range(nW[1]):
            for x in range(nW[0]):
                assert np.array_equal(_bits(oracle.cell_homography(sc, g, x, y, z)), _bits(M[z, y, x]))
                # cv::cuda::warpPerspective inverts the forward matrix itself (cv::invert, double) and
                # narrows to float: what nmi_cell_homography_inv returns
                want = np.linalg.inv(M[z, y, x])
                got = search.cell_homography_inv(cam, g, x, y, z).reshape(3, 3)
                assert np.allclose(got, want, rtol=1e-6, atol=1e-9)
                assert np.array_equal(_bits(got), _bits(oracle.cell_homography_inv(sc, g, x, y, z).reshape(3, 3)))


def test_warp_matrices_after_resize(ref, oracle):
    """Image::resizeKernel (image.cpp:236-268) rebuilds the matrices for the halved steps."""
    sc = synth.make_scene("C2", n_points=10)
    M = ref.geom_warp_matrices((3, 3, 3), (0.02, 0.02, 0.05), sc.W, sc.H, sc.fx, sc.fy, sc.cx, sc.cy,
                               resize=((3, 1, 3), (0.01, 0.02, 0.025)))
    g = Grid.make((1, 1, 1), (3, 1, 3), (0.2, 0.2, 0.5), (0.01, 0.02, 0.025))
    assert M.shape[:3] == (3, 1, 3)
    for z in range(3):
        for x in range(3):
            assert np.array_equal(_bits(oracle.cell_homography(sc, g, x, 0, z)), _bits(M[z, 0, x]))


def test_setup_cam_is_the_references(ref):
    for T, _ in _poses(20, seed=3):
        pos, d, up = ref.geom_setup_cam(T)
        assert np.array_equal(pos, T[:3, 3]) and np.array_equal(up, T[:3, 1])
        assert np.array_equal(d, (T[:3, 2] + T[:3, 3]).astype(np.float32))  # float adds


def test_cell_translation_is_the_references(ref, oracle):
    """setupCam -> setCamera -> calculateTranslation with glm::rotate (rendering.hpp:644-665): product
    (nmi_cell_translation) == oracle == the reference's lines, bit for bit, even and odd grids, camera far
    from the origin (where `dir = pos + z` rounds coarsely)."""
    n = 0
    for T, rng in _poses(150, seed=1, spread=2000.0):
        nS = tuple(int(v) for v in rng.integers(1, 7, 3))
        stepT = tuple(float(v) for v in (0.05 + rng.random(3)).astype(np.float32))
        g = Grid.make(nS, (1, 1, 1), stepT, (0.02, 0.02, 0.05))
        for _ in range(4):
            s = tuple(int(rng.integers(0, nS[k])) for k in range(3))
            r = ref.geom_cell_translation(T, nS, stepT, *s)
            assert np.array_equal(_bits(r), _bits(oracle.cell_translation(T, g, *s))), (T, nS, s)
            assert np.array_equal(_bits(r), _bits(search.cell_translation(T, g, *s))), (T, nS, s)
            n += 1
    assert n == 600


def test_apply_winner_is_the_references(ref, oracle):
    """Tracking::CalculateNMIRelocalization (Tracking.cc:2374-2419): rot_k = (best - n/2) * step with the
    INTEGER n/2, R = Rz*Ry*Rx in float, newLoc = Twc * [R | 0] + calculateTranslationCV(best)."""
    for T, rng in _poses(150, seed=2):
        nS = tuple(int(v) for v in rng.integers(1, 6, 3))
        nW = tuple(int(v) for v in rng.integers(1, 6, 3))
        stepT, stepR = (0.2, 0.17, 0.5), (0.02, 0.031, 0.05)
        g = Grid.make(nS, nW, stepT, stepR)
        s = tuple(int(rng.integers(0, nS[k])) for k in range(3))
        w = tuple(int(rng.integers(0, nW[k])) for k in range(3))
        r = ref.geom_apply_winner(T, nS, nW, stepT, stepR, s, w)
        assert np.array_equal(_bits(r), _bits(oracle.apply_winner(T, g, s, w)))
        assert np.array_equal(_bits(r), _bits(search.apply_winner(T, g, s, w)))


def test_shim_arithmetic_is_opencvs(ref):
    """The shim's gemm / invert rules against the real OpenCV (Python cv2): K*R*K.inv() formed with
    cv2.gemm / cv2.invert from the same R gives the matrices the shim-compiled image.cpp produced."""
    cv2 = pytest.importorskip("cv2")
    sc = synth.make_scene("C2", n_points=10)
    nW, stepR = (3, 2, 3), (0.02, 0.03, 0.05)
    M = ref.geom_warp_matrices(nW, stepR, sc.W, sc.H, sc.fx, sc.fy, sc.cx, sc.cy)
    K = np.array([[sc.fx, 0, sc.cx], [0, sc.fy, sc.cy], [0, 0, 1]], np.float64)
    _, Ki = cv2.invert(K)
    for z in range(nW[2]):
        tz = float(np.float32(-((nW[2] - 1) // 2)) * np.float32(stepR[2])) + z * float(np.float32(stepR[2]))
        for y in range(nW[1]):
            ty = float(np.float32(-((nW[1] - 1) // 2)) * np.float32(stepR[1])) + y * float(np.float32(stepR[1]))
            for x in range(nW[0]):
                tx = float(np.float32(-((nW[0] - 1) // 2)) * np.float32(stepR[0])) + x * float(np.float32(stepR[0]))
                Rz = np.array([[np.cos(tz), -np.sin(tz), 0], [np.sin(tz), np.cos(tz), 0], [0, 0, 1]])
                Ry = np.array([[np.cos(ty), 0, np.sin(ty)], [0, 1, 0], [-np.sin(ty), 0, np.cos(ty)]])
                Rx = np.array([[1, 0, 0], [0, np.cos(tx), -np.sin(tx)], [0, np.sin(tx), np.cos(tx)]])
                R = cv2.gemm(cv2.gemm(Rz, Ry, 1, None, 0), Rx, 1, None, 0)
                want = cv2.gemm(cv2.gemm(K, R, 1, None, 0), Ki, 1, None, 0)
                assert np.array_equal(_bits(want), _bits(M[z, y, x])), (x, y, z)
