"""Known-answer and property tests that pin the CPU oracle (no GPU).

The reference ships no tests or golden vectors, so the oracle is
pinned by (i) hand-computable cases, (ii) algebraic properties of the reference's
formulas (NMI.cu:342-362), (iii) an independent numpy restatement of the integer
stages, and (iv) the committed fixtures in tests/golden/ (regression pin).
"""
import json
import math
from pathlib import Path

import numpy as np
import pytest

from orbslam2_nmi_b200 import synth
from orbslam2_nmi_b200.capi import Grid

GOLDEN = Path(__file__).parent / "golden"


# ---------------------------------------------------------------- histogram -----
def np_joint(render, warped, bins=256, bg=True):
    a = render.reshape(-1).astype(np.int64)
    b = warped.reshape(-1).astype(np.int64)
    if not bg:
        keep = (a != 0) & (b != 0)
        a, b = a[keep], b[keep]
    sh = 2 if bins == 64 else 0
    a >>= sh
    b >>= sh
    J = np.bincount(a * bins + b, minlength=bins * bins).reshape(bins, bins).astype(np.uint32)
    return J, J.sum(1).astype(np.uint32), J.sum(0).astype(np.uint32)


def test_hist_hand_2x2(oracle):
    r = np.array([[0, 1], [1, 255]], dtype=np.uint8)
    w = np.array([[3, 3], [4, 0]], dtype=np.uint8)
    J, HA, HB = oracle.joint_hist(r, w)
    assert J[0, 3] == 1 and J[1, 3] == 1 and J[1, 4] == 1 and J[255, 0] == 1 and J.sum() == 4
    assert HA[0] == 1 and HA[1] == 2 and HA[255] == 1 and HB[3] == 2 and HB[4] == 1 and HB[0] == 1
    # joint index is render*256 + camera (NMI.cu:48)
    assert J.reshape(-1)[1 * 256 + 4] == 1
    # BG off (NMI.cu:85): pixels with render==0 or camera==0 are skipped
    J, HA, HB = oracle.joint_hist(r, w, bg=False)
    assert J.sum() == 2 and J[1, 3] == 1 and J[1, 4] == 1


@pytest.mark.parametrize("bins", [256, 64])
@pytest.mark.parametrize("bg", [True, False])
def test_hist_matches_numpy(oracle, bins, bg):
    rng = np.random.default_rng(3)
    r = rng.integers(0, 256, (37, 53), dtype=np.uint8)
    w = rng.integers(0, 256, (37, 53), dtype=np.uint8)
    r[::5, ::3] = 0
    w[1::7, ::2] = 0
    J, HA, HB = oracle.joint_hist(r, w, bins=bins, bg=bg)
    J2, HA2, HB2 = np_joint(r, w, bins, bg)
    assert np.array_equal(J, J2) and np.array_equal(HA, HA2) and np.array_equal(HB, HB2)
    assert np.array_equal(HA, J.sum(1)) and np.array_equal(HB, J.sum(0))


# -------------------------------------------------------------------- score -----
def test_score_hand_values(oracle):
    # two equiprobable values, identical images: H(A)=H(B)=H(AB)=1 bit
    r = np.array([[10, 20], [10, 20]], dtype=np.uint8)
    assert oracle.eval_one(r, r, mode=oracle.SUC) == pytest.approx(1.0, abs=1e-7)
    assert oracle.eval_one(r, r, mode=oracle.ENMI) == pytest.approx(2.0, abs=1e-7)
    # independent: A in {10,20}, B in {1,2} all four combos once: H(A)=H(B)=1, H(AB)=2
    a = np.array([[10, 10], [20, 20]], dtype=np.uint8)
    b = np.array([[1, 2], [1, 2]], dtype=np.uint8)
    assert oracle.eval_one(a, b, mode=oracle.SUC) == pytest.approx(0.0, abs=1e-7)
    assert oracle.eval_one(a, b, mode=oracle.ENMI) == pytest.approx(1.0, abs=1e-7)
    # constant images: all entropies are 0 -> guarded 0 (NMI.cu:344,353)
    c = np.full((4, 4), 7, dtype=np.uint8)
    assert oracle.eval_one(c, c, mode=oracle.SUC) == 0.0
    assert oracle.eval_one(c, c, mode=oracle.ENMI) == 0.0


def test_score_exact_small_case(oracle):
    # 8 pixels: p = {1/2, 1/4, 1/8, 1/8}: every p*log2 p is exact in fp32
    r = np.array([[1, 1, 1, 1, 2, 2, 3, 4]], dtype=np.uint8)
    w = np.array([[9, 9, 9, 9, 9, 9, 9, 9]], dtype=np.uint8)
    J, HA, HB = oracle.joint_hist(r, w)
    ha = 0.5 * 1 + 0.25 * 2 + 0.125 * 3 * 2  # 1.75 bits
    # H(B)=0, H(AB)=H(A) -> SUC = 2(1 - 1.75/1.75) = 0, ENMI = 1
    assert oracle.score_f32(J, HA, HB, 8, oracle.SUC) == 0.0
    assert oracle.score_f32(J, HA, HB, 8, oracle.ENMI) == 1.0
    assert oracle.score_f64(J, HA, HB, 8, oracle.ENMI) == 1.0
    assert ha == 1.75


def test_score_length_is_always_wh(oracle):
    # kernel.cu:85: length = W*H even when BG pixels are skipped -> p's do not sum to 1
    r = np.array([[0, 0, 5, 6]], dtype=np.uint8)
    w = np.array([[1, 1, 5, 6]], dtype=np.uint8)
    J, HA, HB = oracle.joint_hist(r, w, bg=False)
    assert J.sum() == 2
    s = oracle.score_f32(J, HA, HB, 4, oracle.ENMI)
    # each of A, B, AB has two bins with p = 1/4: every sum is 2 * (1/4 * -2) = -1
    assert s == pytest.approx(2.0, abs=1e-7)


def test_suc_enmi_relation_and_f64(oracle):
    rng = np.random.default_rng(11)
    base = rng.integers(0, 256, (64, 80), dtype=np.uint8)
    noisy = np.clip(base.astype(int) + rng.integers(-20, 21, base.shape), 0, 255).astype(np.uint8)
    for bins in (256, 64):
        J, HA, HB = oracle.joint_hist(base, noisy, bins=bins)
        suc = oracle.score_f32(J, HA, HB, base.size, oracle.SUC)
        enmi = oracle.score_f32(J, HA, HB, base.size, oracle.ENMI)
        assert 0.0 < suc < 1.0 and 1.0 < enmi < 2.0
        assert suc == pytest.approx(2 * (1 - 1 / enmi), rel=1e-5)
        assert suc == pytest.approx(oracle.score_f64(J, HA, HB, base.size, oracle.SUC), rel=2e-5)
        # plain numpy float64 entropy as a third opinion
        p = J[J > 0] / base.size
        pa = HA[HA > 0] / base.size
        pb = HB[HB > 0] / base.size
        h = lambda q: -(q * np.log2(q)).sum()
        assert oracle.score_f64(J, HA, HB, base.size, oracle.ENMI) == pytest.approx(
            (h(pa) + h(pb)) / h(p), rel=1e-12)


def test_tree_order_is_reference_order(oracle):
    # The pairwise tree (NMI.cu:270-287) is NOT a left-to-right sum: build counts whose fp32
    # tree sum differs from the sequential sum and check the oracle follows the tree.
    rng = np.random.default_rng(5)
    J = rng.integers(0, 2000, (256, 256)).astype(np.uint32)
    HA = J.sum(1).astype(np.uint32)
    HB = J.sum(0).astype(np.uint32)
    L = int(J.sum())

    def term(c):
        c = np.asarray(c, dtype=np.uint32)
        p = c.astype(np.float32) / np.float32(L)
        with np.errstate(divide="ignore", invalid="ignore"):
            t = p * np.log2(p, dtype=np.float32)
        return np.where(c == 0, np.float32(0), t).astype(np.float32)

    def tree(x):
        x = x.astype(np.float32).copy()
        n = x.shape[-1] // 2
        while n >= 1:
            x[..., :n] = x[..., :n] + x[..., n:2 * n]
            n //= 2
        return x[..., 0]

    rows = tree(term(J))
    sab, sa, sb = tree(rows), tree(term(HA)), tree(term(HB))
    want = np.float32(2) * (np.float32(1) - (-sab) / ((-sa) + (-sb)))
    got = oracle.score_f32(J, HA, HB, L, oracle.SUC)
    # numpy's log2 and glibc's log2f may differ by an ulp per term; the tree order makes the
    # two agree to ~1e-6 relative, far tighter than a sequential sum would
    assert got == pytest.approx(float(want), rel=3e-6)


# ------------------------------------------------------------------- argmax -----
def test_argmax_rules(oracle):
    # strict > from 0, first index of ties (helperFunctions.cpp:50-103, Tracking.cc:1952)
    assert oracle.argmax(np.array([0.1, 0.5, 0.5, 0.2], np.float32)) == (1, 0.5)
    # nothing above 0: max stays 0 and the first exact 0 wins
    i, m = oracle.argmax(np.array([-0.1, 0.0, 0.0], np.float32))
    assert (i, m) == (1, 0.0)
    # all negative: the reference's vector is empty
    assert oracle.argmax(np.array([-0.1, -0.2], np.float32))[0] == -1
    # NaN never compares greater
    assert oracle.argmax(np.array([np.nan, 0.3, np.nan], np.float32))[0] == 1


def test_linear_index_order(oracle):
    g = Grid.make((3, 2, 4), (2, 3, 5), (0.1, 0.1, 0.1), (0.01, 0.01, 0.01))
    seen = []
    for wz in range(5):
        for wy in range(3):
            for wx in range(2):
                for sz in range(4):
                    for sy in range(2):
                        for sx in range(3):
                            l = oracle.load().orc_linear_index(oracle.grid(g), sx, sy, sz, wx, wy, wz)
                            seen.append(l)
                            assert oracle.unravel(g, l) == ((sx, sy, sz), (wx, wy, wz))
    assert seen == list(range(g.n_pose))  # loop order of find_max_elements == increasing l


# ----------------------------------------------------------- grid and warps -----
def test_cell_angles_integer_division(oracle):
    # image.cpp:77: start = -(n-1)/2*step with INTEGER division
    g = Grid.make((1, 1, 1), (4, 3, 5), (0, 0, 0), (0.02, 0.03, 0.05))
    ang = [oracle.cell_angles(g, i, 0, 0)[0] for i in range(4)]
    assert np.allclose(ang, np.array([-1, 0, 1, 2]) * np.float32(0.02), atol=1e-9)
    ang = [oracle.cell_angles(g, 0, i, 0)[1] for i in range(3)]
    assert np.allclose(ang, np.array([-1, 0, 1]) * np.float32(0.03), atol=1e-9)
    ang = [oracle.cell_angles(g, 0, 0, i)[2] for i in range(5)]
    assert np.allclose(ang, np.array([-2, -1, 0, 1, 2]) * np.float32(0.05), atol=1e-9)


def test_homography_centre_cell_is_identity(oracle):
    sc = synth.make_scene("tiny", n_points=10)
    g = synth.default_grid()
    m = oracle.cell_homography_inv(sc, g, 1, 1, 1)
    assert np.allclose(m.reshape(3, 3), np.eye(3), atol=1e-6)
    img = synth.frame_textured(sc.W, sc.H)
    assert np.array_equal(oracle.warp(img, np.eye(3, dtype=np.float32).reshape(-1)), img)
    # centre cell: exact identity coefficients are not guaranteed in fp32, but the image is
    assert np.array_equal(oracle.warp(img, m), img)


def test_homography_matches_numpy(oracle):
    sc = synth.make_scene("tiny", n_points=10)
    g = synth.default_grid()
    K = np.array([[sc.fx, 0, sc.cx], [0, sc.fy, sc.cy], [0, 0, 1]])
    for (ix, iy, iz) in [(0, 0, 0), (2, 1, 0), (1, 2, 2)]:
        tx, ty, tz = oracle.cell_angles(g, ix, iy, iz)
        Rx = np.array([[1, 0, 0], [0, math.cos(tx), -math.sin(tx)], [0, math.sin(tx), math.cos(tx)]])
        Ry = np.array([[math.cos(ty), 0, math.sin(ty)], [0, 1, 0], [-math.sin(ty), 0, math.cos(ty)]])
        Rz = np.array([[math.cos(tz), -math.sin(tz), 0], [math.sin(tz), math.cos(tz), 0], [0, 0, 1]])
        M = K @ (Rz @ Ry @ Rx) @ np.linalg.inv(K)  # image.cpp:103-104
        want = np.linalg.inv(M)
        got = oracle.cell_homography_inv(sc, g, ix, iy, iz).reshape(3, 3)
        assert np.allclose(got, want, rtol=1e-5, atol=1e-7)


def test_warp_translation_and_border(oracle):
    img = np.arange(12 * 10, dtype=np.uint8).reshape(10, 12)
    # dst(x,y) = src(x+2, y+1): integer shift, zeros where the source is outside
    m = np.array([1, 0, 2, 0, 1, 1, 0, 0, 1], dtype=np.float32)
    out = oracle.warp(img, m)
    assert np.array_equal(out[:9, :10], img[1:, 2:])
    assert (out[9, :] == 0).all() and (out[:, 10:] == 0).all()
    # half-pixel shift: average of two neighbours, round-half-even
    m = np.array([1, 0, 0.5, 0, 1, 0, 0, 0, 1], dtype=np.float32)
    out = oracle.warp(img, m)
    want = np.rint((img[:, :-1].astype(np.float64) + img[:, 1:]) / 2)
    assert np.array_equal(out[:, :-1], want.astype(np.uint8))
    assert np.array_equal(out[:, -1], np.rint(img[:, -1] / 2.0).astype(np.uint8))  # tap outside = 0


def test_warp_against_opencv_if_present(oracle):
    cv2 = pytest.importorskip("cv2")
    sc = synth.make_scene("tiny", n_points=10)
    g = synth.default_grid()
    img = synth.frame_textured(sc.W, sc.H)
    minv = oracle.cell_homography_inv(sc, g, 0, 2, 2)
    ours = oracle.warp(img, minv).astype(int)
    M = np.linalg.inv(minv.reshape(3, 3).astype(np.float64))
    ref = cv2.warpPerspective(img, M, (sc.W, sc.H), flags=cv2.INTER_LINEAR,
                              borderMode=cv2.BORDER_CONSTANT, borderValue=0).astype(int)
    # cv2 (CPU) interpolates in 5-bit fixed point: expect +-1 grey level, not bit equality
    # (away from the warped image border, where cv2 blends the constant border differently)
    d = np.abs(ours - ref)
    assert np.mean(d <= 1) > 0.99 and d.mean() < 0.3 and d.max() <= 4
    assert np.mean(ours == ref) > 0.7


# ------------------------------------------------------------------- render -----
def test_cell_translation_axes(oracle):
    sc = synth.make_scene("tiny", n_points=10)
    g = Grid.make((3, 3, 3), (1, 1, 1), (0.2, 0.3, 0.5), (0, 0, 0))
    T = sc.Twc
    assert np.allclose(oracle.cell_translation(T, g, 1, 1, 1), 0, atol=1e-7)
    # +x cell moves along -x_cam (camera-left), +y along +y_cam (down), +z along -z_cam (back)
    assert np.allclose(oracle.cell_translation(T, g, 2, 1, 1), -0.2 * T[:3, 0], atol=1e-6)
    assert np.allclose(oracle.cell_translation(T, g, 1, 2, 1), 0.3 * T[:3, 1], atol=1e-6)
    assert np.allclose(oracle.cell_translation(T, g, 1, 1, 2), -0.5 * T[:3, 2], atol=1e-6)
    # even counts use the half-integer offset (rendering.hpp:646)
    g2 = Grid.make((4, 1, 1), (1, 1, 1), (0.2, 0.2, 0.2), (0, 0, 0))
    assert np.allclose(oracle.cell_translation(T, g2, 0, 0, 0), 1.5 * 0.2 * T[:3, 0], atol=1e-6)


def _cam(W=64, H=48, f=60.0, cx=None, cy=None):
    class Cam:
        pass

    c = Cam()
    c.W, c.H, c.fx, c.fy = W, H, f, f
    c.cx, c.cy = (W / 2 if cx is None else cx), (H / 2 if cy is None else cy)
    c.zn, c.zf, c.point_size = 5.0, 30.0, 3.0
    return c


def test_render_single_point_splat(oracle):
    cam = _cam()
    T = np.eye(4, dtype=np.float32)  # camera at origin looking along +z
    pts = np.array([[0.05, 0.02, 10.0, 100 / 256.0]], dtype=np.float32)
    win, img = oracle.render_points(cam, T, np.zeros(3, np.float32), pts)
    ys, xs = np.nonzero(win != oracle.EMPTY)
    # centre (32.3, 24.1) -> 3x3 block around pixel (32, 24)
    assert sorted(set(xs)) == [31, 32, 33] and sorted(set(ys)) == [23, 24, 25] and len(xs) == 9
    assert (img[win != oracle.EMPTY] == 100).all() and (img[win == oracle.EMPTY] == 255).all()


def test_render_depth_and_ties(oracle):
    cam = _cam()
    T = np.eye(4, dtype=np.float32)
    pts = np.array([[0, 0, 12.0, 10 / 256.0], [0, 0, 8.0, 20 / 256.0], [0, 0, 8.0, 30 / 256.0],
                    [0, 0, 4.0, 40 / 256.0], [0, 0, 31.0, 50 / 256.0]], dtype=np.float32)
    win, img = oracle.render_points(cam, T, np.zeros(3, np.float32), pts)
    # nearest inside [zn, zf] wins; equal depth -> lower index (GL_LESS, in-order)
    assert win[24, 32] == 1 and img[24, 32] == 20
    # near/far clipping drops points 3 and 4 entirely
    assert set(np.unique(win)) == {1, oracle.EMPTY}


def test_render_principal_point_quirk(oracle):
    # the projection forces the principal point to the image centre and scales the focal
    # length by (W/2)/cx (rendering.hpp:196-202): a point on the optical axis lands at W/2
    cam = _cam(cx=20.0, cy=30.0)
    T = np.eye(4, dtype=np.float32)
    pts = np.array([[0, 0, 10.0, 0.5], [1.0, 0, 10.0, 0.25]], dtype=np.float32)
    win, _ = oracle.render_points(cam, T, np.zeros(3, np.float32), pts)
    assert win[24, 32] == 0
    # x = 1 m at 10 m: xw = 32 * (1 + (60/20) * 0.1) = 41.6
    ys, xs = np.nonzero(win == 1)
    assert sorted(set(xs)) == [40, 41, 42]


def test_render_centre_clip(oracle):
    # a point whose centre is outside the clip volume is dropped even if its 3x3 square
    # would touch the image (GL clips point primitives on the vertex)
    cam = _cam()
    T = np.eye(4, dtype=np.float32)
    x_edge = 10.0 * (cam.W / 2) / cam.fx  # nx == 1 exactly at this x for z = 10
    pts = np.array([[x_edge * 1.001, 0, 10.0, 0.5], [x_edge * 0.999, 0, 10.0, 0.5]], np.float32)
    win, _ = oracle.render_points(cam, T, np.zeros(3, np.float32), pts)
    assert not (win == 0).any() and (win == 1).any()


# ------------------------------------------------------------ pose and grid -----
def test_apply_winner_centre_is_identity(oracle):
    sc = synth.make_scene("tiny", n_points=10)
    g = synth.default_grid()
    out = oracle.apply_winner(sc.Twc, g, (1, 1, 1), (1, 1, 1))
    assert np.allclose(out, sc.Twc, atol=1e-6)
    # even warp counts: applied rotation uses n/2 (Tracking.cc:2383), evaluation uses (n-1)/2
    g4 = Grid.make((1, 1, 1), (4, 1, 1), (0.1, 0.1, 0.1), (0.02, 0.02, 0.02))
    out = oracle.apply_winner(sc.Twc, g4, (0, 0, 0), (2, 0, 0))
    assert np.allclose(out, sc.Twc, atol=1e-6)  # index 2 is the "middle" for n = 4


def test_resize_grid_rules(oracle):
    g = Grid.make((3, 3, 3), (3, 3, 3), (0.2, 0.2, 0.5), (0.02, 0.02, 0.05))
    # best on the periphery in sx and wz: those steps stay, the others halve
    o = oracle.resize_grid(g, (0, 1, 1), (1, 1, 2))
    assert [o.stepT[k] for k in range(3)] == pytest.approx([0.2, 0.1, 0.25])
    assert [o.stepR[k] for k in range(3)] == pytest.approx([0.01, 0.01, 0.05])
    # steps under 5 mm / 1 mrad collapse the axis (nmiSearchKernel.cpp:122-138)
    g2 = Grid.make((3, 3, 3), (3, 3, 3), (0.009, 0.2, 0.5), (0.0019, 0.02, 0.05))
    o = oracle.resize_grid(g2, (1, 1, 1), (1, 1, 1))
    assert list(o.nS) == [1, 3, 3] and list(o.nW) == [1, 3, 3]
    assert oracle.is_middle(g, (1, 1, 1), (1, 1, 1)) and not oracle.is_middle(g, (1, 1, 0), (1, 1, 1))
    # n == 1 axes always halve (the periphery test needs n > 1)
    g1 = Grid.make((1, 3, 3), (3, 3, 3), (0.2, 0.2, 0.5), (0.02, 0.02, 0.05))
    o = oracle.resize_grid(g1, (0, 0, 0), (0, 0, 0))
    assert o.stepT[0] == pytest.approx(0.1) and o.stepT[1] == pytest.approx(0.2)


# --------------------------------------------------- planted pose + goldens -----
def test_planted_pose_is_found(oracle):
    sc = synth.make_scene("tiny")
    g = Grid.make((3, 3, 1), (3, 1, 1), (0.3, 0.3, 0.5), (0.03, 0.02, 0.05))
    hidden_s, hidden_w = (2, 0, 0), (1, 0, 0)  # warp centre cell: frame = unrotated render
    t = oracle.cell_translation(sc.Twc, g, *hidden_s)
    _, img = oracle.render_points(sc, sc.Twc, t, sc.xyzi)
    frame = synth.frame_from_render(img)
    scores, _, _ = oracle.search_points(sc, sc.Twc, g, sc.xyzi, frame)
    best, _ = oracle.argmax(scores)
    assert oracle.unravel(g, best) == (hidden_s, hidden_w)


def test_golden_fixtures(oracle):
    """Regression pin: fixtures written by tests/golden/make_golden.py from this oracle."""
    data = json.loads((GOLDEN / "oracle_small.json").read_text())
    sc = synth.make_scene(data["config"], n_points=data["n_points"], seed=data["seed"])
    g = Grid.make(data["nS"], data["nW"], data["stepT"], data["stepR"])
    frame = synth.frame_textured(sc.W, sc.H, seed=data["frame_seed"])
    scores, renders, warps = oracle.search_points(sc, sc.Twc, g, sc.xyzi, frame, keep_images=True)
    import zlib

    assert [zlib.crc32(r.tobytes()) for r in renders] == data["render_crc32"]
    assert [zlib.crc32(w.tobytes()) for w in warps] == data["warp_crc32"]
    J, HA, HB = oracle.joint_hist(renders[0], warps[0])
    assert zlib.crc32(J.tobytes()) == data["joint_crc32_pair00"]
    assert np.allclose(scores, np.array(data["scores"], dtype=np.float32), rtol=1e-6, atol=1e-7)
    assert oracle.argmax(scores)[0] == data["argmax"]


# --------------------------------------------------------------- mesh raster -----
def _mesh_cam():
    cam = _cam(W=64, H=48, f=64.0)
    return cam, np.eye(4, dtype=np.float32), np.zeros(3, np.float32)


def _px_to_xyz(cam, px, py, z):
    """World point (camera at origin, identity pose) that lands on window position (px, py)."""
    return [(px - cam.W / 2) * z / (cam.fx * (cam.W / 2) / cam.cx), (py - cam.H / 2) * z / (cam.fy * (cam.H / 2) / cam.cy), z]


def test_mesh_front_face_and_cull(oracle):
    cam, T, t0 = _mesh_cam()
    z = 10.0
    a, b, c = _px_to_xyz(cam, 10, 10, z), _px_to_xyz(cam, 40, 10, z), _px_to_xyz(cam, 10, 40, z)
    verts = np.array([a + [0.25], b + [0.5], c + [0.75]], dtype=np.float32)
    # GL front face = counter-clockwise with y up = clockwise on a top-down image
    win_cw, img = oracle.render_mesh(cam, T, t0, verts, np.array([[0, 1, 2]], np.uint32))
    win_ccw, _ = oracle.render_mesh(cam, T, t0, verts, np.array([[0, 2, 1]], np.uint32))
    on_cw, on_ccw = (win_cw != oracle.EMPTY).sum(), (win_ccw != oracle.EMPTY).sum()
    assert (on_cw == 0) != (on_ccw == 0)  # exactly one winding is culled (rendering.hpp:300)
    win = win_cw if on_cw else win_ccw
    # right triangle with 30 px legs: ~450 pixel centres
    assert 400 < (win != oracle.EMPTY).sum() < 500
    assert win[15, 15] == 0 and win[45, 5] == oracle.EMPTY


def test_mesh_watertight_shared_edge(oracle):
    """Two triangles sharing an edge cover every pixel of the quad exactly once (top-left rule):
    the union has no hole, and swapping the draw order changes no pixel's coverage."""
    cam, T, t0 = _mesh_cam()
    z = 12.0
    rng = np.random.default_rng(0)
    for _ in range(20):
        x0, y0 = rng.uniform(5, 20, 2)
        x1, y1 = x0 + rng.uniform(8, 30), y0 + rng.uniform(8, 20)
        q = [_px_to_xyz(cam, x0, y0, z), _px_to_xyz(cam, x1, y0, z), _px_to_xyz(cam, x1, y1, z),
             _px_to_xyz(cam, x0, y1, z)]
        verts = np.array([p + [0.5] for p in q], dtype=np.float32)
        for tris in ([[0, 2, 1], [0, 3, 2]], [[0, 1, 2], [0, 2, 3]]):
            win, _ = oracle.render_mesh(cam, T, t0, verts, np.array(tris, np.uint32))
            if (win != oracle.EMPTY).any():
                break
        cover = win != oracle.EMPTY
        # pixel centres strictly inside the rectangle are all covered, exactly the snapped box
        ys, xs = np.nonzero(cover)
        assert cover[ys.min():ys.max() + 1, xs.min():xs.max() + 1].all()  # no hole along the diagonal
        sx0, sx1 = np.rint(x0 * 256), np.rint(x1 * 256)
        want_cols = [i for i in range(cam.W) if sx0 <= i * 256 + 128 < sx1]
        assert xs.min() == want_cols[0] and xs.max() == want_cols[-1]  # left edge in, right edge out
        both = np.array(tris, np.uint32)
        w_ab, _ = oracle.render_mesh(cam, T, t0, verts, both)
        w_ba, _ = oracle.render_mesh(cam, T, t0, verts, both[::-1].copy())
        assert np.array_equal(w_ab != oracle.EMPTY, w_ba != oracle.EMPTY)
        # each pixel belongs to exactly one triangle: drawing order only renames the winner
        assert np.array_equal(np.where(w_ab == oracle.EMPTY, 9, 1 - w_ab), np.where(w_ba == oracle.EMPTY, 9, w_ba))


def test_mesh_depth_ties_and_clip(oracle):
    cam, T, t0 = _mesh_cam()

    def tri(z, grey):
        return [_px_to_xyz(cam, 10, 10, z) + [grey], _px_to_xyz(cam, 10, 40, z) + [grey],
                _px_to_xyz(cam, 40, 10, z) + [grey]]

    verts = np.array(tri(12.0, 0.1) + tri(8.0, 0.2) + tri(8.0, 0.3) + tri(4.0, 0.4) + tri(31.0, 0.5), np.float32)
    tris = np.arange(15, dtype=np.uint32).reshape(5, 3)
    win, img = oracle.render_mesh(cam, T, t0, verts, tris)
    if not (win != oracle.EMPTY).any():
        tris = tris[:, [0, 2, 1]].copy()
        win, img = oracle.render_mesh(cam, T, t0, verts, tris)
    # nearer wins; equal depth -> lower triangle index; outside [zn, zf] dropped
    assert set(np.unique(win)) == {1, oracle.EMPTY}
    assert img[15, 15] == int(np.floor(255 * np.float32(0.2) + 0.5))
    # a triangle with ONE vertex beyond the far plane is CLIPPED like GL clips it (per fragment on the
    # interpolated depth): the part inside [zn, zf] is drawn, the rest is not
    full, _ = oracle.render_mesh(cam, T, t0, np.array(tri(12.0, 0.1), np.float32), tris[:1])
    verts2 = np.array(tri(12.0, 0.1), np.float32)
    verts2[2, :3] = _px_to_xyz(cam, 40, 10, 60.0)
    w2, _ = oracle.render_mesh(cam, T, t0, verts2, tris[:1])
    n_full, n_clip = int((full != oracle.EMPTY).sum()), int((w2 != oracle.EMPTY).sum())
    assert 0 < n_clip < n_full
    assert w2[35, 11] != oracle.EMPTY or w2[11, 11] != oracle.EMPTY   # near the two vertices at 12 m: kept
    assert w2[11, 38] == oracle.EMPTY                                  # next to the vertex at 60 m: beyond zf = 30 m
    # a triangle entirely beyond the far plane, or with a vertex that cannot be projected, draws nothing
    far = np.array(tri(31.0, 0.5), np.float32)
    assert not (oracle.render_mesh(cam, T, t0, far, tris[:1])[0] != oracle.EMPTY).any()


def test_mesh_synthetic_terrain_renders(oracle):
    sc = synth.make_scene("tiny", n_points=10)
    verts, tris = synth.make_mesh(120, 120, extent=24.0)
    win, img = oracle.render_mesh(sc, sc.Twc, np.zeros(3, np.float32), verts, tris)
    covered = (win != oracle.EMPTY).mean()
    assert covered > 0.95  # the terrain fills the view from 15 m up
    assert 20 < img[win != oracle.EMPTY].mean() < 235


# ------------------------------------------------------- textured mesh (A.4, shading) -----
def test_mesh_texture_shading_hand_values(oracle):
    """Per-fragment shading of Rendering<1> (ShadingWithTexture.fragmentshader:16 over a texture that
    loadBMP_custom uploads with its B,G,R bytes as r,g,b, texture.cpp:90)."""
    cam, T, t0 = _mesh_cam()
    quad = [_px_to_xyz(cam, 8, 4, 10.0) + [0.0], _px_to_xyz(cam, 8, 44, 10.0) + [0.0],
            _px_to_xyz(cam, 56, 4, 10.0) + [0.0], _px_to_xyz(cam, 56, 44, 10.0) + [0.0]]
    verts = np.array(quad, np.float32)
    tris = np.array([[0, 1, 2], [2, 1, 3]], np.uint32)
    win, _ = oracle.render_mesh(cam, T, t0, verts, tris)
    if not (win != oracle.EMPTY).any():
        tris = tris[:, [0, 2, 1]].copy()
    uv_of = {0: (0.0, 0.0), 1: (0.0, 1.0), 2: (1.0, 0.0), 3: (1.0, 1.0)}
    uv = np.array([[uv_of[int(k)] for k in t] for t in tris], np.float32)
    # (i) a constant texture: every covered pixel = floor(0.299 c0 + 0.587 c1 + 0.114 c2 + 0.5) on the FILE order
    tex = np.zeros((4, 4, 3), np.uint8)
    tex[...] = (255, 0, 0)
    win, img = oracle.render_mesh_tex(cam, T, t0, verts, tris, uv, tex)
    cov = win != oracle.EMPTY
    assert cov.sum() > 1800 and set(np.unique(img[cov])) == {76} and set(np.unique(img[~cov])) == {255}
    tex[...] = (0, 0, 255)
    assert set(np.unique(oracle.render_mesh_tex(cam, T, t0, verts, tris, uv, tex)[1][cov])) == {29}
    # (ii) a two-texel texture along u: black | white, bilinear + GL_REPEAT: the middle of the quad is
    # the texel border (grey ~128), a quarter in is a texel centre (pure), the quad's edges wrap to ~128
    tex2 = np.zeros((1, 2, 3), np.uint8)
    tex2[0, 1] = 255
    _, img2 = oracle.render_mesh_tex(cam, T, t0, verts, tris, uv, tex2)
    row = img2[24].astype(int)
    assert abs(row[32] - 128) <= 6 and row[20] <= 8 and row[44] >= 247   # pixel 20 / 44: u = 0.26 / 0.76
    # pixel 9: u = 1.5/48 -> x = -0.4375: texel -1 wraps to the white one, weight 0.4375 -> 111.6;
    # pixel 55: u = 47.5/48 -> x = 1.479: white texel blending into the wrapped black one -> 132.8
    assert abs(row[9] - 112) <= 1 and abs(row[55] - 133) <= 1
    # (iii) the flat-shaded entry point is untouched by the texture machinery
    w3, i3 = oracle.render_mesh(cam, T, t0, verts, tris)
    assert np.array_equal(w3, win) and set(np.unique(i3[cov])) == {0}


def test_mesh_texture_perspective_correct(oracle):
    """A quad tilted in depth: UV is interpolated perspective-correctly (1/Zc weights), so the texel
    boundary of a two-texel texture does NOT fall on the quad's screen-space midpoint."""
    cam, T, t0 = _mesh_cam()
    near, far = 6.0, 24.0
    quad = [_px_to_xyz(cam, 8, 4, near) + [0.0], _px_to_xyz(cam, 8, 44, near) + [0.0],
            _px_to_xyz(cam, 56, 4, far) + [0.0], _px_to_xyz(cam, 56, 44, far) + [0.0]]
    verts = np.array(quad, np.float32)
    tris = np.array([[0, 1, 2], [2, 1, 3]], np.uint32)
    uv_of = {0: (0.0, 0.0), 1: (0.0, 1.0), 2: (1.0, 0.0), 3: (1.0, 1.0)}
    for flip in (False, True):
        tt = tris[:, [0, 2, 1]].copy() if flip else tris
        uv = np.array([[uv_of[int(k)] for k in t] for t in tt], np.float32)
        tex2 = np.zeros((1, 2, 3), np.uint8)
        tex2[0, 1] = 255
        win, img = oracle.render_mesh_tex(cam, T, t0, verts, tt, uv, tex2)
        if (win != oracle.EMPTY).any():
            break
    row = img[24].astype(int)
    cross = int(np.argmax(row[10:54] >= 128)) + 10      # first pixel past the u = 0.5 boundary
    # u(s) = (s / far) / ((1 - s) / near + s / far) at screen fraction s -> u = 0.5 at s = far / (near + far)
    want = 8 + (56 - 8) * (far / (near + far))
    assert abs(cross - want) <= 2 and abs(cross - 32) >= 8
