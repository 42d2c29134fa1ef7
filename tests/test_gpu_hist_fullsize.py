"""Full-size (1920x1080) worst cases of the packed-u16 joint histogram, and the integers of the
code path the benchmark times.

VERDICT r1 weak #2 / #3: (i) nmi_get_hist used to fall back to the warp-per-row epilogue, so the
integer histograms were asserted for another code path than the timed one -- path=1 keeps the
persistent kernel's fast epilogue (J is copied out before it runs; HA, HB and the score are its
own).  (ii) the 16-bit-counter crossing / repay scheme was only stressed at 320x200 -- here a
1920x1080 single-bin pair (506 crossings of one counter), sky x sky and sparse-render pairs run
through the BATCHED launch of a grid search (nmi_score_pairs) with hot-bin skipping 0 / 1 / 2 and
are compared with the reference's own kernels (oracle/_ref) and the oracle.
"""
import numpy as np
import pytest

from orbslam2_nmi_b200 import synth
from orbslam2_nmi_b200.capi import Grid

pytestmark = pytest.mark.gpu
W, H = 1920, 1080


@pytest.fixture(scope="module")
def searcher(nmi_lib):
    from orbslam2_nmi_b200.search import NmiSearcher

    s = NmiSearcher(0)
    yield s
    s.close()


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_py

    if not ref_py.available():
        pytest.skip("oracle/_ref/libnmi_ref.so not built")
    ref_py.load()
    return ref_py


def _stacks():
    rng = np.random.default_rng(77)
    tex = synth.frame_textured(W, H, seed=5)
    sky_render = synth.frame_textured(W, H, seed=6)
    sky_render[: int(0.4 * H)] = 255                      # render background over 40 % of the image
    sparse = np.full((H, W), 255, np.uint8)                # sparse cloud: 3 % of the pixels covered
    m = rng.random((H, W)) < 0.03
    sparse[m] = rng.integers(0, 255, int(m.sum()), dtype=np.uint8)
    renders = [np.full((H, W), 255, np.uint8), np.full((H, W), 7, np.uint8), sky_render, sparse, tex]
    border = tex.copy()
    border[:, :400] = 0                                    # warp border: 21 % zeros
    warps = [np.full((H, W), 128, np.uint8), synth.frame_sky(W, H), border,
             rng.integers(0, 256, (H, W), dtype=np.uint8), np.full((H, W), 255, np.uint8)]
    return renders, warps


@pytest.mark.parametrize("skip", [0, 1, 2])
def test_full_size_worst_case_counters(searcher, oracle, ref, skip):
    import torch

    renders, warps = _stacks()
    n_r, n_w = len(renders), len(warps)
    searcher.set_camera(W, H, 870.0, 870.0, W / 2, H / 2, 5.0, 30.0)
    d_r = torch.from_numpy(np.stack(renders)).cuda()
    d_w = torch.from_numpy(np.stack(warps)).cuda()
    torch.cuda.synchronize()
    want = np.zeros((n_w, n_r), np.float32)
    hists = {}
    for wi in range(n_w):
        for ri in range(n_r):
            J, HA, HB = oracle.joint_hist(renders[ri], warps[wi])
            hists[(ri, wi)] = (J, HA, HB)
            want[wi, ri] = oracle.score_stages_f32(J, HA, HB, W * H)["score"]
    assert hists[(0, 0)][0].max() == W * H                 # one counter takes every pixel: 506 crossings
    searcher.set_hist_skip(skip)
    try:
        for rep in range(2):                                # mode 1 decides from the previous launch's levels
            got = searcher.score_pairs(d_r.data_ptr(), n_r, W * H, d_w.data_ptr(), n_w, W * H)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), \
                f"skip {skip} rep {rep}: scores differ from the oracle's at {np.argwhere(got != want)[:4].tolist()}"
        if skip == 0:
            assert searcher.last_hist_path() == 1
        if skip == 2:
            assert searcher.last_hist_path() == 2
        # integer histograms of the nastiest pairs against the REFERENCE's kernels, through every build
        for (ri, wi) in [(0, 0), (0, 4), (2, 1), (3, 2), (4, 3)]:
            st = ref.stages(renders[ri], warps[wi])
            J, HA, HB = hists[(ri, wi)]
            assert np.array_equal(st["J"], J) and np.array_equal(st["HA"], HA) and np.array_equal(st["HB"], HB)
            for path in (1, 2, 0):
                gJ, gHA, gHB, gs = searcher.get_hist(ri, wi, path=path)
                assert np.array_equal(gJ, st["J"]), f"pair {ri},{wi} path {path}: joint histogram"
                assert np.array_equal(gHA, st["HA"]) and np.array_equal(gHB, st["HB"]), f"pair {ri},{wi} path {path}"
                assert np.float32(gs) == want[wi, ri]
                ref_score = oracle.finish_f32(*[float(x) for x in st["sums"]])
                assert np.float32(gs) == np.float32(ref_score), "score bits differ from the reference kernels' totals"
    finally:
        searcher.set_hist_skip(1)


@pytest.mark.parametrize("variant", [0, 9, 10])
def test_timed_path_integers(searcher, oracle, variant):
    """HA / HB / score read back through path=1 come out of rows_epilogue_fast (variants 0, 9, 10 run
    it), J out of the very words it reads -- after a real search, at a ragged size with repaid
    crossings (a sparse render keeps ~90 % of the pixels in row 255)."""
    sc = synth.make_scene("small", n_points=4000)
    g = Grid.make((2, 2, 1), (2, 1, 2), (0.2, 0.2, 0.5), (0.02, 0.02, 0.05))
    frame = synth.frame_sky(sc.W, sc.H)
    searcher.set_scene(sc)
    searcher.set_frame(frame)
    searcher.set_hist_skip(0)
    try:
        fl = searcher.flags(variant=variant)
        res = searcher.search(sc.Twc, g, fl, want_scores=True)
        assert searcher.last_hist_path() == 1
        scores, renders, warps = oracle.search_points(sc, sc.Twc, g, sc.xyzi, frame, keep_images=True)
        crossed = 0
        for s in range(g.n_synth):
            for w in range(g.n_warp):
                J, HA, HB = oracle.joint_hist(renders[s], warps[w])
                crossed += int((J >= 4096).sum())
                gJ, gHA, gHB, gs = searcher.get_hist(s, w, fl, path=1)
                assert np.array_equal(gJ, J) and np.array_equal(gHA, HA) and np.array_equal(gHB, HB)
                assert np.float32(gs) == np.float32(res.scores[w * g.n_synth + s]) == np.float32(scores[w * g.n_synth + s])
        assert crossed > 0  # the repaid-crossing rows (warp-per-row fallback inside the fast epilogue) were exercised
    finally:
        searcher.set_hist_skip(1)
