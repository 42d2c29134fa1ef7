"""world_size-2 (and 3) gloo test of the multi-rank exchange step on CPU.

Each rank scores its nmi_partition slice with the CPU oracle (standing in for its GPU),
packs the local winner key, and the ranks combine with one int64 MAX all-reduce -- the same
collective bench.py runs over NCCL.  The decoded winner must equal the single-process
find_max_elements answer, including ties and the all-zero case.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from orbslam2_nmi_b200 import multigpu, search, synth
from orbslam2_nmi_b200.capi import Grid


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle_py as oracle

        sc = synth.make_scene("tiny", n_points=20000)
        if case == "views":
            g = Grid.make((3, 2, 1), (2, 1, 1), (0.2, 0.2, 0.5), (0.02, 0.02, 0.05))
        else:  # fewer synthetic views than ranks: the warp axis is sharded
            g = Grid.make((1, 1, 1), (3, 2, 1), (0.2, 0.2, 0.5), (0.02, 0.02, 0.05))
        frame = synth.frame_textured(sc.W, sc.H, seed=4)
        idx = multigpu.shard_indices(g, rank, world)
        # the oracle scores the full grid; a rank only LOOKS at its own slice
        scores, _, _ = oracle.search_points(sc, sc.Twc, g, sc.xyzi, frame, threads=1)
        if case == "tie":
            scores[:] = 0.0
        local = np.full(g.n_pose, np.nan, dtype=np.float32)
        local[idx] = scores[idx]
        key = torch.tensor([multigpu.local_key_from_scores(local, idx)], dtype=torch.int64)
        multigpu.allreduce_key(key)
        got = search.decode_key(g, int(key.item()))
        want, wmax = oracle.argmax(scores)
        out_q.put((rank, got.best_index, got.best_score, want, wmax, sorted(idx.tolist())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,case", [(2, "views"), (3, "views"), (2, "warps"), (2, "tie")])
def test_key_allreduce_matches_single_process(nmi_lib, oracle, world, case):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = []
    for rank, got_idx, got_score, want, wmax, idx in res:
        assert got_idx == want, f"rank {rank}: {got_idx} != {want}"
        assert got_score == np.float32(wmax)
        covered += idx
    n = len(covered)
    assert sorted(covered) == list(range(n))  # slices tile the grid exactly once


# ------------------------------------------------- sharded coarse-to-fine driver (C4) ----
def _reloc_worker(rank, world, port, threshold, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle_py as oracle

        sc = synth.make_scene("tiny", n_points=20000)
        g0 = Grid.make((3, 3, 1), (3, 1, 1), (0.4, 0.4, 0.5), (0.04, 0.02, 0.05))
        t = oracle.cell_translation(sc.Twc, g0, 2, 0, 0)
        _, img = oracle.render_points(sc, sc.Twc, t, sc.xyzi)
        frame = synth.frame_from_render(img, seed=3)
        exchanges = []

        def level_scores(T, g, idx):  # the oracle stands in for this rank's GPU
            scores, _, _ = oracle.search_points(sc, T, g, sc.xyzi, frame, threads=1)
            return scores[idx]

        def reduce_max(key):
            k = torch.tensor([key], dtype=torch.int64)
            multigpu.allreduce_key(k)
            exchanges.append(int(k.item()))
            return int(k.item())

        got = multigpu.relocalize_sharded_host(level_scores, sc.Twc, g0, rank, world, reduce_max,
                                               threshold=threshold)
        rc, want = oracle.relocalize_points(sc, sc.Twc, g0, sc.xyzi, frame, threshold, threads=1)
        out_q.put((rank, rc, len(exchanges),
                   (got.iterations, got.relocalized, got.failed, got.n_evals),
                   (want.iterations, want.relocalized, want.failed, want.n_evals),
                   list(got.Twc[:]), list(want.Twc[:]), list(got.best_s) + list(got.best_w),
                   list(want.best_s) + list(want.best_w), list(got.final_grid.stepT), list(want.final_grid.stepT)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,threshold", [(2, 0.05), (3, 0.05), (2, 0.9)])
def test_sharded_level_driver_matches_single_process(nmi_lib, oracle, world, threshold):
    """nmi_relocalize_with (the driver nmi_relocalize_sharded runs on every rank) with each level
    scored in nmi_partition slices and combined by one int64 MAX all-reduce per level: every rank
    takes the same decisions as the oracle's single-process RelocalizeWithNMIStrategy."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_reloc_worker, args=(r, world, port, threshold, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, rc, n_ex, got, want, gT, wT, gb, wb, gstep, wstep in res:
        assert rc == 0
        assert got == want, f"rank {rank}"
        assert n_ex == got[0]          # exactly one 8-byte exchange per level
        assert gT == wT and gb == wb and gstep == wstep
    assert len({tuple(r[5]) for r in res}) == 1  # all ranks end on the same pose


def test_level_driver_callback_errors_surface(nmi_lib):
    sc = synth.make_scene("tiny", n_points=100)
    g0 = Grid.make((3, 1, 1), (1, 1, 1), (0.4, 0.4, 0.5), (0.04, 0.02, 0.05))
    from orbslam2_nmi_b200.capi import NmiError

    with pytest.raises(NmiError) as e:
        search.relocalize_with(lambda T, g: 4, sc.Twc, g0)   # NMI_ERR_NO_WINNER aborts the driver
    assert e.value.code == 4
    with pytest.raises(ZeroDivisionError):
        search.relocalize_with(lambda T, g: 1 // 0, sc.Twc, g0)
    # the retry key decodes to NMI_ERR_RETRY, never to a pose
    from orbslam2_nmi_b200 import capi
    import ctypes as C
    r = capi.Result()
    assert capi.load().nmi_decode_key(C.byref(g0), C.c_uint64(capi.NMI_KEY_RETRY), C.byref(r)) == capi.NMI_ERR_RETRY
    assert r.best_index == -1
