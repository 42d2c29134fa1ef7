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
