"""Multi-GPU pose-grid search: one process per GPU, torch.distributed for the plumbing.

The cloud and frame are replicated, the pose grid is sharded by nmi_partition (synthetic
views first, warps when there are fewer views than ranks), every rank scores its slice
with its own context, and the only exchange is ONE 8-byte max-allreduce of the packed
winner key  (bits(max(score,0)) << 32) | (0xFFFFFFFF - linear_index)  -- the unsigned max
of which is exactly helperFunctions::find_max_elements' rule (first index of the maximum)
across ranks.  Scores are >= 0 so the key is < 2^63 and an int64 MAX-reduce is the same
as the unsigned one (NCCL on GPU, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

from . import search as _search
from .capi import Flags, Grid


def pack_key(max_score: float, index: int) -> int:
    """Host mirror of csrc/argmax.cu's key (used by the CPU tests of the exchange step)."""
    m = np.float32(max(float(max_score), 0.0))
    low = 0 if index < 0 else 0xFFFFFFFF - int(index)
    return (int(m.view(np.uint32)) << 32) | low


def local_key_from_scores(scores: np.ndarray, indices: np.ndarray) -> int:
    """find_max_elements restricted to `indices` (the rank's slice): max from 0, strict >,
    lowest index among the equal maxima; low word 0 when nothing qualifies."""
    s = np.asarray(scores, dtype=np.float32)[indices]
    m = np.float32(0.0)
    if s.size:
        with np.errstate(invalid="ignore"):
            cand = s[s > m]
        if cand.size:
            m = cand.max()
    hits = indices[s == m]
    return pack_key(float(m), int(hits.min()) if hits.size else -1)


def shard_indices(grid: Grid, rank: int, world: int) -> np.ndarray:
    """Linear rating indices this rank evaluates (rating order wz,wy,wx,sz,sy,sx)."""
    axis, b, e = _search.partition(grid, rank, world)
    nS, nW = grid.n_synth, grid.n_warp
    s = np.arange(b, e) if axis == 0 else np.arange(nS)
    w = np.arange(nW) if axis == 0 else np.arange(b, e)
    return (w[:, None] * nS + s[None, :]).reshape(-1)


def allreduce_key(key_tensor):
    """In-place max-allreduce of the int64 key tensor over the default process group."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(key_tensor, op=dist.ReduceOp.MAX)
    return key_tensor


def sharded_search(searcher, Twc, grid: Grid, flags: Flags, key_tensor, rank: int, world: int,
                   stream=None):
    """Enqueue this rank's slice on the searcher's stream, then the 8-byte allreduce on the
    same stream (no host sync in between). Returns the decoded global winner."""
    import torch

    from . import capi

    for attempt in range(2):
        ctx = torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream())
        with ctx:
            searcher.search_enqueue(Twc, grid, flags, rank, world, key_tensor.data_ptr())
            allreduce_key(key_tensor)
        (stream or torch.cuda.current_stream()).synchronize()
        # nmi_read_key also notices a LOCAL bin overflow and arms the exact sizing for the retry
        key = searcher.read_key(key_tensor.data_ptr())
        if key != capi.NMI_KEY_RETRY:
            return searcher.decode(grid, key)
        # some rank's fixed-capacity splat bins filled up: every rank sees the same RETRY key and
        # redoes the level once, like csrc/driver.cpp sharded_level
    raise capi.NmiError(capi.NMI_ERR_RETRY, "splat record bins overflowed twice")


def relocalize_sharded(searcher, Twc, grid: Grid, flags: Flags | None, key_tensor, rank: int, world: int,
                       **kw):
    """BASELINE config 4: the coarse-to-fine driver with every level sharded over the ranks of
    the default process group (csrc/driver.cpp nmi_relocalize_sharded).  The exchange step is one
    8-byte MAX all_reduce of `key_tensor` (int64, on the searcher's GPU) per level, enqueued by
    NCCL on the searcher's own stream right behind the argmax kernel."""
    import torch

    assert key_tensor.is_cuda and key_tensor.numel() == 1 and key_tensor.dtype == torch.int64

    def exchange(key_dev: int, stream: int):
        assert key_dev == key_tensor.data_ptr()
        with torch.cuda.stream(torch.cuda.ExternalStream(stream, device=key_tensor.device)):
            allreduce_key(key_tensor)

    return searcher.relocalize_sharded(Twc, grid, flags, rank, world, key_tensor.data_ptr(), exchange, **kw)


def relocalize_sharded_host(level_scores, Twc, grid: Grid, rank: int, world: int, reduce_max, **kw):
    """The same driver with the per-rank scoring supplied by the caller -- the host-side logic
    of the multi-GPU level driver without a GPU (CPU tests: gloo, world_size 2).
    level_scores(Twc, grid, indices) -> scores of this rank's linear indices;
    reduce_max(int key) -> int is the exchange step."""

    def level(T, g):
        idx = shard_indices(g, rank, world)
        full = np.full(g.n_pose, -1.0, dtype=np.float32)
        full[idx] = np.asarray(level_scores(T, g, idx), dtype=np.float32)
        key = int(reduce_max(local_key_from_scores(full, idx)))
        r = _search.decode_key(g, key)
        if r.best_index < 0:
            return 4  # NMI_ERR_NO_WINNER
        return r.best_s, r.best_w, r.best_score

    return _search.relocalize_with(level, Twc, grid, **kw)
