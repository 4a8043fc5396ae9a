"""Multi-GPU execution of the batched search: one process per GPU, ``torch.distributed`` plumbing.

The reference's only parallelism on this path is a static partition of the query batch over a
thread pool (``VamanaIndex::search``, /root/reference/include/svs/index/vamana/index.h:571-574 with
``threads::balance``, lib/threads/types.h:311-329).  Across GPUs the same partition applies:

* **Mode A -- replicas** (:class:`ReplicatedSearch`): every rank holds the whole index; rank ``r``
  searches the ``balance(nq, world, r)`` slice of the queries; the only communication is the final
  gather of ``(ids, distances)`` rows (NCCL all-gather over NVLink).  No reduction, no collective on
  the data path.
* **Mode B -- sharded index** (:class:`ShardedSearch`): every rank holds a contiguous id range of the
  base vectors with its own graph and entry point, searches *all* queries on it, adds its id offset,
  all-gathers the per-shard top-k and merges ``world x k -> k`` per query with the reference's
  ``TotalOrder`` (distance, then id; lib/neighbor.h:143-155) so the result is deterministic.

The classes are device-agnostic in their plumbing (CPU tensors + ``gloo`` work, which is how the
host logic is tested without GPUs); the local search and the merge are injected callables whose
defaults are the CUDA paths of ``libsvsb200.so``.
"""
from __future__ import annotations

from typing import Callable

import numpy as np
import torch
import torch.distributed as dist


def balance(n: int, nparts: int, i: int) -> tuple[int, int]:
    """``threads::balance`` (lib/threads/types.h:311-329): contiguous ranges whose sizes differ by <= 1,
    the first ``n % nparts`` ranges being one longer."""
    base, rem = divmod(n, nparts)
    start = i * base + min(i, rem)
    return start, start + base + (1 if i < rem else 0)


def _world(group) -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def _gather_rows(local: torch.Tensor, counts: list[int], group) -> torch.Tensor:
    """All-gather row blocks of unequal length (sizes known from ``balance``) into one tensor."""
    rank, world = _world(group)
    if world == 1:
        return local
    width = max(counts)
    padded = local
    if local.shape[0] < width:
        pad = torch.zeros((width - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], dim=0)
    out = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    if all(c == width for c in counts):
        return out
    return torch.cat([out[r * width: r * width + counts[r]] for r in range(world)], dim=0)


class ReplicatedSearch:
    """Mode A: replicate the index, shard the queries, gather the result rows."""

    def __init__(self, local_search: Callable[[torch.Tensor, int], tuple[torch.Tensor, torch.Tensor]], group=None):
        """``local_search(queries, k) -> (ids int64 [m,k], dists float32 [m,k])`` on this rank's device."""
        self.local_search = local_search
        self.group = group

    def search(self, queries: torch.Tensor, k: int) -> tuple[torch.Tensor, torch.Tensor]:
        rank, world = _world(self.group)
        nq = queries.shape[0]
        start, stop = balance(nq, world, rank)
        ids, dists = self.local_search(queries[start:stop], k)
        if world == 1:
            return ids, dists
        counts = [balance(nq, world, r)[1] - balance(nq, world, r)[0] for r in range(world)]
        # one collective for both result arrays: ids and the float32 distances (bit-cast, widened to
        # int64 lanes) travel as a single [rows, 2k] int64 block -- the gather is latency bound
        packed = torch.cat([ids, dists.view(torch.int32).to(torch.int64)], dim=1)
        out = _gather_rows(packed, counts, self.group)
        return out[:, :k].contiguous(), out[:, k:].to(torch.int32).view(torch.float32)


def merge_topk_reference_order(ids: np.ndarray, dists: np.ndarray, k: int, greater: bool):
    """Host restatement of the cross-shard merge (TotalOrder: distance, then id).  Test helper and
    documentation of what ``svsb200_merge_topk_device`` computes; shapes [shards, nq, k]."""
    shards, nq, _ = ids.shape
    out_i = np.empty((nq, k), dtype=ids.dtype)
    out_d = np.empty((nq, k), dtype=np.float32)
    for q in range(nq):
        cand = [(float(-dists[s, q, j]) if greater else float(dists[s, q, j]), int(ids[s, q, j]), float(dists[s, q, j]))
                for s in range(shards) for j in range(ids.shape[2]) if ids[s, q, j] >= 0]
        cand.sort(key=lambda t: (t[0], t[1]))
        for j in range(k):
            if j < len(cand):
                out_i[q, j], out_d[q, j] = cand[j][1], cand[j][2]
            else:
                out_i[q, j], out_d[q, j] = -1, (-np.inf if greater else np.inf)
    return out_i, out_d


def cuda_merge(ids: torch.Tensor, dists: torch.Tensor, k: int, greater: bool):
    """``svsb200_merge_topk_device`` on [shards, nq, k] CUDA tensors (ids int64)."""
    from . import _lib
    lib = _lib.lib()
    shards, nq, kk = ids.shape
    assert kk == k and ids.is_cuda and ids.dtype == torch.int64 and dists.dtype == torch.float32
    out_i = torch.empty((nq, k), dtype=torch.int64, device=ids.device)
    out_d = torch.empty((nq, k), dtype=torch.float32, device=ids.device)
    stream = _stream_handle(ids.device)
    _lib.check(lib.svsb200_merge_topk_device(ids.data_ptr(), dists.data_ptr(), shards, nq, k, 1 if greater else 0,
                                             out_i.data_ptr(), out_d.data_ptr(), ids.device.index, stream))
    return out_i, out_d


class ShardedSearch:
    """Mode B: shard the base vectors (each shard its own graph), search all queries everywhere,
    all-gather the per-shard top-k, merge with TotalOrder."""

    def __init__(self, local_search, id_offset: int, greater: bool, merge=cuda_merge, group=None):
        self.local_search = local_search
        self.id_offset = int(id_offset)
        self.greater = bool(greater)
        self.merge = merge
        self.group = group

    def search(self, queries: torch.Tensor, k: int):
        rank, world = _world(self.group)
        ids, dists = self.local_search(queries, k)
        valid = ids >= 0
        ids = torch.where(valid, ids + self.id_offset, ids)
        if world == 1:
            return ids, dists
        nq = ids.shape[0]
        all_ids = torch.empty((world * nq, k), dtype=ids.dtype, device=ids.device)
        all_d = torch.empty((world * nq, k), dtype=dists.dtype, device=dists.device)
        dist.all_gather_into_tensor(all_ids, ids.contiguous(), group=self.group)
        dist.all_gather_into_tensor(all_d, dists.contiguous(), group=self.group)
        return self.merge(all_ids.view(world, nq, k), all_d.view(world, nq, k), k, self.greater)


def _stream_handle(device) -> int:
    """torch's current stream as a cudaStream_t for the C ABI.  torch reports the legacy default stream
    as 0, which the ABI reads as "use the index's own stream": pass cudaStreamLegacy (0x1) instead so the
    work stays ordered with the caller's torch ops and events."""
    return torch.cuda.current_stream(device).cuda_stream or 1


def cuda_local_search(index, id_dtype=torch.int64):
    """Adapter: a :class:`~scalablevectorsearch_b200.Vamana` as ``local_search`` over CUDA tensors.
    Enqueues on torch's current stream; nothing synchronises."""
    np_dtype = {torch.float32: np.float32, torch.float16: np.float16, torch.int8: np.int8, torch.uint8: np.uint8}
    if id_dtype not in (torch.int32, torch.int64):
        raise TypeError("id_dtype must be torch.int32 or torch.int64")
    id_bytes = 4 if id_dtype == torch.int32 else 8

    def run(queries: torch.Tensor, k: int):
        assert queries.is_cuda and queries.is_contiguous()
        nq = queries.shape[0]
        ids = torch.empty((nq, k), dtype=id_dtype, device=queries.device)
        dists = torch.empty((nq, k), dtype=torch.float32, device=queries.device)
        if nq:
            index.search_device(queries.data_ptr(), np_dtype[queries.dtype], nq, k, ids.data_ptr(), dists.data_ptr(),
                                stream=_stream_handle(queries.device), id_bytes=id_bytes)
        return ids, dists

    return run
