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


class ReplicatedSearch:
    """Mode A: replicate the index, shard the queries, gather the result rows.

    One collective per batch and nothing else on the data path: every rank owns a byte block
    ``[ids | distances]`` of ``width = max rows per rank`` rows; the local search writes its rows straight into
    that block when the callable takes ``out=`` (the CUDA path does), and one ``all_gather_into_tensor`` of the
    blocks (NCCL over NVLink) completes the batch -- no concatenation, no widening of the float distances."""

    def __init__(self, local_search: Callable[..., tuple[torch.Tensor, torch.Tensor]], group=None,
                 id_dtype: torch.dtype = torch.int64):
        """``local_search(queries, k[, out=(ids, dists)]) -> (ids [m,k], dists float32 [m,k])`` on this rank's
        device; ``id_dtype`` is the id element type of the gathered result (int64, or int32 to halve the bytes)."""
        self.local_search = local_search
        self.group = group
        self.id_dtype = id_dtype
        self._takes_out = bool(getattr(local_search, "takes_out", False))
        self._blocks = {}

    def _buffers(self, world, width, k, device):
        key = (world, width, k, str(device))
        if key not in self._blocks:
            isz = torch.empty((), dtype=self.id_dtype).element_size()
            nb_ids, nb_d = width * k * isz, width * k * 4
            mine = torch.empty(nb_ids + nb_d, dtype=torch.uint8, device=device)
            everyone = torch.empty(world * (nb_ids + nb_d), dtype=torch.uint8, device=device)
            self._blocks = {key: (mine, everyone, nb_ids)}   # keep only the current shape
        return self._blocks[key]

    def search(self, queries: torch.Tensor, k: int) -> tuple[torch.Tensor, torch.Tensor]:
        rank, world = _world(self.group)
        nq = queries.shape[0]
        start, stop = balance(nq, world, rank)
        m = stop - start
        if world == 1:
            return self.local_search(queries, k)
        counts = [balance(nq, world, r)[1] - balance(nq, world, r)[0] for r in range(world)]
        width = max(counts)
        mine, everyone, nb_ids = self._buffers(world, width, k, queries.device)
        ids_v = mine[:nb_ids].view(self.id_dtype).view(width, k)
        d_v = mine[nb_ids:].view(torch.float32).view(width, k)
        if self._takes_out:
            self.local_search(queries[start:stop], k, out=(ids_v[:m], d_v[:m]))
        else:
            ids, dists = self.local_search(queries[start:stop], k)
            ids_v[:m].copy_(ids)
            d_v[:m].copy_(dists)
        dist.all_gather_into_tensor(everyone, mine, group=self.group)
        blocks = everyone.view(world, -1)
        all_ids = blocks[:, :nb_ids].view(self.id_dtype).view(world, width, k)
        all_d = blocks[:, nb_ids:].view(torch.float32).view(world, width, k)
        if all(c == width for c in counts):
            return all_ids.reshape(world * width, k), all_d.reshape(world * width, k)
        return (torch.cat([all_ids[r, :counts[r]] for r in range(world)], dim=0),
                torch.cat([all_d[r, :counts[r]] for r in range(world)], dim=0))


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

    def run(queries: torch.Tensor, k: int, out=None):
        assert queries.is_cuda and queries.is_contiguous()
        nq = queries.shape[0]
        if out is None:
            ids = torch.empty((nq, k), dtype=id_dtype, device=queries.device)
            dists = torch.empty((nq, k), dtype=torch.float32, device=queries.device)
        else:
            ids, dists = out
            assert ids.dtype == id_dtype and ids.is_contiguous() and dists.is_contiguous() and ids.shape == (nq, k)
        if nq:
            index.search_device(queries.data_ptr(), np_dtype[queries.dtype], nq, k, ids.data_ptr(), dists.data_ptr(),
                                stream=_stream_handle(queries.device), id_bytes=id_bytes)
        return ids, dists

    run.takes_out = True
    return run
