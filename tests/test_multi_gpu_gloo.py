"""World-size-2 `gloo` tests of the multi-GPU host logic (query sharding, result gather, cross-shard
merge).  The local search is an injected callable (here: the oracle on CPU tensors), so the plumbing that
runs over NCCL on the GPU box is exercised here without GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle_search(data, graph, ep, window):
    from oracle.bindings import OracleLib
    idx = OracleLib().index(data, graph, ep, "l2")

    def run(q: torch.Tensor, k: int):
        if q.shape[0] == 0:
            return torch.empty((0, k), dtype=torch.int64), torch.empty((0, k), dtype=torch.float32)
        ids, d = idx.search(q.numpy(), k, window, window)
        ids = ids.astype(np.int64)
        ids[ids == 0xFFFFFFFF] = -1
        return torch.from_numpy(ids), torch.from_numpy(np.nan_to_num(d, nan=np.inf))
    return run


def _oracle_search_into(data, graph, ep, window, id_dtype):
    """The same, through the `out=` protocol the CUDA local search uses (rows written straight into the rank's block)."""
    plain = _oracle_search(data, graph, ep, window)

    def run(q, k, out=None):
        ids, d = plain(q, k)
        if out is None:
            return ids.to(id_dtype), d
        out[0].copy_(ids.to(id_dtype))
        out[1].copy_(d)
        return out
    run.takes_out = True
    return run


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import Dataset
    from scalablevectorsearch_b200.multi_gpu import ReplicatedSearch, ShardedSearch, merge_topk_reference_order
    ds = Dataset()
    q = torch.from_numpy(ds.queries[:101])          # odd count: unequal shards
    # Mode A: replicas
    a_ids, a_d = ReplicatedSearch(_oracle_search(ds.data, ds.graph, ds.entry_point, 20)).search(q, 10)
    # the packed single-collective path with 4-byte ids, local rows written in place
    a32_ids, a32_d = ReplicatedSearch(_oracle_search_into(ds.data, ds.graph, ds.entry_point, 20, torch.int32),
                                      id_dtype=torch.int32).search(q, 10)
    assert a32_ids.dtype == torch.int32 and torch.equal(a32_ids.to(torch.int64), a_ids) and torch.equal(a32_d, a_d)
    # Mode B: each rank owns half of the base vectors with its own (sub)graph
    n = ds.data.shape[0]
    lo, hi = (0, n // 2) if rank == 0 else (n // 2, n)
    sub = ds.graph[lo:hi].copy()
    for i in range(sub.shape[0]):                      # keep only in-shard edges, re-based
        nb = sub[i, 1:1 + sub[i, 0]]
        nb = nb[(nb >= lo) & (nb < hi)] - lo
        sub[i, 0] = len(nb)
        sub[i, 1:1 + len(nb)] = nb
    merge = lambda i, d, k, g: tuple(torch.from_numpy(x) for x in merge_topk_reference_order(i.numpy(), d.numpy(), k, g))
    b_ids, b_d = ShardedSearch(_oracle_search(ds.data[lo:hi], sub, 17, 20), id_offset=lo, greater=False,
                               merge=merge).search(q, 10)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), a_ids=a_ids.numpy(), a_d=a_d.numpy(), b_ids=b_ids.numpy(),
             b_d=b_d.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_replicated_and_sharded(tmp_path, dataset, oracle):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    # every rank ends with the same, complete result
    for key in ("a_ids", "a_d", "b_ids", "b_d"):
        assert np.array_equal(r0[key], r1[key]), key
    # Mode A == single-process search of the whole batch
    ids, d = oracle.index(dataset.data, dataset.graph, dataset.entry_point, "l2").search(dataset.queries[:101], 10, 20, 20)
    assert np.array_equal(r0["a_ids"], ids.astype(np.int64)) and np.array_equal(r0["a_d"], d)
    # Mode B: merged lists are sorted by (distance, id), ids are global and unique per row
    b_ids, b_d = r0["b_ids"], r0["b_d"]
    assert b_ids.shape == (101, 10) and b_ids.min() >= 0 and b_ids.max() < dataset.data.shape[0]
    assert np.all(np.diff(b_d, axis=1) >= 0)
    assert all(len(set(r.tolist())) == 10 for r in b_ids)
    true_d = ((dataset.data[b_ids[:, 0]] - dataset.queries[:101]) ** 2).sum(1)
    assert np.allclose(true_d, b_d[:, 0], rtol=1e-5)
