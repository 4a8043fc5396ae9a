"""World-size-2 NCCL tests of the multi-GPU paths on real devices (skipped on a single-GPU box):
ReplicatedSearch / ShardedSearch over `torch.distributed` (one process per GPU), and the single-process
multi-device forms of the C ABI (replicated index, svsb200_search_sharded with direct peer writes)."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT, bits, knn_graph

pytestmark = pytest.mark.gpu


def _ngpus():
    import torch
    return torch.cuda.device_count()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from conftest import Dataset
    from scalablevectorsearch_b200 import DistanceType, SearchBufferConfig, Vamana
    from scalablevectorsearch_b200.multi_gpu import ReplicatedSearch, ShardedSearch, cuda_local_search
    ds = Dataset()
    # Mode A: replicas, 901 queries (unequal slices), int32 and int64 ids
    index = Vamana.from_arrays(ds.data, ds.graph, ds.entry_point, DistanceType.L2, device=rank)
    index.search_parameters.buffer_config = SearchBufferConfig(22, 23)
    q = torch.from_numpy(ds.queries[99:]).cuda()
    out = {}
    for name, dt in (("i32", torch.int32), ("i64", torch.int64)):
        ids, d = ReplicatedSearch(cuda_local_search(index, id_dtype=dt), id_dtype=dt).search(q, 10)
        torch.cuda.synchronize()
        out[f"a_{name}_ids"], out[f"a_{name}_d"] = ids.cpu().numpy().astype(np.int64), d.cpu().numpy()
    # Mode B: each rank owns half of the base vectors with its own graph
    n = ds.data.shape[0]
    lo, hi = (0, n // 2) if rank == 0 else (n // 2, n)
    g = knn_graph(ds.data[lo:hi], 24, np.random.default_rng(7 + rank))
    shard = Vamana.from_arrays(ds.data[lo:hi], g, 3, DistanceType.L2, device=rank)
    shard.search_parameters.buffer_config = SearchBufferConfig(32, 48)
    qb = torch.from_numpy(ds.queries[:200]).cuda()
    b_ids, b_d = ShardedSearch(cuda_local_search(shard), id_offset=lo, greater=False).search(qb, 10)
    torch.cuda.synchronize()
    out["b_ids"], out["b_d"] = b_ids.cpu().numpy(), b_d.cpu().numpy()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_world2_nccl_replicated_and_sharded(dataset, ref_outputs, oracle, tmp_path):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from scalablevectorsearch_b200.multi_gpu import merge_topk_reference_order
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (np.load(tmp_path / f"rank{r}.npz") for r in range(2))
    oi, od = oracle.index(dataset.data, dataset.graph, dataset.entry_point, "l2").search(dataset.queries[99:], 10, 22, 23)
    for name in ("i32", "i64"):
        for r in (r0, r1):   # every rank holds the whole result
            assert np.array_equal(r[f"a_{name}_ids"], oi.astype(np.int64)), name
            assert np.array_equal(bits(r[f"a_{name}_d"]), bits(od)), name
    n = dataset.data.shape[0]
    parts = []
    for rank, (lo, hi) in enumerate(((0, n // 2), (n // 2, n))):
        g = knn_graph(dataset.data[lo:hi], 24, np.random.default_rng(7 + rank))
        wi, wd = oracle.index(dataset.data[lo:hi], g, 3, "l2").search(dataset.queries[:200], 10, 32, 48)
        parts.append((wi.astype(np.int64) + lo, wd))
    want_i, want_d = merge_topk_reference_order(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]), 10, False)
    for r in (r0, r1):
        assert np.array_equal(r["b_ids"], want_i) and np.array_equal(bits(r["b_d"]), bits(want_d))


def test_one_process_replicas_and_sharded_over_two_devices(dataset, ref_outputs, oracle):
    """C ABI, one process: svsb200_index_create_multi (batch split over the devices) and svsb200_search_sharded
    (shards on different devices: the shards' kernels write their rows into device 0's gather block over NVLink)."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    from scalablevectorsearch_b200 import DistanceType, SearchBufferConfig, ShardedVamana, Vamana
    from scalablevectorsearch_b200.multi_gpu import merge_topk_reference_order
    index = Vamana.from_arrays(dataset.data, dataset.graph, dataset.entry_point, DistanceType.L2, device=[0, 1])
    assert index.num_devices == 2
    index.search_parameters.buffer_config = SearchBufferConfig(22, 23)
    ids, d = index.search(dataset.queries[100:], 10)
    assert np.array_equal(ids, ref_outputs["l2_f32_f32_w22_c23_ids"].astype(np.uint64))
    assert np.array_equal(bits(d), bits(ref_outputs["l2_f32_f32_w22_c23_dists"]))
    n = dataset.data.shape[0]
    shards, parts, offs = [], [], [0, n // 2]
    q = dataset.queries[:200]
    for dev, (lo, hi) in enumerate(((0, n // 2), (n // 2, n))):
        g = knn_graph(dataset.data[lo:hi], 24, np.random.default_rng(7 + dev))
        shards.append(Vamana.from_arrays(dataset.data[lo:hi], g, 3, DistanceType.L2, device=dev))
        wi, wd = oracle.index(dataset.data[lo:hi], g, 3, "l2").search(q, 10, 32, 48)
        parts.append((wi.astype(np.int64) + lo, wd))
    sv = ShardedVamana(shards, offs)
    sv.search_parameters.buffer_config = SearchBufferConfig(32, 48)
    ids, d = sv.search(q, 10)
    want_i, want_d = merge_topk_reference_order(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]), 10, False)
    assert np.array_equal(ids.astype(np.int64), want_i) and np.array_equal(bits(d), bits(want_d))
