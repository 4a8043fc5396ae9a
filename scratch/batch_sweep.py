"""Kernel time of the search vs batch size on one GPU (the per-GPU share of a strong-scaled 10k batch at N = 1, 2, 4, 8)."""
import json, sys
import numpy as np
sys.path.insert(0, ".")
import bench, torch
from scalablevectorsearch_b200 import DistanceType, SearchBufferConfig, Vamana
name = sys.argv[1] if len(sys.argv) > 1 else "c2-1Mx96-f32-L2-w128"
w, base, queries, graph, ep = bench.load_workload(name, 0, 1, lambda: None)
metric = {"l2": DistanceType.L2, "ip": DistanceType.MIP}[w["metric"]]
index = Vamana.from_arrays(base, graph, ep, metric)
p = index.search_parameters
p.buffer_config = SearchBufferConfig(w["window"], w["window"])
index.search_parameters = p
dq = torch.from_numpy(queries).cuda()
k = 10
out = {}
st = torch.cuda.current_stream().cuda_stream or 1
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for nq in (10000, 5000, 2500, 1250, 625, 148, 32):
    ids = torch.empty((nq, k), dtype=torch.int64, device="cuda"); d = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    ts = []
    for it in range(8):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); index.search_device(dq.data_ptr(), queries.dtype, nq, k, ids.data_ptr(), d.data_ptr(), stream=st); e.record()
        torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    out[nq] = round(float(np.median(ts[3:])), 4)
print(json.dumps({"workload": name, "ms_by_batch": out, "note": "L2 flushed between launches; events around the whole enqueue (prepare + search kernels)"}))
