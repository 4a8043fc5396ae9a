"""Why is a 1250-query batch slower than a 148-query one?  Kernel time of (a) the first nq queries, (b) the slowest query
of the batch replicated nq times (no spread in hop counts: what is left is contention), with the hop statistics."""
import json, sys
import numpy as np
sys.path.insert(0, ".")
import bench, torch
from scalablevectorsearch_b200 import DistanceType, SearchBufferConfig, Vamana
w, base, queries, graph, ep = bench.load_workload("c2-1Mx96-f32-L2-w128", 0, 1, lambda: None)
index = Vamana.from_arrays(base, graph, ep, DistanceType.L2)
p = index.search_parameters; p.buffer_config = SearchBufferConfig(128, 128); index.search_parameters = p
k = 10
index.set_counting(True)
index.search(queries, k)
hops, evals = index.counters(len(queries))[:2]
index.set_counting(False)
hops = np.asarray(hops)
out = {"hops_mean": float(hops.mean()), "hops_p50": float(np.percentile(hops, 50)), "hops_p99": float(np.percentile(hops, 99)),
       "hops_max": int(hops.max()), "hops_max_first": {n: int(hops[:n].max()) for n in (32, 148, 625, 1250, 2500)}}
st = torch.cuda.current_stream().cuda_stream or 1
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(qnp):
    dq = torch.from_numpy(np.ascontiguousarray(qnp)).cuda(); nq = len(qnp)
    ids = torch.empty((nq, k), dtype=torch.int64, device="cuda"); d = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    ts = []
    for it in range(7):
        flush.zero_()
        index.search_device(dq.data_ptr(), qnp.dtype, nq, k, ids.data_ptr(), d.data_ptr(), stream=st); torch.cuda.synchronize()
        ts.append(index.last_kernel_ms())
    return round(float(np.median(ts[2:])), 4)
slow = int(hops.argmax()); med = int(np.argsort(hops)[len(hops) // 2])
for n in (1, 32, 148, 625, 1250, 2500):
    out[f"first_{n}"] = run(queries[:n])
    out[f"slowest_x{n}"] = run(np.repeat(queries[slow:slow + 1], n, axis=0))
    out[f"median_x{n}"] = run(np.repeat(queries[med:med + 1], n, axis=0))
print(json.dumps(out))
