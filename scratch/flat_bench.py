"""Tensor-core flat search timing: exact top-10 of nq queries over a bench workload's base vectors.
usage: flat_bench.py [workload] [nq]"""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import bench, torch
from scalablevectorsearch_b200 import DistanceType, Vamana
from scalablevectorsearch_b200.synthetic import clustered_unit_vectors
name = sys.argv[1] if len(sys.argv) > 1 else "c2-1Mx96-f32-L2-w128"
w = bench.WORKLOADS[name]
nq = int(sys.argv[2]) if len(sys.argv) > 2 else w["nq"]
base, queries = clustered_unit_vectors(w["n"], nq, w["dim"])
if w["dtype"] == "float16":
    base = base.astype(np.float16)
metric = {"l2": DistanceType.L2, "ip": DistanceType.MIP}[w["metric"]]
index = Vamana.from_arrays(base, np.zeros((w["n"], 2), dtype=np.uint32), 0, metric)
dq = torch.from_numpy(queries).cuda()
ids = torch.empty((nq, 10), dtype=torch.int64, device="cuda"); d = torch.empty((nq, 10), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream or 1
fb = index.flat_search_device(dq.data_ptr(), queries.dtype, nq, 10, ids.data_ptr(), d.data_ptr(), stream=st)   # builds the tiles
torch.cuda.synchronize()
times = []
for _ in range(5):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); fb = index.flat_search_device(dq.data_ptr(), queries.dtype, nq, 10, ids.data_ptr(), d.data_ptr(), stream=st); e.record()
    torch.cuda.synchronize(); times.append(s.elapsed_time(e))
ms = float(np.median(times))
flops = 2.0 * nq * w["n"] * w["dim"]
peaks = json.load(open("MEASURED_PEAKS.json")) if __import__("os").path.exists("MEASURED_PEAKS.json") else {"bf16_tflops": 1590.0}
out = {"workload": name, "nq": nq, "flat_ms": round(ms, 3), "fallback_queries": fb, "gemm_tflops_incl_everything": round(flops / ms / 1e9, 1),
       "frac_of_measured_bf16_peak": round(flops / ms / 1e9 / peaks["bf16_tflops"], 4), "queries_per_s": round(nq / ms * 1e3)}
# the exact scan on a sample, for the speed ratio and an equality check
ns = min(nq, 500)
i2 = torch.empty((ns, 10), dtype=torch.int64, device="cuda"); d2 = torch.empty((ns, 10), dtype=torch.float32, device="cuda")
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
index.exhaustive_device(dq.data_ptr(), queries.dtype, ns, 10, i2.data_ptr(), d2.data_ptr(), stream=st); torch.cuda.synchronize()
s.record(); index.exhaustive_device(dq.data_ptr(), queries.dtype, ns, 10, i2.data_ptr(), d2.data_ptr(), stream=st); e.record(); torch.cuda.synchronize()
out.update(scan_ms_per_query=round(s.elapsed_time(e) / ns, 4), flat_ms_per_query=round(ms / nq, 5),
           equal_to_scan=bool(torch.equal(ids[:ns], i2) and torch.equal(d[:ns].view(torch.int32), d2.view(torch.int32))))
print(json.dumps(out))
