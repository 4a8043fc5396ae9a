"""GPU graph build vs the reference CPU builder on a bench workload: build seconds, average degree, recall@10 at a few
windows (ground truth from the exhaustive scan).  usage: build_bench.py [workload] [--no-ref]"""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
import torch
from scalablevectorsearch_b200 import DistanceType, SearchBufferConfig, Vamana, VamanaBuildParameters, build_graph
from scalablevectorsearch_b200.synthetic import clustered_unit_vectors

name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "c2-1Mx96-f32-L2-w128"
w = bench.WORKLOADS[name]
base, queries = clustered_unit_vectors(w["n"], w["nq"], w["dim"])
if w["dtype"] == "float16":
    base = base.astype(np.float16)
metric = {"l2": DistanceType.L2, "ip": DistanceType.MIP}[w["metric"]]
t0 = time.time()
g, ep = build_graph(base, metric, VamanaBuildParameters(alpha=w["alpha"], graph_max_degree=w["max_degree"], window_size=w["build_window"]))
t_gpu = time.time() - t0
out = {"workload": name, "gpu_build_s": round(t_gpu, 2), "gpu_avg_degree": float(g[:, 0].mean()), "gpu_entry_point": ep}


def recalls(graph, ep):
    index = Vamana.from_arrays(base, graph, ep, metric)
    sample = 1000
    qd = torch.from_numpy(queries[:sample]).cuda()
    gi = torch.empty((sample, 10), dtype=torch.int64, device="cuda"); gd = torch.empty((sample, 10), dtype=torch.float32, device="cuda")
    index.exhaustive_device(qd.data_ptr(), queries.dtype, sample, 10, gi.data_ptr(), gd.data_ptr(), stream=torch.cuda.current_stream().cuda_stream or 1)
    torch.cuda.synchronize()
    gt = gi.cpu().numpy()
    r = {}
    for win in (10, 32, 128):
        index.search_parameters.buffer_config = SearchBufferConfig(win)
        ids, _ = index.search(queries[:sample], 10)
        r[win] = round(float(np.mean([len(set(ids[i].tolist()) & set(gt[i].tolist())) for i in range(sample)])) / 10, 4)
    index.search_parameters.buffer_config = SearchBufferConfig(w["window"])
    index.search(queries, 10); index.search(queries, 10)
    r["kernel_ms_full_batch"] = round(index.last_kernel_ms(), 4)
    return r


out["gpu_recall"] = recalls(g, ep)
if "--no-ref" not in sys.argv:
    from oracle.bindings import RefLib
    t0 = time.time()
    gr, epr = RefLib().build(base, w["metric"], w["max_degree"], w["build_window"], alpha=w["alpha"], threads=bench.effective_cpus())
    out.update(ref_build_s=round(time.time() - t0, 2), ref_threads=bench.effective_cpus(), ref_avg_degree=float(gr[:, 0].mean()), ref_entry_point=epr,
               ref_recall=recalls(gr, epr))
print(json.dumps(out))
