"""Cycles per phase of a hop (diagnostic build: scratch/build_variant.sh clocks -DSVSB200_PHASE_CLOCKS; SVSB200_LIB=...)."""
import ctypes as C, json, sys
import numpy as np
sys.path.insert(0, ".")
import bench, torch
from scalablevectorsearch_b200 import DistanceType, SearchBufferConfig, Vamana, _lib
w, base, queries, graph, ep = bench.load_workload("c2-1Mx96-f32-L2-w128", 0, 1, lambda: None)
index = Vamana.from_arrays(base, graph, ep, DistanceType.L2)
p = index.search_parameters; p.buffer_config = SearchBufferConfig(128, 128); index.search_parameters = p
lib = _lib.lib(); k = 10
st = torch.cuda.current_stream().cuda_stream or 1
out = {}
buf = (C.c_ulonglong * 8)()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for cold in (0, 1):
  for nq in (1, 32):
    dq = torch.from_numpy(np.ascontiguousarray(queries[:nq])).cuda()
    ids = torch.empty((nq, k), dtype=torch.int64, device="cuda"); d = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    for it in range(3):
        if cold:
            flush.zero_(); torch.cuda.synchronize()
            lib.svsb200_debug_phase_clocks(buf, 1)
        index.search_device(dq.data_ptr(), queries.dtype, nq, k, ids.data_ptr(), d.data_ptr(), stream=st); torch.cuda.synchronize()
        lib.svsb200_debug_phase_clocks(buf, 1)
    v = list(buf); hops = max(1, v[4])
    out[f"{'cold' if cold else 'warm'}_{nq}"] = {"kernel_ms": round(index.last_kernel_ms(), 4), "hops": v[4], "cycles_per_hop": {"next+adjacency": round(v[0] / hops), "filter": round(v[1] / hops),
               "distances": round(v[2] / hops), "merge": round(v[3] / hops)}}
print(json.dumps(out))
