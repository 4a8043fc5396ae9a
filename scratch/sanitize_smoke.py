"""Tiny searches for compute-sanitizer: every row type, split rows, SQ, LVQ-8, exhaustive, small windows."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scalablevectorsearch_b200 import DistanceType, SearchBufferConfig, Vamana, lvq8_compress
rng = np.random.default_rng(0)
def graph(n, R):
    g = np.zeros((n, R + 1), dtype=np.uint32)
    for i in range(n):
        d = rng.integers(1, R + 1); nb = rng.choice(n, size=d, replace=False); g[i, 0] = d; g[i, 1:1 + d] = nb
    return g
for dim, R in ((17, 5), (96, 64), (300, 33)):
    n = 600
    x = rng.standard_normal((n, dim)).astype(np.float32); q = rng.standard_normal((40, dim)).astype(np.float32)
    g = graph(n, R)
    for data, qq in ((x, q), (x.astype(np.float16), q.astype(np.float16)), ((x * 20).astype(np.int8), (q * 20).astype(np.int8)),
                     ((x * 20 + 128).clip(0, 255).astype(np.uint8), q)):
        for m in (DistanceType.L2, DistanceType.MIP, DistanceType.Cosine):
            ix = Vamana.from_arrays(data, g, 1, m)
            for w, c in ((1, 1), (7, 19), (64, 64)):
                ix.search_parameters.buffer_config = SearchBufferConfig(w, c)
                ix.search(qq, 5)
    rows, mean = lvq8_compress(x)
    ix = Vamana.from_arrays(rows, g, 1, DistanceType.L2, lvq8=(dim, mean)); ix.search_parameters.buffer_config = SearchBufferConfig(16)
    ix.search(q, 5)
    ix = Vamana.from_arrays((x * 20).astype(np.int8), g, 1, DistanceType.MIP, sq=(0.05, 0.1)); ix.search_parameters.buffer_config = SearchBufferConfig(16)
    ix.search(q, 5)
print("sanitize workload done")
