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
if os.environ.get("SANITIZE_LIGHT") == "2":   # one float32 index, the lean kernel only (racecheck analysis)
    x = rng.standard_normal((600, 96)).astype(np.float32); q = rng.standard_normal((40, 96)).astype(np.float32)
    ix = Vamana.from_arrays(x, graph(600, 64), 1, DistanceType.L2)
    for w, c in ((7, 19), (64, 64)):
        ix.search_parameters.buffer_config = SearchBufferConfig(w, c); ix.search(q, 5)
    print("minimal sanitize workload done"); sys.exit(0)
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
# ---- round 2: several entry points, filtered / range search, the tensor-core flat search, the GPU builder ----
from scalablevectorsearch_b200 import VamanaBuildParameters, build_graph
n, dim = 3000, 96
x = rng.standard_normal((n, dim)).astype(np.float32); q = rng.standard_normal((300, dim)).astype(np.float32)
g = graph(n, 32)
ix = Vamana.from_arrays(x, g, 1, DistanceType.L2)
ix.set_entry_points([1, 17, 99, 2048])
ix.search_parameters.buffer_config = SearchBufferConfig(3, 3); ix.search(q, 3)
ix.search_parameters.buffer_config = SearchBufferConfig(40, 50); ix.search(q, 10)
allowed = np.zeros(n, dtype=np.uint8); allowed[::3] = 1
ix.search_filtered(q[:64], 5, allowed)
ix.range_search(q[:16], 150.0)
if os.environ.get("SANITIZE_LIGHT"):
    print("round-2 sanitize workload done (light: no flat search / builder)"); sys.exit(0)
ix.flat_search(q, 10)                        # 3 query tiles: one CTA per segment
ix.flat_search(np.tile(q, (3, 1)), 10)       # 8 query tiles: row groups of 4
ix16 = Vamana.from_arrays(x.astype(np.float16), g, 1, DistanceType.MIP); ix16.flat_search(q[:130].astype(np.float16), 25)
xw = rng.standard_normal((2000, 272)).astype(np.float32)
Vamana.from_arrays(xw, graph(2000, 8), 0, DistanceType.L2).flat_search(rng.standard_normal((600, 272)).astype(np.float32), 10)   # pacing on
for m in (DistanceType.L2, DistanceType.MIP, DistanceType.Cosine):
    gg, ep = build_graph(x, m, VamanaBuildParameters(alpha=1.2 if m == DistanceType.L2 else 0.95, graph_max_degree=32, window_size=40))
    b = Vamana.from_arrays(x, gg, ep, m); b.search_parameters.buffer_config = SearchBufferConfig(20); b.search(q[:50], 10)
print("round-2 sanitize workload done")
