"""Deterministic synthetic datasets of SURVEY.md §8(d): a 1024-cluster Gaussian mixture, unit-normalised
(Deep1B descriptors are unit-norm).  Fixed seeds so every box regenerates identical bytes."""
from __future__ import annotations

import numpy as np

CENTRES_SEED, BASE_SEED, QUERY_SEED = 1001, 1002, 2001


def _mixture(n: int, dim: int, centres: np.ndarray, seed: int, sigma: float, chunk: int = 1 << 18) -> np.ndarray:
    rng = np.random.default_rng(seed)
    out = np.empty((n, dim), dtype=np.float32)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        which = rng.integers(0, centres.shape[0], size=hi - lo)
        x = centres[which] + np.float32(sigma) * rng.standard_normal((hi - lo, dim), dtype=np.float32)
        x /= np.sqrt((x * x).sum(axis=1, keepdims=True, dtype=np.float32))
        out[lo:hi] = x
    return out


def clustered_unit_vectors(n: int, nq: int, dim: int, clusters: int = 1024, sigma: float = 0.35):
    """Returns (base float32 [n,dim], queries float32 [nq,dim])."""
    centres = np.random.default_rng(CENTRES_SEED).standard_normal((clusters, dim), dtype=np.float32)
    return _mixture(n, dim, centres, BASE_SEED, sigma), _mixture(nq, dim, centres, QUERY_SEED, sigma)


def clustered_queries(nq: int, dim: int, shard: int = 0, clusters: int = 1024, sigma: float = 0.35) -> np.ndarray:
    """Query block number ``shard`` of the same law (block 0 == the queries of :func:`clustered_unit_vectors`);
    used to give every GPU its own batch when per-GPU work is held fixed."""
    centres = np.random.default_rng(CENTRES_SEED).standard_normal((clusters, dim), dtype=np.float32)
    return _mixture(nq, dim, centres, QUERY_SEED + shard, sigma)


def clustered_base_block(n: int, dim: int, block: int, clusters: int = 1024, sigma: float = 0.35) -> np.ndarray:
    """Base-vector block number ``block`` of the same law: sharded datasets are defined shard by shard (shard s
    of a G-way sharded index is block s), so every rank generates only its own rows."""
    centres = np.random.default_rng(CENTRES_SEED).standard_normal((clusters, dim), dtype=np.float32)
    return _mixture(n, dim, centres, BASE_SEED + 7919 * (block + 1), sigma)
