"""Readers/writers for the on-disk formats the reference harness uses around the search path.

Formats (reference file:line under /root/reference/include/svs):
  * ``*.fvecs / *.ivecs / *.bvecs`` -- per-row ``int32 dim`` header followed by ``dim``
    elements (core/io/vecs.h:137-273).
  * native ``*.svs`` v1 -- 1024-byte header {magic 0xcad4a6b2579980fe, 16-byte uuid,
    u64 num_vectors, u64 dims} then raw row-major rows (core/io/native.h:315-345).  The
    Vamana graph file is the same container holding ``uint32[n][max_degree+1]`` rows with
    the out-degree in element 0 (core/graph/graph.h:103-114).
"""
from __future__ import annotations

import os
import struct

import numpy as np

SVS_MAGIC = 0xCAD4A6B2579980FE
SVS_HEADER_BYTES = 1024

_VECS_DTYPES = {".fvecs": np.float32, ".ivecs": np.uint32, ".bvecs": np.uint8, ".hvecs": np.float16}


def read_vecs(path: str, dtype=None) -> np.ndarray:
    """Read an ``[fibh]vecs`` file into an ``(n, dim)`` array."""
    ext = os.path.splitext(path)[1]
    dtype = np.dtype(dtype or _VECS_DTYPES[ext])
    raw = np.fromfile(path, dtype=np.uint8)
    if raw.size == 0:
        return np.zeros((0, 0), dtype=dtype)
    dim = int(raw[:4].view(np.int32)[0])
    row_bytes = 4 + dim * dtype.itemsize
    if raw.size % row_bytes:
        raise ValueError(f"{path}: size {raw.size} is not a multiple of row size {row_bytes}")
    rows = raw.reshape(-1, row_bytes)
    if not np.all(rows[:, :4].view(np.int32) == dim):
        raise ValueError(f"{path}: ragged vecs file")
    return np.ascontiguousarray(rows[:, 4:]).view(dtype).reshape(-1, dim)


def write_vecs(path: str, array: np.ndarray) -> None:
    array = np.ascontiguousarray(array)
    n, dim = array.shape
    out = np.empty((n, 4 + dim * array.dtype.itemsize), dtype=np.uint8)
    out[:, :4] = np.frombuffer(struct.pack("<i", dim), dtype=np.uint8)
    out[:, 4:] = array.view(np.uint8).reshape(n, -1)
    out.tofile(path)


def read_svs(path: str, dtype) -> np.ndarray:
    """Read a native v1 ``.svs`` container as ``(num_vectors, dims)`` of ``dtype``."""
    dtype = np.dtype(dtype)
    with open(path, "rb") as f:
        header = f.read(SVS_HEADER_BYTES)
        magic, = struct.unpack_from("<Q", header, 0)
        if magic != SVS_MAGIC:
            raise ValueError(f"{path}: bad magic {magic:#x}")
        n, dims = struct.unpack_from("<QQ", header, 24)
        body = np.fromfile(f, dtype=dtype, count=n * dims)
    if body.size != n * dims:
        raise ValueError(f"{path}: truncated ({body.size} of {n * dims} elements)")
    return body.reshape(n, dims)


def write_svs(path: str, array: np.ndarray) -> None:
    array = np.ascontiguousarray(array)
    n, dims = array.shape
    header = bytearray(SVS_HEADER_BYTES)
    struct.pack_into("<Q", header, 0, SVS_MAGIC)
    struct.pack_into("<QQ", header, 24, n, dims)
    with open(path, "wb") as f:
        f.write(header)
        array.tofile(f)


def read_graph(path: str) -> np.ndarray:
    """Read a native graph file: ``uint32[n][max_degree+1]`` rows, degree first."""
    return read_svs(path, np.uint32)
