"""Host-side logic that needs no GPU: file formats, parameter objects, partitioning."""
import numpy as np
import pytest

from scalablevectorsearch_b200 import SearchBufferConfig, VamanaSearchParameters, io
from scalablevectorsearch_b200.multi_gpu import balance, merge_topk_reference_order
from scalablevectorsearch_b200.synthetic import clustered_unit_vectors


def test_vecs_and_native_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    for dt, ext in ((np.float32, ".fvecs"), (np.uint32, ".ivecs"), (np.uint8, ".bvecs")):
        a = (rng.random((7, 5)) * 100).astype(dt)
        io.write_vecs(str(tmp_path / f"a{ext}"), a)
        assert np.array_equal(io.read_vecs(str(tmp_path / f"a{ext}")), a)
    a = rng.standard_normal((9, 6)).astype(np.float32)
    io.write_svs(str(tmp_path / "a.svs"), a)
    assert np.array_equal(io.read_svs(str(tmp_path / "a.svs"), np.float32), a)
    raw = open(tmp_path / "a.svs", "rb").read()
    assert len(raw) == 1024 + a.nbytes and raw[:8] == (0xCAD4A6B2579980FE).to_bytes(8, "little")
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.svs"
        bad.write_bytes(b"\0" * 2048)
        io.read_svs(str(bad), np.float32)


def test_known_fixture_layout(dataset):
    """tests/utils/test_dataset.h: 10 000 x 128 data, 1000 queries, graph max degree 128, entry 9426."""
    assert dataset.data.shape == (10000, 128) and dataset.queries.shape == (1000, 128)
    assert dataset.graph.shape == (10000, 129) and dataset.entry_point == 9426
    assert int(dataset.graph[:, 0].max()) <= 128
    assert float(dataset.data.sum()) == 28887.0        # SUM_OF_ALL_VECTORS, test_dataset.h:125


def test_search_buffer_config_invariants():
    """tests/svs/index/vamana/search_buffer.cpp:314-342."""
    assert SearchBufferConfig(10).search_buffer_capacity == 10
    assert SearchBufferConfig(10, 20).search_buffer_capacity == 20
    with pytest.raises(ValueError):
        SearchBufferConfig(20, 10)
    p = VamanaSearchParameters()
    assert (p.buffer_config.search_window_size, p.search_buffer_visited_set, p.prefetch_lookahead, p.prefetch_step) == \
        (0, False, 4, 1)   # search_params.h:35-48


def test_balance_matches_reference_partition():
    """lib/threads/types.h:311-329: contiguous, sizes differ by at most one, longer ranges first."""
    for n in (0, 1, 7, 10, 10000, 10007):
        for parts in (1, 2, 3, 8):
            ranges = [balance(n, parts, i) for i in range(parts)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_total_order_merge_model():
    ids = np.array([[[3, 9, -1]], [[4, 7, 8]]], dtype=np.int64)
    d = np.array([[[0.5, 2.0, np.inf]], [[0.5, 1.0, 3.0]]], dtype=np.float32)
    oi, od = merge_topk_reference_order(ids, d, 3, greater=False)
    assert oi.tolist() == [[3, 4, 7]] and od.tolist() == [[0.5, 0.5, 1.0]]   # equal distance -> smaller id first
    oi, od = merge_topk_reference_order(ids[:, :, :2], d[:, :, :2], 2, greater=True)
    assert oi.tolist() == [[9, 7]]


def test_synthetic_generator_is_deterministic_and_unit_norm():
    a, qa = clustered_unit_vectors(2000, 50, 96)
    b, qb = clustered_unit_vectors(2000, 50, 96)
    assert np.array_equal(a, b) and np.array_equal(qa, qb)
    assert np.allclose(np.linalg.norm(a, axis=1), 1.0, atol=1e-5)
    assert a.dtype == np.float32 and qa.shape == (50, 96)


def test_toml_subset_reader(tmp_path):
    """The library's own TOML reader (svsb200_toml_get) on the reference's saved-configuration shape
    (data/test_dataset/vamana_config.toml): nested tables, indentation, comments, strings, booleans, floats."""
    from scalablevectorsearch_b200 import Svsb200Error
    from scalablevectorsearch_b200.vamana import toml_get
    p = tmp_path / "svs_config.toml"
    p.write_text("""# comment
__version__ = 'v0.0.2'   # trailing comment

[object]
__schema__ = 'vamana_index_parameters'
entry_point = 9426
name = "vamana # not a comment"

    [object.build_parameters]
    alpha = 1.2000000476837158
    use_full_search_history = true

    [object.search_parameters]
    search_buffer_capacity = 40
    search_window_size = 24
[[ignored.array]]
x = 1
""")
    assert toml_get(str(p), "object.entry_point") == "9426"
    assert toml_get(str(p), "__version__") == "v0.0.2"
    assert toml_get(str(p), "object.name") == "vamana # not a comment"
    assert float(toml_get(str(p), "object.build_parameters.alpha")) == pytest.approx(1.2)
    assert toml_get(str(p), "object.build_parameters.use_full_search_history") == "true"
    assert toml_get(str(p), "object.search_parameters.search_window_size") == "24"
    with pytest.raises(Svsb200Error):
        toml_get(str(p), "ignored.array.x")
    with pytest.raises(Svsb200Error):
        toml_get(str(p), "object.missing")
