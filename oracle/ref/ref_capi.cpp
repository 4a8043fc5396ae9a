// oracle/_ref driver -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// A thin `extern "C"` façade that instantiates the *unmodified* reference headers
// (`-I/root/reference/include`, never copied) so that tests and `bench.py --impl
// reference` can run the reference's own AVX-512 Vamana search / build on host cores.
//
// Reference entry points exercised (file:line under /root/reference):
//   index/vamana/index.h:564-611     VamanaIndex::search (batch, thread pool)
//   index/vamana/greedy_search.h:124 greedy_search (per-query, with counting tracker)
//   index/vamana/index.h:968-994     auto_build
//   core/distance/*.h                distance::compute / maybe_fix_argument
//   quantization/scalar/scalar.h     SQDataset::compress + compressed distances
//
// Nothing in the product (libsvsb200.so / scalablevectorsearch_b200) links or loads this.

#include "svs/core/data.h"
#include "svs/core/distance.h"
#include "svs/core/graph.h"
#include "svs/extensions/vamana/scalar.h"
#include "svs/index/vamana/index.h"
#include "svs/lib/float16.h"
#include "svs/quantization/scalar/scalar.h"

#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace {

thread_local std::string g_error;

enum DType : int { F32 = 0, F16 = 1, I8 = 2, U8 = 3 };
enum Metric : int { L2 = 0, IP = 1, COS = 2 };

using Graph = svs::graphs::SimpleGraph<uint32_t, svs::lib::Allocator<uint32_t>>;
template <typename T> using Data = svs::data::SimpleData<T, svs::Dynamic, svs::lib::Allocator<T>>;
template <typename T> using SQData =
    svs::quantization::scalar::SQDataset<T, svs::Dynamic, svs::lib::Allocator<T>>;

struct CountTracker {
    size_t hops = 0;
    size_t dists = 0;
    template <class N> void visited(N, size_t n) {
        ++hops;
        dists += n;
    }
};

Graph make_graph(const uint32_t* rows, size_t n, size_t max_degree) {
    // `rows` is the reference's own in-memory layout: n x (max_degree + 1), element 0 is
    // the out-degree (core/graph/graph.h:103-114).
    auto g = Graph(n, max_degree);
    for (size_t i = 0; i < n; ++i) {
        const uint32_t* r = rows + i * (max_degree + 1);
        g.replace_node(i, std::span<const uint32_t>(r + 1, r[0]));
    }
    return g;
}

template <typename T> Data<T> make_data(const void* src, size_t n, size_t dim) {
    auto d = Data<T>(n, dim);
    const T* p = static_cast<const T*>(src);
    for (size_t i = 0; i < n; ++i) {
        d.set_datum(i, std::span<const T>(p + i * dim, dim));
    }
    return d;
}

struct IndexBase {
    virtual ~IndexBase() = default;
    virtual void search(
        int qtype,
        const void* queries,
        size_t nq,
        size_t k,
        size_t window,
        size_t capacity,
        bool visited,
        uint64_t* ids,
        float* dists
    ) = 0;
    virtual void counts(
        int qtype,
        const void* queries,
        size_t nq,
        size_t window,
        size_t capacity,
        uint64_t* hops,
        uint64_t* evals
    ) = 0;
    virtual void set_threads(size_t n) = 0;
    virtual size_t dims() const = 0;
};

template <typename DataT, typename Dist> struct IndexImpl : IndexBase {
    using Index = svs::index::vamana::VamanaIndex<Graph, DataT, Dist>;
    Index index;

    IndexImpl(Graph g, DataT d, uint32_t ep, size_t threads)
        : index(std::move(g), std::move(d), ep, Dist{}, threads) {}

    size_t dims() const override { return index.dimensions(); }
    void set_threads(size_t n) override {
        index.set_threadpool(svs::threads::DefaultThreadPool(n));
    }

    template <typename Q>
    void search_typed(
        const Q* queries,
        size_t nq,
        size_t k,
        size_t window,
        size_t capacity,
        bool visited,
        uint64_t* ids,
        float* dists
    ) {
        auto view = svs::data::ConstSimpleDataView<Q>(queries, nq, index.dimensions());
        auto sp = index.get_search_parameters();
        sp.buffer_config({window, capacity});
        sp.search_buffer_visited_set(visited);
        auto result = svs::QueryResultView<size_t>(
            svs::MatrixView<size_t>(
                svs::make_dims(nq, k), reinterpret_cast<size_t*>(ids)
            ),
            svs::MatrixView<float>(svs::make_dims(nq, k), dists)
        );
        index.search(result, view, sp);
    }

    template <typename Q>
    void counts_typed(
        const Q* queries,
        size_t nq,
        size_t window,
        size_t capacity,
        uint64_t* hops,
        uint64_t* evals
    ) {
        size_t dim = index.dimensions();
        index.experimental_escape_hatch([&](const auto& graph,
                                            const auto& data,
                                            const auto& distance,
                                            auto eps) {
            std::vector<uint32_t> ep(eps.begin(), eps.end());
            namespace v = svs::index::vamana;
            for (size_t q = 0; q < nq; ++q) {
                auto buffer = typename Index::search_buffer_type(
                    v::SearchBufferConfig(window, capacity),
                    svs::distance::comparator(distance)
                );
                auto scratch = v::extensions::single_search_setup(data, distance);
                auto accessor = svs::data::GetDatumAccessor{};
                auto tracker = CountTracker{};
                auto query = std::span<const Q>(queries + q * dim, dim);
                v::greedy_search(
                    graph,
                    data,
                    accessor,
                    query,
                    scratch,
                    buffer,
                    v::EntryPointInitializer<uint32_t>{svs::lib::as_const_span(ep)},
                    v::NeighborBuilder{},
                    tracker
                );
                // The initializer reports one `visited` per entry point (1 evaluation
                // each); every later call is a node expansion.
                hops[q] = tracker.hops - ep.size();
                evals[q] = tracker.dists;
            }
        });
    }

    void search(
        int qtype,
        const void* q,
        size_t nq,
        size_t k,
        size_t w,
        size_t c,
        bool visited,
        uint64_t* ids,
        float* dists
    ) override {
        using E = typename DataT::element_type;
        if (qtype == F32) {
            search_typed(static_cast<const float*>(q), nq, k, w, c, visited, ids, dists);
        } else if (qtype == F16) {
            search_typed(
                static_cast<const svs::Float16*>(q), nq, k, w, c, visited, ids, dists
            );
        } else if constexpr (std::is_same_v<E, int8_t>) {
            if (qtype != I8) {
                throw ANNEXCEPTION("unsupported query type {}", qtype);
            }
            search_typed(static_cast<const int8_t*>(q), nq, k, w, c, visited, ids, dists);
        } else if constexpr (std::is_same_v<E, uint8_t>) {
            if (qtype != U8) {
                throw ANNEXCEPTION("unsupported query type {}", qtype);
            }
            search_typed(static_cast<const uint8_t*>(q), nq, k, w, c, visited, ids, dists);
        } else {
            throw ANNEXCEPTION("unsupported query type {}", qtype);
        }
    }

    void counts(
        int qtype,
        const void* q,
        size_t nq,
        size_t w,
        size_t c,
        uint64_t* hops,
        uint64_t* evals
    ) override {
        using E = typename DataT::element_type;
        if (qtype == F32) {
            counts_typed(static_cast<const float*>(q), nq, w, c, hops, evals);
        } else if (qtype == F16) {
            counts_typed(static_cast<const svs::Float16*>(q), nq, w, c, hops, evals);
        } else if constexpr (std::is_same_v<E, int8_t>) {
            counts_typed(static_cast<const int8_t*>(q), nq, w, c, hops, evals);
        } else if constexpr (std::is_same_v<E, uint8_t>) {
            counts_typed(static_cast<const uint8_t*>(q), nq, w, c, hops, evals);
        } else {
            throw ANNEXCEPTION("unsupported query type {}", qtype);
        }
    }
};

template <typename DataT>
IndexBase* make_index(DataT data, Graph g, uint32_t ep, int metric, size_t threads) {
    switch (metric) {
        case L2:
            return new IndexImpl<DataT, svs::distance::DistanceL2>(
                std::move(g), std::move(data), ep, threads
            );
        case IP:
            return new IndexImpl<DataT, svs::distance::DistanceIP>(
                std::move(g), std::move(data), ep, threads
            );
        case COS:
            return new IndexImpl<DataT, svs::distance::DistanceCosineSimilarity>(
                std::move(g), std::move(data), ep, threads
            );
        default:
            throw ANNEXCEPTION("bad metric {}", metric);
    }
}

template <typename Q, typename T, typename Dist>
void distance_rows(const Q* q, const T* rows, size_t nrows, size_t dim, float* out) {
    auto dist = Dist{};
    auto qs = std::span<const Q>(q, dim);
    svs::distance::maybe_fix_argument(dist, qs);
    for (size_t i = 0; i < nrows; ++i) {
        out[i] = svs::distance::compute(dist, qs, std::span<const T>(rows + i * dim, dim));
    }
}

template <typename Q, typename T>
void distance_rows_metric(
    int metric, const void* q, const void* rows, size_t nrows, size_t dim, float* out
) {
    auto* qq = static_cast<const Q*>(q);
    auto* rr = static_cast<const T*>(rows);
    switch (metric) {
        case L2: distance_rows<Q, T, svs::distance::DistanceL2>(qq, rr, nrows, dim, out); break;
        case IP: distance_rows<Q, T, svs::distance::DistanceIP>(qq, rr, nrows, dim, out); break;
        case COS:
            distance_rows<Q, T, svs::distance::DistanceCosineSimilarity>(
                qq, rr, nrows, dim, out
            );
            break;
        default: throw ANNEXCEPTION("bad metric {}", metric);
    }
}

template <typename F> int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    } catch (...) {
        g_error = "unknown exception";
        return 2;
    }
}

} // namespace

extern "C" {

const char* svsref_last_error() { return g_error.c_str(); }

/// 1 when the runtime dispatcher selects the AVX-512 kernels (lib/avx_detection.h:24-57);
/// parity with the GPU path is defined against that tree.
int svsref_avx512() {
    return svs::detail::avx_runtime_flags.is_avx512f_supported() ? 1 : 0;
}
int svsref_avx512vnni() {
    return svs::detail::avx_runtime_flags.is_avx512vnni_supported() ? 1 : 0;
}

int svsref_distance_rows(
    int metric,
    int qtype,
    int dtype,
    const void* query,
    const void* rows,
    size_t nrows,
    size_t dim,
    float* out
) {
    return guarded([&] {
        using svs::Float16;
        int key = qtype * 4 + dtype;
        switch (key) {
            case F32 * 4 + F32: distance_rows_metric<float, float>(metric, query, rows, nrows, dim, out); break;
            case F32 * 4 + F16: distance_rows_metric<float, Float16>(metric, query, rows, nrows, dim, out); break;
            case F32 * 4 + I8: distance_rows_metric<float, int8_t>(metric, query, rows, nrows, dim, out); break;
            case F32 * 4 + U8: distance_rows_metric<float, uint8_t>(metric, query, rows, nrows, dim, out); break;
            case F16 * 4 + F32: distance_rows_metric<Float16, float>(metric, query, rows, nrows, dim, out); break;
            case F16 * 4 + F16: distance_rows_metric<Float16, Float16>(metric, query, rows, nrows, dim, out); break;
            case I8 * 4 + I8: distance_rows_metric<int8_t, int8_t>(metric, query, rows, nrows, dim, out); break;
            case U8 * 4 + U8: distance_rows_metric<uint8_t, uint8_t>(metric, query, rows, nrows, dim, out); break;
            default: throw ANNEXCEPTION("unsupported (query,data) pair {} {}", qtype, dtype);
        }
    });
}

void* svsref_index_create(
    int dtype,
    const void* data,
    size_t n,
    size_t dim,
    const uint32_t* graph_rows,
    size_t max_degree,
    uint32_t entry_point,
    int metric,
    size_t threads
) {
    IndexBase* out = nullptr;
    int rc = guarded([&] {
        auto g = make_graph(graph_rows, n, max_degree);
        switch (dtype) {
            case F32: out = make_index(make_data<float>(data, n, dim), std::move(g), entry_point, metric, threads); break;
            case F16: out = make_index(make_data<svs::Float16>(data, n, dim), std::move(g), entry_point, metric, threads); break;
            case I8: out = make_index(make_data<int8_t>(data, n, dim), std::move(g), entry_point, metric, threads); break;
            case U8: out = make_index(make_data<uint8_t>(data, n, dim), std::move(g), entry_point, metric, threads); break;
            default: throw ANNEXCEPTION("bad dtype {}", dtype);
        }
    });
    return rc == 0 ? out : nullptr;
}

/// Scalar-quantised index: compress `data` (f32) with the reference's own
/// `SQDataset::compress` (quantization/scalar/scalar.h:447-469), return scale/bias and
/// the codes so the GPU side can be fed the identical bytes.
void* svsref_sq_index_create(
    int code_type, // I8 or U8
    const float* data,
    size_t n,
    size_t dim,
    const uint32_t* graph_rows,
    size_t max_degree,
    uint32_t entry_point,
    int metric,
    size_t threads,
    float* scale_out,
    float* bias_out,
    void* codes_out
) {
    IndexBase* out = nullptr;
    int rc = guarded([&] {
        auto g = make_graph(graph_rows, n, max_degree);
        auto raw = make_data<float>(data, n, dim);
        auto finish = [&](auto sq) {
            *scale_out = sq.get_scale();
            *bias_out = sq.get_bias();
            using E = typename decltype(sq)::element_type;
            auto* dst = static_cast<E*>(codes_out);
            for (size_t i = 0; i < n; ++i) {
                auto row = sq.get_datum(i);
                std::memcpy(dst + i * dim, row.data(), dim * sizeof(E));
            }
            out = make_index(std::move(sq), std::move(g), entry_point, metric, threads);
        };
        if (code_type == I8) {
            finish(SQData<int8_t>::compress(raw, threads));
        } else if (code_type == U8) {
            finish(SQData<uint8_t>::compress(raw, threads));
        } else {
            throw ANNEXCEPTION("bad SQ code type {}", code_type);
        }
    });
    return rc == 0 ? out : nullptr;
}

void svsref_index_destroy(void* h) { delete static_cast<IndexBase*>(h); }

int svsref_index_set_threads(void* h, size_t threads) {
    return guarded([&] { static_cast<IndexBase*>(h)->set_threads(threads); });
}

int svsref_index_search(
    void* h,
    int qtype,
    const void* queries,
    size_t nq,
    size_t k,
    size_t window,
    size_t capacity,
    int visited_set,
    uint64_t* ids,
    float* dists
) {
    return guarded([&] {
        static_cast<IndexBase*>(h)->search(
            qtype, queries, nq, k, window, capacity, visited_set != 0, ids, dists
        );
    });
}

/// Per-query work counters from a counting `GreedySearchTracker`
/// (index/vamana/greedy_search.h:38-42,165): `hops[q]` = expanded nodes, `evals[q]` =
/// distance evaluations including entry points.  These define the ALGORITHMIC bytes of
/// DESIGN.md §measurement.
int svsref_index_counts(
    void* h,
    int qtype,
    const void* queries,
    size_t nq,
    size_t window,
    size_t capacity,
    uint64_t* hops,
    uint64_t* evals
) {
    return guarded([&] {
        static_cast<IndexBase*>(h)->counts(qtype, queries, nq, window, capacity, hops, evals);
    });
}

/// Reference graph construction (index/vamana/index.h:968-994 auto_build: medoid entry
/// point, two passes alpha=1 then alpha).  Output layout = n x (max_degree+1), degree first.
int svsref_build(
    int dtype,
    const void* data,
    size_t n,
    size_t dim,
    int metric,
    float alpha,
    size_t max_degree,
    size_t window,
    size_t max_candidates,
    size_t prune_to,
    size_t threads,
    uint32_t* graph_out,
    uint32_t* entry_point_out
) {
    return guarded([&] {
        auto params = svs::index::vamana::VamanaBuildParameters{
            alpha, max_degree, window, max_candidates, prune_to, true};
        auto extract = [&](const auto& index) {
            index.experimental_escape_hatch(
                [&](const auto& graph, const auto&, const auto&, auto eps) {
                    *entry_point_out = eps[0];
                    for (size_t i = 0; i < n; ++i) {
                        auto nb = graph.get_node(i);
                        uint32_t* row = graph_out + i * (max_degree + 1);
                        row[0] = static_cast<uint32_t>(nb.size());
                        for (size_t j = 0; j < max_degree; ++j) {
                            row[1 + j] = j < nb.size() ? nb[j] : 0;
                        }
                    }
                }
            );
        };
        auto run = [&](auto data_typed) {
            auto alloc = svs::lib::Allocator<uint32_t>{};
            switch (metric) {
                case L2: extract(svs::index::vamana::auto_build(params, std::move(data_typed), svs::distance::DistanceL2{}, threads, alloc)); break;
                case IP: extract(svs::index::vamana::auto_build(params, std::move(data_typed), svs::distance::DistanceIP{}, threads, alloc)); break;
                case COS: extract(svs::index::vamana::auto_build(params, std::move(data_typed), svs::distance::DistanceCosineSimilarity{}, threads, alloc)); break;
                default: throw ANNEXCEPTION("bad metric {}", metric);
            }
        };
        switch (dtype) {
            case F32: run(make_data<float>(data, n, dim)); break;
            case F16: run(make_data<svs::Float16>(data, n, dim)); break;
            default: throw ANNEXCEPTION("build supports f32/f16 data, got {}", dtype);
        }
    });
}

/// float -> Float16 with the reference's own conversion (lib/float16.h:54-79), used to
/// generate f16 fixtures that are bit-identical to what the reference would store.
void svsref_to_float16(const float* src, size_t n, uint16_t* dst) {
    for (size_t i = 0; i < n; ++i) {
        dst[i] = svs::Float16(src[i]).raw();
    }
}

} // extern "C"
