// runtime_demo.cpp -- a program written against the reference's runtime header only
// (bindings/cpp/include/svs/runtime/vamana_index.h), linked with libsvsb200_runtime.so instead of libsvs_runtime:
// build from a float file, plain / filtered / range search; prints labels so a test can compare them.
#include "svs/runtime/vamana_index.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

namespace rt = svs::runtime::v0;

struct EvenIds : rt::IDFilter {
    bool is_member(size_t id) const override { return id % 2 == 0; }
};
struct VecAllocator : rt::ResultsAllocator {
    mutable std::vector<size_t> labels;
    mutable std::vector<float> distances;
    rt::SearchResultsStorage allocate(std::span<size_t> counts) const override {
        size_t total = 0;
        for (size_t c : counts) total += c;
        labels.assign(total, 0);
        distances.assign(total, 0.f);
        return {std::span<size_t>(labels), std::span<float>(distances)};
    }
};

int main(int argc, char** argv) {
    if (argc != 6) {
        std::fprintf(stderr, "usage: %s <data.f32> <n> <dim> <queries.f32> <nq>\n", argv[0]);
        return 2;
    }
    const size_t n = std::strtoul(argv[2], nullptr, 10), dim = std::strtoul(argv[3], nullptr, 10), nq = std::strtoul(argv[5], nullptr, 10);
    std::vector<float> data(n * dim), queries(nq * dim);
    FILE* f = std::fopen(argv[1], "rb");
    if (!f || std::fread(data.data(), 4, n * dim, f) != n * dim) return 3;
    std::fclose(f);
    f = std::fopen(argv[4], "rb");
    if (!f || std::fread(queries.data(), 4, nq * dim, f) != nq * dim) return 3;
    std::fclose(f);
    rt::VamanaIndex* index = nullptr;
    rt::VamanaIndex::BuildParams bp;
    bp.graph_max_degree = 32;
    bp.construction_window_size = 64;
    rt::VamanaIndex::SearchParams sp;
    sp.search_window_size = 40;
    rt::Status st = rt::VamanaIndex::build(&index, dim, rt::MetricType::L2, rt::StorageKind::FP32, bp, sp);
    if (st.ok()) st = index->add(n, data.data());
    if (!st.ok()) {
        std::fprintf(stderr, "error: %s\n", st.message());
        return 1;
    }
    const size_t k = 5;
    std::vector<float> d(nq * k);
    std::vector<size_t> l(nq * k);
    st = index->search(nq, queries.data(), k, d.data(), l.data());
    if (!st.ok()) return 1;
    for (size_t i = 0; i < nq * k; ++i) std::printf("plain %zu %.9g\n", l[i], d[i]);
    EvenIds even;
    st = index->search(nq, queries.data(), k, d.data(), l.data(), nullptr, &even);
    if (!st.ok()) return 1;
    for (size_t i = 0; i < nq * k; ++i) std::printf("even %zu %.9g\n", l[i], d[i]);
    VecAllocator alloc;
    st = index->range_search(2, queries.data(), d[k - 1], alloc);
    if (!st.ok()) return 1;
    for (size_t i = 0; i < alloc.labels.size(); ++i) std::printf("range %zu %.9g\n", alloc.labels[i], alloc.distances[i]);
    float one = 0.f;
    st = index->get_distance(l[0], queries.data(), &one);
    std::printf("get_distance %zu %.9g\n", l[0], one);
    rt::VamanaIndex::destroy(index);
    return 0;
}
