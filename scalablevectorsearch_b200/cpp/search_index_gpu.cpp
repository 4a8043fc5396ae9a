// search_index_gpu.cpp -- the reference's `utils/search_index.cpp` CLI with one change: the
// index handed to the orchestrator is a svsb200::GpuVamanaIndex.  Same positional arguments,
// same output file (`<prefix>_idx.ivecs`), so BASELINE config #1 (data/test_dataset through
// the CLI) runs unchanged.  Everything except GpuVamanaIndex is the reference's public API:
// svs::index::vamana::auto_assemble (index/vamana/index.h:1022-1046), svs::VamanaImpl / svs::Vamana
// (orchestrators/vamana.h:108-305), svs::load_data, QueryResult::save_vecs.
#include "gpu_vamana_index.h"

#include "svs/core/distance.h"
#include "svs/core/graph.h"
#include "svs/core/data.h"
#include "svs/orchestrators/vamana.h"

#include <cstdio>
#include <filesystem>
#include <string>

namespace {

template <typename Eq, typename Edb, typename Dist>
void run(
    const std::string& query_file,
    size_t window,
    size_t k,
    size_t threads,
    const std::filesystem::path& config,
    const std::filesystem::path& graph,
    const std::filesystem::path& data,
    const std::string& prefix,
    const std::vector<int>& devices
) {
    auto cpu = svs::index::vamana::auto_assemble(
        config,
        svs::GraphLoader(graph),
        svs::VectorDataLoader<Edb, svs::Dynamic>(data),
        Dist{},
        threads
    );
    auto index = svsb200::make_gpu_vamana<svs::lib::Types<Eq>>(svsb200::GpuVamanaIndex{std::move(cpu), devices});
    index.set_search_parameters(index.get_search_parameters().buffer_config({window}));
    const auto queries = svs::load_data<Eq>(query_file);
    auto tic = svs::lib::now();
    auto result = index.search(queries, k);
    std::printf(
        "backend: %s\nGlobal search time: %g seconds (%zu queries)\n",
        index.experimental_backend_string().c_str(),
        svs::lib::time_difference(tic),
        queries.size()
    );
    result.save_vecs(prefix + "_idx.ivecs");
}

template <typename Eq, typename Edb, typename... Args> void by_distance(const std::string& d, Args&&... a) {
    if (d == "L2") {
        run<Eq, Edb, svs::distance::DistanceL2>(a...);
    } else if (d == "MIP") {
        run<Eq, Edb, svs::distance::DistanceIP>(a...);
    } else if (d == "Cosine") {
        run<Eq, Edb, svs::distance::DistanceCosineSimilarity>(a...);
    } else {
        throw ANNEXCEPTION("Unknown distance {}", d);
    }
}

} // namespace

int main(int argc, char** argv) {
    if (argc != 12 && argc != 13) {
        std::fprintf(
            stderr,
            "usage: %s <query type: float|float16|int8|uint8> <data type> <query file> <search window> "
            "<neighbors> <threads> <config dir> <graph dir> <data dir> <result prefix> <L2|MIP|Cosine> "
            "[device[,device...]]   (several devices: one replica each, the batch is split over them)\n",
            argv[0]
        );
        return 2;
    }
    try {
        const std::string qt = argv[1], dt = argv[2], qfile = argv[3];
        const size_t window = std::stoul(argv[4]), k = std::stoul(argv[5]), threads = std::stoul(argv[6]);
        const std::filesystem::path config = argv[7], graph = argv[8], data = argv[9];
        const std::string prefix = argv[10], dist = argv[11];
        std::vector<int> devices;
        for (std::string rest = argc == 13 ? argv[12] : "0"; !rest.empty();) {
            const size_t comma = rest.find(',');
            devices.push_back(std::stoi(rest.substr(0, comma)));
            rest = comma == std::string::npos ? "" : rest.substr(comma + 1);
        }
        auto go = [&]<typename Eq, typename Edb>() {
            by_distance<Eq, Edb>(dist, qfile, window, k, threads, config, graph, data, prefix, devices);
        };
        if (qt == "float" && dt == "float") {
            go.template operator()<float, float>();
        } else if (qt == "float" && dt == "float16") {
            go.template operator()<float, svs::Float16>();
        } else if (qt == "float" && dt == "int8") {
            go.template operator()<float, int8_t>();
        } else if (qt == "float" && dt == "uint8") {
            go.template operator()<float, uint8_t>();
        } else if (qt == "int8" && dt == "int8") {
            go.template operator()<int8_t, int8_t>();
        } else if (qt == "uint8" && dt == "uint8") {
            go.template operator()<uint8_t, uint8_t>();
        } else {
            throw ANNEXCEPTION("Unsupported (query, data) element types ({}, {})", qt, dt);
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
