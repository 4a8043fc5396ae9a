// fast_lvq8.cu -- lean search-kernel instantiations (search_fast.cuh) for lvq8 rows.
#include "search_fast.cuh"

namespace svsb200 {

template <> cudaError_t launch_search_fast<ROW_LVQ8>(int op, const SearchParams& p, const LaunchConfig& cfg) {
    switch (op) {
        case OP_L2F: return launch_fast_dims<ROW_LVQ8, OP_L2F>(p, cfg);
        case OP_IPF: return launch_fast_dims<ROW_LVQ8, OP_IPF>(p, cfg);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace svsb200
