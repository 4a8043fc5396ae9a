// fast_i8.cu -- lean search-kernel instantiations (search_fast.cuh) for i8 rows.
#include "search_fast.cuh"

namespace svsb200 {

template <> cudaError_t launch_search_fast<SVSB200_I8>(int op, const SearchParams& p, const LaunchConfig& cfg) {
    switch (op) {
        case OP_L2F: return launch_fast_dims<SVSB200_I8, OP_L2F>(p, cfg);
        case OP_IPF: return launch_fast_dims<SVSB200_I8, OP_IPF>(p, cfg);
        case OP_COSF: return launch_fast_dims<SVSB200_I8, OP_COSF>(p, cfg);
        case OP_L2I: return launch_fast_dims<SVSB200_I8, OP_L2I>(p, cfg);
        case OP_IPI: return launch_fast_dims<SVSB200_I8, OP_IPI>(p, cfg);
        case OP_COSI: return launch_fast_dims<SVSB200_I8, OP_COSI>(p, cfg);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace svsb200
