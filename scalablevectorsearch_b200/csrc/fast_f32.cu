// fast_f32.cu -- lean search-kernel instantiations (search_fast.cuh) for f32 rows.
#include "search_fast.cuh"

namespace svsb200 {

template <> cudaError_t launch_search_fast<SVSB200_F32>(int op, const SearchParams& p, const LaunchConfig& cfg) {
    switch (op) {
        case OP_L2F: return launch_fast_dims<SVSB200_F32, OP_L2F>(p, cfg);
        case OP_IPF: return launch_fast_dims<SVSB200_F32, OP_IPF>(p, cfg);
        case OP_COSF: return launch_fast_dims<SVSB200_F32, OP_COSF>(p, cfg);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace svsb200

#ifdef SVSB200_PHASE_CLOCKS
// diagnostic build only: cycles per hop phase {next+adjacency, filter, distances, merge, hops} summed over the f32 launches
extern "C" int svsb200_debug_phase_clocks(unsigned long long* out, int reset) {
    if (cudaMemcpyFromSymbol(out, svsb200::g_phase_clocks, 8 * sizeof(unsigned long long)) != cudaSuccess) return -1;
    if (reset) {
        unsigned long long zero[8] = {};
        if (cudaMemcpyToSymbol(svsb200::g_phase_clocks, zero, sizeof(zero)) != cudaSuccess) return -1;
    }
    return 0;
}
#endif
