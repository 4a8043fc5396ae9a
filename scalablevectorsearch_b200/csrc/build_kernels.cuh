// build_kernels.cuh -- GPU Vamana graph construction: the reference's VamanaBuilder
// (index/vamana/vamana_build.h:169-599) restated as rounds of whole-batch kernels.
//
// Per round over a batch of nodes (vamana_build.h:221-310 `construct`):
//   1. generate_neighbors (:320-470): greedy search from the medoid for every batch node -- the lean search
//      kernel in HIST mode (search_fast.cuh) with the visited filter on and the full search history kept --
//      then candidate pool = history + the node's current neighbours, sorted by TotalOrder, truncated to
//      max_candidate_pool_size, and alpha-robust pruning to graph_max_degree (prune.h: Progressive strategy
//      for L2 :183-240, Iterative for inner product :107-170): `build_prune_kernel`, mode FORWARD;
//   2. add_reverse_edges (:475-580): every new edge v->u is mirrored as u->v when u has room
//      (`reverse_insert_kernel`), otherwise recorded in a per-vertex overflow list; vertices with overflow are
//      re-pruned over (overflow + adjacency) to prune_to: `build_prune_kernel`, mode REVERSE.
// Two passes (index.h:436-439): reverse pruning with alpha = 1 first, then with the configured alpha; the
// forward pruning always uses the configured alpha (vamana_build.h:268-272).
//
// The reference's result depends on thread timing (vertex locks, bucket order), so the parity bar is the
// reference's own: recall equivalence of the built index (tests/integration/vamana/index_build.cpp:96,139).
// Distances inside the builder use the search path's bit-exact code (eval_pass), keys = sign * distance.
#pragma once

#include "search_fast.cuh"

#include <cfloat>

namespace svsb200 {

constexpr uint32_t kPoolMax = 512;   // candidate pool entries per node held in shared memory

struct BuildParams {
    int reverse;                 // 0: forward (pool = search history + adjacency), 1: reverse (overflow + adjacency)
    uint32_t first;              // forward: first node of the batch (node = first + i, query row i)
    uint32_t count;              // forward: batch size
    const uint32_t* count_ptr;   // reverse: number of touched vertices (device)
    const uint32_t* nodes;       // reverse: the touched vertices
    const uint2* hist;           // forward: [count][hist_cap] {key bits, id}
    const uint32_t* hist_count;
    uint32_t hist_cap;
    const int* head;             // reverse: overflow linked lists (head per vertex, -1 = empty)
    const int* next;
    const uint32_t* pair_v;
    uint32_t* graph;             // [n][gstride], neighbours first, kNoNeighbor padding
    uint32_t* deg;               // [n]
    uint32_t gstride;
    uint32_t max_degree;         // R
    uint32_t limit;              // size of the pruned list: R (forward) / prune_to (reverse)
    uint32_t max_candidates;
    float alpha;
    int iterative;               // prune strategy: 0 = Progressive (L2), 1 = Iterative (inner product, cosine)
    unsigned int* work_counter;
};

// Per-CTA (one warp) shared memory of build_prune_kernel.
__host__ __device__ inline size_t build_smem_bytes(uint32_t qstride) {
    // query | candidate ids | candidate keys | prune state | compact list ids, keys, pool positions | result
    return size_t(qstride) * 4 + size_t(kPoolMax) * 4 * 6 + 256 * 4;
}

// Row `id` of the dataset as the fp32 operand the distance code expects in q_s (what prepare_queries_kernel
// does for a query: exact conversion, zero padding).
// Returns the row's Euclidean norm (the `a_norm` of CosineSimilarity::fix_argument, cosine.h:117-119).
template <int ROWT>
__device__ __forceinline__ float stage_row(const SearchParams& p, uint32_t id, float* q_s, int lane) {
    const char* row = reinterpret_cast<const char*>(p.vectors) + size_t(id) * p.row_stride;
    float sq = 0.f;
    for (uint32_t i = lane; i < p.qstride; i += 32) {
        float v = 0.f;
        if (i < p.dim) {
            if constexpr (ROWT == SVSB200_F32) v = reinterpret_cast<const float*>(row)[i];
            else v = __half2float(reinterpret_cast<const __half*>(row)[i]);
        }
        q_s[i] = v;
        sq = fmaf(v, v, sq);
    }
    for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(0xFFFFFFFFu, sq, o);
    return sqrtf(sq);
}

template <int ROWT, int OP, int DS, int KS>
__global__ void __launch_bounds__(32, 16) build_prune_kernel(const __grid_constant__ SearchParams p,
                                                            const __grid_constant__ BuildParams bp) {
    constexpr int G = KS * 16 / Row<ROWT>::LPT;
    constexpr int GROUPS = 32 / G;
    constexpr unsigned FULL = 0xFFFFFFFFu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const int g = lane / G, t = lane % G;
    const unsigned lt_mask = (1u << lane) - 1u;
    float* q_s = reinterpret_cast<float*>(smem_raw);
    uint32_t* cid = reinterpret_cast<uint32_t*>(q_s + p.qstride);   // [kPoolMax] pool ids
    float* ckey = reinterpret_cast<float*>(cid + kPoolMax);           // [kPoolMax] key of (node, candidate)
    float* state = ckey + kPoolMax;                                   // [kPoolMax] prune state
    uint32_t* lid = reinterpret_cast<uint32_t*>(state + kPoolMax);    // [kPoolMax] compact list: ids
    float* lkey = reinterpret_cast<float*>(lid + kPoolMax);           // [kPoolMax] compact list: keys
    uint32_t* lpos = reinterpret_cast<uint32_t*>(lkey + kPoolMax);    // [kPoolMax] compact list: pool positions
    uint32_t* res = lpos + kPoolMax;                                  // [256] result
    const char* vectors = reinterpret_cast<const char*>(p.vectors);
    const float ksign = p.greater ? -1.0f : 1.0f;
    const uint32_t total = bp.reverse ? *bp.count_ptr : bp.count;

    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(bp.work_counter, 1u);
        w = __shfl_sync(FULL, w, 0);
        if (w >= total) break;
        const uint32_t node = bp.reverse ? bp.nodes[w] : bp.first + w;
        const uint32_t deg0 = min(bp.deg[node], bp.max_degree);
        uint32_t* row = bp.graph + size_t(node) * bp.gstride;

        // ---- candidate pool ----
        uint32_t P = 0;       // entries whose key is known
        uint32_t M = 0;       // entries of lid[] whose distance to `node` still has to be computed
        float qnorm = 1.0f;   // cosine: norm of the vector currently staged as the query
        if (!bp.reverse) {
            // the node as a query: prepared by the search launch of this round
            for (uint32_t i = lane; i < p.qstride; i += 32) q_s[i] = p.qf[size_t(w) * p.qstride + i];
            qnorm = p.qaux[2 * size_t(w)];
            P = min(bp.hist_count[w], min(bp.hist_cap, kPoolMax));
            for (uint32_t i = lane; i < P; i += 32) {
                const uint2 e = bp.hist[size_t(w) * bp.hist_cap + i];
                ckey[i] = __uint_as_float(e.x);
                cid[i] = e.y;
            }
            __syncwarp();
            // neighbours of the node that the search did not visit (vamana_build.h:424-441)
            for (uint32_t j0 = 0; j0 < deg0; j0 += 32) {
                const uint32_t j = j0 + lane;
                const uint32_t id = j < deg0 ? row[j] : kNoNeighbor;
                bool fresh = id != kNoNeighbor && id != node;
                for (uint32_t i = 0; fresh && i < P; ++i) fresh = cid[i] != id;
                const unsigned m = __ballot_sync(FULL, fresh);
                if (fresh && P + M + __popc(m & lt_mask) < kPoolMax) lid[M + __popc(m & lt_mask)] = id;
                M = min(M + __popc(m), kPoolMax - P);
            }
        } else {
            qnorm = stage_row<ROWT>(p, node, q_s, lane);
            // overflow edges of this round (vamana_build.h:528-531), walked by one lane
            if (lane == 0) {
                uint32_t c = 0;
                for (int i = bp.head[node]; i >= 0 && c < kPoolMax; i = bp.next[i]) {
                    const uint32_t v = bp.pair_v[i];
                    bool dup = v == node;
                    for (uint32_t z = 0; !dup && z < c; ++z) dup = lid[z] == v;
                    if (!dup) lid[c++] = v;
                }
                res[0] = c;
            }
            __syncwarp();
            M = res[0];
            __syncwarp();
            // the old adjacency list, minus what the overflow already holds (:534-538)
            const uint32_t nover = M;
            for (uint32_t j0 = 0; j0 < deg0; j0 += 32) {
                const uint32_t j = j0 + lane;
                const uint32_t id = j < deg0 ? row[j] : kNoNeighbor;
                bool fresh = id != kNoNeighbor && id != node;
                for (uint32_t i = 0; fresh && i < nover; ++i) fresh = lid[i] != id;
                const unsigned m = __ballot_sync(FULL, fresh);
                if (fresh && M + __popc(m & lt_mask) < kPoolMax) lid[M + __popc(m & lt_mask)] = id;
                M = min(M + __popc(m), kPoolMax);
            }
        }
        __syncwarp();
        // distances node -> lid[0..M), appended to the pool
        for (uint32_t base = 0; base < M; base += 2 * GROUPS)
            eval_pass<ROWT, OP, DS, 2, KS, true>(p, q_s, vectors, lid, lkey, base, M, g, t, qnorm, 0.f, ksign);
        __syncwarp();
        for (uint32_t i = lane; i < M; i += 32) {
            cid[P + i] = lid[i];
            ckey[P + i] = lkey[i];
        }
        P += M;
        __syncwarp();

        // ---- sort by TotalOrder (key, id): bitonic over the next power of two ----
        uint32_t N = 32;
        while (N < P) N <<= 1;
        for (uint32_t i = P + lane; i < N; i += 32) {
            ckey[i] = INFINITY;
            cid[i] = 0xFFFFFFFFu;
        }
        __syncwarp();
        for (uint32_t k = 2; k <= N; k <<= 1) {
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t i = lane; i < N / 2; i += 32) {
                    const uint32_t a = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                    const uint32_t b = a | j;
                    const bool up = (a & k) == 0;
                    const float ka = ckey[a], kb = ckey[b];
                    const uint32_t ia = cid[a], ib = cid[b];
                    const bool a_gt_b = (kb < ka) || (!(ka < kb) && ib < ia);
                    if (a_gt_b == up) {
                        ckey[a] = kb; ckey[b] = ka;
                        cid[a] = ib; cid[b] = ia;
                    }
                }
                __syncwarp();
            }
        }
        P = min(P, bp.max_candidates);

        // ---- heuristic_prune_neighbors (prune.h): state[] is the Progressive strategy's `pruned` ratio
        // (tombstone = lowest, sentinel = max) or the Iterative strategy's state (0 available, 1 added, 2 pruned)
        for (uint32_t i = lane; i < P; i += 32) state[i] = bp.iterative ? 0.0f : -FLT_MAX;
        __syncwarp();
        uint32_t nres = 0;
        float cur_alpha = 1.0f;
        // while (result.size() < max && !cmp(alpha, current_alpha)): cmp is < for L2, > for inner product
        while (nres < bp.limit && !(p.greater ? bp.alpha > cur_alpha : bp.alpha < cur_alpha)) {
            uint32_t start = 0;
            while (nres < bp.limit && start < P) {
                // next pool entry that is not excluded at this alpha level
                uint32_t j = start + lane;
                bool ok = false;
                if (j < P) {
                    const float st = state[j];
                    const bool excluded = bp.iterative ? st != 0.0f : cur_alpha < st;
                    ok = !excluded && cid[j] != node;
                }
                const unsigned m = __ballot_sync(FULL, ok);
                if (m == 0) {
                    start += 32;
                    continue;
                }
                j = start + __ffs(m) - 1;
                start = j + 1;
                if (lane == 0) {
                    state[j] = bp.iterative ? 1.0f : FLT_MAX;
                    res[nres] = cid[j];
                }
                ++nres;
                qnorm = stage_row<ROWT>(p, cid[j], q_s, lane);
                __syncwarp();
                // the later entries still available at this level, compacted
                uint32_t L = 0;
                for (uint32_t t0 = start; t0 < P; t0 += 32) {
                    const uint32_t tt = t0 + lane;
                    bool av = false;
                    if (tt < P) {
                        const float st = state[tt];
                        av = bp.iterative ? st == 0.0f : !(cur_alpha < st);
                    }
                    const unsigned mm = __ballot_sync(FULL, av);
                    if (av) {
                        lid[L + __popc(mm & lt_mask)] = cid[tt];
                        lpos[L + __popc(mm & lt_mask)] = tt;
                    }
                    L += __popc(mm);
                }
                __syncwarp();
                if (L == 0) continue;
                // djk for every listed entry: key(selected, candidate)
                for (uint32_t base = 0; base < L; base += 2 * GROUPS)
                    eval_pass<ROWT, OP, DS, 2, KS, true>(p, q_s, vectors, lid, lkey, base, L, g, t, qnorm, 0.f, ksign);
                __syncwarp();
                for (uint32_t i = lane; i < L; i += 32) {
                    const uint32_t tt = lpos[i];
                    const float djk = lkey[i] * ksign;          // distance(selected, candidate)
                    const float dvt = ckey[tt] * ksign;         // distance(node, candidate)
                    if (bp.iterative) {
                        // if (cmp(current_alpha * djk, candidate.distance())) pruned   (prune.h:150)
                        const float lhs = cur_alpha * djk;
                        if (p.greater ? lhs > dvt : lhs < dvt) state[tt] = 2.0f;
                    } else {
                        // pruned[t] = std::max(pruned[t], candidate.distance() / djk, cmp)   (prune.h:227)
                        const float r = dvt / djk;
                        if (state[tt] < r) state[tt] = r;
                    }
                }
                __syncwarp();
            }
            if (bp.alpha == 1.0f) break;
            if (bp.iterative)
                for (uint32_t i = lane; i < P; i += 32)
                    if (state[i] == 2.0f) state[i] = 0.0f;   // reenable (prune.h:160-163)
            cur_alpha *= bp.alpha;
            __syncwarp();
        }
        __syncwarp();
        // ---- graph.replace_node(node, result) ----
        for (uint32_t j = lane; j < bp.gstride; j += 32) row[j] = j < nres ? res[j] : kNoNeighbor;
        if (lane == 0) bp.deg[node] = nres;
        __syncwarp();
    }
}

template <int ROWT, int OP> cudaError_t launch_build_prune(const SearchParams& p, const BuildParams& bp, int grid, cudaStream_t stream) {
    const size_t smem = build_smem_bytes(p.qstride);
    auto go = [&](auto kernel) -> cudaError_t {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (err != cudaSuccess) return err;
        kernel<<<grid, 32, smem, stream>>>(p, bp);
        count_launch();
        return cudaGetLastError();
    };
    if (p.dim == 96) return go(build_prune_kernel<ROWT, OP, 96, 1>);
    if (p.dim >= 256) return go(build_prune_kernel<ROWT, OP, 0, 4>);
    return go(build_prune_kernel<ROWT, OP, 0, 1>);
}

// Defined in build_f32.cu / build_f16.cu.
template <int ROWT> cudaError_t launch_build_search(int op, const SearchParams& p, const LaunchConfig& cfg);
template <int ROWT> cudaError_t launch_build_prune_op(int op, const SearchParams& p, const BuildParams& bp, int grid, cudaStream_t stream);

}  // namespace svsb200
