// svsb200.cu -- C ABI (include/svsb200.h) of the B200 Vamana search library: index upload,
// query preparation (the device-side maybe_fix_argument), launch plumbing, result gather.
//
// No CPU fallback lives here: every entry point either runs CUDA kernels on an sm_100
// device or fails with an error.
#include "common.cuh"

#include <atomic>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

namespace svsb200 {

static thread_local std::string g_error;
static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static int fail(const std::string& msg) {
    g_error = msg;
    return 1;
}
#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t err__ = (expr);                                                            \
        if (err__ != cudaSuccess) {                                                            \
            return fail(std::string(#expr) + ": " + cudaGetErrorString(err__));                \
        }                                                                                      \
    } while (0)

static size_t esize(int dtype) { return dtype == SVSB200_F32 ? 4 : dtype == SVSB200_F16 ? 2 : 1; }
static size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }

template <typename T> struct DeviceBuffer {
    T* ptr = nullptr;
    size_t count = 0;
    cudaError_t ensure(size_t n) {
        if (n <= count) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        count = 0;
        cudaError_t err = cudaMalloc(&ptr, n * sizeof(T));
        if (err == cudaSuccess) count = n;
        return err;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        count = 0;
    }
};

}  // namespace svsb200

using namespace svsb200;

struct svsb200_index {
    int device = 0;
    int sm_count = 0;
    int dtype = 0, metric = 0, storage = 0;
    size_t n = 0, dim = 0, max_degree = 0;
    uint32_t row_stride = 0, gstride = 0, entry_point = 0;
    float scale = 1.f, bias = 0.f;
    void* d_vectors = nullptr;
    uint32_t* d_graph = nullptr;
    uint16_t* d_ref_degree = nullptr;
    float* d_mean = nullptr;          // LVQ-8: dataset mean
    uint32_t lvq_const_offset = 0;
    size_t device_bytes = 0;
    // scratch, grown on demand
    DeviceBuffer<unsigned char> q_raw, q_codes, ids;
    DeviceBuffer<float> q_f32, q_aux, dists;
    DeviceBuffer<uint32_t> hops, evals, fetched;
    unsigned int* d_counter = nullptr;
    int counting = 0;
    size_t counted_nq = 0;
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    bool timed = false;
    // options
    long warps_per_cta = 0, ctas_per_sm = 0, rows_in_flight = 0, filter_slots = -1, filter_tag16 = 1, no_split = 0;
    long generic_kernel = 0;   // 1: force the generic (round-1) kernel instead of the lean one
    int last_kernel = 0;       // 1 = lean kernel, 0 = generic kernel (introspection for tests)
    std::mutex mutex;
};

namespace svsb200 {

// ---------------------------------------------------------------------------------------
// Upload kernels
// ---------------------------------------------------------------------------------------

// Reference adjacency rows (degree first, core/graph/graph.h:103-114) -> HBM layout:
// neighbours first, kNoNeighbor padding, row length a multiple of 4 words so rows stay
// 16-byte aligned.  Repeated ids inside a row keep their first occurrence only: the
// reference's insert() rejects the later copy as a duplicate (search_buffer.h:380-391) or
// drops it off the end, so removing it up front cannot change any result.
__global__ void repack_graph_kernel(const uint32_t* __restrict__ src, size_t row_len, uint32_t n,
                                    uint32_t* __restrict__ dst, uint32_t gstride, uint16_t* __restrict__ ref_degree,
                                    int* __restrict__ bad) {
    const uint32_t row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const uint32_t* in = src + size_t(row) * row_len;
    uint32_t* out = dst + size_t(row) * gstride;
    const uint32_t deg = in[0];
    if (deg > row_len - 1) {
        if (lane == 0) atomicExch(bad, 1);
        return;
    }
    if (lane == 0) ref_degree[row] = uint16_t(deg);
    uint32_t written = 0;
    for (uint32_t j0 = 0; j0 < deg; j0 += 32) {
        const uint32_t j = j0 + lane;
        uint32_t id = j < deg ? in[1 + j] : kNoNeighbor;
        bool keep = j < deg;
        if (keep && id >= n) {
            atomicExch(bad, 2);
            keep = false;
        }
        if (keep) {
            for (uint32_t i = 0; i < j; ++i) {
                if (in[1 + i] == id) {
                    keep = false;
                    break;
                }
            }
        }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
        if (keep) out[written + __popc(m & ((1u << lane) - 1u))] = id;
        written += __popc(m);
    }
    for (uint32_t j = written + lane; j < gstride; j += 32) out[j] = kNoNeighbor;
}

// ---------------------------------------------------------------------------------------
// Query preparation == distance::maybe_fix_argument for the whole batch
// (concepts/distance.h:90-130), one warp per query.
// ---------------------------------------------------------------------------------------
enum PrepMode : int {
    PREP_FLOAT = 0,   // float tree: operands converted exactly like the SIMD loads
    PREP_INT = 1,     // exact integer kernels: raw int8/uint8 query
    PREP_SQ_L2 = 2,   // EuclideanCompressed::fix_argument  (scalar.h:75-82)
    PREP_SQ_IP = 3,   // InnerProductCompressed::fix_argument (scalar.h:123-131)
    PREP_SQ_COS = 4,  // CosineSimilarityCompressed::fix_argument (scalar.h:168-171)
    PREP_LVQ_L2 = 5,  // LVQ-8, L2: query with the dataset mean removed (own spec)
    PREP_LVQ_IP = 6,  // LVQ-8, IP: raw query + <q, mean>
};

// Float16 -> float the way non-SIMD reference code does it (lib/float16.h:45-52):
// subnormals flush to signed zero.
__device__ __forceinline__ float f16_scalar(uint16_t x) {
    if ((x & 0x7C00u) == 0) return __uint_as_float(uint32_t(x & 0x8000u) << 16);
    return __half2float(__ushort_as_half(x));
}

template <int QT> __device__ __forceinline__ float q_simd(const void* q, uint32_t i) {
    if constexpr (QT == SVSB200_F32) return static_cast<const float*>(q)[i];
    if constexpr (QT == SVSB200_F16) return __half2float(__ushort_as_half(static_cast<const uint16_t*>(q)[i]));
    if constexpr (QT == SVSB200_I8) return float(static_cast<const int8_t*>(q)[i]);
    return float(static_cast<const uint8_t*>(q)[i]);
}
template <int QT> __device__ __forceinline__ float q_scalar(const void* q, uint32_t i) {
    if constexpr (QT == SVSB200_F16) return f16_scalar(static_cast<const uint16_t*>(q)[i]);
    return q_simd<QT>(q, i);
}

template <int QT>
__global__ void prepare_queries_kernel(const void* __restrict__ queries, uint32_t nq, uint32_t dim, uint32_t qstride,
                                       int mode, int metric, int code_type, float scale, float bias,
                                       const float* __restrict__ mean, float* __restrict__ qf,
                                       uint8_t* __restrict__ qcodes, float* __restrict__ qaux) {
    const uint32_t q = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    constexpr size_t QES = QT == SVSB200_F32 ? 4 : QT == SVSB200_F16 ? 2 : 1;
    const void* src = static_cast<const char*>(queries) + size_t(q) * dim * QES;
    float* f = qf + size_t(q) * qstride;
    uint8_t* c = qcodes + size_t(q) * qstride;

    for (uint32_t i = lane; i < qstride; i += 32) {
        float fv = 0.f;
        uint8_t cv = 0;
        if (i < dim) {
            if (mode == PREP_FLOAT || mode == PREP_LVQ_IP) {
                fv = q_simd<QT>(src, i);
            } else if (mode == PREP_LVQ_L2) {
                fv = __fsub_rn(q_simd<QT>(src, i), mean[i]);
            } else if (mode == PREP_INT) {
                if constexpr (QT == SVSB200_I8 || QT == SVSB200_U8) cv = static_cast<const uint8_t*>(src)[i];
            } else if (mode == PREP_SQ_L2) {
                // detail::compress (scalar.h:38-42)
                const float lo = code_type == SVSB200_I8 ? -128.f : 0.f, hi = code_type == SVSB200_I8 ? 127.f : 255.f;
                float r = roundf(__fdiv_rn(__fsub_rn(q_scalar<QT>(src, i), bias), scale));
                r = fminf(fmaxf(r, lo), hi);
                cv = code_type == SVSB200_I8 ? uint8_t(int8_t(int(r))) : uint8_t(int(r));
            } else {
                fv = q_scalar<QT>(src, i);
            }
        }
        f[i] = fv;
        c[i] = cv;
    }
    __syncwarp();
    if (lane != 0) return;

    float aux0 = 0.f, aux1 = 0.f;
    if (mode == PREP_INT || mode == PREP_SQ_L2) {
        int xx = 0;
        const bool is_signed = (mode == PREP_INT) ? (QT == SVSB200_I8) : (code_type == SVSB200_I8);
        for (uint32_t i = 0; i < dim; ++i) {
            int v = is_signed ? int(int8_t(c[i])) : int(c[i]);
            xx += v * v;
        }
        aux1 = __int_as_float(xx);
    }
    const bool need_norm = (metric == SVSB200_COSINE) && (mode == PREP_FLOAT || mode == PREP_INT || mode == PREP_SQ_COS);
    if (need_norm) {
        // distance::norm (distance_core.h:45-66): sequential fp32 `accum += v * v`, sqrt.
        float acc = 0.f;
        for (uint32_t i = 0; i < dim; ++i) {
            float sq;
            if constexpr (QT == SVSB200_I8 || QT == SVSB200_U8) {
                int v = QT == SVSB200_I8 ? int(static_cast<const int8_t*>(src)[i]) : int(static_cast<const uint8_t*>(src)[i]);
                sq = float(v * v);
            } else {
                float v = q_scalar<QT>(src, i);
                sq = __fmul_rn(v, v);
            }
            acc = __fadd_rn(acc, sq);
        }
        aux0 = __fsqrt_rn(acc);
    } else if (mode == PREP_LVQ_IP) {
        float acc = 0.f;
        for (uint32_t i = 0; i < dim; ++i) acc = __fmaf_rn(f[i], mean[i], acc);
        aux0 = acc;
    } else if (mode == PREP_SQ_IP) {
        // std::reduce over the fp32 query (libstdc++: four at a time, then the tail).
        float acc = 0.f;
        uint32_t i = 0;
        for (; i + 4 <= dim; i += 4) {
            float v1 = __fadd_rn(f[i], f[i + 1]);
            float v2 = __fadd_rn(f[i + 2], f[i + 3]);
            acc = __fadd_rn(acc, __fadd_rn(v1, v2));
        }
        for (; i < dim; ++i) acc = __fadd_rn(acc, f[i]);
        aux0 = __fmul_rn(bias, acc);
    }
    qaux[2 * size_t(q)] = aux0;
    qaux[2 * size_t(q) + 1] = aux1;
}

// ---------------------------------------------------------------------------------------
// Cross-shard top-k merge with TotalOrder (lib/neighbor.h:143-155): distance, then id.
// ---------------------------------------------------------------------------------------
// One warp per query selects the k smallest (key, id) pairs among all nshards * k candidates by repeated
// "smallest entry greater than the previous output" -- a full TotalOrder sort of the candidates' prefix, so
// the result does not depend on how ties are ordered inside a shard's list (they come out of the search
// buffer in insertion order, not id order).  Padding entries (id = all-ones) are ignored.
__device__ __forceinline__ uint32_t total_order_key(float d, int greater) {
    float k = greater ? -d : d;
    k = __fadd_rn(k, 0.0f);                       // -0 == +0 under operator<
    const uint32_t u = __float_as_uint(k);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // monotone float -> uint
}
__global__ void merge_topk_kernel(const uint64_t* __restrict__ ids, const float* __restrict__ dists, uint32_t nshards,
                                  uint32_t nq, uint32_t k, int greater, uint64_t* __restrict__ out_ids,
                                  float* __restrict__ out_dists) {
    const uint32_t q = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const uint32_t n = nshards * k;
    // (key, id) of the previous output; "nothing yet" sorts before everything
    bool have_last = false;
    uint32_t last_key = 0;
    uint64_t last_id = 0;
    for (uint32_t j = 0; j < k; ++j) {
        uint32_t best_key = 0xFFFFFFFFu;
        uint64_t best_id = ~uint64_t(0);
        float best_d = 0.f;
        bool found = false;
        for (uint32_t c = lane; c < n; c += 32) {
            const uint32_t sh = c / k, jj = c - sh * k;
            const size_t o = (size_t(sh) * nq + q) * k + jj;
            const uint64_t id = ids[o];
            if (id == ~uint64_t(0)) continue;
            const float d = dists[o];
            const uint32_t key = total_order_key(d, greater);
            if (have_last && (key < last_key || (key == last_key && id <= last_id))) continue;
            if (!found || key < best_key || (key == best_key && id < best_id)) {
                best_key = key;
                best_id = id;
                best_d = d;
                found = true;
            }
        }
        // warp argmin over (found, key, id)
        for (int off = 16; off; off >>= 1) {
            const uint32_t okey = __shfl_xor_sync(FULL, best_key, off);
            const uint64_t oid = __shfl_xor_sync(FULL, best_id, off);
            const float od = __shfl_xor_sync(FULL, best_d, off);
            const bool ofound = __shfl_xor_sync(FULL, int(found), off) != 0;
            if (ofound && (!found || okey < best_key || (okey == best_key && oid < best_id))) {
                best_key = okey;
                best_id = oid;
                best_d = od;
                found = true;
            }
        }
        if (lane == 0) {
            const size_t o = size_t(q) * k + j;
            out_ids[o] = found ? best_id : ~uint64_t(0);
            out_dists[o] = found ? best_d : (greater ? -INFINITY : INFINITY);
        }
        if (!found) {
            for (uint32_t r = j + 1 + lane; r < k; r += 32) {   // nothing left: pad the tail
                out_ids[size_t(q) * k + r] = ~uint64_t(0);
                out_dists[size_t(q) * k + r] = greater ? -INFINITY : INFINITY;
            }
            return;
        }
        have_last = true;
        last_key = best_key;
        last_id = best_id;
    }
}

// LVQ-8 encoder (own specification), one warp per vector:
//   r_i = x_i - mean_i;  lower = min r, upper = max r;  delta = (upper - lower) / 255
//   {delta, lower} are stored as float16 (round to nearest even) and the codes are computed against
//   the *stored* constants:  c_i = clamp(rint((r_i - lower16) / delta16), 0, 255)   (0 when delta16 == 0)
// so that decode y_i = fma(delta16, c_i, lower16) is the nearest representable grid point.
__global__ void lvq8_compress_kernel(const float* __restrict__ data, uint32_t n, uint32_t dim,
                                     const float* __restrict__ mean, uint8_t* __restrict__ rows, uint32_t stride,
                                     uint32_t const_offset) {
    const uint32_t row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const float* x = data + size_t(row) * dim;
    uint8_t* out = rows + size_t(row) * stride;
    float lo = INFINITY, hi = -INFINITY;
    for (uint32_t i = lane; i < dim; i += 32) {
        const float r = __fsub_rn(x[i], mean[i]);
        lo = fminf(lo, r);
        hi = fmaxf(hi, r);
    }
    for (int o = 16; o; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, o));
    }
    const __half dh = __float2half_rn(__fdiv_rn(__fsub_rn(hi, lo), 255.0f));
    const __half lh = __float2half_rn(lo);
    const float d = __half2float(dh), l = __half2float(lh);
    for (uint32_t i = lane; i < stride; i += 32) {
        uint8_t c = 0;
        if (i < dim && d > 0.0f) {
            const float r = __fsub_rn(x[i], mean[i]);
            float q = rintf(__fdiv_rn(__fsub_rn(r, l), d));
            q = fminf(fmaxf(q, 0.0f), 255.0f);
            c = uint8_t(int(q));
        }
        if (i < const_offset || i >= const_offset + 4) out[i] = c;
    }
    if (lane == 0) {
        __half2 h = __halves2half2(dh, lh);
        *reinterpret_cast<__half2*>(out + const_offset) = h;
    }
}

template <int QT>
static cudaError_t launch_prepare(const void* d_queries, uint32_t nq, uint32_t dim, uint32_t qstride, int mode, int metric,
                                  int code_type, float scale, float bias, const float* mean, float* qf, uint8_t* qcodes,
                                  float* qaux, cudaStream_t stream) {
    const int warps = 8;
    const unsigned grid = (nq + warps - 1) / warps;
    prepare_queries_kernel<QT><<<grid, warps * 32, 0, stream>>>(d_queries, nq, dim, qstride, mode, metric, code_type, scale,
                                                               bias, mean, qf, qcodes, qaux);
    count_launch();
    return cudaGetLastError();
}

}  // namespace svsb200

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

const char* svsb200_last_error(void) { return g_error.c_str(); }
int svsb200_version(void) { return SVSB200_VERSION; }
uint64_t svsb200_launch_count(void) { return g_launches.load(); }

int svsb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int svsb200_device_sm(int device, int* sm) {
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (sm) *sm = prop.major * 10 + prop.minor;
    return 0;
}

int svsb200_index_create(const void* vectors, int dtype, size_t n, size_t dim, size_t row_stride_bytes,
                         const uint32_t* graph_rows, size_t graph_row_len, uint32_t entry_point, int metric, int storage,
                         const float* aux, int device, svsb200_index** out) {
    if (!out) return fail("svsb200_index_create: out is NULL");
    *out = nullptr;
    if (!vectors || !graph_rows) return fail("svsb200_index_create: NULL input");
    if (dtype < SVSB200_F32 || dtype > SVSB200_U8) return fail("svsb200_index_create: bad dtype");
    if (metric < SVSB200_L2 || metric > SVSB200_COSINE) return fail("svsb200_index_create: bad metric");
    if (n == 0 || dim == 0) return fail("svsb200_index_create: empty dataset");
    if (n >= (size_t(1) << 31)) return fail("svsb200_index_create: more than 2^31-1 vectors per index");
    if (graph_row_len < 2) return fail("svsb200_index_create: graph rows need a degree word and one slot");
    if (graph_row_len > 65536) return fail("svsb200_index_create: max_degree above 65535");
    if (entry_point >= n) return fail("svsb200_index_create: entry point out of range");
    if (storage == SVSB200_SQ) {
        if (dtype != SVSB200_I8 && dtype != SVSB200_U8) return fail("svsb200_index_create: SQ codes must be int8/uint8");
        if (!aux) return fail("svsb200_index_create: SQ needs aux = {scale, bias}");
    } else if (storage == SVSB200_LVQ8) {
        if (dtype != SVSB200_U8) return fail("svsb200_index_create: LVQ-8 rows are uint8 codes (dtype SVSB200_U8)");
        if (!aux) return fail("svsb200_index_create: LVQ-8 needs aux = mean[dim]");
        if (row_stride_bytes != svsb200_lvq8_row_stride(dim))
            return fail("svsb200_index_create: LVQ-8 rows must use svsb200_lvq8_row_stride(dim)");
    } else if (storage != SVSB200_PLAIN) {
        return fail("svsb200_index_create: unsupported storage kind");
    }
    int ndev = svsb200_device_count();
    if (ndev == 0) return fail("svsb200_index_create: no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return fail("svsb200_index_create: bad device ordinal");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10 || prop.minor != 0)
        return fail("svsb200_index_create: device is not sm_100 (this binary holds sm_100a code only)");

    auto* ix = new svsb200_index();
    ix->device = device;
    ix->sm_count = prop.multiProcessorCount;
    ix->dtype = dtype;
    ix->metric = metric;
    ix->storage = storage;
    ix->n = n;
    ix->dim = dim;
    ix->max_degree = graph_row_len - 1;
    ix->entry_point = entry_point;
    if (storage == SVSB200_SQ) {
        ix->scale = aux[0];
        ix->bias = aux[1];
    }
    size_t row_bytes = dim * esize(dtype);
    if (storage == SVSB200_LVQ8) {
        ix->lvq_const_offset = uint32_t(round_up(dim, 4));
        row_bytes = ix->lvq_const_offset + 4;
    }
    const size_t src_stride = row_stride_bytes ? row_stride_bytes : row_bytes;
    ix->row_stride = uint32_t(storage == SVSB200_LVQ8 ? svsb200_lvq8_row_stride(dim) : round_up(row_bytes, 16));
    // rows of up to 128 neighbours are padded to whole 32-word groups (one coalesced load per group and lane in
    // the lean kernel, no per-lane bounds checks); wider rows stay 16-byte aligned only
    ix->gstride = uint32_t(round_up(ix->max_degree, ix->max_degree <= 32u * kFastMaxGW ? 32 : 4));

    auto cleanup = [&](int rc) {
        svsb200_index_destroy(ix);
        return rc;
    };
#define CUDA_TRY_IX(expr)                                                                      \
    do {                                                                                       \
        cudaError_t err__ = (expr);                                                            \
        if (err__ != cudaSuccess) {                                                            \
            fail(std::string(#expr) + ": " + cudaGetErrorString(err__));                       \
            return cleanup(1);                                                                 \
        }                                                                                      \
    } while (0)

    const size_t vbytes = n * size_t(ix->row_stride);
    const size_t gbytes = n * size_t(ix->gstride) * sizeof(uint32_t);
    CUDA_TRY_IX(cudaMalloc(&ix->d_vectors, vbytes));
    CUDA_TRY_IX(cudaMalloc(&ix->d_graph, gbytes));
    CUDA_TRY_IX(cudaMalloc(&ix->d_ref_degree, n * sizeof(uint16_t)));
    ix->device_bytes = vbytes + gbytes + n * sizeof(uint16_t);
    CUDA_TRY_IX(cudaMemset(ix->d_vectors, 0, vbytes));
    CUDA_TRY_IX(cudaMemcpy2D(ix->d_vectors, ix->row_stride, vectors, src_stride, row_bytes, n, cudaMemcpyHostToDevice));
    {
        uint32_t* d_src = nullptr;
        int* d_bad = nullptr;
        const size_t sbytes = n * graph_row_len * sizeof(uint32_t);
        CUDA_TRY_IX(cudaMalloc(&d_src, sbytes));
        cudaError_t err = cudaMalloc(&d_bad, sizeof(int));
        if (err == cudaSuccess) err = cudaMemset(d_bad, 0, sizeof(int));
        if (err == cudaSuccess) err = cudaMemcpy(d_src, graph_rows, sbytes, cudaMemcpyHostToDevice);
        int bad = 0;
        if (err == cudaSuccess) {
            const int warps = 8;
            repack_graph_kernel<<<unsigned((n + warps - 1) / warps), warps * 32>>>(d_src, graph_row_len, uint32_t(n),
                                                                                  ix->d_graph, ix->gstride,
                                                                                  ix->d_ref_degree, d_bad);
            count_launch();
            err = cudaGetLastError();
        }
        if (err == cudaSuccess) err = cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost);
        cudaFree(d_src);
        if (d_bad) cudaFree(d_bad);
        CUDA_TRY_IX(err);
        if (bad) {
            fail(bad == 1 ? "svsb200_index_create: adjacency row with degree > max_degree"
                          : "svsb200_index_create: neighbour id out of range");
            return cleanup(1);
        }
    }
    if (storage == SVSB200_LVQ8) {
        CUDA_TRY_IX(cudaMalloc(&ix->d_mean, dim * sizeof(float)));
        CUDA_TRY_IX(cudaMemcpy(ix->d_mean, aux, dim * sizeof(float), cudaMemcpyHostToDevice));
    }
    CUDA_TRY_IX(cudaMalloc(&ix->d_counter, sizeof(unsigned int)));
    CUDA_TRY_IX(cudaStreamCreateWithFlags(&ix->own_stream, cudaStreamNonBlocking));
    CUDA_TRY_IX(cudaEventCreate(&ix->ev_start));
    CUDA_TRY_IX(cudaEventCreate(&ix->ev_stop));
#undef CUDA_TRY_IX
    *out = ix;
    return 0;
}

int svsb200_index_destroy(svsb200_index* ix) {
    if (!ix) return 0;
    cudaSetDevice(ix->device);
    if (ix->d_vectors) cudaFree(ix->d_vectors);
    if (ix->d_graph) cudaFree(ix->d_graph);
    if (ix->d_ref_degree) cudaFree(ix->d_ref_degree);
    if (ix->d_counter) cudaFree(ix->d_counter);
    if (ix->d_mean) cudaFree(ix->d_mean);
    ix->q_raw.release();
    ix->q_codes.release();
    ix->ids.release();
    ix->q_f32.release();
    ix->q_aux.release();
    ix->dists.release();
    ix->hops.release();
    ix->evals.release();
    ix->fetched.release();
    if (ix->own_stream) cudaStreamDestroy(ix->own_stream);
    if (ix->ev_start) cudaEventDestroy(ix->ev_start);
    if (ix->ev_stop) cudaEventDestroy(ix->ev_stop);
    delete ix;
    return 0;
}

size_t svsb200_index_size(const svsb200_index* ix) { return ix ? ix->n : 0; }
size_t svsb200_index_dimensions(const svsb200_index* ix) { return ix ? ix->dim : 0; }
size_t svsb200_index_max_degree(const svsb200_index* ix) { return ix ? ix->max_degree : 0; }
size_t svsb200_index_device_bytes(const svsb200_index* ix) { return ix ? ix->device_bytes : 0; }
int svsb200_index_device(const svsb200_index* ix) { return ix ? ix->device : -1; }

int svsb200_set_counting(svsb200_index* ix, int enabled) {
    if (!ix) return fail("svsb200_set_counting: NULL index");
    ix->counting = enabled;
    return 0;
}

int svsb200_set_option(svsb200_index* ix, const char* name, long value) {
    if (!ix || !name) return fail("svsb200_set_option: NULL argument");
    const std::string key(name);
    if (key == "warps_per_cta") {
        if (value < 0 || value > 8) return fail("warps_per_cta must be in [0, 8]");
        ix->warps_per_cta = value;
    } else if (key == "ctas_per_sm") {
        if (value < 0 || value > 32) return fail("ctas_per_sm must be in [0, 32]");
        ix->ctas_per_sm = value;
    } else if (key == "rows_in_flight") {
        if (value < 0 || value > 2) return fail("rows_in_flight must be in [0, 2]");
        ix->rows_in_flight = value;
    } else if (key == "no_split") {
        ix->no_split = value;
    } else if (key == "filter_tag16") {
        ix->filter_tag16 = value;
    } else if (key == "generic_kernel") {
        ix->generic_kernel = value;
    } else if (key == "visited_filter_slots") {
        // -1 = default; 0 = off; otherwise a power of two
        if (value > 0 && (value & (value - 1))) return fail("visited_filter_slots must be a power of two");
        // the filter words sit in front of 16-byte aligned arrays in shared memory
        if (value > 0 && value < 8) return fail("visited_filter_slots must be 0 (off) or at least 8");
        if (value > 16384) return fail("visited_filter_slots must be <= 16384");
        ix->filter_slots = value;
    } else {
        return fail("svsb200_set_option: unknown option " + key);
    }
    return 0;
}

int svsb200_get_option(svsb200_index* ix, const char* name, long* value) {
    if (!ix || !name || !value) return fail("svsb200_get_option: NULL argument");
    const std::string key(name);
    if (key == "last_kernel") *value = ix->last_kernel;          // 1 = lean kernel, 0 = generic kernel
    else if (key == "warps_per_cta") *value = ix->warps_per_cta;
    else if (key == "ctas_per_sm") *value = ix->ctas_per_sm;
    else if (key == "rows_in_flight") *value = ix->rows_in_flight;
    else if (key == "visited_filter_slots") *value = ix->filter_slots;
    else if (key == "generic_kernel") *value = ix->generic_kernel;
    else return fail("svsb200_get_option: unknown option " + key);
    return 0;
}

// Shared body of svsb200_search / svsb200_search_device: everything on the device.
static int search_on_device(svsb200_index* ix, const void* d_queries, int qdtype, size_t nq, size_t k, size_t window,
                            size_t capacity, void* d_out_ids, int id_bytes, float* d_out_dists, cudaStream_t stream,
                            bool exhaustive = false) {
    if (id_bytes != 4 && id_bytes != 8) return fail("id_bytes must be 4 or 8");
    if (qdtype < SVSB200_F32 || qdtype > SVSB200_U8) return fail("bad query dtype");
    if (window > capacity) {
        // SearchBufferConfig::check_invariants (search_buffer.h:87-96)
        return fail("Improper configuration for search buffer! search window size cannot exceed capacity");
    }
    if (capacity < k) window = capacity = k;   // index/vamana/index.h:590-592
    if (capacity == 0) return fail("search buffer capacity is zero");
    if (nq == 0) return 0;
    if (nq >= (size_t(1) << 31)) return fail("too many queries in one batch");

    // Which (query, data) pairs exist, mirroring the SIMD specialisations
    // (euclidean.h:293-358) and the SQ CPOs (extensions/vamana/scalar.h:32-43).
    int op, mode;
    const int metric = ix->metric;
    if (ix->storage == SVSB200_LVQ8) {
        if (qdtype != SVSB200_F32 && qdtype != SVSB200_F16) return fail("LVQ-8 datasets take float32/float16 queries");
        if (metric == SVSB200_COSINE) return fail("LVQ-8: cosine is not supported (L2 and MIP are)");
        op = metric == SVSB200_L2 ? OP_L2F : OP_IPF;
        mode = metric == SVSB200_L2 ? PREP_LVQ_L2 : PREP_LVQ_IP;
    } else if (ix->storage == SVSB200_SQ) {
        if (qdtype != SVSB200_F32 && qdtype != SVSB200_F16) return fail("SQ datasets take float32/float16 queries");
        op = metric == SVSB200_L2 ? OP_L2I : metric == SVSB200_IP ? OP_IPF : OP_COSF;
        mode = metric == SVSB200_L2 ? PREP_SQ_L2 : metric == SVSB200_IP ? PREP_SQ_IP : PREP_SQ_COS;
    } else if (qdtype == SVSB200_I8 || qdtype == SVSB200_U8) {
        if (qdtype != ix->dtype) return fail("int8/uint8 queries need a dataset of the same type");
        op = metric == SVSB200_L2 ? OP_L2I : metric == SVSB200_IP ? OP_IPI : OP_COSI;
        mode = PREP_INT;
    } else {
        if (qdtype == SVSB200_F16 && ix->dtype != SVSB200_F32 && ix->dtype != SVSB200_F16)
            return fail("float16 queries need a float32/float16 dataset");
        op = metric == SVSB200_L2 ? OP_L2F : metric == SVSB200_IP ? OP_IPF : OP_COSF;
        mode = PREP_FLOAT;
    }

    const uint32_t qstride = uint32_t(round_up(ix->dim, 16));
    CUDA_TRY(ix->q_f32.ensure(nq * qstride));
    CUDA_TRY(ix->q_codes.ensure(nq * qstride));
    CUDA_TRY(ix->q_aux.ensure(nq * 2));
    if (ix->counting) {
        CUDA_TRY(ix->hops.ensure(nq));
        CUDA_TRY(ix->evals.ensure(nq));
        CUDA_TRY(ix->fetched.ensure(nq));
        ix->counted_nq = nq;
    }

    cudaError_t err;
    switch (qdtype) {
        case SVSB200_F32:
            err = launch_prepare<SVSB200_F32>(d_queries, uint32_t(nq), uint32_t(ix->dim), qstride, mode, metric, ix->dtype,
                                              ix->scale, ix->bias, ix->d_mean, ix->q_f32.ptr, ix->q_codes.ptr, ix->q_aux.ptr, stream);
            break;
        case SVSB200_F16:
            err = launch_prepare<SVSB200_F16>(d_queries, uint32_t(nq), uint32_t(ix->dim), qstride, mode, metric, ix->dtype,
                                              ix->scale, ix->bias, ix->d_mean, ix->q_f32.ptr, ix->q_codes.ptr, ix->q_aux.ptr, stream);
            break;
        case SVSB200_I8:
            err = launch_prepare<SVSB200_I8>(d_queries, uint32_t(nq), uint32_t(ix->dim), qstride, mode, metric, ix->dtype,
                                             ix->scale, ix->bias, ix->d_mean, ix->q_f32.ptr, ix->q_codes.ptr, ix->q_aux.ptr, stream);
            break;
        default:
            err = launch_prepare<SVSB200_U8>(d_queries, uint32_t(nq), uint32_t(ix->dim), qstride, mode, metric, ix->dtype,
                                             ix->scale, ix->bias, ix->d_mean, ix->q_f32.ptr, ix->q_codes.ptr, ix->q_aux.ptr, stream);
    }
    CUDA_TRY(err);
    CUDA_TRY(cudaMemsetAsync(ix->d_counter, 0, sizeof(unsigned int), stream));

    SearchParams p{};
    p.vectors = ix->d_vectors;
    p.graph = ix->d_graph;
    p.ref_degree = ix->d_ref_degree;
    p.n = uint32_t(ix->n);
    p.dim = uint32_t(ix->dim);
    p.row_stride = ix->row_stride;
    p.gstride = ix->gstride;
    p.entry_point = ix->entry_point;
    p.greater = metric != SVSB200_L2;
    p.sq = ix->storage == SVSB200_SQ;
    p.lvq = ix->storage == SVSB200_LVQ8;
    p.no_split = int(ix->no_split);
    p.lvq_const_offset = ix->lvq_const_offset;
    p.scale = ix->scale;
    p.bias = ix->bias;
    p.scale_sq = ix->scale * ix->scale;   // EuclideanCompressed ctor (scalar.h:68-72)
    p.qf = ix->q_f32.ptr;
    p.qcodes = ix->q_codes.ptr;
    p.qaux = ix->q_aux.ptr;
    p.qstride = qstride;
    p.nq = uint32_t(nq);
    p.k = uint32_t(k);
    p.window = uint32_t(window);
    p.capacity = uint32_t(capacity);
    p.cap_pad = uint32_t(round_up(capacity + 1, 32));
    p.deg_pad = uint32_t(round_up(ix->gstride, 32));
    p.out_ids = d_out_ids;
    p.id_bytes = id_bytes;
    p.out_dists = d_out_dists;
    p.work_counter = ix->d_counter;
    p.hops = ix->counting ? ix->hops.ptr : nullptr;
    p.evals = ix->counting ? ix->evals.ptr : nullptr;
    p.fetched = ix->counting ? ix->fetched.ptr : nullptr;
    p.filter_slots = exhaustive ? 0u : (ix->filter_slots < 0 ? 4096u : uint32_t(ix->filter_slots));
    // 16-bit tags (two per 32-bit set, 2-way LRU) are exact as long as every id >> log2(sets) fits below
    // the 0xFFFF "empty" mark; larger indexes fall back to direct-mapped 32-bit entries.
    p.filter_shift = 0;
    while ((2u << p.filter_shift) < p.filter_slots) ++p.filter_shift;   // log2(sets) with sets = slots / 2
    p.filter_tag16 = p.filter_slots >= 2 && ((uint64_t(ix->n - 1) >> p.filter_shift) < 0xFFFFull) && ix->filter_tag16 != 0;
    if (!p.filter_tag16) {
        p.filter_shift = 0;
        while ((1u << p.filter_shift) < p.filter_slots) ++p.filter_shift;
    }

    LaunchConfig cfg{};
    const size_t smem_limit = 227 * 1024;
    // The lean kernel (search_fast.cuh) covers the common shape: 16-bit-tag filter on, adjacency rows of up to
    // 128 neighbours; anything else (and the exhaustive scan) runs on the generic kernel.
    const uint32_t fast_cap_pad = uint32_t(round_up(capacity, 32));
    // its filter: sets of eight 16-bit tags; default 256 sets (4 KB), more when n needs it for exactness
    uint32_t fast_slots = ix->filter_slots < 0 ? 2048u : uint32_t(ix->filter_slots);
    if (fast_slots && fast_slots < 64) fast_slots = 64;
    uint32_t fast_shift = 0;
    while ((8u << fast_shift) < fast_slots) ++fast_shift;                       // log2(sets)
    while (fast_slots && (uint64_t(ix->n - 1) >> fast_shift) >= 0xFFFFull && fast_shift < 20) {
        ++fast_shift;
        fast_slots <<= 1;
    }
    const size_t fast_bytes = fast_smem_bytes(p.qstride, fast_cap_pad, p.deg_pad, fast_slots * 2u);
    const bool use_fast = !exhaustive && !ix->generic_kernel && fast_slots >= 64 && ix->filter_tag16 &&
                          (uint64_t(ix->n - 1) >> fast_shift) < 0xFFFFull && p.deg_pad <= 32u * kFastMaxGW &&
                          p.gstride % 32u == 0 &&
                          fast_bytes <= smem_limit;
    ix->last_kernel = use_fast ? 1 : 0;
    const int nrows = ix->rows_in_flight ? int(ix->rows_in_flight) : 2;
    if (use_fast) {
        p.cap_pad = fast_cap_pad;
        p.filter_slots = fast_slots;
        p.filter_shift = fast_shift;
        p.filter_tag16 = 1;
        cfg.warps_per_cta = 1;
        cfg.smem_bytes = fast_bytes;
        cfg.stream = stream;
        cfg.grid = ix->ctas_per_sm ? ix->sm_count * int(ix->ctas_per_sm) : -ix->sm_count;
    } else {
        const size_t per_warp = warp_smem_bytes(p.qstride, p.cap_pad, p.deg_pad, p.filter_slots * (p.filter_tag16 ? 2u : 4u));
        int warps = ix->warps_per_cta ? int(ix->warps_per_cta) : 4;
        while (warps > 1 && per_warp * warps > smem_limit) warps >>= 1;
        if (per_warp * warps > smem_limit) return fail("search buffer capacity too large for shared memory");
        cfg.warps_per_cta = warps;
        cfg.smem_bytes = per_warp * warps;
        cfg.stream = stream;
        // grid: persistent CTAs; the launcher clamps to what is resident.  ctas_per_sm == 0
        // means "as many as fit" (computed by the launcher through the occupancy API).
        cfg.grid = ix->sm_count * (ix->ctas_per_sm ? int(ix->ctas_per_sm) : 0);
        if (cfg.grid == 0 || exhaustive) cfg.grid = -ix->sm_count;   // negative: launcher multiplies by occupancy
    }

    CUDA_TRY(cudaEventRecord(ix->ev_start, stream));
    if (exhaustive) {
        switch (ix->storage == SVSB200_LVQ8 ? ROW_LVQ8 : ix->dtype) {
            case ROW_LVQ8: err = launch_search_exhaustive<ROW_LVQ8>(op, p, cfg); break;
            case SVSB200_F32: err = launch_search_exhaustive<SVSB200_F32>(op, p, cfg); break;
            case SVSB200_F16: err = launch_search_exhaustive<SVSB200_F16>(op, p, cfg); break;
            case SVSB200_I8: err = launch_search_exhaustive<SVSB200_I8>(op, p, cfg); break;
            default: err = launch_search_exhaustive<SVSB200_U8>(op, p, cfg);
        }
        CUDA_TRY(err);
        CUDA_TRY(cudaEventRecord(ix->ev_stop, stream));
        ix->timed = true;
        return 0;
    }
    if (use_fast) {
        switch (ix->storage == SVSB200_LVQ8 ? ROW_LVQ8 : ix->dtype) {
            case ROW_LVQ8: err = launch_search_fast<ROW_LVQ8>(op, p, cfg); break;
            case SVSB200_F32: err = launch_search_fast<SVSB200_F32>(op, p, cfg); break;
            case SVSB200_F16: err = launch_search_fast<SVSB200_F16>(op, p, cfg); break;
            case SVSB200_I8: err = launch_search_fast<SVSB200_I8>(op, p, cfg); break;
            default: err = launch_search_fast<SVSB200_U8>(op, p, cfg);
        }
    } else
    switch (ix->storage == SVSB200_LVQ8 ? ROW_LVQ8 : ix->dtype) {
        case ROW_LVQ8: err = launch_search<ROW_LVQ8>(op, p, cfg, nrows); break;
        case SVSB200_F32: err = launch_search<SVSB200_F32>(op, p, cfg, nrows); break;
        case SVSB200_F16: err = launch_search<SVSB200_F16>(op, p, cfg, nrows); break;
        case SVSB200_I8: err = launch_search<SVSB200_I8>(op, p, cfg, nrows); break;
        default: err = launch_search<SVSB200_U8>(op, p, cfg, nrows);
    }
    CUDA_TRY(err);
    CUDA_TRY(cudaEventRecord(ix->ev_stop, stream));
    ix->timed = true;
    return 0;
}

int svsb200_search_device(svsb200_index* ix, const void* d_queries, int qdtype, size_t nq, size_t k, size_t window,
                          size_t capacity, int use_visited_set, void* d_out_ids, int id_bytes, float* d_out_dists,
                          void* stream) {
    (void)use_visited_set;   // performance-only in the reference (search_buffer.h:420); results identical
    if (!ix) return fail("svsb200_search_device: NULL index");
    if (nq && (!d_queries || !d_out_ids || !d_out_dists)) return fail("svsb200_search_device: NULL buffer");
    std::lock_guard<std::mutex> lock(ix->mutex);
    CUDA_TRY(cudaSetDevice(ix->device));
    return search_on_device(ix, d_queries, qdtype, nq, k, window, capacity, d_out_ids, id_bytes, d_out_dists,
                            stream ? static_cast<cudaStream_t>(stream) : ix->own_stream);
}

int svsb200_search(svsb200_index* ix, const void* queries, int qdtype, size_t nq, size_t k, size_t window,
                   size_t capacity, int use_visited_set, void* out_ids, int id_bytes, float* out_dists, void* stream_) {
    (void)use_visited_set;
    if (!ix) return fail("svsb200_search: NULL index");
    if (nq == 0) return 0;
    if (!queries || !out_ids || !out_dists) return fail("svsb200_search: NULL buffer");
    if (qdtype < SVSB200_F32 || qdtype > SVSB200_U8) return fail("bad query dtype");
    if (id_bytes != 4 && id_bytes != 8) return fail("id_bytes must be 4 or 8");
    std::lock_guard<std::mutex> lock(ix->mutex);
    CUDA_TRY(cudaSetDevice(ix->device));
    cudaStream_t stream = stream_ ? static_cast<cudaStream_t>(stream_) : ix->own_stream;
    const size_t qbytes = nq * ix->dim * esize(qdtype);
    CUDA_TRY(ix->q_raw.ensure(qbytes));
    CUDA_TRY(ix->ids.ensure(nq * k * size_t(id_bytes)));
    CUDA_TRY(ix->dists.ensure(nq * k));
    CUDA_TRY(cudaMemcpyAsync(ix->q_raw.ptr, queries, qbytes, cudaMemcpyHostToDevice, stream));
    int rc = search_on_device(ix, ix->q_raw.ptr, qdtype, nq, k, window, capacity, ix->ids.ptr, id_bytes, ix->dists.ptr,
                              stream);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(out_ids, ix->ids.ptr, nq * k * size_t(id_bytes), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(out_dists, ix->dists.ptr, nq * k * sizeof(float), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return 0;
}

int svsb200_get_fetched(svsb200_index* ix, size_t nq, uint32_t* fetched) {
    if (!ix || !fetched) return fail("svsb200_get_fetched: NULL argument");
    if (!ix->counting || ix->counted_nq < nq) return fail("svsb200_get_fetched: counting was not enabled for that many queries");
    CUDA_TRY(cudaSetDevice(ix->device));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(fetched, ix->fetched.ptr, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return 0;
}

int svsb200_get_counters(svsb200_index* ix, size_t nq, uint32_t* hops, uint32_t* evals) {
    if (!ix) return fail("svsb200_get_counters: NULL index");
    if (!ix->counting || ix->counted_nq < nq) return fail("svsb200_get_counters: counting was not enabled for that many queries");
    CUDA_TRY(cudaSetDevice(ix->device));
    CUDA_TRY(cudaDeviceSynchronize());
    if (hops) CUDA_TRY(cudaMemcpy(hops, ix->hops.ptr, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (evals) CUDA_TRY(cudaMemcpy(evals, ix->evals.ptr, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return 0;
}

int svsb200_last_kernel_ms(svsb200_index* ix, float* ms) {
    if (!ix || !ms) return fail("svsb200_last_kernel_ms: NULL argument");
    if (!ix->timed) return fail("svsb200_last_kernel_ms: no search has run yet");
    CUDA_TRY(cudaSetDevice(ix->device));
    CUDA_TRY(cudaEventSynchronize(ix->ev_stop));
    CUDA_TRY(cudaEventElapsedTime(ms, ix->ev_start, ix->ev_stop));
    return 0;
}

int svsb200_merge_topk_device(const uint64_t* d_ids, const float* d_dists, size_t nshards, size_t nq, size_t k, int metric,
                              uint64_t* d_out_ids, float* d_out_dists, int device, void* stream) {
    if (nshards == 0 || nshards > 1024) return fail("svsb200_merge_topk_device: 1..1024 shards supported");
    if (nq == 0 || k == 0) return 0;
    CUDA_TRY(cudaSetDevice(device));
    const unsigned warps = 4;
    merge_topk_kernel<<<unsigned((nq + warps - 1) / warps), warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        d_ids, d_dists, uint32_t(nshards), uint32_t(nq), uint32_t(k), metric != SVSB200_L2, d_out_ids, d_out_dists);
    count_launch();
    CUDA_TRY(cudaGetLastError());
    return 0;
}

size_t svsb200_lvq8_row_stride(size_t dim) { return round_up(round_up(dim, 4) + 4, 32); }

int svsb200_lvq8_compress(const float* data, size_t n, size_t dim, const float* mean, void* out_rows, int device) {
    if (!data || !mean || !out_rows) return fail("svsb200_lvq8_compress: NULL argument");
    if (n == 0 || dim == 0 || n >= (size_t(1) << 31)) return fail("svsb200_lvq8_compress: bad shape");
    if (svsb200_device_count() == 0) return fail("svsb200_lvq8_compress: no CUDA device (there is no CPU fallback)");
    CUDA_TRY(cudaSetDevice(device));
    const size_t stride = svsb200_lvq8_row_stride(dim);
    float *d_data = nullptr, *d_mean = nullptr;
    uint8_t* d_rows = nullptr;
    cudaError_t err = cudaMalloc(&d_data, n * dim * sizeof(float));
    if (err == cudaSuccess) err = cudaMalloc(&d_mean, dim * sizeof(float));
    if (err == cudaSuccess) err = cudaMalloc(&d_rows, n * stride);
    if (err == cudaSuccess) err = cudaMemcpy(d_data, data, n * dim * sizeof(float), cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMemcpy(d_mean, mean, dim * sizeof(float), cudaMemcpyHostToDevice);
    if (err == cudaSuccess) {
        const int warps = 8;
        lvq8_compress_kernel<<<unsigned((n + warps - 1) / warps), warps * 32>>>(d_data, uint32_t(n), uint32_t(dim), d_mean,
                                                                              d_rows, uint32_t(stride),
                                                                              uint32_t(round_up(dim, 4)));
        count_launch();
        err = cudaGetLastError();
    }
    if (err == cudaSuccess) err = cudaMemcpy(out_rows, d_rows, n * stride, cudaMemcpyDeviceToHost);
    if (d_data) cudaFree(d_data);
    if (d_mean) cudaFree(d_mean);
    if (d_rows) cudaFree(d_rows);
    CUDA_TRY(err);
    return 0;
}

int svsb200_exhaustive_device(svsb200_index* ix, const void* d_queries, int qdtype, size_t nq, size_t k, uint64_t* d_out_ids,
                              float* d_out_dists, void* stream) {
    if (!ix) return fail("svsb200_exhaustive_device: NULL index");
    if (nq && (!d_queries || !d_out_ids || !d_out_dists)) return fail("svsb200_exhaustive_device: NULL buffer");
    if (k == 0 || k > 1024) return fail("svsb200_exhaustive_device: k must be in [1, 1024]");
    std::lock_guard<std::mutex> lock(ix->mutex);
    CUDA_TRY(cudaSetDevice(ix->device));
    return search_on_device(ix, d_queries, qdtype, nq, k, k, k, d_out_ids, 8, d_out_dists,
                            stream ? static_cast<cudaStream_t>(stream) : ix->own_stream, true);
}

}  // extern "C"
