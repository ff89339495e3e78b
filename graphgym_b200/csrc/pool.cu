// Graph-level pooling of node rows: global_add/mean/max_pool of
// graphgym/models/pooling.py:12-33 — scatter(x[id], batch[id], reduce) with torch_scatter in the reference.
// DeepSNAP batches are block-diagonal, so `batch` is non-decreasing: the segments are found by binary search
// (no sort), every (graph, column) is reduced by one thread walking its rows in order — deterministic, unlike the
// reference's atomic scatter — and adjacent threads read adjacent columns (coalesced).  max keeps the arg-max row
// for the backward.
#include <float.h>

#include "common.cuh"

namespace gg {

// seg_ptr[g] = first position p with key[p] >= g (key non-decreasing, values in [0, G)); flag counts violations
__global__ void __launch_bounds__(256) segment_bounds_kernel(const int64_t* __restrict__ key, int64_t m, int64_t G,
                                                             int32_t* __restrict__ seg_ptr, int32_t* __restrict__ flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g <= G; g += stride) {
        int64_t lo = 0, hi = m;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (key[mid] < g) lo = mid + 1;
            else hi = mid;
        }
        seg_ptr[g] = (int32_t)lo;
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const int64_t k = key[i];
        if (k < 0 || k >= G || (i + 1 < m && key[i + 1] < k)) atomicAdd(flag, 1);
    }
}

__global__ void __launch_bounds__(256) segment_pool_kernel(const float* __restrict__ x, int64_t ldx,
                                                           const int64_t* __restrict__ row_index,
                                                           const int32_t* __restrict__ seg_ptr, int64_t G, int64_t f,
                                                           int mode, float* __restrict__ out, int64_t ldo,
                                                           int32_t* __restrict__ argmax) {
    const int64_t total = G * f;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = e / f, c = e - g * f;
        const int beg = seg_ptr[g], end = seg_ptr[g + 1];
        float acc = mode == GG_POOL_MAX ? -FLT_MAX : 0.f;
        int best = -1;
        for (int p = beg; p < end; ++p) {
            const int64_t r = row_index ? row_index[p] : p;
            const float v = __ldg(x + r * ldx + c);
            if (mode == GG_POOL_MAX) {
                if (v > acc || best < 0) { acc = v; best = p; }
            } else {
                acc += v;
            }
        }
        if (mode == GG_POOL_MEAN && end > beg) acc /= (float)(end - beg);
        if (mode == GG_POOL_MAX && best < 0) acc = 0.f;  // empty graph: torch_scatter leaves the zero-initialised row
        out[g * ldo + c] = acc;
        if (argmax) argmax[e] = best;
    }
}

// gx[row(p), c] = d out[key[p], c] / d x: sum -> g, mean -> g / count, max -> g where p is the arg-max
__global__ void __launch_bounds__(256) segment_pool_bwd_kernel(const float* __restrict__ g, int64_t ldg,
                                                               const int64_t* __restrict__ row_index,
                                                               const int64_t* __restrict__ key,
                                                               const int32_t* __restrict__ seg_ptr, int64_t m, int64_t f,
                                                               int mode, const int32_t* __restrict__ argmax,
                                                               float* __restrict__ gx, int64_t ldgx) {
    const int64_t total = m * f;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = e / f, c = e - p * f;
        const int64_t gr = key[p];
        float v = __ldg(g + gr * ldg + c);
        if (mode == GG_POOL_MEAN) v /= (float)(seg_ptr[gr + 1] - seg_ptr[gr]);
        if (mode == GG_POOL_MAX && argmax[gr * f + c] != (int32_t)p) v = 0.f;
        const int64_t r = row_index ? row_index[p] : p;
        gx[r * ldgx + c] = v;   // rows are distinct (node_id_index is a set of rows): plain store
    }
}

static inline int pool_grid(int64_t total) {
    int64_t b = ceil_div(total, 256);
    if (b > (int64_t)kNumSMs * 16) b = (int64_t)kNumSMs * 16;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace gg

using namespace gg;

extern "C" {

int gg_segment_bounds_i64(const int64_t* key, int64_t m, int64_t num_segments, int32_t* seg_ptr,
                          int32_t* violations, gg_stream_t stream) {
    GG_REQUIRE(m >= 0 && num_segments >= 0 && m < ((int64_t)1 << 31), "gg_segment_bounds_i64: size out of range");
    GG_REQUIRE(seg_ptr && violations && (m == 0 || key), "gg_segment_bounds_i64: null pointer");
    cudaStream_t st = as_stream(stream);
    GG_CUDA(cudaMemsetAsync(violations, 0, sizeof(int32_t), st));
    const int64_t work = m > num_segments + 1 ? m : num_segments + 1;
    segment_bounds_kernel<<<pool_grid(work), 256, 0, st>>>(key, m, num_segments, seg_ptr, violations);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_segment_pool_f32(const float* x, int64_t ldx, const int64_t* row_index, const int32_t* seg_ptr,
                        int64_t num_segments, int64_t f, int mode, float* out, int64_t ldo, int32_t* argmax,
                        gg_stream_t stream) {
    GG_REQUIRE(num_segments >= 0 && f >= 0, "gg_segment_pool_f32: negative size");
    GG_REQUIRE(mode == GG_POOL_SUM || mode == GG_POOL_MEAN || mode == GG_POOL_MAX, "gg_segment_pool_f32: mode=%d", mode);
    if (num_segments == 0 || f == 0) return GG_OK;
    GG_REQUIRE(x && seg_ptr && out && ldx >= f && ldo >= f, "gg_segment_pool_f32: bad operands");
    segment_pool_kernel<<<pool_grid(num_segments * f), 256, 0, as_stream(stream)>>>(x, ldx, row_index, seg_ptr,
                                                                                   num_segments, f, mode, out, ldo,
                                                                                   argmax);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_segment_pool_bwd_f32(const float* g, int64_t ldg, const int64_t* row_index, const int64_t* key,
                            const int32_t* seg_ptr, int64_t m, int64_t f, int mode, const int32_t* argmax, float* gx,
                            int64_t ldgx, gg_stream_t stream) {
    GG_REQUIRE(m >= 0 && f >= 0, "gg_segment_pool_bwd_f32: negative size");
    GG_REQUIRE(mode == GG_POOL_SUM || mode == GG_POOL_MEAN || mode == GG_POOL_MAX, "gg_segment_pool_bwd_f32: mode=%d", mode);
    if (m == 0 || f == 0) return GG_OK;
    GG_REQUIRE(g && key && seg_ptr && gx && ldg >= f && ldgx >= f && (mode != GG_POOL_MAX || argmax),
               "gg_segment_pool_bwd_f32: bad operands");
    segment_pool_bwd_kernel<<<pool_grid(m * f), 256, 0, as_stream(stream)>>>(g, ldg, row_index, key, seg_ptr, m, f, mode,
                                                                            argmax, gx, ldgx);
    GG_LAUNCHED();
    return GG_OK;
}

}  // extern "C"
