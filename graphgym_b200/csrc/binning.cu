// Label / feature binning on the device: the reference computes the node-task label by binning a
// scalar node statistic over the WHOLE dataset on the host with numpy — "balanced" bin edges are order statistics of the
// sorted values (ref: graphgym/models/feature_augment.py:219-231: np.sort, linspace indices, np.unique) and the label is
// np.digitize(arr, bins) - 1 (ref: feature_augment.py:139-140).  The values are float64 (nx.clustering returns Python
// floats), so ties and edges must be decided in float64 to reproduce the labels bit for bit:
//   gg_f64_sort_keys   order-preserving 64-bit keys of the doubles, split into two u32 words (+ the identity permutation)
//                      -> two stable LSD passes of gg_sort_pairs_u32 (low word, then high word via gg_gather_u32) give
//                      the ascending order; the caller reads the handful of order statistics it needs
//   gg_digitize_f64    out[i] = #{bins[j] <= x[i]} - 1   (bins ascending; np.digitize(x, bins) - 1)
#include "common.cuh"

namespace gg {

__global__ void __launch_bounds__(256) f64_sort_keys_kernel(const double* __restrict__ x, int64_t n,
                                                            uint32_t* __restrict__ hi, uint32_t* __restrict__ lo,
                                                            uint32_t* __restrict__ idx) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long b = (unsigned long long)__double_as_longlong(x[i]);
        b = (b >> 63) ? ~b : (b | 0x8000000000000000ull);   // negative: flip all; positive: set the sign bit
        hi[i] = (uint32_t)(b >> 32);
        lo[i] = (uint32_t)b;
        idx[i] = (uint32_t)i;
    }
}

__global__ void __launch_bounds__(256) gather_u32_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx,
                                                         int64_t n, uint32_t* __restrict__ dst) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[idx[i]];
}

__global__ void __launch_bounds__(256) digitize_f64_kernel(const double* __restrict__ x, int64_t n,
                                                           const double* __restrict__ bins, int m,
                                                           int64_t* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = x[i];
        int lo = 0, hi = m;   // first j with bins[j] > v
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (bins[mid] <= v) lo = mid + 1;
            else hi = mid;
        }
        out[i] = (int64_t)lo - 1;
    }
}

static inline int bin_grid(int64_t n) {
    int64_t b = ceil_div(n, 256);
    if (b > (int64_t)kNumSMs * 8) b = (int64_t)kNumSMs * 8;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace gg

using namespace gg;

extern "C" {

int gg_f64_sort_keys(const double* x, int64_t n, uint32_t* key_hi, uint32_t* key_lo, uint32_t* index, gg_stream_t stream) {
    GG_REQUIRE(n >= 0 && n < ((int64_t)1 << 31), "gg_f64_sort_keys: n=%lld out of range", (long long)n);
    if (n == 0) return GG_OK;
    GG_REQUIRE(x && key_hi && key_lo && index, "gg_f64_sort_keys: null pointer");
    f64_sort_keys_kernel<<<bin_grid(n), 256, 0, as_stream(stream)>>>(x, n, key_hi, key_lo, index);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_gather_u32(const uint32_t* src, const uint32_t* index, int64_t n, uint32_t* dst, gg_stream_t stream) {
    GG_REQUIRE(n >= 0, "gg_gather_u32: negative size");
    if (n == 0) return GG_OK;
    GG_REQUIRE(src && index && dst, "gg_gather_u32: null pointer");
    gather_u32_kernel<<<bin_grid(n), 256, 0, as_stream(stream)>>>(src, index, n, dst);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_digitize_f64(const double* x, int64_t n, const double* bins, int num_bins, int64_t* out, gg_stream_t stream) {
    GG_REQUIRE(n >= 0 && num_bins >= 1, "gg_digitize_f64: bad sizes");
    if (n == 0) return GG_OK;
    GG_REQUIRE(x && bins && out, "gg_digitize_f64: null pointer");
    digitize_f64_kernel<<<bin_grid(n), 256, 0, as_stream(stream)>>>(x, n, bins, num_bins, out);
    GG_LAUNCHED();
    return GG_OK;
}

}  // extern "C"
