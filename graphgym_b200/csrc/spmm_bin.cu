// EXPERIMENTAL, opt-in (GG_SPMM_ALGO=bin), NOT YET TIMED — written at the end of round 1 as the candidate for DESIGN.md §8
// "next (1)"; the default paths never call it and its tests run only with GG_TEST_EXPERIMENTAL=1.  State: the parity tests
// against the merge-path kernel passed on a B200 for f = 4..128 (hubs, mean, self term, bias) with the round's last GPU
// seconds; the empty-layout edge case was fixed afterwards and has not been re-run; no timing yet.
//
// Degree-binned aggregation.  The merge-path kernels pay for row ends: 10-26 instructions per slot go into segment
// sweeps, cross-group butterflies and split-row partials, while the narrow SDDMM — same gathers, no row ends — streams
// 1.5x more slots per second.  Here a row end costs nothing:
//   * the rows of a layout are ordered by DESCENDING degree once (radix sort of 0x7fffffff - degree, stable) and the
//     slot arrays are re-laid in that order (gg_degree_keys, gg_sort_pairs_u32, gg_permute_rows_u32);
//   * a warp takes 32/G consecutive rows of that order, one row per group of G lanes; the rows have (nearly) the same
//     degree, so all groups run the same trip count — the degree of the chunk's first row — with no row-end test, no
//     combine across groups and no partial sums; slots past a shorter row's end are masked (weight 0, row 0 gathered);
//   * chunks are handed out by an atomic counter in descending-degree order (longest first);
//   * the few hub rows above a degree threshold form the prefix of the order and stay on the merge-path kernel.
// Every row is summed by one group in slot order: the same fixed order as the one-warp-per-row kernel.
#include "common.cuh"

namespace gg {

constexpr int kBinThreads = 256;

__global__ void __launch_bounds__(256) degree_keys_kernel(const int32_t* __restrict__ rowptr, int64_t n,
                                                          uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        keys[r] = 0x7fffffffu - (uint32_t)(rowptr[r + 1] - rowptr[r]);
        vals[r] = (uint32_t)r;
    }
}

// one warp per permuted row i: dst[rp2[i] + k] = src[rowptr[order[i]] + k]
__global__ void __launch_bounds__(256) permute_rows_kernel(const int32_t* __restrict__ rowptr,
                                                           const int32_t* __restrict__ order,
                                                           const int32_t* __restrict__ rp2,
                                                           const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                                                           int64_t n) {
    const int lane = threadIdx.x & 31;
    for (int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += (int64_t)gridDim.x * 8) {
        const int r = order[i];
        const int s = rowptr[r], d = rowptr[r + 1] - s, t = rp2[i];
        for (int k = lane; k < d; k += 32) dst[t + k] = src[s + k];
    }
}

struct BinArgs {
    const int32_t* rp;       // permuted row pointer [n + 1]
    const int32_t* nbr;      // permuted neighbour ids
    const float* w;          // permuted weights (nullable)
    const int32_t* row_map;  // permuted row -> original row
    int64_t row_begin;       // first permuted row this launch handles (the hub prefix is skipped)
    int64_t n;
    const float* x;
    int64_t ldx;
    float* out;
    int64_t ldo;
    int f;
    int reduce;
    const float* x_self;
    int64_t ld_self;
    float self_scale;
    const float* bias;
    int* counter;
};

template <int G, bool WEIGHTED>
__global__ void __launch_bounds__(kBinThreads, 4) spmm_bin_kernel(const __grid_constant__ BinArgs a) {
    constexpr int S = 32 / G, U = 4;
    const int lane = threadIdx.x & 31;
    const int grp = lane / G, gl = lane % G;
    const int nvec = a.f >> 2;
    const bool act = gl < nvec;
    const char* __restrict__ xg = reinterpret_cast<const char*>(a.x) + (act ? gl : nvec - 1) * 16;
    const uint32_t row_bytes = (uint32_t)a.ldx * 4u;
    const int64_t chunks = (a.n - a.row_begin + S - 1) / S;

    int64_t chunk = 0;
    if (lane == 0) chunk = atomicAdd(a.counter, 1);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    while (chunk < chunks) {
        int next = 0;
        if (lane == 0) next = atomicAdd(a.counter, 1);
        const int64_t i = a.row_begin + chunk * S + grp;
        const bool row_ok = i < a.n;
        const int start = row_ok ? __ldg(a.rp + i) : 0;
        const int deg = row_ok ? __ldg(a.rp + i + 1) - start : 0;
        const int trip = __shfl_sync(0xffffffffu, deg, 0);  // the chunk's first row is its longest
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < trip; k += U) {
            float4 v[U];
            float wv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool ok = k + u < deg;
                const int j = ok ? __ldg(a.nbr + start + k + u) : 0;
                wv[u] = ok ? (WEIGHTED ? __ldg(a.w + start + k + u) : 1.f) : 0.f;
                v[u] = ldg_nc_f4(reinterpret_cast<const float4*>(xg + (uint64_t)(uint32_t)j * row_bytes));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) fma4(acc, wv[u], v[u]);
        }
        if (row_ok && act) {
            const int row = __ldg(a.row_map + i);
            if (a.reduce == GG_MEAN && deg > 0) {
                const float inv = 1.0f / (float)deg;
                acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
            }
            if (a.x_self) fma4(acc, a.self_scale, __ldg(reinterpret_cast<const float4*>(a.x_self + (int64_t)row * a.ld_self) + gl));
            if (a.bias) add4(acc, __ldg(reinterpret_cast<const float4*>(a.bias) + gl));
            reinterpret_cast<float4*>(a.out + (int64_t)row * a.ldo)[gl] = acc;
        }
        chunk = __shfl_sync(0xffffffffu, next, 0);
    }
}

// hub rows: out[row_map[h], :] = tmp[h, :] + self_scale * x_self[row, :] + bias
__global__ void __launch_bounds__(256) finish_rows_kernel(const float* __restrict__ tmp, int64_t ld_tmp,
                                                          const int32_t* __restrict__ row_map, int64_t h, int f,
                                                          const float* __restrict__ x_self, int64_t ld_self,
                                                          float self_scale, const float* __restrict__ bias,
                                                          float* __restrict__ out, int64_t ldo) {
    const int64_t total = h * f;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / f;
        const int c = (int)(e - r * f);
        const int row = row_map[r];
        float v = tmp[r * ld_tmp + c];
        if (x_self) v = fmaf(self_scale, x_self[(int64_t)row * ld_self + c], v);
        if (bias) v += bias[c];
        out[(int64_t)row * ldo + c] = v;
    }
}

static inline int bin_grid(int64_t total, int per_block) {
    int64_t b = ceil_div(total, per_block);
    if (b > (int64_t)kNumSMs * 8) b = (int64_t)kNumSMs * 8;
    return (int)(b < 1 ? 1 : b);
}
static inline bool bin_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int G>
static void launch_bin(const BinArgs& a, cudaStream_t st) {
    const int64_t chunks = ceil_div(a.n - a.row_begin, 32 / G);
    int grid = (int)ceil_div(chunks, kBinThreads / 32);
    if (grid > kNumSMs * 4) grid = kNumSMs * 4;
    if (grid < 1) grid = 1;
    if (a.w) spmm_bin_kernel<G, true><<<grid, kBinThreads, 0, st>>>(a);
    else spmm_bin_kernel<G, false><<<grid, kBinThreads, 0, st>>>(a);
    count_launch();
}

}  // namespace gg

using namespace gg;

extern "C" {

int gg_degree_keys(const int32_t* rowptr, int64_t num_rows, uint32_t* keys, uint32_t* vals, gg_stream_t stream) {
    GG_REQUIRE(num_rows >= 0, "gg_degree_keys: negative size");
    if (num_rows == 0) return GG_OK;
    GG_REQUIRE(rowptr && keys && vals, "gg_degree_keys: null pointer");
    degree_keys_kernel<<<bin_grid(num_rows, 256), 256, 0, as_stream(stream)>>>(rowptr, num_rows, keys, vals);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_permute_rows_u32(const int32_t* rowptr, const int32_t* order, const int32_t* rowptr_perm, const uint32_t* src,
                        uint32_t* dst, int64_t num_rows, gg_stream_t stream) {
    GG_REQUIRE(num_rows >= 0, "gg_permute_rows_u32: negative size");
    if (num_rows == 0) return GG_OK;
    GG_REQUIRE(rowptr && order && rowptr_perm && src && dst, "gg_permute_rows_u32: null pointer");
    permute_rows_kernel<<<bin_grid(num_rows, 8), 256, 0, as_stream(stream)>>>(rowptr, order, rowptr_perm, src, dst,
                                                                             num_rows);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_spmm_bin_f32(const int32_t* rowptr_perm, const int32_t* nbr_perm, const float* w_perm, const int32_t* row_map,
                    int64_t row_begin, int64_t num_rows, const float* x, int64_t ldx, float* out, int64_t ldo, int64_t f,
                    int reduce, const float* x_self, int64_t ld_self, float self_scale, const float* bias,
                    int32_t* counter, gg_stream_t stream) {
    GG_REQUIRE(num_rows >= 0 && row_begin >= 0 && row_begin <= num_rows && f >= 0, "gg_spmm_bin_f32: bad sizes");
    GG_REQUIRE(reduce == GG_SUM || reduce == GG_MEAN, "gg_spmm_bin_f32: reduce=%d", reduce);
    if (num_rows == row_begin || f == 0) return GG_OK;
    if (f % 4 != 0 || f > 128) {
        set_error("gg_spmm_bin_f32: needs f %% 4 == 0 and f <= 128 (got %lld)", (long long)f);
        return GG_ERR_UNSUPPORTED;
    }
    // nbr_perm may be null for a layout without slots (every degree is 0: it is never dereferenced)
    GG_REQUIRE(rowptr_perm && row_map && x && out && counter, "gg_spmm_bin_f32: null pointer");
    GG_REQUIRE(ldx % 4 == 0 && ldx >= f && ldx < ((int64_t)1 << 30) && ldo % 4 == 0 && ldo >= f && bin_al16(x) &&
                   bin_al16(out) && (!bias || bin_al16(bias)) && (!x_self || (bin_al16(x_self) && ld_self % 4 == 0)),
               "gg_spmm_bin_f32: rows must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    GG_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
    BinArgs a{rowptr_perm, nbr_perm, w_perm, row_map, row_begin, num_rows, x, ldx, out, ldo, (int)f, reduce,
              x_self, ld_self, self_scale, bias, counter};
    const int nvec = (int)(f / 4);
    if (nvec <= 4) launch_bin<4>(a, st);
    else if (nvec <= 8) launch_bin<8>(a, st);
    else if (nvec <= 16) launch_bin<16>(a, st);
    else launch_bin<32>(a, st);
    GG_CUDA(cudaPeekAtLastError());
    return GG_OK;
}

int gg_finish_rows_f32(const float* tmp, int64_t ld_tmp, const int32_t* row_map, int64_t num_rows, int64_t f,
                       const float* x_self, int64_t ld_self, float self_scale, const float* bias, float* out,
                       int64_t ldo, gg_stream_t stream) {
    GG_REQUIRE(num_rows >= 0 && f >= 0, "gg_finish_rows_f32: negative size");
    if (num_rows == 0 || f == 0) return GG_OK;
    GG_REQUIRE(tmp && row_map && out && ld_tmp >= f && ldo >= f, "gg_finish_rows_f32: bad operands");
    finish_rows_kernel<<<bin_grid(num_rows * f, 256), 256, 0, as_stream(stream)>>>(tmp, ld_tmp, row_map, num_rows, (int)f,
                                                                                  x_self, ld_self, self_scale, bias, out,
                                                                                  ldo);
    GG_LAUNCHED();
    return GG_OK;
}

}  // extern "C"
