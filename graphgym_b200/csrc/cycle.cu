// ID-GNN Fast: cycle-count augmentation diag(A^p), p = 1..k.
//
// Reference: graphgym/contrib/transform/identity.py:25-35 densifies the normalised adjacency
// A_hat = D^-1/2 (A+I) D^-1/2 to n x n fp32 and takes torch.diag(A_hat^p) by repeated dense matmul:
// O(k n^3) flops, O(n^2) memory (4 TB at n = 1M).  Here the diagonal is obtained WITHOUT forming any
// power: for a block of B source nodes the one-hot block E_B is propagated through the sparse matrix,
// V_t = A^T V_{t-1} (a [rows, B] SpMM with 512-byte rows — the same warp-per-row, 128-bit-gather
// shape as the aggregation kernel), and
//     symmetric A :  diag(A^2t)_i = <V_t[:,i], V_t[:,i]>,  diag(A^2t+1)_i = <V_t[:,i], V_t+1[:,i]>
//                    (half-power trick: ceil(k/2) hops instead of k)
//     general A   :  diag(A^p)_i = V_p[i, i]  (k hops)
// Two modes:
//     float  A_hat weights per slot, fp32 vectors, fp64 dot accumulation  -> parity with
//            compute_identity within 1e-5 relative;
//     int64  unweighted A (duplicate edges counted), exact closed-walk counts, overflow reported
//            (the reference has no integer mode — SURVEY D2).
// `rows` is a node range [row_begin, row_end) closed under adjacency (one graph, or the graphs of a
// block-diagonal batch that hold the source block), so batches of small graphs cost O(their own size).
// HBM-bound: per hop E_range * 512 B gathered + rows * 512 B written.
#include "common.cuh"

namespace gg {

constexpr int kWalkWarps = 8;
constexpr int kWalkThreads = kWalkWarps * 32;
constexpr int kWalkUnroll = 4;

template <typename T> struct Walk;
template <> struct Walk<float> {
    using Vec = float4;
    using Acc = double;
    static constexpr int kPerLane = 4;
    static constexpr int kCols = 128;
};
template <> struct Walk<long long> {
    using Vec = longlong2;
    using Acc = long long;
    static constexpr int kPerLane = 2;
    static constexpr int kCols = 64;
};

__device__ __forceinline__ void vset(float4& v, int i, float x) { (&v.x)[i] = x; }
__device__ __forceinline__ void vset(longlong2& v, int i, long long x) { (&v.x)[i] = x; }
__device__ __forceinline__ float vget(const float4& v, int i) { return (&v.x)[i]; }
__device__ __forceinline__ long long vget(const longlong2& v, int i) { return (&v.x)[i]; }

template <typename T>
__global__ void __launch_bounds__(kWalkThreads)
    walk_init_kernel(typename Walk<T>::Vec* __restrict__ v, int64_t rows, int64_t src_local, int src_count) {
    using W = Walk<T>;
    const int64_t total = rows * 32;
    for (int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; id < total;
         id += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = id >> 5;
        const int lane = (int)(id & 31);
        typename W::Vec z;
#pragma unroll
        for (int i = 0; i < W::kPerLane; ++i) {
            const int col = lane * W::kPerLane + i;
            vset(z, i, (col < src_count && r == src_local + col) ? (T)1 : (T)0);
        }
        v[id] = z;
    }
}

// Vout[r,:] = sum_{s in row r} w[s] * Vin[nbr[s] - row_begin, :]   (one warp per row of the range)
template <typename T, bool WEIGHTED>
__global__ void __launch_bounds__(kWalkThreads)
    walk_step_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr,
                     const float* __restrict__ w, int64_t row_begin, int64_t rows,
                     const typename Walk<T>::Vec* __restrict__ vin, typename Walk<T>::Vec* __restrict__ vout) {
    using W = Walk<T>;
    using Vec = typename W::Vec;
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * kWalkWarps + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int beg = __ldg(rowptr + row_begin + r), end = __ldg(rowptr + row_begin + r + 1);
    Vec acc;
#pragma unroll
    for (int i = 0; i < W::kPerLane; ++i) vset(acc, i, (T)0);
    for (int base = beg; base < end; base += 32) {
        const int mine = base + lane;
        const int c = mine < end ? __ldg(nbr + mine) - (int)row_begin : 0;
        float wv = 1.f;
        if (WEIGHTED) wv = mine < end ? __ldg(w + mine) : 0.f;
        const int cnt = min(32, end - base);
        for (int k = 0; k < cnt; k += kWalkUnroll) {
            Vec v[kWalkUnroll];
            float ww[kWalkUnroll];
#pragma unroll
            for (int u = 0; u < kWalkUnroll; ++u) {
                const int sl = k + u;
                const int j = __shfl_sync(0xffffffffu, c, sl & 31);
                ww[u] = WEIGHTED ? __shfl_sync(0xffffffffu, wv, sl & 31) : 1.f;
                if (sl < cnt) v[u] = __ldg(vin + (int64_t)j * 32 + lane);
                else {
#pragma unroll
                    for (int i = 0; i < W::kPerLane; ++i) vset(v[u], i, (T)0);
                }
            }
#pragma unroll
            for (int u = 0; u < kWalkUnroll; ++u)
#pragma unroll
                for (int i = 0; i < W::kPerLane; ++i) {
                    if (WEIGHTED) vset(acc, i, (T)fmaf(ww[u], (float)vget(v[u], i), (float)vget(acc, i)));
                    else vset(acc, i, vget(acc, i) + vget(v[u], i));
                }
        }
    }
    vout[r * 32 + lane] = acc;
}

// partial[block][col] = sum over the block's rows of Va[r,col] * Vb[r,col]  (+ fp64 shadow for int64)
template <typename T>
__global__ void __launch_bounds__(kWalkThreads)
    walk_dot_kernel(const typename Walk<T>::Vec* __restrict__ va, const typename Walk<T>::Vec* __restrict__ vb,
                    int64_t rows, int64_t chunk, typename Walk<T>::Acc* __restrict__ partial,
                    double* __restrict__ shadow) {
    using W = Walk<T>;
    using Acc = typename W::Acc;
    __shared__ Acc s_acc[kWalkWarps][W::kCols];
    __shared__ double s_sh[kWalkWarps][W::kCols];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t r_beg = (int64_t)blockIdx.x * chunk;
    const int64_t r_end = r_beg + chunk < rows ? r_beg + chunk : rows;
    Acc acc[W::kPerLane];
    double sh[W::kPerLane];
#pragma unroll
    for (int i = 0; i < W::kPerLane; ++i) acc[i] = (Acc)0, sh[i] = 0.0;
    for (int64_t r = r_beg + wid; r < r_end; r += kWalkWarps) {
        const typename W::Vec a = __ldg(va + r * 32 + lane);
        const typename W::Vec b = __ldg(vb + r * 32 + lane);
#pragma unroll
        for (int i = 0; i < W::kPerLane; ++i) {
            acc[i] += (Acc)vget(a, i) * (Acc)vget(b, i);
            sh[i] += (double)vget(a, i) * (double)vget(b, i);
        }
    }
#pragma unroll
    for (int i = 0; i < W::kPerLane; ++i) {
        s_acc[wid][lane * W::kPerLane + i] = acc[i];
        s_sh[wid][lane * W::kPerLane + i] = sh[i];
    }
    __syncthreads();
    for (int col = threadIdx.x; col < W::kCols; col += blockDim.x) {
        Acc t = (Acc)0;
        double d = 0.0;
        for (int q = 0; q < kWalkWarps; ++q) t += s_acc[q][col], d += s_sh[q][col];
        partial[(int64_t)blockIdx.x * W::kCols + col] = t;
        shadow[(int64_t)blockIdx.x * W::kCols + col] = d;
    }
}

template <typename T, typename Out>
__global__ void walk_dot_final_kernel(const typename Walk<T>::Acc* __restrict__ partial,
                                      const double* __restrict__ shadow, int64_t blocks, int src_count,
                                      Out* __restrict__ out, int64_t ld_out, int p, int* __restrict__ overflow) {
    using W = Walk<T>;
    for (int col = threadIdx.x; col < src_count; col += blockDim.x) {
        typename W::Acc t = 0;
        double d = 0.0;
        for (int64_t b = 0; b < blocks; ++b) t += partial[b * W::kCols + col], d += shadow[b * W::kCols + col];
        if (sizeof(Out) == 8 && !(d < 9.0e18)) {  // int64 mode: the exact value does not fit
            atomicAdd(overflow, 1);
            out[(int64_t)col * ld_out + p] = (Out)0x7fffffffffffffffLL;
        } else {
            out[(int64_t)col * ld_out + p] = (Out)t;
        }
    }
}

template <typename T, typename Out>
__global__ void walk_diag_read_kernel(const typename Walk<T>::Vec* __restrict__ v, int64_t src_local,
                                      int src_count, Out* __restrict__ out, int64_t ld_out, int p) {
    using W = Walk<T>;
    for (int col = threadIdx.x; col < src_count; col += blockDim.x) {
        const typename W::Vec x = v[(src_local + col) * 32 + col / W::kPerLane];
        out[(int64_t)col * ld_out + p] = (Out)vget(x, col % W::kPerLane);
    }
}

static int64_t dot_blocks(int64_t rows) {
    int64_t b = ceil_div(rows, 256);
    if (b > kNumSMs * 4) b = kNumSMs * 4;
    return b < 1 ? 1 : b;
}

template <typename T>
size_t cycle_ws_bytes(int64_t rows) {
    return 2 * align_up((size_t)rows * 512, 256) + 2 * align_up((size_t)dot_blocks(rows) * Walk<T>::kCols * 8, 256) +
           512;
}

// merge-path plan of the whole-graph layout: the float propagation step then runs on gg_spmm_mp_f32
struct WalkPlan {
    const int32_t* item_row;
    const int32_t* item_slot;
    int64_t items;
    // sliced-ELL layout of the same graph (gg_sell_build): when given, the hops run on gg_spmm_sell_f32
    const uint32_t* chunk_ptr = nullptr;
    int64_t chunks = 0;
    const int32_t* idx = nullptr;
    const float* w_sell = nullptr;
    const int32_t* vdst = nullptr;
    const int32_t* hub_rows = nullptr;
    const int32_t* hub_pptr = nullptr;
    int64_t hubs = 0, partial_rows = 0;
};

template <typename T, typename Out>
int cycle_diag(const int32_t* rowptr, const int32_t* nbr, const float* w, int64_t row_begin, int64_t row_end,
               int k, int symmetric, int64_t src_begin, int src_count, Out* out, int64_t ld_out, int* overflow,
               void* workspace, size_t workspace_bytes, cudaStream_t st, const WalkPlan* plan = nullptr) {
    using W = Walk<T>;
    using Vec = typename W::Vec;
    const int64_t rows = row_end - row_begin;
    GG_REQUIRE(rows > 0 && k >= 1 && src_count >= 1 && src_count <= W::kCols,
               "gg_cycle_diag: need 1 <= src_count <= %d, k >= 1, a non-empty row range", W::kCols);
    GG_REQUIRE(src_begin >= row_begin && src_begin + src_count <= row_end,
               "gg_cycle_diag: sources outside the row range");
    GG_REQUIRE(rowptr && out && workspace && ld_out >= k, "gg_cycle_diag: bad operands");
    const size_t mp_bytes = !plan ? 0
                            : plan->chunk_ptr ? align_up(gg_spmm_sell_workspace_bytes(plan->partial_rows, W::kCols), 256)
                                              : align_up(gg_spmm_mp_workspace_bytes(plan->items, W::kCols), 256);
    if (workspace_bytes < cycle_ws_bytes<T>(rows) + mp_bytes) {
        set_error("gg_cycle_diag: workspace %zu < %zu", workspace_bytes, cycle_ws_bytes<T>(rows) + mp_bytes);
        return GG_ERR_WORKSPACE;
    }
    Carver c(workspace);
    Vec* prev = reinterpret_cast<Vec*>(c.take<char>((size_t)rows * 512));
    Vec* cur = reinterpret_cast<Vec*>(c.take<char>((size_t)rows * 512));
    const int64_t nb = dot_blocks(rows);
    typename W::Acc* partial = c.take<typename W::Acc>((size_t)nb * W::kCols);
    double* shadow = c.take<double>((size_t)nb * W::kCols);
    const int64_t chunk = ceil_div(rows, nb);
    const int64_t src_local = src_begin - row_begin;

    int init_grid = (int)(ceil_div(rows * 32, kWalkThreads) < kNumSMs * 16 ? ceil_div(rows * 32, kWalkThreads)
                                                                            : kNumSMs * 16);
    walk_init_kernel<T><<<init_grid, kWalkThreads, 0, st>>>(prev, rows, src_local, src_count);
    GG_LAUNCHED();
    const int step_grid = (int)ceil_div(rows, kWalkWarps);
    void* mp_ws = plan ? c.take<char>(mp_bytes) : nullptr;
    int step_rc = GG_OK;
    auto step = [&](const Vec* in, Vec* o) {
        if (plan && plan->chunk_ptr) {   // degree-sorted sliced-ELL aggregation (round 2: 3.3 vs 4.3 ms at 128 columns)
            const int rc = gg_spmm_sell_f32(plan->chunk_ptr, plan->chunks, plan->idx, plan->w_sell, plan->vdst, rowptr,
                                            plan->hub_rows, plan->hub_pptr, plan->hubs, plan->partial_rows,
                                            reinterpret_cast<const float*>(in), W::kCols, reinterpret_cast<float*>(o),
                                            W::kCols, nullptr, 1, rows, rows, W::kCols, GG_SUM, nullptr, 0, 0.f, nullptr,
                                            nullptr, nullptr, nullptr, nullptr, mp_ws, mp_bytes, 4,
                                            reinterpret_cast<gg_stream_t>(st));
            if (rc != GG_OK) step_rc = rc;
            return;
        }
        if (plan) {  // load-balanced, TMA-staged aggregation kernel (hub rows no longer serialise one warp)
            const int rc = gg_spmm_mpg_f32(rowptr, nbr, w, plan->item_row, plan->item_slot, plan->items,
                                           reinterpret_cast<const float*>(in), W::kCols, reinterpret_cast<float*>(o),
                                           W::kCols, nullptr, 1, rows, rows, W::kCols, GG_SUM, nullptr, 0, 0.f, nullptr,
                                           nullptr, nullptr, nullptr, nullptr, mp_ws, mp_bytes, 4,
                                           reinterpret_cast<gg_stream_t>(st));
            if (rc != GG_OK) step_rc = rc;
            return;
        }
        if (w) walk_step_kernel<T, true><<<step_grid, kWalkThreads, 0, st>>>(rowptr, nbr, w, row_begin, rows, in, o);
        else walk_step_kernel<T, false><<<step_grid, kWalkThreads, 0, st>>>(rowptr, nbr, w, row_begin, rows, in, o);
        count_launch();
    };
    auto dot = [&](const Vec* a, const Vec* b, int p) {
        walk_dot_kernel<T><<<(int)nb, kWalkThreads, 0, st>>>(a, b, rows, chunk, partial, shadow);
        walk_dot_final_kernel<T, Out><<<1, 128, 0, st>>>(partial, shadow, nb, src_count, out, ld_out, p, overflow);
        count_launch(2);
    };
    if (symmetric) {
        int p = 0;  // column index = power - 1
        while (p < k) {
            step(prev, cur);            // cur = V_{t+1}, prev = V_t
            dot(prev, cur, p++);        // power 2t+1
            if (p < k) dot(cur, cur, p++);  // power 2t+2
            Vec* t = prev; prev = cur; cur = t;
        }
    } else {
        for (int p = 0; p < k; ++p) {
            step(prev, cur);
            walk_diag_read_kernel<T, Out><<<1, 128, 0, st>>>(cur, src_local, src_count, out, ld_out, p);
            count_launch();
            Vec* t = prev; prev = cur; cur = t;
        }
    }
    if (step_rc != GG_OK) return step_rc;
    GG_CUDA(cudaPeekAtLastError());
    return GG_OK;
}

}  // namespace gg

using namespace gg;

extern "C" {

size_t gg_cycle_diag_workspace_bytes(int64_t num_rows_in_range) {
    size_t a = cycle_ws_bytes<float>(num_rows_in_range > 0 ? num_rows_in_range : 1);
    size_t b = cycle_ws_bytes<long long>(num_rows_in_range > 0 ? num_rows_in_range : 1);
    return a > b ? a : b;
}

int gg_cycle_diag_f32(const int32_t* rowptr, const int32_t* nbr, const float* w_slot, int64_t row_begin,
                      int64_t row_end, int k, int symmetric, int64_t src_begin, int src_count, float* out,
                      int64_t ld_out, void* workspace, size_t workspace_bytes, gg_stream_t stream) {
    return cycle_diag<float, float>(rowptr, nbr, w_slot, row_begin, row_end, k, symmetric, src_begin, src_count,
                                    out, ld_out, nullptr, workspace, workspace_bytes, as_stream(stream));
}

size_t gg_cycle_diag_mp_workspace_bytes(int64_t num_rows, int64_t items) {
    return gg_cycle_diag_workspace_bytes(num_rows) + align_up(gg_spmm_mp_workspace_bytes(items, 128), 256) + 256;
}

int gg_cycle_diag_mp_f32(const int32_t* rowptr, const int32_t* nbr, const float* w_slot, const int32_t* item_row,
                         const int32_t* item_slot, int64_t items, int64_t num_rows, int k, int symmetric,
                         int64_t src_begin, int src_count, float* out, int64_t ld_out, void* workspace,
                         size_t workspace_bytes, gg_stream_t stream) {
    GG_REQUIRE(item_row && item_slot && items >= 1, "gg_cycle_diag_mp_f32: missing merge-path plan");
    if constexpr (sizeof(float4) != 16) return GG_ERR_UNSUPPORTED;
    WalkPlan plan{item_row, item_slot, items};
    return cycle_diag<float, float>(rowptr, nbr, w_slot, 0, num_rows, k, symmetric, src_begin, src_count, out, ld_out,
                                    nullptr, workspace, workspace_bytes, as_stream(stream), &plan);
}

size_t gg_cycle_diag_sell_workspace_bytes(int64_t num_rows, int64_t partial_rows) {
    return gg_cycle_diag_workspace_bytes(num_rows) + align_up(gg_spmm_sell_workspace_bytes(partial_rows, 128), 256) + 256;
}

int gg_cycle_diag_sell_f32(const int32_t* rowptr, const uint32_t* chunk_ptr, int64_t chunks, const int32_t* idx,
                           const float* w_sell, const int32_t* vdst, const int32_t* hub_rows, const int32_t* hub_pptr,
                           int64_t hubs, int64_t partial_rows, int64_t num_rows, int k, int symmetric, int64_t src_begin,
                           int src_count, float* out, int64_t ld_out, void* workspace, size_t workspace_bytes,
                           gg_stream_t stream) {
    GG_REQUIRE(chunk_ptr && idx && vdst && chunks >= 1, "gg_cycle_diag_sell_f32: missing sliced-ELL layout");
    WalkPlan plan{nullptr, nullptr, 0};
    plan.chunk_ptr = chunk_ptr; plan.chunks = chunks; plan.idx = idx; plan.w_sell = w_sell; plan.vdst = vdst;
    plan.hub_rows = hub_rows; plan.hub_pptr = hub_pptr; plan.hubs = hubs; plan.partial_rows = partial_rows;
    return cycle_diag<float, float>(rowptr, nullptr, w_sell, 0, num_rows, k, symmetric, src_begin, src_count, out, ld_out,
                                    nullptr, workspace, workspace_bytes, as_stream(stream), &plan);
}

int gg_cycle_diag_i64(const int32_t* rowptr, const int32_t* nbr, int64_t row_begin, int64_t row_end, int k,
                      int symmetric, int64_t src_begin, int src_count, int64_t* out, int64_t ld_out,
                      int32_t* overflow_count, void* workspace, size_t workspace_bytes, gg_stream_t stream) {
    GG_REQUIRE(overflow_count, "gg_cycle_diag_i64: null overflow counter");
    return cycle_diag<long long, long long>(rowptr, nbr, nullptr, row_begin, row_end, k, symmetric, src_begin,
                                            src_count, reinterpret_cast<long long*>(out), ld_out, overflow_count,
                                            workspace, workspace_bytes, as_stream(stream));
}

}  // extern "C"
