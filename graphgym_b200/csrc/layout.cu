// Graph layout: COO edge_index -> CSR (grouped by target) / CSC (grouped by source), with the
// reference layer's self-loop policy folded in.
//
// The reference never builds a compressed layout: SparseAdj is a COO holder (ref:
// sparse_adj.py:16-56) and every forward re-runs the PyG COO edits (ref: idconv.py:52-60,140-148,
// 232-233,302-304,370).  Here the edit + grouping happens once per edge_index:
//   prep      key[e] = group endpoint (or the sentinel N for a removed self loop), appended loops
//   sort      stable LSD radix sort by key (sort.cu) -> rowid[], perm[]
//   bounds    rowptr[] from the sorted keys (one pass, no atomics)
//   gather    nbr[s] = the other endpoint of edge perm[s]
// All integer, all HBM-bound, all deterministic.
#include "common.cuh"

namespace gg {
int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, void* ws, cudaStream_t st);
size_t scan_workspace_bytes();

static inline bool policy_adds(int policy) {
    return policy == GG_LOOPS_ADD_REMAINING || policy == GG_LOOPS_REMOVE_ADD || policy == GG_LOOPS_ADD;
}
static inline bool policy_removes(int policy) {
    return policy == GG_LOOPS_ADD_REMAINING || policy == GG_LOOPS_REMOVE_ADD ||
           policy == GG_LOOPS_REMOVE;
}
static inline int bits_for(int64_t max_value) {  // bits needed to hold values 0..max_value
    int b = 0;
    while (((int64_t)1 << b) <= max_value) ++b;
    return b;
}

constexpr int kThreads = 256;
static inline int grid_for(int64_t n, int per_thread = 1) {
    int64_t g = ceil_div(n, (int64_t)kThreads * per_thread);
    if (g < 1) g = 1;
    int64_t cap = (int64_t)kNumSMs * 32;
    return (int)(g < cap ? g : cap);
}

__global__ void __launch_bounds__(kThreads)
    layout_prep_kernel(const int64_t* __restrict__ ei, int64_t E, int64_t N, int removes, int adds,
                       int by_source, int64_t range_lo, int64_t range_hi, int64_t nbr_lo, int64_t nbr_hi,
                       uint32_t* __restrict__ keys, int32_t* __restrict__ bad_count) {
    int64_t total = E + (adds ? N : 0);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        uint32_t key;
        if (e < E) {
            int64_t s = ei[e], t = ei[E + e];
            bool bad = s < 0 || s >= N || t < 0 || t >= N;
            if (bad) atomicAdd(bad_count, 1);
            int64_t k = by_source ? s : t;
            int64_t o = by_source ? t : s;  // the other endpoint (what nbr[] will hold)
            // row-partitioned layouts keep only the groups this rank owns, halo-pipelined ones additionally
            // only the neighbours that live in one peer's block
            bool drop = bad || (removes && s == t) || k < range_lo || k >= range_hi || o < nbr_lo || o >= nbr_hi;
            key = drop ? (uint32_t)N : (uint32_t)k;
        } else {
            int64_t i = e - E;
            bool keep = i >= range_lo && i < range_hi && i >= nbr_lo && i < nbr_hi;
            key = keep ? (uint32_t)i : (uint32_t)N;
        }
        keys[e] = key;
    }
}

// rowptr[r] = first slot whose key >= r; the virtual slot M carries key N, so rowptr[N] = E'.
__global__ void __launch_bounds__(kThreads)
    layout_bounds_kernel(const uint32_t* __restrict__ sorted_keys, int64_t M, int64_t N,
                         int32_t* __restrict__ rowptr) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s <= M;
         s += (int64_t)gridDim.x * blockDim.x) {
        int64_t k = s < M ? (int64_t)sorted_keys[s] : N;
        int64_t kp = s > 0 ? (int64_t)sorted_keys[s - 1] : -1;
        for (int64_t r = kp + 1; r <= k; ++r) rowptr[r] = (int32_t)s;
    }
}

__global__ void __launch_bounds__(kThreads)
    layout_gather_kernel(const int64_t* __restrict__ ei, int64_t E, int64_t N, int by_source,
                         const uint32_t* __restrict__ sorted_keys, const int32_t* __restrict__ perm,
                         int64_t M, int32_t* __restrict__ nbr) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < M;
         s += (int64_t)gridDim.x * blockDim.x) {
        if (sorted_keys[s] >= (uint32_t)N) continue;  // removed edges sort past the end
        int64_t e = perm[s];
        nbr[s] = e < E ? (int32_t)(by_source ? ei[E + e] : ei[e]) : (int32_t)(e - E);
    }
}

__global__ void __launch_bounds__(kThreads)
    fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        p[i] = v;
}

__global__ void __launch_bounds__(kThreads)
    slot_inverse_kernel(const int32_t* __restrict__ perm_a, int64_t S, int32_t* __restrict__ inv) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < S;
         s += (int64_t)gridDim.x * blockDim.x)
        inv[perm_a[s]] = (int32_t)s;
}
__global__ void __launch_bounds__(kThreads)
    slot_map_kernel(const int32_t* __restrict__ perm_b, int64_t S, const int32_t* __restrict__ inv,
                    int32_t* __restrict__ map) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < S;
         t += (int64_t)gridDim.x * blockDim.x)
        map[t] = inv[perm_b[t]];
}

// last removed (i,i) edge per node: PyG add_remaining_self_loops keeps that edge's weight
__global__ void __launch_bounds__(kThreads)
    last_loop_kernel(const int64_t* __restrict__ ei, int64_t E, int32_t* __restrict__ last_loop) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E;
         e += (int64_t)gridDim.x * blockDim.x) {
        int64_t s = ei[e];
        if (s == ei[E + e]) atomicMax(&last_loop[s], (int32_t)e);  // max is order-independent
    }
}
__global__ void __launch_bounds__(kThreads)
    slot_weights_kernel(const int32_t* __restrict__ perm, int64_t S, const float* __restrict__ w_edge,
                        int64_t E, const int32_t* __restrict__ last_loop, float loop_fill,
                        float* __restrict__ w_slot) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < S;
         s += (int64_t)gridDim.x * blockDim.x) {
        int64_t e = perm[s];
        float w;
        if (e < E) {
            w = w_edge ? w_edge[e] : 1.0f;
        } else {
            int32_t le = last_loop ? last_loop[e - E] : -1;
            w = le >= 0 ? w_edge[le] : loop_fill;
        }
        w_slot[s] = w;
    }
}

// one warp per segment, lanes stride the slots, fixed-order butterfly: deterministic
__global__ void __launch_bounds__(kThreads)
    segment_degree_kernel(const int32_t* __restrict__ rowptr, const float* __restrict__ w_slot,
                          int64_t N, float* __restrict__ deg) {
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < N; r += nwarps) {
        int beg = rowptr[r], end = rowptr[r + 1];
        float s;
        if (w_slot == nullptr) {
            s = (float)(end - beg);
        } else {
            s = 0.f;
            for (int i = beg + lane; i < end; i += 32) s += w_slot[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        }
        if (lane == 0) deg[r] = s;
    }
}

__device__ __forceinline__ float inv_sqrt_or_zero(float d) {
    // deg.pow(-0.5) with inf -> 0 (ref: idconv.py:57-58,145-146)
    return d > 0.f ? 1.0f / sqrtf(d) : 0.f;
}
__global__ void __launch_bounds__(kThreads)
    gcn_norm_kernel(const int32_t* __restrict__ rowid, const int32_t* __restrict__ nbr,
                    const float* __restrict__ w_in, const float* __restrict__ deg, int64_t S,
                    float* __restrict__ w_out) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < S;
         s += (int64_t)gridDim.x * blockDim.x) {
        float w = w_in ? w_in[s] : 1.0f;
        // same association as the reference: (dis[row] * w) * dis[col]
        w_out[s] = inv_sqrt_or_zero(deg[rowid[s]]) * w * inv_sqrt_or_zero(deg[nbr[s]]);
    }
}

// backward weights of a mean aggregation: slot (source j <- target i) carries 1/indeg(i)
__global__ void __launch_bounds__(kThreads)
    mean_weights_kernel(const int32_t* __restrict__ nbr, const float* __restrict__ deg, int64_t S,
                        float* __restrict__ w_out) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < S;
         s += (int64_t)gridDim.x * blockDim.x) {
        float d = deg[nbr[s]];
        w_out[s] = d > 0.f ? 1.0f / d : 0.f;
    }
}

__global__ void __launch_bounds__(kThreads)
    id_count_kernel(const int64_t* __restrict__ id, int64_t m, int64_t n, float* __restrict__ count) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = id[i];
        if (r >= 0 && r < n) atomicAdd(&count[r], 1.0f);  // small integers: exact in any order
    }
}

}  // namespace gg

using namespace gg;

extern "C" {

int64_t gg_layout_capacity(int64_t E, int64_t N, int policy) {
    return E + (policy_adds(policy) ? N : 0);
}

size_t gg_layout_build_workspace_bytes(int64_t E, int64_t N, int policy) {
    int64_t M = gg_layout_capacity(E, N, policy);
    size_t b = 256;                                        // bad-edge counter
    b += align_up((size_t)(M > 0 ? M : 1) * 4, 256) * 2;   // keys, rowid fallback
    b += gg_sort_pairs_workspace_bytes(M);
    return b + 256;
}

int gg_layout_build(const int64_t* edge_index, int64_t E, int64_t N, int policy, int group_by,
                    int32_t* rowptr, int32_t* nbr, int32_t* perm, int32_t* rowid, void* workspace,
                    size_t workspace_bytes, gg_stream_t stream) {
    return gg_layout_build_range(edge_index, E, N, policy, group_by, 0, N, 0, N, rowptr, nbr, perm, rowid,
                                 workspace, workspace_bytes, stream);
}

int gg_layout_build_range(const int64_t* edge_index, int64_t E, int64_t N, int policy, int group_by,
                          int64_t range_begin, int64_t range_end, int64_t nbr_begin, int64_t nbr_end,
                          int32_t* rowptr, int32_t* nbr, int32_t* perm, int32_t* rowid, void* workspace,
                          size_t workspace_bytes, gg_stream_t stream) {
    GG_REQUIRE(E >= 0 && N >= 0, "gg_layout_build: negative size");
    GG_REQUIRE(nbr_begin >= 0 && nbr_begin <= nbr_end && nbr_end <= N,
               "gg_layout_build_range: bad neighbour range [%lld, %lld) of %lld", (long long)nbr_begin,
               (long long)nbr_end, (long long)N);
    GG_REQUIRE(range_begin >= 0 && range_begin <= range_end && range_end <= N,
               "gg_layout_build_range: bad range [%lld, %lld) of %lld", (long long)range_begin,
               (long long)range_end, (long long)N);
    GG_REQUIRE(policy >= GG_LOOPS_KEEP && policy <= GG_LOOPS_ADD, "gg_layout_build: policy=%d", policy);
    GG_REQUIRE(group_by == GG_BY_TARGET || group_by == GG_BY_SOURCE, "gg_layout_build: group_by=%d",
               group_by);
    int64_t M = gg_layout_capacity(E, N, policy);
    GG_REQUIRE(M < ((int64_t)1 << 31) - 1 && N < ((int64_t)1 << 31) - 1,
               "gg_layout_build: %lld slots / %lld nodes exceed int32 indices", (long long)M,
               (long long)N);
    GG_REQUIRE(rowptr && workspace, "gg_layout_build: null pointer");
    GG_REQUIRE(M == 0 || (nbr && perm), "gg_layout_build: null nbr/perm");
    GG_REQUIRE(E == 0 || edge_index, "gg_layout_build: null edge_index");
    if (workspace_bytes < gg_layout_build_workspace_bytes(E, N, policy)) {
        set_error("gg_layout_build: workspace %zu < %zu", workspace_bytes,
                  gg_layout_build_workspace_bytes(E, N, policy));
        return GG_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    Carver c(workspace);
    int32_t* bad = c.take<int32_t>(64);
    uint32_t* keys = c.take<uint32_t>(M > 0 ? M : 1);
    uint32_t* sorted = rowid ? reinterpret_cast<uint32_t*>(rowid) : c.take<uint32_t>(M > 0 ? M : 1);
    void* sort_ws = c.take<char>(gg_sort_pairs_workspace_bytes(M));

    GG_CUDA(cudaMemsetAsync(bad, 0, sizeof(int32_t), st));
    if (M == 0) {
        fill_i32_kernel<<<grid_for(N + 1), kThreads, 0, st>>>(rowptr, N + 1, 0);
        GG_LAUNCHED();
        return GG_OK;
    }
    layout_prep_kernel<<<grid_for(M, 4), kThreads, 0, st>>>(edge_index, E, N, policy_removes(policy),
                                                           policy_adds(policy),
                                                           group_by == GG_BY_SOURCE, range_begin,
                                                           range_end, nbr_begin, nbr_end, keys, bad);
    GG_LAUNCHED();
    int rc = gg_sort_pairs_u32(keys, nullptr, sorted, reinterpret_cast<uint32_t*>(perm), M,
                               bits_for(N), sort_ws, gg_sort_pairs_workspace_bytes(M), stream);
    if (rc != GG_OK) return rc;
    layout_bounds_kernel<<<grid_for(M + 1, 4), kThreads, 0, st>>>(sorted, M, N, rowptr);
    GG_LAUNCHED();
    layout_gather_kernel<<<grid_for(M, 4), kThreads, 0, st>>>(edge_index, E, N,
                                                             group_by == GG_BY_SOURCE, sorted, perm,
                                                             M, nbr);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_layout_slot_map(const int32_t* perm_a, const int32_t* perm_b, int64_t S, int64_t E, int64_t N,
                       int32_t* scratch, int32_t* map, gg_stream_t stream) {
    GG_REQUIRE(S >= 0 && E >= 0 && N >= 0, "gg_layout_slot_map: negative size");
    if (S == 0) return GG_OK;
    GG_REQUIRE(perm_a && perm_b && scratch && map, "gg_layout_slot_map: null pointer");
    cudaStream_t st = as_stream(stream);
    slot_inverse_kernel<<<grid_for(S, 4), kThreads, 0, st>>>(perm_a, S, scratch);
    GG_LAUNCHED();
    slot_map_kernel<<<grid_for(S, 4), kThreads, 0, st>>>(perm_b, S, scratch, map);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_layout_slot_weights(const int32_t* perm, int64_t S, const int64_t* edge_index,
                           const float* w_edge, int64_t E, int64_t N, int policy, float loop_fill,
                           int32_t* scratch_nodes, float* w_slot, gg_stream_t stream) {
    GG_REQUIRE(S >= 0 && E >= 0 && N >= 0, "gg_layout_slot_weights: negative size");
    if (S == 0) return GG_OK;
    GG_REQUIRE(perm && w_slot, "gg_layout_slot_weights: null pointer");
    cudaStream_t st = as_stream(stream);
    const int32_t* last_loop = nullptr;
    if (policy == GG_LOOPS_ADD_REMAINING && w_edge != nullptr) {
        GG_REQUIRE(edge_index && scratch_nodes,
                   "gg_layout_slot_weights: edge_index/scratch needed for weighted remaining loops");
        fill_i32_kernel<<<grid_for(N), kThreads, 0, st>>>(scratch_nodes, N, -1);
        GG_LAUNCHED();
        if (E > 0) {
            last_loop_kernel<<<grid_for(E, 4), kThreads, 0, st>>>(edge_index, E, scratch_nodes);
            GG_LAUNCHED();
        }
        last_loop = scratch_nodes;
    }
    slot_weights_kernel<<<grid_for(S, 4), kThreads, 0, st>>>(perm, S, w_edge, E, last_loop,
                                                            loop_fill, w_slot);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_segment_degree(const int32_t* rowptr, const float* w_slot, int64_t N, float* deg,
                      gg_stream_t stream) {
    GG_REQUIRE(N >= 0, "gg_segment_degree: negative size");
    if (N == 0) return GG_OK;
    GG_REQUIRE(rowptr && deg, "gg_segment_degree: null pointer");
    segment_degree_kernel<<<grid_for(N * 32), kThreads, 0, as_stream(stream)>>>(rowptr, w_slot, N,
                                                                               deg);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_gcn_norm(const int32_t* rowid, const int32_t* nbr, const float* w_in, const float* deg,
                int64_t S, float* w_out, gg_stream_t stream) {
    GG_REQUIRE(S >= 0, "gg_gcn_norm: negative size");
    if (S == 0) return GG_OK;
    GG_REQUIRE(rowid && nbr && deg && w_out, "gg_gcn_norm: null pointer");
    gcn_norm_kernel<<<grid_for(S, 2), kThreads, 0, as_stream(stream)>>>(rowid, nbr, w_in, deg, S,
                                                                       w_out);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_mean_weights(const int32_t* nbr, const float* deg, int64_t S, float* w_out,
                    gg_stream_t stream) {
    GG_REQUIRE(S >= 0, "gg_mean_weights: negative size");
    if (S == 0) return GG_OK;
    GG_REQUIRE(nbr && deg && w_out, "gg_mean_weights: null pointer");
    mean_weights_kernel<<<grid_for(S, 2), kThreads, 0, as_stream(stream)>>>(nbr, deg, S, w_out);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_id_count(const int64_t* id, int64_t m, int64_t n, float* count, gg_stream_t stream) {
    GG_REQUIRE(m >= 0 && n >= 0, "gg_id_count: negative size");
    if (n == 0) return GG_OK;
    GG_REQUIRE(count, "gg_id_count: null pointer");
    cudaStream_t st = as_stream(stream);
    GG_CUDA(cudaMemsetAsync(count, 0, (size_t)n * sizeof(float), st));
    if (m > 0) {
        GG_REQUIRE(id, "gg_id_count: null id");
        id_count_kernel<<<grid_for(m), kThreads, 0, st>>>(id, m, n, count);
        GG_LAUNCHED();
    }
    return GG_OK;
}

}  // extern "C"
