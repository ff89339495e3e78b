// ID-GNN Full: batched k-hop ego-net extraction.
//
// Reference: graphgym/models/transform.py:11-38 — for every centre i of a graph, nx.ego_graph (BFS +
// induced subgraph copy) in Python, then a dict relabel: centre copies keep ids 0..n-1, every other
// member of ego i gets a fresh id from a running counter that starts at n, egos in centre order;
// node_id_index = arange(n).  GraphGym applies it per graph and then collates graphs block-diagonally.
//
// Here a whole block-diagonal batch of graphs is expanded in two launches (sizes -> exclusive scans ->
// fill).  One warp owns one centre: a level-synchronous frontier BFS over bitmaps held in the warp's
// shared-memory slice (visited / frontier / next, one bit per node of the centre's graph), neighbour
// lists read 32 slots at a time.  Members are enumerated in ASCENDING original id (the canonical form,
// SURVEY D8), induced edges in (member ascending, adjacency-slot) order, so the output is a pure
// function of the input — no atomics decide a position.
//
// Output numbering for graph g with n_g nodes and centre range [lo, hi):
//   base_g = out_node_ptr[g];  centre c -> base_g + (c - lo);
//   j-th non-centre member of ego c -> base_g + n_g + (ego_ptr[c] - ego_ptr[lo]) + j.
#include "common.cuh"

namespace gg {
int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, void* ws, cudaStream_t st);
size_t scan_workspace_bytes();

constexpr int kEgoWarps = 4;
constexpr int kEgoThreads = kEgoWarps * 32;

struct EgoArgs {
    const int32_t* rowptr;  // adjacency grouped by source (CSC layout of the symmetric edge list)
    const int32_t* nbr;
    int64_t n;
    int radius;
    const int32_t* graph_ptr;  // [G+1] node ranges of the block-diagonal batch
    const int32_t* graph_of;   // [n] graph id of every node
    int words;                 // bitmap words per warp = ceil(max graph nodes / 32)
    // sizes
    int32_t* node_count;  // [n] non-centre members
    int32_t* edge_count;  // [n] directed induced edges
    // fill
    const uint32_t* ego_ptr;       // [n+1] exclusive scan of node_count
    const uint32_t* edge_ptr;      // [n+1] exclusive scan of edge_count
    const int64_t* out_node_ptr;   // [G+1] first output node id of every graph
    int64_t* orig_id;              // [total nodes] original node of every output node
    int64_t* edge_src;             // [total edges]
    int64_t* edge_tgt;
};

// BFS from centre c inside [lo, hi); leaves the member bitmap in `vis` (local ids = node - lo).
__device__ __forceinline__ void ego_bfs(const EgoArgs& a, int c, int lo, int hi, uint32_t* vis,
                                        uint32_t* cur, uint32_t* nxt, int lane) {
    const int nloc = hi - lo;
    const int words = (nloc + 31) >> 5;
    if (a.radius > 4) {  // reference quirk: radius > 4 => the whole graph (transform.py:18-19)
        for (int w = lane; w < words; w += 32) {
            int rem = nloc - w * 32;
            vis[w] = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
        }
        __syncwarp();
        return;
    }
    for (int w = lane; w < words; w += 32) vis[w] = cur[w] = nxt[w] = 0u;
    __syncwarp();
    if (lane == 0) {
        const int lc = c - lo;
        vis[lc >> 5] = cur[lc >> 5] = 1u << (lc & 31);
    }
    __syncwarp();
    for (int level = 0; level < a.radius; ++level) {
        // every frontier node in turn (warp-uniform), its adjacency 32 slots at a time
        for (int w = 0; w < words; ++w) {
            uint32_t bits = cur[w];
            while (bits) {
                const int u = lo + w * 32 + (__ffs(bits) - 1);
                bits &= bits - 1;
                const int beg = __ldg(a.rowptr + u), end = __ldg(a.rowptr + u + 1);
                for (int s = beg + lane; s < end; s += 32) {
                    const int v = __ldg(a.nbr + s) - lo;
                    const uint32_t m = 1u << (v & 31);
                    if (!(vis[v >> 5] & m)) atomicOr(&nxt[v >> 5], m);  // set-union: order-free
                }
            }
        }
        __syncwarp();
        uint32_t any = 0;
        for (int w = lane; w < words; w += 32) {
            const uint32_t fresh = nxt[w] & ~vis[w];
            vis[w] |= fresh;
            cur[w] = fresh;
            nxt[w] = 0u;
            any |= fresh;
        }
        __syncwarp();
        if (!__any_sync(0xffffffffu, any != 0u)) break;
    }
}

template <bool FILL>
__global__ void __launch_bounds__(kEgoThreads) egonet_kernel(EgoArgs a) {
    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t c64 = (int64_t)blockIdx.x * kEgoWarps + wid;
    if (c64 >= a.n) return;
    const int c = (int)c64;
    uint32_t* vis = smem + (size_t)wid * 4 * a.words;
    uint32_t* cur = vis + a.words;
    uint32_t* nxt = cur + a.words;
    uint32_t* pre = nxt + a.words;  // exclusive popcount prefix of `vis` (FILL only)
    const int g = __ldg(a.graph_of + c);
    const int lo = __ldg(a.graph_ptr + g), hi = __ldg(a.graph_ptr + g + 1);
    const int nloc = hi - lo, words = (nloc + 31) >> 5;
    ego_bfs(a, c, lo, hi, vis, cur, nxt, lane);

    if (!FILL) {
        int members = 0, edges = 0;
        for (int w = lane; w < words; w += 32) members += __popc(vis[w]);
        for (int w = 0; w < words; ++w) {
            uint32_t bits = vis[w];
            while (bits) {
                const int u = lo + w * 32 + (__ffs(bits) - 1);
                bits &= bits - 1;
                const int beg = __ldg(a.rowptr + u), end = __ldg(a.rowptr + u + 1);
                for (int s = beg + lane; s < end; s += 32) {
                    const int v = __ldg(a.nbr + s) - lo;
                    edges += (vis[v >> 5] >> (v & 31)) & 1u;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            members += __shfl_xor_sync(0xffffffffu, members, o);
            edges += __shfl_xor_sync(0xffffffffu, edges, o);
        }
        if (lane == 0) {
            a.node_count[c] = members - 1;
            a.edge_count[c] = edges;
        }
        return;
    }

    // ---- fill ----
    if (lane == 0) {  // exclusive prefix of popcounts: rank of a member = pre[word] + popc(bits below)
        uint32_t run = 0;
        for (int w = 0; w < words; ++w) {
            pre[w] = run;
            run += __popc(vis[w]);
        }
    }
    __syncwarp();
    const int lc = c - lo;
    const int64_t base = a.out_node_ptr[g];
    const int64_t ego_base = base + nloc + ((int64_t)a.ego_ptr[c] - (int64_t)a.ego_ptr[lo]);
    const uint32_t centre_rank = pre[lc >> 5] + __popc(vis[lc >> 5] & ((1u << (lc & 31)) - 1u));
    auto new_id = [&](int v_loc) -> int64_t {
        if (v_loc == lc) return base + lc;
        uint32_t rank = pre[v_loc >> 5] + __popc(vis[v_loc >> 5] & ((1u << (v_loc & 31)) - 1u));
        return ego_base + (int64_t)(rank - (rank > centre_rank ? 1u : 0u));  // skip the centre's slot
    };
    if (lane == 0) a.orig_id[base + lc] = c;
    // non-centre members in ascending original id
    for (int w = lane; w < words; w += 32) {
        uint32_t bits = vis[w];
        while (bits) {
            const int v_loc = w * 32 + (__ffs(bits) - 1);
            bits &= bits - 1;
            if (v_loc != lc) a.orig_id[new_id(v_loc)] = lo + v_loc;
        }
    }
    // induced edges: members ascending, adjacency slots in order, compacted with ballots
    int64_t pos = (int64_t)a.edge_ptr[c];
    // edges of all graphs are concatenated in centre order, so edge_ptr is already global
    for (int w = 0; w < words; ++w) {
        uint32_t bits = vis[w];
        while (bits) {
            const int u_loc = w * 32 + (__ffs(bits) - 1);
            bits &= bits - 1;
            const int u = lo + u_loc;
            const int64_t nu = new_id(u_loc);
            const int beg = __ldg(a.rowptr + u), end = __ldg(a.rowptr + u + 1);
            for (int s0 = beg; s0 < end; s0 += 32) {
                const int s = s0 + lane;
                int v_loc = -1;
                bool keep = false;
                if (s < end) {
                    v_loc = __ldg(a.nbr + s) - lo;
                    keep = (vis[v_loc >> 5] >> (v_loc & 31)) & 1u;
                }
                const uint32_t mask = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const int64_t p = pos + __popc(mask & ((1u << lane) - 1u));
                    a.edge_src[p] = nu;
                    a.edge_tgt[p] = new_id(v_loc);
                }
                pos += __popc(mask);
            }
        }
    }
}

// out_node_ptr[g+1] = n_g + members of the graph's egos; then a serial prefix (G is a batch size)
__global__ void __launch_bounds__(256) ego_graph_sizes_kernel(const int32_t* __restrict__ graph_ptr, int64_t G,
                                                              const uint32_t* __restrict__ ego_ptr,
                                                              int64_t* __restrict__ out_node_ptr) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < G; g += (int64_t)gridDim.x * blockDim.x) {
        const int lo = graph_ptr[g], hi = graph_ptr[g + 1];
        out_node_ptr[g + 1] = (int64_t)(hi - lo) + ((int64_t)ego_ptr[hi] - (int64_t)ego_ptr[lo]);
    }
}
__global__ void ego_graph_prefix_kernel(int64_t G, int64_t* __restrict__ out_node_ptr) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int64_t run = 0;
        out_node_ptr[0] = 0;
        for (int64_t g = 1; g <= G; ++g) {
            run += out_node_ptr[g];
            out_node_ptr[g] = run;
        }
    }
}

}  // namespace gg

using namespace gg;

extern "C" {

size_t gg_egonet_workspace_bytes(int64_t n, int64_t num_graphs) {
    return align_up((size_t)(n + 1) * 4, 256) * 2 + align_up(scan_workspace_bytes(), 256) + 512;
}

static int ego_check(const char* who, int64_t n, int64_t G, int radius, int max_graph_nodes) {
    if (n < 0 || G < 1 || radius < 0 || max_graph_nodes < 1) {
        set_error("%s: bad sizes n=%lld G=%lld radius=%d max_graph_nodes=%d", who, (long long)n, (long long)G,
                  radius, max_graph_nodes);
        return GG_ERR_INVALID;
    }
    // 4 bitmaps per warp, kEgoWarps warps per CTA, <= 200 KB of shared memory
    size_t smem = (size_t)kEgoWarps * 4 * (size_t)((max_graph_nodes + 31) / 32) * 4;
    if (smem > 200 * 1024) {
        set_error("%s: graphs of %d nodes need %zu B of bitmap per CTA (> 200 KB): ego-nets on graphs above "
                  "~100K nodes are out of this kernel's range", who, max_graph_nodes, smem);
        return GG_ERR_UNSUPPORTED;
    }
    return GG_OK;
}

static size_t ego_smem(int max_graph_nodes) {
    return (size_t)kEgoWarps * 4 * (size_t)((max_graph_nodes + 31) / 32) * 4;
}

// Phase 1: node_count[c], edge_count[c]; then ego_ptr / edge_ptr (exclusive scans, n+1 entries) and
// out_node_ptr[G+1].  The caller reads ego_ptr[n], edge_ptr[n], out_node_ptr[G] to size phase 2.
int gg_egonet_sizes(const int32_t* rowptr, const int32_t* nbr, int64_t n, int radius,
                    const int32_t* graph_ptr, const int32_t* graph_of, int64_t num_graphs,
                    int max_graph_nodes, uint32_t* ego_ptr, uint32_t* edge_ptr, int64_t* out_node_ptr,
                    void* workspace, size_t workspace_bytes, gg_stream_t stream) {
    int rc = ego_check("gg_egonet_sizes", n, num_graphs, radius, max_graph_nodes);
    if (rc != GG_OK) return rc;
    GG_REQUIRE(rowptr && graph_ptr && graph_of && ego_ptr && edge_ptr && out_node_ptr && workspace,
               "gg_egonet_sizes: null pointer");
    if (workspace_bytes < gg_egonet_workspace_bytes(n, num_graphs)) {
        set_error("gg_egonet_sizes: workspace too small");
        return GG_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    Carver c(workspace);
    int32_t* node_count = c.take<int32_t>(n + 1);
    int32_t* edge_count = c.take<int32_t>(n + 1);
    void* scan_ws = c.take<char>(scan_workspace_bytes());
    GG_CUDA(cudaMemsetAsync(node_count + n, 0, 4, st));
    GG_CUDA(cudaMemsetAsync(edge_count + n, 0, 4, st));
    EgoArgs a{};
    a.rowptr = rowptr; a.nbr = nbr; a.n = n; a.radius = radius; a.graph_ptr = graph_ptr;
    a.graph_of = graph_of; a.words = (max_graph_nodes + 31) / 32;
    a.node_count = node_count; a.edge_count = edge_count;
    size_t smem = ego_smem(max_graph_nodes);
    if (n > 0) {
        GG_CUDA(cudaFuncSetAttribute(egonet_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
        egonet_kernel<false><<<(int)ceil_div(n, kEgoWarps), kEgoThreads, smem, st>>>(a);
        GG_LAUNCHED();
    }
    rc = exclusive_scan_u32(reinterpret_cast<uint32_t*>(node_count), ego_ptr, n + 1, scan_ws, st);
    if (rc != GG_OK) return rc;
    rc = exclusive_scan_u32(reinterpret_cast<uint32_t*>(edge_count), edge_ptr, n + 1, scan_ws, st);
    if (rc != GG_OK) return rc;
    ego_graph_sizes_kernel<<<(int)(ceil_div(num_graphs, 256) < 1024 ? ceil_div(num_graphs, 256) : 1024), 256, 0,
                             st>>>(graph_ptr, num_graphs, ego_ptr, out_node_ptr);
    GG_LAUNCHED();
    ego_graph_prefix_kernel<<<1, 32, 0, st>>>(num_graphs, out_node_ptr);
    GG_LAUNCHED();
    return GG_OK;
}

// Phase 2: orig_id[total_nodes], edge_index_out[2, total_edges] (row 0 = source, row 1 = target).
int gg_egonet_fill(const int32_t* rowptr, const int32_t* nbr, int64_t n, int radius,
                   const int32_t* graph_ptr, const int32_t* graph_of, int64_t num_graphs,
                   int max_graph_nodes, const uint32_t* ego_ptr, const uint32_t* edge_ptr,
                   const int64_t* out_node_ptr, int64_t total_edges, int64_t* orig_id,
                   int64_t* edge_index_out, gg_stream_t stream) {
    int rc = ego_check("gg_egonet_fill", n, num_graphs, radius, max_graph_nodes);
    if (rc != GG_OK) return rc;
    if (n == 0) return GG_OK;
    GG_REQUIRE(rowptr && graph_ptr && graph_of && ego_ptr && edge_ptr && out_node_ptr && orig_id,
               "gg_egonet_fill: null pointer");
    GG_REQUIRE(total_edges == 0 || edge_index_out, "gg_egonet_fill: null edge output");
    EgoArgs a{};
    a.rowptr = rowptr; a.nbr = nbr; a.n = n; a.radius = radius; a.graph_ptr = graph_ptr;
    a.graph_of = graph_of; a.words = (max_graph_nodes + 31) / 32;
    a.ego_ptr = ego_ptr; a.edge_ptr = edge_ptr; a.out_node_ptr = out_node_ptr; a.orig_id = orig_id;
    a.edge_src = edge_index_out; a.edge_tgt = edge_index_out + total_edges;
    size_t smem = ego_smem(max_graph_nodes);
    cudaStream_t st = as_stream(stream);
    GG_CUDA(cudaFuncSetAttribute(egonet_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    egonet_kernel<true><<<(int)ceil_div(n, kEgoWarps), kEgoThreads, smem, st>>>(a);
    GG_LAUNCHED();
    return GG_OK;
}

}  // extern "C"
