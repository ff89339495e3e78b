// Peer memory over NVLink / NVSwitch for the row-partitioned path: buffers that the other
// GPUs of the box can read and write directly, a stream-ordered barrier between the ranks, and the
// column-slice scatter that turns a rank's rows [rows_r, F] into every peer's feature slice [N, F/P].
//
// One process per GPU: a buffer is cudaMalloc'ed by its owner, exported as a CUDA IPC handle (64 opaque
// bytes the host side passes around with torch.distributed) and opened by the peers; the peers' stores
// then travel over NVLink as ordinary global stores.  Nothing here calls NCCL.
#include <string.h>

#include "common.cuh"

namespace gg {

struct PeerPtrs {
    void* p[GG_PEER_MAX];
};

// ---- barrier --------------------------------------------------------------------------------------
// flags[r] lives on rank r: u32 slots [0, GG_PEER_MAX) are arrival epochs, slot s written by rank s only;
// slot GG_PEER_MAX is rank r's own barrier count.  The kernel takes the next epoch from that counter, so a
// barrier launch carries no host-side state and can be replayed from a CUDA graph.  Thread t signals peer t
// (release, system scope: every store this stream issued before the barrier is visible to whoever acquires
// the flag) and then waits until peer t's signal for the same epoch has landed here.
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(32) peer_barrier_kernel(const __grid_constant__ PeerPtrs flags, int world, int rank,
                                                          long long timeout_cycles) {
    const int t = threadIdx.x;
    uint32_t* mine = static_cast<uint32_t*>(flags.p[rank]);
    uint32_t epoch = 0;
    if (t == 0) {
        epoch = mine[GG_PEER_MAX] + 1u;  // only this kernel touches the counter, launches are stream-ordered
        mine[GG_PEER_MAX] = epoch;
    }
    epoch = __shfl_sync(0xffffffffu, epoch, 0);
    if (t >= world || t == rank) return;
    __threadfence_system();
    st_release_sys(static_cast<uint32_t*>(flags.p[t]) + rank, epoch);
    const long long t0 = clock64();
    // epochs only grow; the signed difference keeps the comparison right across a u32 wrap
    while ((int32_t)(ld_acquire_sys(mine + t) - epoch) < 0) {
        if (clock64() - t0 > timeout_cycles) __trap();  // a lost peer fails the launch instead of hanging the GPU
        __nanosleep(64);
    }
}

// ---- column-slice scatter -------------------------------------------------------------------------
// dst[c][(row_base + i) * fs + k] = src[i * ld + c * fs + k],  c = 0..world-1, fs = f / world.
// One 16-byte vector per lane; a WARP owns a chunk of 32 consecutive vectors of ONE peer's destination (512 contiguous
// bytes = full NVLink packets) and consecutive warps go to different peers, starting from rank + 1: at any moment every
// rank writes to every peer.  History (8 GPUs, 137 MB per rank, profiles/r02 traces): walking the source rows left
// fs*4 = 64-byte stores (0.34 ms + 0.02 ms drain at the barrier); enumerating peer-major made all ranks sweep the same
// peer at the same time (0.23 ms + 0.27 ms drain: the destination's ingress is shared by the 8 senders).
__global__ void __launch_bounds__(256) peer_scatter_cols_kernel(const float* __restrict__ src, int64_t ld,
                                                                int64_t rows, int f, int fs,
                                                                const __grid_constant__ PeerPtrs dst,
                                                                int64_t row_base, int rank) {
    const int svec = fs >> 2, world = f / fs;
    const int64_t per_peer = rows * svec;                       // vectors per peer
    const int64_t chunks_per_peer = (per_peer + 31) >> 5;
    const int64_t total_chunks = chunks_per_peer * world;
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < total_chunks; q += nwarps) {
        const int c = (int)((q % world + rank + 1) % world);
        const int64_t v = (q / world) * 32 + lane;
        if (v >= per_peer) continue;
        const int64_t i = v / svec;
        const int k = (int)(v - i * svec);
        const float4 val = __ldg(reinterpret_cast<const float4*>(src + i * ld + (int64_t)c * fs) + k);
        reinterpret_cast<float4*>(static_cast<float*>(dst.p[c]) + (row_base + i) * fs)[k] = val;
    }
}

// Return leg of the feature-sliced exchange as a bulk push: this rank's finished slice src[n, fs] (all rows, its fs
// columns) goes to the rows' owners; owner o receives rows [o * per, (o + 1) * per) as ONE contiguous block
// recv_o[rank][i][k] (slice-major).  Same striping as the scatter: a warp stores 512 contiguous bytes into one owner,
// consecutive warps serve different owners.  (Storing the rows from the aggregation kernel's epilogue — 64-byte pieces at
// random rows — cost 0.45 ms on top of a 0.61 ms kernel at 8 GPUs.)
__global__ void __launch_bounds__(256) peer_push_rows_kernel(const float* __restrict__ src, int64_t n, int fs, int64_t per,
                                                             int rank, int owner_begin, int owners,
                                                             const __grid_constant__ PeerPtrs dst) {
    const int svec = fs >> 2;
    const int64_t per_vec = per * svec;
    const int64_t chunks_per_owner = (per_vec + 31) >> 5;
    const int64_t total_chunks = chunks_per_owner * owners;
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < total_chunks; q += nwarps) {
        const int o = owner_begin + (int)((q % owners + rank + 1) % owners);
        const int64_t rem = (q / owners) * 32 + lane;            // vector inside the owner's block
        const int64_t e = (int64_t)o * per_vec + rem;            // vector inside src
        if (rem >= per_vec || e >= n * svec) continue;
        const float4 val = __ldg(reinterpret_cast<const float4*>(src) + e);
        reinterpret_cast<float4*>(static_cast<float*>(dst.p[o]) + (int64_t)rank * per * fs)[rem] = val;
    }
}

// owner side: out[i, c * fs + k] = recv[c][i][k]  (local; replaces the copy out of the reused peer block)
__global__ void __launch_bounds__(256) peer_gather_slices_kernel(const float* __restrict__ recv, int64_t per, int64_t rows,
                                                                 int fs, int world, float* __restrict__ out, int64_t ldo) {
    const int svec = fs >> 2, nvec = svec * world;
    const int64_t total = rows * nvec;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / nvec;
        const int v = (int)(e - i * nvec);
        const int c = v / svec, k = v - c * svec;
        const float4 val = __ldg(reinterpret_cast<const float4*>(recv + ((int64_t)c * per + i) * fs) + k);
        reinterpret_cast<float4*>(out + i * ldo)[v] = val;
    }
}

}  // namespace gg

using namespace gg;

extern "C" {

int gg_peer_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int gg_peer_alloc(size_t bytes, void** ptr_host, unsigned char* handle_host) {
    GG_REQUIRE(ptr_host && handle_host && bytes > 0, "gg_peer_alloc: bad arguments");
    void* p = nullptr;
    GG_CUDA(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error("gg_peer_alloc: %s", cudaGetErrorString(e));
        return GG_ERR_CUDA;
    }
    memcpy(handle_host, &h, sizeof(h));
    *ptr_host = p;
    return GG_OK;
}

int gg_peer_open(const unsigned char* handle_host, void** ptr_host) {
    GG_REQUIRE(handle_host && ptr_host, "gg_peer_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_host, sizeof(h));
    GG_CUDA(cudaIpcOpenMemHandle(ptr_host, h, cudaIpcMemLazyEnablePeerAccess));
    return GG_OK;
}

int gg_peer_close(void* ptr) {
    if (ptr) GG_CUDA(cudaIpcCloseMemHandle(ptr));
    return GG_OK;
}

int gg_peer_free(void* ptr) {
    if (ptr) GG_CUDA(cudaFree(ptr));
    return GG_OK;
}

int gg_peer_barrier(void* const* flags_host, int world, int rank, gg_stream_t stream) {
    GG_REQUIRE(flags_host && world >= 1 && world <= GG_PEER_MAX && rank >= 0 && rank < world,
               "gg_peer_barrier: world=%d rank=%d (at most %d peers)", world, rank, GG_PEER_MAX);
    if (world == 1) return GG_OK;
    PeerPtrs f{};
    for (int i = 0; i < world; ++i) {
        GG_REQUIRE(flags_host[i], "gg_peer_barrier: null flag block for rank %d", i);
        f.p[i] = flags_host[i];
    }
    // ~20 s at 2 GHz: far beyond any healthy skew between ranks, short enough to fail before a box-level timeout
    peer_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(f, world, rank, 40000000000LL);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_peer_scatter_cols_f32(const float* src, int64_t ld, int64_t rows, int64_t f, float* const* dst_host,
                             int world, int rank, int64_t row_base, gg_stream_t stream) {
    GG_REQUIRE(rows >= 0 && f >= 0 && row_base >= 0, "gg_peer_scatter_cols_f32: negative size");
    GG_REQUIRE(dst_host && world >= 1 && world <= GG_PEER_MAX && rank >= 0 && rank < world,
               "gg_peer_scatter_cols_f32: world=%d rank=%d", world, rank);
    if (rows == 0 || f == 0) return GG_OK;
    GG_REQUIRE(f % (4 * world) == 0, "gg_peer_scatter_cols_f32: f=%lld must be a multiple of 4*world=%d",
               (long long)f, 4 * world);
    GG_REQUIRE(src && ld >= f && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0,
               "gg_peer_scatter_cols_f32: source rows must be 16-byte aligned");
    PeerPtrs d{};
    for (int i = 0; i < world; ++i) {
        GG_REQUIRE(dst_host[i] && (reinterpret_cast<uintptr_t>(dst_host[i]) & 15) == 0,
                   "gg_peer_scatter_cols_f32: destination %d null or misaligned", i);
        d.p[i] = dst_host[i];
    }
    int64_t total = rows * (f / 4);
    int64_t grid = ceil_div(total, 256 * 4);
    if (grid > (int64_t)kNumSMs * 8) grid = (int64_t)kNumSMs * 8;
    if (grid < 1) grid = 1;
    peer_scatter_cols_kernel<<<(int)grid, 256, 0, as_stream(stream)>>>(src, ld, rows, (int)f, (int)(f / world), d,
                                                                      row_base, rank);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_peer_push_rows_f32(const float* src, int64_t n, int64_t fs, int64_t rows_per_rank, int world, int rank,
                          int owner_begin, int owner_end, float* const* recv_host, gg_stream_t stream) {
    GG_REQUIRE(n >= 0 && fs >= 0 && rows_per_rank >= 1, "gg_peer_push_rows_f32: bad sizes");
    GG_REQUIRE(owner_begin >= 0 && owner_begin <= owner_end && owner_end <= world,
               "gg_peer_push_rows_f32: owners [%d, %d) outside [0, %d)", owner_begin, owner_end, world);
    if (owner_begin == owner_end) return GG_OK;
    GG_REQUIRE(recv_host && world >= 1 && world <= GG_PEER_MAX && rank >= 0 && rank < world &&
                   rows_per_rank * world >= n,
               "gg_peer_push_rows_f32: world=%d rank=%d rows_per_rank=%lld do not cover %lld rows", world, rank,
               (long long)rows_per_rank, (long long)n);
    if (n == 0 || fs == 0) return GG_OK;
    GG_REQUIRE(fs % 4 == 0 && src && (reinterpret_cast<uintptr_t>(src) & 15) == 0,
               "gg_peer_push_rows_f32: the slice must be 16-byte aligned with fs %% 4 == 0");
    PeerPtrs d{};
    for (int i = 0; i < world; ++i) {
        GG_REQUIRE(recv_host[i] && (reinterpret_cast<uintptr_t>(recv_host[i]) & 15) == 0,
                   "gg_peer_push_rows_f32: destination %d null or misaligned", i);
        d.p[i] = recv_host[i];
    }
    int64_t total = rows_per_rank * (owner_end - owner_begin) * (fs / 4);
    int64_t grid = ceil_div(total, 256 * 4);
    if (grid > (int64_t)kNumSMs * 8) grid = (int64_t)kNumSMs * 8;
    if (grid < 1) grid = 1;
    peer_push_rows_kernel<<<(int)grid, 256, 0, as_stream(stream)>>>(src, n, (int)fs, rows_per_rank, rank, owner_begin,
                                                                   owner_end - owner_begin, d);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_peer_gather_slices_f32(const float* recv, int64_t rows_per_rank, int64_t rows, int64_t fs, int world, float* out,
                              int64_t ldo, gg_stream_t stream) {
    GG_REQUIRE(rows >= 0 && rows <= rows_per_rank && fs >= 0 && world >= 1 && world <= GG_PEER_MAX,
               "gg_peer_gather_slices_f32: bad sizes");
    if (rows == 0 || fs == 0) return GG_OK;
    GG_REQUIRE(fs % 4 == 0 && recv && out && ldo >= fs * world && ldo % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(recv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "gg_peer_gather_slices_f32: operands must be 16-byte aligned with fs %% 4 == 0");
    int64_t total = rows * (fs / 4) * world;
    int64_t grid = ceil_div(total, 256 * 4);
    if (grid > (int64_t)kNumSMs * 8) grid = (int64_t)kNumSMs * 8;
    peer_gather_slices_kernel<<<(int)grid, 256, 0, as_stream(stream)>>>(recv, rows_per_rank, rows, (int)fs, world, out, ldo);
    GG_LAUNCHED();
    return GG_OK;
}

}  // extern "C"
