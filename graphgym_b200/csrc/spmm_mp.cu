// Merge-path CSR aggregation (SpMM v2) — the load-balanced, persistent version of spmm.cu.
//
// Why: on power-law graphs (hub rows of 10^4 slots next to rows of 3) one-warp-per-row leaves the SMs
// 18 % idle and the resident warps at 24 % of peak (profiles/r01_spmm_v1_ncu_raw.csv): the kernel is
// latency/imbalance-bound at 57 % of the HBM roofline, DRAM only 35 % busy.
//
// How: the rows' slots and one end-of-row marker per row form one merged sequence of E' + N units
// (Merrill & Garland merge-path).  The plan (built once per layout, gg_spmm_plan_build) cuts it into
// items of `units` consecutive units; item k starts at (item_row[k], item_slot[k]).  A persistent grid
// of warps pulls items from an atomic counter.  Per item a warp
//   1. stages the item's neighbour indices, weights and row pointers in its private shared-memory
//      tile — the 16-byte-aligned interior with ONE cp.async.bulk (TMA 1-D bulk copy) per array,
//      completion on an mbarrier;
//   2. streams the slots: batches of independent 128-bit feature-row gathers (indices come from
//      shared memory, so the gather addresses never wait on a global load), accumulating in slot order;
//   3. at every marker writes the finished row (epilogue fused), except for the two rows an item can
//      share with its neighbours: their partial sums go to head[k] / carry[k].
// A second tiny kernel adds the partials of split rows in item order.  No atomics touch the data, the
// sum order per row is fixed => bitwise run-to-run deterministic, like v1.
#include "common.cuh"

namespace gg {

constexpr int kMpWarps = 8;
constexpr int kMpThreads = kMpWarps * 32;
constexpr int kMpMaxUnits = 480;           // units per item (slots + markers)
constexpr int kMpTile = kMpMaxUnits + 8;   // + alignment slack

struct MpArgs {
    const int32_t* rowptr;
    const int32_t* nbr;
    const float* w;
    const int32_t* item_row;   // [items + 1]
    const int32_t* item_slot;  // [items + 1]
    int items;
    const float* x;
    int64_t ldx;
    float* out;
    int64_t ldo;
    int64_t n;
    int f;
    int reduce;
    const float* x_self;
    int64_t ld_self;
    float self_scale;
    const float* bias;
    int* counter;
    float* carry;  // [items, f]
    float* head;   // [items, f]
    // optional rank-1 epilogue terms: out[row,:] += r1_s[row] * r1_v[:] + r2_s[row] * r2_v[:]
    // (GAT backward: da_src[j] * att_src + da_tgt[j] * att_tgt, ref: idconv.py:324)
    const float* r1_s;
    const float* r1_v;
    const float* r2_s;
    const float* r2_v;
    int l2_hint;  // 1: gathers evict_last, index / weight / output streams evict_first
};

// ---- mbarrier / bulk-copy PTX -----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (int spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (spin > (1 << 22)) __trap();  // a lost copy must fail the launch, never hang the GPU
    }
}

// Stage global[s0, s1) into tile so that tile[s - sa] holds element s, sa = s0 & ~3.
template <typename T, bool TMA>
__device__ __forceinline__ uint32_t stage_begin(T* tile, const T* __restrict__ g, int s0, int s1, int lane,
                                                uint64_t* bar, uint64_t pol) {
    const int sa = s0 & ~3;
    const int ea = s1 & ~3;  // interior [sa, ea) is 16-byte aligned on both sides
    uint32_t bytes = 0;
    if (TMA) {
        if (ea > sa) {
            bytes = (uint32_t)(ea - sa) * 4u;
            if (lane == 0) bulk_g2s(tile, g + sa, bytes, bar, pol);
        }
        const int t = ea + lane;  // <= 3 tail elements past the aligned interior (ea >= sa always)
        if (lane < 4 && t < s1) tile[t - sa] = g[t];
    } else {
        for (int q = s0 + lane; q < s1; q += 32) tile[q - sa] = g[q];
    }
    return bytes;
}

// U = independent feature-row gathers per lane and batch.  DEEP doubles it for the 128-wide case: 16 x 512 B
// in flight per warp at 3 CTAs/SM (24 warps) instead of 8 x 512 B at 4 CTAs/SM (32 warps).
template <int VPL, bool DEEP> struct MpUnroll {
    static constexpr int value = VPL == 1 ? (DEEP ? 16 : 8) : (VPL == 2 ? (DEEP ? 8 : 4) : 2);
};

template <int VPL, bool WEIGHTED, bool TMA, bool DEEP, bool RANK1>
__global__ void __launch_bounds__(kMpThreads, VPL == 1 ? (DEEP ? 3 : 4) : 2) spmm_mp_kernel(MpArgs a) {
    constexpr int U = MpUnroll<VPL, DEEP>::value;
    __shared__ __align__(16) int32_t s_nbr_all[kMpWarps][kMpTile];
    __shared__ __align__(16) float s_w_all[WEIGHTED ? kMpWarps : 1][WEIGHTED ? kMpTile : 4];
    __shared__ __align__(16) int32_t s_rp_all[kMpWarps][kMpTile];
    __shared__ __align__(8) uint64_t s_bar[kMpWarps];

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int32_t* s_nbr = s_nbr_all[wid];
    float* s_w = s_w_all[WEIGHTED ? wid : 0];
    int32_t* s_rp = s_rp_all[wid];
    uint64_t* bar = &s_bar[wid];
    if (TMA) {
        if (lane == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncwarp();
    }
    uint32_t phase = 0;
    const uint64_t pol_keep = l2_policy(a.l2_hint ? 1 : 0), pol_stream = l2_policy(a.l2_hint ? 2 : 0);

    const int nvec = a.f >> 2;
    const float4* __restrict__ x4 = reinterpret_cast<const float4*>(a.x);
    const int64_t ldx4 = a.ldx >> 2;
    bool act[VPL];
#pragma unroll
    for (int q = 0; q < VPL; ++q) act[q] = lane + q * 32 < nvec;

    int item = 0;
    if (lane == 0) item = atomicAdd(a.counter, 1);
    item = __shfl_sync(0xffffffffu, item, 0);

    while (item < a.items) {
        int next = 0;
        if (lane == 0) next = atomicAdd(a.counter, 1);  // in flight while this item is processed
        const int r0 = __ldg(a.item_row + item), s0 = __ldg(a.item_slot + item);
        const int r1 = __ldg(a.item_row + item + 1), s1 = __ldg(a.item_slot + item + 1);
        const int sa = s0 & ~3;

        // ---- stage: neighbour ids, weights (TMA bulk copies) and row pointers r0 .. r1+1 ----
        __syncwarp();  // everyone is done with the previous item's tiles
        uint32_t bytes = 0;
        if (TMA && lane == 0) {
            const int ea = s1 & ~3;
            if (ea > sa) mbar_expect_tx(bar, (uint32_t)(ea - sa) * 4u * (WEIGHTED ? 2u : 1u));
        }
        bytes += stage_begin<int32_t, TMA>(s_nbr, a.nbr, s0, s1, lane, bar, pol_stream);
        if (WEIGHTED) bytes += stage_begin<float, TMA>(s_w, a.w, s0, s1, lane, bar, pol_stream);
        const int nrp = (r1 < a.n ? r1 + 1 : (int)a.n) - r0 + 1;  // rowptr[r0 .. min(r1+1, n)]
        for (int i = lane; i < nrp; i += 32) s_rp[i] = __ldg(a.rowptr + r0 + i);
        __syncwarp();
        if (TMA && bytes) {
            mbar_wait(bar, phase);
            phase ^= 1;
        }

        // ---- stream the item's slots ----
        const bool cont_first = s0 > s_rp[0];
        int cur = r0;
        int re = r0 < a.n ? s_rp[1] : 0x7fffffff;  // slot index at which the current row ends
        float4 acc[VPL];
#pragma unroll
        for (int q = 0; q < VPL; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);

        auto finalize = [&](int row) {
            if (row == r0 && cont_first) {  // the row began in an earlier item: partial only
#pragma unroll
                for (int q = 0; q < VPL; ++q)
                    if (act[q]) reinterpret_cast<float4*>(a.head + (int64_t)item * a.f)[lane + q * 32] = acc[q];
            } else {
                const int deg = s_rp[row - r0 + 1] - s_rp[row - r0];
                const float inv = (a.reduce == GG_MEAN && deg > 0) ? 1.0f / (float)deg : 1.0f;
#pragma unroll
                for (int q = 0; q < VPL; ++q) {
                    if (!act[q]) continue;
                    const int vi = lane + q * 32;
                    float4 r = acc[q];
                    if (a.reduce == GG_MEAN) { r.x *= inv; r.y *= inv; r.z *= inv; r.w *= inv; }
                    if (a.x_self)
                        fma4(r, a.self_scale,
                             __ldg(reinterpret_cast<const float4*>(a.x_self + (int64_t)row * a.ld_self) + vi));
                    if (a.bias) add4(r, __ldg(reinterpret_cast<const float4*>(a.bias) + vi));
                    if (RANK1) {
                        if (a.r1_s) fma4(r, __ldg(a.r1_s + row), __ldg(reinterpret_cast<const float4*>(a.r1_v) + vi));
                        if (a.r2_s) fma4(r, __ldg(a.r2_s + row), __ldg(reinterpret_cast<const float4*>(a.r2_v) + vi));
                    }
                    stg_f4_hint(reinterpret_cast<float4*>(a.out + (int64_t)row * a.ldo) + vi, r, pol_stream);
                }
            }
#pragma unroll
            for (int q = 0; q < VPL; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        };

        for (int s = s0; s < s1; s += U) {
            float4 v[U][VPL];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool ok = s + u < s1;
                const int j = ok ? s_nbr[s + u - sa] : 0;
                const float4* p = x4 + (int64_t)j * ldx4 + lane;
#pragma unroll
                for (int q = 0; q < VPL; ++q)
                    v[u][q] = (ok && act[q]) ? ldg_nc_f4_hint(p + q * 32, pol_keep) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (s + u < s1) {  // warp-uniform
                    while (s + u == re) {  // markers in front of this slot (also empty rows)
                        finalize(cur);
                        ++cur;
                        re = s_rp[cur - r0 + 1];
                    }
                    const float wv = WEIGHTED ? s_w[s + u - sa] : 1.f;
#pragma unroll
                    for (int q = 0; q < VPL; ++q) {
                        if (WEIGHTED) fma4(acc[q], wv, v[u][q]);
                        else add4(acc[q], v[u][q]);
                    }
                }
            }
        }
        while (cur < r1) {  // markers after the last slot of the item
            finalize(cur);
            ++cur;
        }
        // the row still open at the end of the item (row r1): its partial, possibly all zero
#pragma unroll
        for (int q = 0; q < VPL; ++q)
            if (act[q]) reinterpret_cast<float4*>(a.carry + (int64_t)item * a.f)[lane + q * 32] = acc[q];

        item = __shfl_sync(0xffffffffu, next, 0);
    }
}

// Rows split over several items: out[row] = epi( carry[k0] + ... + carry[f-1] + head[f] ), f = the item
// that consumed the row's marker.  One warp per item f; almost all exit at once.
template <int VPL>
__global__ void __launch_bounds__(kMpThreads) spmm_mp_fixup_kernel(MpArgs a) {
    const int lane = threadIdx.x & 31;
    const int f_item = blockIdx.x * kMpWarps + (threadIdx.x >> 5);
    if (f_item < 1 || f_item >= a.items) return;
    const int r0 = __ldg(a.item_row + f_item), s0 = __ldg(a.item_slot + f_item);
    const int r1 = __ldg(a.item_row + f_item + 1);
    if (r0 >= a.n || r0 >= r1) return;  // the item does not finish its first row
    const int rb = __ldg(a.rowptr + r0);
    if (s0 <= rb) return;                // the row starts with the item: written directly
    int k0 = f_item - 1;
    while (k0 >= 1 && __ldg(a.item_row + k0) == r0 && __ldg(a.item_slot + k0) > rb) --k0;
    const int nvec = a.f >> 2;
    const int deg = __ldg(a.rowptr + r0 + 1) - rb;
    const float inv = (a.reduce == GG_MEAN && deg > 0) ? 1.0f / (float)deg : 1.0f;
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
        const int vi = lane + q * 32;
        if (vi >= nvec) continue;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = k0; k < f_item; ++k) add4(r, reinterpret_cast<const float4*>(a.carry + (int64_t)k * a.f)[vi]);
        add4(r, reinterpret_cast<const float4*>(a.head + (int64_t)f_item * a.f)[vi]);
        if (a.reduce == GG_MEAN) { r.x *= inv; r.y *= inv; r.z *= inv; r.w *= inv; }
        if (a.x_self)
            fma4(r, a.self_scale, __ldg(reinterpret_cast<const float4*>(a.x_self + (int64_t)r0 * a.ld_self) + vi));
        if (a.bias) add4(r, __ldg(reinterpret_cast<const float4*>(a.bias) + vi));
        if (a.r1_s) fma4(r, __ldg(a.r1_s + r0), __ldg(reinterpret_cast<const float4*>(a.r1_v) + vi));
        if (a.r2_s) fma4(r, __ldg(a.r2_s + r0), __ldg(reinterpret_cast<const float4*>(a.r2_v) + vi));
        reinterpret_cast<float4*>(a.out + (int64_t)r0 * a.ldo)[vi] = r;
    }
}

// item k starts at diagonal d = k * units of the merged (slots + markers) sequence:
//   row  = number of markers before d = first r with rowptr[r+1] + r >= d,   slot = d - row
__global__ void __launch_bounds__(256) spmm_plan_kernel(const int32_t* __restrict__ rowptr, int64_t n,
                                                        int64_t slots, int units, int items,
                                                        int32_t* __restrict__ item_row,
                                                        int32_t* __restrict__ item_slot) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k <= items; k += gridDim.x * blockDim.x) {
        int64_t d = (int64_t)k * units;
        if (d > n + slots) d = n + slots;
        int64_t lo = 0, hi = n;  // answer in [0, n]
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if ((int64_t)rowptr[mid + 1] + mid >= d) hi = mid;
            else lo = mid + 1;
        }
        item_row[k] = (int32_t)lo;
        item_slot[k] = (int32_t)(d - lo);
    }
}

// =====================================================================================================
// Narrow rows (f <= 128) and peer-memory output.
//
// With f/4 < 32 vectors per row the kernel above would leave most lanes idle.  Here the 32 lanes are cut
// into S = 32/G groups of G lanes (G = 4, 8, 16 or 32 >= f/4) and one warp instruction gathers S DIFFERENT
// slots of the warp's item: a batch is S x 8 consecutive slots, group g owns slots [8g, 8g+8) of it (two
// 128-bit shared loads fetch its 8 neighbour ids), every group accumulates privately, and when a row ends
// the partial sums are combined with a fixed xor-butterfly across the groups.  Everything is warp-uniform:
// a row end costs one predicated sweep + log2(S) shuffle stages, never a divergent branch.
// (The first design made every group an independent item consumer; the union of 8 groups' row-end
// branches serialised the warp at 65 instructions per slot and 33 G slots/s whatever the row width —
// profiles/r01_mpg16_v0_ncu_raw.csv.  A third design — private walks per group, one segmented scan of the
// groups' tails per batch, row ends finished by the owning group — was correct but slower, 1.73 vs 1.46 ms at
// F/P = 16: the divergent out-of-line row stores and spills cost more than the sweeps they replaced.)
//
// PEER: the row-partitioned path's feature-sliced exchange (parallel.py).  This rank aggregates its
// F/P-wide column slice for ALL rows; a finished row is stored straight into the memory of the rank that
// owns the row (peer pointer table, plain 16-byte stores that travel over NVLink), so the aggregation
// and the return leg of the exchange are one kernel.
// =====================================================================================================
// slots per group and batch: 4 (measured on the products graph, all rows: F=16 1.41 vs 1.51 ms with 8, F=32 1.44 vs 1.50,
// F=64 2.40 vs 2.53, F=128 4.17 vs 4.46 — shorter row-end sweeps and 32 resident warps outweigh the shallower batches;
// at F=128 this kernel also beats the whole-warp TMA kernel above, 4.17 vs 4.55 ms)
#ifndef GG_MPG_U
#define GG_MPG_U 4
#endif
struct PeerOut {
    float* out[GG_PEER_MAX];  // out[o] = rank o's [rows_per_rank, ldo] block, already offset to this rank's columns
    int per;                  // rows per rank
};

template <bool PEER>
__device__ __forceinline__ float4* mpg_out_row(const MpArgs& a, const PeerOut& po, int row) {
    if (PEER) {
        const int o = row / po.per;
        return reinterpret_cast<float4*>(po.out[o] + (int64_t)(row - o * po.per) * a.ldo);
    }
    return reinterpret_cast<float4*>(a.out + (int64_t)row * a.ldo);
}

template <int G>
__device__ __forceinline__ void mpg_reduce_groups(float4& r) {
#pragma unroll
    for (int m = G; m < 32; m <<= 1) {
        r.x += __shfl_xor_sync(0xffffffffu, r.x, m);
        r.y += __shfl_xor_sync(0xffffffffu, r.y, m);
        r.z += __shfl_xor_sync(0xffffffffu, r.z, m);
        r.w += __shfl_xor_sync(0xffffffffu, r.w, m);
    }
}

template <int G, bool WEIGHTED, bool PEER>
__global__ void __launch_bounds__(kMpThreads, 4)
    spmm_mpg_kernel(const __grid_constant__ MpArgs a, const __grid_constant__ PeerOut po) {
    constexpr int S = 32 / G;  // slots per warp instruction
    constexpr int U = GG_MPG_U;  // slots per group and batch
    constexpr int B = S * U;   // slots per batch
    __shared__ __align__(16) int32_t s_nbr_all[kMpWarps][kMpTile];
    __shared__ __align__(16) float s_w_all[WEIGHTED ? kMpWarps : 1][WEIGHTED ? kMpTile : 4];
    __shared__ __align__(16) int32_t s_rp_all[kMpWarps][kMpTile];

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int grp = lane / G, gl = lane % G;
    int32_t* s_nbr = s_nbr_all[wid];
    float* s_w = s_w_all[WEIGHTED ? wid : 0];
    int32_t* s_rp = s_rp_all[wid];

    const int nvec = a.f >> 2;
    const bool act = gl < nvec;
    const bool writer = act && grp == 0;
    // gather address = xg + j * row_bytes: one IMAD.WIDE.U32 per slot (the host checks ldx * 4 < 2^32);
    // lanes past the row's last vector (f/4 < G) gather a duplicate of it and never store: no predicates
    const char* __restrict__ xg = reinterpret_cast<const char*>(a.x) + (act ? gl : nvec - 1) * 16;
    const uint32_t row_bytes = (uint32_t)a.ldx * 4u;
    const uint64_t pol_keep = l2_policy(a.l2_hint ? 1 : 0), pol_stream = l2_policy(a.l2_hint ? 2 : 0);
    auto gather = [&](int j) {
        return ldg_nc_f4_hint(reinterpret_cast<const float4*>(xg + (uint64_t)(uint32_t)j * row_bytes), pol_keep);
    };

    int item = 0;
    if (lane == 0) item = atomicAdd(a.counter, 1);
    item = __shfl_sync(0xffffffffu, item, 0);

    while (item < a.items) {
        int next = 0;
        if (lane == 0) next = atomicAdd(a.counter, 1);
        const int r0 = __ldg(a.item_row + item), s0 = __ldg(a.item_slot + item);
        const int r1 = __ldg(a.item_row + item + 1), s1 = __ldg(a.item_slot + item + 1);

        __syncwarp();  // everyone is done with the previous item's tiles
        for (int q = s0 + lane; q < s1; q += 32) {
            s_nbr[q - s0] = ldg_nc_s32_hint(a.nbr + q, pol_stream);
            if (WEIGHTED) s_w[q - s0] = ldg_nc_f32_hint(a.w + q, pol_stream);
        }
        const int nrp = (r1 < a.n ? r1 + 1 : (int)a.n) - r0 + 1;
        for (int i = lane; i < nrp; i += 32) s_rp[i] = __ldg(a.rowptr + r0 + i);
        __syncwarp();

        const bool cont_first = s0 > s_rp[0];
        int cur = r0;
        int re = r0 < a.n ? s_rp[1] : 0x7fffffff;  // slot index at which the current row ends
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);

        // acc holds this group's share of row `row`: combine the groups, apply the epilogue, store
        auto finalize = [&](int row) {
            mpg_reduce_groups<G>(acc);
            if (row == r0 && cont_first) {  // the row began in an earlier item: partial only
                if (writer) reinterpret_cast<float4*>(a.head + (int64_t)item * a.f)[gl] = acc;
            } else if (writer) {
                float4 r = acc;
                if (a.reduce == GG_MEAN) {
                    const int deg = s_rp[row - r0 + 1] - s_rp[row - r0];
                    const float inv = deg > 0 ? 1.0f / (float)deg : 1.0f;
                    r.x *= inv; r.y *= inv; r.z *= inv; r.w *= inv;
                }
                if (a.x_self)
                    fma4(r, a.self_scale, __ldg(reinterpret_cast<const float4*>(a.x_self + (int64_t)row * a.ld_self) + gl));
                if (a.bias) add4(r, __ldg(reinterpret_cast<const float4*>(a.bias) + gl));
                if (a.r1_s) fma4(r, __ldg(a.r1_s + row), __ldg(reinterpret_cast<const float4*>(a.r1_v) + gl));
                if (a.r2_s) fma4(r, __ldg(a.r2_s + row), __ldg(reinterpret_cast<const float4*>(a.r2_v) + gl));
                stg_f4_hint(mpg_out_row<PEER>(a, po, row) + gl, r, pol_stream);
            }
            acc = make_float4(0.f, 0.f, 0.f, 0.f);
        };

        for (int s = s0; s < s1; s += B) {
            const int g0 = s + grp * U;  // this group's first slot of the batch
            const int t = g0 - s0;       // multiple of U (4 or 8): 16-byte aligned inside the tile
            const int e = s + B < s1 ? s + B : s1;
            const bool full = s + B <= s1;
            float4 v[U];
            float wv[U];
            if (full) {
                int j[U];
#pragma unroll
                for (int q4 = 0; q4 < U / 4; ++q4) {
                    const int4 i4 = *reinterpret_cast<const int4*>(s_nbr + t + 4 * q4);
                    j[4 * q4] = i4.x; j[4 * q4 + 1] = i4.y; j[4 * q4 + 2] = i4.z; j[4 * q4 + 3] = i4.w;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) v[u] = gather(j[u]);
                if (WEIGHTED) {
#pragma unroll
                    for (int q4 = 0; q4 < U / 4; ++q4) {
                        const float4 w4 = *reinterpret_cast<const float4*>(s_w + t + 4 * q4);
                        wv[4 * q4] = w4.x; wv[4 * q4 + 1] = w4.y; wv[4 * q4 + 2] = w4.z; wv[4 * q4 + 3] = w4.w;
                    }
                }
            } else {  // the item's last, partial batch
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool ok = g0 + u < s1;
                    v[u] = ok ? gather(s_nbr[t + u]) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (WEIGHTED) wv[u] = ok ? s_w[t + u] : 0.f;
                }
            }
            if (re >= e && !(re == e && cur < r1)) {  // no row ends inside the batch
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (WEIGHTED) fma4(acc, wv[u], v[u]);
                    else add4(acc, v[u]);
                }
            } else {
                int pos = s;
                while (true) {  // warp-uniform: one round per row segment of the batch
                    const int seg_end = re < e ? re : e;
                    const int lo = pos - g0, hi = seg_end - g0;  // this group's slots u in [lo, hi) belong to the segment
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (u >= lo && u < hi) {
                            if (WEIGHTED) fma4(acc, wv[u], v[u]);
                            else add4(acc, v[u]);
                        }
                    }
                    if (re > e || cur >= r1) break;  // the row continues (or its marker belongs to the next item)
                    finalize(cur);
                    ++cur;
                    re = cur < a.n ? s_rp[cur - r0 + 1] : 0x7fffffff;
                    pos = seg_end;
                }
            }
        }
        while (cur < r1) {  // markers after the last slot of the item (empty rows, marker-only items)
            finalize(cur);
            ++cur;
        }
        // the row still open at the end of the item (row r1): its partial, possibly all zero
        mpg_reduce_groups<G>(acc);
        if (writer) reinterpret_cast<float4*>(a.carry + (int64_t)item * a.f)[gl] = acc;

        item = __shfl_sync(0xffffffffu, next, 0);
    }
}

// G lanes per item; same combination order as spmm_mp_fixup_kernel.
template <int G, bool PEER>
__global__ void __launch_bounds__(kMpThreads)
    spmm_mpg_fixup_kernel(const __grid_constant__ MpArgs a, const __grid_constant__ PeerOut po) {
    const int64_t idx = (int64_t)blockIdx.x * kMpThreads + threadIdx.x;
    const int64_t fi = idx / G;
    const int gl = (int)(idx % G);
    if (fi < 1 || fi >= a.items) return;
    const int f_item = (int)fi;
    const int r0 = __ldg(a.item_row + f_item), s0 = __ldg(a.item_slot + f_item);
    const int r1 = __ldg(a.item_row + f_item + 1);
    if (r0 >= a.n || r0 >= r1) return;
    const int rb = __ldg(a.rowptr + r0);
    if (s0 <= rb) return;
    if (gl >= (a.f >> 2)) return;
    int k0 = f_item - 1;
    while (k0 >= 1 && __ldg(a.item_row + k0) == r0 && __ldg(a.item_slot + k0) > rb) --k0;
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = k0; k < f_item; ++k) add4(r, reinterpret_cast<const float4*>(a.carry + (int64_t)k * a.f)[gl]);
    add4(r, reinterpret_cast<const float4*>(a.head + (int64_t)f_item * a.f)[gl]);
    if (a.reduce == GG_MEAN) {
        const int deg = __ldg(a.rowptr + r0 + 1) - rb;
        const float inv = deg > 0 ? 1.0f / (float)deg : 1.0f;
        r.x *= inv; r.y *= inv; r.z *= inv; r.w *= inv;
    }
    if (a.x_self) fma4(r, a.self_scale, __ldg(reinterpret_cast<const float4*>(a.x_self + (int64_t)r0 * a.ld_self) + gl));
    if (a.bias) add4(r, __ldg(reinterpret_cast<const float4*>(a.bias) + gl));
    if (a.r1_s) fma4(r, __ldg(a.r1_s + r0), __ldg(reinterpret_cast<const float4*>(a.r1_v) + gl));
    if (a.r2_s) fma4(r, __ldg(a.r2_s + r0), __ldg(reinterpret_cast<const float4*>(a.r2_v) + gl));
    mpg_out_row<PEER>(a, po, r0)[gl] = r;
}

template <int G, bool WEIGHTED, bool PEER>
static void launch_mpg_one(const MpArgs& a, const PeerOut& po, cudaStream_t st) {
    static bool done = false;
    if (!done) {
        cudaFuncSetAttribute(spmm_mpg_kernel<G, WEIGHTED, PEER>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
        done = true;
    }
    int grid = (int)ceil_div(a.items, kMpWarps);
    if (grid > kNumSMs * 4) grid = kNumSMs * 4;
    spmm_mpg_kernel<G, WEIGHTED, PEER><<<grid, kMpThreads, 0, st>>>(a, po);
    count_launch();
    spmm_mpg_fixup_kernel<G, PEER><<<(int)ceil_div((int64_t)a.items * G, kMpThreads), kMpThreads, 0, st>>>(a, po);
    count_launch();
}

template <int G>
static void launch_mpg(const MpArgs& a, const PeerOut& po, bool peer, cudaStream_t st) {
    if (a.w) {
        if (peer) launch_mpg_one<G, true, true>(a, po, st);
        else launch_mpg_one<G, true, false>(a, po, st);
    } else {
        if (peer) launch_mpg_one<G, false, true>(a, po, st);
        else launch_mpg_one<G, false, false>(a, po, st);
    }
}

static inline int mpg_lanes(int64_t f) {
    if (f <= 0 || f % 4 != 0 || f > 128) return 0;
    const int nvec = (int)(f / 4);
    return nvec <= 4 ? 4 : nvec <= 8 ? 8 : nvec <= 16 ? 16 : 32;
}

static inline bool mp_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int VPL, bool WEIGHTED, bool TMA, bool DEEP, bool RANK1 = false>
static void launch_one(const MpArgs& a, int grid, cudaStream_t st) {
    static bool done = false;  // up to 4 CTAs x 47 KB of tiles per SM: ask for the large carve-out once
    if (!done) {
        cudaFuncSetAttribute(spmm_mp_kernel<VPL, WEIGHTED, TMA, DEEP, RANK1>,
                             cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        done = true;
    }
    spmm_mp_kernel<VPL, WEIGHTED, TMA, DEEP, RANK1><<<grid, kMpThreads, 0, st>>>(a);
}

template <int VPL>
static void launch_mp(const MpArgs& a, bool tma, bool deep, int grid, cudaStream_t st) {
    if (a.r1_s || a.r2_s) {  // rank-1 epilogue terms: a separate instantiation keeps the common kernel lean
        if (a.w) launch_one<VPL, true, false, false, true>(a, grid, st);
        else launch_one<VPL, false, false, false, true>(a, grid, st);
    } else if (deep && VPL <= 2) {
        if (a.w) {
            if (tma) launch_one<VPL, true, true, true>(a, grid, st);
            else launch_one<VPL, true, false, true>(a, grid, st);
        } else {
            if (tma) launch_one<VPL, false, true, true>(a, grid, st);
            else launch_one<VPL, false, false, true>(a, grid, st);
        }
    } else if (a.w) {
        if (tma) launch_one<VPL, true, true, false>(a, grid, st);
        else launch_one<VPL, true, false, false>(a, grid, st);
    } else {
        if (tma) launch_one<VPL, false, true, false>(a, grid, st);
        else launch_one<VPL, false, false, false>(a, grid, st);
    }
    count_launch();
    spmm_mp_fixup_kernel<VPL><<<(int)ceil_div(a.items, kMpWarps), kMpThreads, 0, st>>>(a);
    count_launch();
}

}  // namespace gg

using namespace gg;

extern "C" {

int gg_spmm_plan_units(int64_t num_rows, int64_t num_slots) {
    // enough items to give every resident warp (148 SMs x 32) a few, capped by the shared-memory tile
    int64_t total = num_rows + num_slots;
    int64_t u = total / ((int64_t)kNumSMs * 32 * 4);
    if (u > kMpMaxUnits) u = kMpMaxUnits;
    if (u < 64) u = 64;
    return (int)(u / 32 * 32);
}

int64_t gg_spmm_plan_items(int64_t num_rows, int64_t num_slots, int units) {
    if (units <= 0) return 0;
    int64_t it = ceil_div(num_rows + num_slots, units);
    return it < 1 ? 1 : it;
}

int gg_spmm_plan_build(const int32_t* rowptr, int64_t num_rows, int64_t num_slots, int units,
                       int32_t* item_row, int32_t* item_slot, gg_stream_t stream) {
    GG_REQUIRE(num_rows >= 0 && num_slots >= 0, "gg_spmm_plan_build: negative size");
    GG_REQUIRE(units >= 32 && units <= kMpMaxUnits, "gg_spmm_plan_build: units=%d not in [32, %d]", units,
               kMpMaxUnits);
    GG_REQUIRE(rowptr && item_row && item_slot, "gg_spmm_plan_build: null pointer");
    int64_t items = gg_spmm_plan_items(num_rows, num_slots, units);
    GG_REQUIRE(items < ((int64_t)1 << 31) - 1, "gg_spmm_plan_build: too many items");
    int grid = (int)ceil_div(items + 1, 256);
    if (grid > kNumSMs * 8) grid = kNumSMs * 8;
    spmm_plan_kernel<<<grid, 256, 0, as_stream(stream)>>>(rowptr, num_rows, num_slots, units, (int)items,
                                                         item_row, item_slot);
    GG_LAUNCHED();
    return GG_OK;
}

size_t gg_spmm_mp_workspace_bytes(int64_t items, int64_t f) {
    return 256 + 2 * align_up((size_t)items * (size_t)f * sizeof(float), 256);
}

int gg_spmm_mp_f32(const int32_t* rowptr, const int32_t* nbr, const float* w_slot,
                   const int32_t* item_row, const int32_t* item_slot, int64_t items, const float* x,
                   int64_t ldx, float* out, int64_t ldo, int64_t num_rows, int64_t f, int reduce,
                   const float* x_self, int64_t ld_self, float self_scale, const float* bias,
                   const float* r1_s, const float* r1_v, const float* r2_s, const float* r2_v,
                   void* workspace, size_t workspace_bytes, int stage_mode, gg_stream_t stream) {
    GG_REQUIRE(num_rows >= 0 && f >= 0 && items >= 0, "gg_spmm_mp_f32: negative size");
    GG_REQUIRE((!r1_s || r1_v) && (!r2_s || r2_v), "gg_spmm_mp_f32: rank-1 term without its vector");
    GG_REQUIRE(reduce == GG_SUM || reduce == GG_MEAN, "gg_spmm_mp_f32: reduce=%d", reduce);
    if (num_rows == 0 || f == 0) return GG_OK;
    GG_REQUIRE(rowptr && item_row && item_slot && x && out && workspace, "gg_spmm_mp_f32: null pointer");
    GG_REQUIRE(items >= 1 && items < ((int64_t)1 << 31) - 1, "gg_spmm_mp_f32: items out of range");
    bool vec = (f % 4 == 0) && f <= 1024 && (ldx % 4 == 0) && (ldo % 4 == 0) && mp_aligned16(x) &&
               mp_aligned16(out) && (!x_self || (ld_self % 4 == 0 && mp_aligned16(x_self))) &&
               (!bias || mp_aligned16(bias)) && (!r1_v || mp_aligned16(r1_v)) && (!r2_v || mp_aligned16(r2_v));
    if (!vec) {
        set_error("gg_spmm_mp_f32: needs f %% 4 == 0, f <= 1024 and 16-byte aligned rows (use gg_spmm_f32)");
        return GG_ERR_UNSUPPORTED;
    }
    if (workspace_bytes < gg_spmm_mp_workspace_bytes(items, f)) {
        set_error("gg_spmm_mp_f32: workspace %zu < %zu", workspace_bytes, gg_spmm_mp_workspace_bytes(items, f));
        return GG_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    Carver c(workspace);
    int* counter = c.take<int>(64);
    float* carry = c.take<float>((size_t)items * f);
    float* head = c.take<float>((size_t)items * f);
    GG_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
    MpArgs a{rowptr, nbr, w_slot, item_row, item_slot, (int)items, x, ldx, out, ldo, num_rows, (int)f,
             reduce, x_self, ld_self, self_scale, bias, counter, carry, head, r1_s, r1_v, r2_s, r2_v,
             (stage_mode & 4) ? 1 : 0};
    // TMA staging needs 16-byte aligned index / weight arrays; otherwise plain loads
    // stage_mode bit 0: 0 = cp.async.bulk staging, 1 = plain loads; bit 1: deep gather batches
    bool tma = (stage_mode & 1) == 0 && mp_aligned16(nbr) && (!w_slot || mp_aligned16(w_slot));
    bool deep = (stage_mode & 2) != 0;
    const int nvec = (int)(f / 4);
    int per_sm = nvec <= 32 ? (deep ? 3 : 4) : 2;
    int grid = (int)ceil_div(items, kMpWarps);
    if (grid > kNumSMs * per_sm) grid = kNumSMs * per_sm;
    if (nvec <= 32) launch_mp<1>(a, tma, deep, grid, st);
    else if (nvec <= 64) launch_mp<2>(a, tma, deep, grid, st);
    else if (nvec <= 128) launch_mp<4>(a, tma, deep, grid, st);
    else launch_mp<8>(a, tma, deep, grid, st);
    GG_CUDA(cudaPeekAtLastError());
    return GG_OK;
}

int gg_spmm_group_lanes(int64_t f) { return mpg_lanes(f); }

int gg_spmm_mpg_f32(const int32_t* rowptr, const int32_t* nbr, const float* w_slot, const int32_t* item_row,
                    const int32_t* item_slot, int64_t items, const float* x, int64_t ldx, float* out,
                    int64_t ldo, float* const* out_peers_host, int world, int64_t rows_per_rank,
                    int64_t num_rows, int64_t f, int reduce, const float* x_self, int64_t ld_self,
                    float self_scale, const float* bias, const float* r1_s, const float* r1_v, const float* r2_s,
                    const float* r2_v, void* workspace, size_t workspace_bytes, int flags, gg_stream_t stream) {
    GG_REQUIRE(num_rows >= 0 && f >= 0 && items >= 0, "gg_spmm_mpg_f32: negative size");
    GG_REQUIRE((!r1_s || (r1_v && mp_aligned16(r1_v))) && (!r2_s || (r2_v && mp_aligned16(r2_v))),
               "gg_spmm_mpg_f32: rank-1 term without its (16-byte aligned) vector");
    GG_REQUIRE(reduce == GG_SUM || reduce == GG_MEAN, "gg_spmm_mpg_f32: reduce=%d", reduce);
    if (num_rows == 0 || f == 0) return GG_OK;
    const int g = mpg_lanes(f);
    if (!g) {
        set_error("gg_spmm_mpg_f32: needs f %% 4 == 0 and f <= 128 (got %lld)", (long long)f);
        return GG_ERR_UNSUPPORTED;
    }
    const bool peer = out_peers_host != nullptr;
    GG_REQUIRE(rowptr && item_row && item_slot && x && workspace && (peer || out), "gg_spmm_mpg_f32: null pointer");
    GG_REQUIRE(items >= 1 && items < ((int64_t)1 << 31) - 1, "gg_spmm_mpg_f32: items out of range");
    GG_REQUIRE(num_rows < ((int64_t)1 << 31), "gg_spmm_mpg_f32: too many rows");
    GG_REQUIRE(ldx >= f && ldx < ((int64_t)1 << 30), "gg_spmm_mpg_f32: ldx=%lld out of range", (long long)ldx);
    GG_REQUIRE((ldx % 4 == 0) && (ldo % 4 == 0) && mp_aligned16(x) && (!x_self || (ld_self % 4 == 0 && mp_aligned16(x_self))) &&
                   (!bias || mp_aligned16(bias)), "gg_spmm_mpg_f32: rows must be 16-byte aligned");
    PeerOut po{};
    po.per = 1;
    if (peer) {
        GG_REQUIRE(world >= 1 && world <= GG_PEER_MAX && rows_per_rank >= 1 && rows_per_rank < ((int64_t)1 << 31) &&
                       rows_per_rank * world >= num_rows,
                   "gg_spmm_mpg_f32: world=%d rows_per_rank=%lld do not cover %lld rows", world,
                   (long long)rows_per_rank, (long long)num_rows);
        for (int i = 0; i < world; ++i) {
            GG_REQUIRE(out_peers_host[i] && mp_aligned16(out_peers_host[i]), "gg_spmm_mpg_f32: peer block %d null or misaligned", i);
            po.out[i] = out_peers_host[i];
        }
        po.per = (int)rows_per_rank;
    } else {
        GG_REQUIRE(mp_aligned16(out), "gg_spmm_mpg_f32: out must be 16-byte aligned");
    }
    if (workspace_bytes < gg_spmm_mp_workspace_bytes(items, f)) {
        set_error("gg_spmm_mpg_f32: workspace %zu < %zu", workspace_bytes, gg_spmm_mp_workspace_bytes(items, f));
        return GG_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    Carver c(workspace);
    int* counter = c.take<int>(64);
    float* carry = c.take<float>((size_t)items * f);
    float* head = c.take<float>((size_t)items * f);
    GG_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
    MpArgs a{rowptr, nbr, w_slot, item_row, item_slot, (int)items, x, ldx, out, ldo, num_rows, (int)f,
             reduce, x_self, ld_self, self_scale, bias, counter, carry, head, r1_s, r1_v, r2_s, r2_v,
             (flags & 4) ? 1 : 0};
    if (g == 4) launch_mpg<4>(a, po, peer, st);
    else if (g == 8) launch_mpg<8>(a, po, peer, st);
    else if (g == 16) launch_mpg<16>(a, po, peer, st);
    else launch_mpg<32>(a, po, peer, st);
    GG_CUDA(cudaPeekAtLastError());
    return GG_OK;
}

}  // extern "C"
