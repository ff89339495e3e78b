// GAT (heads = 1) fused into the sliced-ELL aggregation (ref: GATIDConvLayer.message / update,
// graphgym/contrib/layer/idconv.py:317-342; PyG softmax: exp(z - max) / (sum + 1e-16)).
//
// gat_mp.cu materialises alpha: gat_alpha (E' floats out) -> SpMM (alpha in) -> SDDMM (dalpha out) -> gat_dz (alpha,
// dalpha in, dz out) -> gat_csc_gather (alpha, dz gathered through the slot map, alpha_T out) -> SpMM (alpha_T in):
// 20.0 ms per step on the products graph against 11.0 for GCN (profiles/r01_bench_products_gat.json).  Here the edge
// softmax lives inside the three heavy passes and alpha never touches memory:
//   forward      gat_sell_fwd       one walk over the CSR sliced-ELL layout: logits from the per-node halves, ONLINE softmax
//                                   (running max / sum, accumulator rescaled when the max grows) fused with the weighted
//                                   aggregation; saves (max, sum) per row — 8 bytes per node instead of 4 per edge
//   backward     gat_sell_bwd_edge  same layout: dalpha = <g_i, h_j> (4 slots reduced across the row's lanes by one
//                                   multi-value butterfly), alpha recomputed from (max, sum), D_i = <g_i, out_i - b>
//                                   (= sum_e alpha_e dalpha_e), dz written once in CSR slot order, da_tgt per row
//                gat_sell_bwd_src   CSC sliced-ELL layout: alpha recomputed per slot from a 16-byte per-target record
//                                   (a_tgt, max, 1/sum: a 39 MB array, L2 resident), dH_j = sum alpha g_i + rank-1 terms,
//                                   da_src_j = sum dz gathered through a precomposed slot map
// Rows cut into virtual rows (layout `seg`) carry (max, sum, partial accumulator) / partial sums to small fix-up kernels.
// Every row is walked by one group of lanes in slot order: deterministic.
#include <math_constants.h>
#include <stdlib.h>

#include "common.cuh"

namespace gg {

constexpr int kGsRows = 8;        // virtual rows per chunk (= spmm_sell.cu)
constexpr int kGsThreads = 256;
constexpr int32_t kGsNoRow = INT32_MIN;

struct GatSellArgs {
    // layout (CSR for fwd / bwd_edge, CSC for bwd_src)
    const uint32_t* chunk_ptr;
    int chunks;
    const int4* idx4;
    const int4* slot4;       // bwd_edge: CSR slot of every entry; bwd_src: CSR slot of the same edge (composed map)
    const int32_t* vdst;
    const int32_t* hub_rows;
    const int32_t* hub_pptr;
    int hubs;
    int64_t n;
    int f;
    float slope;
    // operands
    const float* h;          // fwd / bwd_edge: gathered rows; bwd_src: g
    int64_t ldh;
    const float* g;          // bwd_edge: gradient rows of the targets
    int64_t ldg;
    const float* fout;       // bwd_edge: forward output (for D_i)
    int64_t ldfo;
    const float* a_tgt;
    const float* a_src;
    const float* bias;
    float2* rowstat;         // [n] (max, sum)
    const float4* tstat;     // bwd_src: [n] (a_tgt, max, 1 / (sum + 1e-16), 0)
    const float* dz_in;      // bwd_src
    const float* da_tgt_in;  // bwd_src: complete da_tgt
    const float* att_src;    // bwd_src: [f]
    const float* att_tgt;
    float* out;              // fwd: out; bwd_src: dh
    int64_t ldo;
    float* dz;               // bwd_edge out (CSR slot order)
    float* da;               // bwd_edge: da_tgt; bwd_src: da_src
    // split rows
    float* pacc;             // [partial rows, f]
    float2* pstat;           // [partial rows]: fwd (max, sum); bwd: (sum, -)
    int* counter;
    int l2_hint;             // feature rows evict_last, index / map / dz streams evict_first (GG_GAT_L2HINT, default on)
    const float* own;        // bwd_one: the row owner's matrix (h of the SOURCE rows); a.h holds the gathered g
    int64_t ld_own;
    // training forward: the part of the aggregation that comes through positive logits (see gat_sell_bwd_one_kernel)
    float* out_pos;          // [n, ld_pos]: sum over {e : z_e > 0} alpha_e h_j
    int64_t ld_pos;
    float* a_pos;            // [n]: sum over {e : z_e > 0} alpha_e
    float* pacc2;            // split rows
    float* ps2;
};

__device__ __forceinline__ float gs_leaky(float z, float slope) { return z > 0.f ? z : slope * z; }
// exp(e - m) with the conventions of a running softmax: a masked slot (e = -inf) weighs 0 even while m is still -inf
// __expf (ex2.approx, ~2 ulp): the full-range expf costs ~10 instructions per slot and lane, which made the first version
// issue bound
__device__ __forceinline__ float gs_p(float e, float m) { return e == -CUDART_INF_F ? 0.f : __expf(e - m); }
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
// index / map units are streamed once: evict_first in L2, so that they do not push out the feature rows (evict_last);
// the first version used plain loads and moved 24-39 GB of DRAM traffic per pass against 17.5 GB for the plain aggregation
__device__ __forceinline__ int4 gs_ldg_i4_pol(const int4* p, uint64_t pol) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
#define gs_ldg_i4(p) gs_ldg_i4_pol((p), pol_stream)

// the row a virtual row belongs to: itself, or (piece of a split row) the hub whose partial range holds the piece
__device__ __forceinline__ int gs_row_of(const GatSellArgs& a, int d) {
    if (d >= 0 || d == kGsNoRow) return d;
    const int p = -d - 1;
    int lo = 0, hi = a.hubs - 1;   // last h with hub_pptr[h] <= p
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(a.hub_pptr + mid) <= p) lo = mid;
        else hi = mid - 1;
    }
    return __ldg(a.hub_rows + lo);
}

template <int G>
__device__ __forceinline__ float gs_group_sum(float v) {
#pragma unroll
    for (int m = G / 2; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// Four per-lane partial sums -> every lane ends with the complete sum (over its group of G lanes) of ONE of them:
// value kcls = 2 * bit(G/2) + bit(G/4) of its lane id.  log2(G) + 1 shuffles instead of 4 log2(G).
template <int G>
__device__ __forceinline__ float gs_reduce4(float p0, float p1, float p2, float p3, int gl) {
    const bool up = (gl & (G / 2)) != 0;
    const float s0 = up ? p0 : p2, s1 = up ? p1 : p3;          // what the partner keeps
    float q0 = (up ? p2 : p0) + __shfl_xor_sync(0xffffffffu, s0, G / 2);
    float q1 = (up ? p3 : p1) + __shfl_xor_sync(0xffffffffu, s1, G / 2);
    const bool up2 = (gl & (G / 4)) != 0;
    float r = (up2 ? q1 : q0) + __shfl_xor_sync(0xffffffffu, up2 ? q0 : q1, G / 4);
#pragma unroll
    for (int m = G / 8; m > 0; m >>= 1) r += __shfl_xor_sync(0xffffffffu, r, m);
    return r;
}
__device__ __forceinline__ int gs_pick(const int4& v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }
__device__ __forceinline__ int gs_pick8(const int4& a, const int4& b, int k) { return k < 4 ? gs_pick(a, k) : gs_pick(b, k - 4); }

// Per-slot SCALAR work of a pair of units (8 slots: a logit, a weight, a dz gather) is done by ONE lane per slot and
// broadcast with shuffles: lane gl of a group owns slot (gl & 7) (groups of 4 lanes: slots gl and gl + 4).  Computing
// every slot's scalars in every lane — 8 loads and 8 exps per lane and pair — made the first version issue bound
// (5.3 / 8.4 / 10.4 ms for the three passes against 3.6 ms for the plain aggregation).
template <int G> struct GsOwn {
    static constexpr int kLanes = G >= 8 ? 8 : G;    // lanes of a group that own distinct slots
    static constexpr int kPer = 8 / kLanes;          // slots per owning lane
};
template <int G>
__device__ __forceinline__ float gs_bcast(const float (&loc)[GsOwn<G>::kPer], int k) {   // slot k's value, in every lane
    return __shfl_sync(0xffffffffu, loc[k / GsOwn<G>::kLanes], k % GsOwn<G>::kLanes, G);
}

// ---- forward -------------------------------------------------------------------------------------------------------
// TRAIN: also accumulates out_pos / a_pos, the share of the aggregation that arrives through positive logits.  The backward
// needs da_tgt_i = sum_e dz_e; because sum_e alpha_e (dalpha_e - D_i) = 0 and lrelu' is 1 or `slope`, that sum equals
// (1 - slope) (<g_i, out_pos_i> - D_i a_pos_i): a per-node dot product instead of a transposition of 62 M per-edge values.
template <int G, bool TRAIN>
__device__ __forceinline__ void gat_sell_fwd_body(const GatSellArgs& a) {
    constexpr int S = 32 / G, P = kGsRows / S;
    const int lane = threadIdx.x & 31, grp = lane / G, gl = lane % G;
    const int nvec = a.f >> 2;
    const bool act = gl < nvec;
    const char* __restrict__ xg = reinterpret_cast<const char*>(a.h) + (act ? gl : nvec - 1) * 16;
    uint32_t row_bytes = (uint32_t)a.ldh * 4u;
    asm volatile("" : "+l"(xg), "+r"(row_bytes));
    const uint64_t pol_keep = l2_policy(a.l2_hint ? 1 : 0), pol_stream = l2_policy(a.l2_hint ? 2 : 0);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto gather = [&](int j) {
        float4 v = zero4;
        if (j >= 0) v = ldg_nc_f4_hint(reinterpret_cast<const float4*>(xg + (uint64_t)(uint32_t)j * row_bytes), pol_keep);
        return v;
    };
    int chunk = 0;
    if (lane == 0) chunk = atomicAdd(a.counter, 1);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    while (chunk < a.chunks) {
        int next = 0;
        if (lane == 0) next = atomicAdd(a.counter, 1);
        const uint32_t base = __ldg(a.chunk_ptr + chunk);
        const int nk = (int)(__ldg(a.chunk_ptr + chunk + 1) - base) / kGsRows;   // 4-slot units per row
        const int4 none4 = make_int4(-1, -1, -1, -1);
        // index units are fetched one pair ahead (across the rows of the chunk too): without it every pair paid two
        // serial memory latencies, index load then gathers (5.1 ms against 3.6 ms for the plain aggregation)
        const int4* __restrict__ ip0 = a.idx4 + base + grp;
        int4 na = nk > 0 ? gs_ldg_i4(ip0) : none4;
        int4 nb = nk > 1 ? gs_ldg_i4(ip0 + kGsRows) : none4;
        for (int p = 0; p < P; ++p) {
            const int q = p * S + grp;
            const int d = __ldg(a.vdst + (int64_t)chunk * kGsRows + q);
            const int row = gs_row_of(a, d);
            const float at = row >= 0 ? __ldg(a.a_tgt + row) : 0.f;
            const int4* __restrict__ ip = a.idx4 + base + q;
            float m = -CUDART_INF_F, s = 0.f, s2 = 0.f;
            float4 acc = zero4, acc2 = zero4;
            for (int k4 = 0; k4 < nk; k4 += 2) {
                const int4 ia = na, ib = nb;
                {   // next pair: of this row, or the first of the chunk's next row
                    const bool same = k4 + 2 < nk;
                    const int4* np_ = same ? ip + (int64_t)(k4 + 2) * kGsRows : ip + S;
                    const bool more = same || p + 1 < P;
                    na = more ? gs_ldg_i4(np_) : none4;
                    nb = (same ? k4 + 3 < nk : (p + 1 < P && nk > 1)) ? gs_ldg_i4(np_ + kGsRows) : none4;
                }
                // this lane's slot(s): the source halves of the logits are requested before the feature rows
                float e[GsOwn<G>::kPer];
#pragma unroll
                for (int t = 0; t < GsOwn<G>::kPer; ++t) {
                    const int j = gs_pick8(ia, ib, (gl & 7) + t * GsOwn<G>::kLanes);
                    e[t] = j >= 0 ? __ldg(a.a_src + j) : -CUDART_INF_F;
                }
                const float4 v0 = gather(ia.x), v1 = gather(ia.y), v2 = gather(ia.z), v3 = gather(ia.w);
                const float4 v4 = gather(ib.x), v5 = gather(ib.y), v6 = gather(ib.z), v7 = gather(ib.w);
                float mx = -CUDART_INF_F;
#pragma unroll
                for (int t = 0; t < GsOwn<G>::kPer; ++t) {
                    if (e[t] != -CUDART_INF_F) e[t] = gs_leaky(at + e[t], a.slope);
                    mx = fmaxf(mx, e[t]);
                }
#pragma unroll
                for (int o = GsOwn<G>::kLanes / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                const float mn = fmaxf(m, mx);
                const float sc = m == mn ? 1.f : gs_p(m, mn);     // m = -inf (first slots of the row): 0
                float pl[GsOwn<G>::kPer];
#pragma unroll
                for (int t = 0; t < GsOwn<G>::kPer; ++t) pl[t] = gs_p(e[t], mn);
                const float p0 = gs_bcast<G>(pl, 0), p1 = gs_bcast<G>(pl, 1), p2 = gs_bcast<G>(pl, 2), p3 = gs_bcast<G>(pl, 3);
                const float p4 = gs_bcast<G>(pl, 4), p5 = gs_bcast<G>(pl, 5), p6 = gs_bcast<G>(pl, 6), p7 = gs_bcast<G>(pl, 7);
                acc.x *= sc; acc.y *= sc; acc.z *= sc; acc.w *= sc;
                fma4(acc, p0, v0); fma4(acc, p1, v1); fma4(acc, p2, v2); fma4(acc, p3, v3);
                fma4(acc, p4, v4); fma4(acc, p5, v5); fma4(acc, p6, v6); fma4(acc, p7, v7);
                s = fmaf(s, sc, ((p0 + p1) + (p2 + p3)) + ((p4 + p5) + (p6 + p7)));
                if constexpr (TRAIN) {
                    // which slots have a positive logit (leaky(z) > 0 <=> z > 0): one ballot per owned slot, read at the
                    // owner's lane, instead of eight more broadcasts
                    unsigned pm[GsOwn<G>::kPer];
#pragma unroll
                    for (int t = 0; t < GsOwn<G>::kPer; ++t) pm[t] = __ballot_sync(0xffffffffu, e[t] > 0.f);
                    auto pos = [&](int k, float pk) {
                        return ((pm[k / GsOwn<G>::kLanes] >> (grp * G + k % GsOwn<G>::kLanes)) & 1u) ? pk : 0.f;
                    };
                    const float q0 = pos(0, p0), q1 = pos(1, p1), q2 = pos(2, p2), q3 = pos(3, p3);
                    const float q4 = pos(4, p4), q5 = pos(5, p5), q6 = pos(6, p6), q7 = pos(7, p7);
                    acc2.x *= sc; acc2.y *= sc; acc2.z *= sc; acc2.w *= sc;
                    fma4(acc2, q0, v0); fma4(acc2, q1, v1); fma4(acc2, q2, v2); fma4(acc2, q3, v3);
                    fma4(acc2, q4, v4); fma4(acc2, q5, v5); fma4(acc2, q6, v6); fma4(acc2, q7, v7);
                    s2 = fmaf(s2, sc, ((q0 + q1) + (q2 + q3)) + ((q4 + q5) + (q6 + q7)));
                }
                m = mn;
            }
            if (d >= 0) {
                const float inv = 1.0f / (s + 1e-16f);
                if (act) {
                    float4 r = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
                    if (a.bias) add4(r, __ldg(reinterpret_cast<const float4*>(a.bias) + gl));
                    reinterpret_cast<float4*>(a.out + (int64_t)row * a.ldo)[gl] = r;
                    if constexpr (TRAIN)
                        reinterpret_cast<float4*>(a.out_pos + (int64_t)row * a.ld_pos)[gl] =
                            make_float4(acc2.x * inv, acc2.y * inv, acc2.z * inv, acc2.w * inv);
                }
                if (gl == 0) {
                    a.rowstat[row] = make_float2(m, s);
                    if constexpr (TRAIN) a.a_pos[row] = s2 * inv;
                }
            } else if (d != kGsNoRow) {
                const int pr = -d - 1;
                if (act) {
                    reinterpret_cast<float4*>(a.pacc + (int64_t)pr * a.f)[gl] = acc;
                    if constexpr (TRAIN) reinterpret_cast<float4*>(a.pacc2 + (int64_t)pr * a.f)[gl] = acc2;
                }
                if (gl == 0) {
                    a.pstat[pr] = make_float2(m, s);
                    if constexpr (TRAIN) a.ps2[pr] = s2;
                }
            }
        }
        chunk = __shfl_sync(0xffffffffu, next, 0);
    }
}

template <int G>
__global__ void __launch_bounds__(kGsThreads, 3) gat_sell_fwd_eval_kernel(const __grid_constant__ GatSellArgs a) {
    gat_sell_fwd_body<G, false>(a);
}
template <int G>
__global__ void __launch_bounds__(kGsThreads, 3) gat_sell_fwd_train_kernel(const __grid_constant__ GatSellArgs a) {
    gat_sell_fwd_body<G, true>(a);
}

// split rows: merge the pieces' (max, sum, accumulator) in piece order
__global__ void __launch_bounds__(256) gat_sell_fwd_fixup_kernel(const __grid_constant__ GatSellArgs a) {
    const int nvec = a.f >> 2;
    const int64_t total = (int64_t)a.hubs * nvec;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int hb = (int)(e / nvec), gl = (int)(e - (int64_t)hb * nvec);
        const int p0 = __ldg(a.hub_pptr + hb), p1 = __ldg(a.hub_pptr + hb + 1);
        float M = -CUDART_INF_F;
        for (int p = p0; p < p1; ++p) M = fmaxf(M, a.pstat[p].x);
        float S = 0.f, S2 = 0.f;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f), r2 = r;
        for (int p = p0; p < p1; ++p) {
            const float2 st = a.pstat[p];
            const float w = gs_p(st.x, M);
            S = fmaf(st.y, w, S);
            fma4(r, w, reinterpret_cast<const float4*>(a.pacc + (int64_t)p * a.f)[gl]);
            if (a.out_pos) {
                S2 = fmaf(a.ps2[p], w, S2);
                fma4(r2, w, reinterpret_cast<const float4*>(a.pacc2 + (int64_t)p * a.f)[gl]);
            }
        }
        const float inv = 1.0f / (S + 1e-16f);
        r.x *= inv; r.y *= inv; r.z *= inv; r.w *= inv;
        if (a.bias) add4(r, __ldg(reinterpret_cast<const float4*>(a.bias) + gl));
        const int row = __ldg(a.hub_rows + hb);
        reinterpret_cast<float4*>(a.out + (int64_t)row * a.ldo)[gl] = r;
        if (gl == 0) a.rowstat[row] = make_float2(M, S);
        if (a.out_pos) {
            reinterpret_cast<float4*>(a.out_pos + (int64_t)row * a.ld_pos)[gl] =
                make_float4(r2.x * inv, r2.y * inv, r2.z * inv, r2.w * inv);
            if (gl == 0) a.a_pos[row] = S2 * inv;
        }
    }
}

// ---- backward, edge side (CSR layout) ------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(kGsThreads, 3) gat_sell_bwd_edge_kernel(const __grid_constant__ GatSellArgs a) {
    constexpr int S = 32 / G, P = kGsRows / S;
    const int lane = threadIdx.x & 31, grp = lane / G, gl = lane % G;
    const int nvec = a.f >> 2;
    const bool act = gl < nvec;
    const char* __restrict__ xg = reinterpret_cast<const char*>(a.h) + (act ? gl : nvec - 1) * 16;
    uint32_t row_bytes = (uint32_t)a.ldh * 4u;
    asm volatile("" : "+l"(xg), "+r"(row_bytes));
    const uint64_t pol_keep = l2_policy(a.l2_hint ? 1 : 0), pol_stream = l2_policy(a.l2_hint ? 2 : 0);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto gather = [&](int j) {
        float4 v = zero4;
        if (j >= 0) v = ldg_nc_f4_hint(reinterpret_cast<const float4*>(xg + (uint64_t)(uint32_t)j * row_bytes), pol_keep);
        return v;
    };
    const int kcls = ((gl & (G / 2)) ? 2 : 0) + ((gl & (G / 4)) ? 1 : 0);   // the slot of a unit this lane finishes
    const bool writer = (gl & (G / 4 - 1)) == 0;                           // one lane per class and group
    int chunk = 0;
    if (lane == 0) chunk = atomicAdd(a.counter, 1);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    while (chunk < a.chunks) {
        int next = 0;
        if (lane == 0) next = atomicAdd(a.counter, 1);
        const uint32_t base = __ldg(a.chunk_ptr + chunk);
        const int nk = (int)(__ldg(a.chunk_ptr + chunk + 1) - base) / kGsRows;
        const int4 none4 = make_int4(-1, -1, -1, -1);
        int4 na = none4, nb = none4;
        for (int p = 0; p < P; ++p) {
            const int q = p * S + grp;
            const int d = __ldg(a.vdst + (int64_t)chunk * kGsRows + q);
            const int row = gs_row_of(a, d);
            const bool live = row >= 0;
            const float at = live ? __ldg(a.a_tgt + row) : 0.f;
            const float2 st = live ? a.rowstat[row] : make_float2(0.f, 1.f);
            const float inv = 1.0f / (st.y + 1e-16f);
            float4 gi = zero4;
            float dpart = 0.f;
            if (live && act) {
                gi = __ldg(reinterpret_cast<const float4*>(a.g + (int64_t)row * a.ldg) + gl);
                float4 o = __ldg(reinterpret_cast<const float4*>(a.fout + (int64_t)row * a.ldfo) + gl);
                if (a.bias) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias) + gl);
                    o.x -= b.x; o.y -= b.y; o.z -= b.z; o.w -= b.w;
                }
                dpart = dot4(gi, o);
            }
            const float D = gs_group_sum<G>(dpart);    // = sum_e alpha_e dalpha_e of the WHOLE row
            const int4* __restrict__ ip = a.idx4 + base + q;
            const int4* __restrict__ sp = a.slot4 + base + q;
            float dsum = 0.f;
            if (p == 0) {   // index units run one pair ahead, across the rows of the chunk too
                na = nk > 0 ? gs_ldg_i4(ip) : none4;
                nb = nk > 1 ? gs_ldg_i4(ip + kGsRows) : none4;
            }
            for (int k4 = 0; k4 < nk; k4 += 2) {
                const int4 ia = na, ib = nb;
                {
                    const bool same = k4 + 2 < nk;
                    const int4* np_ = same ? ip + (int64_t)(k4 + 2) * kGsRows : ip + S;
                    const bool more = same || p + 1 < P;
                    na = more ? gs_ldg_i4(np_) : none4;
                    nb = (same ? k4 + 3 < nk : (p + 1 < P && nk > 1)) ? gs_ldg_i4(np_ + kGsRows) : none4;
                }
                // the finishing lanes' scalars (source logit halves, destination slots) travel with the feature rows:
                // requested here, consumed after the reduction (a dependent load there cost a memory latency per pair)
                const int ja = gs_pick(ia, kcls), jb = gs_pick(ib, kcls);
                float za = 0.f, zb = 0.f;
                int sa = -1, sb = -1;
                if (writer) {
                    if (ja >= 0) { za = __ldg(a.a_src + ja); sa = __ldg(reinterpret_cast<const int*>(sp + (int64_t)k4 * kGsRows) + kcls); }
                    if (jb >= 0) { zb = __ldg(a.a_src + jb); sb = __ldg(reinterpret_cast<const int*>(sp + (int64_t)(k4 + 1) * kGsRows) + kcls); }
                }
                const float4 v0 = gather(ia.x), v1 = gather(ia.y), v2 = gather(ia.z), v3 = gather(ia.w);
                const float4 v4 = gather(ib.x), v5 = gather(ib.y), v6 = gather(ib.z), v7 = gather(ib.w);
                const float da0 = gs_reduce4<G>(dot4(gi, v0), dot4(gi, v1), dot4(gi, v2), dot4(gi, v3), gl);
                const float da1 = gs_reduce4<G>(dot4(gi, v4), dot4(gi, v5), dot4(gi, v6), dot4(gi, v7), gl);
                if (writer) {
                    if (ja >= 0) {
                        const float z = at + za;
                        const float v = __expf(gs_leaky(z, a.slope) - st.x) * inv * (da0 - D) * (z > 0.f ? 1.f : a.slope);
                        a.dz[sa] = v;
                        dsum += v;
                    }
                    if (jb >= 0) {
                        const float z = at + zb;
                        const float v = __expf(gs_leaky(z, a.slope) - st.x) * inv * (da1 - D) * (z > 0.f ? 1.f : a.slope);
                        a.dz[sb] = v;
                        dsum += v;
                    }
                }
            }
            const float da = gs_group_sum<G>(dsum);
            if (gl == 0) {
                if (d >= 0) a.da[row] = da;
                else if (d != kGsNoRow) a.pstat[-d - 1] = make_float2(da, 0.f);
            }
        }
        chunk = __shfl_sync(0xffffffffu, next, 0);
    }
}

__global__ void __launch_bounds__(256) gat_sell_bwd_edge_fixup_kernel(const __grid_constant__ GatSellArgs a) {
    for (int hb = blockIdx.x * blockDim.x + threadIdx.x; hb < a.hubs; hb += gridDim.x * blockDim.x) {
        const int p0 = __ldg(a.hub_pptr + hb), p1 = __ldg(a.hub_pptr + hb + 1);
        float s = 0.f;
        for (int p = p0; p < p1; ++p) s += a.pstat[p].x;
        a.da[__ldg(a.hub_rows + hb)] = s;
    }
}

// ---- backward, source side (CSC layout) ----------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(kGsThreads, 3) gat_sell_bwd_src_kernel(const __grid_constant__ GatSellArgs a) {
    constexpr int S = 32 / G, P = kGsRows / S;
    const int lane = threadIdx.x & 31, grp = lane / G, gl = lane % G;
    const int nvec = a.f >> 2;
    const bool act = gl < nvec;
    const char* __restrict__ xg = reinterpret_cast<const char*>(a.h) + (act ? gl : nvec - 1) * 16;   // a.h = g here
    uint32_t row_bytes = (uint32_t)a.ldh * 4u;
    asm volatile("" : "+l"(xg), "+r"(row_bytes));
    const uint64_t pol_keep = l2_policy(a.l2_hint ? 1 : 0), pol_stream = l2_policy(a.l2_hint ? 2 : 0);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto gather = [&](int j) {
        float4 v = zero4;
        if (j >= 0) v = ldg_nc_f4_hint(reinterpret_cast<const float4*>(xg + (uint64_t)(uint32_t)j * row_bytes), pol_keep);
        return v;
    };
    int chunk = 0;
    if (lane == 0) chunk = atomicAdd(a.counter, 1);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    while (chunk < a.chunks) {
        int next = 0;
        if (lane == 0) next = atomicAdd(a.counter, 1);
        const uint32_t base = __ldg(a.chunk_ptr + chunk);
        const int nk = (int)(__ldg(a.chunk_ptr + chunk + 1) - base) / kGsRows;
        const int4 none4 = make_int4(-1, -1, -1, -1);
        int4 na = none4, nb = none4, nma = none4, nmb = none4;
        for (int p = 0; p < P; ++p) {
            const int q = p * S + grp;
            const int d = __ldg(a.vdst + (int64_t)chunk * kGsRows + q);
            const int row = gs_row_of(a, d);
            const float as = row >= 0 ? __ldg(a.a_src + row) : 0.f;
            const int4* __restrict__ ip = a.idx4 + base + q;
            const int4* __restrict__ mp = a.slot4 + base + q;
            float4 acc = zero4;
            float dsum = 0.f;
            if (p == 0) {   // index and map units run one pair ahead, across the rows of the chunk too
                na = nk > 0 ? gs_ldg_i4(ip) : none4;
                nb = nk > 1 ? gs_ldg_i4(ip + kGsRows) : none4;
                nma = nk > 0 ? gs_ldg_i4(mp) : none4;
                nmb = nk > 1 ? gs_ldg_i4(mp + kGsRows) : none4;
            }
            for (int k4 = 0; k4 < nk; k4 += 2) {
                const int4 ia = na, ib = nb, ma = nma, mb = nmb;
                {
                    const bool same = k4 + 2 < nk;
                    const int64_t off = same ? (int64_t)(k4 + 2) * kGsRows : (int64_t)S;
                    const bool more = same || p + 1 < P;
                    const bool more2 = same ? k4 + 3 < nk : (p + 1 < P && nk > 1);
                    na = more ? gs_ldg_i4(ip + off) : none4;
                    nb = more2 ? gs_ldg_i4(ip + off + kGsRows) : none4;
                    nma = more ? gs_ldg_i4(mp + off) : none4;
                    nmb = more2 ? gs_ldg_i4(mp + off + kGsRows) : none4;
                }
                // this lane's slot(s): the target's record and the edge's dz are requested before the feature rows
                float4 ts[GsOwn<G>::kPer];
                float dzl[GsOwn<G>::kPer];
#pragma unroll
                for (int t = 0; t < GsOwn<G>::kPer; ++t) {
                    const int k = (gl & 7) + t * GsOwn<G>::kLanes;
                    const int i = gs_pick8(ia, ib, k), ms = gs_pick8(ma, mb, k);
                    ts[t] = i >= 0 ? __ldg(a.tstat + i) : make_float4(0.f, 0.f, 0.f, -1.f);   // .w < 0 marks padding
                    dzl[t] = ms >= 0 ? ldg_nc_f32_hint(a.dz_in + ms, pol_stream) : 0.f;
                }
                const float4 v0 = gather(ia.x), v1 = gather(ia.y), v2 = gather(ia.z), v3 = gather(ia.w);
                const float4 v4 = gather(ib.x), v5 = gather(ib.y), v6 = gather(ib.z), v7 = gather(ib.w);
                float wl[GsOwn<G>::kPer];
#pragma unroll
                for (int t = 0; t < GsOwn<G>::kPer; ++t) {
                    // alpha of the edge (source = this row, target = the slot's node), recomputed
                    wl[t] = ts[t].w < 0.f ? 0.f : __expf(gs_leaky(ts[t].x + as, a.slope) - ts[t].y) * ts[t].z;
                    if (gl < GsOwn<G>::kLanes) dsum += dzl[t];       // lanes past the owners hold duplicates
                }
                fma4(acc, gs_bcast<G>(wl, 0), v0); fma4(acc, gs_bcast<G>(wl, 1), v1);
                fma4(acc, gs_bcast<G>(wl, 2), v2); fma4(acc, gs_bcast<G>(wl, 3), v3);
                fma4(acc, gs_bcast<G>(wl, 4), v4); fma4(acc, gs_bcast<G>(wl, 5), v5);
                fma4(acc, gs_bcast<G>(wl, 6), v6); fma4(acc, gs_bcast<G>(wl, 7), v7);
            }
            dsum = gs_group_sum<G>(dsum);
            if (d >= 0) {
                if (act) {   // dH_j = sum alpha g_i + da_src_j att_src + da_tgt_j att_tgt
                    fma4(acc, dsum, __ldg(reinterpret_cast<const float4*>(a.att_src) + gl));
                    fma4(acc, __ldg(a.da_tgt_in + row), __ldg(reinterpret_cast<const float4*>(a.att_tgt) + gl));
                    reinterpret_cast<float4*>(a.out + (int64_t)row * a.ldo)[gl] = acc;
                }
                if (gl == 0) a.da[row] = dsum;
            } else if (d != kGsNoRow) {
                const int pr = -d - 1;
                if (act) reinterpret_cast<float4*>(a.pacc + (int64_t)pr * a.f)[gl] = acc;
                if (gl == 0) a.pstat[pr] = make_float2(dsum, 0.f);
            }
        }
        chunk = __shfl_sync(0xffffffffu, next, 0);
    }
}

__global__ void __launch_bounds__(256) gat_sell_bwd_src_fixup_kernel(const __grid_constant__ GatSellArgs a) {
    const int nvec = a.f >> 2;
    const int64_t total = (int64_t)a.hubs * nvec;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int hb = (int)(e / nvec), gl = (int)(e - (int64_t)hb * nvec);
        const int p0 = __ldg(a.hub_pptr + hb), p1 = __ldg(a.hub_pptr + hb + 1);
        float dsum = 0.f;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = p0; p < p1; ++p) {
            dsum += a.pstat[p].x;
            add4(r, reinterpret_cast<const float4*>(a.pacc + (int64_t)p * a.f)[gl]);
        }
        const int row = __ldg(a.hub_rows + hb);
        fma4(r, dsum, __ldg(reinterpret_cast<const float4*>(a.att_src) + gl));
        fma4(r, __ldg(a.da_tgt_in + row), __ldg(reinterpret_cast<const float4*>(a.att_tgt) + gl));
        reinterpret_cast<float4*>(a.out + (int64_t)row * a.ldo)[gl] = r;
        if (gl == 0) a.da[row] = dsum;
    }
}


// ---- backward in ONE heavy pass (CSC layout) -----------------------------------------------------------------------------
// bwd_edge gathers h_j for every edge to form dalpha = <g_i, h_j>, bwd_src gathers g_i for every edge to form
// dH_j = sum alpha g_i: 2 x 33 GB of gathers on the products graph.  Both need the same pair (g_i, h_j), so one walk over
// the CSC layout does both: the row owner j keeps h_j in registers, every slot gathers g_i once, the dot product gives
// dalpha, alpha is recomputed from the target's record (a_tgt, max, 1/sum, D_i) with D_i = <g_i, out_i - bias> =
// sum_e alpha_e dalpha_e precomputed per node, dz_e = alpha_e (dalpha_e - D_i) lrelu'(z_e) is summed into da_src_j and
// alpha_e g_i into dH_j.  dz never goes to memory: its other marginal, da_tgt_i = sum over the edges INTO i, is known
// before the walk from the training forward's out_pos / a_pos (gat_tstat_d_kernel), so the epilogue also adds
// da_tgt_j att_tgt.  (Two earlier versions moved dz through memory for that sum: scattered into CSR slot order from the
// walk — 2 GB of 32-byte sector write-backs for 250 MB of values, +1.3 ms — or written in place and gathered by a CSR-side
// pass through an inverse slot map — 100 bytes of DRAM per 4-byte read, 2.0 ms.)
template <int G>
__global__ void __launch_bounds__(kGsThreads, 3) gat_sell_bwd_one_kernel(const __grid_constant__ GatSellArgs a) {
    constexpr int S = 32 / G, P = kGsRows / S;
    const int lane = threadIdx.x & 31, grp = lane / G, gl = lane % G;
    const int nvec = a.f >> 2;
    const bool act = gl < nvec;
    const char* __restrict__ xg = reinterpret_cast<const char*>(a.h) + (act ? gl : nvec - 1) * 16;   // a.h = g
    uint32_t row_bytes = (uint32_t)a.ldh * 4u;
    asm volatile("" : "+l"(xg), "+r"(row_bytes));
    const uint64_t pol_keep = l2_policy(a.l2_hint ? 1 : 0), pol_stream = l2_policy(a.l2_hint ? 2 : 0);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto gather = [&](int j) {
        float4 v = zero4;
        if (j >= 0) v = ldg_nc_f4_hint(reinterpret_cast<const float4*>(xg + (uint64_t)(uint32_t)j * row_bytes), pol_keep);
        return v;
    };
    const int kcls = ((gl & (G / 2)) ? 2 : 0) + ((gl & (G / 4)) ? 1 : 0);   // the slot of a unit this lane finishes
    const bool writer = (gl & (G / 4 - 1)) == 0;                           // one lane per class and group
    auto rep = [](int k) { return (k >> 1) * (G / 2) + (k & 1) * (G / 4); };  // group-relative writer lane of class k
    int chunk = 0;
    if (lane == 0) chunk = atomicAdd(a.counter, 1);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    while (chunk < a.chunks) {
        int next = 0;
        if (lane == 0) next = atomicAdd(a.counter, 1);
        const uint32_t base = __ldg(a.chunk_ptr + chunk);
        const int nk = (int)(__ldg(a.chunk_ptr + chunk + 1) - base) / kGsRows;
        const int4 none4 = make_int4(-1, -1, -1, -1);
        int4 na = none4, nb = none4;
        for (int p = 0; p < P; ++p) {
            const int q = p * S + grp;
            const int d = __ldg(a.vdst + (int64_t)chunk * kGsRows + q);
            const int row = gs_row_of(a, d);
            const bool live = row >= 0;
            const float as = live ? __ldg(a.a_src + row) : 0.f;
            float4 hj = zero4;
            if (live && act) hj = ldg_nc_f4_hint(reinterpret_cast<const float4*>(a.own + (int64_t)row * a.ld_own) + gl, pol_stream);
            const int4* __restrict__ ip = a.idx4 + base + q;
            float4 acc = zero4;
            float dsum = 0.f;
            if (p == 0) {
                na = nk > 0 ? gs_ldg_i4(ip) : none4;
                nb = nk > 1 ? gs_ldg_i4(ip + kGsRows) : none4;
            }
            for (int k4 = 0; k4 < nk; k4 += 2) {
                const int4 ia = na, ib = nb;
                {
                    const bool same = k4 + 2 < nk;
                    const int4* np_ = same ? ip + (int64_t)(k4 + 2) * kGsRows : ip + S;
                    const bool more = same || p + 1 < P;
                    na = more ? gs_ldg_i4(np_) : none4;
                    nb = (same ? k4 + 3 < nk : (p + 1 < P && nk > 1)) ? gs_ldg_i4(np_ + kGsRows) : none4;
                }
                // this lane finishes slot kcls of both units: the targets' records and the edges' CSR slots travel with the rows
                const int ta = gs_pick(ia, kcls), tb = gs_pick(ib, kcls);
                float4 tsa = make_float4(0.f, 0.f, -1.f, 0.f), tsb = tsa;    // .z < 0 marks padding
                if (ta >= 0) tsa = ldg_nc_f4_hint(a.tstat + ta, pol_keep);
                if (tb >= 0) tsb = ldg_nc_f4_hint(a.tstat + tb, pol_keep);
                const float4 v0 = gather(ia.x), v1 = gather(ia.y), v2 = gather(ia.z), v3 = gather(ia.w);
                const float4 v4 = gather(ib.x), v5 = gather(ib.y), v6 = gather(ib.z), v7 = gather(ib.w);
                const float da0 = gs_reduce4<G>(dot4(hj, v0), dot4(hj, v1), dot4(hj, v2), dot4(hj, v3), gl);
                const float da1 = gs_reduce4<G>(dot4(hj, v4), dot4(hj, v5), dot4(hj, v6), dot4(hj, v7), gl);
                float wa = 0.f, wb = 0.f;
                if (tsa.z >= 0.f) {
                    const float z = tsa.x + as;
                    wa = __expf(gs_leaky(z, a.slope) - tsa.y) * tsa.z;
                    if (writer) dsum += wa * (da0 - tsa.w) * (z > 0.f ? 1.f : a.slope);
                }
                if (tsb.z >= 0.f) {
                    const float z = tsb.x + as;
                    wb = __expf(gs_leaky(z, a.slope) - tsb.y) * tsb.z;
                    if (writer) dsum += wb * (da1 - tsb.w) * (z > 0.f ? 1.f : a.slope);
                }
                fma4(acc, __shfl_sync(0xffffffffu, wa, rep(0), G), v0); fma4(acc, __shfl_sync(0xffffffffu, wa, rep(1), G), v1);
                fma4(acc, __shfl_sync(0xffffffffu, wa, rep(2), G), v2); fma4(acc, __shfl_sync(0xffffffffu, wa, rep(3), G), v3);
                fma4(acc, __shfl_sync(0xffffffffu, wb, rep(0), G), v4); fma4(acc, __shfl_sync(0xffffffffu, wb, rep(1), G), v5);
                fma4(acc, __shfl_sync(0xffffffffu, wb, rep(2), G), v6); fma4(acc, __shfl_sync(0xffffffffu, wb, rep(3), G), v7);
            }
            dsum = gs_group_sum<G>(dsum);
            if (d >= 0) {
                if (act) {   // dH_j = sum alpha g_i + da_src_j att_src + da_tgt_j att_tgt
                    fma4(acc, dsum, __ldg(reinterpret_cast<const float4*>(a.att_src) + gl));
                    fma4(acc, __ldg(a.da_tgt_in + row), __ldg(reinterpret_cast<const float4*>(a.att_tgt) + gl));
                    __stcs(reinterpret_cast<float4*>(a.out + (int64_t)row * a.ldo) + gl, acc);
                }
                if (gl == 0) a.da[row] = dsum;
            } else if (d != kGsNoRow) {
                const int pr = -d - 1;
                if (act) reinterpret_cast<float4*>(a.pacc + (int64_t)pr * a.f)[gl] = acc;
                if (gl == 0) a.pstat[pr] = make_float2(dsum, 0.f);
            }
        }
        chunk = __shfl_sync(0xffffffffu, next, 0);
    }
}

__global__ void __launch_bounds__(256) gat_sell_bwd_one_fixup_kernel(const __grid_constant__ GatSellArgs a) {
    const int nvec = a.f >> 2;
    const int64_t total = (int64_t)a.hubs * nvec;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int hb = (int)(e / nvec), gl = (int)(e - (int64_t)hb * nvec);
        const int p0 = __ldg(a.hub_pptr + hb), p1 = __ldg(a.hub_pptr + hb + 1);
        float dsum = 0.f;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = p0; p < p1; ++p) {
            dsum += a.pstat[p].x;
            add4(r, reinterpret_cast<const float4*>(a.pacc + (int64_t)p * a.f)[gl]);
        }
        const int row = __ldg(a.hub_rows + hb);
        fma4(r, dsum, __ldg(reinterpret_cast<const float4*>(a.att_src) + gl));
        fma4(r, __ldg(a.da_tgt_in + row), __ldg(reinterpret_cast<const float4*>(a.att_tgt) + gl));
        reinterpret_cast<float4*>(a.out + (int64_t)row * a.ldo)[gl] = r;
        if (gl == 0) a.da[row] = dsum;
    }
}

// per-target record with D_i = <g_i, out_i - bias>, and da_tgt_i = (1 - slope) (<g_i, out_pos_i> - D_i a_pos_i)
// (one warp per row)
__global__ void __launch_bounds__(256) gat_tstat_d_kernel(const float* __restrict__ a_tgt, const float2* __restrict__ rowstat,
                                                          const float* __restrict__ g, int64_t ldg, const float* __restrict__ out,
                                                          int64_t ldo, const float* __restrict__ bias,
                                                          const float* __restrict__ out_pos, int64_t ld_pos,
                                                          const float* __restrict__ a_pos, float slope, int64_t n, int f,
                                                          float4* __restrict__ tstat, float* __restrict__ da_tgt) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * 8;
    // two rows per warp and trip: six 16-byte loads in flight per lane (one row per trip ran at 4.0 TB/s)
    for (int64_t i0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); i0 < n; i0 += 2 * warps) {
        const int64_t i1 = i0 + warps;
        const bool two = i1 < n;
        float ds[2] = {0.f, 0.f}, ps[2] = {0.f, 0.f};
        for (int c = lane * 4; c < f; c += 128) {
            float4 gv[2], ov[2], pv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int64_t i = u ? i1 : i0;
                if (u && !two) { gv[u] = ov[u] = pv[u] = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
                gv[u] = ldg_nc_f4(reinterpret_cast<const float4*>(g + i * ldg + c));
                ov[u] = ldg_nc_f4(reinterpret_cast<const float4*>(out + i * ldo + c));
                pv[u] = ldg_nc_f4(reinterpret_cast<const float4*>(out_pos + i * ld_pos + c));
            }
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bias) b = *reinterpret_cast<const float4*>(bias + c);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                ov[u].x -= b.x; ov[u].y -= b.y; ov[u].z -= b.z; ov[u].w -= b.w;
                ds[u] += dot4(gv[u], ov[u]);
                ps[u] += dot4(gv[u], pv[u]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                ds[u] += __shfl_xor_sync(0xffffffffu, ds[u], o);
                ps[u] += __shfl_xor_sync(0xffffffffu, ps[u], o);
            }
        }
        if (lane < 2 && (lane == 0 || two)) {
            const int64_t i = lane ? i1 : i0;
            const float dsum = lane ? ds[1] : ds[0], psum = lane ? ps[1] : ps[0];
            const float2 st = rowstat[i];
            tstat[i] = make_float4(a_tgt[i], st.x, 1.0f / (st.y + 1e-16f), dsum);
            da_tgt[i] = (1.0f - slope) * (psum - dsum * a_pos[i]);
        }
    }
}

// per-target record of the source-side pass; composed slot map of the CSC layout
__global__ void __launch_bounds__(256) gat_tstat_kernel(const float* __restrict__ a_tgt, const float2* __restrict__ rowstat,
                                                        int64_t n, float4* __restrict__ tstat) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float2 st = rowstat[i];
        tstat[i] = make_float4(a_tgt[i], st.x, 1.0f / (st.y + 1e-16f), 0.f);
    }
}
__global__ void __launch_bounds__(256) sell_compose_map_kernel(const int32_t* __restrict__ slot_of, int64_t total,
                                                               const int32_t* __restrict__ map, int32_t* __restrict__ out) {
    for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < total; d += (int64_t)gridDim.x * blockDim.x) {
        const int s = slot_of[d];
        out[d] = s >= 0 ? map[s] : -1;
    }
}

static inline int gs_grid(int64_t total, int per_block) {
    int64_t b = ceil_div(total, per_block);
    if (b > (int64_t)kNumSMs * 8) b = (int64_t)kNumSMs * 8;
    return (int)(b < 1 ? 1 : b);
}
static inline int gs_lanes(int64_t f) {
    if (f <= 0 || f % 4 != 0 || f > 128) return 0;
    const int nvec = (int)(f / 4);
    return nvec <= 4 ? 4 : nvec <= 8 ? 8 : nvec <= 16 ? 16 : 32;
}
static inline bool gs_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline int gs_l2_hint() {
    static const int v = [] {
        const char* e = getenv("GG_GAT_L2HINT");
        return (e && e[0] == '0') ? 0 : 1;
    }();
    return v;
}

#define GS_DISPATCH(kernel, g, grid, st, a)                                   \
    do {                                                                      \
        if (g == 4) kernel<4><<<grid, kGsThreads, 0, st>>>(a);               \
        else if (g == 8) kernel<8><<<grid, kGsThreads, 0, st>>>(a);          \
        else if (g == 16) kernel<16><<<grid, kGsThreads, 0, st>>>(a);        \
        else kernel<32><<<grid, kGsThreads, 0, st>>>(a);                     \
    } while (0)

static int gs_main_grid(int64_t chunks) {
    int grid = (int)ceil_div(chunks, kGsThreads / 32);
    if (grid > kNumSMs * 3) grid = kNumSMs * 3;
    return grid < 1 ? 1 : grid;
}

}  // namespace gg

using namespace gg;

extern "C" {

size_t gg_gat_sell_workspace_bytes(int64_t partial_rows, int64_t f) {
    const size_t pr = (size_t)(partial_rows > 0 ? partial_rows : 0);
    return 512 + 2 * align_up(pr * (size_t)f * sizeof(float), 256) + align_up(pr * sizeof(float2), 256) +
           align_up(pr * sizeof(float), 256);
}

int gg_sell_compose_map(const int32_t* slot_of, int64_t total, const int32_t* map, int32_t* out, gg_stream_t stream) {
    GG_REQUIRE(total >= 0, "gg_sell_compose_map: negative size");
    if (total == 0) return GG_OK;
    GG_REQUIRE(slot_of && map && out, "gg_sell_compose_map: null pointer");
    sell_compose_map_kernel<<<gs_grid(total, 256 * 4), 256, 0, as_stream(stream)>>>(slot_of, total, map, out);
    GG_LAUNCHED();
    return GG_OK;
}

#define GS_COMMON_CHECKS(who)                                                                                          \
    GG_REQUIRE(n >= 0 && f >= 0 && chunks >= 0 && hubs >= 0 && partial_rows >= 0, who ": negative size");              \
    if (n == 0 || f == 0) return GG_OK;                                                                                \
    const int g_lanes = gs_lanes(f);                                                                                   \
    if (!g_lanes) {                                                                                                    \
        set_error(who ": needs f %% 4 == 0 and f <= 128 (got %lld)", (long long)f);                                    \
        return GG_ERR_UNSUPPORTED;                                                                                     \
    }                                                                                                                  \
    GG_REQUIRE(chunk_ptr && idx && vdst && workspace && (hubs == 0 || (hub_rows && hub_pptr)), who ": null pointer");  \
    GG_REQUIRE(chunks < ((int64_t)1 << 31) && n < ((int64_t)1 << 31), who ": sizes out of range");                     \
    if (workspace_bytes < gg_gat_sell_workspace_bytes(partial_rows, f)) {                                              \
        set_error(who ": workspace %zu < %zu", workspace_bytes, gg_gat_sell_workspace_bytes(partial_rows, f));          \
        return GG_ERR_WORKSPACE;                                                                                       \
    }                                                                                                                  \
    cudaStream_t st = as_stream(stream);                                                                               \
    Carver c(workspace);                                                                                               \
    int* counter = c.take<int>(64);                                                                                    \
    float* pacc = c.take<float>((size_t)partial_rows * f);                                                             \
    float2* pstat = c.take<float2>((size_t)partial_rows);                                                              \
    float* pacc2 = c.take<float>((size_t)partial_rows * f);                                                            \
    float* ps2 = c.take<float>((size_t)partial_rows);                                                                  \
    (void)pacc2; (void)ps2;                                                                                            \
    GG_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));

int gg_gat_sell_fwd_f32(const uint32_t* chunk_ptr, int64_t chunks, const int32_t* idx, const int32_t* vdst,
                        const int32_t* hub_rows, const int32_t* hub_pptr, int64_t hubs, int64_t partial_rows,
                        const float* h, int64_t ldh, const float* a_tgt, const float* a_src, int64_t n, int64_t f,
                        float slope, const float* bias, float* out, int64_t ldo, float* rowstat, float* out_pos,
                        int64_t ld_pos, float* a_pos, void* workspace, size_t workspace_bytes, gg_stream_t stream) {
    GS_COMMON_CHECKS("gg_gat_sell_fwd_f32")
    GG_REQUIRE((out_pos == nullptr) == (a_pos == nullptr) && (!out_pos || (gs_al16(out_pos) && ld_pos >= f && ld_pos % 4 == 0)),
               "gg_gat_sell_fwd_f32: out_pos and a_pos come together, rows 16-byte aligned");
    GG_REQUIRE(h && a_tgt && a_src && out && rowstat && ldh >= f && ldh < ((int64_t)1 << 30) && ldh % 4 == 0 &&
                   ldo % 4 == 0 && gs_al16(h) && gs_al16(out) && gs_al16(idx) && (!bias || gs_al16(bias)) &&
                   (reinterpret_cast<uintptr_t>(rowstat) & 7) == 0,
               "gg_gat_sell_fwd_f32: bad operands (rows must be 16-byte aligned)");
    GatSellArgs a{};
    a.chunk_ptr = chunk_ptr; a.chunks = (int)chunks; a.idx4 = reinterpret_cast<const int4*>(idx); a.vdst = vdst;
    a.hub_rows = hub_rows; a.hub_pptr = hub_pptr; a.hubs = (int)hubs; a.n = n; a.f = (int)f; a.slope = slope;
    a.h = h; a.ldh = ldh; a.a_tgt = a_tgt; a.a_src = a_src; a.bias = bias; a.out = out; a.ldo = ldo;
    a.rowstat = reinterpret_cast<float2*>(rowstat); a.pacc = pacc; a.pstat = pstat; a.counter = counter; a.l2_hint = gs_l2_hint();
    a.out_pos = out_pos; a.ld_pos = ld_pos; a.a_pos = a_pos; a.pacc2 = pacc2; a.ps2 = ps2;
    if (out_pos) {
        GS_DISPATCH(gat_sell_fwd_train_kernel, g_lanes, gs_main_grid(chunks), st, a);
    } else {
        GS_DISPATCH(gat_sell_fwd_eval_kernel, g_lanes, gs_main_grid(chunks), st, a);
    }
    GG_LAUNCHED();
    if (hubs > 0) {
        gat_sell_fwd_fixup_kernel<<<gs_grid(hubs * (f / 4), 256), 256, 0, st>>>(a);
        GG_LAUNCHED();
    }
    return GG_OK;
}

int gg_gat_sell_bwd_edge_f32(const uint32_t* chunk_ptr, int64_t chunks, const int32_t* idx, const int32_t* slot_of,
                             const int32_t* vdst, const int32_t* hub_rows, const int32_t* hub_pptr, int64_t hubs,
                             int64_t partial_rows, const float* h, int64_t ldh, const float* g, int64_t ldg,
                             const float* fwd_out, int64_t ld_out, const float* bias, const float* a_tgt,
                             const float* a_src, const float* rowstat, int64_t n, int64_t f, float slope, float* dz,
                             float* da_tgt, void* workspace, size_t workspace_bytes, gg_stream_t stream) {
    GS_COMMON_CHECKS("gg_gat_sell_bwd_edge_f32")
    GG_REQUIRE(h && g && fwd_out && a_tgt && a_src && rowstat && dz && da_tgt && slot_of && ldh >= f &&
                   ldh < ((int64_t)1 << 30) && ldh % 4 == 0 && ldg % 4 == 0 && ld_out % 4 == 0 && gs_al16(h) && gs_al16(g) &&
                   gs_al16(fwd_out) && gs_al16(idx) && gs_al16(slot_of) && (!bias || gs_al16(bias)),
               "gg_gat_sell_bwd_edge_f32: bad operands (rows must be 16-byte aligned)");
    GatSellArgs a{};
    a.chunk_ptr = chunk_ptr; a.chunks = (int)chunks; a.idx4 = reinterpret_cast<const int4*>(idx);
    a.slot4 = reinterpret_cast<const int4*>(slot_of); a.vdst = vdst; a.hub_rows = hub_rows; a.hub_pptr = hub_pptr;
    a.hubs = (int)hubs; a.n = n; a.f = (int)f; a.slope = slope; a.h = h; a.ldh = ldh; a.g = g; a.ldg = ldg;
    a.fout = fwd_out; a.ldfo = ld_out; a.bias = bias; a.a_tgt = a_tgt; a.a_src = a_src;
    a.rowstat = reinterpret_cast<float2*>(const_cast<float*>(rowstat)); a.dz = dz; a.da = da_tgt; a.pacc = pacc;
    a.pstat = pstat; a.counter = counter; a.l2_hint = gs_l2_hint();
    GS_DISPATCH(gat_sell_bwd_edge_kernel, g_lanes, gs_main_grid(chunks), st, a);
    GG_LAUNCHED();
    if (hubs > 0) {
        gat_sell_bwd_edge_fixup_kernel<<<gs_grid(hubs, 256), 256, 0, st>>>(a);
        GG_LAUNCHED();
    }
    return GG_OK;
}

int gg_gat_sell_bwd_src_f32(const uint32_t* chunk_ptr, int64_t chunks, const int32_t* idx, const int32_t* edge_map,
                            const int32_t* vdst, const int32_t* hub_rows, const int32_t* hub_pptr, int64_t hubs,
                            int64_t partial_rows, const float* g, int64_t ldg, const float* a_tgt, const float* a_src,
                            const float* rowstat, const float* dz, const float* da_tgt, const float* att_src,
                            const float* att_tgt, int64_t n, int64_t f, float slope, float* dh, int64_t ld_dh,
                            float* da_src, float* tstat_scratch, void* workspace, size_t workspace_bytes,
                            gg_stream_t stream) {
    GS_COMMON_CHECKS("gg_gat_sell_bwd_src_f32")
    GG_REQUIRE(g && a_tgt && a_src && rowstat && dz && da_tgt && att_src && att_tgt && dh && da_src && tstat_scratch &&
                   edge_map && ldg >= f && ldg < ((int64_t)1 << 30) && ldg % 4 == 0 && ld_dh % 4 == 0 && gs_al16(g) &&
                   gs_al16(dh) && gs_al16(idx) && gs_al16(edge_map) && gs_al16(att_src) && gs_al16(att_tgt) &&
                   gs_al16(tstat_scratch),
               "gg_gat_sell_bwd_src_f32: bad operands (rows must be 16-byte aligned)");
    gat_tstat_kernel<<<gs_grid(n, 256), 256, 0, st>>>(a_tgt, reinterpret_cast<const float2*>(rowstat), n,
                                                      reinterpret_cast<float4*>(tstat_scratch));
    GG_LAUNCHED();
    GatSellArgs a{};
    a.chunk_ptr = chunk_ptr; a.chunks = (int)chunks; a.idx4 = reinterpret_cast<const int4*>(idx);
    a.slot4 = reinterpret_cast<const int4*>(edge_map); a.vdst = vdst; a.hub_rows = hub_rows; a.hub_pptr = hub_pptr;
    a.hubs = (int)hubs; a.n = n; a.f = (int)f; a.slope = slope; a.h = g; a.ldh = ldg; a.a_src = a_src;
    a.tstat = reinterpret_cast<const float4*>(tstat_scratch); a.dz_in = dz; a.da_tgt_in = da_tgt; a.att_src = att_src;
    a.att_tgt = att_tgt; a.out = dh; a.ldo = ld_dh; a.da = da_src; a.pacc = pacc; a.pstat = pstat; a.counter = counter;
    a.l2_hint = gs_l2_hint();
    GS_DISPATCH(gat_sell_bwd_src_kernel, g_lanes, gs_main_grid(chunks), st, a);
    GG_LAUNCHED();
    if (hubs > 0) {
        gat_sell_bwd_src_fixup_kernel<<<gs_grid(hubs * (f / 4), 256), 256, 0, st>>>(a);
        GG_LAUNCHED();
    }
    return GG_OK;
}

int gg_gat_sell_bwd_one_f32(const uint32_t* chunk_ptr, int64_t chunks, const int32_t* idx, const int32_t* vdst,
                            const int32_t* hub_rows, const int32_t* hub_pptr, int64_t hubs, int64_t partial_rows,
                            const float* h, int64_t ldh, const float* g, int64_t ldg, const float* fwd_out, int64_t ld_out,
                            const float* bias, const float* out_pos, int64_t ld_pos, const float* a_pos, const float* a_tgt,
                            const float* a_src, const float* rowstat, const float* att_src, const float* att_tgt, int64_t n,
                            int64_t f, float slope, float* dh, int64_t ld_dh, float* da_tgt, float* da_src,
                            float* tstat_scratch, void* workspace, size_t workspace_bytes, gg_stream_t stream) {
    GS_COMMON_CHECKS("gg_gat_sell_bwd_one_f32")
    GG_REQUIRE(h && g && fwd_out && out_pos && a_pos && a_tgt && a_src && rowstat && att_src && att_tgt && dh && da_tgt && da_src &&
                   tstat_scratch && ldg >= f && ldg < ((int64_t)1 << 30) && ldg % 4 == 0 && ldh % 4 == 0 && ld_out % 4 == 0 &&
                   ld_pos % 4 == 0 && ld_dh % 4 == 0 && gs_al16(h) && gs_al16(g) && gs_al16(fwd_out) && gs_al16(out_pos) &&
                   gs_al16(dh) && gs_al16(idx) && gs_al16(att_src) && gs_al16(att_tgt) && gs_al16(tstat_scratch) &&
                   (!bias || gs_al16(bias)),
               "gg_gat_sell_bwd_one_f32: bad operands (rows must be 16-byte aligned)");
    gat_tstat_d_kernel<<<gs_grid(n, 8), 256, 0, st>>>(a_tgt, reinterpret_cast<const float2*>(rowstat), g, ldg, fwd_out, ld_out,
                                                      bias, out_pos, ld_pos, a_pos, slope, n, (int)f,
                                                      reinterpret_cast<float4*>(tstat_scratch), da_tgt);
    GG_LAUNCHED();
    GatSellArgs a{};
    a.chunk_ptr = chunk_ptr; a.chunks = (int)chunks; a.idx4 = reinterpret_cast<const int4*>(idx);
    a.vdst = vdst; a.hub_rows = hub_rows; a.hub_pptr = hub_pptr;
    a.hubs = (int)hubs; a.n = n; a.f = (int)f; a.slope = slope; a.h = g; a.ldh = ldg; a.own = h; a.ld_own = ldh;
    a.a_src = a_src; a.tstat = reinterpret_cast<const float4*>(tstat_scratch); a.att_src = att_src; a.att_tgt = att_tgt;
    a.out = dh; a.ldo = ld_dh; a.da_tgt_in = da_tgt; a.da = da_src; a.pacc = pacc; a.pstat = pstat; a.counter = counter;
    a.l2_hint = gs_l2_hint();
    GS_DISPATCH(gat_sell_bwd_one_kernel, g_lanes, gs_main_grid(chunks), st, a);
    GG_LAUNCHED();
    if (hubs > 0) {
        gat_sell_bwd_one_fixup_kernel<<<gs_grid(hubs * (f / 4), 256), 256, 0, st>>>(a);
        GG_LAUNCHED();
    }
    return GG_OK;
}

}  // extern "C"
