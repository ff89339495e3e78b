// GAT at scale (heads = 1): the three gather passes of a GAT layer on the merge-path machinery.
//
// gat.cu fuses softmax and aggregation in one warp-per-row kernel; on power-law graphs that inherits the
// imbalance measured for SpMM v1 (hub rows of 10^4 slots serialise one warp).  Here the layer is split
// into light per-slot passes (4-12 B per edge) and heavy feature-row passes that run on the balanced
// merge-path kernels:
//   forward   gat_alpha        alpha[s] = softmax_i(leaky_relu(a_tgt[i] + a_src[j_s]))    light, warp per row
//             gg_spmm_mp_f32   out = sum_s alpha[s] H[j_s] + bias                           heavy (spmm_mp.cu)
//   backward  gat_sddmm_mp     dalpha[s] = <g_i, H[j_s]>  (merge-path stream over the slots, g_i of the
//                              current row in registers, 8 gathers in flight, 9-shuffle reduction)      heavy
//             gat_dz           D_i = sum alpha dalpha;  dz = alpha (dalpha - D_i) lrelu'(z);  da_tgt      light
//             gat_csc_gather   alpha_T[t] = alpha[map[t]],  da_src[j] = sum_t dz[map[t]]                  light
//             gg_spmm_mp_f32   dH = sum_t alpha_T[t] g[i_t] + da_src att_src + da_tgt att_tgt (rank-1)   heavy
// ref: GATIDConvLayer.message/update graphgym/contrib/layer/idconv.py:317-342, PyG softmax.
#include <math_constants.h>

#include "common.cuh"

namespace gg {

constexpr int kGmWarps = 8;
constexpr int kGmThreads = kGmWarps * 32;

__device__ __forceinline__ float gm_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float gm_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float gm_leaky(float z, float slope) { return z > 0.f ? z : slope * z; }

// one warp per target row: max, sum of exp, alpha (PyG softmax: exp(z - max) / (sum + 1e-16))
__global__ void __launch_bounds__(kGmThreads)
    gat_alpha_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr,
                     const float* __restrict__ a_tgt, const float* __restrict__ a_src, int64_t n, float slope,
                     float* __restrict__ alpha) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kGmWarps + (threadIdx.x >> 5);
    if (row >= n) return;
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    const float at = __ldg(a_tgt + row);
    if (end - beg <= 32) {  // the common case: the whole row in registers, one pass
        const int s = beg + lane;
        const bool in = s < end;
        const float z = in ? gm_leaky(at + __ldg(a_src + __ldg(nbr + s)), slope) : -CUDART_INF_F;
        const float m = gm_warp_max(z);
        const float e = in ? expf(z - m) : 0.f;
        const float inv = 1.0f / (gm_warp_sum(e) + 1e-16f);
        if (in) alpha[s] = e * inv;
        return;
    }
    float m = -CUDART_INF_F;
    for (int s = beg + lane; s < end; s += 32) m = fmaxf(m, gm_leaky(at + __ldg(a_src + __ldg(nbr + s)), slope));
    m = gm_warp_max(m);
    float sum = 0.f;
    for (int s = beg + lane; s < end; s += 32) sum += expf(gm_leaky(at + __ldg(a_src + __ldg(nbr + s)), slope) - m);
    const float inv = 1.0f / (gm_warp_sum(sum) + 1e-16f);
    for (int s = beg + lane; s < end; s += 32)
        alpha[s] = expf(gm_leaky(at + __ldg(a_src + __ldg(nbr + s)), slope) - m) * inv;
}

// one warp per target row: D_i = sum_s alpha dalpha (fixed order), dz, da_tgt
__global__ void __launch_bounds__(kGmThreads)
    gat_dz_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr,
                  const float* __restrict__ a_tgt, const float* __restrict__ a_src,
                  const float* __restrict__ alpha, const float* __restrict__ dalpha, int64_t n, float slope,
                  float* __restrict__ dz, float* __restrict__ da_tgt) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kGmWarps + (threadIdx.x >> 5);
    if (row >= n) return;
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    const float at = __ldg(a_tgt + row);
    float d = 0.f;
    for (int s = beg + lane; s < end; s += 32) d = fmaf(__ldg(alpha + s), __ldg(dalpha + s), d);
    d = gm_warp_sum(d);
    float acc = 0.f;
    for (int s = beg + lane; s < end; s += 32) {
        const float z = at + __ldg(a_src + __ldg(nbr + s));
        const float v = __ldg(alpha + s) * (__ldg(dalpha + s) - d) * (z > 0.f ? 1.f : slope);
        dz[s] = v;
        acc += v;
    }
    acc = gm_warp_sum(acc);
    if (lane == 0) da_tgt[row] = acc;
}

// one warp per source row of the CSC layout: alpha in CSC order, da_src
__global__ void __launch_bounds__(kGmThreads)
    gat_csc_gather_kernel(const int32_t* __restrict__ rowptr_t, const int32_t* __restrict__ map,
                          const float* __restrict__ alpha, const float* __restrict__ dz, int64_t n,
                          float* __restrict__ alpha_t, float* __restrict__ da_src) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kGmWarps + (threadIdx.x >> 5);
    if (row >= n) return;
    const int beg = __ldg(rowptr_t + row), end = __ldg(rowptr_t + row + 1);
    float acc = 0.f;
    for (int t = beg + lane; t < end; t += 32) {
        const int m = __ldg(map + t);
        alpha_t[t] = __ldg(alpha + m);
        acc += __ldg(dz + m);
    }
    acc = gm_warp_sum(acc);
    if (lane == 0) da_src[row] = acc;
}

// ---- merge-path SDDMM: dalpha[s] = <g[row(s),:], h[nbr[s],:]> ---------------------------------------
constexpr int kSdMaxUnits = 480;
constexpr int kSdTile = kSdMaxUnits + 8;

struct SddmmArgs {
    const int32_t* rowptr;
    const int32_t* nbr;
    const int32_t* item_row;
    const int32_t* item_slot;
    int items;
    const float* h;
    int64_t ldh;
    const float* g;
    int64_t ldg;
    int64_t n;
    int f;
    int* counter;
    float* dalpha;
};

template <int VPL>
__global__ void __launch_bounds__(kGmThreads, VPL == 1 ? 4 : 2) gat_sddmm_mp_kernel(SddmmArgs a) {
    constexpr int U = 8;
    __shared__ int32_t s_nbr_all[kGmWarps][kSdTile];
    __shared__ int32_t s_rp_all[kGmWarps][kSdTile];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int32_t* s_nbr = s_nbr_all[wid];
    int32_t* s_rp = s_rp_all[wid];
    const int nvec = a.f >> 2;
    const float4* __restrict__ h4 = reinterpret_cast<const float4*>(a.h);
    const float4* __restrict__ g4 = reinterpret_cast<const float4*>(a.g);
    const int64_t ldh4 = a.ldh >> 2, ldg4 = a.ldg >> 2;
    bool act[VPL];
#pragma unroll
    for (int q = 0; q < VPL; ++q) act[q] = lane + q * 32 < nvec;

    int item = 0;
    if (lane == 0) item = atomicAdd(a.counter, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    while (item < a.items) {
        int next = 0;
        if (lane == 0) next = atomicAdd(a.counter, 1);
        const int r0 = __ldg(a.item_row + item), s0 = __ldg(a.item_slot + item);
        const int r1 = __ldg(a.item_row + item + 1), s1 = __ldg(a.item_slot + item + 1);
        __syncwarp();
        for (int q = s0 + lane; q < s1; q += 32) s_nbr[q - s0] = __ldg(a.nbr + q);
        const int nrp = (r1 < a.n ? r1 + 1 : (int)a.n) - r0 + 1;
        for (int i = lane; i < nrp; i += 32) s_rp[i] = __ldg(a.rowptr + r0 + i);
        __syncwarp();

        int cur = r0;
        int re = r0 < a.n ? s_rp[1] : 0x7fffffff;
        // g of the current row, and of the next one so that a row switch never waits on global memory
        float4 gc[VPL], gn[VPL];
#pragma unroll
        for (int q = 0; q < VPL; ++q) {
            gc[q] = (act[q] && cur < a.n) ? __ldg(g4 + (int64_t)cur * ldg4 + lane + q * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
            gn[q] = (act[q] && cur + 1 < a.n) ? __ldg(g4 + (int64_t)(cur + 1) * ldg4 + lane + q * 32)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int s = s0; s < s1; s += U) {
            float4 v[U][VPL];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool ok = s + u < s1;
                const int j = ok ? s_nbr[s + u - s0] : 0;
                const float4* p = h4 + (int64_t)j * ldh4 + lane;
#pragma unroll
                for (int q = 0; q < VPL; ++q)
                    v[u][q] = (ok && act[q]) ? ldg_nc_f4(p + q * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float p8[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                p8[u] = 0.f;
                if (s + u < s1) {  // warp-uniform
                    while (s + u >= re) {  // this slot belongs to a later row: switch g
                        ++cur;
                        re = s_rp[cur - r0 + 1];
#pragma unroll
                        for (int q = 0; q < VPL; ++q) {
                            gc[q] = gn[q];
                            gn[q] = (act[q] && cur + 1 < a.n) ? __ldg(g4 + (int64_t)(cur + 1) * ldg4 + lane + q * 32)
                                                               : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < VPL; ++q)
                        p8[u] += gc[q].x * v[u][q].x + gc[q].y * v[u][q].y + gc[q].z * v[u][q].z + gc[q].w * v[u][q].w;
                }
            }
            // reduce 8 per-lane partials over the warp with 9 shuffles: halve the value set while folding
            // lane bits 4, 3, 2; the sum of slot k ends up in the 4 lanes with (lane >> 2) == k
            float w4[4], w2[2], w1;
            {
                const bool hi = lane & 16;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float send = hi ? p8[k] : p8[k + 4];
                    const float keep = hi ? p8[k + 4] : p8[k];
                    w4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
            }
            {
                const bool hi = lane & 8;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float send = hi ? w4[k] : w4[k + 2];
                    const float keep = hi ? w4[k + 2] : w4[k];
                    w2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
            }
            {
                const bool hi = lane & 4;
                const float send = hi ? w2[0] : w2[1];
                const float keep = hi ? w2[1] : w2[0];
                w1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            w1 += __shfl_xor_sync(0xffffffffu, w1, 2);
            w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
            // lane bits (4,3,2) select the slot: bit4 -> +4, bit3 -> +2, bit2 -> +1
            const int k = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
            if ((lane & 3) == 0 && s + k < s1) a.dalpha[s + k] = w1;
        }
        item = __shfl_sync(0xffffffffu, next, 0);
    }
}

// ---- narrow-row SDDMM (f <= 128): the feature-sliced exchange's share of dalpha ---------------------------
// Rank c of the row-partitioned path holds column slice c of H and of g for ALL nodes; its contribution to
// dalpha[s] is the dot product over its f = F/P columns (the ranks' contributions are summed afterwards).  Same
// warp layout as spmm_mpg_kernel: 32/G groups of G lanes, a batch = 32/G x 8 consecutive slots, group k owns
// slots [8k, 8k+8); the row of every slot is staged next to its neighbour id, the g row comes through L1
// (consecutive slots share it), the H row is a no-allocate gather; G-lane shuffle reduction per slot.
template <int G>
__global__ void __launch_bounds__(kGmThreads, 4) gat_sddmm_mpg_kernel(const __grid_constant__ SddmmArgs a) {
    constexpr int S = 32 / G, U = 8, B = S * U;
    __shared__ __align__(16) int32_t s_nbr_all[kGmWarps][kSdTile];
    __shared__ __align__(16) int32_t s_row_all[kGmWarps][kSdTile];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int grp = lane / G, gl = lane % G;
    int32_t* s_nbr = s_nbr_all[wid];
    int32_t* s_row = s_row_all[wid];
    const int nvec = a.f >> 2;
    const int col = (gl < nvec ? gl : nvec - 1) * 16;  // idle lanes (f/4 < G) read a duplicate, masked below
    const float keep = gl < nvec ? 1.f : 0.f;
    const char* __restrict__ hb = reinterpret_cast<const char*>(a.h) + col;
    const char* __restrict__ gb = reinterpret_cast<const char*>(a.g) + col;
    const uint32_t h_bytes = (uint32_t)a.ldh * 4u, g_bytes = (uint32_t)a.ldg * 4u;

    int item = 0;
    if (lane == 0) item = atomicAdd(a.counter, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    while (item < a.items) {
        int next = 0;
        if (lane == 0) next = atomicAdd(a.counter, 1);
        const int r0 = __ldg(a.item_row + item), s0 = __ldg(a.item_slot + item);
        const int s1 = __ldg(a.item_slot + item + 1);
        __syncwarp();
        int rr = r0;  // row of this lane's current slot: slots ascend with the lane's stride, rows follow
        for (int q = s0 + lane; q < s1; q += 32) {
            while (__ldg(a.rowptr + rr + 1) <= q) ++rr;
            s_nbr[q - s0] = __ldg(a.nbr + q);
            s_row[q - s0] = rr;
        }
        __syncwarp();
        for (int s = s0; s < s1; s += B) {
            const int g0 = s + grp * U;
            const int t = g0 - s0;
            float dot[U];
            if (s + B <= s1) {
                const int4 j0 = *reinterpret_cast<const int4*>(s_nbr + t), j1 = *reinterpret_cast<const int4*>(s_nbr + t + 4);
                const int4 i0 = *reinterpret_cast<const int4*>(s_row + t), i1 = *reinterpret_cast<const int4*>(s_row + t + 4);
                const int j[U] = {j0.x, j0.y, j0.z, j0.w, j1.x, j1.y, j1.z, j1.w};
                const int i[U] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
                float4 v[U], w[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    v[u] = ldg_nc_f4(reinterpret_cast<const float4*>(hb + (uint64_t)(uint32_t)j[u] * h_bytes));
                    w[u] = __ldg(reinterpret_cast<const float4*>(gb + (uint64_t)(uint32_t)i[u] * g_bytes));
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    dot[u] = keep * (v[u].x * w[u].x + v[u].y * w[u].y + v[u].z * w[u].z + v[u].w * w[u].w);
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    dot[u] = 0.f;
                    if (g0 + u < s1) {
                        const float4 v = ldg_nc_f4(reinterpret_cast<const float4*>(hb + (uint64_t)(uint32_t)s_nbr[t + u] * h_bytes));
                        const float4 w = __ldg(reinterpret_cast<const float4*>(gb + (uint64_t)(uint32_t)s_row[t + u] * g_bytes));
                        dot[u] = keep * (v.x * w.x + v.y * w.y + v.z * w.z + v.w * w.w);
                    }
                }
            }
#pragma unroll
            for (int m = 1; m < G; m <<= 1) {
#pragma unroll
                for (int u = 0; u < U; ++u) dot[u] += __shfl_xor_sync(0xffffffffu, dot[u], m);
            }
            // every lane of the group holds the 8 sums; lane gl stores slot gl (and gl + 4 when G = 4)
#pragma unroll
            for (int u = 0; u < U; ++u)
                if ((u % G) == gl && g0 + u < s1) a.dalpha[g0 + u] = dot[u];
        }
        item = __shfl_sync(0xffffffffu, next, 0);
    }
}

static inline bool gm_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace gg

using namespace gg;

extern "C" {

int gg_gat_alpha_f32(const int32_t* rowptr, const int32_t* nbr, const float* a_tgt, const float* a_src,
                     int64_t n, float slope, float* alpha, gg_stream_t stream) {
    GG_REQUIRE(n >= 0, "gg_gat_alpha_f32: negative size");
    if (n == 0) return GG_OK;
    GG_REQUIRE(rowptr && a_tgt && a_src && alpha, "gg_gat_alpha_f32: null pointer");
    gat_alpha_kernel<<<(int)ceil_div(n, kGmWarps), kGmThreads, 0, as_stream(stream)>>>(rowptr, nbr, a_tgt, a_src, n,
                                                                                       slope, alpha);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_gat_dz_f32(const int32_t* rowptr, const int32_t* nbr, const float* a_tgt, const float* a_src,
                  const float* alpha, const float* dalpha, int64_t n, float slope, float* dz, float* da_tgt,
                  gg_stream_t stream) {
    GG_REQUIRE(n >= 0, "gg_gat_dz_f32: negative size");
    if (n == 0) return GG_OK;
    GG_REQUIRE(rowptr && a_tgt && a_src && alpha && dalpha && dz && da_tgt, "gg_gat_dz_f32: null pointer");
    gat_dz_kernel<<<(int)ceil_div(n, kGmWarps), kGmThreads, 0, as_stream(stream)>>>(rowptr, nbr, a_tgt, a_src, alpha,
                                                                                    dalpha, n, slope, dz, da_tgt);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_gat_csc_gather_f32(const int32_t* rowptr_t, const int32_t* slot_map, const float* alpha, const float* dz,
                          int64_t n, float* alpha_t, float* da_src, gg_stream_t stream) {
    GG_REQUIRE(n >= 0, "gg_gat_csc_gather_f32: negative size");
    if (n == 0) return GG_OK;
    GG_REQUIRE(rowptr_t && alpha && dz && alpha_t && da_src, "gg_gat_csc_gather_f32: null pointer");
    gat_csc_gather_kernel<<<(int)ceil_div(n, kGmWarps), kGmThreads, 0, as_stream(stream)>>>(rowptr_t, slot_map, alpha,
                                                                                            dz, n, alpha_t, da_src);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_gat_sddmm_mp_f32(const int32_t* rowptr, const int32_t* nbr, const int32_t* item_row,
                        const int32_t* item_slot, int64_t items, const float* h, int64_t ldh, const float* g,
                        int64_t ldg, int64_t n, int64_t f, float* dalpha, int32_t* counter, gg_stream_t stream) {
    GG_REQUIRE(n >= 0 && f >= 0 && items >= 0, "gg_gat_sddmm_mp_f32: negative size");
    if (n == 0 || f == 0 || items == 0) return GG_OK;
    GG_REQUIRE(rowptr && item_row && item_slot && h && g && dalpha && counter, "gg_gat_sddmm_mp_f32: null pointer");
    if (!((f % 4 == 0) && f <= 1024 && ldh % 4 == 0 && ldg % 4 == 0 && gm_al16(h) && gm_al16(g))) {
        set_error("gg_gat_sddmm_mp_f32: needs f %% 4 == 0, f <= 1024 and 16-byte aligned rows");
        return GG_ERR_UNSUPPORTED;
    }
    cudaStream_t st = as_stream(stream);
    GG_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
    SddmmArgs a{rowptr, nbr, item_row, item_slot, (int)items, h, ldh, g, ldg, n, (int)f, counter, dalpha};
    const int nvec = (int)(f / 4);
    int per_sm = nvec <= 32 ? 4 : 2;
    int grid = (int)ceil_div(items, kGmWarps);
    if (grid > kNumSMs * per_sm) grid = kNumSMs * per_sm;
    if (nvec <= 32) gat_sddmm_mp_kernel<1><<<grid, kGmThreads, 0, st>>>(a);
    else if (nvec <= 64) gat_sddmm_mp_kernel<2><<<grid, kGmThreads, 0, st>>>(a);
    else if (nvec <= 128) gat_sddmm_mp_kernel<4><<<grid, kGmThreads, 0, st>>>(a);
    else gat_sddmm_mp_kernel<8><<<grid, kGmThreads, 0, st>>>(a);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_gat_sddmm_mpg_f32(const int32_t* rowptr, const int32_t* nbr, const int32_t* item_row,
                         const int32_t* item_slot, int64_t items, const float* h, int64_t ldh, const float* g,
                         int64_t ldg, int64_t n, int64_t f, float* dalpha, int32_t* counter, gg_stream_t stream) {
    GG_REQUIRE(n >= 0 && f >= 0 && items >= 0, "gg_gat_sddmm_mpg_f32: negative size");
    if (n == 0 || f == 0 || items == 0) return GG_OK;
    GG_REQUIRE(rowptr && item_row && item_slot && h && g && dalpha && counter, "gg_gat_sddmm_mpg_f32: null pointer");
    if (!((f % 4 == 0) && f <= 128 && ldh % 4 == 0 && ldg % 4 == 0 && gm_al16(h) && gm_al16(g) &&
          ldh < ((int64_t)1 << 30) && ldg < ((int64_t)1 << 30))) {
        set_error("gg_gat_sddmm_mpg_f32: needs f %% 4 == 0, f <= 128 and 16-byte aligned rows");
        return GG_ERR_UNSUPPORTED;
    }
    cudaStream_t st = as_stream(stream);
    GG_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
    SddmmArgs a{rowptr, nbr, item_row, item_slot, (int)items, h, ldh, g, ldg, n, (int)f, counter, dalpha};
    const int nvec = (int)(f / 4);
    int grid = (int)ceil_div(items, kGmWarps);
    if (grid > kNumSMs * 4) grid = kNumSMs * 4;
    if (nvec <= 4) gat_sddmm_mpg_kernel<4><<<grid, kGmThreads, 0, st>>>(a);
    else if (nvec <= 8) gat_sddmm_mpg_kernel<8><<<grid, kGmThreads, 0, st>>>(a);
    else if (nvec <= 16) gat_sddmm_mpg_kernel<16><<<grid, kGmThreads, 0, st>>>(a);
    else gat_sddmm_mpg_kernel<32><<<grid, kGmThreads, 0, st>>>(a);
    GG_LAUNCHED();
    return GG_OK;
}

}  // extern "C"

// ---- generic per-row softmax over given logits (Tfg scaled dot-product attention, ref: TfgIDLayer.py:336-345,
// sparse_adj.py:136-151 -> tf_geometric segment_softmax: exp(z - max) / sum, no epsilon) -------------------------------
namespace gg {

// one warp per row: alpha[s] = exp(scale * z[s] - max) / sum
__global__ void __launch_bounds__(kGmThreads)
    segment_softmax_kernel(const int32_t* __restrict__ rowptr, const float* __restrict__ z, int64_t n, float scale,
                           float* __restrict__ alpha) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kGmWarps + (threadIdx.x >> 5);
    if (row >= n) return;
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    float m = -CUDART_INF_F;
    for (int s = beg + lane; s < end; s += 32) m = fmaxf(m, scale * __ldg(z + s));
    m = gm_warp_max(m);
    float sum = 0.f;
    for (int s = beg + lane; s < end; s += 32) sum += expf(scale * __ldg(z + s) - m);
    const float inv = 1.0f / gm_warp_sum(sum);
    for (int s = beg + lane; s < end; s += 32) alpha[s] = expf(scale * __ldg(z + s) - m) * inv;
}

// dz[s] = scale * alpha[s] * (dalpha[s] - sum_row alpha dalpha)
__global__ void __launch_bounds__(kGmThreads)
    segment_softmax_bwd_kernel(const int32_t* __restrict__ rowptr, const float* __restrict__ alpha,
                               const float* __restrict__ dalpha, int64_t n, float scale, float* __restrict__ dz) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kGmWarps + (threadIdx.x >> 5);
    if (row >= n) return;
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    float d = 0.f;
    for (int s = beg + lane; s < end; s += 32) d = fmaf(__ldg(alpha + s), __ldg(dalpha + s), d);
    d = gm_warp_sum(d);
    for (int s = beg + lane; s < end; s += 32) dz[s] = scale * __ldg(alpha + s) * (__ldg(dalpha + s) - d);
}

}  // namespace gg

extern "C" {

int gg_segment_softmax_f32(const int32_t* rowptr, const float* z, int64_t n, float scale, float* alpha,
                           gg_stream_t stream) {
    GG_REQUIRE(n >= 0, "gg_segment_softmax_f32: negative size");
    if (n == 0) return GG_OK;
    GG_REQUIRE(rowptr && z && alpha, "gg_segment_softmax_f32: null pointer");
    gg::segment_softmax_kernel<<<(int)gg::ceil_div(n, gg::kGmWarps), gg::kGmThreads, 0, gg::as_stream(stream)>>>(rowptr, z, n, scale,
                                                                                                                alpha);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_segment_softmax_bwd_f32(const int32_t* rowptr, const float* alpha, const float* dalpha, int64_t n, float scale,
                               float* dz, gg_stream_t stream) {
    GG_REQUIRE(n >= 0, "gg_segment_softmax_bwd_f32: negative size");
    if (n == 0) return GG_OK;
    GG_REQUIRE(rowptr && alpha && dalpha && dz, "gg_segment_softmax_bwd_f32: null pointer");
    gg::segment_softmax_bwd_kernel<<<(int)gg::ceil_div(n, gg::kGmWarps), gg::kGmThreads, 0, gg::as_stream(stream)>>>(
        rowptr, alpha, dalpha, n, scale, dz);
    GG_LAUNCHED();
    return GG_OK;
}

}  // extern "C"
