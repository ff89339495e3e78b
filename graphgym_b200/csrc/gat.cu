// GAT: edge-softmax fused into the aggregation.
//
// Reference op sequence (ref: idconv.py:317-342, PyG softmax): materialise x_i, x_j and their
// concat per edge ([E,2C]), dot with att, leaky_relu, scatter_max, exp, scatter_add, divide, scale the
// [E,C] messages, scatter_add.  Here:
//   gat_scores   a_tgt[n,h] = <att[h,:C], H[n,h]>,  a_src[n,h] = <att[h,C:], H[n,h]>  (two dots per node,
//                so the per-edge logit is one add: z = a_tgt[i] + a_src[j])
//   gat_fwd      one warp per target row: segment max / sum of exp with warp shuffles over the row's
//                slots (4-byte a_src gathers), alpha written once for the backward, then the same
//                128-bit feature-row gather loop as the SpMM with alpha as the weight.
//   gat_bwd_edge one warp per target row: d_alpha = <g_i, H_j> per slot (the SDDMM, same gather
//                traffic as the forward), softmax + leaky_relu backward using
//                sum_e alpha_e d_alpha_e = <g_i, out_i - bias>, writes dz per slot and da_tgt per node
//   gat_bwd_src  one warp per source row of the CSC layout: dH_j = sum alpha_e g_i (weights fetched
//                through the CSC->CSR slot map) + da_src[j] att_src + da_tgt[j] att_tgt
//   gat_att_grad d att = [da_tgt^T H | da_src^T H] per head, fixed-order two-stage reduction
// HBM-bound: per pass E'*F*4 gathered + N*F*4 written + E'*(4 idx + 4 alpha) (+ logits 2*N*H*4).
#include <math_constants.h>

#include "common.cuh"

namespace gg {

constexpr int kGatWarps = 8;
constexpr int kGatThreads = kGatWarps * 32;
constexpr int kGatUnroll = 4;
constexpr int kMaxHeads = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float leaky(float z, float slope) { return z > 0.f ? z : slope * z; }
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}

struct GatArgs {
    const int32_t* rowptr;
    const int32_t* nbr;
    const float* h;
    int64_t ldh;
    const float* a_tgt;
    const float* a_src;
    int64_t n;
    int heads;
    int c;  // channels per head; f = heads * c, f % 4 == 0 and c % 4 == 0
    float slope;
    const float* bias;
    float* alpha;
    float* out;
    int64_t ldo;
    // backward
    const float* g;
    int64_t ldg;
    float* dz;
    float* da_tgt;
    const int32_t* map;  // CSC slot -> CSR slot
    const float* att;    // [heads, 2c]
    float* da_src;
    float* dh;
    int64_t lddh;
};

// a_tgt / a_src: one warp per node
__global__ void __launch_bounds__(kGatThreads) gat_scores_kernel(GatArgs a, float* a_tgt, float* a_src) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kGatWarps + (threadIdx.x >> 5);
    if (row >= a.n) return;
    const float4* hr = reinterpret_cast<const float4*>(a.h + row * a.ldh);
    const int cv = a.c >> 2;
    for (int hd = 0; hd < a.heads; ++hd) {
        const float4* at = reinterpret_cast<const float4*>(a.att + (int64_t)hd * 2 * a.c);
        float st = 0.f, ss = 0.f;
        for (int v = lane; v < cv; v += 32) {
            float4 x = __ldg(hr + hd * cv + v);
            st += dot4(x, __ldg(at + v));
            ss += dot4(x, __ldg(at + cv + v));
        }
        st = warp_sum(st);
        ss = warp_sum(ss);
        if (lane == 0) {
            a_tgt[row * a.heads + hd] = st;
            a_src[row * a.heads + hd] = ss;
        }
    }
}

template <int VPL, bool ONE_HEAD>
__global__ void __launch_bounds__(kGatThreads) gat_fwd_kernel(GatArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kGatWarps + (threadIdx.x >> 5);
    if (row >= a.n) return;
    const int beg = __ldg(a.rowptr + row), end = __ldg(a.rowptr + row + 1);
    const int H = ONE_HEAD ? 1 : a.heads;
    const int nvec = (H * a.c) >> 2;

    // ---- segment softmax: max, sum of exp, alpha (PyG softmax: exp(z - max) / (sum + 1e-16)) ----
    for (int hd = 0; hd < H; ++hd) {
        const float at = __ldg(a.a_tgt + row * H + hd);
        float m = -CUDART_INF_F;
        for (int s = beg + lane; s < end; s += 32)
            m = fmaxf(m, leaky(at + __ldg(a.a_src + (int64_t)__ldg(a.nbr + s) * H + hd), a.slope));
        m = warp_max(m);
        float sum = 0.f;
        for (int s = beg + lane; s < end; s += 32)
            sum += expf(leaky(at + __ldg(a.a_src + (int64_t)__ldg(a.nbr + s) * H + hd), a.slope) - m);
        sum = warp_sum(sum);
        const float inv = 1.0f / (sum + 1e-16f);
        for (int s = beg + lane; s < end; s += 32)
            a.alpha[(int64_t)s * H + hd] =
                expf(leaky(at + __ldg(a.a_src + (int64_t)__ldg(a.nbr + s) * H + hd), a.slope) - m) * inv;
    }
    __syncwarp();

    // ---- aggregation with alpha as the per-slot weight ----
    const float4* __restrict__ x4 = reinterpret_cast<const float4*>(a.h);
    const int64_t ld4 = a.ldh >> 2;
    bool act[VPL];
    int hq[VPL];
    float4 acc[VPL];
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
        act[q] = lane + q * 32 < nvec;
        hq[q] = ONE_HEAD ? 0 : ((lane + q * 32) * 4) / a.c;
        acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int base = beg; base < end; base += 32) {
        const int mine = base + lane;
        const int cidx = mine < end ? __ldg(a.nbr + mine) : 0;
        float wv = 0.f;
        if (ONE_HEAD) wv = mine < end ? a.alpha[mine] : 0.f;  // written by this warp above
        const int cnt = min(32, end - base);
        for (int k = 0; k < cnt; k += kGatUnroll) {
            float4 v[kGatUnroll][VPL];
            float ww[kGatUnroll][VPL];
#pragma unroll
            for (int u = 0; u < kGatUnroll; ++u) {
                const int sl = k + u;
                const int j = __shfl_sync(0xffffffffu, cidx, sl & 31);
                const float w1 = ONE_HEAD ? __shfl_sync(0xffffffffu, wv, sl & 31) : 0.f;
                const bool ok = sl < cnt;
                const float4* p = x4 + (int64_t)j * ld4 + lane;
#pragma unroll
                for (int q = 0; q < VPL; ++q) {
                    const bool on = ok && act[q];
                    v[u][q] = on ? ldg_nc_f4(p + q * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
                    ww[u][q] = ONE_HEAD ? w1 : (on ? a.alpha[(int64_t)(base + sl) * H + hq[q]] : 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < kGatUnroll; ++u)
#pragma unroll
                for (int q = 0; q < VPL; ++q) fma4(acc[q], ww[u][q], v[u][q]);
        }
    }
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
        if (!act[q]) continue;
        const int vi = lane + q * 32;
        float4 r = acc[q];
        if (a.bias) add4(r, __ldg(reinterpret_cast<const float4*>(a.bias) + vi));
        reinterpret_cast<float4*>(a.out + row * a.ldo)[vi] = r;
    }
}

template <int VPL, bool ONE_HEAD>
__global__ void __launch_bounds__(kGatThreads) gat_bwd_edge_kernel(GatArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kGatWarps + (threadIdx.x >> 5);
    if (row >= a.n) return;
    const int beg = __ldg(a.rowptr + row), end = __ldg(a.rowptr + row + 1);
    const int H = ONE_HEAD ? 1 : a.heads;
    const int nvec = (H * a.c) >> 2;
    const float4* __restrict__ x4 = reinterpret_cast<const float4*>(a.h);
    const int64_t ld4 = a.ldh >> 2;

    bool act[VPL];
    int hq[VPL];
    float4 gi[VPL];
    float D[kMaxHeads];
    float at[kMaxHeads];
    float da[kMaxHeads];
#pragma unroll
    for (int hd = 0; hd < kMaxHeads; ++hd) D[hd] = 0.f, da[hd] = 0.f, at[hd] = 0.f;
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
        const int vi = lane + q * 32;
        act[q] = vi < nvec;
        hq[q] = ONE_HEAD ? 0 : (vi * 4) / a.c;
        gi[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (act[q]) {
            gi[q] = __ldg(reinterpret_cast<const float4*>(a.g + row * a.ldg) + vi);
            float4 o = __ldg(reinterpret_cast<const float4*>(a.out + row * a.ldo) + vi);
            if (a.bias) {
                float4 b = __ldg(reinterpret_cast<const float4*>(a.bias) + vi);
                o.x -= b.x; o.y -= b.y; o.z -= b.z; o.w -= b.w;
            }
            const float d = dot4(gi[q], o);  // sum_e alpha_e d_alpha_e = <g_i, out_i - bias>
#pragma unroll
            for (int hd = 0; hd < kMaxHeads; ++hd)
                if (hd < H && hq[q] == hd) D[hd] += d;
        }
    }
#pragma unroll
    for (int hd = 0; hd < kMaxHeads; ++hd)
        if (hd < H) {
            D[hd] = warp_sum(D[hd]);
            at[hd] = __ldg(a.a_tgt + row * H + hd);
        }

    for (int base = beg; base < end; base += 32) {
        const int mine = base + lane;
        const int cidx = mine < end ? __ldg(a.nbr + mine) : 0;
        const int cnt = min(32, end - base);
        for (int k = 0; k < cnt; k += kGatUnroll) {
            float4 v[kGatUnroll][VPL];
            int jj[kGatUnroll];
#pragma unroll
            for (int u = 0; u < kGatUnroll; ++u) {
                const int sl = k + u;
                jj[u] = __shfl_sync(0xffffffffu, cidx, sl & 31);
                const bool ok = sl < cnt;
                const float4* p = x4 + (int64_t)jj[u] * ld4 + lane;
#pragma unroll
                for (int q = 0; q < VPL; ++q)
                    v[u][q] = (ok && act[q]) ? ldg_nc_f4(p + q * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < kGatUnroll; ++u) {
                const int sl = k + u;
                if (sl >= cnt) break;  // warp-uniform
                const int64_t s = base + sl;
#pragma unroll
                for (int hd = 0; hd < kMaxHeads; ++hd) {
                    if (hd >= H) break;
                    float d = 0.f;
#pragma unroll
                    for (int q = 0; q < VPL; ++q)
                        if (ONE_HEAD || hq[q] == hd) d += dot4(gi[q], v[u][q]);
                    d = warp_sum(d);  // d_alpha of this slot and head
                    const float al = a.alpha[s * H + hd];
                    const float z = at[hd] + __ldg(a.a_src + (int64_t)jj[u] * H + hd);
                    const float dzv = al * (d - D[hd]) * (z > 0.f ? 1.f : a.slope);
                    if (lane == 0) a.dz[s * H + hd] = dzv;
                    da[hd] += dzv;
                }
            }
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int hd = 0; hd < kMaxHeads; ++hd)
            if (hd < H) a.da_tgt[row * H + hd] = da[hd];
    }
}

// CSC rows: dH_j = sum_t alpha[map[t]] g[nbr_t] + da_src[j] * att_src + da_tgt[j] * att_tgt
template <int VPL, bool ONE_HEAD>
__global__ void __launch_bounds__(kGatThreads) gat_bwd_src_kernel(GatArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kGatWarps + (threadIdx.x >> 5);
    if (row >= a.n) return;
    const int beg = __ldg(a.rowptr + row), end = __ldg(a.rowptr + row + 1);
    const int H = ONE_HEAD ? 1 : a.heads;
    const int nvec = (H * a.c) >> 2;
    const float4* __restrict__ g4 = reinterpret_cast<const float4*>(a.g);
    const int64_t ld4 = a.ldg >> 2;

    bool act[VPL];
    int hq[VPL];
    float4 acc[VPL];
    float ds[kMaxHeads];
#pragma unroll
    for (int hd = 0; hd < kMaxHeads; ++hd) ds[hd] = 0.f;
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
        act[q] = lane + q * 32 < nvec;
        hq[q] = ONE_HEAD ? 0 : ((lane + q * 32) * 4) / a.c;
        acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int base = beg; base < end; base += 32) {
        const int mine = base + lane;
        const bool in = mine < end;
        const int cidx = in ? __ldg(a.nbr + mine) : 0;
        const int mp = in ? __ldg(a.map + mine) : 0;
        float wv = 0.f;
        if (ONE_HEAD) {
            wv = in ? __ldg(a.alpha + mp) : 0.f;
            ds[0] += in ? __ldg(a.dz + mp) : 0.f;
        } else {
#pragma unroll
            for (int hd = 0; hd < kMaxHeads; ++hd)
                if (hd < H && in) ds[hd] += __ldg(a.dz + (int64_t)mp * H + hd);
        }
        const int cnt = min(32, end - base);
        for (int k = 0; k < cnt; k += kGatUnroll) {
            float4 v[kGatUnroll][VPL];
            float ww[kGatUnroll][VPL];
#pragma unroll
            for (int u = 0; u < kGatUnroll; ++u) {
                const int sl = k + u;
                const int j = __shfl_sync(0xffffffffu, cidx, sl & 31);
                const int m = __shfl_sync(0xffffffffu, mp, sl & 31);
                const float w1 = ONE_HEAD ? __shfl_sync(0xffffffffu, wv, sl & 31) : 0.f;
                const bool ok = sl < cnt;
                const float4* p = g4 + (int64_t)j * ld4 + lane;
#pragma unroll
                for (int q = 0; q < VPL; ++q) {
                    const bool on = ok && act[q];
                    v[u][q] = on ? ldg_nc_f4(p + q * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
                    ww[u][q] = ONE_HEAD ? w1 : (on ? __ldg(a.alpha + (int64_t)m * H + hq[q]) : 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < kGatUnroll; ++u)
#pragma unroll
                for (int q = 0; q < VPL; ++q) fma4(acc[q], ww[u][q], v[u][q]);
        }
    }
#pragma unroll
    for (int hd = 0; hd < kMaxHeads; ++hd)
        if (hd < H) ds[hd] = warp_sum(ds[hd]);
    if (lane == 0) {
#pragma unroll
        for (int hd = 0; hd < kMaxHeads; ++hd)
            if (hd < H) a.da_src[row * H + hd] = ds[hd];
    }
    const int cv = a.c >> 2;
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
        if (!act[q]) continue;
        const int vi = lane + q * 32;
        const int hd = hq[q];
        const int cin = vi - hd * cv;  // float4 index inside the head
        const float4* at = reinterpret_cast<const float4*>(a.att + (int64_t)hd * 2 * a.c);
        float4 r = acc[q];
        float dsv = 0.f;
#pragma unroll
        for (int x = 0; x < kMaxHeads; ++x)
            if (x == hd) dsv = ds[x];
        fma4(r, dsv, __ldg(at + cv + cin));
        fma4(r, __ldg(a.da_tgt + row * H + hd), __ldg(at + cin));
        reinterpret_cast<float4*>(a.dh + row * a.lddh)[vi] = r;
    }
}

// partial[b][0][f] = sum_rows da_tgt[r,h(f)] H[r,f];  partial[b][1][f] likewise with da_src
__global__ void __launch_bounds__(256)
    gat_att_partial_kernel(const float* __restrict__ h, int64_t ldh, const float* __restrict__ da_tgt,
                           const float* __restrict__ da_src, int64_t n, int heads, int c, int64_t chunk,
                           float* __restrict__ part) {
    const int f = heads * c;
    int64_t r_beg = (int64_t)blockIdx.x * chunk;
    int64_t r_end = r_beg + chunk < n ? r_beg + chunk : n;
    for (int col = threadIdx.x; col < f; col += blockDim.x) {
        const int hd = col / c;
        // four rows in flight per thread (a single dependent load + fma chain ran at 2.7 TB/s)
        float st[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
        int64_t r = r_beg;
        for (; r + 4 <= r_end; r += 4) {
            float x[4], dt[4], dsv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                x[u] = __ldg(h + (r + u) * ldh + col);
                dt[u] = __ldg(da_tgt + (r + u) * heads + hd);
                dsv[u] = __ldg(da_src + (r + u) * heads + hd);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                st[u] = fmaf(dt[u], x[u], st[u]);
                ss[u] = fmaf(dsv[u], x[u], ss[u]);
            }
        }
        for (; r < r_end; ++r) {
            const float x = __ldg(h + r * ldh + col);
            st[0] = fmaf(__ldg(da_tgt + r * heads + hd), x, st[0]);
            ss[0] = fmaf(__ldg(da_src + r * heads + hd), x, ss[0]);
        }
        part[((int64_t)blockIdx.x * 2 + 0) * f + col] = (st[0] + st[1]) + (st[2] + st[3]);
        part[((int64_t)blockIdx.x * 2 + 1) * f + col] = (ss[0] + ss[1]) + (ss[2] + ss[3]);
    }
}
// one warp per output element: lanes stride over the blocks' partials (fixed order: deterministic)
__global__ void __launch_bounds__(256)
    gat_att_final_kernel(const float* __restrict__ part, int64_t blocks, int heads, int c,
                         float* __restrict__ datt) {
    const int f = heads * c;
    const int lane = threadIdx.x & 31;
    for (int e = blockIdx.x * 8 + (threadIdx.x >> 5); e < 2 * f; e += gridDim.x * 8) {
        const int which = e / f, col = e % f;
        float s = 0.f;
        for (int64_t b = lane; b < blocks; b += 32) s += part[(b * 2 + which) * f + col];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const int hd = col / c, ci = col % c;
        if (lane == 0) datt[(int64_t)hd * 2 * c + which * c + ci] = s;  // [h, 0:c] = target half, [h, c:2c] = source half
    }
}

static int64_t att_blocks(int64_t n) {
    int64_t b = ceil_div(n, 128);
    int64_t cap = (int64_t)kNumSMs * 8;
    if (b > cap) b = cap;
    return b < 1 ? 1 : b;
}

static int check_shape(const char* who, int64_t n, int heads, int c) {
    if (n < 0 || heads < 1 || heads > kMaxHeads || c < 4 || (c % 4) != 0 || (int64_t)heads * c > 1024) {
        set_error("%s: unsupported shape n=%lld heads=%d c=%d (need 1<=heads<=%d, c%%4==0, heads*c<=1024)",
                  who, (long long)n, heads, c, kMaxHeads);
        return GG_ERR_UNSUPPORTED;
    }
    return GG_OK;
}

template <template <int, bool> class K>
struct Dispatch;

#define GG_GAT_DISPATCH(KERNEL, args, nvec, one_head, grid, st)                                   \
    do {                                                                                          \
        if ((nvec) <= 32) {                                                                       \
            if (one_head) KERNEL<1, true><<<grid, kGatThreads, 0, st>>>(args);                    \
            else KERNEL<1, false><<<grid, kGatThreads, 0, st>>>(args);                            \
        } else if ((nvec) <= 64) {                                                                \
            if (one_head) KERNEL<2, true><<<grid, kGatThreads, 0, st>>>(args);                    \
            else KERNEL<2, false><<<grid, kGatThreads, 0, st>>>(args);                            \
        } else if ((nvec) <= 128) {                                                               \
            if (one_head) KERNEL<4, true><<<grid, kGatThreads, 0, st>>>(args);                    \
            else KERNEL<4, false><<<grid, kGatThreads, 0, st>>>(args);                            \
        } else {                                                                                  \
            if (one_head) KERNEL<8, true><<<grid, kGatThreads, 0, st>>>(args);                    \
            else KERNEL<8, false><<<grid, kGatThreads, 0, st>>>(args);                            \
        }                                                                                         \
    } while (0)

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace gg

using namespace gg;

extern "C" {

int gg_gat_scores_f32(const float* h, int64_t ldh, const float* att, int64_t n, int heads, int c,
                      float* a_tgt, float* a_src, gg_stream_t stream) {
    int rc = check_shape("gg_gat_scores_f32", n, heads, c);
    if (rc != GG_OK) return rc;
    if (n == 0) return GG_OK;
    GG_REQUIRE(h && att && a_tgt && a_src && ldh >= (int64_t)heads * c && ldh % 4 == 0 && al16(h) && al16(att),
               "gg_gat_scores_f32: bad operands");
    GatArgs a{};
    a.h = h; a.ldh = ldh; a.att = att; a.n = n; a.heads = heads; a.c = c;
    gat_scores_kernel<<<(int)ceil_div(n, kGatWarps), kGatThreads, 0, as_stream(stream)>>>(a, a_tgt, a_src);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_gat_fwd_f32(const int32_t* rowptr, const int32_t* nbr, const float* h, int64_t ldh,
                   const float* a_tgt, const float* a_src, int64_t n, int heads, int c, float slope,
                   const float* bias, float* alpha, float* out, int64_t ldo, gg_stream_t stream) {
    int rc = check_shape("gg_gat_fwd_f32", n, heads, c);
    if (rc != GG_OK) return rc;
    if (n == 0) return GG_OK;
    const int64_t f = (int64_t)heads * c;
    GG_REQUIRE(rowptr && h && a_tgt && a_src && alpha && out, "gg_gat_fwd_f32: null pointer");
    GG_REQUIRE(ldh >= f && ldo >= f && ldh % 4 == 0 && ldo % 4 == 0 && al16(h) && al16(out) &&
                   (!bias || al16(bias)), "gg_gat_fwd_f32: rows must be 16-byte aligned");
    GatArgs a{};
    a.rowptr = rowptr; a.nbr = nbr; a.h = h; a.ldh = ldh; a.a_tgt = a_tgt; a.a_src = a_src; a.n = n;
    a.heads = heads; a.c = c; a.slope = slope; a.bias = bias; a.alpha = alpha; a.out = out; a.ldo = ldo;
    int grid = (int)ceil_div(n, kGatWarps);
    cudaStream_t st = as_stream(stream);
    GG_GAT_DISPATCH(gat_fwd_kernel, a, f / 4, heads == 1, grid, st);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_gat_bwd_edge_f32(const int32_t* rowptr, const int32_t* nbr, const float* h, int64_t ldh,
                        const float* a_tgt, const float* a_src, const float* alpha, const float* g,
                        int64_t ldg, const float* out, int64_t ldo, const float* bias, int64_t n,
                        int heads, int c, float slope, float* dz, float* da_tgt, gg_stream_t stream) {
    int rc = check_shape("gg_gat_bwd_edge_f32", n, heads, c);
    if (rc != GG_OK) return rc;
    if (n == 0) return GG_OK;
    const int64_t f = (int64_t)heads * c;
    GG_REQUIRE(rowptr && h && a_tgt && a_src && alpha && g && out && dz && da_tgt,
               "gg_gat_bwd_edge_f32: null pointer");
    GG_REQUIRE(ldh >= f && ldg >= f && ldo >= f && ldh % 4 == 0 && ldg % 4 == 0 && ldo % 4 == 0 &&
                   al16(h) && al16(g) && al16(out) && (!bias || al16(bias)),
               "gg_gat_bwd_edge_f32: rows must be 16-byte aligned");
    GatArgs a{};
    a.rowptr = rowptr; a.nbr = nbr; a.h = h; a.ldh = ldh; a.a_tgt = a_tgt; a.a_src = a_src; a.n = n;
    a.heads = heads; a.c = c; a.slope = slope; a.bias = bias; a.alpha = const_cast<float*>(alpha);
    a.out = const_cast<float*>(out); a.ldo = ldo; a.g = g; a.ldg = ldg; a.dz = dz; a.da_tgt = da_tgt;
    int grid = (int)ceil_div(n, kGatWarps);
    cudaStream_t st = as_stream(stream);
    GG_GAT_DISPATCH(gat_bwd_edge_kernel, a, f / 4, heads == 1, grid, st);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_gat_bwd_src_f32(const int32_t* rowptr_t, const int32_t* nbr_t, const int32_t* slot_map,
                       const float* alpha, const float* dz, const float* g, int64_t ldg,
                       const float* da_tgt, const float* att, int64_t n, int heads, int c,
                       float* da_src, float* dh, int64_t lddh, gg_stream_t stream) {
    int rc = check_shape("gg_gat_bwd_src_f32", n, heads, c);
    if (rc != GG_OK) return rc;
    if (n == 0) return GG_OK;
    const int64_t f = (int64_t)heads * c;
    GG_REQUIRE(rowptr_t && alpha && dz && g && da_tgt && att && da_src && dh,
               "gg_gat_bwd_src_f32: null pointer");
    GG_REQUIRE(ldg >= f && lddh >= f && ldg % 4 == 0 && lddh % 4 == 0 && al16(g) && al16(dh) && al16(att),
               "gg_gat_bwd_src_f32: rows must be 16-byte aligned");
    GatArgs a{};
    a.rowptr = rowptr_t; a.nbr = nbr_t; a.map = slot_map; a.alpha = const_cast<float*>(alpha);
    a.dz = const_cast<float*>(dz); a.g = g; a.ldg = ldg; a.da_tgt = const_cast<float*>(da_tgt);
    a.att = att; a.n = n; a.heads = heads; a.c = c; a.da_src = da_src; a.dh = dh; a.lddh = lddh;
    int grid = (int)ceil_div(n, kGatWarps);
    cudaStream_t st = as_stream(stream);
    GG_GAT_DISPATCH(gat_bwd_src_kernel, a, f / 4, heads == 1, grid, st);
    GG_LAUNCHED();
    return GG_OK;
}

size_t gg_gat_att_grad_workspace_bytes(int64_t n, int heads, int c) {
    return (size_t)att_blocks(n) * 2 * (size_t)heads * (size_t)c * sizeof(float) + 256;
}

int gg_gat_att_grad_f32(const float* h, int64_t ldh, const float* da_tgt, const float* da_src,
                        int64_t n, int heads, int c, float* datt, void* workspace,
                        size_t workspace_bytes, gg_stream_t stream) {
    int rc = check_shape("gg_gat_att_grad_f32", n, heads, c);
    if (rc != GG_OK) return rc;
    GG_REQUIRE(datt, "gg_gat_att_grad_f32: null output");
    cudaStream_t st = as_stream(stream);
    const int64_t f = (int64_t)heads * c;
    if (n == 0) {
        GG_CUDA(cudaMemsetAsync(datt, 0, (size_t)2 * f * 4, st));
        return GG_OK;
    }
    GG_REQUIRE(h && da_tgt && da_src && workspace && ldh >= f, "gg_gat_att_grad_f32: bad operands");
    if (workspace_bytes < gg_gat_att_grad_workspace_bytes(n, heads, c)) {
        set_error("gg_gat_att_grad_f32: workspace too small");
        return GG_ERR_WORKSPACE;
    }
    int64_t blocks = att_blocks(n);
    int64_t chunk = ceil_div(n, blocks);
    blocks = ceil_div(n, chunk);
    float* part = static_cast<float*>(workspace);
    int threads = f >= 256 ? 256 : (int)(ceil_div(f, 32) * 32);
    gat_att_partial_kernel<<<(int)blocks, threads, 0, st>>>(h, ldh, da_tgt, da_src, n, heads, c, chunk, part);
    GG_LAUNCHED();
    gat_att_final_kernel<<<(int)ceil_div(2 * f, 8), 256, 0, st>>>(part, blocks, heads, c, datt);
    GG_LAUNCHED();
    return GG_OK;
}

}  // extern "C"
