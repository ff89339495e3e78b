// CSR aggregation:  out[i,:] = epi( sum_{s in seg i} w[s] * x[nbr[s],:] ).
//
// Replaces the reference's gather -> scale -> scatter-add sequence (ref: idconv.py:89-92,177-180,
// sparse_adj.py:91-97), which materialises the [E,F] message tensor and reduces it with atomics.
// Here one warp owns one output row: it streams the row's neighbour indices (coalesced 128 B per
// 32 slots), broadcasts them with shuffles and issues up to 8 independent 128-bit feature-row
// gathers per lane before accumulating, so a warp keeps up to 8 * F * 4 bytes in flight.  The sum
// order is the slot order => results are run-to-run deterministic (no atomics).
//
// HBM roofline: algorithmic bytes per launch = E'*F*4 (gathered rows) + N*F*4 (output)
//               + E'*4 (nbr) + (N+1)*4 (rowptr) [+ E'*4 weights].
#include "common.cuh"

namespace gg {

constexpr int kSpmmWarps = 8;
constexpr int kSpmmThreads = kSpmmWarps * 32;
// independent gathers in flight per lane-group; wide rows (VPL float4 per lane) unroll less so the
// staging registers stay <= 16 float4 per thread
template <int VPL> struct Unroll { static constexpr int value = VPL <= 2 ? 8 : (VPL == 4 ? 4 : 2); };

struct SpmmArgs {
    const int32_t* rowptr;
    const int32_t* nbr;
    const float* w;
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
};

// LPR lanes cover one feature row with VPL float4 each (VPL > 1 only when LPR == 32); a warp
// processes 32/LPR edges of the row at once and folds the groups with xor-shuffles at the end.
template <int LPR, int VPL, bool WEIGHTED>
__global__ void __launch_bounds__(kSpmmThreads) spmm_vec_kernel(SpmmArgs a) {
    constexpr int G = 32 / LPR;
    constexpr int kUnroll = Unroll<VPL>::value;
    const int lane = threadIdx.x & 31;
    const int grp = lane / LPR;
    const int vl = lane % LPR;
    const int nvec = a.f >> 2;
    const int64_t row = (int64_t)blockIdx.x * kSpmmWarps + (threadIdx.x >> 5);
    if (row >= a.n) return;

    const int beg = __ldg(a.rowptr + row), end = __ldg(a.rowptr + row + 1);
    const float4* __restrict__ x4 = reinterpret_cast<const float4*>(a.x);
    const int64_t ldx4 = a.ldx >> 2;

    bool act[VPL];
    float4 acc[VPL];
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
        act[q] = vl + q * LPR < nvec;
        acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    for (int base = beg; base < end; base += 32) {
        const int mine = base + lane;
        const int c = mine < end ? __ldg(a.nbr + mine) : 0;
        float wv = 1.f;
        if (WEIGHTED) wv = mine < end ? __ldg(a.w + mine) : 0.f;
        const int cnt = min(32, end - base);
        for (int k = 0; k < cnt; k += G * kUnroll) {
            float4 v[kUnroll][VPL];
            float ww[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int sl = k + u * G + grp;
                const int j = __shfl_sync(0xffffffffu, c, sl & 31);
                ww[u] = WEIGHTED ? __shfl_sync(0xffffffffu, wv, sl & 31) : 1.f;
                const bool ok = sl < cnt;
                const float4* p = x4 + (int64_t)j * ldx4 + vl;
#pragma unroll
                for (int q = 0; q < VPL; ++q) {
                    v[u][q] = (ok && act[q]) ? ldg_nc_f4(p + q * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
                for (int q = 0; q < VPL; ++q) {
                    if (WEIGHTED) fma4(acc[q], ww[u], v[u][q]);
                    else add4(acc[q], v[u][q]);
                }
            }
        }
    }
    if (G > 1) {
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) {
#pragma unroll
            for (int q = 0; q < VPL; ++q) {
                acc[q].x += __shfl_xor_sync(0xffffffffu, acc[q].x, o);
                acc[q].y += __shfl_xor_sync(0xffffffffu, acc[q].y, o);
                acc[q].z += __shfl_xor_sync(0xffffffffu, acc[q].z, o);
                acc[q].w += __shfl_xor_sync(0xffffffffu, acc[q].w, o);
            }
        }
    }
    if (grp != 0) return;
    const float inv = (a.reduce == GG_MEAN && end > beg) ? 1.0f / (float)(end - beg) : 1.0f;
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
        if (!act[q]) continue;
        const int vi = vl + q * LPR;
        float4 r = acc[q];
        if (a.reduce == GG_MEAN) {
            r.x *= inv; r.y *= inv; r.z *= inv; r.w *= inv;
        }
        if (a.x_self) {
            float4 s = __ldg(reinterpret_cast<const float4*>(a.x_self + row * a.ld_self) + vi);
            fma4(r, a.self_scale, s);
        }
        if (a.bias) {
            float4 b = __ldg(reinterpret_cast<const float4*>(a.bias) + vi);
            add4(r, b);
        }
        reinterpret_cast<float4*>(a.out + row * a.ldo)[vi] = r;
    }
}

// Any f / any alignment: lanes stride the features, slots are walked in order.
template <bool WEIGHTED>
__global__ void __launch_bounds__(kSpmmThreads) spmm_scalar_kernel(SpmmArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kSpmmWarps + (threadIdx.x >> 5);
    if (row >= a.n) return;
    const int beg = __ldg(a.rowptr + row), end = __ldg(a.rowptr + row + 1);
    const float inv = (a.reduce == GG_MEAN && end > beg) ? 1.0f / (float)(end - beg) : 1.0f;
    for (int f0 = lane; f0 < a.f; f0 += 32) {
        float acc = 0.f;
        for (int s = beg; s < end; ++s) {
            const int j = __ldg(a.nbr + s);
            const float xv = __ldg(a.x + (int64_t)j * a.ldx + f0);
            if (WEIGHTED) acc = fmaf(__ldg(a.w + s), xv, acc);
            else acc += xv;
        }
        if (a.reduce == GG_MEAN) acc *= inv;
        if (a.x_self) acc = fmaf(a.self_scale, __ldg(a.x_self + row * a.ld_self + f0), acc);
        if (a.bias) acc += __ldg(a.bias + f0);
        a.out[row * a.ldo + f0] = acc;
    }
}

template <int LPR, int VPL>
static void launch_vec(const SpmmArgs& a, cudaStream_t st) {
    int grid = (int)ceil_div(a.n, kSpmmWarps);
    if (a.w) spmm_vec_kernel<LPR, VPL, true><<<grid, kSpmmThreads, 0, st>>>(a);
    else spmm_vec_kernel<LPR, VPL, false><<<grid, kSpmmThreads, 0, st>>>(a);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace gg

using namespace gg;

extern "C" int gg_spmm_f32(const int32_t* rowptr, const int32_t* nbr, const float* w_slot,
                           const float* x, int64_t ldx, float* out, int64_t ldo, int64_t num_rows,
                           int64_t f, int reduce, const float* x_self, int64_t ld_self,
                           float self_scale, const float* bias, gg_stream_t stream) {
    GG_REQUIRE(num_rows >= 0 && f >= 0, "gg_spmm_f32: negative size");
    GG_REQUIRE(reduce == GG_SUM || reduce == GG_MEAN, "gg_spmm_f32: reduce=%d", reduce);
    if (num_rows == 0 || f == 0) return GG_OK;
    GG_REQUIRE(rowptr && x && out, "gg_spmm_f32: null pointer");
    GG_REQUIRE(ldx >= f && ldo >= f, "gg_spmm_f32: leading dimension smaller than f");
    GG_REQUIRE(!x_self || ld_self >= f, "gg_spmm_f32: ld_self smaller than f");
    GG_REQUIRE(num_rows < ((int64_t)1 << 31) && f < (1 << 20), "gg_spmm_f32: size out of range");
    cudaStream_t st = as_stream(stream);
    SpmmArgs a{rowptr, nbr, w_slot, x, ldx, out, ldo, num_rows, (int)f, reduce,
               x_self, ld_self, self_scale, bias};
    bool vec = (f % 4 == 0) && (ldx % 4 == 0) && (ldo % 4 == 0) && aligned16(x) && aligned16(out) &&
               (!x_self || (ld_self % 4 == 0 && aligned16(x_self))) && (!bias || aligned16(bias)) &&
               f <= 1024;
    if (vec) {
        int nvec = (int)(f / 4);
        if (nvec <= 4) launch_vec<4, 1>(a, st);
        else if (nvec <= 8) launch_vec<8, 1>(a, st);
        else if (nvec <= 16) launch_vec<16, 1>(a, st);
        else if (nvec <= 32) launch_vec<32, 1>(a, st);
        else if (nvec <= 64) launch_vec<32, 2>(a, st);
        else if (nvec <= 128) launch_vec<32, 4>(a, st);
        else launch_vec<32, 8>(a, st);
    } else {
        int grid = (int)ceil_div(num_rows, kSpmmWarps);
        if (w_slot) spmm_scalar_kernel<true><<<grid, kSpmmThreads, 0, st>>>(a);
        else spmm_scalar_kernel<false><<<grid, kSpmmThreads, 0, st>>>(a);
    }
    GG_LAUNCHED();
    return GG_OK;
}
