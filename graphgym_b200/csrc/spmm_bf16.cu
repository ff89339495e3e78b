// bf16-gather aggregation: the 1e-2 mode of the north star ("fp32, or 1e-2 (bf16)"; SURVEY §8a row 4,
// §8b `gg_spmm_*` fp32 + bf16).  The gathered operand X is stored in bf16 (half the bytes of the pass that
// binds the layer), every product and sum is fp32, the output is fp32:
//     out[i,:] = reduce_{s in row i} w[s] * float(X_bf16[nbr[s],:])  (+ bias)
// Same merge-path plan, same fixed summation order and the same warp layout as spmm_mpg_kernel (spmm_mp.cu):
// a 16-byte gather now carries 8 columns, so G = 4..32 lanes cover f = 32..256 columns and one warp instruction
// gathers 32/G different slots.  gg_cast_f32_bf16 is the round-to-nearest-even conversion of a feature matrix.
#include <cuda_bf16.h>

#include "common.cuh"

namespace gg {

constexpr int kHWarps = 8;
constexpr int kHThreads = kHWarps * 32;
constexpr int kHTile = 488;  // plan units <= 480, + row-pointer slack

struct HArgs {
    const int32_t* rowptr;
    const int32_t* nbr;
    const float* w;
    const int32_t* item_row;
    const int32_t* item_slot;
    int items;
    const uint16_t* x;  // bf16 bits
    int64_t ldx;        // elements
    float* out;
    int64_t ldo;
    int64_t n;
    int f;
    int reduce;
    const float* bias;
    const float* x_self;  // fp32 self term (GIN's (1 + eps) x_i), nullable
    int64_t ld_self;
    float self_scale;
    int* counter;
    float* carry;  // [items, f]
    float* head;   // [items, f]
    int l2_hint;
};

struct Acc8 {
    float v[8];
};
__device__ __forceinline__ void zero8(Acc8& a) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a.v[i] = 0.f;
}
// two bf16 per 32-bit word, element 0 in the low half
__device__ __forceinline__ void fma8(Acc8& a, float w, const uint4& r) {
    const uint32_t q[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a.v[2 * i] = fmaf(w, __uint_as_float(q[i] << 16), a.v[2 * i]);
        a.v[2 * i + 1] = fmaf(w, __uint_as_float(q[i] & 0xffff0000u), a.v[2 * i + 1]);
    }
}
template <int G>
__device__ __forceinline__ void reduce8(Acc8& a) {
#pragma unroll
    for (int m = G; m < 32; m <<= 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a.v[i] += __shfl_xor_sync(0xffffffffu, a.v[i], m);
    }
}
__device__ __forceinline__ uint4 ldg_nc_u4_hint(const void* p, uint64_t pol) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void store8(float* p, const Acc8& a) {
    reinterpret_cast<float4*>(p)[0] = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(a.v[4], a.v[5], a.v[6], a.v[7]);
}
__device__ __forceinline__ void load_add8(Acc8& a, const float* p) {
    const float4 x = reinterpret_cast<const float4*>(p)[0], y = reinterpret_cast<const float4*>(p)[1];
    a.v[0] += x.x; a.v[1] += x.y; a.v[2] += x.z; a.v[3] += x.w;
    a.v[4] += y.x; a.v[5] += y.y; a.v[6] += y.z; a.v[7] += y.w;
}

__device__ __forceinline__ void h_epilogue(const HArgs& a, int row, int gl, Acc8& r, int deg) {
    if (a.reduce == GG_MEAN) {
        const float inv = deg > 0 ? 1.0f / (float)deg : 1.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] *= inv;
    }
    if (a.x_self) {
        const float* p = a.x_self + (int64_t)row * a.ld_self + gl * 8;
        const float4 x = __ldg(reinterpret_cast<const float4*>(p)), y = __ldg(reinterpret_cast<const float4*>(p) + 1);
        r.v[0] = fmaf(a.self_scale, x.x, r.v[0]); r.v[1] = fmaf(a.self_scale, x.y, r.v[1]);
        r.v[2] = fmaf(a.self_scale, x.z, r.v[2]); r.v[3] = fmaf(a.self_scale, x.w, r.v[3]);
        r.v[4] = fmaf(a.self_scale, y.x, r.v[4]); r.v[5] = fmaf(a.self_scale, y.y, r.v[5]);
        r.v[6] = fmaf(a.self_scale, y.z, r.v[6]); r.v[7] = fmaf(a.self_scale, y.w, r.v[7]);
    }
    if (a.bias) load_add8(r, a.bias + gl * 8);
    store8(a.out + (int64_t)row * a.ldo + gl * 8, r);
}

// Batches of S x 4 slots: the four gathered 16-byte vectors of a lane are unpacked ONCE into 32 fp32 registers, so the
// no-row-end path is 32 plain FMAs and a row-end sweep re-uses the unpacked values (the first version kept 8 packed
// vectors and unpacked inside every sweep: 1.48 G instructions, issue-bound at 3.6 ms for F = 128 on the products
// graph — profiles/r01_spmm_bf16_ncu_raw.csv).
struct Row8 {
    float v[8];
};
__device__ __forceinline__ void unpack8(Row8& o, const uint4& r) {
    const uint32_t q[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        o.v[2 * i] = __uint_as_float(q[i] << 16);
        o.v[2 * i + 1] = __uint_as_float(q[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ void fma8f(Acc8& a, float w, const Row8& r) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a.v[i] = fmaf(w, r.v[i], a.v[i]);
}

template <int G, bool WEIGHTED>
__global__ void __launch_bounds__(kHThreads, 4) spmm_h_kernel(const __grid_constant__ HArgs a) {
    constexpr int S = 32 / G, U = 4, B = S * U;
    __shared__ __align__(16) int32_t s_nbr_all[kHWarps][kHTile];
    __shared__ __align__(16) float s_w_all[WEIGHTED ? kHWarps : 1][WEIGHTED ? kHTile : 4];
    __shared__ __align__(16) int32_t s_rp_all[kHWarps][kHTile];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int grp = lane / G, gl = lane % G;
    int32_t* s_nbr = s_nbr_all[wid];
    float* s_w = s_w_all[WEIGHTED ? wid : 0];
    int32_t* s_rp = s_rp_all[wid];
    const int nvec = a.f >> 3;  // 16-byte vectors of 8 bf16 per row
    const bool act = gl < nvec;
    const bool writer = act && grp == 0;
    const char* __restrict__ xg = reinterpret_cast<const char*>(a.x) + (act ? gl : nvec - 1) * 16;
    const uint32_t row_bytes = (uint32_t)a.ldx * 2u;
    const uint64_t pol_keep = l2_policy(a.l2_hint ? 1 : 0), pol_stream = l2_policy(a.l2_hint ? 2 : 0);
    auto gather = [&](int j) { return ldg_nc_u4_hint(xg + (uint64_t)(uint32_t)j * row_bytes, pol_keep); };

    int item = 0;
    if (lane == 0) item = atomicAdd(a.counter, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    while (item < a.items) {
        int next = 0;
        if (lane == 0) next = atomicAdd(a.counter, 1);
        const int r0 = __ldg(a.item_row + item), s0 = __ldg(a.item_slot + item);
        const int r1 = __ldg(a.item_row + item + 1), s1 = __ldg(a.item_slot + item + 1);
        __syncwarp();
        for (int q = s0 + lane; q < s1; q += 32) {
            s_nbr[q - s0] = ldg_nc_s32_hint(a.nbr + q, pol_stream);
            if (WEIGHTED) s_w[q - s0] = ldg_nc_f32_hint(a.w + q, pol_stream);
        }
        const int nrp = (r1 < a.n ? r1 + 1 : (int)a.n) - r0 + 1;
        for (int i = lane; i < nrp; i += 32) s_rp[i] = __ldg(a.rowptr + r0 + i);
        __syncwarp();

        const bool cont_first = s0 > s_rp[0];
        int cur = r0;
        int re = r0 < a.n ? s_rp[1] : 0x7fffffff;
        Acc8 acc;
        zero8(acc);
        auto finalize = [&](int row) {
            reduce8<G>(acc);
            if (row == r0 && cont_first) {
                if (writer) store8(a.head + (int64_t)item * a.f + gl * 8, acc);
            } else if (writer) {
                Acc8 r = acc;
                h_epilogue(a, row, gl, r, s_rp[row - r0 + 1] - s_rp[row - r0]);
            }
            zero8(acc);
        };
        for (int s = s0; s < s1; s += B) {
            const int g0 = s + grp * U;
            const int t = g0 - s0;  // multiple of 4: 16-byte aligned inside the tile
            const int e = s + B < s1 ? s + B : s1;
            uint4 raw[U];
            float wv[U];
            if (s + B <= s1) {
                const int4 i0 = *reinterpret_cast<const int4*>(s_nbr + t);
                raw[0] = gather(i0.x); raw[1] = gather(i0.y); raw[2] = gather(i0.z); raw[3] = gather(i0.w);
                if (WEIGHTED) {
                    const float4 w0 = *reinterpret_cast<const float4*>(s_w + t);
                    wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w;
                } else {
                    wv[0] = wv[1] = wv[2] = wv[3] = 1.f;
                }
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool ok = g0 + u < s1;
                    raw[u] = ok ? gather(s_nbr[t + u]) : make_uint4(0u, 0u, 0u, 0u);
                    wv[u] = ok ? (WEIGHTED ? s_w[t + u] : 1.f) : 0.f;
                }
            }
            Row8 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) unpack8(v[u], raw[u]);
            if (re >= e && !(re == e && cur < r1)) {
#pragma unroll
                for (int u = 0; u < U; ++u) fma8f(acc, wv[u], v[u]);
            } else {
                int pos = s;
                while (true) {
                    const int seg_end = re < e ? re : e;
                    const int lo = pos - g0, hi = seg_end - g0;
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (u >= lo && u < hi) fma8f(acc, wv[u], v[u]);
                    if (re > e || cur >= r1) break;
                    finalize(cur);
                    ++cur;
                    re = cur < a.n ? s_rp[cur - r0 + 1] : 0x7fffffff;
                    pos = seg_end;
                }
            }
        }
        while (cur < r1) {
            finalize(cur);
            ++cur;
        }
        reduce8<G>(acc);
        if (writer) store8(a.carry + (int64_t)item * a.f + gl * 8, acc);
        item = __shfl_sync(0xffffffffu, next, 0);
    }
}

// rows split over several items: carry[k0..f-1] + head[f] in item order, then the epilogue (as spmm_mp_fixup_kernel)
template <int G>
__global__ void __launch_bounds__(kHThreads) spmm_h_fixup_kernel(const __grid_constant__ HArgs a) {
    const int64_t idx = (int64_t)blockIdx.x * kHThreads + threadIdx.x;
    const int64_t fi = idx / G;
    const int gl = (int)(idx % G);
    if (fi < 1 || fi >= a.items) return;
    const int f_item = (int)fi;
    const int r0 = __ldg(a.item_row + f_item), s0 = __ldg(a.item_slot + f_item);
    const int r1 = __ldg(a.item_row + f_item + 1);
    if (r0 >= a.n || r0 >= r1) return;
    const int rb = __ldg(a.rowptr + r0);
    if (s0 <= rb) return;
    if (gl >= (a.f >> 3)) return;
    int k0 = f_item - 1;
    while (k0 >= 1 && __ldg(a.item_row + k0) == r0 && __ldg(a.item_slot + k0) > rb) --k0;
    Acc8 r;
    zero8(r);
    for (int k = k0; k < f_item; ++k) load_add8(r, a.carry + (int64_t)k * a.f + gl * 8);
    load_add8(r, a.head + (int64_t)f_item * a.f + gl * 8);
    h_epilogue(a, r0, gl, r, __ldg(a.rowptr + r0 + 1) - rb);
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src, int64_t ld, int64_t rows, int f,
                                                        uint16_t* __restrict__ dst, int64_t ld_dst) {
    const int nv = f >> 2;  // 4 floats -> 4 bf16 (8 bytes) per thread and iteration
    const int64_t total = rows * nv;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / nv;
        const int c = (int)(e - r * nv) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + r * ld + c));
        const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&lo);
        o.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(dst + r * ld_dst + c) = o;
    }
}

template <int G>
static void launch_h(const HArgs& a, cudaStream_t st) {
    int grid = (int)ceil_div(a.items, kHWarps);
    if (grid > kNumSMs * 4) grid = kNumSMs * 4;
    if (a.w) {
        static bool done = false;
        if (!done) {
            cudaFuncSetAttribute(spmm_h_kernel<G, true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);
            done = true;
        }
        spmm_h_kernel<G, true><<<grid, kHThreads, 0, st>>>(a);
    } else {
        static bool done = false;
        if (!done) {
            cudaFuncSetAttribute(spmm_h_kernel<G, false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);
            done = true;
        }
        spmm_h_kernel<G, false><<<grid, kHThreads, 0, st>>>(a);
    }
    count_launch();
    spmm_h_fixup_kernel<G><<<(int)ceil_div((int64_t)a.items * G, kHThreads), kHThreads, 0, st>>>(a);
    count_launch();
}

static inline bool h_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace gg

using namespace gg;

extern "C" {

int gg_cast_f32_bf16(const float* src, int64_t ld, int64_t rows, int64_t f, uint16_t* dst, int64_t ld_dst,
                     gg_stream_t stream) {
    GG_REQUIRE(rows >= 0 && f >= 0, "gg_cast_f32_bf16: negative size");
    if (rows == 0 || f == 0) return GG_OK;
    GG_REQUIRE(src && dst && ld >= f && ld_dst >= f, "gg_cast_f32_bf16: bad operands");
    GG_REQUIRE(f % 4 == 0 && ld % 4 == 0 && ld_dst % 4 == 0 && h_al16(src) && (reinterpret_cast<uintptr_t>(dst) & 7) == 0,
               "gg_cast_f32_bf16: needs f %% 4 == 0 and aligned rows");
    int64_t grid = ceil_div(rows * (f / 4), 256 * 4);
    if (grid > (int64_t)kNumSMs * 8) grid = (int64_t)kNumSMs * 8;
    if (grid < 1) grid = 1;
    cast_bf16_kernel<<<(int)grid, 256, 0, as_stream(stream)>>>(src, ld, rows, (int)f, dst, ld_dst);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_spmm_mp_bf16(const int32_t* rowptr, const int32_t* nbr, const float* w_slot, const int32_t* item_row,
                    const int32_t* item_slot, int64_t items, const uint16_t* x_bf16, int64_t ldx, float* out,
                    int64_t ldo, int64_t num_rows, int64_t f, int reduce, const float* x_self, int64_t ld_self,
                    float self_scale, const float* bias, void* workspace, size_t workspace_bytes, int flags,
                    gg_stream_t stream) {
    GG_REQUIRE(num_rows >= 0 && f >= 0 && items >= 0, "gg_spmm_mp_bf16: negative size");
    GG_REQUIRE(reduce == GG_SUM || reduce == GG_MEAN, "gg_spmm_mp_bf16: reduce=%d", reduce);
    if (num_rows == 0 || f == 0) return GG_OK;
    if (f % 8 != 0 || f > 256) {
        set_error("gg_spmm_mp_bf16: needs f %% 8 == 0 and f <= 256 (got %lld)", (long long)f);
        return GG_ERR_UNSUPPORTED;
    }
    GG_REQUIRE(rowptr && item_row && item_slot && x_bf16 && out && workspace, "gg_spmm_mp_bf16: null pointer");
    GG_REQUIRE(items >= 1 && items < ((int64_t)1 << 31) - 1 && num_rows < ((int64_t)1 << 31),
               "gg_spmm_mp_bf16: size out of range");
    GG_REQUIRE(ldx % 8 == 0 && ldx >= f && ldx < ((int64_t)1 << 30) && ldo % 4 == 0 && ldo >= f && h_al16(x_bf16) &&
                   h_al16(out) && (!bias || h_al16(bias)) && (!x_self || (h_al16(x_self) && ld_self % 4 == 0 && ld_self >= f)),
               "gg_spmm_mp_bf16: rows must be 16-byte aligned");
    const size_t need = 256 + 2 * align_up((size_t)items * (size_t)f * sizeof(float), 256);
    if (workspace_bytes < need) {
        set_error("gg_spmm_mp_bf16: workspace %zu < %zu", workspace_bytes, need);
        return GG_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    Carver c(workspace);
    int* counter = c.take<int>(64);
    float* carry = c.take<float>((size_t)items * f);
    float* head = c.take<float>((size_t)items * f);
    GG_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
    HArgs a{rowptr, nbr, w_slot, item_row, item_slot, (int)items, x_bf16, ldx, out, ldo, num_rows, (int)f, reduce,
            bias, x_self, ld_self, self_scale, counter, carry, head, (flags & 4) ? 1 : 0};
    const int nvec = (int)(f / 8);
    if (nvec <= 4) launch_h<4>(a, st);
    else if (nvec <= 8) launch_h<8>(a, st);
    else if (nvec <= 16) launch_h<16>(a, st);
    else launch_h<32>(a, st);
    GG_CUDA(cudaPeekAtLastError());
    return GG_OK;
}

}  // extern "C"
