// Dense transform with ID-GNN heterogeneous weights — fp32 CUDA-core path.
//
//   out[N,F] = act( sum_g diag(scale_g) A_g[N,K_g] B_g + bias ) .* (mask > 0)
//
// The reference computes X*W, gathers X[id], multiplies by W_id and index_add_s the M rows back
// (ref: idconv.py:64-67,152-155,248-251,307-310,372-375): 2 GEMM launches + gather + scatter.
// Here every K-segment accumulates into the same 128x128 register tile, the ID segment scales its
// A rows by the multiplicity of the row in `id` while loading them, and a tile with no centre row
// skips that segment, so the heterogeneous transform is one pass that writes H once.
//
// This file is the exact-fp32 path (parity mode, <= 1e-5 rel against the oracle); the tensor-core
// path lives in gemm_tc.cu.  128x128x16 tiles, 256 threads, 8x8 outputs per thread, global->register
// prefetch of the next k-slab while the current one is multiplied out of shared memory.
#include "common.cuh"

namespace gg {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int kGemmThreads = 256;
constexpr int kPad = 4;  // keeps float4 alignment of the smem rows

struct GemmArgs {
    gg_gemm_segment seg[GG_GEMM_MAX_SEGMENTS];
    int num_segments;
    int b_trans;
    int64_t n;
    int f;
    const float* bias;
    int act;
    const float* relu_mask;
    int64_t ld_mask;
    float* out;
    int64_t ldo;
};

struct Stage {  // what one thread holds of the next k-slab
    float a[8];
    float b[8];
};

static inline __host__ __device__ bool ptr_aligned16(const void* p) {
    return (reinterpret_cast<uintptr_t>(p) & 15) == 0;
}

// Tile whose smem second index runs over matrix rows and whose global rows are contiguous along k:
// used for A ([n,k] row-major) and for B when b_trans ([f,k] row-major).  dst[e] layout:
//   vec   : thread t owns rows r = t/4 + 64*i (i<2), k-quad kq = t%4 -> values [i*4 + c]
//   scalar: element e = t + 256*i (i<8): r = e/16, kk = e%16
template <bool VEC>
__device__ __forceinline__ void load_kmajor(const float* __restrict__ p, int64_t ld, int64_t rows,
                                            int64_t kdim, int64_t row0, int64_t k0,
                                            const float* __restrict__ scale, float* dst) {
    const int t = threadIdx.x;
    if (VEC) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int64_t r = row0 + (t >> 2) + 64 * i;
            int64_t k = k0 + (t & 3) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < rows && k < kdim) {  // kdim % 4 == 0 on this path: the quad is all in or all out
                v = __ldg(reinterpret_cast<const float4*>(p + r * ld + k));
                if (scale) {
                    float s = __ldg(scale + r);
                    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
                }
            }
            dst[i * 4 + 0] = v.x; dst[i * 4 + 1] = v.y; dst[i * 4 + 2] = v.z; dst[i * 4 + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int e = t + 256 * i;
            int64_t r = row0 + (e >> 4);
            int64_t k = k0 + (e & 15);
            float v = 0.f;
            if (r < rows && k < kdim) {
                v = __ldg(p + r * ld + k);
                if (scale) v *= __ldg(scale + r);
            }
            dst[i] = v;
        }
    }
}
template <bool VEC>
__device__ __forceinline__ void store_kmajor(float (*s)[BM + kPad], const float* src) {
    const int t = threadIdx.x;
    if (VEC) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int r = (t >> 2) + 64 * i;
            int kk = (t & 3) * 4;
#pragma unroll
            for (int c = 0; c < 4; ++c) s[kk + c][r] = src[i * 4 + c];
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int e = t + 256 * i;
            s[e & 15][e >> 4] = src[i];
        }
    }
}

// Tile whose global rows run over k and are contiguous along the smem second index:
// B [k,f] row-major, and both operands of the weight-gradient kernel.
//   vec   : e = t + 256*i (i<2): kk = e/32, quad = e%32
//   scalar: e = t + 256*i (i<8): kk = e/128, c = e%128
template <bool VEC>
__device__ __forceinline__ void load_nmajor(const float* __restrict__ p, int64_t ld, int64_t kdim,
                                            int64_t cols, int64_t k0, int64_t col0, float* dst) {
    const int t = threadIdx.x;
    if (VEC) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int e = t + 256 * i;
            int64_t k = k0 + (e >> 5);
            int64_t c = col0 + (e & 31) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < kdim && c < cols) v = __ldg(reinterpret_cast<const float4*>(p + k * ld + c));
            dst[i * 4 + 0] = v.x; dst[i * 4 + 1] = v.y; dst[i * 4 + 2] = v.z; dst[i * 4 + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int e = t + 256 * i;
            int64_t k = k0 + (e >> 7);
            int64_t c = col0 + (e & 127);
            dst[i] = (k < kdim && c < cols) ? __ldg(p + k * ld + c) : 0.f;
        }
    }
}
template <bool VEC>
__device__ __forceinline__ void store_nmajor(float (*s)[BN + kPad], const float* src) {
    const int t = threadIdx.x;
    if (VEC) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int e = t + 256 * i;
            *reinterpret_cast<float4*>(&s[e >> 5][(e & 31) * 4]) =
                make_float4(src[i * 4 + 0], src[i * 4 + 1], src[i * 4 + 2], src[i * 4 + 3]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int e = t + 256 * i;
            s[e >> 7][e & 127] = src[i];
        }
    }
}

// acc[i][j] += As[k][rows(i)] * Bs[k][cols(j)];  rows(i) = ty*4 + (i&3) + 64*(i>>2), same for cols
__device__ __forceinline__ void tile_fma(const float (*As)[BM + kPad], const float (*Bs)[BN + kPad],
                                         float acc[8][8]) {
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
    for (int k = 0; k < BK; ++k) {
        float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
        float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
        float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
}

template <bool VEC_A, bool VEC_B>
__global__ void __launch_bounds__(kGemmThreads, 2) id_gemm_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[2][BK][BM + kPad];
    __shared__ __align__(16) float Bs[2][BK][BN + kPad];
    const int64_t row0 = (int64_t)blockIdx.x * BM;
    const int64_t col0 = (int64_t)blockIdx.y * BN;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int sg = 0; sg < g.num_segments; ++sg) {
        const gg_gemm_segment s = g.seg[sg];
        if (s.k <= 0) continue;
        if (s.scale) {  // ID segment: skip the tile when it holds no centre row
            int64_t r = row0 + threadIdx.x;
            int any = (threadIdx.x < BM && r < g.n) ? (__ldg(s.scale + r) != 0.f) : 0;
            if (!__syncthreads_or(any)) continue;
        }
        Stage st;
        auto fetch = [&](int64_t k0) {
            load_kmajor<VEC_A>(s.a, s.lda, g.n, s.k, row0, k0, s.scale, st.a);
            if (g.b_trans) load_kmajor<VEC_B>(s.b, s.ldb, g.f, s.k, col0, k0, nullptr, st.b);
            else load_nmajor<VEC_B>(s.b, s.ldb, s.k, g.f, k0, col0, st.b);
        };
        auto commit = [&](int buf) {
            store_kmajor<VEC_A>(As[buf], st.a);
            if (g.b_trans) store_kmajor<VEC_B>(Bs[buf], st.b);
            else store_nmajor<VEC_B>(Bs[buf], st.b);
        };
        fetch(0);
        __syncthreads();  // previous segment's readers are done with buffer 0
        commit(0);
        __syncthreads();
        int buf = 0;
        for (int64_t k0 = 0; k0 < s.k; k0 += BK) {
            const bool more = k0 + BK < s.k;
            if (more) fetch(k0 + BK);
            tile_fma(As[buf], Bs[buf], acc);
            if (more) commit(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }

    // epilogue: bias -> act -> mask -> store
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const bool vec_out = (g.f % 4 == 0) && (g.ldo % 4 == 0) && ptr_aligned16(g.out) &&
                         (!g.relu_mask || (g.ld_mask % 4 == 0 && ptr_aligned16(g.relu_mask)));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = row0 + ty * 4 + (i & 3) + 64 * (i >> 2);
        if (r >= g.n) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t c = col0 + tx * 4 + 64 * h;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[j] = acc[i][h * 4 + j];
                if (g.bias && c + j < g.f) v[j] += __ldg(g.bias + c + j);
                if (g.act == GG_ACT_RELU) v[j] = fmaxf(v[j], 0.f);
                if (g.relu_mask && c + j < g.f)
                    v[j] = __ldg(g.relu_mask + r * g.ld_mask + c + j) > 0.f ? v[j] : 0.f;
            }
            if (vec_out && c + 3 < g.f) {
                *reinterpret_cast<float4*>(g.out + r * g.ldo + c) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c + j < g.f) g.out[r * g.ldo + c + j] = v[j];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// weight gradient: out[K,F] = sum_r A[row(r),:]^T G[row(r),:], split over rows
// ---------------------------------------------------------------------------------------------
struct TnArgs {
    const float* a;
    int64_t lda;
    const float* g;
    int64_t ldg;
    const int64_t* row_index;  // nullable: row(r) = row_index[r]
    int64_t n, k, f;
    int64_t rows_per_split;
    float* dst;  // [splits, k, f] partials (or the output itself when splits == 1)
    int64_t ld_dst;
    int64_t split_stride;
};

// one reduction slab: 16 matrix rows x 128 columns; thread t loads 8 values
template <bool VEC>
__device__ __forceinline__ void load_rows(const float* __restrict__ p, int64_t ld, int64_t cols,
                                          const int64_t* __restrict__ row_index, int64_t r0,
                                          int64_t r_end, int64_t col0, float* dst) {
    const int t = threadIdx.x;
    if (VEC) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int e = t + 256 * i;
            int64_t r = r0 + (e >> 5);
            int64_t c = col0 + (e & 31) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < r_end && c < cols) {
                int64_t rr = row_index ? __ldg(row_index + r) : r;
                v = __ldg(reinterpret_cast<const float4*>(p + rr * ld + c));
            }
            dst[i * 4 + 0] = v.x; dst[i * 4 + 1] = v.y; dst[i * 4 + 2] = v.z; dst[i * 4 + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int e = t + 256 * i;
            int64_t r = r0 + (e >> 7);
            int64_t c = col0 + (e & 127);
            float v = 0.f;
            if (r < r_end && c < cols) {
                int64_t rr = row_index ? __ldg(row_index + r) : r;
                v = __ldg(p + rr * ld + c);
            }
            dst[i] = v;
        }
    }
}

template <bool VEC_A, bool VEC_G>
__global__ void __launch_bounds__(kGemmThreads, 2) gemm_tn_kernel(TnArgs g) {
    __shared__ __align__(16) float As[2][BK][BM + kPad];
    __shared__ __align__(16) float Bs[2][BK][BN + kPad];
    const int tiles_f = (int)((g.f + BN - 1) / BN);
    const int64_t m0 = (int64_t)(blockIdx.x / tiles_f) * BM;  // offset in K (rows of out)
    const int64_t c0 = (int64_t)(blockIdx.x % tiles_f) * BN;  // offset in F
    const int64_t r_beg = (int64_t)blockIdx.y * g.rows_per_split;
    const int64_t r_end = r_beg + g.rows_per_split < g.n ? r_beg + g.rows_per_split : g.n;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    Stage st;
    auto fetch = [&](int64_t r0) {
        load_rows<VEC_A>(g.a, g.lda, g.k, g.row_index, r0, r_end, m0, st.a);
        load_rows<VEC_G>(g.g, g.ldg, g.f, g.row_index, r0, r_end, c0, st.b);
    };
    auto commit = [&](int buf) {
        store_nmajor<VEC_A>(As[buf], st.a);
        store_nmajor<VEC_G>(Bs[buf], st.b);
    };
    if (r_beg < r_end) {
        fetch(r_beg);
        commit(0);
        __syncthreads();
        int buf = 0;
        for (int64_t r0 = r_beg; r0 < r_end; r0 += BK) {
            const bool more = r0 + BK < r_end;
            if (more) fetch(r0 + BK);
            tile_fma(As[buf], Bs[buf], acc);
            if (more) commit(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    float* dst = g.dst + (int64_t)blockIdx.y * g.split_stride;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = m0 + ty * 4 + (i & 3) + 64 * (i >> 2);
        if (r >= g.k) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t c = c0 + tx * 4 + (j & 3) + 64 * (j >> 2);
            if (c < g.f) dst[r * g.ld_dst + c] = acc[i][j];
        }
    }
}

// out[e] = sum_s part[s][e], s ascending (fixed order)
__global__ void __launch_bounds__(256) split_reduce_kernel(const float* __restrict__ part,
                                                           int64_t splits, int64_t rows, int64_t cols,
                                                           int64_t split_stride, float* __restrict__ out,
                                                           int64_t ldo) {
    int64_t total = rows * cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int64_t p = 0; p < splits; ++p) s += part[p * split_stride + e];
        out[(e / cols) * ldo + (e % cols)] = s;
    }
}

// column sums: block b sums rows [b*chunk, (b+1)*chunk) for every column
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ g, int64_t ldg,
                                                             int64_t n, int64_t f, int64_t chunk,
                                                             float* __restrict__ part) {
    int64_t r_beg = (int64_t)blockIdx.x * chunk;
    int64_t r_end = r_beg + chunk < n ? r_beg + chunk : n;
    for (int64_t c = threadIdx.x; c < f; c += blockDim.x) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int64_t r = r_beg;
        for (; r + 3 < r_end; r += 4) {
            s0 += __ldg(g + r * ldg + c);
            s1 += __ldg(g + (r + 1) * ldg + c);
            s2 += __ldg(g + (r + 2) * ldg + c);
            s3 += __ldg(g + (r + 3) * ldg + c);
        }
        for (; r < r_end; ++r) s0 += __ldg(g + r * ldg + c);
        part[(int64_t)blockIdx.x * f + c] = (s0 + s1) + (s2 + s3);
    }
}

// Vector variant (f % 4 == 0, f <= 128, 16-byte aligned rows): a warp covers one row per 128-bit load, eight
// rows in flight per warp; warps are combined in shared memory in a fixed order.
__global__ void __launch_bounds__(256) colsum_partial_vec_kernel(const float* __restrict__ g, int64_t ldg, int64_t n,
                                                                 int nvec, int64_t chunk, float* __restrict__ part) {
    __shared__ float4 s_acc[8][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t r_beg = (int64_t)blockIdx.x * chunk;
    const int64_t r_end = r_beg + chunk < n ? r_beg + chunk : n;
    float4 acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < nvec) {
        int64_t r = r_beg + wid;
        for (; r + 56 < r_end; r += 64) {  // rows r, r+8, ..., r+56: eight independent loads
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = ldg_nc_f4(reinterpret_cast<const float4*>(g + (r + 8 * u) * ldg) + lane);
#pragma unroll
            for (int u = 0; u < 8; ++u) add4(acc[u & 3], v[u]);
        }
        for (; r < r_end; r += 8) add4(acc[0], ldg_nc_f4(reinterpret_cast<const float4*>(g + r * ldg) + lane));
    }
    add4(acc[0], acc[1]);
    add4(acc[2], acc[3]);
    add4(acc[0], acc[2]);
    s_acc[wid][lane] = acc[0];
    __syncthreads();
    if (wid == 0 && lane < nvec) {
        float4 t = s_acc[0][lane];
#pragma unroll
        for (int q = 1; q < 8; ++q) add4(t, s_acc[q][lane]);
        reinterpret_cast<float4*>(part + (int64_t)blockIdx.x * (nvec * 4))[lane] = t;
    }
}
// out[c] = sum_b part[b][c] in block order, four independent chains per thread
__global__ void __launch_bounds__(128) colsum_final_kernel(const float* __restrict__ part, int64_t blocks, int f,
                                                           float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= f) return;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int64_t b = 0;
    for (; b + 3 < blocks; b += 4) {
        a0 += part[b * f + c];
        a1 += part[(b + 1) * f + c];
        a2 += part[(b + 2) * f + c];
        a3 += part[(b + 3) * f + c];
    }
    for (; b < blocks; ++b) a0 += part[b * f + c];
    out[c] = (a0 + a1) + (a2 + a3);
}

static int64_t colsum_blocks(int64_t n) {
    int64_t b = ceil_div(n, 256);
    int64_t cap = (int64_t)kNumSMs * 8;
    if (b > cap) b = cap;
    return b < 1 ? 1 : b;
}

static void tn_plan(int64_t n, int64_t k, int64_t f, int64_t* splits, int64_t* rows_per_split) {
    int64_t tiles = ceil_div(k, BM) * ceil_div(f, BN);
    int64_t want = ceil_div((int64_t)kNumSMs * 2, tiles);
    int64_t max_by_rows = ceil_div(n, (int64_t)BK * 8);
    int64_t s = want < max_by_rows ? want : max_by_rows;
    if (s < 1) s = 1;
    int64_t rps = ceil_div(ceil_div(n, s), BK) * BK;
    if (rps < BK) rps = BK;
    *splits = ceil_div(n > 0 ? n : 1, rps);
    *rows_per_split = rps;
}

}  // namespace gg

using namespace gg;

extern "C" {

int gg_id_gemm_f32(const gg_gemm_segment* segs, int num_segments, int b_trans, int64_t n, int64_t f,
                   const float* bias, int act, const float* relu_mask, int64_t ld_mask, float* out,
                   int64_t ldo, gg_stream_t stream) {
    GG_REQUIRE(segs && num_segments >= 1 && num_segments <= GG_GEMM_MAX_SEGMENTS,
               "gg_id_gemm_f32: num_segments=%d", num_segments);
    GG_REQUIRE(n >= 0 && f >= 0, "gg_id_gemm_f32: negative size");
    GG_REQUIRE(act == GG_ACT_NONE || act == GG_ACT_RELU, "gg_id_gemm_f32: act=%d", act);
    if (n == 0 || f == 0) return GG_OK;
    GG_REQUIRE(out && ldo >= f, "gg_id_gemm_f32: bad output");
    GG_REQUIRE(!relu_mask || ld_mask >= f, "gg_id_gemm_f32: bad mask stride");
    GG_REQUIRE(f < ((int64_t)1 << 30), "gg_id_gemm_f32: f out of range");
    GemmArgs g{};
    bool vec_a = true, vec_b = true;
    for (int i = 0; i < num_segments; ++i) {
        const gg_gemm_segment& s = segs[i];
        GG_REQUIRE(s.k >= 0, "gg_id_gemm_f32: segment %d has negative k", i);
        if (s.k > 0) {
            GG_REQUIRE(s.a && s.b, "gg_id_gemm_f32: segment %d has a null operand", i);
            GG_REQUIRE(s.lda >= s.k, "gg_id_gemm_f32: segment %d lda < k", i);
            GG_REQUIRE(s.ldb >= (b_trans ? s.k : f), "gg_id_gemm_f32: segment %d ldb too small", i);
            vec_a = vec_a && (s.k % 4 == 0) && (s.lda % 4 == 0) && ptr_aligned16(s.a);
            if (b_trans) vec_b = vec_b && (s.k % 4 == 0) && (s.ldb % 4 == 0) && ptr_aligned16(s.b);
            else vec_b = vec_b && (f % 4 == 0) && (s.ldb % 4 == 0) && ptr_aligned16(s.b);
        }
        g.seg[i] = s;
    }
    g.num_segments = num_segments;
    g.b_trans = b_trans;
    g.n = n;
    g.f = (int)f;
    g.bias = bias;
    g.act = act;
    g.relu_mask = relu_mask;
    g.ld_mask = ld_mask;
    g.out = out;
    g.ldo = ldo;
    dim3 grid((unsigned)ceil_div(n, BM), (unsigned)ceil_div(f, BN));
    cudaStream_t st = as_stream(stream);
    if (vec_a && vec_b) id_gemm_kernel<true, true><<<grid, kGemmThreads, 0, st>>>(g);
    else if (vec_a) id_gemm_kernel<true, false><<<grid, kGemmThreads, 0, st>>>(g);
    else if (vec_b) id_gemm_kernel<false, true><<<grid, kGemmThreads, 0, st>>>(g);
    else id_gemm_kernel<false, false><<<grid, kGemmThreads, 0, st>>>(g);
    GG_LAUNCHED();
    return GG_OK;
}

size_t gg_gemm_tn_workspace_bytes(int64_t n, int64_t k, int64_t f) {
    int64_t splits, rps;
    tn_plan(n, k, f, &splits, &rps);
    return (size_t)splits * (size_t)k * (size_t)f * sizeof(float) + 256;
}

static int gemm_tn_impl(const float* a, int64_t lda, const int64_t* row_index, const float* g,
                        int64_t ldg, int64_t n, int64_t k, int64_t f, float* out, int64_t ldo,
                        void* workspace, size_t workspace_bytes, gg_stream_t stream) {
    GG_REQUIRE(n >= 0 && k >= 0 && f >= 0, "gg_gemm_tn_f32: negative size");
    if (k == 0 || f == 0) return GG_OK;
    GG_REQUIRE(out && ldo >= f, "gg_gemm_tn_f32: bad output");
    cudaStream_t st = as_stream(stream);
    if (n == 0) {
        GG_CUDA(cudaMemset2DAsync(out, (size_t)ldo * 4, 0, (size_t)f * 4, (size_t)k, st));
        return GG_OK;
    }
    GG_REQUIRE(a && g && lda >= k && ldg >= f, "gg_gemm_tn_f32: bad operands");
    int64_t splits, rps;
    tn_plan(n, k, f, &splits, &rps);
    TnArgs t{};
    t.a = a; t.lda = lda; t.g = g; t.ldg = ldg; t.row_index = row_index;
    t.n = n; t.k = k; t.f = f; t.rows_per_split = rps;
    if (splits == 1) {
        t.dst = out; t.ld_dst = ldo; t.split_stride = 0;
    } else {
        if (!workspace || workspace_bytes < gg_gemm_tn_workspace_bytes(n, k, f)) {
            set_error("gg_gemm_tn_f32: workspace %zu < %zu", workspace_bytes,
                      gg_gemm_tn_workspace_bytes(n, k, f));
            return GG_ERR_WORKSPACE;
        }
        t.dst = static_cast<float*>(workspace); t.ld_dst = f; t.split_stride = k * f;
    }
    bool vec_a = (k % 4 == 0) && (lda % 4 == 0) && ptr_aligned16(a);
    bool vec_g = (f % 4 == 0) && (ldg % 4 == 0) && ptr_aligned16(g);
    dim3 grid((unsigned)(ceil_div(k, BM) * ceil_div(f, BN)), (unsigned)splits);
    if (vec_a && vec_g) gemm_tn_kernel<true, true><<<grid, kGemmThreads, 0, st>>>(t);
    else if (vec_a) gemm_tn_kernel<true, false><<<grid, kGemmThreads, 0, st>>>(t);
    else if (vec_g) gemm_tn_kernel<false, true><<<grid, kGemmThreads, 0, st>>>(t);
    else gemm_tn_kernel<false, false><<<grid, kGemmThreads, 0, st>>>(t);
    GG_LAUNCHED();
    if (splits > 1) {
        int64_t total = k * f;
        int blocks = (int)(ceil_div(total, 256) < kNumSMs * 8 ? ceil_div(total, 256) : kNumSMs * 8);
        split_reduce_kernel<<<blocks, 256, 0, st>>>(t.dst, splits, k, f, t.split_stride, out, ldo);
        GG_LAUNCHED();
    }
    return GG_OK;
}

int gg_gemm_tn_f32(const float* a, int64_t lda, const int64_t* row_index, const float* g, int64_t ldg,
                   int64_t n, int64_t k, int64_t f, float* out, int64_t ldo, void* workspace,
                   size_t workspace_bytes, gg_stream_t stream) {
    return gemm_tn_impl(a, lda, row_index, g, ldg, n, k, f, out, ldo, workspace, workspace_bytes,
                        stream);
}

size_t gg_colsum_workspace_bytes(int64_t n, int64_t f) {
    return (size_t)colsum_blocks(n) * (size_t)(f > 0 ? f : 1) * sizeof(float) + 256;
}

int gg_colsum_f32(const float* g, int64_t ldg, int64_t n, int64_t f, float* out, void* workspace,
                  size_t workspace_bytes, gg_stream_t stream) {
    GG_REQUIRE(n >= 0 && f >= 0, "gg_colsum_f32: negative size");
    if (f == 0) return GG_OK;
    GG_REQUIRE(out, "gg_colsum_f32: null output");
    cudaStream_t st = as_stream(stream);
    if (n == 0) {
        GG_CUDA(cudaMemsetAsync(out, 0, (size_t)f * 4, st));
        return GG_OK;
    }
    GG_REQUIRE(g && ldg >= f && workspace, "gg_colsum_f32: bad operands");
    if (workspace_bytes < gg_colsum_workspace_bytes(n, f)) {
        set_error("gg_colsum_f32: workspace %zu < %zu", workspace_bytes,
                  gg_colsum_workspace_bytes(n, f));
        return GG_ERR_WORKSPACE;
    }
    int64_t blocks = colsum_blocks(n);
    int64_t chunk = ceil_div(n, blocks);
    blocks = ceil_div(n, chunk);
    float* part = static_cast<float*>(workspace);
    if (f % 4 == 0 && f <= 128 && ldg % 4 == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        int64_t vb = ceil_div(n, 512);  // >= 64 rows per warp before a block is worth launching
        if (vb > (int64_t)kNumSMs * 4) vb = (int64_t)kNumSMs * 4;
        if (vb > blocks) vb = blocks;   // the workspace is sized for `blocks` partial rows
        if (vb < 1) vb = 1;
        const int64_t vchunk = ceil_div(n, vb);
        vb = ceil_div(n, vchunk);
        colsum_partial_vec_kernel<<<(int)vb, 256, 0, st>>>(g, ldg, n, (int)(f / 4), vchunk, part);
        GG_LAUNCHED();
        colsum_final_kernel<<<(int)ceil_div(f, 128), 128, 0, st>>>(part, vb, (int)f, out);
        GG_LAUNCHED();
        return GG_OK;
    }
    int threads = f >= 256 ? 256 : (int)(ceil_div(f, 32) * 32);
    colsum_partial_kernel<<<(int)blocks, threads, 0, st>>>(g, ldg, n, f, chunk, part);
    GG_LAUNCHED();
    int rb = (int)(ceil_div(f, 256));
    split_reduce_kernel<<<rb, 256, 0, st>>>(part, blocks, 1, f, f, out, f);
    GG_LAUNCHED();
    return GG_OK;
}

}  // extern "C"
