// Dense transform with ID-GNN heterogeneous weights on the 5th-generation tensor cores
// (SURVEY §8a row 9, K7): tcgen05.mma, accumulators in TMEM, fp32 in / fp32 out.
//
//   out[N,F] = act( sum_g diag(scale_g) A_g[N,K_g] B_g + bias ) .* (mask > 0)       (same contract as gemm.cu)
//
// fp32 parity (<= 1e-5 relative) through a tensor core: 3xTF32.  Every operand is split
//   x = hi + lo,  hi = rna_tf32(x), lo = rna_tf32(x - hi)   (round-to-nearest: unbiased, both exact TF32 values)
// and the product is  hi_a*hi_b + hi_a*lo_b + lo_a*hi_b  (the dropped lo*lo term is ~2^-22 relative),
// three kind::tf32 MMAs accumulating into the same fp32 TMEM tile.
//
// Structure of one CTA (128 x 128 output tile, 256 threads, up to 3 CTAs per SM, 128 TMEM columns each):
//   * the K-segments (X W, then the centre rows' X W_id with the row multiplicity as scale; tiles with
//     no centre row skip it) are walked in slabs of 16 k;
//   * A slab: each thread loads 2 x 16 B of its row, applies the scale, splits hi/lo in registers and
//     stores both into shared memory in the UMMA canonical K-major no-swizzle layout
//     (16-byte chunks, chunk-major: address = chunk * 2048 + row * 16, i.e. LBO = 2048 B, SBO = 128 B);
//   * B slab: pre-split once per call by b_image_kernel into exactly that layout, copied linearly;
//   * fence.proxy.async + __syncthreads, then ONE thread issues the 6 tcgen05.mma of the slab and a
//     tcgen05.commit onto the stage's mbarrier; two stages, the next slab's global loads are in flight
//     while the tensor core works;
//   * epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> bias / ReLU / mask -> global.
#include "common.cuh"

namespace gg {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 16;
constexpr int TC_THREADS = 256;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;        // 8192: one operand slab (hi or lo)
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;       // A_hi, A_lo, B_hi, B_lo
constexpr int TC_SMEM_BYTES = 2 * TC_STAGE_BYTES;       // two stages
constexpr uint32_t TC_LBO = TC_BM * 16;                 // bytes between 16-byte k-chunks
constexpr uint32_t TC_SBO = 128;                        // bytes between 8-row groups
constexpr int TC_TMEM_COLS = 128;

struct TcSegment {
    const float* a;
    int64_t lda;
    const float* scale;
    const float* b_image;  // [n_tiles][k_slabs][2][TC_TILE_BYTES / 4]
    int k;
    int k_slabs;
};
struct TcArgs {
    TcSegment seg[GG_GEMM_MAX_SEGMENTS];
    int num_segments;
    int64_t n;
    int f;
    const float* bias;
    int act;
    const float* relu_mask;
    int64_t ld_mask;
    float* out;
    int64_t ldo;
    int accumulate_out;  // epilogue adds the tile already in `out` (a previous K-chunk's partial)
    int final_chunk;     // bias / act / mask are applied by the last chunk only
};

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (int spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(tc_smem_u32(bar)), "r"(parity)
            : "memory");
        if (spin > (1 << 22)) __trap();  // fail the launch instead of hanging the GPU
    }
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version = 1, [61,64) layout type = 0
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(TC_LBO >> 4) << 16) |
           ((uint64_t)(TC_SBO >> 4) << 32) | (1ull << 46);
}
// kind::tf32 instruction descriptor: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major,
// N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
                              ((uint32_t)(TC_BM >> 4) << 24);

__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(TC_IDESC), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u),
        "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     tc_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
          "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// hi = x rounded to nearest TF32, lo = (x - hi) rounded to nearest TF32: both are exact TF32 values, so
// the tensor core's own operand truncation never fires and the split error is unbiased.
__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = rna_tf32(x);
    lo = rna_tf32(x - hi);
}

// B operand image: for N-tile nt, k-slab ks: [hi | lo] slabs in the canonical layout,
//   element (n, k) of the slab at float offset (k/4)*512 + n*4 + (k%4)   (chunk-major, 16-byte chunks)
__global__ void __launch_bounds__(256)
    b_image_kernel(const float* __restrict__ b, int64_t ldb, int b_trans, int k, int f, int k_slabs, int n_tiles,
                   float* __restrict__ image) {
    const int64_t total = (int64_t)n_tiles * k_slabs * (TC_BK / 4) * TC_BN;  // one thread per 16-byte chunk
    for (int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; id < total;
         id += (int64_t)gridDim.x * blockDim.x) {
        const int nn = (int)(id % TC_BN);
        const int c = (int)((id / TC_BN) % (TC_BK / 4));
        const int ks = (int)((id / (TC_BN * (TC_BK / 4))) % k_slabs);
        const int nt = (int)(id / ((int64_t)TC_BN * (TC_BK / 4) * k_slabs));
        const int col = nt * TC_BN + nn;
        float hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int kk = ks * TC_BK + c * 4 + i;
            float v = 0.f;
            if (kk < k && col < f) v = b_trans ? __ldg(b + (int64_t)col * ldb + kk) : __ldg(b + (int64_t)kk * ldb + col);
            split_tf32(v, hi[i], lo[i]);
        }
        float* slab = image + ((int64_t)nt * k_slabs + ks) * (2 * TC_TILE_BYTES / 4);
        const int off = c * (TC_BN * 4) + nn * 4;
        *reinterpret_cast<float4*>(slab + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(slab + TC_TILE_BYTES / 4 + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}

template <bool VEC_A>
__global__ void __launch_bounds__(TC_THREADS, 3) tc_gemm_kernel(TcArgs g) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t mma_done[2];
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t row0 = (int64_t)blockIdx.x * TC_BM;
    const int nt = blockIdx.y;
    const int n_tiles = gridDim.y;

    if (tid == 0) {
        tc_mbar_init(&mma_done[0], 1);
        tc_mbar_init(&mma_done[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // TMEM: 128 fp32 accumulator columns for this CTA
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         tc_smem_u32(&tmem_base_smem)),
                     "r"(TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    // this thread's share of a slab: A row (tid & 127), chunks 2*(tid>>7) and +1; 64 B of the B image
    const int a_row = tid & 127;
    const int a_c0 = (tid >> 7) * 2;
    const int64_t grow = row0 + a_row;
    const bool row_ok = grow < g.n;

    int it = 0;  // slabs issued so far
    for (int sg = 0; sg < g.num_segments; ++sg) {
        const TcSegment s = g.seg[sg];
        if (s.k_slabs <= 0) continue;
        float sc = 1.f;
        if (s.scale) {
            sc = row_ok ? __ldg(s.scale + grow) : 0.f;
            if (!__syncthreads_or(sc != 0.f)) continue;  // no centre row in this tile
        }
        const float* arow = s.a + grow * s.lda;
        const float4* bimg = reinterpret_cast<const float4*>(s.b_image) +
                             (int64_t)nt * s.k_slabs * (2 * TC_TILE_BYTES / 16);
        float4 ra[2], rb[4];
        auto fetch = [&](int ks) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int k = ks * TC_BK + (a_c0 + i) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row_ok) {
                    if (VEC_A) {
                        if (k < s.k) v = __ldg(reinterpret_cast<const float4*>(arow + k));
                    } else {
                        if (k + 0 < s.k) v.x = __ldg(arow + k + 0);
                        if (k + 1 < s.k) v.y = __ldg(arow + k + 1);
                        if (k + 2 < s.k) v.z = __ldg(arow + k + 2);
                        if (k + 3 < s.k) v.w = __ldg(arow + k + 3);
                    }
                }
                ra[i] = v;
            }
            const float4* src = bimg + (int64_t)ks * (2 * TC_TILE_BYTES / 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) rb[j] = __ldg(src + tid + j * TC_THREADS);
        };
        fetch(0);
        for (int ks = 0; ks < s.k_slabs; ++ks, ++it) {
            const int stage = it & 1;
            uint8_t* st = smem + stage * TC_STAGE_BYTES;
            if (it >= 2) tc_mbar_wait(&mma_done[stage], (uint32_t)((it >> 1) - 1) & 1u);  // slab it-2 consumed
            // A: scale, split, store hi / lo;  B: linear copy of the prepared image
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                float4 h, l;
                split_tf32(ra[i].x * sc, h.x, l.x);
                split_tf32(ra[i].y * sc, h.y, l.y);
                split_tf32(ra[i].z * sc, h.z, l.z);
                split_tf32(ra[i].w * sc, h.w, l.w);
                const int off = (a_c0 + i) * (int)TC_LBO + a_row * 16;
                *reinterpret_cast<float4*>(st + off) = h;
                *reinterpret_cast<float4*>(st + TC_TILE_BYTES + off) = l;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4*>(st + 2 * TC_TILE_BYTES + (tid + j * TC_THREADS) * 16) = rb[j];
            if (ks + 1 < s.k_slabs) fetch(ks + 1);  // in flight while the tensor core works
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> async proxy
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = tc_smem_u32(st), a_lo = a_hi + TC_TILE_BYTES;
                const uint32_t b_hi = a_hi + 2 * TC_TILE_BYTES, b_lo = a_hi + 3 * TC_TILE_BYTES;
#pragma unroll
                for (int j = 0; j < TC_BK / 8; ++j) {  // one MMA consumes 8 k = two 16-byte chunks
                    const uint32_t ko = j * 2 * TC_LBO;
                    tc_mma(tmem_base, tc_smem_desc(a_hi + ko), tc_smem_desc(b_hi + ko), (it > 0 || j > 0) ? 1u : 0u);
                    tc_mma(tmem_base, tc_smem_desc(a_hi + ko), tc_smem_desc(b_lo + ko), 1u);
                    tc_mma(tmem_base, tc_smem_desc(a_lo + ko), tc_smem_desc(b_hi + ko), 1u);
                }
                tc_commit(&mma_done[stage]);
            }
        }
    }
    // all MMAs were issued by one thread in order: the last commit covers them all
    if (it > 0) {
        const int last = it - 1;
        tc_mbar_wait(&mma_done[last & 1], (uint32_t)(last >> 1) & 1u);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: TMEM -> registers -> bias / act / mask -> global ----
    const int q = warp & 3;        // TMEM lane quarter this warp may read
    const int half = warp >> 2;    // column half
    const int64_t r = row0 + q * 32 + lane;
    const bool vec_out = (g.f % 4 == 0) && (g.ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.out) & 15) == 0) &&
                         (!g.relu_mask || (g.ld_mask % 4 == 0 && (reinterpret_cast<uintptr_t>(g.relu_mask) & 15) == 0));
#pragma unroll
    for (int part = 0; part < 2; ++part) {
        const int cbase = half * 64 + part * 32;
        uint32_t acc[32];
        if (it > 0) {
            tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cbase, acc);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = 0u;
        }
        if (r < g.n) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const int c = nt * TC_BN + cbase + j;
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    v[e] = __uint_as_float(acc[j + e]);
                    if (c + e < g.f) {
                        if (g.accumulate_out) v[e] += g.out[r * g.ldo + c + e];  // fp32 round-to-nearest add
                        if (g.final_chunk) {
                            if (g.bias) v[e] += __ldg(g.bias + c + e);
                            if (g.act == GG_ACT_RELU) v[e] = fmaxf(v[e], 0.f);
                            if (g.relu_mask) v[e] = __ldg(g.relu_mask + r * g.ld_mask + c + e) > 0.f ? v[e] : 0.f;
                        }
                    }
                }
                if (vec_out && c + 3 < g.f) {
                    *reinterpret_cast<float4*>(g.out + r * g.ldo + c) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (c + e < g.f) g.out[r * g.ldo + c + e] = v[e];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS)
                     : "memory");
    }
    (void)n_tiles;
}

static inline int tc_slabs(int64_t k) { return (int)ceil_div(k, TC_BK); }
static inline int tc_ntiles(int64_t f) { return (int)ceil_div(f, TC_BN); }
static inline size_t tc_image_floats(int64_t k, int64_t f) {
    return (size_t)tc_ntiles(f) * (size_t)tc_slabs(k) * (2 * TC_TILE_BYTES / 4);
}

}  // namespace gg

using namespace gg;

// The tensor core accumulates with truncation: the result is biased towards zero by ~6e-9 * K relative
// (measured, scratch/tc_err_probe.py).  To stay inside 1e-5 for any K the reduction is cut into launches
// of at most kTcMaxKPerLaunch k; each launch's partial tile is added to `out` in fp32 (round to nearest)
// by the next launch's epilogue.  GNN hidden sizes (K <= 512) take one launch.
constexpr int kTcMaxKPerLaunch = 512;

struct TcPiece {  // a sub-range of a caller segment
    int seg;
    int64_t k_off, k_len;
};

static int tc_plan(const gg_gemm_segment* segs, int num_segments, TcPiece* pieces, int* launch_of, int max_pieces) {
    int np = 0, launch = 0, in_launch = 0;
    int64_t k_in_launch = 0;
    for (int i = 0; i < num_segments; ++i) {
        int64_t off = 0;
        while (off < segs[i].k) {
            int64_t room = kTcMaxKPerLaunch - k_in_launch;
            if (room < TC_BK || in_launch == GG_GEMM_MAX_SEGMENTS) {
                ++launch; in_launch = 0; k_in_launch = 0; room = kTcMaxKPerLaunch;
            }
            int64_t len = segs[i].k - off;
            if (len > room) len = room / TC_BK * TC_BK;  // keep slab alignment inside a segment
            if (np >= max_pieces) return -1;
            pieces[np] = TcPiece{i, off, len};
            launch_of[np] = launch;
            ++np; ++in_launch; k_in_launch += len; off += len;
        }
    }
    return np;
}
constexpr int kTcMaxPieces = 256;

extern "C" {

size_t gg_id_gemm_tc_workspace_bytes(const gg_gemm_segment* segs, int num_segments, int64_t f) {
    size_t b = 256;
    TcPiece pieces[kTcMaxPieces];
    int launch_of[kTcMaxPieces];
    int np = tc_plan(segs, num_segments, pieces, launch_of, kTcMaxPieces);
    for (int i = 0; i < np; ++i) b += align_up(tc_image_floats(pieces[i].k_len, f) * 4, 256);
    return b;
}

int gg_id_gemm_tc_f32(const gg_gemm_segment* segs, int num_segments, int b_trans, int64_t n, int64_t f,
                      const float* bias, int act, const float* relu_mask, int64_t ld_mask, float* out,
                      int64_t ldo, void* workspace, size_t workspace_bytes, gg_stream_t stream) {
    GG_REQUIRE(segs && num_segments >= 1 && num_segments <= GG_GEMM_MAX_SEGMENTS,
               "gg_id_gemm_tc_f32: num_segments=%d", num_segments);
    GG_REQUIRE(n >= 0 && f >= 0, "gg_id_gemm_tc_f32: negative size");
    GG_REQUIRE(act == GG_ACT_NONE || act == GG_ACT_RELU, "gg_id_gemm_tc_f32: act=%d", act);
    if (n == 0 || f == 0) return GG_OK;
    GG_REQUIRE(out && ldo >= f && workspace, "gg_id_gemm_tc_f32: bad output / workspace");
    GG_REQUIRE(!relu_mask || ld_mask >= f, "gg_id_gemm_tc_f32: bad mask stride");
    GG_REQUIRE(f < (1 << 20) && tc_ntiles(f) <= 65535, "gg_id_gemm_tc_f32: f out of range");
    for (int i = 0; i < num_segments; ++i) {
        const gg_gemm_segment& s = segs[i];
        GG_REQUIRE(s.k >= 0 && s.k < (1 << 24), "gg_id_gemm_tc_f32: segment %d k out of range", i);
        if (s.k == 0) continue;
        GG_REQUIRE(s.a && s.b && s.lda >= s.k, "gg_id_gemm_tc_f32: segment %d has a bad operand", i);
        GG_REQUIRE(s.ldb >= (b_trans ? s.k : f), "gg_id_gemm_tc_f32: segment %d ldb too small", i);
    }
    TcPiece pieces[kTcMaxPieces];
    int launch_of[kTcMaxPieces];
    const int np = tc_plan(segs, num_segments, pieces, launch_of, kTcMaxPieces);
    GG_REQUIRE(np >= 0, "gg_id_gemm_tc_f32: reduction too long (more than %d pieces of %d)", kTcMaxPieces,
               kTcMaxKPerLaunch);
    if (workspace_bytes < gg_id_gemm_tc_workspace_bytes(segs, num_segments, f)) {
        set_error("gg_id_gemm_tc_f32: workspace %zu < %zu", workspace_bytes,
                  gg_id_gemm_tc_workspace_bytes(segs, num_segments, f));
        return GG_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    static bool attr_done = false;
    if (!attr_done) {
        GG_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        GG_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        attr_done = true;
    }
    Carver c(workspace);
    const int num_launches = np > 0 ? launch_of[np - 1] + 1 : 1;
    dim3 grid((unsigned)ceil_div(n, TC_BM), (unsigned)tc_ntiles(f));
    int p = 0;
    for (int l = 0; l < num_launches; ++l) {
        TcArgs g{};
        bool vec_a = true;
        int ns = 0;
        for (; p < np && launch_of[p] == l; ++p, ++ns) {
            const gg_gemm_segment& s = segs[pieces[p].seg];
            const int64_t k_off = pieces[p].k_off, k_len = pieces[p].k_len;
            TcSegment& t = g.seg[ns];
            t.k = (int)k_len;
            t.k_slabs = tc_slabs(k_len);
            t.a = s.a + k_off; t.lda = s.lda; t.scale = s.scale;
            vec_a = vec_a && (k_len % 4 == 0) && (s.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(t.a) & 15) == 0);
            float* image = c.take<float>(tc_image_floats(k_len, f));
            const float* bsrc = b_trans ? s.b + k_off : s.b + k_off * s.ldb;
            int64_t chunks = (int64_t)tc_ntiles(f) * t.k_slabs * (TC_BK / 4) * TC_BN;
            int bgrid = (int)(ceil_div(chunks, 256) < kNumSMs * 8 ? ceil_div(chunks, 256) : kNumSMs * 8);
            b_image_kernel<<<bgrid, 256, 0, st>>>(bsrc, s.ldb, b_trans, (int)k_len, (int)f, t.k_slabs, tc_ntiles(f),
                                                  image);
            GG_LAUNCHED();
            t.b_image = image;
        }
        g.num_segments = ns;
        g.n = n; g.f = (int)f; g.bias = bias; g.act = act; g.relu_mask = relu_mask; g.ld_mask = ld_mask;
        g.out = out; g.ldo = ldo;
        g.accumulate_out = l > 0;
        g.final_chunk = l == num_launches - 1;
        if (vec_a) tc_gemm_kernel<true><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(g);
        else tc_gemm_kernel<false><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(g);
        GG_LAUNCHED();
    }
    return GG_OK;
}

}  // extern "C"
