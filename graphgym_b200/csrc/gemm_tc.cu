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
// Structure of one CTA (128 x 128 output tile, 256 threads, 2 CTAs per SM, 128 TMEM columns each):
//   * the K-segments (X W, then the centre rows' X W_id with the row multiplicity as scale; tiles with
//     no centre row skip it) are walked in slabs of 16 k;
//   * A slab: each thread loads 2 x 16 B of its row, applies the scale, splits hi/lo in registers and
//     stores both into shared memory in the UMMA canonical K-major no-swizzle layout
//     (16-byte chunks, chunk-major: address = chunk * 2048 + row * 16, i.e. LBO = 2048 B, SBO = 128 B);
//   * B slab: pre-split once per call by b_image_kernel into exactly that layout; one cp.async.bulk
//     (TMA) of 16 KB per slab, one slab ahead, completing on the stage's b_full mbarrier;
//   * fence.proxy.async + __syncthreads, then ONE thread issues the 6 tcgen05.mma of the slab and a
//     tcgen05.commit onto the stage's mma_done mbarrier; three stages, A rows four slabs ahead in registers;
//   * epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> shared memory -> bias / ReLU / mask ->
//     coalesced 512-byte row stores.
#include <cuda.h>

#include "common.cuh"

namespace gg {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 16;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;        // 8192: one operand slab (hi or lo)
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;       // A_hi, A_lo, B_hi, B_lo
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

__device__ __forceinline__ void tc_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tc_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
// RULE for releasing a shared-memory slot that this thread has read with ld.shared (the producer refills it by TMA or
// cp.async): arrive on the slot's barrier only AFTER an instruction that consumes the loaded registers has issued
// (the tcgen05.st / st.shared that carries the data on).  An arrive placed right after the loads is NOT ordered behind
// their execution: the refill landed under loads still in flight and whole warps' rows of the persistent kernels came out
// wrong (0.3 % of the rows of a 2.45 M-row product with a 4-slot ring; found by profiles/probes/gemm_stress.py, pinned
// down with profiles/probes/gemm_diag.py).  A fake register dependency in inline PTX does not help: ptxas folds it away.
__device__ __forceinline__ void tc_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     tc_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(tc_smem_u32(bar))
                 : "memory");
}
// 16-byte / 4-byte asynchronous global -> shared copies (LDGSTS); src_bytes < size zero-fills the rest
__device__ __forceinline__ void tc_cp_async16(void* dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tc_smem_u32(dst)), "l"(src), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void tc_cp_async4(void* dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(tc_smem_u32(dst)), "l"(src), "r"(src_bytes)
                 : "memory");
}
// the mbarrier receives one arrival from this thread once all its earlier cp.async have landed
__device__ __forceinline__ void tc_cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
// ring bookkeeping: item i lives in slot i % Z and is the (i / Z)-th use of that slot
__device__ __forceinline__ uint32_t tc_full_parity(int i, int z) { return (uint32_t)(i / z) & 1u; }
__device__ __forceinline__ uint32_t tc_free_parity(int i, int z) { return (uint32_t)(i / z - 1) & 1u; }

// Warp roles.  A memory fence waits for the thread's outstanding loads, so the threads that fence
// (fence.proxy.async, needed between their st.shared and the tensor core's reads) must not be the ones
// that prefetch from global memory — otherwise every slab pays a DRAM round trip (measured: 1.45 us/slab).
//   warps 0-3  loaders    cp.async of the raw fp32 slab into a staging ring; never wait on data
//   warps 4-7  converters staging ring -> scale / hi-lo split -> operand stage, fence.proxy.async
//   warp  8    issuer     TMA bulk copies of the B image, tcgen05.mma, tcgen05.commit
constexpr int TC_ROLE_THREADS = 128;
constexpr int TC_BLOCK = 2 * TC_ROLE_THREADS + 32;

// ---- NN kernel shared-memory plan: raw ring (5 x 8 KB) | A_lo stages (2 x 8 KB) | B ring (3 x 16 KB)
// The raw fp32 slab IS the hi operand: the loaders already write it in the canonical layout and the tensor
// core reads only the TF32 bits of every word (truncation), so hi = trunc_tf32(x) costs nothing.  The
// converters only produce lo = x - trunc_tf32(x) (exact in fp32; <= 13 significant bits, of which the
// hardware keeps the top 11): |x - hi - lo_tf32| <= 2^-21 |x|, biased towards zero by ~2^-22 on average —
// two orders of magnitude inside the 1e-5 parity band.  A scaled (ID) segment is scaled in place first.
#ifndef GG_TC_CTAS_PER_SM
#define GG_TC_CTAS_PER_SM 3
#endif
#if GG_TC_CTAS_PER_SM == 3
constexpr int NN_RAW_SLOTS = 3, NN_A_STAGES = 2, NN_B_SLOTS = 2;   // 72 KB: three CTAs per SM
#else
constexpr int NN_RAW_SLOTS = 5, NN_A_STAGES = 2, NN_B_SLOTS = 3;   // 104 KB: two CTAs per SM
#endif
constexpr int NN_RAW_BYTES = TC_TILE_BYTES;
constexpr int NN_OFF_A = NN_RAW_SLOTS * NN_RAW_BYTES;
constexpr int NN_OFF_B = NN_OFF_A + NN_A_STAGES * TC_TILE_BYTES;
constexpr int NN_RING_BYTES = NN_OFF_B + NN_B_SLOTS * 2 * TC_TILE_BYTES;
constexpr int TC_EPI_STRIDE = TC_BN + 4;
constexpr int NN_EPI_BYTES = TC_BM * TC_EPI_STRIDE * 4;  // the epilogue tile reuses the rings
constexpr int NN_SMEM_BYTES = NN_RING_BYTES > NN_EPI_BYTES ? NN_RING_BYTES : NN_EPI_BYTES;

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

template <bool VEC_A>
__global__ void __launch_bounds__(TC_BLOCK, GG_TC_CTAS_PER_SM) tc_gemm_kernel(TcArgs g) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t raw_full[NN_RAW_SLOTS], raw_empty[NN_RAW_SLOTS];
    __shared__ __align__(8) uint64_t op_full[NN_A_STAGES], a_free[NN_A_STAGES];
    __shared__ __align__(8) uint64_t b_full[NN_B_SLOTS], b_free[NN_B_SLOTS];
    __shared__ __align__(8) uint64_t all_done;
    __shared__ uint32_t tmem_base_smem;
    __shared__ int s_slabs[GG_GEMM_MAX_SEGMENTS + 1];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t row0 = (int64_t)blockIdx.x * TC_BM;
    const int nt = blockIdx.y;

    if (tid == 0) {
        for (int i = 0; i < NN_RAW_SLOTS; ++i) {
            tc_mbar_init(&raw_full[i], TC_ROLE_THREADS);
            tc_mbar_init(&raw_empty[i], 1);
        }
        for (int i = 0; i < NN_A_STAGES; ++i) {
            tc_mbar_init(&op_full[i], TC_ROLE_THREADS);
            tc_mbar_init(&a_free[i], 1);
        }
        for (int i = 0; i < NN_B_SLOTS; ++i) {
            tc_mbar_init(&b_full[i], 1);
            tc_mbar_init(&b_free[i], 1);
        }
        tc_mbar_init(&all_done, 1);
        s_slabs[GG_GEMM_MAX_SEGMENTS] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // TMEM: 128 fp32 accumulator columns for this CTA
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         tc_smem_u32(&tmem_base_smem)),
                     "r"(TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // active segments of this tile: an ID segment whose rows all have scale 0 (no centre) is skipped
    const int my_row = tid & (TC_ROLE_THREADS - 1);
    const int64_t grow = row0 + my_row;
    const bool row_ok = grow < g.n;
    float seg_scale[GG_GEMM_MAX_SEGMENTS];
#pragma unroll
    for (int sg = 0; sg < GG_GEMM_MAX_SEGMENTS; ++sg) {
        seg_scale[sg] = 1.f;
        int slabs = 0;
        if (sg < g.num_segments && g.seg[sg].k_slabs > 0) {
            bool on = true;
            if (g.seg[sg].scale) {
                seg_scale[sg] = (row_ok && tid < 2 * TC_ROLE_THREADS) ? __ldg(g.seg[sg].scale + grow) : 0.f;
                on = __syncthreads_or(seg_scale[sg] != 0.f);
            }
            if (on) slabs = g.seg[sg].k_slabs;
        }
        if (tid == 0) s_slabs[sg] = slabs;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;
    int total = 0;
#pragma unroll
    for (int sg = 0; sg < GG_GEMM_MAX_SEGMENTS; ++sg) total += s_slabs[sg];
    // walking the flattened slab sequence: (sg, ks) = segment and slab inside it; s_slabs[MAX] = 0 ends the skip
    int sg = 0, ks = 0, seg_left = s_slabs[0];
    auto skip_empty = [&]() {
        while (seg_left == 0 && sg < GG_GEMM_MAX_SEGMENTS - 1) seg_left = s_slabs[++sg];
    };
    auto advance = [&]() {
        ++ks;
        if (--seg_left == 0) { ks = 0; skip_empty(); }
    };
    skip_empty();

    if (warp < 4) {
        // ===== loaders: this thread's row, 64 B per slab, straight into the raw ring (= the hi operand) =====
        int slot = 0, round = 0;
        for (int i = 0; i < total; ++i) {
            if (round > 0) tc_mbar_wait(&raw_empty[slot], (uint32_t)(round - 1) & 1u);
            const TcSegment& s = g.seg[sg];
            const float* arow = s.a + (row_ok ? grow : 0) * s.lda;
            uint8_t* dst = smem + slot * NN_RAW_BYTES + my_row * 16;
            const int k0 = ks * TC_BK;
            if (VEC_A && row_ok && k0 + TC_BK <= s.k) {  // interior slab: four full 16-byte copies
#pragma unroll
                for (int c = 0; c < TC_BK / 4; ++c) tc_cp_async16(dst + c * TC_LBO, arow + k0 + c * 4, 16u);
            } else {
#pragma unroll
                for (int c = 0; c < TC_BK / 4; ++c) {
                    const int k = k0 + c * 4;
                    if (VEC_A) {
                        const int left = row_ok ? (s.k - k) * 4 : 0;
                        tc_cp_async16(dst + c * TC_LBO, arow + (k < s.k ? k : 0), left >= 16 ? 16u : (left > 0 ? (uint32_t)left : 0u));
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const bool ok = row_ok && k + e < s.k;
                            tc_cp_async4(dst + c * TC_LBO + e * 4, arow + (ok ? k + e : 0), ok ? 4u : 0u);
                        }
                    }
                }
            }
            tc_cp_async_arrive(&raw_full[slot]);
            if (++slot == NN_RAW_SLOTS) { slot = 0; ++round; }
            advance();
        }
    } else if (warp < 8) {
        // ===== converters: lo = x - trunc_tf32(x) into the lo stage (an ID segment scales the slab in place first) =====
        int slot = 0, round = 0;
        for (int i = 0; i < total; ++i) {
            const int stage = i & (NN_A_STAGES - 1);
            const float sc = seg_scale[0] * (sg == 0) + seg_scale[1] * (sg == 1) + seg_scale[2] * (sg == 2) +
                             seg_scale[3] * (sg == 3);
            const bool scaled = g.seg[sg].scale != nullptr;
            tc_mbar_wait(&raw_full[slot], (uint32_t)round & 1u);
            uint8_t* raw = smem + slot * NN_RAW_BYTES + my_row * 16;
            float4 v[TC_BK / 4];
#pragma unroll
            for (int c = 0; c < TC_BK / 4; ++c) v[c] = *reinterpret_cast<const float4*>(raw + c * TC_LBO);
            if (scaled) {
#pragma unroll
                for (int c = 0; c < TC_BK / 4; ++c) {
                    v[c].x *= sc; v[c].y *= sc; v[c].z *= sc; v[c].w *= sc;
                    *reinterpret_cast<float4*>(raw + c * TC_LBO) = v[c];
                }
            }
            if (i >= NN_A_STAGES) tc_mbar_wait(&a_free[stage], (uint32_t)(i / NN_A_STAGES - 1) & 1u);
            uint8_t* st = smem + NN_OFF_A + stage * TC_TILE_BYTES + my_row * 16;
#pragma unroll
            for (int c = 0; c < TC_BK / 4; ++c) {
                float4 l;
                l.x = v[c].x - trunc_tf32(v[c].x);
                l.y = v[c].y - trunc_tf32(v[c].y);
                l.z = v[c].z - trunc_tf32(v[c].z);
                l.w = v[c].w - trunc_tf32(v[c].w);
                *reinterpret_cast<float4*>(st + c * TC_LBO) = l;
            }
            // generic-proxy writes (the loaders' cp.async data this thread has observed, and the stores above)
            // -> visible to the tensor core's async-proxy reads
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_arrive(&op_full[stage]);
            if (++slot == NN_RAW_SLOTS) { slot = 0; ++round; }
            advance();
        }
    } else if (lane == 0) {
        // ===== issuer: B image slabs by TMA (two ahead), then the six MMAs of every slab =====
        int bsg = sg, bks = ks, bleft = seg_left;  // the B prefetch runs two slabs ahead with its own cursor
        auto issue_b = [&](int i) {
            const TcSegment& s = g.seg[bsg];
            const float* src = s.b_image + ((int64_t)nt * s.k_slabs + bks) * (2 * TC_TILE_BYTES / 4);
            const int slot = i % NN_B_SLOTS;
            tc_expect_tx(&b_full[slot], 2 * TC_TILE_BYTES);
            tc_bulk_g2s(smem + NN_OFF_B + slot * 2 * TC_TILE_BYTES, src, 2 * TC_TILE_BYTES, &b_full[slot]);
            ++bks;
            if (--bleft == 0) {
                bks = 0;
                while (bleft == 0 && bsg < GG_GEMM_MAX_SEGMENTS - 1) bleft = s_slabs[++bsg];
            }
        };
        constexpr int kAhead = NN_B_SLOTS - 1;  // B slabs in flight ahead of the MMA
        for (int i = 0; i < kAhead && i < total; ++i) issue_b(i);
        int rslot = 0;
        for (int i = 0; i < total; ++i) {
            const int stage = i & (NN_A_STAGES - 1), slot = i % NN_B_SLOTS;
            if (i + kAhead < total) {  // slot (i+kAhead) % NN_B_SLOTS was last read by slab i-1
                if (i + kAhead >= NN_B_SLOTS)
                    tc_mbar_wait(&b_free[(i + kAhead) % NN_B_SLOTS], tc_free_parity(i + kAhead, NN_B_SLOTS));
                issue_b(i + kAhead);
            }
            tc_mbar_wait(&op_full[stage], tc_full_parity(i, NN_A_STAGES));
            tc_mbar_wait(&b_full[slot], tc_full_parity(i, NN_B_SLOTS));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = tc_smem_u32(smem + rslot * NN_RAW_BYTES);
            const uint32_t a_lo = tc_smem_u32(smem + NN_OFF_A + stage * TC_TILE_BYTES);
            const uint32_t b_hi = tc_smem_u32(smem + NN_OFF_B + slot * 2 * TC_TILE_BYTES), b_lo = b_hi + TC_TILE_BYTES;
#pragma unroll
            for (int j = 0; j < TC_BK / 8; ++j) {  // one MMA consumes 8 k = two 16-byte chunks
                const uint32_t ko = j * 2 * TC_LBO;
                tc_mma(tmem_base, tc_smem_desc(a_hi + ko), tc_smem_desc(b_hi + ko), (i > 0 || j > 0) ? 1u : 0u);
                tc_mma(tmem_base, tc_smem_desc(a_hi + ko), tc_smem_desc(b_lo + ko), 1u);
                tc_mma(tmem_base, tc_smem_desc(a_lo + ko), tc_smem_desc(b_hi + ko), 1u);
            }
            tc_commit(&raw_empty[rslot]);
            tc_commit(&a_free[stage]);
            tc_commit(&b_free[slot]);
            if (++rslot == NN_RAW_SLOTS) rslot = 0;
        }
        tc_commit(&all_done);  // completes when every MMA issued above has completed
    }
    // ---- epilogue: TMEM -> registers -> shared (row stride 132 floats) -> coalesced 512-byte rows ----
    tc_mbar_wait(&all_done, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();  // every role is done with the rings
    float* tile = reinterpret_cast<float*>(smem);
    if (warp < 8) {
        const int q = warp & 3;      // TMEM lane quarter this warp may read
        const int half = warp >> 2;  // column half
        const int trow = q * 32 + lane;
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            const int cbase = half * 64 + part * 32;
            uint32_t acc[32];
            if (total > 0) {
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cbase, acc);
            } else {
#pragma unroll
                for (int e = 0; e < 32; ++e) acc[e] = 0u;
            }
#pragma unroll
            for (int e = 0; e < 32; e += 4)
                *reinterpret_cast<float4*>(tile + trow * TC_EPI_STRIDE + cbase + e) =
                    make_float4(__uint_as_float(acc[e]), __uint_as_float(acc[e + 1]), __uint_as_float(acc[e + 2]),
                                __uint_as_float(acc[e + 3]));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS)
                     : "memory");
    }
    if (warp >= 8) return;
    const bool vec_out = (g.f % 4 == 0) && (g.ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.out) & 15) == 0) &&
                         (!g.relu_mask || (g.ld_mask % 4 == 0 && (reinterpret_cast<uintptr_t>(g.relu_mask) & 15) == 0)) &&
                         (!g.bias || (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0);
    const int c = nt * TC_BN + lane * 4;  // this lane's 4 columns
    const bool plain = vec_out && !g.accumulate_out &&
                       (!g.final_chunk || (!g.bias && g.act != GG_ACT_RELU && !g.relu_mask));
    if (plain) {  // nothing to fuse: per row one 128-bit shared load and one 128-bit coalesced store per lane
        if (c + 3 < g.f) {
            const int rows = (int)(g.n - row0 < TC_BM ? g.n - row0 : TC_BM);
            float* orow = g.out + (row0 + warp) * g.ldo + c;
#pragma unroll 4
            for (int rr = warp; rr < rows; rr += 8, orow += 8 * g.ldo)
                *reinterpret_cast<float4*>(orow) = *reinterpret_cast<const float4*>(tile + rr * TC_EPI_STRIDE + lane * 4);
        }
        return;
    }
    for (int rr = warp; rr < TC_BM; rr += 8) {
        const int64_t r = row0 + rr;
        if (r >= g.n) break;
        const float4 a4 = *reinterpret_cast<const float4*>(tile + rr * TC_EPI_STRIDE + lane * 4);
        float v[4] = {a4.x, a4.y, a4.z, a4.w};
        if (vec_out && c + 3 < g.f) {
            if (g.accumulate_out) {
                const float4 o = *reinterpret_cast<const float4*>(g.out + r * g.ldo + c);
                v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;  // fp32 round-to-nearest add of partials
            }
            if (g.final_chunk) {
                if (g.bias) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + c));
                    v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
                }
                if (g.act == GG_ACT_RELU) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
                }
                if (g.relu_mask) {
                    const float4 m = __ldg(reinterpret_cast<const float4*>(g.relu_mask + r * g.ld_mask + c));
                    v[0] = m.x > 0.f ? v[0] : 0.f; v[1] = m.y > 0.f ? v[1] : 0.f;
                    v[2] = m.z > 0.f ? v[2] : 0.f; v[3] = m.w > 0.f ? v[3] : 0.f;
                }
            }
            *reinterpret_cast<float4*>(g.out + r * g.ldo + c) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (c + e >= g.f) continue;
                float x = v[e];
                if (g.accumulate_out) x += g.out[r * g.ldo + c + e];
                if (g.final_chunk) {
                    if (g.bias) x += __ldg(g.bias + c + e);
                    if (g.act == GG_ACT_RELU) x = fmaxf(x, 0.f);
                    if (g.relu_mask) x = __ldg(g.relu_mask + r * g.ld_mask + c + e) > 0.f ? x : 0.f;
                }
                g.out[r * g.ldo + c + e] = x;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Persistent NN kernel for the layer-sized transforms (one unscaled segment, K <= 128, F <= 128: X W and dH W^T of
// gcnconv / gatconv / ginconv).  The per-tile kernel above is a chain of dependent latencies per CTA (barrier and
// TMEM set-up, first slab from DRAM, seven slabs through a three-slot ring, accumulator drain, stores) with at most
// three CTAs per SM to overlap them: 0.82 ms for 2.45 M x 100 -> 128 = 0.41 of the HBM copy peak, and every CTA
// re-reads the 112 KB B image from L2.  Here one CTA per SM walks row tiles without ever draining its pipeline:
//   warp 0     producer   the B image ONCE (resident in shared memory), then A slabs of 128 rows x 32 k by TMA
//                         (2-D tensor map, SWIZZLE_128B) into a ring that stays full across tile boundaries
//   warps 2-5  converters thread = row: the slab row into registers (the ring slot is free again at once), then
//                         hi = x and lo = x - trunc_tf32(x) into TENSOR MEMORY (tcgen05.st, lane = row, column = k):
//                         the A operand never goes back to shared memory
//   warp 1     issuer     tcgen05.mma with A from TMEM and B from the resident image, into one of TWO 128-column
//                         accumulators (the tensor core reads the TF32 bits of hi)
//   warps 6-9  epilogue   TMEM -> registers (bias / ReLU) -> per-warp 4 KB staging -> ONE TMA store of 32 rows x 32
//                         columns (the tensor map clips rows >= n and columns >= f), overlapped with the next tile's
//                         MMAs through the second accumulator; with a ReLU mask: staged, coalesced 128-byte row pieces
// Shared-memory traffic decides the design: an MMA of 8 k with both operands in shared memory reads 8 KB in its
// 64 cycles — the whole 128 B/clk of the SM — so the first version (A hi = raw slab, A lo = a second shared slab:
// 0.56 ms, 0.61 of the copy peak) had no bandwidth left for the TMA writes, the converters and the epilogue staging.
// With A in TMEM the MMAs read 4 KB each, the lo slabs disappear (the ring grows from 4 to 6 slots) and a ring slot
// is recycled after the converters' read instead of after the MMAs (0.52 ms).  ncu then showed the issuer waiting
// for a free accumulator and the epilogue warps in a chain of shared loads, parameter loads and predicated row stores:
// the stores went to TMA.  TMEM: 2 x 128 accumulator columns + 4 operand stages x (32 hi + 32 lo) columns = all 512.
// ---------------------------------------------------------------------------------------------
constexpr int PS_BK = 32;                              // k per A slab (one 128-byte swizzle row)
constexpr int PS_SLAB_BYTES = TC_BM * PS_BK * 4;       // 16384
constexpr int PS_A_STAGES = 4;                         // TMEM operand stages of 64 columns (hi | lo)
constexpr int PS_MAX_RAW = 8;
constexpr int PS_STAGING_BYTES = 4 * 32 * 32 * 4;      // four epilogue warps x 32 rows x 32 columns
constexpr int PS_THREADS = 10 * 32;
constexpr int PS_SMEM_MAX = 232448 - 2048;             // dynamic budget (static barriers and the base alignment aside)

struct PsArgs {
    const float* b_image;   // [k_slabs16][2][TC_TILE_BYTES / 4] of N-tile 0
    int k;                  // reduction length (multiple of 4)
    int k_slabs16;          // 16-k slabs of the B image
    int raw_slots;
    int64_t n;
    int f;
    const float* bias;
    int act;
    const float* relu_mask;
    int64_t ld_mask;
    float* out;
    int64_t ldo;
};

// K-major SWIZZLE_128B descriptor: 8-row groups 1024 B apart, layout type 2; the k-step inside the 128-byte row is
// selected by advancing the start address (32 B per 8 tf32)
__device__ __forceinline__ uint64_t ps_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
__device__ __forceinline__ void ps_tma_load_2d(void* dst, const void* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            tc_smem_u32(dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(tc_smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(TC_IDESC), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

__device__ __forceinline__ void ps_tma_store_2d(const void* tmap, int c0, int c1, const void* src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(c0), "r"(c1),
                 "r"(tc_smem_u32(src))
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

template <bool MASKED>
__global__ void __launch_bounds__(PS_THREADS, 1) tc_gemm_persist_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                        const __grid_constant__ CUtensorMap tmap_out,
                                                                        const PsArgs g) {
    extern __shared__ uint8_t ps_smem_raw[];
    __shared__ __align__(8) uint64_t raw_full[PS_MAX_RAW], raw_free[PS_MAX_RAW];
    __shared__ __align__(8) uint64_t a_full[PS_A_STAGES], a_free[PS_A_STAGES];
    __shared__ __align__(8) uint64_t acc_full[2], acc_free[2];
    __shared__ __align__(8) uint64_t b_ready;
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* smem = ps_smem_raw + ((1024u - (tc_smem_u32(ps_smem_raw) & 1023u)) & 1023u);
    const int R = g.raw_slots;
    uint8_t* s_raw = smem;
    uint8_t* s_b = s_raw + R * PS_SLAB_BYTES;
    uint8_t* s_stage = s_b + g.k_slabs16 * 2 * TC_TILE_BYTES;

    if (tid == 0) {
        for (int i = 0; i < PS_MAX_RAW; ++i) {
            tc_mbar_init(&raw_full[i], 1);
            tc_mbar_init(&raw_free[i], 128);
        }
        for (int i = 0; i < PS_A_STAGES; ++i) {
            tc_mbar_init(&a_full[i], 128);
            tc_mbar_init(&a_free[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc_mbar_init(&acc_full[i], 1);
            tc_mbar_init(&acc_free[i], 128);
        }
        tc_mbar_init(&b_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // all 512 columns: two accumulators + four operand stages
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         tc_smem_u32(&tmem_base_smem)),
                     "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    const int64_t tiles = (g.n + TC_BM - 1) / TC_BM;
    const int slabs = (g.k + PS_BK - 1) / PS_BK;     // A slabs per tile
    const int ksteps = (g.k + 7) / 8;                // MMAs of 8 k per operand pair and tile

    if (warp == 0) {
        if (lane == 0) {
            // ===== producer =====
            const uint32_t b_bytes = (uint32_t)g.k_slabs16 * 2u * TC_TILE_BYTES;
            tc_expect_tx(&b_ready, b_bytes);
            for (int s = 0; s < g.k_slabs16; ++s)
                tc_bulk_g2s(s_b + s * 2 * TC_TILE_BYTES, g.b_image + (int64_t)s * (2 * TC_TILE_BYTES / 4), 2 * TC_TILE_BYTES,
                            &b_ready);
            int item = 0;
            for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
                for (int s = 0; s < slabs; ++s, ++item) {
                    const int slot = item % R, use = item / R;
                    if (use > 0) tc_mbar_wait(&raw_free[slot], (uint32_t)(use - 1) & 1u);
                    tc_expect_tx(&raw_full[slot], PS_SLAB_BYTES);
                    ps_tma_load_2d(s_raw + slot * PS_SLAB_BYTES, &tmap_a, s * PS_BK, (int)(t * TC_BM), &raw_full[slot]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== issuer =====
            tc_mbar_wait(&b_ready, 0u);
            int item = 0, it = 0;
            for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
                const int buf = it & 1, buse = it >> 1;
                if (buse > 0) tc_mbar_wait(&acc_free[buf], (uint32_t)(buse - 1) & 1u);
                const uint32_t d_tmem = tmem_base + (uint32_t)buf * TC_TMEM_COLS;
                for (int s = 0; s < slabs; ++s, ++item) {
                    const int stage = item % PS_A_STAGES;
                    tc_mbar_wait(&a_full[stage], (uint32_t)(item / PS_A_STAGES) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = tmem_base + 2 * TC_TMEM_COLS + (uint32_t)stage * 64u, a_lo = a_hi + 32u;
#pragma unroll
                    for (int j = 0; j < PS_BK / 8; ++j) {
                        const int gs = s * (PS_BK / 8) + j;          // k-step of the tile
                        if (gs < ksteps) {
                            const uint32_t b_hi = tc_smem_u32(s_b) + (uint32_t)(gs >> 1) * (2 * TC_TILE_BYTES) +
                                                  (uint32_t)(gs & 1) * 2 * TC_LBO;
                            const uint32_t b_lo = b_hi + TC_TILE_BYTES;
                            tc_mma_ts(d_tmem, a_hi + j * 8, tc_smem_desc(b_hi), gs > 0 ? 1u : 0u);
                            tc_mma_ts(d_tmem, a_hi + j * 8, tc_smem_desc(b_lo), 1u);
                            tc_mma_ts(d_tmem, a_lo + j * 8, tc_smem_desc(b_hi), 1u);
                        }
                    }
                    tc_commit(&a_free[stage]);
                }
                tc_commit(&acc_full[buf]);
            }
        }
    } else if (warp < 6) {
        // ===== converters: thread = row = TMEM lane (a warp reaches the lanes of its quarter, warp % 4) =====
        const int q = warp & 3, row = q * 32 + lane;
        const uint32_t t_lane = tmem_base + 2 * TC_TMEM_COLS + ((uint32_t)(q * 32) << 16);
        int item = 0;
        for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
            for (int s = 0; s < slabs; ++s, ++item) {
                const int slot = item % R, stage = item % PS_A_STAGES;
                tc_mbar_wait(&raw_full[slot], (uint32_t)(item / R) & 1u);
                const uint8_t* src = s_raw + slot * PS_SLAB_BYTES + row * 128;
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {   // logical 16-byte chunk c of the row sits at chunk c ^ (row % 8)
                    const float4 v = *reinterpret_cast<const float4*>(src + ((c ^ (row & 7)) * 16));
                    hi[4 * c] = __float_as_uint(v.x); hi[4 * c + 1] = __float_as_uint(v.y);
                    hi[4 * c + 2] = __float_as_uint(v.z); hi[4 * c + 3] = __float_as_uint(v.w);
                }
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const float x = __uint_as_float(hi[e]);
                    lo[e] = __float_as_uint(x - trunc_tf32(x));
                }
                if (item >= PS_A_STAGES) {
                    tc_mbar_wait(&a_free[stage], (uint32_t)(item / PS_A_STAGES - 1) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                tc_st32(t_lane + (uint32_t)stage * 64u, hi);
                tc_st32(t_lane + (uint32_t)stage * 64u + 32u, lo);
                tc_arrive(&raw_free[slot]);      // the stores have consumed the loaded registers (see the RULE above)
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                tc_arrive(&a_full[stage]);
            }
        }
    } else {
        // ===== epilogue: warp q of the four reads TMEM lanes [32 q, 32 q + 32) =====
        const int q = warp & 3;
        uint8_t* stg = s_stage + (warp - 6) * (32 * 32 * 4);
        int it = 0;
        for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
            const int buf = it & 1;
            tc_mbar_wait(&acc_full[buf], (uint32_t)(it >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + (uint32_t)buf * TC_TMEM_COLS + ((uint32_t)(q * 32) << 16);
            const int64_t row_base = t * TC_BM + q * 32;
#pragma unroll 1
            for (int part = 0; part < 4; ++part) {
                if (part * 32 >= g.f) {   // columns past f: nothing to store (warp-uniform)
                    if (part == 3) {
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        tc_arrive(&acc_free[buf]);
                    }
                    continue;
                }
                uint32_t acc[32];
                tc_ld32(taddr + (uint32_t)(part * 32), acc);
                if (part == 3 || (part + 1) * 32 >= g.f) {   // last read of this accumulator: hand it back to the issuer
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    if (part == 3) tc_arrive(&acc_free[buf]);
                }
                if constexpr (!MASKED) {
                    // this thread's row, 32 columns: bias / ReLU in registers, then the swizzled staging row
                    if (g.bias) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (part * 32 + e * 4 < g.f) b4 = __ldg(reinterpret_cast<const float4*>(g.bias + part * 32 + e * 4));
                            acc[4 * e] = __float_as_uint(__uint_as_float(acc[4 * e]) + b4.x);
                            acc[4 * e + 1] = __float_as_uint(__uint_as_float(acc[4 * e + 1]) + b4.y);
                            acc[4 * e + 2] = __float_as_uint(__uint_as_float(acc[4 * e + 2]) + b4.z);
                            acc[4 * e + 3] = __float_as_uint(__uint_as_float(acc[4 * e + 3]) + b4.w);
                        }
                    }
                    if (g.act == GG_ACT_RELU) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) acc[e] = __float_as_uint(fmaxf(__uint_as_float(acc[e]), 0.f));
                    }
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous store has read the staging
                    __syncwarp();
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        *reinterpret_cast<float4*>(stg + lane * 128 + ((e ^ (lane & 7)) * 16)) =
                            make_float4(__uint_as_float(acc[4 * e]), __uint_as_float(acc[4 * e + 1]),
                                        __uint_as_float(acc[4 * e + 2]), __uint_as_float(acc[4 * e + 3]));
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) ps_tma_store_2d(&tmap_out, part * 32, (int)row_base, stg);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        *reinterpret_cast<float4*>(stg + lane * 128 + ((e ^ (lane & 7)) * 16)) =
                            make_float4(__uint_as_float(acc[4 * e]), __uint_as_float(acc[4 * e + 1]),
                                        __uint_as_float(acc[4 * e + 2]), __uint_as_float(acc[4 * e + 3]));
                    __syncwarp();
                    const int piece = lane & 7;
                    const int c = part * 32 + piece * 4;
                    const bool col_ok = c < g.f;
                    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (g.bias && col_ok) b4 = __ldg(reinterpret_cast<const float4*>(g.bias + c));
                    float4 v[8], m[8];
#pragma unroll
                    for (int r4 = 0; r4 < 8; ++r4) {   // all loads first: eight independent shared and mask loads in flight
                        const int rr = r4 * 4 + (lane >> 3);
                        const int64_t r = row_base + rr;
                        v[r4] = *reinterpret_cast<const float4*>(stg + rr * 128 + ((piece ^ (rr & 7)) * 16));
                        m[r4] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (r < g.n && col_ok) m[r4] = __ldg(reinterpret_cast<const float4*>(g.relu_mask + r * g.ld_mask + c));
                    }
#pragma unroll
                    for (int r4 = 0; r4 < 8; ++r4) {
                        const int64_t r = row_base + r4 * 4 + (lane >> 3);
                        float4 x = v[r4];
                        x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
                        if (g.act == GG_ACT_RELU) {
                            x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
                        }
                        x.x = m[r4].x > 0.f ? x.x : 0.f; x.y = m[r4].y > 0.f ? x.y : 0.f;
                        x.z = m[r4].z > 0.f ? x.z : 0.f; x.w = m[r4].w > 0.f ? x.w : 0.f;
                        if (r < g.n && col_ok) *reinterpret_cast<float4*>(g.out + r * g.ldo + c) = x;
                    }
                    __syncwarp();
                }
            }
        }
        if (!MASKED && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before exit
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// Weight gradient on the tensor cores:  partial[split][K,F] = sum_{r in split} A[row(r),:]^T G[row(r),:]
// (dW = X^T dH; dW_id = X[id]^T dH[id]).  M = K, N = F, the reduction runs over matrix ROWS, so both
// operands are transposed on their way to the operand stage:
//   loaders     cp.async 16 rows x 512 B of A and of G (row segments, coalesced) into the raw ring
//   converters  thread c reads COLUMN c of the raw A slab (16 values, conflict-free), packs 4 consecutive
//               rows per 16-byte k-chunk, splits hi/lo, stores the K-major operand slab; same for G
//   issuer      six tcgen05.mma per slab
// A split covers kTnRowsPerSplit = 256 rows so that the truncating tensor-core accumulation stays
// below 1.6e-6 relative; several splits run back to back in one CTA, their tiles are added in fp32
// (round to nearest) in the CTA's shared accumulator, and the per-CTA partials are reduced in fp64.
// ---------------------------------------------------------------------------------------------
constexpr int kTnRowsPerSplit = 256;
constexpr int TN_RAW_SLOTS = 2, TN_STAGES = 2;
constexpr int TN_RAW_BYTES = 2 * TC_BK * TC_BM * 4;                   // raw A (16 x 128) + raw G (16 x 128)
constexpr int TN_OFF_OP = TN_RAW_SLOTS * TN_RAW_BYTES;                // 32 KB
constexpr int TN_SMEM_BYTES = TN_OFF_OP + TN_STAGES * TC_STAGE_BYTES;  // 96 KB
static_assert(TC_BM * TC_EPI_STRIDE * 4 <= TN_SMEM_BYTES, "epilogue tile must fit");

struct TnTcArgs {
    const float* a;
    int64_t lda;
    const float* g;
    int64_t ldg;
    const int64_t* row_index;
    int64_t n;
    int k, f;
    int64_t rows_per_cta;  // multiple of kTnRowsPerSplit
    float* partial;        // [gridDim.y][k][f]
};

template <bool VEC>
__global__ void __launch_bounds__(TC_BLOCK, 2) tc_gemm_tn_kernel(TnTcArgs g) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t raw_full[TN_RAW_SLOTS], raw_empty[TN_RAW_SLOTS];
    __shared__ __align__(8) uint64_t op_full[TN_STAGES], op_free[TN_STAGES];
    __shared__ __align__(8) uint64_t seg_done, tmem_free;
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = (g.f + TC_BN - 1) / TC_BN;
    const int m0 = (blockIdx.x / n_tiles) * TC_BM;  // offset in K (rows of the output)
    const int c0 = (blockIdx.x % n_tiles) * TC_BN;  // offset in F
    const int64_t r_beg = (int64_t)blockIdx.y * g.rows_per_cta;
    const int64_t r_end = r_beg + g.rows_per_cta < g.n ? r_beg + g.rows_per_cta : g.n;
    const int total = r_end > r_beg ? (int)((r_end - r_beg + TC_BK - 1) / TC_BK) : 0;
    constexpr int kSlabsPerSeg = kTnRowsPerSplit / TC_BK;  // 32 slabs, then the accumulator is drained

    if (tid == 0) {
        for (int i = 0; i < TN_RAW_SLOTS; ++i) {
            tc_mbar_init(&raw_full[i], TC_ROLE_THREADS);
            tc_mbar_init(&raw_empty[i], TC_ROLE_THREADS);
        }
        for (int i = 0; i < TN_STAGES; ++i) {
            tc_mbar_init(&op_full[i], TC_ROLE_THREADS);
            tc_mbar_init(&op_free[i], 1);
        }
        tc_mbar_init(&seg_done, 1);
        tc_mbar_init(&tmem_free, 2 * TC_ROLE_THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
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
    const int rt = tid & (TC_ROLE_THREADS - 1);

    // fp32 accumulator of the CTA across its 512-row segments: this thread's share of the 128x128 tile
    // in the epilogue mapping (row = warp-quarter*32 + lane, 64 columns) -- kept in registers
    float racc[64];
#pragma unroll
    for (int e = 0; e < 64; ++e) racc[e] = 0.f;

    // At the end of every 256-row segment warps 0-7 drain the TMEM tile into racc (each warp its lane
    // quarter, loaders the low / converters the high 64 columns) and release the accumulator.  Both
    // roles have finished their part of the segment when they get here, so nothing can deadlock.
    auto drain = [&](int sgi) {
        const int q = warp & 3, half = warp >> 2;
        tc_mbar_wait(&seg_done, (uint32_t)sgi & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            uint32_t acc[32];
            tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64 + part * 32), acc);
#pragma unroll
            for (int e = 0; e < 32; ++e) racc[part * 32 + e] += __uint_as_float(acc[e]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        tc_arrive(&tmem_free);  // the issuer may overwrite the accumulator
    };

    if (warp < 4) {
        // ===== loaders =====
        for (int i = 0; i < total; ++i) {
            const int slot = i % TN_RAW_SLOTS;
            if (i >= TN_RAW_SLOTS) tc_mbar_wait(&raw_empty[slot], tc_free_parity(i, TN_RAW_SLOTS));
            const int64_t r0 = r_beg + (int64_t)i * TC_BK;
            uint8_t* raw_a = smem + slot * TN_RAW_BYTES;
            uint8_t* raw_g = raw_a + TC_BK * TC_BM * 4;
#pragma unroll
            for (int j = 0; j < (TC_BK * TC_BM / 4) / TC_ROLE_THREADS; ++j) {  // 4 x 16 B per operand per thread
                const int q = rt + j * TC_ROLE_THREADS;
                const int kk = q >> 5, cq = q & 31;  // slab row, 16-byte chunk inside the 512-byte row segment
                const int64_t r = r0 + kk;
                const bool rok = r < r_end;
                const int64_t rr = rok ? (g.row_index ? __ldg(g.row_index + r) : r) : 0;
                const int ca = m0 + cq * 4, cg = c0 + cq * 4;
                if (VEC) {
                    const int la = rok ? (g.k - ca) * 4 : 0, lg = rok ? (g.f - cg) * 4 : 0;
                    tc_cp_async16(raw_a + q * 16, g.a + rr * g.lda + (ca < g.k ? ca : 0),
                                  la >= 16 ? 16u : (la > 0 ? (uint32_t)la : 0u));
                    tc_cp_async16(raw_g + q * 16, g.g + rr * g.ldg + (cg < g.f ? cg : 0),
                                  lg >= 16 ? 16u : (lg > 0 ? (uint32_t)lg : 0u));
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const bool oka = rok && ca + e < g.k, okg = rok && cg + e < g.f;
                        tc_cp_async4(raw_a + q * 16 + e * 4, g.a + rr * g.lda + (oka ? ca + e : 0), oka ? 4u : 0u);
                        tc_cp_async4(raw_g + q * 16 + e * 4, g.g + rr * g.ldg + (okg ? cg + e : 0), okg ? 4u : 0u);
                    }
                }
            }
            tc_cp_async_arrive(&raw_full[slot]);
            if (i % kSlabsPerSeg == kSlabsPerSeg - 1 || i == total - 1) drain(i / kSlabsPerSeg);
        }
    } else if (warp < 8) {
        // ===== converters: column rt of raw A and raw G -> K-major hi/lo slabs =====
        for (int i = 0; i < total; ++i) {
            const int slot = i % TN_RAW_SLOTS, stage = i % TN_STAGES;
            tc_mbar_wait(&raw_full[slot], tc_full_parity(i, TN_RAW_SLOTS));
            const float* raw_a = reinterpret_cast<const float*>(smem + slot * TN_RAW_BYTES);
            const float* raw_g = raw_a + TC_BK * TC_BM;
            float va[TC_BK], vg[TC_BK];
#pragma unroll
            for (int kk = 0; kk < TC_BK; ++kk) {
                va[kk] = raw_a[kk * TC_BM + rt];
                vg[kk] = raw_g[kk * TC_BM + rt];
            }
            if (i >= TN_STAGES) tc_mbar_wait(&op_free[stage], tc_free_parity(i, TN_STAGES));
            uint8_t* st = smem + TN_OFF_OP + stage * TC_STAGE_BYTES;
#pragma unroll
            for (int c = 0; c < TC_BK / 4; ++c) {
                float4 h, l;
                const int off = c * (int)TC_LBO + rt * 16;
                split_tf32(va[c * 4 + 0], h.x, l.x);
                split_tf32(va[c * 4 + 1], h.y, l.y);
                split_tf32(va[c * 4 + 2], h.z, l.z);
                split_tf32(va[c * 4 + 3], h.w, l.w);
                *reinterpret_cast<float4*>(st + off) = h;
                *reinterpret_cast<float4*>(st + TC_TILE_BYTES + off) = l;
                split_tf32(vg[c * 4 + 0], h.x, l.x);
                split_tf32(vg[c * 4 + 1], h.y, l.y);
                split_tf32(vg[c * 4 + 2], h.z, l.z);
                split_tf32(vg[c * 4 + 3], h.w, l.w);
                *reinterpret_cast<float4*>(st + 2 * TC_TILE_BYTES + off) = h;
                *reinterpret_cast<float4*>(st + 3 * TC_TILE_BYTES + off) = l;
            }
            tc_arrive(&raw_empty[slot]);         // after the stores that consume the loaded registers (RULE above)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_arrive(&op_full[stage]);
            if (i % kSlabsPerSeg == kSlabsPerSeg - 1 || i == total - 1) drain(i / kSlabsPerSeg);
        }
    } else if (lane == 0) {
        // ===== issuer =====
        for (int i = 0; i < total; ++i) {
            const int stage = i % TN_STAGES;
            const int in_seg = i % kSlabsPerSeg;
            if (in_seg == 0 && i > 0) tc_mbar_wait(&tmem_free, (uint32_t)(i / kSlabsPerSeg - 1) & 1u);  // drained
            tc_mbar_wait(&op_full[stage], tc_full_parity(i, TN_STAGES));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = tc_smem_u32(smem + TN_OFF_OP + stage * TC_STAGE_BYTES), a_lo = a_hi + TC_TILE_BYTES;
            const uint32_t b_hi = a_hi + 2 * TC_TILE_BYTES, b_lo = a_hi + 3 * TC_TILE_BYTES;
#pragma unroll
            for (int j = 0; j < TC_BK / 8; ++j) {
                const uint32_t ko = j * 2 * TC_LBO;
                tc_mma(tmem_base, tc_smem_desc(a_hi + ko), tc_smem_desc(b_hi + ko), (in_seg > 0 || j > 0) ? 1u : 0u);
                tc_mma(tmem_base, tc_smem_desc(a_hi + ko), tc_smem_desc(b_lo + ko), 1u);
                tc_mma(tmem_base, tc_smem_desc(a_lo + ko), tc_smem_desc(b_hi + ko), 1u);
            }
            tc_commit(&op_free[stage]);
            if (in_seg == kSlabsPerSeg - 1 || i == total - 1) tc_commit(&seg_done);  // this 512-row segment is in TMEM
        }
    }
    __syncthreads();  // all roles done with the rings
    float* tile = reinterpret_cast<float*>(smem);
    if (warp < 8) {
        const int q = warp & 3, half = warp >> 2;
        const int tr = q * 32 + lane;
#pragma unroll
        for (int e = 0; e < 64; e += 4)
            *reinterpret_cast<float4*>(tile + tr * TC_EPI_STRIDE + half * 64 + e) =
                make_float4(racc[e], racc[e + 1], racc[e + 2], racc[e + 3]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS)
                     : "memory");
    }
    if (warp >= 8) return;
    float* dst = g.partial + (int64_t)blockIdx.y * g.k * g.f;
    const int c = c0 + lane * 4;
    const bool vec = (g.f % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.partial) & 15) == 0);
    for (int rr = warp; rr < TC_BM; rr += 8) {
        const int m = m0 + rr;
        if (m >= g.k) break;
        const float4 a4 = *reinterpret_cast<const float4*>(tile + rr * TC_EPI_STRIDE + lane * 4);
        if (vec && c + 3 < g.f) {
            *reinterpret_cast<float4*>(dst + (int64_t)m * g.f + c) = a4;
        } else {
            const float v[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (c + e < g.f) dst[(int64_t)m * g.f + c + e] = v[e];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Persistent weight-gradient kernel (dW = X^T dH for K <= 128, F <= 128, plain rows: every layer-sized call).
// The kernel above moves each 16-row slab through shared memory five times (raw, two column reads, four operand
// slabs, MMA reads of both operands: 110 KB per 16 rows = 0.47 ms of shared-memory time alone), keeps 32 KB in flight
// per CTA and empties its whole pipeline every 256 rows to drain the accumulator: 0.83 ms for 2.45 M rows = 0.41 of
// the copy peak.  Here, one CTA per SM:
//   warp 0       producer     X and dH slabs of 32 rows x 128 columns by TMA (columns past K / F and rows past n
//                             arrive as zeros), four 32 KB slots that stay full across segment boundaries
//   warps 4-7    A side       thread m reads COLUMN m of the X slab (conflict-free), splits hi / lo and stores them
//                             into tensor memory (lane = m, column = row of the slab): the transposed A operand
//                             never goes back to shared memory
//   warps 8-11   B side       thread c reads column c of the dH slab and writes the K-major hi / lo operand slabs
//   warp 1       issuer       tcgen05.mma (A from TMEM), 256-row segments alternating between two accumulators
//   warps 12-15  drain        accumulator -> registers -> 4 KB staging -> TMA store (first segment) or TMA
//                             reduce-add (later ones) into the CTA's partial tile, which stays in L2; segment s is
//                             drained while segment s + 1 is multiplied.  The adds of one tile are issued in segment
//                             order and each is complete before the next is issued: deterministic.
// The per-CTA partials are reduced in fp64 by tn_reduce_kernel as before.
// History: cvt.rna hi / lo splits 0.56 ms (ncu: the issuer waits for the B side, which is busy 80 % of the time;
// cvt.rna runs at 16 per clock and SM) -> truncation splits 0.52 ms; eight warps per converter side instead of four:
// slower (0.55 ms), dropped.  Tensor pipe 44 % active, shared-memory pipe 35 %: the remaining gap to the 0.34 ms DRAM
// time is in the hand-offs between the roles (two B stages), not in any one unit.
// ---------------------------------------------------------------------------------------------
constexpr int TP_ROWS = 32;                              // reduction rows per slab
constexpr int TP_HALF_BYTES = TP_ROWS * TC_BM * 4;       // 16384: one operand's raw slab, and one hi or lo B slab
constexpr int TP_RAW_BYTES = 2 * TP_HALF_BYTES;
constexpr int TP_RAW_SLOTS = 3;
constexpr int TP_B_STAGES = 3;
constexpr int TP_A_STAGES = 4;
constexpr int TP_SEG_SLABS = kTnRowsPerSplit / TP_ROWS;  // 8 slabs per accumulator segment
constexpr int TP_THREADS = 16 * 32;
constexpr int TP_SMEM_BYTES = 1024 + TP_RAW_SLOTS * TP_RAW_BYTES + TP_B_STAGES * 2 * TP_HALF_BYTES + PS_STAGING_BYTES;
static_assert(TP_SMEM_BYTES <= PS_SMEM_MAX, "persistent TN kernel: shared memory plan too large");

struct TpArgs {
    int64_t n;
    int k, f;
    int64_t rows_per_cta;   // multiple of kTnRowsPerSplit
};

__device__ __forceinline__ void tp_tma_reduce_add_2d(const void* tmap, int c0, int c1, const void* src) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap),
                 "r"(c0), "r"(c1), "r"(tc_smem_u32(src))
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(TP_THREADS, 1) tc_gemm_tn_persist_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                           const __grid_constant__ CUtensorMap tmap_g,
                                                                           const __grid_constant__ CUtensorMap tmap_part,
                                                                           const TpArgs g) {
    extern __shared__ uint8_t tp_smem_raw[];
    __shared__ __align__(8) uint64_t raw_full[TP_RAW_SLOTS], raw_free[TP_RAW_SLOTS];
    __shared__ __align__(8) uint64_t a_full[TP_A_STAGES], a_free[TP_A_STAGES];
    __shared__ __align__(8) uint64_t b_full[TP_B_STAGES], b_free[TP_B_STAGES];
    __shared__ __align__(8) uint64_t acc_full[2], acc_free[2];
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* smem = tp_smem_raw + ((1024u - (tc_smem_u32(tp_smem_raw) & 1023u)) & 1023u);
    uint8_t* s_raw = smem;
    uint8_t* s_b = s_raw + TP_RAW_SLOTS * TP_RAW_BYTES;
    uint8_t* s_stage = s_b + TP_B_STAGES * 2 * TP_HALF_BYTES;

    if (tid == 0) {
        for (int i = 0; i < TP_RAW_SLOTS; ++i) {
            tc_mbar_init(&raw_full[i], 1);
            tc_mbar_init(&raw_free[i], 256);
        }
        for (int i = 0; i < TP_A_STAGES; ++i) {
            tc_mbar_init(&a_full[i], 128);
            tc_mbar_init(&a_free[i], 1);
        }
        for (int i = 0; i < TP_B_STAGES; ++i) {
            tc_mbar_init(&b_full[i], 128);
            tc_mbar_init(&b_free[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc_mbar_init(&acc_full[i], 1);
            tc_mbar_init(&acc_free[i], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         tc_smem_u32(&tmem_base_smem)),
                     "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    const int64_t r_beg = (int64_t)blockIdx.x * g.rows_per_cta;
    const int64_t r_end = r_beg + g.rows_per_cta < g.n ? r_beg + g.rows_per_cta : g.n;
    const int total = r_end > r_beg ? (int)((r_end - r_beg + TP_ROWS - 1) / TP_ROWS) : 0;
    const int nseg = (total + TP_SEG_SLABS - 1) / TP_SEG_SLABS;

    if (warp == 0) {
        if (lane == 0) {
            // ===== producer =====
            for (int i = 0; i < total; ++i) {
                const int slot = i % TP_RAW_SLOTS, use = i / TP_RAW_SLOTS;
                if (use > 0) tc_mbar_wait(&raw_free[slot], (uint32_t)(use - 1) & 1u);
                const int r0 = (int)(r_beg + (int64_t)i * TP_ROWS);
                tc_expect_tx(&raw_full[slot], TP_RAW_BYTES);
                ps_tma_load_2d(s_raw + slot * TP_RAW_BYTES, &tmap_a, 0, r0, &raw_full[slot]);
                ps_tma_load_2d(s_raw + slot * TP_RAW_BYTES + TP_HALF_BYTES, &tmap_g, 0, r0, &raw_full[slot]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== issuer =====
            for (int i = 0; i < total; ++i) {
                const int seg = i / TP_SEG_SLABS, in_seg = i % TP_SEG_SLABS, buf = seg & 1;
                if (in_seg == 0 && seg >= 2) tc_mbar_wait(&acc_free[buf], (uint32_t)((seg >> 1) - 1) & 1u);
                const int as = i % TP_A_STAGES, bs = i % TP_B_STAGES;
                tc_mbar_wait(&a_full[as], (uint32_t)(i / TP_A_STAGES) & 1u);
                tc_mbar_wait(&b_full[bs], (uint32_t)(i / TP_B_STAGES) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)buf * TC_TMEM_COLS;
                const uint32_t a_hi = tmem_base + 2 * TC_TMEM_COLS + (uint32_t)as * 64u, a_lo = a_hi + 32u;
                const uint32_t b_hi = tc_smem_u32(s_b + bs * 2 * TP_HALF_BYTES), b_lo = b_hi + TP_HALF_BYTES;
#pragma unroll
                for (int j = 0; j < TP_ROWS / 8; ++j) {
                    const uint32_t ko = j * 2 * TC_LBO;
                    tc_mma_ts(d_tmem, a_hi + j * 8, tc_smem_desc(b_hi + ko), (in_seg > 0 || j > 0) ? 1u : 0u);
                    tc_mma_ts(d_tmem, a_hi + j * 8, tc_smem_desc(b_lo + ko), 1u);
                    tc_mma_ts(d_tmem, a_lo + j * 8, tc_smem_desc(b_hi + ko), 1u);
                }
                tc_commit(&a_free[as]);
                tc_commit(&b_free[bs]);
                if (in_seg == TP_SEG_SLABS - 1 || i == total - 1) tc_commit(&acc_full[buf]);
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===== A side: column m of the X slab -> TMEM lane m (hi | lo) =====
        const int q = warp & 3, m = q * 32 + lane;
        const uint32_t t_lane = tmem_base + 2 * TC_TMEM_COLS + ((uint32_t)(q * 32) << 16);
        for (int i = 0; i < total; ++i) {
            const int slot = i % TP_RAW_SLOTS, stage = i % TP_A_STAGES;
            tc_mbar_wait(&raw_full[slot], (uint32_t)(i / TP_RAW_SLOTS) & 1u);
            const float* raw = reinterpret_cast<const float*>(s_raw + slot * TP_RAW_BYTES);
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int r = 0; r < TP_ROWS; ++r) {   // hi = the raw bits (the tensor core reads the TF32 part), lo = x - trunc(x):
                const float x = raw[r * TC_BM + m];   // two full-rate instructions per element (cvt.rna runs at 16 / clk / SM)
                hi[r] = __float_as_uint(x);
                lo[r] = __float_as_uint(x - trunc_tf32(x));
            }
            if (i >= TP_A_STAGES) {
                tc_mbar_wait(&a_free[stage], (uint32_t)(i / TP_A_STAGES - 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            tc_st32(t_lane + (uint32_t)stage * 64u, hi);
            tc_st32(t_lane + (uint32_t)stage * 64u + 32u, lo);
            tc_arrive(&raw_free[slot]);          // after the stores that consume the loaded registers (RULE above)
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            tc_arrive(&a_full[stage]);
        }
    } else if (warp >= 8 && warp < 12) {
        // ===== B side: column c of the dH slab -> K-major hi / lo operand slabs (chunk-major, no swizzle) =====
        const int c = tid - 256;
        for (int i = 0; i < total; ++i) {
            const int slot = i % TP_RAW_SLOTS, stage = i % TP_B_STAGES;
            tc_mbar_wait(&raw_full[slot], (uint32_t)(i / TP_RAW_SLOTS) & 1u);
            const float* raw = reinterpret_cast<const float*>(s_raw + slot * TP_RAW_BYTES + TP_HALF_BYTES);
            float v[TP_ROWS];
#pragma unroll
            for (int r = 0; r < TP_ROWS; ++r) v[r] = raw[r * TC_BN + c];
            if (i >= TP_B_STAGES) tc_mbar_wait(&b_free[stage], (uint32_t)(i / TP_B_STAGES - 1) & 1u);
            uint8_t* st = s_b + stage * 2 * TP_HALF_BYTES + c * 16;
#pragma unroll
            for (int ch = 0; ch < TP_ROWS / 4; ++ch) {
                const float4 h = make_float4(v[ch * 4], v[ch * 4 + 1], v[ch * 4 + 2], v[ch * 4 + 3]);
                const float4 l = make_float4(h.x - trunc_tf32(h.x), h.y - trunc_tf32(h.y), h.z - trunc_tf32(h.z),
                                             h.w - trunc_tf32(h.w));
                *reinterpret_cast<float4*>(st + ch * TC_LBO) = h;
                *reinterpret_cast<float4*>(st + TP_HALF_BYTES + ch * TC_LBO) = l;
            }
            tc_arrive(&raw_free[slot]);          // after the stores that consume the loaded registers (RULE above)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_arrive(&b_full[stage]);
        }
    } else if (warp >= 12) {
        // ===== drain: warp q owns output rows [32 q, 32 q + 32) of the CTA's partial tile =====
        const int q = warp & 3;
        uint8_t* stg = s_stage + q * (32 * 32 * 4);
        const int prow = (int)blockIdx.x * TC_BM + q * 32;      // row of the [ctas * 128, f] partial matrix
        for (int seg = 0; seg < nseg; ++seg) {
            const int buf = seg & 1;
            tc_mbar_wait(&acc_full[buf], (uint32_t)(seg >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + (uint32_t)buf * TC_TMEM_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int part = 0; part < 4; ++part) {
                const bool live = part * 32 < g.f && q * 32 < g.k;      // warp-uniform
                uint32_t acc[32];
                if (live) tc_ld32(taddr + (uint32_t)(part * 32), acc);
                if (part == 3) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    tc_arrive(&acc_free[buf]);
                }
                if (!live) continue;
                // first part of a segment: every earlier store / add of this warp is COMPLETE, so the adds to one tile stay
                // in segment order; later parts (other columns) only need the staging buffer back
                if (lane == 0) {
                    if (part == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                    else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
                __syncwarp();
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    *reinterpret_cast<float4*>(stg + lane * 128 + ((e ^ (lane & 7)) * 16)) =
                        make_float4(__uint_as_float(acc[4 * e]), __uint_as_float(acc[4 * e + 1]),
                                    __uint_as_float(acc[4 * e + 2]), __uint_as_float(acc[4 * e + 3]));
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    if (seg == 0) ps_tma_store_2d(&tmap_part, part * 32, prow, stg);
                    else tp_tma_reduce_add_2d(&tmap_part, part * 32, prow, stg);
                }
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// out[e] = sum_s part[s][e] in fp64; split s starts at part + s * split_stride.  One warp per output element: lane l
// adds the splits l, l + 32, ... in ascending order, then a fixed shuffle tree (deterministic).  (One thread per element
// walked up to 148 dependent L2 loads: 80 us for a 100 x 128 gradient.)
__global__ void __launch_bounds__(256) tn_reduce_kernel(const float* __restrict__ part, int64_t splits, int64_t split_stride,
                                                        int64_t rows, int64_t cols, float* __restrict__ out, int64_t ldo) {
    const int64_t total = rows * cols;
    const int lane = threadIdx.x & 31;
    for (int64_t e = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); e < total; e += (int64_t)gridDim.x * 8) {
        double s = 0.0;
        for (int64_t p = lane; p < splits; p += 32) s += (double)part[p * split_stride + e];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) out[(e / cols) * ldo + (e % cols)] = (float)s;
    }
}

static inline int tc_slabs(int64_t k) { return (int)ceil_div(k, TC_BK); }
static inline int tc_ntiles(int64_t f) { return (int)ceil_div(f, TC_BN); }
static inline size_t tc_image_floats(int64_t k, int64_t f) {
    return (size_t)tc_ntiles(f) * (size_t)tc_slabs(k) * (2 * TC_TILE_BYTES / 4);
}

}  // namespace gg

using namespace gg;

// The tensor core accumulates with truncation: the result is biased towards zero by ~6e-9 * K relative
// (measured, profiles/probes/tc_err_probe.py).  To stay inside 1e-5 for any K the reduction is cut into launches
// of at most kTcMaxKPerLaunch = 256 k (bias <= 1.6e-6); each launch's partial tile is added to `out` in
// fp32 (round to nearest) by the next launch's epilogue.  GNN hidden sizes (K <= 256) take one launch.
constexpr int kTcMaxKPerLaunch = 256;

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

// ---- persistent path: eligibility, tensor map, launch ----
typedef CUresult (*PsEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PsEncodeTiled ps_encoder() {
    static PsEncodeTiled fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        (void)cudaGetLastError();
        return reinterpret_cast<PsEncodeTiled>(p);
    }();
    return fn;
}
static int ps_raw_slots(int k_slabs16) {
    const int64_t left = (int64_t)PS_SMEM_MAX - 1024 - (int64_t)k_slabs16 * 2 * TC_TILE_BYTES - PS_STAGING_BYTES;
    int64_t r = left / PS_SLAB_BYTES;
    return (int)(r > PS_MAX_RAW ? PS_MAX_RAW : r);
}
static bool ps_eligible(const gg_gemm_segment& s, const TcPiece& pc, int64_t n, int64_t f, const float* bias,
                        const float* relu_mask, int64_t ld_mask, const float* out, int64_t ldo) {
    static const bool on = [] { const char* e = getenv("GG_GEMM_PERSIST"); return !e || e[0] != '0'; }();
    if (!on || s.scale || pc.k_off != 0 || pc.k_len != s.k) return false;
    if (s.k < 4 || s.k > 128 || s.k % 4 || s.lda % 4 || (reinterpret_cast<uintptr_t>(s.a) & 15)) return false;
    if (f > TC_BN || f % 4 || ldo % 4 || (reinterpret_cast<uintptr_t>(out) & 15)) return false;
    if (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) return false;
    if (relu_mask && (ld_mask % 4 || (reinterpret_cast<uintptr_t>(relu_mask) & 15))) return false;
    if (n < (int64_t)TC_BM * kNumSMs || n >= ((int64_t)1 << 31)) return false;   // small problems: the per-tile kernel
    return ps_raw_slots(tc_slabs(s.k)) >= 2 && ps_encoder() != nullptr;
}
static int ps_launch(const gg_gemm_segment& s, int b_trans, int64_t n, int64_t f, const float* bias, int act,
                     const float* relu_mask, int64_t ld_mask, float* out, int64_t ldo, void* workspace, cudaStream_t st) {
    PsEncodeTiled enc = ps_encoder();
    if (!enc) return GG_ERR_UNSUPPORTED;
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)s.k, (cuuint64_t)n};
    const cuuint64_t gstride[1] = {(cuuint64_t)s.lda * 4};
    const cuuint32_t box[2] = {(cuuint32_t)PS_BK, (cuuint32_t)TC_BM};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(s.a), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return GG_ERR_UNSUPPORTED;
    CUtensorMap tmap_out;
    const cuuint64_t odim[2] = {(cuuint64_t)f, (cuuint64_t)n};
    const cuuint64_t ostride[1] = {(cuuint64_t)ldo * 4};
    const cuuint32_t obox[2] = {32, 32};
    const CUresult ro = enc(&tmap_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, odim, ostride, obox, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (ro != CUDA_SUCCESS) return GG_ERR_UNSUPPORTED;
    Carver c(workspace);
    const int k16 = tc_slabs(s.k);
    float* image = c.take<float>(tc_image_floats(s.k, f));
    const int64_t chunks = (int64_t)k16 * (TC_BK / 4) * TC_BN;
    b_image_kernel<<<(int)ceil_div(chunks, 256), 256, 0, st>>>(s.b, s.ldb, b_trans, (int)s.k, (int)f, k16, 1, image);
    GG_LAUNCHED();
    PsArgs g{};
    g.b_image = image; g.k = (int)s.k; g.k_slabs16 = k16; g.raw_slots = ps_raw_slots(k16);
    g.n = n; g.f = (int)f; g.bias = bias; g.act = act; g.relu_mask = relu_mask; g.ld_mask = ld_mask; g.out = out; g.ldo = ldo;
    const int smem = 1024 + g.raw_slots * PS_SLAB_BYTES + k16 * 2 * TC_TILE_BYTES + PS_STAGING_BYTES;
    static bool attr_done = false;
    if (!attr_done) {
        GG_CUDA(cudaFuncSetAttribute(tc_gemm_persist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PS_SMEM_MAX));
        GG_CUDA(cudaFuncSetAttribute(tc_gemm_persist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PS_SMEM_MAX));
        attr_done = true;
    }
    const int64_t tiles = ceil_div(n, TC_BM);
    const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
    if (relu_mask) tc_gemm_persist_kernel<true><<<grid, PS_THREADS, smem, st>>>(tmap, tmap_out, g);
    else tc_gemm_persist_kernel<false><<<grid, PS_THREADS, smem, st>>>(tmap, tmap_out, g);
    GG_LAUNCHED();
    return GG_OK;
}

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
    if (np == 1 && ps_eligible(segs[pieces[0].seg], pieces[0], n, f, bias, relu_mask, ld_mask, out, ldo)) {
        const int rc = ps_launch(segs[pieces[0].seg], b_trans, n, f, bias, act, relu_mask, ld_mask, out, ldo, workspace, st);
        if (rc != GG_ERR_UNSUPPORTED) return rc;   // no tensor-map entry point in this driver: the per-tile kernel serves
    }
    static bool attr_done = false;
    if (!attr_done) {
        GG_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NN_SMEM_BYTES));
        GG_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NN_SMEM_BYTES));
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
        if (vec_a) tc_gemm_kernel<true><<<grid, TC_BLOCK, NN_SMEM_BYTES, st>>>(g);
        else tc_gemm_kernel<false><<<grid, TC_BLOCK, NN_SMEM_BYTES, st>>>(g);
        GG_LAUNCHED();
    }
    return GG_OK;
}

static void tn_tc_plan(int64_t n, int64_t k, int64_t f, int64_t* ctas, int64_t* rows_per_cta) {
    // about two waves of CTAs over the 148 SMs x 2; every CTA runs whole 512-row segments back to back
    int64_t tiles = ceil_div(k, TC_BM) * ceil_div(f, TC_BN);
    int64_t segs = ceil_div(n > 0 ? n : 1, kTnRowsPerSplit);
    int64_t want = ceil_div((int64_t)kNumSMs * 4, tiles);
    int64_t c = segs < want ? segs : want;
    if (c < 1) c = 1;
    int64_t rpc = ceil_div(segs, c) * kTnRowsPerSplit;
    *rows_per_cta = rpc;
    *ctas = ceil_div(n > 0 ? n : 1, rpc);
}

static void tp_plan(int64_t n, int64_t* ctas, int64_t* rows_per_cta) {
    const int64_t segs = ceil_div(n > 0 ? n : 1, kTnRowsPerSplit);
    const int64_t rpc = ceil_div(segs, (int64_t)kNumSMs) * kTnRowsPerSplit;
    *rows_per_cta = rpc;
    *ctas = ceil_div(n > 0 ? n : 1, rpc);
}
static bool tp_eligible(const float* a, int64_t lda, const int64_t* row_index, const float* g, int64_t ldg, int64_t n, int64_t k,
                        int64_t f) {
    static const bool on = [] { const char* e = getenv("GG_GEMM_PERSIST"); return !e || e[0] != '0'; }();
    if (!on || row_index || k < 4 || k > TC_BM || f < 4 || f > TC_BN || k % 4 || f % 4 || lda % 4 || ldg % 4) return false;
    if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(g) & 15)) return false;
    if (n < (int64_t)kNumSMs * kTnRowsPerSplit || n >= ((int64_t)1 << 31)) return false;
    return ps_encoder() != nullptr;
}

size_t gg_gemm_tn_tc_workspace_bytes(int64_t n, int64_t k, int64_t f) {
    int64_t ctas, rpc;
    tn_tc_plan(n, k, f, &ctas, &rpc);
    size_t b = (size_t)ctas * (size_t)k * (size_t)f * sizeof(float) + 256;
    if (k <= TC_BM && f <= TC_BN) {   // the persistent kernel's padded partial tiles: [ctas][128][f]
        tp_plan(n, &ctas, &rpc);
        const size_t pb = (size_t)ctas * TC_BM * (size_t)f * sizeof(float) + 256;
        if (pb > b) b = pb;
    }
    return b;
}

static int tp_launch(const float* a, int64_t lda, const float* g, int64_t ldg, int64_t n, int64_t k, int64_t f, float* out,
                     int64_t ldo, void* workspace, cudaStream_t st) {
    PsEncodeTiled enc = ps_encoder();
    int64_t ctas, rpc;
    tp_plan(n, &ctas, &rpc);
    float* partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    CUtensorMap ta, tg, tp;
    const cuuint32_t estr[2] = {1, 1};
    const cuuint32_t box[2] = {(cuuint32_t)TC_BM, (cuuint32_t)TP_ROWS};
    const cuuint64_t adim[2] = {(cuuint64_t)k, (cuuint64_t)n}, astr[1] = {(cuuint64_t)lda * 4};
    const cuuint64_t gdim[2] = {(cuuint64_t)f, (cuuint64_t)n}, gstr[1] = {(cuuint64_t)ldg * 4};
    const cuuint64_t pdim[2] = {(cuuint64_t)f, (cuuint64_t)ctas * TC_BM}, pstr[1] = {(cuuint64_t)f * 4};
    const cuuint32_t pbox[2] = {32, 32};
    if (enc(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(a), adim, astr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        enc(&tg, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(g), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        enc(&tp, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, partial, pdim, pstr, pbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return GG_ERR_UNSUPPORTED;
    static bool attr_done = false;
    if (!attr_done) {
        GG_CUDA(cudaFuncSetAttribute(tc_gemm_tn_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM_BYTES));
        attr_done = true;
    }
    TpArgs t{n, (int)k, (int)f, rpc};
    tc_gemm_tn_persist_kernel<<<(int)ctas, TP_THREADS, TP_SMEM_BYTES, st>>>(ta, tg, tp, t);
    GG_LAUNCHED();
    const int64_t total = k * f;
    const int blocks = (int)(ceil_div(total, 8) < kNumSMs * 16 ? ceil_div(total, 8) : kNumSMs * 16);
    tn_reduce_kernel<<<blocks, 256, 0, st>>>(partial, ctas, (int64_t)TC_BM * f, k, f, out, ldo);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_gemm_tn_tc_f32(const float* a, int64_t lda, const int64_t* row_index, const float* g, int64_t ldg,
                      int64_t n, int64_t k, int64_t f, float* out, int64_t ldo, void* workspace,
                      size_t workspace_bytes, gg_stream_t stream) {
    GG_REQUIRE(n >= 0 && k >= 0 && f >= 0, "gg_gemm_tn_tc_f32: negative size");
    if (k == 0 || f == 0) return GG_OK;
    GG_REQUIRE(out && ldo >= f, "gg_gemm_tn_tc_f32: bad output");
    cudaStream_t st = as_stream(stream);
    if (n == 0) {
        GG_CUDA(cudaMemset2DAsync(out, (size_t)ldo * 4, 0, (size_t)f * 4, (size_t)k, st));
        return GG_OK;
    }
    GG_REQUIRE(a && g && lda >= k && ldg >= f && workspace, "gg_gemm_tn_tc_f32: bad operands");
    GG_REQUIRE(k < (1 << 20) && f < (1 << 20), "gg_gemm_tn_tc_f32: k / f out of range");
    if (workspace_bytes < gg_gemm_tn_tc_workspace_bytes(n, k, f)) {
        set_error("gg_gemm_tn_tc_f32: workspace %zu < %zu", workspace_bytes, gg_gemm_tn_tc_workspace_bytes(n, k, f));
        return GG_ERR_WORKSPACE;
    }
    if (tp_eligible(a, lda, row_index, g, ldg, n, k, f)) {
        const int rc = tp_launch(a, lda, g, ldg, n, k, f, out, ldo, workspace, st);
        if (rc != GG_ERR_UNSUPPORTED) return rc;
    }
    int64_t ctas, rpc;
    tn_tc_plan(n, k, f, &ctas, &rpc);
    static bool attr_done = false;
    if (!attr_done) {
        GG_CUDA(cudaFuncSetAttribute(tc_gemm_tn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TN_SMEM_BYTES));
        GG_CUDA(cudaFuncSetAttribute(tc_gemm_tn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TN_SMEM_BYTES));
        attr_done = true;
    }
    TnTcArgs t{a, lda, g, ldg, row_index, n, (int)k, (int)f, rpc, static_cast<float*>(workspace)};
    const int64_t tiles = ceil_div(k, TC_BM) * ceil_div(f, TC_BN);
    const bool vec = (k % 4 == 0) && (f % 4 == 0) && (lda % 4 == 0) && (ldg % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(a) & 15) == 0) && ((reinterpret_cast<uintptr_t>(g) & 15) == 0);
    dim3 grid((unsigned)tiles, (unsigned)ctas);
    if (vec) tc_gemm_tn_kernel<true><<<grid, TC_BLOCK, TN_SMEM_BYTES, st>>>(t);
    else tc_gemm_tn_kernel<false><<<grid, TC_BLOCK, TN_SMEM_BYTES, st>>>(t);
    GG_LAUNCHED();
    const int64_t total = k * f;
    int blocks = (int)(ceil_div(total, 8) < kNumSMs * 16 ? ceil_div(total, 8) : kNumSMs * 16);
    tn_reduce_kernel<<<blocks, 256, 0, st>>>(t.partial, ctas, k * f, k, f, out, ldo);
    GG_LAUNCHED();
    return GG_OK;
}

}  // extern "C"
