// Degree-sorted sliced-ELL aggregation for narrow rows (f <= 128) — the kernel behind the feature-sliced exchange of the
// row-partitioned path, where every rank walks ALL slots of the graph at F/P columns.
//
// Why: the merge-path kernels (spmm_mp.cu) pay for row ends.  At F/P = 16 the grouped-slot kernel executes 628 M warp
// instructions for 64.3 M slots (profiles/r01_narrow_ncu_raw.csv) — segment sweeps, cross-group butterflies, split-row
// partials — and streams 46 G slots/s whether a slot is 64 or 128 bytes wide: instruction/latency bound, DRAM at 23 %.
// Round 1's degree-binned experiment (one group of lanes per row, rows sorted by degree) removed the row ends but read its
// indices with dependent 4-byte loads per lane: 58 G slots/s at F = 16 and slower than merge-path from F = 32 up.
//
// Here the layout itself is rebuilt once per graph (gg_sell_build) so that the kernel has nothing left to decide:
//   * rows longer than `seg` slots are cut into virtual rows of <= seg slots (their partial sums are combined by a tiny
//     fix-up kernel in a fixed order); every other row is one virtual row;
//   * virtual rows are sorted by DESCENDING length (stable radix sort) and taken 8 at a time: a chunk.  All rows of a
//     chunk are padded to the chunk's longest row rounded up to 4 slots (sorted order => < 7 % padding on a power-law
//     graph, bounded by slots + 3 V + 7 seg);
//   * inside a chunk the neighbour ids are stored as int4 "units": unit (k4, row) = slots 4 k4 .. 4 k4 + 3 of that row,
//     laid out k4-major, so that one warp-wide 128-bit load fetches 4 slots for each of the 32/G rows a warp works on and
//     consecutive loads walk consecutive memory (padding = -1, weight 0).
// A warp owns a chunk: G lanes per row (G = 4, 8, 16, 32 >= f/4), 32/G rows per pass, 8 / (32/G) passes interleaved so
// that 8 independent 128-bit feature gathers are in flight per lane while the next pair of index units is already being
// fetched.  No row-end test, no shuffles, no shared memory, no atomics on data; every row is summed in slot order by one
// group => bitwise run-to-run deterministic.  Chunks are handed out longest-first by an atomic counter.
// The epilogue (mean, self term, bias, rank-1 terms, peer-memory output) is the one of the merge-path kernels.
#include "common.cuh"

namespace gg {

int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, void* ws, cudaStream_t st);
size_t scan_workspace_bytes();

constexpr int kSellRows = 8;     // virtual rows per chunk
constexpr int kSellThreads = 256;
constexpr int32_t kSellNoRow = INT32_MIN;

static inline int sell_bits_for(int64_t v) {
    int b = 1;
    while (((int64_t)1 << b) <= v) ++b;
    return b;
}
static inline int sell_grid(int64_t total, int per_block) {
    int64_t b = ceil_div(total, per_block);
    if (b > (int64_t)kNumSMs * 8) b = (int64_t)kNumSMs * 8;
    return (int)(b < 1 ? 1 : b);
}

// ---- build -------------------------------------------------------------------------------------------------
// cnt[r] = virtual rows of row r (>= 1), flag[r] = row is split; entry n of both = 0 (so the scans' entry n = totals)
__global__ void __launch_bounds__(256) sell_count_kernel(const int32_t* __restrict__ rowptr, int64_t n, int seg,
                                                         uint32_t* __restrict__ cnt, uint32_t* __restrict__ flag) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n; r += (int64_t)gridDim.x * blockDim.x) {
        uint32_t c = 0;
        if (r < n) {
            const int deg = rowptr[r + 1] - rowptr[r];
            c = deg <= seg ? 1u : (uint32_t)((deg + seg - 1) / seg);
        }
        cnt[r] = c;
        flag[r] = c > 1 ? 1u : 0u;
    }
}

// keys / vals of the sort, per-virtual-row slot offset and destination; the split rows' list
__global__ void __launch_bounds__(256) sell_emit_kernel(const int32_t* __restrict__ rowptr, int64_t n, int seg,
                                                        int64_t vcap, const uint32_t* __restrict__ vstart,
                                                        const uint32_t* __restrict__ hstart, uint32_t* __restrict__ keys,
                                                        uint32_t* __restrict__ vals, int32_t* __restrict__ voff,
                                                        int32_t* __restrict__ vdst_tmp, int32_t* __restrict__ hub_rows,
                                                        int32_t* __restrict__ hub_pptr, int32_t* __restrict__ info) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n; r += stride) {
        const uint32_t v0 = vstart[r], h0 = hstart[r];
        const int32_t p0 = (int32_t)(v0 - ((uint32_t)r - h0));   // partial rows before row r
        if (r == n) {
            hub_pptr[h0] = p0;
            info[0] = (int32_t)v0;   // V
            info[3] = (int32_t)h0;   // split rows
            info[4] = p0;            // partial rows
            continue;
        }
        const int b = rowptr[r], deg = rowptr[r + 1] - b;
        const uint32_t c = vstart[r + 1] - v0;
        if (c > 1) {
            hub_rows[h0] = (int32_t)r;
            hub_pptr[h0] = p0;
        }
        for (uint32_t s = 0; s < c; ++s) {
            const int len = min(seg, deg - (int)s * seg);
            const uint32_t v = v0 + s;
            keys[v] = (uint32_t)(seg - len);
            vals[v] = v;
            voff[v] = b + (int)s * seg;
            vdst_tmp[v] = c == 1 ? (int32_t)r : -(p0 + (int32_t)s) - 1;
        }
    }
    // entries past V (the capacity is a host-side bound): sort to the end, never read
    for (int64_t v = (int64_t)vstart[n] + blockIdx.x * blockDim.x + threadIdx.x; v < vcap; v += stride) {
        keys[v] = (uint32_t)seg + 1u;
        vals[v] = (uint32_t)v;
    }
}

// int4 units of chunk c = 8 * ceil(len(first row) / 4); entry `chunks` = 0
__global__ void __launch_bounds__(256) sell_units_kernel(const uint32_t* __restrict__ keys_sorted, int64_t chunks,
                                                         int seg, uint32_t* __restrict__ units) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c <= chunks; c += (int64_t)gridDim.x * blockDim.x) {
        uint32_t u = 0;
        if (c < chunks) {
            const int len = seg - (int)keys_sorted[c * kSellRows];   // -1 for capacity padding
            u = len > 0 ? (uint32_t)((len + 3) / 4) * kSellRows : 0u;
        }
        units[c] = u;
    }
}

// one warp per chunk: lane = (row q = lane / 4, slot kk = lane % 4 of the unit)
__global__ void __launch_bounds__(256) sell_fill_kernel(const int32_t* __restrict__ nbr,
                                                        const uint32_t* __restrict__ keys_sorted,
                                                        const uint32_t* __restrict__ order, const int32_t* __restrict__ voff,
                                                        const int32_t* __restrict__ vdst_tmp,
                                                        const uint32_t* __restrict__ chunk_ptr, int64_t chunks, int seg,
                                                        int32_t* __restrict__ idx, int32_t* __restrict__ slot_of,
                                                        int32_t* __restrict__ vdst, int32_t* __restrict__ info) {
    const int lane = threadIdx.x & 31, q = lane >> 2, kk = lane & 3;
    for (int64_t c = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); c < chunks; c += (int64_t)gridDim.x * 8) {
        const int64_t i = c * kSellRows + q;
        const int len = seg - (int)keys_sorted[i];
        const uint32_t v = order[i];
        const int off = len > 0 ? voff[v] : 0;
        if (kk == 0) vdst[i] = len >= 0 ? vdst_tmp[v] : kSellNoRow;
        const uint32_t base = chunk_ptr[c];
        const int nk = (int)((chunk_ptr[c + 1] - base) / kSellRows);
        int32_t* di = idx + (int64_t)base * 4 + lane;
        int32_t* ds = slot_of + (int64_t)base * 4 + lane;
        for (int k4 = 0; k4 < nk; ++k4) {
            const int k = k4 * 4 + kk;
            const bool ok = k < len;
            di[(int64_t)k4 * 32] = ok ? nbr[off + k] : -1;
            ds[(int64_t)k4 * 32] = ok ? off + k : -1;
        }
        if (c == chunks - 1 && lane == 0) {
            info[1] = (int32_t)chunks;
            info[2] = (int32_t)chunk_ptr[chunks];   // int4 units in use
        }
    }
}

// ---- hub hint ----------------------------------------------------------------------------------------------------
// The feature matrix is ~10x the L2 on the products graph and every gather used to be tagged evict_last: rows referenced
// a handful of times (most of a power-law graph) evicted the rows referenced thousands of times just as often as the
// other way round — L2 hit rate 35 %, 18.6 GB of DRAM reads for 33 GB of gathers at the DRAM roofline (0.93 of the copy
// peak, profiles/r02_spmm_sell128_bench_ncu_raw.csv).  With the hint, bit 30 of an index marks a HUB source — one of the
// ~`hubs` most referenced nodes, whose rows fit the L2 together — and the kernel keeps hub rows (evict_last) while it
// streams the others through (evict_first).  Caching only: the results are bitwise unchanged.
constexpr int32_t kSellHubBit = 1 << 30;
constexpr int kSellFreqBins = 4096;

__global__ void __launch_bounds__(256) sell_freq_kernel(const int32_t* __restrict__ nbr, int64_t slots, int32_t* __restrict__ freq) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(freq + nbr[s], 1);      // integer counts: the order of the additions does not matter
}
// CTA-local histogram in shared memory (most nodes of a power-law graph share a handful of low counts: global atomics on
// those bins serialised, 0.58 ms at 2.45 M nodes), merged into the global bins once per CTA
__global__ void __launch_bounds__(256) sell_freq_hist_kernel(const int32_t* __restrict__ freq, int64_t n, int32_t* __restrict__ bins) {
    __shared__ int32_t sb[kSellFreqBins];
    for (int b = threadIdx.x; b < kSellFreqBins; b += blockDim.x) sb[b] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int f = freq[i];
        atomicAdd(sb + (f < kSellFreqBins ? f : kSellFreqBins - 1), 1);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < kSellFreqBins; b += blockDim.x)
        if (sb[b]) atomicAdd(bins + b, sb[b]);
}
// smallest threshold T >= 1 with #{freq >= T} <= hubs: 256 threads sum 16 bins each, thread 0 walks the partial sums
// from the top and then the 16 bins of the run that crosses the budget
__global__ void __launch_bounds__(256) sell_freq_threshold_kernel(const int32_t* __restrict__ bins, int64_t hubs,
                                                                  int32_t* __restrict__ threshold) {
    __shared__ int32_t sb[kSellFreqBins];
    __shared__ long long part[256];
    constexpr int kPer = kSellFreqBins / 256;
    long long acc = 0;
    for (int q = 0; q < kPer; ++q) {
        const int b = threadIdx.x * kPer + q;
        sb[b] = bins[b];
        acc += sb[b];
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x != 0) return;
    long long above = 0;
    int t = kSellFreqBins;          // nothing qualifies
    for (int p = 255; p >= 0; --p) {
        if (above + part[p] <= hubs && p > 0) {   // the whole run fits (bin 0 never counts: p = 0 is walked bin by bin)
            above += part[p];
            t = p * kPer;
            continue;
        }
        for (int b = p * kPer + kPer - 1; b >= (p == 0 ? 1 : p * kPer); --b) {
            if (above + sb[b] > hubs) { *threshold = t; return; }
            above += sb[b];
            t = b;
        }
        if (p == 0) break;
    }
    *threshold = t;
}
__global__ void __launch_bounds__(256) sell_hub_flag_kernel(const int32_t* __restrict__ idx, int64_t total,
                                                            const int32_t* __restrict__ freq,
                                                            const int32_t* __restrict__ threshold, int32_t* __restrict__ out) {
    const int t = *threshold;
    for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < total; d += (int64_t)gridDim.x * blockDim.x) {
        const int j = idx[d];
        out[d] = (j >= 0 && freq[j] >= t) ? (j | kSellHubBit) : j;
    }
}

__global__ void __launch_bounds__(256) sell_permute_kernel(const int32_t* __restrict__ slot_of, int64_t total,
                                                           const float* __restrict__ src, float* __restrict__ dst) {
    for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < total; d += (int64_t)gridDim.x * blockDim.x) {
        const int s = slot_of[d];
        dst[d] = s >= 0 ? src[s] : 0.f;
    }
}

// ---- aggregation -------------------------------------------------------------------------------------------
struct SellPeerOut {
    float* out[GG_PEER_MAX];
    int per;
};

struct SellArgs {
    const uint32_t* chunk_ptr;
    int chunks;
    const int4* idx4;
    const float4* w4;
    const int32_t* vdst;
    const int32_t* rowptr;   // original CSR (mean: degree of the whole row)
    const int32_t* hub_rows;
    const int32_t* hub_pptr;
    int hubs;
    const float* x;
    int64_t ldx;
    float* out;
    int64_t ldo;
    int f;
    int reduce;
    const float* x_self;
    int64_t ld_self;
    float self_scale;
    const float* bias;
    const float* r1_s;
    const float* r1_v;
    const float* r2_s;
    const float* r2_v;
    float* partial;   // [partial rows, f]
    int* counter;
    int l2_hint;
    int hub_hint;     // bit 30 of an index marks a hub source (gg_sell_hub_hint): hubs evict_last, the rest evict_first
};

template <bool PEER>
__device__ __forceinline__ float4* sell_out_row(const SellArgs& a, const SellPeerOut& po, int row) {
    if (PEER) {
        const int o = row / po.per;
        return reinterpret_cast<float4*>(po.out[o] + (int64_t)(row - o * po.per) * a.ldo);
    }
    return reinterpret_cast<float4*>(a.out + (int64_t)row * a.ldo);
}

template <bool PEER>
__device__ __forceinline__ void sell_epilogue(const SellArgs& a, const SellPeerOut& po, int row, int gl, float4 r,
                                              uint64_t pol_stream) {
    if (a.reduce == GG_MEAN) {
        const int deg = __ldg(a.rowptr + row + 1) - __ldg(a.rowptr + row);
        const float inv = deg > 0 ? 1.0f / (float)deg : 1.0f;
        r.x *= inv; r.y *= inv; r.z *= inv; r.w *= inv;
    }
    if (a.x_self) fma4(r, a.self_scale, __ldg(reinterpret_cast<const float4*>(a.x_self + (int64_t)row * a.ld_self) + gl));
    if (a.bias) add4(r, __ldg(reinterpret_cast<const float4*>(a.bias) + gl));
    if (a.r1_s) fma4(r, __ldg(a.r1_s + row), __ldg(reinterpret_cast<const float4*>(a.r1_v) + gl));
    if (a.r2_s) fma4(r, __ldg(a.r2_s + row), __ldg(reinterpret_cast<const float4*>(a.r2_v) + gl));
    stg_f4_hint(sell_out_row<PEER>(a, po, row) + gl, r, pol_stream);
}

__device__ __forceinline__ int4 ldg_nc_i4_hint(const int4* p, uint64_t pol) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}

template <int G, bool WEIGHTED, bool PEER>
__global__ void __launch_bounds__(kSellThreads, G >= 16 ? 3 : 4)
    spmm_sell_kernel(const __grid_constant__ SellArgs a, const __grid_constant__ SellPeerOut po) {
    constexpr int S = 32 / G;              // rows per pass
    constexpr int P = kSellRows / S;       // passes per chunk (= accumulators per lane)
    constexpr int PAIRS = P >= 2 ? P / 2 : 1;
    const int lane = threadIdx.x & 31;
    const int grp = lane / G, gl = lane % G;
    const int nvec = a.f >> 2;
    const bool act = gl < nvec;
    // lanes past the row's last vector (f/4 < G) gather a duplicate of it and never store: no predicates
    const char* __restrict__ xg = reinterpret_cast<const char*>(a.x) + (act ? gl : nvec - 1) * 16;
    uint32_t row_bytes = (uint32_t)a.ldx * 4u;
    // opaque to the compiler: under the 64-register cap it otherwise RE-DERIVES both values (tid, f, ldx, a.x from the
    // constant bank: ~15 instructions) in front of every gather instead of keeping three registers live
    asm volatile("" : "+l"(xg), "+r"(row_bytes));
    const uint64_t pol_keep = l2_policy(a.l2_hint ? 1 : 0), pol_stream = l2_policy(a.l2_hint ? 2 : 0);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int4 none4 = make_int4(-1, -1, -1, -1);
    const bool hub_hint = a.hub_hint != 0;
    auto gather = [&](int j) {
        float4 v = zero4;
        if (j >= 0) {
            const uint64_t pol = (hub_hint && !(j & kSellHubBit)) ? pol_stream : pol_keep;
            v = ldg_nc_f4_hint(reinterpret_cast<const float4*>(xg + (uint64_t)(uint32_t)(j & ~kSellHubBit) * row_bytes), pol);
        }
        return v;
    };

    int chunk = 0;
    if (lane == 0) chunk = atomicAdd(a.counter, 1);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    while (chunk < a.chunks) {
        int next = 0;
        if (lane == 0) next = atomicAdd(a.counter, 1);   // in flight while this chunk is processed
        const uint32_t base = __ldg(a.chunk_ptr + chunk);
        const int nun = (int)(__ldg(a.chunk_ptr + chunk + 1) - base) / S;   // units per group (multiple of P)
        const int npairs = (nun + 1) >> 1;
        int dst[P];
#pragma unroll
        for (int p = 0; p < P; ++p) dst[p] = __ldg(a.vdst + (int64_t)chunk * kSellRows + p * S + grp);
        const int4* __restrict__ ip = a.idx4 + base + grp;
        const float4* __restrict__ wp = a.w4 + base + grp;
        float4 acc[P];
#pragma unroll
        for (int p = 0; p < P; ++p) acc[p] = zero4;

        int4 ia = none4, ib = none4;
        if (nun > 0) {
            ia = ldg_nc_i4_hint(ip, pol_stream);
            if (P >= 2 || nun > 1) ib = ldg_nc_i4_hint(ip + S, pol_stream);
        }
        for (int t0 = 0; t0 < npairs; t0 += PAIRS) {
#pragma unroll
            for (int jj = 0; jj < PAIRS; ++jj) {
                const int t = t0 + jj;
                // the next pair's index units are requested before this pair's gathers
                int4 na = none4, nb = none4;
                if (t + 1 < npairs) {
                    const int u = 2 * (t + 1);
                    na = ldg_nc_i4_hint(ip + (int64_t)u * S, pol_stream);
                    if (P >= 2 || u + 1 < nun) nb = ldg_nc_i4_hint(ip + (int64_t)(u + 1) * S, pol_stream);
                }
                // this pair's weights travel with its gathers (padding entries: idx = -1 gathers nothing, v = 0)
                float4 wa = zero4, wb = zero4;
                if (WEIGHTED) {
                    wa = ldg_nc_f4_hint(wp + (int64_t)(2 * t) * S, pol_stream);
                    if (P >= 2 || 2 * t + 1 < nun) wb = ldg_nc_f4_hint(wp + (int64_t)(2 * t + 1) * S, pol_stream);
                }
                const float4 v0 = gather(ia.x), v1 = gather(ia.y), v2 = gather(ia.z), v3 = gather(ia.w);
                const float4 v4 = gather(ib.x), v5 = gather(ib.y), v6 = gather(ib.z), v7 = gather(ib.w);
                float4& A = acc[(2 * jj) % P];
                float4& B = acc[(2 * jj + 1) % P];
                if (WEIGHTED) {
                    fma4(A, wa.x, v0); fma4(A, wa.y, v1); fma4(A, wa.z, v2); fma4(A, wa.w, v3);
                    fma4(B, wb.x, v4); fma4(B, wb.y, v5); fma4(B, wb.z, v6); fma4(B, wb.w, v7);
                } else {
                    add4(A, v0); add4(A, v1); add4(A, v2); add4(A, v3);
                    add4(B, v4); add4(B, v5); add4(B, v6); add4(B, v7);
                }
                ia = na; ib = nb;
            }
        }
        if (act) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const int d = dst[p];
                if (d >= 0) sell_epilogue<PEER>(a, po, d, gl, acc[p], pol_stream);
                else if (d != kSellNoRow)
                    reinterpret_cast<float4*>(a.partial + (int64_t)(-d - 1) * a.f)[gl] = acc[p];
            }
        }
        chunk = __shfl_sync(0xffffffffu, next, 0);
    }
}

// split rows: out[row] = epilogue(partial[p0] + ... + partial[p1 - 1]) in segment order; one thread per (row, vector)
template <bool PEER>
__global__ void __launch_bounds__(256)
    spmm_sell_fixup_kernel(const __grid_constant__ SellArgs a, const __grid_constant__ SellPeerOut po) {
    const int nvec = a.f >> 2;
    const int64_t total = (int64_t)a.hubs * nvec;
    const uint64_t pol = l2_policy(0);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int h = (int)(e / nvec), gl = (int)(e - (int64_t)h * nvec);
        const int p0 = __ldg(a.hub_pptr + h), p1 = __ldg(a.hub_pptr + h + 1);
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = p0; p < p1; ++p) add4(r, reinterpret_cast<const float4*>(a.partial + (int64_t)p * a.f)[gl]);
        sell_epilogue<PEER>(a, po, __ldg(a.hub_rows + h), gl, r, pol);
    }
}

template <int G, bool WEIGHTED, bool PEER>
static void launch_sell_one(const SellArgs& a, const SellPeerOut& po, cudaStream_t st) {
    int grid = (int)ceil_div(a.chunks, kSellThreads / 32);
    constexpr int kPerSm = G >= 16 ? 3 : 4;
    if (grid > kNumSMs * kPerSm) grid = kNumSMs * kPerSm;
    if (grid < 1) grid = 1;
    spmm_sell_kernel<G, WEIGHTED, PEER><<<grid, kSellThreads, 0, st>>>(a, po);
    count_launch();
    if (a.hubs > 0) {
        spmm_sell_fixup_kernel<PEER><<<sell_grid((int64_t)a.hubs * (a.f >> 2), 256), 256, 0, st>>>(a, po);
        count_launch();
    }
}

template <int G>
static void launch_sell(const SellArgs& a, const SellPeerOut& po, bool peer, cudaStream_t st) {
    if (a.w4) {
        if (peer) launch_sell_one<G, true, true>(a, po, st);
        else launch_sell_one<G, true, false>(a, po, st);
    } else {
        if (peer) launch_sell_one<G, false, true>(a, po, st);
        else launch_sell_one<G, false, false>(a, po, st);
    }
}

static inline bool sell_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace gg

using namespace gg;

extern "C" {

int64_t gg_sell_vrow_capacity(int64_t num_rows, int64_t num_slots, int seg) {
    if (seg < 4) return 0;
    int64_t v = num_rows + num_slots / seg + 1;
    return ceil_div(v, kSellRows) * kSellRows;
}

int64_t gg_sell_unit_capacity(int64_t num_rows, int64_t num_slots, int seg) {
    // padded slots <= slots + 3 per virtual row (rounding to 4) + 7 * seg (the chunks' length ranges telescope)
    const int64_t vcap = gg_sell_vrow_capacity(num_rows, num_slots, seg);
    return ceil_div(num_slots + 3 * vcap + 7 * (int64_t)seg, 4) + kSellRows;
}

int64_t gg_sell_split_capacity(int64_t num_slots, int seg) { return seg < 4 ? 0 : num_slots / seg + 1; }

size_t gg_sell_build_workspace_bytes(int64_t num_rows, int64_t num_slots, int seg) {
    const int64_t vcap = gg_sell_vrow_capacity(num_rows, num_slots, seg);
    size_t b = 0;
    b += 2 * align_up((size_t)(num_rows + 1) * 4, 256);              // cnt -> vstart, flag -> hstart
    b += 6 * align_up((size_t)vcap * 4, 256);                        // keys, vals, keys_sorted, order, voff, vdst_tmp
    b += align_up((size_t)(vcap / kSellRows + 1) * 4, 256);          // units (scanned in place into chunk_ptr's source)
    b += align_up(scan_workspace_bytes(), 256);
    b += align_up(gg_sort_pairs_workspace_bytes(vcap), 256);
    return b + 256;
}

int gg_sell_build(const int32_t* rowptr, const int32_t* nbr, int64_t num_rows, int64_t num_slots, int seg,
                  uint32_t* chunk_ptr, int32_t* idx, int32_t* slot_of, int32_t* vdst, int32_t* hub_rows,
                  int32_t* hub_pptr, int32_t* info, void* workspace, size_t workspace_bytes, gg_stream_t stream) {
    GG_REQUIRE(num_rows >= 0 && num_slots >= 0, "gg_sell_build: negative size");
    GG_REQUIRE(seg >= 4 && seg <= 4096 && seg % 4 == 0, "gg_sell_build: seg=%d must be a multiple of 4 in [4, 4096]", seg);
    GG_REQUIRE(rowptr && chunk_ptr && idx && slot_of && vdst && hub_rows && hub_pptr && info && workspace,
               "gg_sell_build: null pointer");
    GG_REQUIRE(num_slots == 0 || nbr, "gg_sell_build: null neighbour array");
    const int64_t vcap = gg_sell_vrow_capacity(num_rows, num_slots, seg);
    GG_REQUIRE(vcap < ((int64_t)1 << 31) && gg_sell_unit_capacity(num_rows, num_slots, seg) < ((int64_t)1 << 29),
               "gg_sell_build: layout too large for 32-bit offsets");
    if (workspace_bytes < gg_sell_build_workspace_bytes(num_rows, num_slots, seg)) {
        set_error("gg_sell_build: workspace %zu < %zu", workspace_bytes,
                  gg_sell_build_workspace_bytes(num_rows, num_slots, seg));
        return GG_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    const int64_t chunks = vcap / kSellRows;
    Carver c(workspace);
    uint32_t* vstart = c.take<uint32_t>(num_rows + 1);
    uint32_t* hstart = c.take<uint32_t>(num_rows + 1);
    uint32_t* keys = c.take<uint32_t>(vcap);
    uint32_t* vals = c.take<uint32_t>(vcap);
    uint32_t* keys_sorted = c.take<uint32_t>(vcap);
    uint32_t* order = c.take<uint32_t>(vcap);
    int32_t* voff = c.take<int32_t>(vcap);
    int32_t* vdst_tmp = c.take<int32_t>(vcap);
    uint32_t* units = c.take<uint32_t>(chunks + 1);
    void* scan_ws = c.take<char>(scan_workspace_bytes());
    void* sort_ws = c.take<char>(gg_sort_pairs_workspace_bytes(vcap));

    GG_CUDA(cudaMemsetAsync(info, 0, 8 * sizeof(int32_t), st));
    sell_count_kernel<<<sell_grid(num_rows + 1, 256), 256, 0, st>>>(rowptr, num_rows, seg, vstart, hstart);
    GG_LAUNCHED();
    int rc = exclusive_scan_u32(vstart, vstart, num_rows + 1, scan_ws, st);
    if (rc != GG_OK) return rc;
    rc = exclusive_scan_u32(hstart, hstart, num_rows + 1, scan_ws, st);
    if (rc != GG_OK) return rc;
    sell_emit_kernel<<<sell_grid(num_rows + 1, 256), 256, 0, st>>>(rowptr, num_rows, seg, vcap, vstart, hstart, keys, vals,
                                                                  voff, vdst_tmp, hub_rows, hub_pptr, info);
    GG_LAUNCHED();
    rc = gg_sort_pairs_u32(keys, vals, keys_sorted, order, vcap, sell_bits_for(seg + 1), sort_ws,
                           gg_sort_pairs_workspace_bytes(vcap), stream);
    if (rc != GG_OK) return rc;
    sell_units_kernel<<<sell_grid(chunks + 1, 256), 256, 0, st>>>(keys_sorted, chunks, seg, units);
    GG_LAUNCHED();
    rc = exclusive_scan_u32(units, chunk_ptr, chunks + 1, scan_ws, st);
    if (rc != GG_OK) return rc;
    sell_fill_kernel<<<sell_grid(chunks, 8), 256, 0, st>>>(nbr, keys_sorted, order, voff, vdst_tmp, chunk_ptr, chunks, seg,
                                                           idx, slot_of, vdst, info);
    GG_LAUNCHED();
    return GG_OK;
}

size_t gg_sell_hub_hint_workspace_bytes(int64_t num_nodes) {
    return 512 + align_up((size_t)(num_nodes > 0 ? num_nodes : 1) * 4, 256) + align_up((size_t)kSellFreqBins * 4, 256);
}

int gg_sell_hub_hint(const int32_t* nbr, int64_t num_slots, int64_t num_nodes, const int32_t* idx, int64_t total,
                     int64_t hubs, int32_t* idx_hint, void* workspace, size_t workspace_bytes, gg_stream_t stream) {
    GG_REQUIRE(num_slots >= 0 && num_nodes >= 0 && total >= 0 && hubs >= 0, "gg_sell_hub_hint: negative size");
    GG_REQUIRE(num_nodes < kSellHubBit, "gg_sell_hub_hint: node ids need bit 30");
    if (total == 0) return GG_OK;
    GG_REQUIRE(idx && idx_hint && workspace && (num_slots == 0 || nbr), "gg_sell_hub_hint: null pointer");
    if (workspace_bytes < gg_sell_hub_hint_workspace_bytes(num_nodes)) {
        set_error("gg_sell_hub_hint: workspace %zu < %zu", workspace_bytes, gg_sell_hub_hint_workspace_bytes(num_nodes));
        return GG_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    Carver c(workspace);
    int32_t* threshold = c.take<int32_t>(64);
    int32_t* freq = c.take<int32_t>(num_nodes > 0 ? num_nodes : 1);
    int32_t* bins = c.take<int32_t>(kSellFreqBins);
    GG_CUDA(cudaMemsetAsync(workspace, 0, gg_sell_hub_hint_workspace_bytes(num_nodes), st));
    if (num_slots > 0) {
        sell_freq_kernel<<<sell_grid(num_slots, 256 * 4), 256, 0, st>>>(nbr, num_slots, freq);
        GG_LAUNCHED();
    }
    sell_freq_hist_kernel<<<sell_grid(num_nodes, 256 * 16), 256, 0, st>>>(freq, num_nodes, bins);
    GG_LAUNCHED();
    sell_freq_threshold_kernel<<<1, 256, 0, st>>>(bins, hubs, threshold);
    GG_LAUNCHED();
    sell_hub_flag_kernel<<<sell_grid(total, 256 * 4), 256, 0, st>>>(idx, total, freq, threshold, idx_hint);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_sell_permute_f32(const int32_t* slot_of, int64_t total, const float* src, float* dst, gg_stream_t stream) {
    GG_REQUIRE(total >= 0, "gg_sell_permute_f32: negative size");
    if (total == 0) return GG_OK;
    GG_REQUIRE(slot_of && src && dst, "gg_sell_permute_f32: null pointer");
    sell_permute_kernel<<<sell_grid(total, 256 * 4), 256, 0, as_stream(stream)>>>(slot_of, total, src, dst);
    GG_LAUNCHED();
    return GG_OK;
}

size_t gg_spmm_sell_workspace_bytes(int64_t partial_rows, int64_t f) {
    return 512 + align_up((size_t)(partial_rows > 0 ? partial_rows : 0) * (size_t)f * sizeof(float), 256);
}

int gg_spmm_sell_f32(const uint32_t* chunk_ptr, int64_t chunks, const int32_t* idx, const float* w_sell,
                     const int32_t* vdst, const int32_t* rowptr, const int32_t* hub_rows, const int32_t* hub_pptr,
                     int64_t hubs, int64_t partial_rows, const float* x, int64_t ldx, float* out, int64_t ldo,
                     float* const* out_peers_host, int world, int64_t rows_per_rank, int64_t num_rows, int64_t f,
                     int reduce, const float* x_self, int64_t ld_self, float self_scale, const float* bias,
                     const float* r1_s, const float* r1_v, const float* r2_s, const float* r2_v, void* workspace,
                     size_t workspace_bytes, int flags, gg_stream_t stream) {
    GG_REQUIRE(num_rows >= 0 && f >= 0 && chunks >= 0 && hubs >= 0 && partial_rows >= 0, "gg_spmm_sell_f32: negative size");
    GG_REQUIRE(reduce == GG_SUM || reduce == GG_MEAN, "gg_spmm_sell_f32: reduce=%d", reduce);
    GG_REQUIRE((!r1_s || (r1_v && sell_al16(r1_v))) && (!r2_s || (r2_v && sell_al16(r2_v))),
               "gg_spmm_sell_f32: rank-1 term without its (16-byte aligned) vector");
    if (num_rows == 0 || f == 0) return GG_OK;
    if (f % 4 != 0 || f > 128) {
        set_error("gg_spmm_sell_f32: needs f %% 4 == 0 and f <= 128 (got %lld)", (long long)f);
        return GG_ERR_UNSUPPORTED;
    }
    const bool peer = out_peers_host != nullptr;
    GG_REQUIRE(chunk_ptr && idx && vdst && rowptr && x && workspace && (peer || out) && (hubs == 0 || (hub_rows && hub_pptr)),
               "gg_spmm_sell_f32: null pointer");
    GG_REQUIRE(chunks < ((int64_t)1 << 31) && num_rows < ((int64_t)1 << 31), "gg_spmm_sell_f32: sizes out of range");
    GG_REQUIRE(ldx >= f && ldx < ((int64_t)1 << 30) && ldx % 4 == 0 && ldo % 4 == 0 && sell_al16(x) && sell_al16(idx) &&
                   (!w_sell || sell_al16(w_sell)) && (!x_self || (ld_self % 4 == 0 && sell_al16(x_self))) &&
                   (!bias || sell_al16(bias)),
               "gg_spmm_sell_f32: rows and unit arrays must be 16-byte aligned");
    SellPeerOut po{};
    po.per = 1;
    if (peer) {
        GG_REQUIRE(world >= 1 && world <= GG_PEER_MAX && rows_per_rank >= 1 && rows_per_rank < ((int64_t)1 << 31) &&
                       rows_per_rank * world >= num_rows,
                   "gg_spmm_sell_f32: world=%d rows_per_rank=%lld do not cover %lld rows", world,
                   (long long)rows_per_rank, (long long)num_rows);
        for (int i = 0; i < world; ++i) {
            GG_REQUIRE(out_peers_host[i] && sell_al16(out_peers_host[i]), "gg_spmm_sell_f32: peer block %d null or misaligned", i);
            po.out[i] = out_peers_host[i];
        }
        po.per = (int)rows_per_rank;
    } else {
        GG_REQUIRE(sell_al16(out), "gg_spmm_sell_f32: out must be 16-byte aligned");
    }
    if (workspace_bytes < gg_spmm_sell_workspace_bytes(partial_rows, f)) {
        set_error("gg_spmm_sell_f32: workspace %zu < %zu", workspace_bytes, gg_spmm_sell_workspace_bytes(partial_rows, f));
        return GG_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    Carver c(workspace);
    int* counter = c.take<int>(64);
    float* partial = c.take<float>((size_t)partial_rows * f);
    GG_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
    SellArgs a{chunk_ptr, (int)chunks, reinterpret_cast<const int4*>(idx), reinterpret_cast<const float4*>(w_sell), vdst,
               rowptr, hub_rows, hub_pptr, (int)hubs, x, ldx, out, ldo, (int)f, reduce, x_self, ld_self, self_scale,
               bias, r1_s, r1_v, r2_s, r2_v, partial, counter, (flags & 4) ? 1 : 0, (flags & 8) ? 1 : 0};
    const int nvec = (int)(f / 4);
    if (nvec <= 4) launch_sell<4>(a, po, peer, st);
    else if (nvec <= 8) launch_sell<8>(a, po, peer, st);
    else if (nvec <= 16) launch_sell<16>(a, po, peer, st);
    else launch_sell<32>(a, po, peer, st);
    GG_CUDA(cudaPeekAtLastError());
    return GG_OK;
}

}  // extern "C"
