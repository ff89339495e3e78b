// On-device batch collation: DeepSNAP's Batch.collate() concatenates the graphs of a mini-batch
// block-diagonally on the HOST, every iteration (ref: graphgym/loader.py:245-250 -> deepsnap Batch.collate;
// graphgym/train.py:21 then copies the batch to the device), and Preprocess concatenates the augmented feature
// blocks column-wise (ref: graphgym/models/feature_augment.py:329-333).  Both are segmented copies:
//   gg_collate_index_i64   out[dst_off_g + k] = (src_g ? src_g[k] : 0) + add_g     node / edge index arrays with the
//                          graph's node offset added (edge_index rows, node_id_index), the `batch` vector (src null,
//                          add = g), label vectors (add = 0)
//   gg_collate_rows_f32    out[dst_row_g + i, col_off : col_off + f_g] = src_g[i, 0 : f_g]   feature blocks stacked
//                          by graph and placed side by side by key; int64 / uint8 sources are converted to float
//                          (`.float()` of Preprocess)
// ONE launch per output array whatever the number of graphs: the segment table lives in device memory.
#include "common.cuh"

namespace gg {

// first g with ends[g] > e  (ends = inclusive prefix of the segment sizes)
__device__ __forceinline__ int seg_find(const int64_t* __restrict__ ends, int num, int64_t e) {
    int lo = 0, hi = num - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ends[mid] > e) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

__global__ void __launch_bounds__(256) collate_index_kernel(const gg_index_segment* __restrict__ seg,
                                                            const int64_t* __restrict__ ends, int num, int64_t total,
                                                            int64_t* __restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int g = seg_find(ends, num, e);
        const gg_index_segment s = seg[g];
        const int64_t k = e - (ends[g] - s.count);
        out[s.dst_offset + k] = (s.src ? s.src[k] : 0) + s.add;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) collate_rows_kernel(const gg_rows_segment* __restrict__ seg,
                                                           const int64_t* __restrict__ ends, int num, int64_t total,
                                                           float* __restrict__ out, int64_t ldo, int64_t col_off) {
    // flattened over the elements of all segments: ends = inclusive prefix of rows_g * f_g
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int g = seg_find(ends, num, e);
        const gg_rows_segment s = seg[g];
        const int64_t k = e - (ends[g] - s.rows * s.f);
        const int64_t i = k / s.f, c = k - i * s.f;
        out[(s.dst_row + i) * ldo + col_off + c] = (float)static_cast<const T*>(s.src)[i * s.ld + c];
    }
}

__global__ void __launch_bounds__(256) seg_ends_index_kernel(const gg_index_segment* __restrict__ seg, int num,
                                                             int64_t* __restrict__ ends) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {   // tables hold a mini-batch's graphs: tens to hundreds of entries
        int64_t acc = 0;
        for (int g = 0; g < num; ++g) ends[g] = (acc += seg[g].count);
    }
}
__global__ void __launch_bounds__(256) seg_ends_rows_kernel(const gg_rows_segment* __restrict__ seg, int num,
                                                            int64_t* __restrict__ ends) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int64_t acc = 0;
        for (int g = 0; g < num; ++g) ends[g] = (acc += seg[g].rows * seg[g].f);
    }
}

static inline int collate_grid(int64_t total) {
    int64_t b = ceil_div(total, 256 * 4);
    if (b > (int64_t)kNumSMs * 8) b = (int64_t)kNumSMs * 8;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace gg

using namespace gg;

extern "C" {

int gg_collate_index_i64(const gg_index_segment* segments_dev, int num_segments, int64_t total, int64_t* out,
                         int64_t* scratch_ends_dev, gg_stream_t stream) {
    GG_REQUIRE(num_segments >= 0 && total >= 0, "gg_collate_index_i64: negative size");
    if (num_segments == 0 || total == 0) return GG_OK;
    GG_REQUIRE(segments_dev && out && scratch_ends_dev, "gg_collate_index_i64: null pointer");
    cudaStream_t st = as_stream(stream);
    seg_ends_index_kernel<<<1, 32, 0, st>>>(segments_dev, num_segments, scratch_ends_dev);
    GG_LAUNCHED();
    collate_index_kernel<<<collate_grid(total), 256, 0, st>>>(segments_dev, scratch_ends_dev, num_segments, total, out);
    GG_LAUNCHED();
    return GG_OK;
}

int gg_collate_rows_f32(const gg_rows_segment* segments_dev, int num_segments, int64_t total_elements, int src_dtype,
                        float* out, int64_t ldo, int64_t col_offset, int64_t* scratch_ends_dev, gg_stream_t stream) {
    GG_REQUIRE(num_segments >= 0 && total_elements >= 0 && col_offset >= 0, "gg_collate_rows_f32: negative size");
    GG_REQUIRE(src_dtype == GG_DTYPE_F32 || src_dtype == GG_DTYPE_I64 || src_dtype == GG_DTYPE_U8,
               "gg_collate_rows_f32: src_dtype=%d", src_dtype);
    if (num_segments == 0 || total_elements == 0) return GG_OK;
    GG_REQUIRE(segments_dev && out && scratch_ends_dev && ldo >= 1, "gg_collate_rows_f32: null pointer");
    cudaStream_t st = as_stream(stream);
    seg_ends_rows_kernel<<<1, 32, 0, st>>>(segments_dev, num_segments, scratch_ends_dev);
    GG_LAUNCHED();
    const int grid = collate_grid(total_elements);
    if (src_dtype == GG_DTYPE_F32)
        collate_rows_kernel<float><<<grid, 256, 0, st>>>(segments_dev, scratch_ends_dev, num_segments, total_elements, out, ldo, col_offset);
    else if (src_dtype == GG_DTYPE_I64)
        collate_rows_kernel<int64_t><<<grid, 256, 0, st>>>(segments_dev, scratch_ends_dev, num_segments, total_elements, out, ldo, col_offset);
    else
        collate_rows_kernel<uint8_t><<<grid, 256, 0, st>>>(segments_dev, scratch_ends_dev, num_segments, total_elements, out, ldo, col_offset);
    GG_LAUNCHED();
    return GG_OK;
}

}  // extern "C"
