// Row gather / scatter-add / ReLU-gradient helpers for the ID branches (ref: idconv.py:372-375).
#include "common.cuh"

namespace gg {
constexpr int kRowThreads = 256;

__global__ void __launch_bounds__(kRowThreads)
    gather_rows_kernel(const float* __restrict__ x, int64_t ldx, const int64_t* __restrict__ id,
                       int64_t m, int64_t f, float* __restrict__ out, int64_t ldo) {
    int64_t total = m * f;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = e / f, c = e % f;
        out[r * ldo + c] = __ldg(x + id[r] * ldx + c);
    }
}
__global__ void __launch_bounds__(kRowThreads)
    scatter_add_rows_kernel(const float* __restrict__ x, int64_t ldx, const int64_t* __restrict__ id,
                            int64_t m, int64_t f, float* __restrict__ out, int64_t ldo) {
    int64_t total = m * f;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = e / f, c = e % f;
        atomicAdd(out + id[r] * ldo + c, __ldg(x + r * ldx + c));
    }
}
__global__ void __launch_bounds__(kRowThreads)
    relu_grad_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ y, int64_t ldy,
                     int64_t n, int64_t f, float* __restrict__ out, int64_t ldo) {
    int64_t total = n * f;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = e / f, c = e % f;
        out[r * ldo + c] = __ldg(y + r * ldy + c) > 0.f ? __ldg(g + r * ldg + c) : 0.f;
    }
}
static inline int row_grid(int64_t total) {
    int64_t g = ceil_div(total, kRowThreads);
    int64_t cap = (int64_t)kNumSMs * 16;
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}
}  // namespace gg

using namespace gg;

extern "C" {
int gg_gather_rows_f32(const float* x, int64_t ldx, const int64_t* id, int64_t m, int64_t f, float* out,
                       int64_t ldo, gg_stream_t stream) {
    GG_REQUIRE(m >= 0 && f >= 0, "gg_gather_rows_f32: negative size");
    if (m == 0 || f == 0) return GG_OK;
    GG_REQUIRE(x && id && out && ldx >= f && ldo >= f, "gg_gather_rows_f32: bad operands");
    gather_rows_kernel<<<row_grid(m * f), kRowThreads, 0, as_stream(stream)>>>(x, ldx, id, m, f, out, ldo);
    GG_LAUNCHED();
    return GG_OK;
}
int gg_scatter_add_rows_f32(const float* x, int64_t ldx, const int64_t* id, int64_t m, int64_t f,
                            float* out, int64_t ldo, gg_stream_t stream) {
    GG_REQUIRE(m >= 0 && f >= 0, "gg_scatter_add_rows_f32: negative size");
    if (m == 0 || f == 0) return GG_OK;
    GG_REQUIRE(x && id && out && ldx >= f && ldo >= f, "gg_scatter_add_rows_f32: bad operands");
    scatter_add_rows_kernel<<<row_grid(m * f), kRowThreads, 0, as_stream(stream)>>>(x, ldx, id, m, f, out,
                                                                                  ldo);
    GG_LAUNCHED();
    return GG_OK;
}
int gg_relu_grad_f32(const float* g, int64_t ldg, const float* y, int64_t ldy, int64_t n, int64_t f,
                     float* out, int64_t ldo, gg_stream_t stream) {
    GG_REQUIRE(n >= 0 && f >= 0, "gg_relu_grad_f32: negative size");
    if (n == 0 || f == 0) return GG_OK;
    GG_REQUIRE(g && y && out && ldg >= f && ldy >= f && ldo >= f, "gg_relu_grad_f32: bad operands");
    relu_grad_kernel<<<row_grid(n * f), kRowThreads, 0, as_stream(stream)>>>(g, ldg, y, ldy, n, f, out, ldo);
    GG_LAUNCHED();
    return GG_OK;
}
}
