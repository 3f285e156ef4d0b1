// Stand-alone dropout entry points (the encoder's dropout sites are fused into ln.cu / attn.cu, see dropout.cuh):
//   ur_dropout_rows_f32   X[r, :] *= mask(row position, :) in place  -- GRU item-embedding dropout, unirec/model/sequential/gru.py:29,
//                         and its backward (same call on the gradient rows)
//   ur_dropout_mask_f32   writes the multipliers themselves (tests feed them to the CPU oracle as explicit factors)
//   ur_rng_advance        rng[1] += 1 on the device (one new mask set per training step, CUDA-graph replayable)
#include "common.cuh"
#include "dropout.cuh"

namespace ur {

__global__ void __launch_bounds__(256) dropout_rows_kernel(float4* __restrict__ X, int64_t ld4, int64_t rows, int d4,
                                                           const int32_t* __restrict__ row_pos, int64_t pos_mul, int64_t pos_add,
                                                           const int32_t* __restrict__ rows_dev, const long long* __restrict__ rng,
                                                           float p, int site, int write_mask) {
    const DropCfg dc = drop_cfg(rng, p, site);
    const int64_t n_live = rows_dev ? min((int64_t)*rows_dev, rows) : rows;
    const int64_t total = n_live * d4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / d4;
        const int c = (int)(i - r * d4);
        const unsigned long long prow = (unsigned long long)(row_pos ? (int64_t)__ldg(row_pos + r) : r * pos_mul + pos_add);
        const float4 m = dc.on ? drop_mask4(dc, prow * d4 + c) : make_float4(1.f, 1.f, 1.f, 1.f);
        float4* x = X + r * ld4 + c;
        *x = write_mask ? m : f4_mul(*x, m);
    }
}

// attention-probability site: element ((bh)*L + i)*L + j
__global__ void __launch_bounds__(256) dropout_mask_flat_kernel(float* __restrict__ out, int64_t n, const long long* __restrict__ rng,
                                                                float p, int site) {
    const DropCfg dc = drop_cfg(rng, p, site);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = dc.on ? drop_mask1(dc, (unsigned long long)i) : 1.f;
}

__global__ void rng_advance_kernel(long long* rng) { rng[1] += 1; }

static inline unsigned ew_grid(int64_t n) {
    int64_t g = (n + 255) / 256;
    const int64_t cap = (int64_t)kNumSMs * 8;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace ur

extern "C" {

int ur_dropout_rows_f32(float* X, int64_t ld, int64_t rows, int d, const int32_t* row_pos, int64_t pos_mul, int64_t pos_add,
                        const int32_t* rows_dev, const int64_t* rng, float p, int site, void* stream) {
    if (d <= 0 || (d & 3) || (ld & 3) || p < 0.f || p >= 1.f) return UR_ERR_BAD_ARG;
    if (rows == 0 || p == 0.f || rng == nullptr) return UR_OK;
    ur::dropout_rows_kernel<<<ur::ew_grid(rows * (d / 4)), 256, 0, (cudaStream_t)stream>>>(
        (float4*)X, ld / 4, rows, d / 4, row_pos, pos_mul, pos_add, rows_dev, (const long long*)rng, p, site, 0);
    UR_RETURN_LAST_ERROR();
}

int ur_dropout_mask_f32(float* out, int64_t rows, int d, const int32_t* row_pos, int64_t pos_mul, int64_t pos_add,
                        const int64_t* rng, float p, int site, int flat, void* stream) {
    if (p < 0.f || p >= 1.f || rng == nullptr) return UR_ERR_BAD_ARG;
    if (rows == 0) return UR_OK;
    if (flat) {
        ur::dropout_mask_flat_kernel<<<ur::ew_grid(rows * d), 256, 0, (cudaStream_t)stream>>>(out, rows * d, (const long long*)rng, p, site);
    } else {
        if (d <= 0 || (d & 3)) return UR_ERR_BAD_ARG;
        ur::dropout_rows_kernel<<<ur::ew_grid(rows * (d / 4)), 256, 0, (cudaStream_t)stream>>>(
            (float4*)out, d / 4, rows, d / 4, row_pos, pos_mul, pos_add, nullptr, (const long long*)rng, p, site, 1);
    }
    UR_RETURN_LAST_ERROR();
}

int ur_rng_advance(int64_t* rng, void* stream) {
    if (rng == nullptr) return UR_ERR_BAD_ARG;
    ur::rng_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((long long*)rng);
    UR_RETURN_LAST_ERROR();
}

}  // extern "C"
