// Embedding-row kernels: gather (K1/K2), dense scatter-add (K11, exact-dense mode), sum-pool tower (K10).
// All HBM-bound: 128-bit coalesced accesses, several independent loads in flight per thread.
#include "common.cuh"

namespace ur {

// ---------------------------------------------------------------------------------------------
// out[i, :] = table[idx[i], :]      (bit-exact copy; reference: nn.Embedding forward,
// unirec/model/base/recommender.py:67,137)
// One thread moves one float4; UNROLL independent 16-byte loads per thread.
// ---------------------------------------------------------------------------------------------
template <int UNROLL>
__global__ void __launch_bounds__(256) gather_rows_kernel(const float4* __restrict__ table, const void* __restrict__ idx,
                                                          int idx64, int64_t n, int d4, float4* __restrict__ out) {
    const int64_t total = n * d4;
    int64_t base = ((int64_t)blockIdx.x * blockDim.x) * UNROLL + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * UNROLL;
    for (; base < total; base += stride) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            int64_t e = base + (int64_t)u * blockDim.x;
            if (e < total) {
                int64_t r = e / d4;
                int c = (int)(e - r * d4);
                int64_t id = load_index(idx, idx64, r);
                v[u] = ldg_stream(table + id * d4 + c);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            int64_t e = base + (int64_t)u * blockDim.x;
            if (e < total) stg_stream(out + e, v[u]);
        }
    }
}

// out[i, :] = float(table_bf16[idx[i], :]): gather from a BF16 copy of a table (half the HBM bytes per row; north_star's opt-in
// 1e-2 mode).  One thread moves 8 elements: one 16-byte load -> two float4 stores.  Exact widening (bf16 -> fp32 is a shift).
__global__ void __launch_bounds__(256) gather_rows_bf16_kernel(const uint4* __restrict__ table, const void* __restrict__ idx, int idx64,
                                                               int64_t n, int d8, float4* __restrict__ out) {
    const int64_t total = n * d8;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / d8;
        const int c = (int)(e - r * d8);
        const int64_t id = load_index(idx, idx64, r);
        const uint4 v = __ldg(table + id * d8 + c);
        float4 lo, hi;
        lo.x = __uint_as_float(v.x << 16); lo.y = __uint_as_float(v.x & 0xffff0000u);
        lo.z = __uint_as_float(v.y << 16); lo.w = __uint_as_float(v.y & 0xffff0000u);
        hi.x = __uint_as_float(v.z << 16); hi.y = __uint_as_float(v.z & 0xffff0000u);
        hi.z = __uint_as_float(v.w << 16); hi.w = __uint_as_float(v.w & 0xffff0000u);
        out[e * 2] = lo;
        out[e * 2 + 1] = hi;
    }
}

// ---------------------------------------------------------------------------------------------
// grad[idx[i], :] += src[i, :] * (coef ? coef[i / group] : 1), rows with idx == pad_id skipped.
// Exact-dense mode of K11 (what embedding_dense_backward produces), one RED.v4 per 16 bytes.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) scatter_add_rows_kernel(float* __restrict__ grad, const void* __restrict__ idx, int idx64,
                                                               int64_t n, int d4, const float4* __restrict__ src,
                                                               int64_t src_group, const float* __restrict__ coef,
                                                               int64_t coef_group, int64_t pad_id) {
    const int64_t total = n * d4;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = e / d4;
        int c = (int)(e - r * d4);
        int64_t id = load_index(idx, idx64, r);
        if (id == pad_id) continue;
        float4 v = __ldg(src + (r / src_group) * d4 + c);
        if (coef) v = f4_scale(v, __ldg(coef + r / coef_group));
        red_add_v4(grad + (id * d4 + c) * 4, v);
    }
}

// out[idx[e / idx_group] & mask] += src[e] for e < n  (scalar tables: item_bias / user_bias gradients, recommender.py:79-90).
// idx_group = 1: one index per entry (item bias); idx_group = N: one index per sample, its N entries summed (user bias).
__global__ void __launch_bounds__(256) scatter_add_scalar_kernel(float* __restrict__ out, const void* __restrict__ idx, int idx64,
                                                                 int64_t idx_group, int64_t mask, const float* __restrict__ src, int64_t n) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const float v = src[e];
        if (v != 0.f) atomicAdd(out + (load_index(idx, idx64, e / idx_group) & mask), v);
    }
}

// ---------------------------------------------------------------------------------------------
// Sum-pool tower (AvgHist / SVD++ / MF user side):
//   u[b,:] = (U ? U[user_id[b],:] : 0) + coeff[b] * sum_l E[seq[b,l],:],  coeff[b] = (len[b]+1)^-alpha
// reference: unirec/model/sequential/avghist.py:34-42, svdplusplus.py:31-39, recommender.py:42-44.
// One CTA (4 warps) per sample: warps split the L rows, each lane-group owns a float4 column slice.
// Never materialises [B,L,d].
// ---------------------------------------------------------------------------------------------
template <int D4>   // d / 4
__global__ void __launch_bounds__(128) pool_sum_kernel(const float4* __restrict__ E, const int32_t* __restrict__ seq, int L,
                                                       const int64_t* __restrict__ seq_len, float alpha,
                                                       const float4* __restrict__ U, const int64_t* __restrict__ user_id,
                                                       float4* __restrict__ out, float* __restrict__ coeff_out, int W, int r) {
    // W > 1: E / U are the row shards of rank r (rows id % W == r at local index id / W); rows of other ranks contribute nothing
    // here -- the per-rank partial sums are added by a reduce-scatter (sharding.py)
    constexpr int LPR = D4 < 32 ? D4 : 32;      // lanes per row
    constexpr int VPL = D4 / LPR;               // float4 per lane
    constexpr int RPW = 32 / LPR;               // rows per warp per step
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane / LPR, col = lane % LPR;
    const int32_t* s = seq + (int64_t)b * L;
    float4 acc[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int rows_per_step = 4 * RPW;
    for (int l0 = 0; l0 < L; l0 += rows_per_step * 4) {
        float4 t[4][VPL];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int l = l0 + u * rows_per_step + warp * RPW + sub;
            int32_t id = l < L ? __ldg(s + l) : 0;
            ok[u] = id > 0 && (id % W) == r;          // the padding row holds zeros: skipped (adds nothing)
#pragma unroll
            for (int v = 0; v < VPL; ++v)
                if (ok[u]) t[u][v] = ldg_stream(E + (int64_t)(id / W) * D4 + v * LPR + col);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < VPL; ++v)
                if (ok[u]) acc[v] = f4_add(acc[v], t[u][v]);
    }
    // reduce the RPW sub-rows inside the warp, then the 4 warps through shared memory
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) {
            acc[v].x += __shfl_xor_sync(0xffffffffu, acc[v].x, o);
            acc[v].y += __shfl_xor_sync(0xffffffffu, acc[v].y, o);
            acc[v].z += __shfl_xor_sync(0xffffffffu, acc[v].z, o);
            acc[v].w += __shfl_xor_sync(0xffffffffu, acc[v].w, o);
        }
    }
    __shared__ float4 sm[4][D4];
    if (sub == 0) {
#pragma unroll
        for (int v = 0; v < VPL; ++v) sm[warp][v * LPR + col] = acc[v];
    }
    __syncthreads();
    const float cf = powf((float)(seq_len[b] + 1), -alpha);
    if (threadIdx.x == 0 && coeff_out) coeff_out[b] = cf;
    for (int c = threadIdx.x; c < D4; c += blockDim.x) {
        float4 o = f4_scale(f4_add(f4_add(sm[0][c], sm[1][c]), f4_add(sm[2][c], sm[3][c])), cf);
        if (U) {
            const int64_t uid = user_id[b];
            if (uid % W == r) o = f4_add(o, __ldg(U + (uid / W) * D4 + c));
        }
        out[(int64_t)b * D4 + c] = o;
    }
}

}  // namespace ur

extern "C" {

int ur_gather_rows_f32(const float* table, int64_t n_rows, int d, const void* idx, int idx_bits, int64_t n, float* out,
                       void* stream) {
    if (d <= 0 || (d & 3) || (idx_bits != 32 && idx_bits != 64)) return UR_ERR_BAD_ARG;
    if (n == 0) return UR_OK;
    (void)n_rows;
    const int d4 = d / 4;
    const int64_t total = n * d4;
    constexpr int UNROLL = 4;
    int64_t blocks = (total + 256 * UNROLL - 1) / (256 * UNROLL);
    const int64_t cap = (int64_t)ur::kNumSMs * 8 * 4;
    if (blocks > cap) blocks = cap;
    ur::gather_rows_kernel<UNROLL><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(table), idx, idx_bits == 64, n, d4, reinterpret_cast<float4*>(out));
    UR_RETURN_LAST_ERROR();
}

int ur_scatter_add_rows_f32(float* grad, int64_t n_rows, int d, const void* idx, int idx_bits, int64_t n, const float* src,
                            int64_t src_group, const float* coef, int64_t coef_group, int64_t pad_id, void* stream) {
    if (d <= 0 || (d & 3) || (idx_bits != 32 && idx_bits != 64) || src_group <= 0 || coef_group <= 0) return UR_ERR_BAD_ARG;
    if (n == 0) return UR_OK;
    (void)n_rows;
    const int d4 = d / 4;
    int64_t blocks = (n * d4 + 255) / 256;
    const int64_t cap = (int64_t)ur::kNumSMs * 8 * 4;
    if (blocks > cap) blocks = cap;
    ur::scatter_add_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        grad, idx, idx_bits == 64, n, d4, reinterpret_cast<const float4*>(src), src_group, coef, coef_group, pad_id);
    UR_RETURN_LAST_ERROR();
}

int ur_scatter_add_scalar_f32(float* out, const void* idx, int idx_bits, int64_t idx_group, int64_t idx_mask, const float* src, int64_t n,
                              void* stream) {
    if ((idx_bits != 32 && idx_bits != 64) || idx_group < 1) return UR_ERR_BAD_ARG;
    if (n == 0) return UR_OK;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)ur::kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    ur::scatter_add_scalar_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(out, idx, idx_bits == 64, idx_group, idx_mask, src, n);
    UR_RETURN_LAST_ERROR();
}

int ur_gather_rows_bf16(const void* table_bf16, int64_t n_rows, int d, const void* idx, int idx_bits, int64_t n, float* out, void* stream) {
    if (d <= 0 || (d & 7) || (idx_bits != 32 && idx_bits != 64)) return UR_ERR_BAD_ARG;
    if (n == 0) return UR_OK;
    (void)n_rows;
    int64_t blocks = (n * (d / 8) + 255) / 256;
    const int64_t cap = (int64_t)ur::kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    ur::gather_rows_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)table_bf16, idx, idx_bits == 64, n, d / 8,
                                                                                    (float4*)out);
    UR_RETURN_LAST_ERROR();
}

int ur_pool_sum_fwd_f32(const float* table, int d, const int32_t* item_seq, int64_t B, int L, const int64_t* item_seq_len,
                        float alpha, const float* user_table, const int64_t* user_id, float* user_emb, float* coeff_out,
                        int world, int rank, void* stream) {
    if (world < 1 || rank < 0 || rank >= world) return UR_ERR_BAD_ARG;
    if (B == 0) return UR_OK;
    auto E = reinterpret_cast<const float4*>(table);
    auto U = reinterpret_cast<const float4*>(user_table);
    auto O = reinterpret_cast<float4*>(user_emb);
    cudaStream_t st = (cudaStream_t)stream;
    switch (d) {
#define UR_CASE(D)                                                                                                         \
    case D:                                                                                                                \
        ur::pool_sum_kernel<D / 4><<<(unsigned)B, 128, 0, st>>>(E, item_seq, L, item_seq_len, alpha, U, user_id, O, coeff_out, world, rank); \
        break;
        UR_CASE(16) UR_CASE(32) UR_CASE(64) UR_CASE(128) UR_CASE(256) UR_CASE(512)
#undef UR_CASE
        default: return UR_ERR_UNSUPPORTED;
    }
    UR_RETURN_LAST_ERROR();
}

}  // extern "C"
