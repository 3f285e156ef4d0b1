// Row-sharded table kernels (multi-GPU): rank r of W owns rows {id : id % W == r} at local index id / W.
// "Move queries, not rows": every rank sees ALL samples' ids and user vectors (all-gathered) and touches only the rows
// it owns, so per-rank HBM traffic stays that of one GPU's batch while only [samples, O(d)] states cross NVLink.
//
//   ur_shard_gather_rows   owned rows -> dense [n, d] buffer (zeros elsewhere)          -> reduce_scatter = sequence rows
//   ur_shard_localize      global ids -> local row ids (-1 where not owned / padding)   -> keys of the row-sparse update
// The owner-side scoring kernels live in shard_ring.cu.
#include "common.cuh"

namespace ur {

__global__ void __launch_bounds__(256) shard_gather_rows_kernel(const float4* __restrict__ table, const void* __restrict__ idx, int idx64,
                                                                int64_t n, int d4, int W, int r, float4* __restrict__ out) {
    const int64_t total = n * d4;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / d4;
        const int c = (int)(e - row * d4);
        const int64_t id = load_index(idx, idx64, row);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (id % W == r) v = ldg_stream(table + (id / W) * d4 + c);
        out[e] = v;
    }
}

__global__ void __launch_bounds__(256) shard_localize_kernel(const void* __restrict__ idx, int idx64, int64_t n, int W, int r,
                                                             int64_t pad_id, int32_t* __restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t id = load_index(idx, idx64, e);
        out[e] = (id != pad_id && id % W == r) ? (int32_t)(id / W) : -1;
    }
}

static inline unsigned grid_for(int64_t n, int per_block = 256, int max_waves = 8) {
    int64_t b = (n + per_block - 1) / per_block;
    const int64_t cap = (int64_t)kNumSMs * max_waves;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace ur

extern "C" {

int ur_shard_gather_rows_f32(const float* table_local, int d, const void* idx, int idx_bits, int64_t n, int world, int rank,
                             float* out, void* stream) {
    if (d <= 0 || (d & 3) || world < 1 || rank < 0 || rank >= world || (idx_bits != 32 && idx_bits != 64)) return UR_ERR_BAD_ARG;
    if (n == 0) return UR_OK;
    ur::shard_gather_rows_kernel<<<ur::grid_for(n * (d / 4)), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)table_local, idx, idx_bits == 64, n, d / 4, world, rank, (float4*)out);
    UR_RETURN_LAST_ERROR();
}

int ur_shard_localize(const void* idx, int idx_bits, int64_t n, int world, int rank, int64_t pad_id, int32_t* out, void* stream) {
    if (world < 1 || rank < 0 || rank >= world || (idx_bits != 32 && idx_bits != 64)) return UR_ERR_BAD_ARG;
    if (n == 0) return UR_OK;
    ur::shard_localize_kernel<<<ur::grid_for(n), 256, 0, (cudaStream_t)stream>>>(idx, idx_bits == 64, n, world, rank, pad_id, out);
    UR_RETURN_LAST_ERROR();
}

}  // extern "C"
