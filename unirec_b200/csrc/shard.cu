// Row-sharded table kernels (multi-GPU): rank r of W owns rows {id : id % W == r} at local index id / W.
// "Move queries, not rows": every rank sees ALL samples' ids and user vectors (all-gathered) and touches only the rows
// it owns, so per-rank HBM traffic stays that of one GPU's batch while only [samples, O(d)] states cross NVLink.
//
//   ur_shard_gather_rows   owned rows -> dense [n, d] buffer (zeros elsewhere)          -> reduce_scatter = sequence rows
//   ur_shard_localize      global ids -> local row ids (-1 where not owned / padding)   -> keys of the row-sparse update
//   ur_score_partial       one pass over OWNED target rows: per-sample online-softmax partial (m, l, sum y s, sum y,
//                          sum p m e, sum y m e) + raw scores of owned entries
//   ur_score_rescale       re-bases every partial onto the global max (after all_reduce MAX) -> partials add up
//   ur_score_finish        home rank: summed partials -> lse, per-sample loss, dLoss/du
//   ur_score_dscore        owner: dLoss/d(dot) for owned entries from the gathered lse
// Arithmetic = scoreloss.cu (reference: recommender.py:76-96, reco_abc.py:260-265).
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

struct PartialParams {
    const float4* table;      // local shard [ceil(V/W), d]
    const float4* user_emb;   // [S, d]  all samples of all ranks
    const int64_t* item_id;   // [S, N]  global ids
    const int32_t* label;     // [S, N] or null
    const float* item_bias;   // [V] replicated or null (global id)
    const float* user_bias;   // replicated or null
    const int64_t* user_id;   // [S]
    float inv_tau, clip;
    int N, W, r;
    int64_t S;
    float* z;                 // [S, N] scaled unclamped scores, written for owned entries only
    float* state;             // [S, 4 + 2d]
};

// One CTA (8 warps) per sample.  Warps take 32-id chunks; owned ids are ballot-compacted and processed 4 rows at a time
// with all 32 lanes on one row (d >= 128) or 32/LPR rows side by side (d < 128).
template <int D4>
__global__ void __launch_bounds__(256) score_partial_kernel(const PartialParams p) {
    constexpr int LPR = D4 < 32 ? D4 : 32;
    constexpr int VPL = D4 / LPR;
    constexpr int RPW = 32 / LPR;
    constexpr int D = D4 * 4;
    constexpr int U = 4;
    constexpr int G = 8 * RPW;
    __shared__ float gstate[G][4];
    __shared__ __align__(16) float gacc[G][D];
    __shared__ float accy[D];
    const int64_t b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane / LPR, col = lane % LPR;
    const int g = warp * RPW + sub;
    const int N = p.N;
    for (int i = threadIdx.x; i < D; i += blockDim.x) accy[i] = 0.f;
    __syncthreads();

    float4 uvec[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) uvec[v] = __ldg(p.user_emb + b * D4 + v * LPR + col);
    const float ub = p.user_bias ? __ldg(p.user_bias + __ldg(p.user_id + b)) : 0.f;
    const int64_t* ids = p.item_id + b * N;
    const int32_t* lab = p.label ? p.label + b * N : nullptr;
    const bool has_clip = p.clip > 0.f;
    float4 acc[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    float st_m = -INFINITY, st_l = 0.f, st_a = 0.f, st_b = 0.f;

    for (int c0 = warp * 32; c0 < N; c0 += 8 * 32) {
        const int jl = c0 + lane;
        const int64_t myid = jl < N ? __ldg(ids + jl) : -1;
        const bool own = jl < N && (myid % p.W) == p.r;
        const unsigned mask = __ballot_sync(0xffffffffu, own);
        const int cnt = __popc(mask);
        for (int k0 = 0; k0 < cnt; k0 += RPW * U) {
            float4 row[U][VPL];
            float bias[U];
            int jj[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = k0 + u * RPW + sub;
                const int src = k < cnt ? (int)__fns(mask, 0, k + 1) : 0;
                const int64_t id = __shfl_sync(0xffffffffu, myid, src);
                jj[u] = k < cnt ? c0 + src : -1;
                bias[u] = 0.f;
                if (jj[u] >= 0) {
                    const int64_t lrow = id / p.W;
#pragma unroll
                    for (int v = 0; v < VPL; ++v) row[u][v] = ldg_stream(p.table + lrow * D4 + v * LPR + col);
                    if (p.item_bias) bias[u] = __ldg(p.item_bias + id);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                // every lane takes part in the shuffles; only live groups update their state
                float dot = 0.f;
                if (jj[u] >= 0) {
#pragma unroll
                    for (int v = 0; v < VPL; ++v) dot += f4_dot(row[u][v], uvec[v]);
                }
                dot = group_sum<LPR>(dot);
                if (jj[u] >= 0) {
                    const int j = jj[u];
                    const float z = (dot + ub + bias[u]) * p.inv_tau;
                    const float s = has_clip ? fminf(fmaxf(z, -p.clip), p.clip) : z;
                    const float mk = (has_clip && (z < -p.clip || z > p.clip)) ? 0.f : 1.f;
                    if (col == 0) p.z[b * N + j] = z;
                    if (s > st_m) {
                        const float sc = __expf(st_m - s);
                        st_l *= sc;
#pragma unroll
                        for (int v = 0; v < VPL; ++v) acc[v] = f4_scale(acc[v], sc);
                        st_m = s;
                    }
                    const float pj = __expf(s - st_m);
                    st_l += pj;
                    const float w = pj * mk;
#pragma unroll
                    for (int v = 0; v < VPL; ++v) acc[v] = f4_fma(w, row[u][v], acc[v]);
                    const int yj = lab ? __ldg(lab + j) : (j == 0);
                    if (yj > 0) {
                        st_a += s; st_b += 1.f;
                        if (mk != 0.f) {
#pragma unroll
                            for (int v = 0; v < VPL; ++v) {
                                float* a = accy + (v * LPR + col) * 4;
                                atomicAdd(a + 0, row[u][v].x); atomicAdd(a + 1, row[u][v].y);
                                atomicAdd(a + 2, row[u][v].z); atomicAdd(a + 3, row[u][v].w);
                            }
                        }
                    }
                }
            }
        }
    }
    if (col == 0) { gstate[g][0] = st_m; gstate[g][1] = st_l; gstate[g][2] = st_a; gstate[g][3] = st_b; }
#pragma unroll
    for (int v = 0; v < VPL; ++v) reinterpret_cast<float4*>(gacc[g])[v * LPR + col] = acc[v];
    __syncthreads();
    float m_all = -INFINITY, l_all = 0.f, a_all = 0.f, b_all = 0.f;
    for (int q = 0; q < G; ++q) m_all = fmaxf(m_all, gstate[q][0]);
    for (int q = 0; q < G; ++q) {
        const float sc = gstate[q][1] > 0.f ? __expf(gstate[q][0] - m_all) : 0.f;
        l_all += gstate[q][1] * sc; a_all += gstate[q][2]; b_all += gstate[q][3];
    }
    float* st = p.state + b * (4 + 2 * D);
    if (threadIdx.x == 0) { st[0] = m_all; st[1] = l_all; st[2] = a_all; st[3] = b_all; }
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        float a = 0.f;
        for (int q = 0; q < G; ++q) {
            const float sc = gstate[q][1] > 0.f ? __expf(gstate[q][0] - m_all) : 0.f;
            a += gacc[q][c] * sc;
        }
        st[4 + c] = a;
        st[4 + D + c] = accy[c];
    }
}

// state[s] (m, l, ., ., acc[d], .) -> re-based on gmax[s]; the m slot is zeroed so that partials add up
__global__ void __launch_bounds__(256) score_rescale_kernel(float* __restrict__ state, const float* __restrict__ gmax, int64_t S, int d) {
    const int stride = 4 + 2 * d;
    const int64_t total = S * (d + 1);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = e / (d + 1);
        const int c = (int)(e - s * (d + 1));
        float* st = state + s * stride;
        const float m = st[0], l = st[1];
        const float sc = l > 0.f ? __expf(m - gmax[s]) : 0.f;   // reads of st[0]/st[1] race only with the c==0 writer below:
        if (c == 0) continue;                                    // handled by the second kernel pass (see host wrapper)
        st[3 + c] *= sc;
    }
}
__global__ void __launch_bounds__(256) score_rescale_head_kernel(float* __restrict__ state, const float* __restrict__ gmax, int64_t S, int d) {
    const int stride = 4 + 2 * d;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
        float* st = state + s * stride;
        const float sc = st[1] > 0.f ? __expf(st[0] - gmax[s]) : 0.f;
        st[1] *= sc;
        st[0] = 0.f;
    }
}

// home rank: summed, re-based partials of its own B samples -> lse, loss term, dLoss/du
__global__ void __launch_bounds__(128) score_finish_kernel(const float* __restrict__ state, const float* __restrict__ gmax, int64_t B, int d,
                                                           float inv_tau, const float* __restrict__ norm_dev, float* __restrict__ loss_vec,
                                                           float* __restrict__ lse_ny, float* __restrict__ grad_user) {
    const int64_t b = blockIdx.x;
    const float* st = state + b * (4 + 2 * d);
    const float l = st[1], ys = st[2], ny = st[3];
    const float lse = gmax[b] + __logf(l);
    const float gscale = inv_tau / *norm_dev;
    if (threadIdx.x == 0) {
        loss_vec[b] = ny * lse - ys;
        lse_ny[2 * b] = lse;
        lse_ny[2 * b + 1] = ny;
    }
    for (int c = threadIdx.x; c < d; c += blockDim.x)
        grad_user[b * d + c] = (ny * st[4 + c] / l - st[4 + d + c]) * gscale;
}

__global__ void __launch_bounds__(256) score_dscore_kernel(const float* __restrict__ z, const int64_t* __restrict__ item_id,
                                                           const int32_t* __restrict__ label, const float* __restrict__ lse_ny, int64_t S,
                                                           int N, int W, int r, float inv_tau, float clip, const float* __restrict__ norm_dev,
                                                           float* __restrict__ dscore) {
    const int64_t total = S * N;
    const float gscale = inv_tau / *norm_dev;
    const bool has_clip = clip > 0.f;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        float g = 0.f;
        if (item_id[e] % W == r) {
            const int64_t s = e / N;
            const int j = (int)(e - s * N);
            const float zz = z[e];
            const float sc = has_clip ? fminf(fmaxf(zz, -clip), clip) : zz;
            const float mk = (has_clip && (zz < -clip || zz > clip)) ? 0.f : 1.f;
            const float y = label ? (float)(label[e] > 0) : (float)(j == 0);
            g = (lse_ny[2 * s + 1] * __expf(sc - lse_ny[2 * s]) - y) * mk * gscale;
        }
        dscore[e] = g;
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

int ur_score_partial_f32(const float* table_local, int d, const float* user_emb, const int64_t* item_id, int64_t S, int N,
                         const int32_t* label, const float* item_bias, const float* user_bias, const int64_t* user_id, float tau,
                         float score_clip, int world, int rank, float* z, float* state, void* stream) {
    if (d <= 0 || (d & 3) || N <= 0 || tau == 0.f || world < 1 || rank < 0 || rank >= world) return UR_ERR_BAD_ARG;
    if (S == 0) return UR_OK;
    ur::PartialParams p;
    p.table = (const float4*)table_local; p.user_emb = (const float4*)user_emb; p.item_id = item_id; p.label = label;
    p.item_bias = item_bias; p.user_bias = user_bias; p.user_id = user_id; p.inv_tau = 1.f / tau; p.clip = score_clip;
    p.N = N; p.W = world; p.r = rank; p.S = S; p.z = z; p.state = state;
    cudaStream_t st = (cudaStream_t)stream;
    switch (d) {
        case 32: ur::score_partial_kernel<8><<<(unsigned)S, 256, 0, st>>>(p); break;
        case 64: ur::score_partial_kernel<16><<<(unsigned)S, 256, 0, st>>>(p); break;
        case 128: ur::score_partial_kernel<32><<<(unsigned)S, 256, 0, st>>>(p); break;
        case 256: ur::score_partial_kernel<64><<<(unsigned)S, 256, 0, st>>>(p); break;
        default: return UR_ERR_UNSUPPORTED;
    }
    UR_RETURN_LAST_ERROR();
}

int ur_score_rescale_f32(float* state, const float* gmax, int64_t S, int d, void* stream) {
    if (S == 0) return UR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    ur::score_rescale_kernel<<<ur::grid_for(S * (d + 1)), 256, 0, st>>>(state, gmax, S, d);
    ur::score_rescale_head_kernel<<<ur::grid_for(S), 256, 0, st>>>(state, gmax, S, d);
    UR_RETURN_LAST_ERROR();
}

int ur_score_finish_f32(const float* state, const float* gmax, int64_t B, int d, float tau, const float* norm_dev, float* loss_vec,
                        float* lse_ny, float* grad_user, void* stream) {
    if (B == 0) return UR_OK;
    ur::score_finish_kernel<<<(unsigned)B, 128, 0, (cudaStream_t)stream>>>(state, gmax, B, d, 1.f / tau, norm_dev, loss_vec, lse_ny, grad_user);
    UR_RETURN_LAST_ERROR();
}

int ur_score_dscore_f32(const float* z, const int64_t* item_id, const int32_t* label, const float* lse_ny, int64_t S, int N, int world,
                        int rank, float tau, float score_clip, const float* norm_dev, float* dscore, void* stream) {
    if (S == 0) return UR_OK;
    ur::score_dscore_kernel<<<ur::grid_for(S * N), 256, 0, (cudaStream_t)stream>>>(z, item_id, label, lse_ny, S, N, world, rank, 1.f / tau,
                                                                                    score_clip, norm_dev, dscore);
    UR_RETURN_LAST_ERROR();
}

}  // extern "C"
