// K8: fused target/negative gather -> inner-product scores -> bias / tau / clamp -> sampled-softmax or BPR loss
//     -> dLoss/dscore and dLoss/du, in ONE pass over the table rows.  The [B,1+K,d] tensor of the reference
//     (unirec/model/base/recommender.py:55,66-67) is never materialised: every row is read from HBM exactly once.
//
// Reference arithmetic restated here:
//   scores  : unirec/model/modules.py:59-66 (InnerProductScorer), recommender.py:76-96 (bias, /tau, clamp)
//   softmax : unirec/model/base/reco_abc.py:260-265   loss = mean over {label>0} of (logsumexp_j s_bj - s_bj)
//   bpr     : reco_abc.py:252-255, modules.py:15-21   loss = mean_{b,j>=1} -log(1e-8 + sigmoid(s_b0 - s_bj))
//
// The softmax branch is a single-query attention: an online (max, sum, weighted-row-sum) recurrence gives
// dL/du = (n_b * sum_j p_j m_j e_j - sum_j y_j m_j e_j) / (P tau) without a second pass over the rows
// (m_j = clamp pass-through mask, n_b = positives in row b, P = positives in the batch).
#include <stdlib.h>
#include "common.cuh"

namespace ur {

struct ScoreLossParams {
    const float4* table;      // [V, d]
    const float4* user_emb;   // [B, d]
    const int64_t* item_id;   // [B, N]
    const int32_t* label;     // [B, N] or null (column 0 positive)
    const float* item_bias;   // [V] or null
    const float* user_bias;   // [n_users] or null
    const int64_t* user_id;   // [B] (only with user_bias)
    const float* norm_dev;    // device scalar normaliser (softmax: P) or null -> norm_host
    float norm_host;
    float inv_tau, clip;      // clip <= 0: no clamp
    int N;
    int64_t B;
    int wps;                  // warps per sample (1,2,4,8)
    float* scores;            // [B, N] final scores (may be null)
    float* loss_vec;          // [B]
    float* dscore;            // [B, N] dLoss/d(dot)   (may be null -> forward only)
    float4* grad_user;        // [B, d]                (may be null -> forward only)
    int l2_evict_first;       // v3 kernel: stream the rows through L2 with evict-first priority (must match scoreloss_v3.cu)
};

constexpr float kBprEps = 1e-8f;   // unirec/constants/global_variables.py:4

template <int D4, int LOSS>   // LOSS 0 = softmax, 1 = bpr
__global__ void __launch_bounds__(256) score_loss_kernel(const ScoreLossParams p) {
    constexpr int LPR = D4 < 32 ? D4 : 32;   // lanes per row
    constexpr int VPL = D4 / LPR;            // float4 per lane
    constexpr int RPW = 32 / LPR;            // row groups per warp
    constexpr int D = D4 * 4;
    constexpr int U = 4;                     // rows in flight per group
    extern __shared__ float smem[];

    const int wps = p.wps, spb = 8 / wps, G = wps * RPW, N = p.N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ls = warp / wps;                       // local sample
    const int wis = warp - ls * wps;                 // warp in sample
    const int sub = lane / LPR, col = lane % LPR;
    const int g = wis * RPW + sub;                   // group in sample
    const int64_t b = (int64_t)blockIdx.x * spb + ls;
    const bool live = b < p.B;

    // shared layout
    float* zbuf = smem;                                  // [spb][N]      scaled, unclamped scores
    float* gstate = zbuf + (((size_t)spb * N + 3) & ~(size_t)3);   // [spb][G][4]  (16-byte aligned: gacc/aux take float4)
    float* gacc = gstate + spb * G * 4;                  // [spb][G][D]
    float* aux = gacc + (size_t)spb * G * D;             // [spb][D]      softmax: sum_j y_j m_j e_j ; bpr: e_0

    for (int i = threadIdx.x; i < spb * D; i += blockDim.x) aux[i] = 0.f;
    __syncthreads();

    float4 uvec[VPL];
    float ub = 0.f;
    const int64_t* ids = p.item_id + (live ? b : 0) * N;
    const int32_t* lab = p.label ? p.label + (live ? b : 0) * N : nullptr;
    if (live) {
#pragma unroll
        for (int v = 0; v < VPL; ++v) uvec[v] = __ldg(p.user_emb + b * D4 + v * LPR + col);
        if (p.user_bias) ub = __ldg(p.user_bias + __ldg(p.user_id + b));
    }
    const float clip = p.clip;
    const bool has_clip = clip > 0.f;

    float4 acc[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    float st_m = -INFINITY, st_l = 0.f;   // softmax: running max / sum.    bpr: st_l = sum_j loss_j
    float st_a = 0.f, st_b = 0.f;         // softmax: sum y_j s_j, sum y_j. bpr: st_a = sum_j c_j
    float s0 = 0.f, mask0 = 1.f;
    float4 e0[VPL];

    if (LOSS == 1 && live) {               // BPR needs s_0 before any negative: every group reads row 0 (L2 hit)
        const int64_t id0 = __ldg(ids);
        float dot = 0.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            e0[v] = ldg_stream(p.table + id0 * D4 + v * LPR + col);
            dot += f4_dot(e0[v], uvec[v]);
        }
        dot = group_sum<LPR>(dot);
        float z = (dot + ub + (p.item_bias ? __ldg(p.item_bias + id0) : 0.f)) * p.inv_tau;
        s0 = has_clip ? fminf(fmaxf(z, -clip), clip) : z;
        mask0 = (has_clip && (z < -clip || z > clip)) ? 0.f : 1.f;
        if (g == 0) {
            if (col == 0) zbuf[(size_t)ls * N] = z;
#pragma unroll
            for (int v = 0; v < VPL; ++v)
                reinterpret_cast<float4*>(aux + ls * D)[v * LPR + col] = e0[v];
        }
    }

    if (live) {
        const int jstart = LOSS == 1 ? 1 : 0;
        // trip count is uniform across the warp (row groups differ only in the offset g): shuffles stay convergent
        for (int jb = jstart; jb < N; jb += G * U) {
            const int j0 = jb + g;
            float4 row[U][VPL];
            float bias[U];
            int32_t y[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = j0 + u * G;
                bias[u] = 0.f; y[u] = 0;
                if (j < N) {
                    const int64_t id = __ldg(ids + j);
#pragma unroll
                    for (int v = 0; v < VPL; ++v) row[u][v] = ldg_stream(p.table + id * D4 + v * LPR + col);
                    if (p.item_bias) bias[u] = __ldg(p.item_bias + id);
                    if (LOSS == 0) y[u] = lab ? __ldg(lab + j) : (j == 0);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = j0 + u * G;
                // the shuffle reduction is executed by ALL lanes (row groups of one warp may disagree on j < N)
                float dot = 0.f;
                if (j < N) {
#pragma unroll
                    for (int v = 0; v < VPL; ++v) dot += f4_dot(row[u][v], uvec[v]);
                }
                dot = group_sum<LPR>(dot);
                if (j < N) {
                    const float z = (dot + ub + bias[u]) * p.inv_tau;
                    const float s = has_clip ? fminf(fmaxf(z, -clip), clip) : z;
                    const float mask = (has_clip && (z < -clip || z > clip)) ? 0.f : 1.f;
                    if (col == 0) zbuf[(size_t)ls * N + j] = z;
                    if (LOSS == 0) {
                        if (s > st_m) {
                            const float sc = __expf(st_m - s);
                            st_l *= sc;
#pragma unroll
                            for (int v = 0; v < VPL; ++v) acc[v] = f4_scale(acc[v], sc);
                            st_m = s;
                        }
                        const float pj = __expf(s - st_m);
                        st_l += pj;
                        const float w = pj * mask;
#pragma unroll
                        for (int v = 0; v < VPL; ++v) acc[v] = f4_fma(w, row[u][v], acc[v]);
                        if (y[u] > 0) {
                            st_a += s; st_b += 1.f;
                            if (mask != 0.f) {
#pragma unroll
                                for (int v = 0; v < VPL; ++v) {
                                    float* a = aux + ls * D + (v * LPR + col) * 4;
                                    atomicAdd(a + 0, row[u][v].x); atomicAdd(a + 1, row[u][v].y);
                                    atomicAdd(a + 2, row[u][v].z); atomicAdd(a + 3, row[u][v].w);
                                }
                            }
                        }
                    } else {
                        const float x = s0 - s;
                        const float sig = 1.f / (1.f + __expf(-x));
                        st_l += -__logf(kBprEps + sig);
                        const float c = sig * (1.f - sig) / (kBprEps + sig);
                        st_a += c;
                        const float w = c * mask;
#pragma unroll
                        for (int v = 0; v < VPL; ++v) acc[v] = f4_fma(w, row[u][v], acc[v]);
                    }
                }
            }
        }
        // publish group state
        if (col == 0) {
            float* gs = gstate + (ls * G + g) * 4;
            gs[0] = st_m; gs[1] = st_l; gs[2] = st_a; gs[3] = st_b;
        }
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            reinterpret_cast<float4*>(gacc + (size_t)(ls * G + g) * D)[v * LPR + col] = acc[v];
    }
    __syncthreads();

    // ---- combine the G groups of each sample; threads of the sample's warps cooperate ----
    const int tis = wis * 32 + lane;          // thread in sample
    const int tps = wps * 32;                 // threads per sample
    float norm = p.norm_dev ? __ldg(p.norm_dev) : p.norm_host;
    if (live) {
        float m_all = -INFINITY;
        if (LOSS == 0)
            for (int q = 0; q < G; ++q) m_all = fmaxf(m_all, gstate[(ls * G + q) * 4]);
        float l_all = 0.f, a_all = 0.f, b_all = 0.f;
        for (int q = 0; q < G; ++q) {
            const float* gs = gstate + (ls * G + q) * 4;
            if (LOSS == 0) {
                const float sc = gs[1] > 0.f ? __expf(gs[0] - m_all) : 0.f;
                l_all += gs[1] * sc;
            } else {
                l_all += gs[1];
            }
            a_all += gs[2]; b_all += gs[3];
        }
        if (LOSS == 0) {
            const float lse = m_all + __logf(l_all);
            const float gscale = p.inv_tau / norm;
            if (p.grad_user) {
                for (int c = tis; c < D; c += tps) {
                    float a = 0.f;
                    for (int q = 0; q < G; ++q) {
                        const float* gs = gstate + (ls * G + q) * 4;
                        const float sc = gs[1] > 0.f ? __expf(gs[0] - m_all) : 0.f;
                        a += gacc[(size_t)(ls * G + q) * D + c] * sc;
                    }
                    reinterpret_cast<float*>(p.grad_user)[b * D + c] = (b_all * a / l_all - aux[ls * D + c]) * gscale;
                }
            }
            if (tis == 0) p.loss_vec[b] = b_all * lse - a_all;
            for (int j = tis; j < N; j += tps) {
                const float z = zbuf[(size_t)ls * N + j];
                const float s = has_clip ? fminf(fmaxf(z, -clip), clip) : z;
                if (p.scores) p.scores[b * N + j] = s;
                if (p.dscore) {
                    const float mask = (has_clip && (z < -clip || z > clip)) ? 0.f : 1.f;
                    const float yj = lab ? (float)(__ldg(lab + j) > 0) : (float)(j == 0);
                    p.dscore[b * N + j] = (b_all * __expf(s - lse) - yj) * mask * gscale;
                }
            }
        } else {
            const float K = (float)(N - 1);
            const float gscale = p.inv_tau / norm;        // norm = B*K
            if (p.grad_user) {
                for (int c = tis; c < D; c += tps) {
                    float a = 0.f;
                    for (int q = 0; q < G; ++q) a += gacc[(size_t)(ls * G + q) * D + c];
                    reinterpret_cast<float*>(p.grad_user)[b * D + c] = (a - a_all * mask0 * aux[ls * D + c]) * gscale;
                }
            }
            if (tis == 0) p.loss_vec[b] = l_all / K;
            for (int j = tis; j < N; j += tps) {
                const float z = zbuf[(size_t)ls * N + j];
                const float s = has_clip ? fminf(fmaxf(z, -clip), clip) : z;
                if (p.scores) p.scores[b * N + j] = s;
                if (p.dscore) {
                    float gj;
                    if (j == 0) {
                        gj = -a_all * mask0 * gscale;
                    } else {
                        const float mask = (has_clip && (z < -clip || z > clip)) ? 0.f : 1.f;
                        const float x = s0 - s;
                        const float sig = 1.f / (1.f + __expf(-x));
                        gj = sig * (1.f - sig) / (kBprEps + sig) * mask * gscale;
                    }
                    p.dscore[b * N + j] = gj;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Large-N variant (one sample per CTA): table rows are staged by the async copy engine.  Each warp owns a ring of
// kStages x kRows row slots in shared memory; one lane per row issues `cp.async.bulk` (UBLKCP) straight from the table
// into its slot, completion is tracked by one mbarrier per stage.  Bytes in flight no longer cost registers:
// 8 warps x (kStages-1) x kRows rows of d*4 bytes per CTA are outstanding while the warp reduces the current stage.
// Arithmetic, outputs and the final combine are identical to score_loss_kernel.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sl_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sl_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "SL_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra SL_DONE;\n\t"
        "bra SL_WAIT;\n\t"
        "SL_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

constexpr int kStages = 3, kRows = 4;
static int g_use_bulk = 1;      // A/B switch (ur_score_loss_set_bulk): 0 = register-staged, 1 = v3 (scoreloss_v3.cu), 2 = v2 bulk ring
int score_loss_v3_try(const ScoreLossParams& p, int d, int loss_type, cudaStream_t st);
extern int g_v3_variant;

template <int D4, int LOSS>
__global__ void __launch_bounds__(256) score_loss_bulk_kernel(const ScoreLossParams p) {
    constexpr int LPR = D4 < 32 ? D4 : 32;
    constexpr int VPL = D4 / LPR;
    constexpr int RPW = 32 / LPR;
    constexpr int D = D4 * 4;
    constexpr int ROWB = D * 4;                              // bytes per table row
    constexpr int G = 8 * RPW;
    extern __shared__ __align__(128) float smem[];
    const int N = p.N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane / LPR, col = lane % LPR;
    const int g = warp * RPW + sub;
    const int64_t b = blockIdx.x;

    // shared layout: ring | ids | barriers | zbuf | aux ; the ring is reused for the group combine after the main loop
    float* ring = smem;                                                     // [8][kStages][kRows][D]
    int64_t* idbuf = reinterpret_cast<int64_t*>(ring + 8 * kStages * kRows * D);   // [8][kStages][kRows]
    uint64_t* bars = reinterpret_cast<uint64_t*>(idbuf + 8 * kStages * kRows);     // [8][kStages]
    float* zbuf = reinterpret_cast<float*>(bars + 8 * kStages);             // [N] (padded to 4)
    float* aux = zbuf + ((N + 3) & ~3);                                     // [D]
    float* gstate = aux + D;                                                // [G][4]

    float* my_ring = ring + (size_t)warp * kStages * kRows * D;
    int64_t* my_ids = idbuf + warp * kStages * kRows;
    uint64_t* my_bars = bars + warp * kStages;
    if (lane < kStages) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sl_smem_u32(my_bars + lane)) : "memory");
    for (int i = threadIdx.x; i < D; i += blockDim.x) aux[i] = 0.f;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    float4 uvec[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) uvec[v] = __ldg(p.user_emb + b * D4 + v * LPR + col);
    const float ub = p.user_bias ? __ldg(p.user_bias + __ldg(p.user_id + b)) : 0.f;
    const int64_t* ids = p.item_id + b * N;
    const int32_t* lab = p.label ? p.label + b * N : nullptr;
    const float clip = p.clip;
    const bool has_clip = clip > 0.f;

    float4 acc[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    float st_m = -INFINITY, st_l = 0.f, st_a = 0.f, st_b = 0.f;
    float s0 = 0.f, mask0 = 1.f;

    if (LOSS == 1) {      // BPR: s_0 first (every group reads row 0 directly; it is one row)
        const int64_t id0 = __ldg(ids);
        float dot = 0.f;
        float4 e0[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            e0[v] = ldg_stream(p.table + id0 * D4 + v * LPR + col);
            dot += f4_dot(e0[v], uvec[v]);
        }
        dot = group_sum<LPR>(dot);
        const float z = (dot + ub + (p.item_bias ? __ldg(p.item_bias + id0) : 0.f)) * p.inv_tau;
        s0 = has_clip ? fminf(fmaxf(z, -clip), clip) : z;
        mask0 = (has_clip && (z < -clip || z > clip)) ? 0.f : 1.f;
        if (g == 0) {
            if (col == 0) zbuf[0] = z;
#pragma unroll
            for (int v = 0; v < VPL; ++v) reinterpret_cast<float4*>(aux)[v * LPR + col] = e0[v];
        }
    }

    const int jstart = LOSS == 1 ? 1 : 0;
    const int total_chunks = (N - jstart + kRows - 1) / kRows;
    const int my_chunks = total_chunks > warp ? (total_chunks - warp + 7) / 8 : 0;      // chunks warp, warp+8, ...

    auto issue = [&](int i) {
        const int slot = i % kStages;
        const int jbase = jstart + (i * 8 + warp) * kRows;
        const int nvalid = min(kRows, N - jbase);
        const uint32_t bar = sl_smem_u32(my_bars + slot);
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(nvalid * ROWB)) : "memory");
        if (lane < nvalid) {
            const int64_t id = __ldg(ids + jbase + lane);
            my_ids[slot * kRows + lane] = id;
            const uint32_t dst = sl_smem_u32(my_ring + (size_t)(slot * kRows + lane) * D);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(reinterpret_cast<const float*>(p.table) + id * D), "r"((uint32_t)ROWB), "r"(bar) : "memory");
        }
    };

    for (int i = 0; i < kStages - 1 && i < my_chunks; ++i) issue(i);
    __syncwarp();
    for (int i = 0; i < my_chunks; ++i) {
        if (i + kStages - 1 < my_chunks) issue(i + kStages - 1);
        const int slot = i % kStages;
        sl_mbar_wait(sl_smem_u32(my_bars + slot), (i / kStages) & 1);
        const int jbase = jstart + (i * 8 + warp) * kRows;
#pragma unroll
        for (int r0 = 0; r0 < kRows; r0 += RPW) {
            const int r = r0 + sub;
            const int j = jbase + r;
            const bool ok = r < kRows && j < N;
            float4 row[VPL];
            float dot = 0.f;
            if (ok) {
                const float4* src = reinterpret_cast<const float4*>(my_ring + (size_t)(slot * kRows + r) * D);
#pragma unroll
                for (int v = 0; v < VPL; ++v) { row[v] = src[v * LPR + col]; dot += f4_dot(row[v], uvec[v]); }
            }
            dot = group_sum<LPR>(dot);
            if (ok) {
                const float bias = p.item_bias ? __ldg(p.item_bias + my_ids[slot * kRows + r]) : 0.f;
                const float z = (dot + ub + bias) * p.inv_tau;
                const float s = has_clip ? fminf(fmaxf(z, -clip), clip) : z;
                const float mask = (has_clip && (z < -clip || z > clip)) ? 0.f : 1.f;
                if (col == 0) zbuf[j] = z;
                if (LOSS == 0) {
                    if (s > st_m) {
                        const float sc = __expf(st_m - s);
                        st_l *= sc;
#pragma unroll
                        for (int v = 0; v < VPL; ++v) acc[v] = f4_scale(acc[v], sc);
                        st_m = s;
                    }
                    const float pj = __expf(s - st_m);
                    st_l += pj;
                    const float w = pj * mask;
#pragma unroll
                    for (int v = 0; v < VPL; ++v) acc[v] = f4_fma(w, row[v], acc[v]);
                    const int yj = lab ? __ldg(lab + j) : (j == 0);
                    if (yj > 0) {
                        st_a += s; st_b += 1.f;
                        if (mask != 0.f) {
#pragma unroll
                            for (int v = 0; v < VPL; ++v) {
                                float* a = aux + (v * LPR + col) * 4;
                                atomicAdd(a + 0, row[v].x); atomicAdd(a + 1, row[v].y);
                                atomicAdd(a + 2, row[v].z); atomicAdd(a + 3, row[v].w);
                            }
                        }
                    }
                } else {
                    const float x = s0 - s;
                    const float sig = 1.f / (1.f + __expf(-x));
                    st_l += -__logf(kBprEps + sig);
                    const float c = sig * (1.f - sig) / (kBprEps + sig);
                    st_a += c;
                    const float w = c * mask;
#pragma unroll
                    for (int v = 0; v < VPL; ++v) acc[v] = f4_fma(w, row[v], acc[v]);
                }
            }
        }
        __syncwarp();      // every lane is done with this slot before it is re-armed by the next issue()
    }
    __syncthreads();       // all rings idle: reuse ring memory for the per-group partial sums
    float* gacc = ring;    // [G][D]
    if (col == 0) { gstate[g * 4 + 0] = st_m; gstate[g * 4 + 1] = st_l; gstate[g * 4 + 2] = st_a; gstate[g * 4 + 3] = st_b; }
#pragma unroll
    for (int v = 0; v < VPL; ++v) reinterpret_cast<float4*>(gacc + (size_t)g * D)[v * LPR + col] = acc[v];
    __syncthreads();

    const int tis = threadIdx.x, tps = blockDim.x;
    const float norm = p.norm_dev ? __ldg(p.norm_dev) : p.norm_host;
    float m_all = -INFINITY;
    if (LOSS == 0)
        for (int q = 0; q < G; ++q) m_all = fmaxf(m_all, gstate[q * 4]);
    float l_all = 0.f, a_all = 0.f, b_all = 0.f;
    for (int q = 0; q < G; ++q) {
        const float* gs = gstate + q * 4;
        if (LOSS == 0) l_all += gs[1] * (gs[1] > 0.f ? __expf(gs[0] - m_all) : 0.f);
        else l_all += gs[1];
        a_all += gs[2]; b_all += gs[3];
    }
    const float gscale = p.inv_tau / norm;
    if (LOSS == 0) {
        const float lse = m_all + __logf(l_all);
        if (p.grad_user) {
            for (int c = tis; c < D; c += tps) {
                float a = 0.f;
                for (int q = 0; q < G; ++q) {
                    const float* gs = gstate + q * 4;
                    a += gacc[(size_t)q * D + c] * (gs[1] > 0.f ? __expf(gs[0] - m_all) : 0.f);
                }
                reinterpret_cast<float*>(p.grad_user)[b * D + c] = (b_all * a / l_all - aux[c]) * gscale;
            }
        }
        if (tis == 0) p.loss_vec[b] = b_all * lse - a_all;
        for (int j = tis; j < N; j += tps) {
            const float z = zbuf[j];
            const float s = has_clip ? fminf(fmaxf(z, -clip), clip) : z;
            if (p.scores) p.scores[b * N + j] = s;
            if (p.dscore) {
                const float mask = (has_clip && (z < -clip || z > clip)) ? 0.f : 1.f;
                const float yj = lab ? (float)(__ldg(lab + j) > 0) : (float)(j == 0);
                p.dscore[b * N + j] = (b_all * __expf(s - lse) - yj) * mask * gscale;
            }
        }
    } else {
        const float K = (float)(N - 1);
        if (p.grad_user) {
            for (int c = tis; c < D; c += tps) {
                float a = 0.f;
                for (int q = 0; q < G; ++q) a += gacc[(size_t)q * D + c];
                reinterpret_cast<float*>(p.grad_user)[b * D + c] = (a - a_all * mask0 * aux[c]) * gscale;
            }
        }
        if (tis == 0) p.loss_vec[b] = l_all / K;
        for (int j = tis; j < N; j += tps) {
            const float z = zbuf[j];
            const float s = has_clip ? fminf(fmaxf(z, -clip), clip) : z;
            if (p.scores) p.scores[b * N + j] = s;
            if (p.dscore) {
                float gj;
                if (j == 0) {
                    gj = -a_all * mask0 * gscale;
                } else {
                    const float mask = (has_clip && (z < -clip || z > clip)) ? 0.f : 1.f;
                    const float x = s0 - s;
                    const float sig = 1.f / (1.f + __expf(-x));
                    gj = sig * (1.f - sig) / (kBprEps + sig) * mask * gscale;
                }
                p.dscore[b * N + j] = gj;
            }
        }
    }
}

static size_t score_loss_bulk_smem(int d, int N) {
    const int d4 = d / 4, lpr = d4 < 32 ? d4 : 32, rpw = 32 / lpr;
    size_t ring = (size_t)8 * kStages * kRows * d * 4;
    const size_t gacc = (size_t)8 * rpw * d * 4;
    if (gacc > ring) ring = gacc;
    return ring + (size_t)8 * kStages * kRows * 8 + (size_t)8 * kStages * 8 + (size_t)((N + 3) & ~3) * 4 + (size_t)d * 4 +
           (size_t)8 * rpw * 16 + 128;
}

// loss = sum(loss_vec) / denom   (denom from device when given); NaN flag for the trainer's skip-step rule
// (unirec/facility/trainer.py:164-168, 344-352) without a host sync in the step.
__global__ void __launch_bounds__(1024) loss_finish_kernel(const float* __restrict__ loss_vec, int64_t B, const float* denom_dev,
                                                           float denom_host, float* __restrict__ loss_out, int32_t* nan_flag) {
    __shared__ double sh[32];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < B; i += blockDim.x) s += (double)loss_vec[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) {
            const float denom = denom_dev ? *denom_dev : denom_host;
            const float loss = (float)(s / (double)denom);
            *loss_out = loss;
            if (nan_flag) *nan_flag = isnan(loss) ? 1 : 0;
        }
    }
}

// out (pre-zeroed) += number of positive labels; per-CTA integer partials, one float atomic each (exact below 2^24)
__global__ void __launch_bounds__(256) count_positive_kernel(const int32_t* __restrict__ label, int64_t n, float* out) {
    __shared__ int sh[8];
    int c = 0;
    const int64_t n4 = n >> 2;
    const int4* l4 = reinterpret_cast<const int4*>(label);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int4 v = __ldg(l4 + i);
        c += (v.x > 0) + (v.y > 0) + (v.z > 0) + (v.w > 0);
    }
    if (blockIdx.x == 0)
        for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) c += label[i] > 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < 8; ++i) t += sh[i];
        if (t) atomicAdd(out, (float)t);
    }
}

static size_t score_loss_smem(int d, int N, int wps) {
    const int d4 = d / 4, lpr = d4 < 32 ? d4 : 32, rpw = 32 / lpr;
    const int spb = 8 / wps, G = wps * rpw;
    return sizeof(float) * ((((size_t)spb * N + 3) & ~(size_t)3) + (size_t)spb * G * 4 + (size_t)spb * G * d + (size_t)spb * d);
}

template <int D4>
static int launch_score_loss(const ScoreLossParams& p, int loss_type, size_t smem, cudaStream_t st) {
    const int spb = 8 / p.wps;
    const unsigned grid = (unsigned)((p.B + spb - 1) / spb);
    // v2 ring kernel: the large-N path of the narrow rows (d <= 64); wide rows (d >= 128) take scoreloss_v3.cu, or the register-staged
    // kernel below for the few shapes outside it -- the v2 kernel is not instantiated for them
    if constexpr (D4 <= 16)
    if (p.wps == 8 && p.N >= 96 && g_use_bulk) {       // one sample per CTA, enough rows to fill the async ring
        const size_t bsm = score_loss_bulk_smem(D4 * 4, p.N);
        if (bsm <= 200 * 1024) {
            if (loss_type == 0) {
                if (bsm > 48 * 1024) cudaFuncSetAttribute(score_loss_bulk_kernel<D4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm);
                score_loss_bulk_kernel<D4, 0><<<(unsigned)p.B, 256, bsm, st>>>(p);
            } else {
                if (bsm > 48 * 1024) cudaFuncSetAttribute(score_loss_bulk_kernel<D4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm);
                score_loss_bulk_kernel<D4, 1><<<(unsigned)p.B, 256, bsm, st>>>(p);
            }
            return 0;
        }
    }
    if (loss_type == 0) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(score_loss_kernel<D4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        score_loss_kernel<D4, 0><<<grid, 256, smem, st>>>(p);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(score_loss_kernel<D4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        score_loss_kernel<D4, 1><<<grid, 256, smem, st>>>(p);
    }
    return 0;
}

}  // namespace ur

extern "C" {

// bring-up / A-B switch between the register-staged and the cp.async.bulk-staged large-N kernels
int ur_score_loss_set_bulk(int on) {
    if (on >= 10) { ur::g_v3_variant = on - 10; ur::g_use_bulk = 1; } else ur::g_use_bulk = on;
    return UR_OK;
}

int ur_count_positive_i32(const int32_t* label, int64_t n, float* out, void* stream) {
    if (reinterpret_cast<uintptr_t>(label) & 15) return UR_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(out, 0, sizeof(float), st);
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > ur::kNumSMs * 4) blocks = ur::kNumSMs * 4;
    if (blocks < 1) blocks = 1;
    ur::count_positive_kernel<<<(unsigned)blocks, 256, 0, st>>>(label, n, out);
    UR_RETURN_LAST_ERROR();
}

int ur_loss_finish_f32(const float* loss_vec, int64_t B, const float* denom_dev, float denom_host, float* loss_out,
                       int32_t* nan_flag, void* stream) {
    ur::loss_finish_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(loss_vec, B, denom_dev, denom_host, loss_out, nan_flag);
    UR_RETURN_LAST_ERROR();
}

int ur_score_loss_fwd_bwd_f32(const float* table, int d, const float* user_emb, const int64_t* item_id, int64_t B, int N,
                              const int32_t* label, const float* item_bias, const float* user_bias, const int64_t* user_id,
                              float tau, float score_clip, int loss_type, const float* norm_dev, float norm_host, float* scores,
                              float* loss_vec, float* dscore, float* grad_user, void* stream) {
    if (d <= 0 || (d & 3) || N <= 0 || (loss_type != 0 && loss_type != 1) || tau == 0.f) return UR_ERR_BAD_ARG;
    if (loss_type == 1 && N < 2) return UR_ERR_BAD_ARG;
    if (B == 0) return UR_OK;
    ur::ScoreLossParams p;
    p.table = (const float4*)table; p.user_emb = (const float4*)user_emb; p.item_id = item_id; p.label = label;
    p.item_bias = item_bias; p.user_bias = user_bias; p.user_id = user_id; p.norm_dev = norm_dev; p.norm_host = norm_host;
    p.inv_tau = 1.f / tau; p.clip = score_clip; p.N = N; p.B = B;
    p.scores = scores; p.loss_vec = loss_vec; p.dscore = dscore; p.grad_user = (float4*)grad_user;
    static const int env_evict = getenv("UR_SCORE_EVICT_FIRST") ? atoi(getenv("UR_SCORE_EVICT_FIRST")) : 1;
    p.l2_evict_first = env_evict;
    const int rpw = d >= 128 ? 1 : 128 / d;
    int wps = 1;
    while (wps < 8 && N > wps * rpw * 8) wps *= 2;     // >= 8 rows per group before adding warps
    size_t smem = ur::score_loss_smem(d, N, wps);
    while (smem > 200 * 1024 && wps < 8) { wps *= 2; smem = ur::score_loss_smem(d, N, wps); }
    if (smem > 200 * 1024) return UR_ERR_UNSUPPORTED;
    p.wps = wps;
    cudaStream_t st = (cudaStream_t)stream;
    if (ur::g_use_bulk == 1) {          // large-N / wide-row path (scoreloss_v3.cu); falls through when the shape is outside it
        const int rc = ur::score_loss_v3_try(p, d, loss_type, st);
        if (rc == UR_OK) UR_RETURN_LAST_ERROR();
    }
    switch (d) {
        case 16: ur::launch_score_loss<4>(p, loss_type, smem, st); break;
        case 32: ur::launch_score_loss<8>(p, loss_type, smem, st); break;
        case 64: ur::launch_score_loss<16>(p, loss_type, smem, st); break;
        case 128: ur::launch_score_loss<32>(p, loss_type, smem, st); break;
        case 256: ur::launch_score_loss<64>(p, loss_type, smem, st); break;
        case 512: ur::launch_score_loss<128>(p, loss_type, smem, st); break;
        default: return UR_ERR_UNSUPPORTED;
    }
    UR_RETURN_LAST_ERROR();
}

}  // extern "C"
