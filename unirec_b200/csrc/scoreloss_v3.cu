// K8, large-N / wide-row variant of the fused score+loss kernel (d in {128, 256, 512}, N >= 128): the HBM-roofline path.
//
// Same arithmetic and outputs as score_loss_kernel (scoreloss.cu); different work decomposition, built to make the kernel
// memory-bound instead of issue-bound (the v1/v2 kernels spent ~100 warp-instructions per 512-byte row):
//   * rows are staged by the async copy engine: one `cp.async.bulk` (UBLKCP) per row into a per-warp shared-memory ring
//     (KS x RPC rows, rows padded by 16 B so that both access patterns below are bank-conflict free), completion on
//     one mbarrier per stage; the ids / labels of a chunk are fetched one iteration ahead into registers;
//   * "QK" phase: LR = d/64 lanes own one row each chunk (RPC = 32/LR rows per chunk): 16 LDS.128 + 64 FMA per lane,
//     log2(LR) shuffles -> score; no 32-lane reduction per row;
//   * softmax statistics once per chunk (one warp max, per-lane partial sums), weights to shared memory;
//   * "PV" phase: every lane owns d/128 float4 columns and accumulates  sum_r w_r * row_r  (dL/du) from the staged rows.
// ~20 warp-instructions per row instead of ~100; 2 CTAs x 4 warps per SM keep ~135 KB of row data in flight.
#include "common.cuh"

namespace ur {

struct ScoreLossParams {     // must match scoreloss.cu
    const float4* table; const float4* user_emb; const int64_t* item_id; const int32_t* label;
    const float* item_bias; const float* user_bias; const int64_t* user_id; const float* norm_dev;
    float norm_host; float inv_tau, clip; int N; int64_t B; int wps;
    float* scores; float* loss_vec; float* dscore; float4* grad_user;
    int l2_evict_first;
};

namespace v3 {

constexpr float kEps = 1e-8f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "V3_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra V3_DONE;\n\t"
        "bra V3_WAIT;\n\t"
        "V3_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

// LR = lanes per row in the dot phase (each lane owns UPL = d/(4 LR) float4 of the row), KW = warps per CTA
template <int D, int LOSS, int LR, int KW, int KS, int MINB>     // KS = ring stages per warp, MINB = resident CTAs per SM
__global__ void __launch_bounds__(KW * 32, MINB) score_loss_v3_kernel(const ScoreLossParams p) {
    constexpr int D4 = D / 4;
    constexpr int UPL = D4 / LR;             // float4 per lane in the dot phase
    constexpr int RPC = 32 / LR;             // rows per chunk
    constexpr int VPL = D4 / 32;             // float4 columns per lane in the accumulate phase
    constexpr int ROWP = D + 4 * LR;         // padded row stride (floats): with the interleaved part mapping below the 8 lanes
                                             // of one LDS.128 phase (8/LR rows x LR parts) hit 8 distinct 16-byte bank groups
    constexpr int ROWB = D * 4;
    extern __shared__ __align__(128) float smem[];
    const int N = p.N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t b = blockIdx.x;

    // shared layout (per warp): ring [KS][RPC][ROWP] | wbuf [RPC] | ybuf [RPC] ; then CTA-wide: bars, ids, zbuf, gstate
    constexpr int WARP_FLOATS = KS * RPC * ROWP + 2 * RPC;
    float* my = smem + (size_t)warp * WARP_FLOATS;
    float* ring = my;
    float* wbuf = my + KS * RPC * ROWP;
    float* ybuf = wbuf + RPC;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)KW * WARP_FLOATS);     // [KW][KS]
    float* zbuf = reinterpret_cast<float*>(bars + KW * KS);                       // [N] padded to 4
    float* gstate = zbuf + ((N + 3) & ~3);                                                 // [KW][4]
    float* e0buf = gstate + KW * 4;                                                    // [D] (bpr: row 0)
    uint8_t* ylab = reinterpret_cast<uint8_t*>(e0buf + D);                                 // [N] positive flags (softmax epilogue)
    uint64_t* my_bars = bars + warp * KS;

    if (lane < KS) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(my_bars + lane)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const int64_t* ids = p.item_id + b * N;
    const int32_t* lab = p.label ? p.label + b * N : nullptr;
    const float clip = p.clip;
    const bool has_clip = clip > 0.f;
    const float ub = p.user_bias ? __ldg(p.user_bias + __ldg(p.user_id + b)) : 0.f;
    const int r_dot = lane / LR, h_dot = lane % LR;
    // The table rows are read exactly once per step: stream them through L2 with evict-first priority, so that they do not push out
    // the encoder activations the backward pass is about to re-read (whose dirty lines would be written back during this kernel).
    uint64_t l2_policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(l2_policy));

    // u: the 16 float4 this lane needs in the dot phase (interleaved parts: float4 index k*LR + h)
    float4 ureg[UPL];
#pragma unroll
    for (int i = 0; i < UPL; ++i) ureg[i] = __ldg(p.user_emb + b * D4 + i * LR + h_dot);

    float s0 = 0.f, mask0 = 1.f;
    if (LOSS == 1) {      // BPR: s_0 first; every warp computes it (row 0 is one L2-resident row)
        const int64_t id0 = __ldg(ids);
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < UPL; ++i) dot += f4_dot(__ldg(p.table + id0 * D4 + i * LR + h_dot), ureg[i]);
#pragma unroll
        for (int o = 1; o < LR; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        const float z = (dot + ub + (p.item_bias ? __ldg(p.item_bias + id0) : 0.f)) * p.inv_tau;
        s0 = has_clip ? fminf(fmaxf(z, -clip), clip) : z;
        mask0 = (has_clip && (z < -clip || z > clip)) ? 0.f : 1.f;
        if (warp == 0) {
            if (lane == 0) zbuf[0] = z;
            for (int c = lane; c < D4; c += 32) reinterpret_cast<float4*>(e0buf)[c] = __ldg(p.table + id0 * D4 + c);
        }
    }

    const int jstart = LOSS == 1 ? 1 : 0;
    const int total_chunks = (N - jstart + RPC - 1) / RPC;
    const int my_chunks = total_chunks > warp ? (total_chunks - warp + KW - 1) / KW : 0;

    // per-lane running state
    float4 acc[VPL], accy[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) { acc[v] = make_float4(0.f, 0.f, 0.f, 0.f); accy[v] = make_float4(0.f, 0.f, 0.f, 0.f); }
    float st_m = -INFINITY;       // common running max of this warp (softmax)
    float l_part = 0.f;           // per-lane partial of sum exp (softmax) / sum loss (bpr)
    float a_part = 0.f, b_part = 0.f;   // softmax: sum y*s, sum y ; bpr: a = sum c

    // ids / labels of the chunk that will be issued NEXT, fetched one iteration ahead (lane r < RPC holds row r)
    int64_t nid = 0;
    int32_t nlab = 0;
    auto fetch_meta = [&](int i) {
        const int j = jstart + (i * KW + warp) * RPC + lane;
        if (i < my_chunks && lane < RPC && j < N) {
            nid = __ldg(ids + j);
            nlab = LOSS == 0 ? (lab ? __ldg(lab + j) : (j == 0)) : 0;
        }
    };
    // meta of issued chunks travels in registers too: slot s's (id,label) of row `lane`
    int64_t sid[KS];
    int32_t slab[KS];
    auto issue = [&](int i, int slot) {
        const int jbase = jstart + (i * KW + warp) * RPC;
        const int nvalid = min(RPC, N - jbase);
        const uint32_t bar = smem_u32(my_bars + slot);
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(nvalid * ROWB)) : "memory");
        sid[slot] = nid; slab[slot] = nlab;
        if (lane < nvalid) {
            const uint32_t dst = smem_u32(ring + (size_t)(slot * RPC + lane) * ROWP);
            if (p.l2_evict_first)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                             ::"r"(dst), "l"(reinterpret_cast<const float*>(p.table) + nid * D), "r"((uint32_t)ROWB), "r"(bar), "l"(l2_policy) : "memory");
            else
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(reinterpret_cast<const float*>(p.table) + nid * D), "r"((uint32_t)ROWB), "r"(bar) : "memory");
        }
    };

#pragma unroll
    for (int c = 0; c < KS - 1; ++c) {
        fetch_meta(c);
        if (c < my_chunks) issue(c, c);
    }
    fetch_meta(KS - 1);
    __syncwarp();

#pragma unroll 1
    for (int i0 = 0; i0 < my_chunks; i0 += KS) {
#pragma unroll
        for (int ss = 0; ss < KS; ++ss) {            // slot index is compile-time: sid[]/slab[] stay in registers
            const int i = i0 + ss;
            if (i >= my_chunks) break;
            const int nslot = (ss + KS - 1) % KS;
            if (i + KS - 1 < my_chunks) issue(i + KS - 1, nslot);
            fetch_meta(i + KS);
            mbar_wait(smem_u32(my_bars + ss), (i / KS) & 1);
            const int jbase = jstart + (i * KW + warp) * RPC;
            const int nvalid = min(RPC, N - jbase);
            const float* slot_rows = ring + (size_t)ss * RPC * ROWP;

            // ---- dot phase: LR lanes per row ----
            const bool ok = r_dot < nvalid;
            float dot = 0.f;
            if (ok) {
                const float4* src = reinterpret_cast<const float4*>(slot_rows + (size_t)r_dot * ROWP) + h_dot;
                float d0 = 0.f, d1 = 0.f;          // two independent FMA chains
#pragma unroll
                for (int k = 0; k < UPL; k += 2) {
                    d0 += f4_dot(src[k * LR], ureg[k]);
                    d1 += f4_dot(src[(k + 1) * LR], ureg[k + 1]);
                }
                dot = d0 + d1;
            }
#pragma unroll
            for (int o = 1; o < LR; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            // (id, label) of row r_dot live in lane r_dot's registers
            const int64_t rid = __shfl_sync(0xffffffffu, sid[ss], r_dot);
            const int32_t rlab = __shfl_sync(0xffffffffu, slab[ss], r_dot);
            float z = 0.f, s = -INFINITY, mk = 0.f;
            if (ok) {
                const float bias = p.item_bias ? __ldg(p.item_bias + rid) : 0.f;
                z = (dot + ub + bias) * p.inv_tau;
                s = has_clip ? fminf(fmaxf(z, -clip), clip) : z;
                mk = (has_clip && (z < -clip || z > clip)) ? 0.f : 1.f;
                if (h_dot == 0) {
                    zbuf[jbase + r_dot] = z;
                    if (LOSS == 0) ylab[jbase + r_dot] = rlab > 0;     // the epilogue never goes back to global labels
                }
            }
            // ---- statistics, once per chunk ----
            float w = 0.f, yw = 0.f;
            if (LOSS == 0) {
                const float m_new = fmaxf(st_m, warp_max(s));
                const float sc = __expf(st_m - m_new);           // st_m = -inf on the first chunk -> 0
                st_m = m_new;
                const float pj = ok ? __expf(s - m_new) : 0.f;
                l_part *= sc;
#pragma unroll
                for (int v = 0; v < VPL; ++v) acc[v] = f4_scale(acc[v], sc);
                if (h_dot == 0) {
                    l_part += pj;
                    if (ok && rlab > 0) { a_part += s; b_part += 1.f; yw = mk; }
                }
                w = pj * mk;
            } else {
                if (ok) {
                    const float x = s0 - s;
                    const float sig = 1.f / (1.f + __expf(-x));
                    const float c = sig * (1.f - sig) / (kEps + sig);
                    if (h_dot == 0) { l_part += -__logf(kEps + sig); a_part += c; }
                    w = c * mk;
                }
            }
            const unsigned any_pos = __ballot_sync(0xffffffffu, yw != 0.f);
            if (h_dot == 0 && r_dot < RPC) { wbuf[r_dot] = w; ybuf[r_dot] = yw; }
            __syncwarp();
            // ---- accumulate phase: lane owns VPL float4 columns ----
#pragma unroll 4
            for (int r = 0; r < nvalid; ++r) {
                const float wr = wbuf[r];
                const float4* row = reinterpret_cast<const float4*>(slot_rows + (size_t)r * ROWP);
#pragma unroll
                for (int v = 0; v < VPL; ++v) acc[v] = f4_fma(wr, row[v * 32 + lane], acc[v]);
            }
            if (any_pos) {
                for (int r = 0; r < nvalid; ++r) {
                    const float yr = ybuf[r];
                    if (yr != 0.f) {
                        const float4* row = reinterpret_cast<const float4*>(slot_rows + (size_t)r * ROWP);
#pragma unroll
                        for (int v = 0; v < VPL; ++v) accy[v] = f4_fma(yr, row[v * 32 + lane], accy[v]);
                    }
                }
            }
            __syncwarp();      // slot and wbuf/ybuf are free again
        }
    }

    // ---- warp totals -> shared, then CTA combine (KW partial states) ----
    l_part = warp_sum(l_part); a_part = warp_sum(a_part); b_part = warp_sum(b_part);
    __syncthreads();           // every ring idle: reuse warp 0..3 ring space for the per-warp accumulators
    float* gacc = smem;        // [KW][D] then [KW][D] for accy   (fits: KW*2*D floats << ring)
    float* gaccy = gacc + KW * D;
    if (lane == 0) { gstate[warp * 4 + 0] = st_m; gstate[warp * 4 + 1] = l_part; gstate[warp * 4 + 2] = a_part; gstate[warp * 4 + 3] = b_part; }
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        reinterpret_cast<float4*>(gacc + warp * D)[v * 32 + lane] = acc[v];
        reinterpret_cast<float4*>(gaccy + warp * D)[v * 32 + lane] = accy[v];
    }
    __syncthreads();

    const int tis = threadIdx.x, tps = blockDim.x;
    const float norm = p.norm_dev ? __ldg(p.norm_dev) : p.norm_host;
    const float gscale = p.inv_tau / norm;
    float m_all = -INFINITY;
    if (LOSS == 0)
        for (int q = 0; q < KW; ++q) m_all = fmaxf(m_all, gstate[q * 4]);
    float l_all = 0.f, a_all = 0.f, b_all = 0.f;
    for (int q = 0; q < KW; ++q) {
        const float* gs = gstate + q * 4;
        if (LOSS == 0) l_all += gs[1] * (gs[1] > 0.f ? __expf(gs[0] - m_all) : 0.f);
        else l_all += gs[1];
        a_all += gs[2]; b_all += gs[3];
    }
    if (LOSS == 0) {
        const float lse = m_all + __logf(l_all);
        if (p.grad_user) {
            for (int c = tis; c < D; c += tps) {
                float a = 0.f, ay = 0.f;
                for (int q = 0; q < KW; ++q) {
                    const float* gs = gstate + q * 4;
                    a += gacc[q * D + c] * (gs[1] > 0.f ? __expf(gs[0] - m_all) : 0.f);
                    ay += gaccy[q * D + c];
                }
                reinterpret_cast<float*>(p.grad_user)[b * D + c] = (b_all * a / l_all - ay) * gscale;
            }
        }
        if (tis == 0) p.loss_vec[b] = b_all * lse - a_all;
        for (int j = tis; j < N; j += tps) {
            const float z = zbuf[j];
            const float s = has_clip ? fminf(fmaxf(z, -clip), clip) : z;
            if (p.scores) p.scores[b * N + j] = s;
            if (p.dscore) {
                const float mk = (has_clip && (z < -clip || z > clip)) ? 0.f : 1.f;
                const float yj = (float)ylab[j];
                p.dscore[b * N + j] = (b_all * __expf(s - lse) - yj) * mk * gscale;
            }
        }
    } else {
        const float K = (float)(N - 1);
        if (p.grad_user) {
            for (int c = tis; c < D; c += tps) {
                float a = 0.f;
                for (int q = 0; q < KW; ++q) a += gacc[q * D + c];
                reinterpret_cast<float*>(p.grad_user)[b * D + c] = (a - a_all * mask0 * e0buf[c]) * gscale;
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
                    const float mk = (has_clip && (z < -clip || z > clip)) ? 0.f : 1.f;
                    const float x = s0 - s;
                    const float sig = 1.f / (1.f + __expf(-x));
                    gj = sig * (1.f - sig) / (kEps + sig) * mk * gscale;
                }
                p.dscore[b * N + j] = gj;
            }
        }
    }
}

template <int D, int LR, int KW, int KS>
static size_t smem_bytes(int N) {
    constexpr int RPC = 32 / LR, ROWP = D + 4 * LR;
    const size_t warp_floats = (size_t)KS * RPC * ROWP + 2 * RPC;
    return sizeof(float) * (KW * warp_floats + (size_t)((N + 3) & ~3) + KW * 4 + D) + sizeof(uint64_t) * KW * KS +
           (size_t)((N + 15) & ~15) + 128;
}

template <int D, int LR, int KW, int KS, int MINB = 2>
static int launch(const ScoreLossParams& p, int loss_type, cudaStream_t st) {
    const size_t sm = smem_bytes<D, LR, KW, KS>(p.N);
    if (sm > (size_t)(227 * 1024) / MINB - 1024) return UR_ERR_UNSUPPORTED;      // MINB CTAs per SM (1 KB reserved per CTA)
    if (loss_type == 0) {
        cudaFuncSetAttribute(score_loss_v3_kernel<D, 0, LR, KW, KS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        score_loss_v3_kernel<D, 0, LR, KW, KS, MINB><<<(unsigned)p.B, KW * 32, sm, st>>>(p);
    } else {
        cudaFuncSetAttribute(score_loss_v3_kernel<D, 1, LR, KW, KS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        score_loss_v3_kernel<D, 1, LR, KW, KS, MINB><<<(unsigned)p.B, KW * 32, sm, st>>>(p);
    }
    return UR_OK;
}

}  // namespace v3

// tuning switch (ur_score_loss_set_bulk(10 + variant)), d = 128:
//   0: 2 lanes/row, 4 warps, 3 stages   1: 4 lanes/row, 8 warps, 3 stages   2: 2 lanes/row, 3 warps, 4 stages
//   3: 2 lanes/row, 6 warps, 2 stages   -1 (default): 0 for N >= 512, 1 below
int g_v3_variant = -1;

// called by ur_score_loss_fwd_bwd_f32 (scoreloss.cu); returns UR_ERR_UNSUPPORTED when the shape is outside this kernel
int score_loss_v3_try(const ScoreLossParams& p, int d, int loss_type, cudaStream_t st) {
    if (p.N < 128) return UR_ERR_UNSUPPORTED;
    // measured on B200 (profiles/score_sweep.py, d=128, N=1025): variant 3 = 76.6% of the HBM copy peak, 0 = 63%, 1 = 48%, 2 = 52%:
    // twelve resident warps per SM with a 2-deep ring beat fewer warps with deeper rings.
    const int var = g_v3_variant >= 0 ? g_v3_variant : 3;
    switch (d) {
        case 128:
            if (var == 0) return v3::launch<128, 2, 4, 3>(p, loss_type, st);
            if (var == 1) return v3::launch<128, 4, 8, 3>(p, loss_type, st);
            if (var == 2) return v3::launch<128, 2, 3, 4>(p, loss_type, st);
            // single-wave shapes: 7 small CTAs per SM hold all of B <= 1036 samples at once (no tail wave)
            if (var == 4) { const int rc = v3::launch<128, 4, 2, 2, 7>(p, loss_type, st); if (rc != UR_ERR_UNSUPPORTED) return rc; }
            if (var == 5) { const int rc = v3::launch<128, 2, 1, 2, 7>(p, loss_type, st); if (rc != UR_ERR_UNSUPPORTED) return rc; }
            if (var == 6) { const int rc = v3::launch<128, 4, 2, 2, 4>(p, loss_type, st); if (rc != UR_ERR_UNSUPPORTED) return rc; }
            if (var == 7) { const int rc = v3::launch<128, 2, 2, 2, 5>(p, loss_type, st); if (rc != UR_ERR_UNSUPPORTED) return rc; }
            // one big CTA per SM (round 2: the d = 256 sweep favoured 8 warps x 1 CTA over 6 warps x 2 CTAs)
            if (var == 8) { const int rc = v3::launch<128, 2, 12, 2, 1>(p, loss_type, st); if (rc != UR_ERR_UNSUPPORTED) return rc; }
            if (var == 9) { const int rc = v3::launch<128, 2, 8, 3, 1>(p, loss_type, st); if (rc != UR_ERR_UNSUPPORTED) return rc; }
            if (var == 16) { const int rc = v3::launch<128, 2, 10, 2, 1>(p, loss_type, st); if (rc != UR_ERR_UNSUPPORTED) return rc; }
            if (var == 17) { const int rc = v3::launch<128, 2, 8, 2, 1>(p, loss_type, st); if (rc != UR_ERR_UNSUPPORTED) return rc; }
            return v3::launch<128, 2, 6, 2>(p, loss_type, st);
        case 256:          // tuning switch for the wide row (profiles/score_sweep.py D=256)
            if (var == 10) { const int rc = v3::launch<256, 4, 6, 3, 1>(p, loss_type, st); if (rc != UR_ERR_UNSUPPORTED) return rc; }
            if (var == 11) { const int rc = v3::launch<256, 4, 4, 3, 2>(p, loss_type, st); if (rc != UR_ERR_UNSUPPORTED) return rc; }
            if (var == 12) { const int rc = v3::launch<256, 4, 3, 4, 2>(p, loss_type, st); if (rc != UR_ERR_UNSUPPORTED) return rc; }
            if (var == 13) { const int rc = v3::launch<256, 8, 6, 2, 2>(p, loss_type, st); if (rc != UR_ERR_UNSUPPORTED) return rc; }
            if (var == 14) { const int rc = v3::launch<256, 4, 8, 2, 1>(p, loss_type, st); if (rc != UR_ERR_UNSUPPORTED) return rc; }
            if (var == 15) { const int rc = v3::launch<256, 4, 4, 2, 3>(p, loss_type, st); if (rc != UR_ERR_UNSUPPORTED) return rc; }
            // default (measured, profiles/r02/score_sweep_d256.txt, N = 4097): 8 warps x 2-deep ring, ONE CTA per SM = 88% of the HBM copy peak
            // (6 warps x 2 CTAs/SM does not fit next to a 4097-entry score buffer and fell back to the register-staged kernel: 39%)
            { const int rc = v3::launch<256, 4, 8, 2, 1>(p, loss_type, st); if (rc != UR_ERR_UNSUPPORTED) return rc; }
            return v3::launch<256, 4, 6, 2>(p, loss_type, st);
        case 512: return v3::launch<512, 8, 6, 2>(p, loss_type, st);
        default: return UR_ERR_UNSUPPORTED;
    }
}

}  // namespace ur
