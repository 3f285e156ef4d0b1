// Row-sharded scoring, owner side (SURVEY 8e "move queries, not rows"), built on the cp.async.bulk row ring of scoreloss_v3.cu.
//
//   ur_pack_ids_i32         (item id, label) -> one int32 per entry: id | (label > 0) << 31.  The id exchange of a step is ONE
//                           all-gather of 4 bytes per entry (was int64 ids + int32 labels = 12 bytes).
//   ur_score_partial_f32    one CTA per sample of ANY rank: the sample's packed ids are scanned once, the entries this rank owns
//                           (id % W == r) are compacted into shared memory, their rows stream through per-warp mbarrier rings
//                           (one UBLKCP per 512 B / 1 KB row) and fold into an online-softmax partial state
//                           (m, l, sum y s, sum y, sum p e [d], sum y e [d]); raw scores of owned entries go to z.
//   ur_score_merge_f32      home rank: merges the W partial states of each of its samples (received by ONE all-to-all)
//                           -> lse, per-sample loss, dLoss/du.  Replaces all-reduce(MAX) + rescale + reduce-scatter + finish.
//   ur_score_dscore_f32     owner: dLoss/d(dot) of owned entries from the gathered (lse, n_y).
//   ur_count_positive_packed  global number of positive labels from the packed ids.
// BPR under sharding (three small phases, K is small there): ur_shard_scores_f32 (owned raw scores), ur_bpr_from_scores_f32 (home: loss,
// dLoss/ds), ur_shard_grad_user_f32 (owner: partial dLoss/du).
// Arithmetic = scoreloss.cu / scoreloss_v3.cu (reference: unirec/model/base/recommender.py:76-96, reco_abc.py:252-265, modules.py:15-21).
#include <stdlib.h>
#include "common.cuh"

namespace ur {
namespace sr {

constexpr float kEps = 1e-8f;
constexpr uint32_t kIdMask = 0x7FFFFFFFu;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "SR_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra SR_DONE;\n\t"
        "bra SR_WAIT;\n\t"
        "SR_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(256) pack_ids_kernel(const int64_t* __restrict__ item_id, const int32_t* __restrict__ label, int64_t n,
                                                       int N, int32_t* __restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t id = (uint32_t)item_id[e] & kIdMask;
        const bool y = label ? label[e] > 0 : (e % N) == 0;
        out[e] = (int32_t)(id | (y ? 0x80000000u : 0u));
    }
}

__global__ void __launch_bounds__(256) count_positive_packed_kernel(const int32_t* __restrict__ ids, int64_t n, float* __restrict__ out) {
    int c = 0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) c += ids[e] < 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    __shared__ int wsum[8];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += wsum[w];
        if (t) atomicAdd(out, (float)t);
    }
}

struct PartialParams {
    const float4* table;      // local shard [ceil(V/W), d]
    const float4* user_emb;   // [S, d]  all samples of all ranks
    const int32_t* ids;       // [S, N]  packed global ids
    const float* item_bias;   // [V] replicated or null (global id)
    const float* user_bias;   // replicated or null
    const int64_t* user_id;   // [S]
    float inv_tau, clip;
    int N, W, r;
    int64_t S;
    float* z;                 // [S, N] scaled unclamped scores, written for owned entries only
    float* state;             // [S, 4 + 2d]
};

// LR lanes per row in the dot phase, KW warps per CTA, KS ring stages per warp.  Softmax partial only.
template <int D, int LR, int KW, int KS, int MINB>
__global__ void __launch_bounds__(KW * 32, MINB) score_partial_ring_kernel(const PartialParams p) {
    constexpr int D4 = D / 4;
    constexpr int UPL = D4 / LR;
    constexpr int RPC = 32 / LR;
    constexpr int VPL = D4 / 32;
    constexpr int ROWP = D + 4 * LR;
    constexpr int ROWB = D * 4;
    extern __shared__ __align__(128) float smem[];
    const int N = p.N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t b = blockIdx.x;

    constexpr int WARP_FLOATS = KS * RPC * ROWP + 2 * RPC;
    float* my = smem + (size_t)warp * WARP_FLOATS;
    float* ring = my;
    float* wbuf = my + KS * RPC * ROWP;
    float* ybuf = wbuf + RPC;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)KW * WARP_FLOATS);        // [KW][KS]
    float* gstate = reinterpret_cast<float*>(bars + KW * KS);                              // [KW][4]
    int* own_meta = reinterpret_cast<int*>(gstate + KW * 4);                               // [N] local row | label << 31
    uint16_t* own_j = reinterpret_cast<uint16_t*>(own_meta + ((N + 3) & ~3));              // [N] slot of the entry in the sample's row
    int* chunk_cnt = reinterpret_cast<int*>(own_j + ((N + 7) & ~7));                       // [ceil(N/32) + 1] prefix of owned counts
    uint64_t* my_bars = bars + warp * KS;

    if (lane < KS) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(my_bars + lane)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");

    // ---- ordered compaction of the owned entries (deterministic: chunk counts -> prefix -> scatter) ----
    // the packed ids are staged in own_meta first: independent coalesced loads (one HBM latency for the whole row instead of one per
    // 32-id chunk), then both passes run out of shared memory
    const int32_t* ids = p.ids + b * N;
    const int n32 = (N + 31) / 32;
    for (int j = threadIdx.x; j < N; j += KW * 32) own_meta[j] = __ldg(ids + j);
    __syncthreads();
    for (int c = warp; c < n32; c += KW) {
        const int j = c * 32 + lane;
        const uint32_t pk = j < N ? (uint32_t)own_meta[j] : 0u;
        const bool own = j < N && ((pk & kIdMask) % (uint32_t)p.W) == (uint32_t)p.r;
        const unsigned m = __ballot_sync(0xffffffffu, own);
        if (lane == 0) chunk_cnt[c + 1] = __popc(m);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        chunk_cnt[0] = 0;
        for (int c = 1; c <= n32; ++c) { tot += chunk_cnt[c]; chunk_cnt[c] = tot; }
    }
    __syncthreads();
    // in-place scatter: entry k <= j always, chunks are processed in order by ONE warp so that no id is overwritten before it is read
    if (warp == 0) {
        for (int c = 0; c < n32; ++c) {
            const int j = c * 32 + lane;
            const uint32_t pk = j < N ? (uint32_t)own_meta[j] : 0u;
            const uint32_t gid = pk & kIdMask;
            const bool own = j < N && (gid % (uint32_t)p.W) == (uint32_t)p.r;
            const unsigned m = __ballot_sync(0xffffffffu, own);
            __syncwarp();
            if (own) {
                const int k = chunk_cnt[c] + __popc(m & ((1u << lane) - 1));
                own_meta[k] = (int)((gid / (uint32_t)p.W) | (pk & 0x80000000u));
                own_j[k] = (uint16_t)j;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    const int n_own = chunk_cnt[n32];

    const float clip = p.clip;
    const bool has_clip = clip > 0.f;
    const float ub = p.user_bias ? __ldg(p.user_bias + __ldg(p.user_id + b)) : 0.f;
    const int r_dot = lane / LR, h_dot = lane % LR;
    const uint64_t l2_policy = l2_evict_first_policy();      // rows are read once per step (see scoreloss_v3.cu)
    float4 ureg[UPL];
#pragma unroll
    for (int i = 0; i < UPL; ++i) ureg[i] = __ldg(p.user_emb + b * D4 + i * LR + h_dot);

    const int total_chunks = (n_own + RPC - 1) / RPC;
    const int my_chunks = total_chunks > warp ? (total_chunks - warp + KW - 1) / KW : 0;

    float4 acc[VPL], accy[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) { acc[v] = make_float4(0.f, 0.f, 0.f, 0.f); accy[v] = make_float4(0.f, 0.f, 0.f, 0.f); }
    float st_m = -INFINITY, l_part = 0.f, a_part = 0.f, b_part = 0.f;

    int smeta[KS];      // (local row | label) of row `lane` of the chunk staged in slot s
    int sj[KS];         // its slot j in the sample's id row
    auto issue = [&](int i, int slot) {
        const int kbase = (i * KW + warp) * RPC;
        const int nvalid = min(RPC, n_own - kbase);
        const uint32_t bar = smem_u32(my_bars + slot);
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(nvalid * ROWB)) : "memory");
        int meta = 0, jj = 0;
        if (lane < nvalid) { meta = own_meta[kbase + lane]; jj = own_j[kbase + lane]; }
        smeta[slot] = meta; sj[slot] = jj;
        if (lane < nvalid) {
            const uint32_t dst = smem_u32(ring + (size_t)(slot * RPC + lane) * ROWP);
            const int64_t lrow = (int64_t)(meta & 0x7FFFFFFF);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                         ::"r"(dst), "l"(reinterpret_cast<const float*>(p.table) + lrow * D), "r"((uint32_t)ROWB), "r"(bar), "l"(l2_policy) : "memory");
        }
    };
#pragma unroll
    for (int c = 0; c < KS - 1; ++c)
        if (c < my_chunks) issue(c, c);
    __syncwarp();

#pragma unroll 1
    for (int i0 = 0; i0 < my_chunks; i0 += KS) {
#pragma unroll
        for (int ss = 0; ss < KS; ++ss) {
            const int i = i0 + ss;
            if (i >= my_chunks) break;
            const int nslot = (ss + KS - 1) % KS;
            if (i + KS - 1 < my_chunks) issue(i + KS - 1, nslot);
            mbar_wait(smem_u32(my_bars + ss), (i / KS) & 1);
            const int kbase = (i * KW + warp) * RPC;
            const int nvalid = min(RPC, n_own - kbase);
            const float* slot_rows = ring + (size_t)ss * RPC * ROWP;
            const bool ok = r_dot < nvalid;
            float dot = 0.f;
            if (ok) {
                const float4* src = reinterpret_cast<const float4*>(slot_rows + (size_t)r_dot * ROWP) + h_dot;
                float d0 = 0.f, d1 = 0.f;
#pragma unroll
                for (int k = 0; k < UPL; k += 2) {
                    d0 += f4_dot(src[k * LR], ureg[k]);
                    d1 += f4_dot(src[(k + 1) * LR], ureg[k + 1]);
                }
                dot = d0 + d1;
            }
#pragma unroll
            for (int o = 1; o < LR; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            const int rmeta = __shfl_sync(0xffffffffu, smeta[ss], r_dot);
            const int rj = __shfl_sync(0xffffffffu, sj[ss], r_dot);
            const bool rlab = rmeta < 0;
            float s = -INFINITY, mk = 0.f;
            if (ok) {
                float bias = 0.f;
                if (p.item_bias) bias = __ldg(p.item_bias + ((int64_t)(rmeta & 0x7FFFFFFF) * p.W + p.r));
                const float z = (dot + ub + bias) * p.inv_tau;
                s = has_clip ? fminf(fmaxf(z, -clip), clip) : z;
                mk = (has_clip && (z < -clip || z > clip)) ? 0.f : 1.f;
                if (h_dot == 0) p.z[b * N + rj] = z;
            }
            float yw = 0.f;
            const float m_new = fmaxf(st_m, warp_max(s));
            const float sc = __expf(st_m - m_new);
            st_m = m_new;
            const float pj = ok ? __expf(s - m_new) : 0.f;
            l_part *= sc;
#pragma unroll
            for (int v = 0; v < VPL; ++v) acc[v] = f4_scale(acc[v], sc);
            if (h_dot == 0) {
                l_part += pj;
                if (ok && rlab) { a_part += s; b_part += 1.f; yw = mk; }
            }
            const float w = pj * mk;
            const unsigned any_pos = __ballot_sync(0xffffffffu, yw != 0.f);
            if (h_dot == 0 && r_dot < RPC) { wbuf[r_dot] = w; ybuf[r_dot] = yw; }
            __syncwarp();
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
            __syncwarp();
        }
    }

    // ---- combine the KW warp partials -> this rank's partial state of sample b ----
    l_part = warp_sum(l_part); a_part = warp_sum(a_part); b_part = warp_sum(b_part);
    __syncthreads();
    float* gacc = smem;
    float* gaccy = gacc + KW * D;
    if (lane == 0) { gstate[warp * 4 + 0] = st_m; gstate[warp * 4 + 1] = l_part; gstate[warp * 4 + 2] = a_part; gstate[warp * 4 + 3] = b_part; }
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        reinterpret_cast<float4*>(gacc + warp * D)[v * 32 + lane] = acc[v];
        reinterpret_cast<float4*>(gaccy + warp * D)[v * 32 + lane] = accy[v];
    }
    __syncthreads();
    float m_all = -INFINITY;
    for (int q = 0; q < KW; ++q) m_all = fmaxf(m_all, gstate[q * 4]);
    float l_all = 0.f, a_all = 0.f, b_all = 0.f;
    for (int q = 0; q < KW; ++q) {
        const float* gs = gstate + q * 4;
        l_all += gs[1] * (gs[1] > 0.f ? __expf(gs[0] - m_all) : 0.f);
        a_all += gs[2]; b_all += gs[3];
    }
    float* st = p.state + b * (4 + 2 * D);
    if (threadIdx.x == 0) { st[0] = m_all; st[1] = l_all; st[2] = a_all; st[3] = b_all; }
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        float a = 0.f, ay = 0.f;
        for (int q = 0; q < KW; ++q) {
            const float* gs = gstate + q * 4;
            a += gacc[q * D + c] * (gs[1] > 0.f ? __expf(gs[0] - m_all) : 0.f);
            ay += gaccy[q * D + c];
        }
        st[4 + c] = a;
        st[4 + D + c] = ay;
    }
}

template <int D, int LR, int KW, int KS>
static size_t ring_smem(int N) {
    constexpr int RPC = 32 / LR, ROWP = D + 4 * LR;
    const size_t warp_floats = (size_t)KS * RPC * ROWP + 2 * RPC;
    return sizeof(float) * (KW * warp_floats + KW * 4) + sizeof(uint64_t) * KW * KS + sizeof(int) * ((N + 3) & ~3) +
           sizeof(uint16_t) * ((N + 7) & ~7) + sizeof(int) * ((N + 31) / 32 + 2) + 128;
}

template <int D, int LR, int KW, int KS, int MINB>
static int launch_ring(const PartialParams& p, cudaStream_t st) {
    const size_t sm = ring_smem<D, LR, KW, KS>(p.N);
    if (sm > (size_t)(227 * 1024) / MINB - 1024) return UR_ERR_UNSUPPORTED;
    cudaFuncSetAttribute(score_partial_ring_kernel<D, LR, KW, KS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    score_partial_ring_kernel<D, LR, KW, KS, MINB><<<(unsigned)p.S, KW * 32, sm, st>>>(p);
    return UR_OK;
}

// generic fallback (any d % 4 == 0, any N): one warp per sample, rows read straight from global memory
__global__ void __launch_bounds__(256) score_partial_simple_kernel(const PartialParams p, int d4) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= p.S) return;
    const int N = p.N, D = d4 * 4;
    const int32_t* ids = p.ids + b * N;
    const float ub = p.user_bias ? __ldg(p.user_bias + __ldg(p.user_id + b)) : 0.f;
    const bool has_clip = p.clip > 0.f;
    float m = -INFINITY, l = 0.f, a = 0.f, ny = 0.f;
    float* st = p.state + b * (4 + 2 * D);
    // acc / accy live in the output state rows (L2): lane owns columns c = lane + 32 i
    for (int c = lane; c < D; c += 32) { st[4 + c] = 0.f; st[4 + D + c] = 0.f; }
    __syncwarp();
    for (int j = 0; j < N; ++j) {
        const uint32_t pk = (uint32_t)__ldg(ids + j);
        const uint32_t gid = pk & kIdMask;
        if (gid % (uint32_t)p.W != (uint32_t)p.r) continue;
        const float4* row = p.table + (int64_t)(gid / (uint32_t)p.W) * d4;
        float dot = 0.f;
        for (int c = lane; c < d4; c += 32) dot += f4_dot(__ldg(row + c), __ldg(p.user_emb + b * d4 + c));
        dot = warp_sum(dot);
        const float z = (dot + ub + (p.item_bias ? __ldg(p.item_bias + gid) : 0.f)) * p.inv_tau;
        const float s = has_clip ? fminf(fmaxf(z, -p.clip), p.clip) : z;
        const float mk = (has_clip && (z < -p.clip || z > p.clip)) ? 0.f : 1.f;
        if (lane == 0) p.z[b * N + j] = z;
        const float m_new = fmaxf(m, s);
        const float sc = __expf(m - m_new);
        const float pj = __expf(s - m_new);
        l = l * sc + pj;
        m = m_new;
        const bool y = (pk & 0x80000000u) != 0u;
        if (y) { a += s; ny += 1.f; }
        for (int c = lane; c < d4; c += 32) {
            const float4 e = __ldg(row + c);
            float4* ac = reinterpret_cast<float4*>(st + 4) + c;
            *ac = f4_fma(pj * mk, e, f4_scale(*ac, sc));
            if (y && mk != 0.f) {
                float4* ay = reinterpret_cast<float4*>(st + 4 + D) + c;
                *ay = f4_add(*ay, e);
            }
        }
    }
    if (lane == 0) { st[0] = m; st[1] = l; st[2] = a; st[3] = ny; }
}

// home rank: states [W, B, 4+2d] (partial of rank w for local sample b) -> lse, loss, dLoss/du
__global__ void __launch_bounds__(128) score_merge_kernel(const float* __restrict__ states, int W, int64_t B, int d, float inv_tau,
                                                          const float* __restrict__ norm_dev, float* __restrict__ loss_vec,
                                                          float* __restrict__ lse_ny, float* __restrict__ grad_user) {
    const int64_t b = blockIdx.x;
    const int stride = 4 + 2 * d;
    float m_all = -INFINITY;
    for (int w = 0; w < W; ++w) {
        const float* st = states + ((int64_t)w * B + b) * stride;
        if (st[1] > 0.f) m_all = fmaxf(m_all, st[0]);
    }
    float l = 0.f, ys = 0.f, ny = 0.f;
    for (int w = 0; w < W; ++w) {
        const float* st = states + ((int64_t)w * B + b) * stride;
        l += st[1] > 0.f ? st[1] * __expf(st[0] - m_all) : 0.f;
        ys += st[2]; ny += st[3];
    }
    const float lse = m_all + __logf(l);
    const float gscale = inv_tau / *norm_dev;
    if (threadIdx.x == 0) {
        loss_vec[b] = ny * lse - ys;
        lse_ny[2 * b] = lse;
        lse_ny[2 * b + 1] = ny;
    }
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        float a = 0.f, ay = 0.f;
        for (int w = 0; w < W; ++w) {
            const float* st = states + ((int64_t)w * B + b) * stride;
            a += st[1] > 0.f ? st[4 + c] * __expf(st[0] - m_all) : 0.f;
            ay += st[4 + d + c];
        }
        grad_user[b * d + c] = (ny * a / l - ay) * gscale;
    }
}

__global__ void __launch_bounds__(256) score_dscore_kernel(const float* __restrict__ z, const int32_t* __restrict__ ids,
                                                           const float* __restrict__ lse_ny, int64_t S, int N, int W, int r, float inv_tau,
                                                           float clip, const float* __restrict__ norm_dev, float* __restrict__ dscore) {
    const int64_t total = S * N;
    const float gscale = inv_tau / *norm_dev;
    const bool has_clip = clip > 0.f;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        float g = 0.f;
        const uint32_t pk = (uint32_t)ids[e];
        if ((pk & kIdMask) % (uint32_t)W == (uint32_t)r) {
            const int64_t s = e / N;
            const float zz = z[e];
            const float sc = has_clip ? fminf(fmaxf(zz, -clip), clip) : zz;
            const float mk = (has_clip && (zz < -clip || zz > clip)) ? 0.f : 1.f;
            const float y = (pk & 0x80000000u) ? 1.f : 0.f;
            g = (lse_ny[2 * s + 1] * __expf(sc - lse_ny[2 * s]) - y) * mk * gscale;
        }
        dscore[e] = g;
    }
}

// ---------------------------------------------------------------- BPR under sharding (three phases)
// phase 1, owner: z[e] = (u_s . e + ub + ib) / tau for owned entries, 0 elsewhere (summed across ranks by a reduce-scatter)
__global__ void __launch_bounds__(256) shard_scores_kernel(const float4* __restrict__ table, int d4, const float4* __restrict__ user,
                                                           const int32_t* __restrict__ ids, int64_t S, int N,
                                                           const float* __restrict__ item_bias, const float* __restrict__ user_bias,
                                                           const int64_t* __restrict__ user_id, float inv_tau, int W, int r,
                                                           float* __restrict__ z) {
    const int lane = threadIdx.x & 31;
    const int64_t total = S * N;
    for (int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < total; e += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const uint32_t gid = (uint32_t)__ldg(ids + e) & kIdMask;
        float out = 0.f;
        if (gid % (uint32_t)W == (uint32_t)r) {          // warp-uniform
            const int64_t s = e / N;
            const float4* row = table + (int64_t)(gid / (uint32_t)W) * d4;
            float dot = 0.f;
            for (int c = lane; c < d4; c += 32) dot += f4_dot(ldg_stream(row + c), __ldg(user + s * d4 + c));
            dot = warp_sum(dot);
            const float ub = user_bias ? __ldg(user_bias + __ldg(user_id + s)) : 0.f;
            out = (dot + ub + (item_bias ? __ldg(item_bias + gid) : 0.f)) * inv_tau;
        }
        if (lane == 0) z[e] = out;
    }
}

// phase 2, home: z [B, N] complete raw scores -> loss_vec[b] = mean_j -log(eps + sigmoid(s0 - sj)), dscore = dLoss/d(dot)
// (modules.py:15-21; the clamp of recommender.py:94-95 masks the gradient of clamped scores)
__global__ void __launch_bounds__(128) bpr_from_scores_kernel(const float* __restrict__ z, int64_t B, int N, float inv_tau, float clip,
                                                              float norm, float* __restrict__ loss_vec, float* __restrict__ dscore,
                                                              float* __restrict__ scores) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const bool has_clip = clip > 0.f;
    const float gscale = inv_tau / norm;
    const float* zb = z + b * N;
    const float z0 = zb[0];
    const float s0 = has_clip ? fminf(fmaxf(z0, -clip), clip) : z0;
    const float mask0 = (has_clip && (z0 < -clip || z0 > clip)) ? 0.f : 1.f;
    float loss = 0.f, csum = 0.f;
    if (scores) scores[b * N] = s0;
    for (int j = 1; j < N; ++j) {
        const float zj = zb[j];
        const float s = has_clip ? fminf(fmaxf(zj, -clip), clip) : zj;
        const float mk = (has_clip && (zj < -clip || zj > clip)) ? 0.f : 1.f;
        const float sig = 1.f / (1.f + __expf(-(s0 - s)));
        const float c = sig * (1.f - sig) / (kEps + sig);
        loss += -__logf(kEps + sig);
        csum += c;
        dscore[b * N + j] = c * mk * gscale;
        if (scores) scores[b * N + j] = s;
    }
    dscore[b * N] = -csum * mask0 * gscale;
    loss_vec[b] = loss / (float)(N - 1);
}

// phase 3, owner: out[s, :] = sum over owned entries j of dscore[s, j] * e_j  (partial dLoss/du; summed by a reduce-scatter)
__global__ void __launch_bounds__(256) shard_grad_user_kernel(const float4* __restrict__ table, int d4, const int32_t* __restrict__ ids,
                                                              const float* __restrict__ dscore, int64_t S, int N, int W, int r,
                                                              float4* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    for (int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < S; s += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        for (int c0 = 0; c0 < d4; c0 += 32) {
            const int c = c0 + lane;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int j = 0; j < N; ++j) {
                const uint32_t gid = (uint32_t)__ldg(ids + s * N + j) & kIdMask;
                if (gid % (uint32_t)W != (uint32_t)r) continue;
                if (c < d4) acc = f4_fma(__ldg(dscore + s * N + j), ldg_stream(table + (int64_t)(gid / (uint32_t)W) * d4 + c), acc);
            }
            if (c < d4) out[s * d4 + c] = acc;
        }
    }
}

static inline unsigned grid_for(int64_t n, int per_block = 256, int max_waves = 8) {
    int64_t b = (n + per_block - 1) / per_block;
    const int64_t cap = (int64_t)kNumSMs * max_waves;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace sr
}  // namespace ur

extern "C" {

int ur_pack_ids_i32(const int64_t* item_id, const int32_t* label, int64_t B, int N, int32_t* out, void* stream) {
    if (N <= 0) return UR_ERR_BAD_ARG;
    if (B == 0) return UR_OK;
    ur::sr::pack_ids_kernel<<<ur::sr::grid_for(B * N), 256, 0, (cudaStream_t)stream>>>(item_id, label, B * N, N, out);
    UR_RETURN_LAST_ERROR();
}

int ur_count_positive_packed(const int32_t* ids, int64_t n, float* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(out, 0, sizeof(float), st);
    if (n == 0) return UR_OK;
    ur::sr::count_positive_packed_kernel<<<ur::sr::grid_for(n, 256, 2), 256, 0, st>>>(ids, n, out);
    UR_RETURN_LAST_ERROR();
}

int ur_score_partial_f32(const float* table_local, int d, const float* user_emb, const int32_t* ids_packed, int64_t S, int N,
                         const float* item_bias, const float* user_bias, const int64_t* user_id, float tau, float score_clip, int world,
                         int rank, float* z, float* state, void* stream) {
    if (d <= 0 || (d & 3) || N <= 0 || N > 65535 || tau == 0.f || world < 1 || rank < 0 || rank >= world) return UR_ERR_BAD_ARG;
    if (user_bias && !user_id) return UR_ERR_BAD_ARG;
    if (S == 0) return UR_OK;
    ur::sr::PartialParams p;
    p.table = (const float4*)table_local; p.user_emb = (const float4*)user_emb; p.ids = ids_packed;
    p.item_bias = item_bias; p.user_bias = user_bias; p.user_id = user_id; p.inv_tau = 1.f / tau; p.clip = score_clip;
    p.N = N; p.W = world; p.r = rank; p.S = S; p.z = z; p.state = state;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = UR_ERR_UNSUPPORTED;
    static const int variant = getenv("UR_PARTIAL_VARIANT") ? atoi(getenv("UR_PARTIAL_VARIANT")) : 0;     // profiles/scripts/partial_sweep.py
    if (N >= 64) {
        // ~N/W owned rows per sample: few warps with a 2-deep ring each, several CTAs per SM
        if (d == 128) {
            // measured at W = 8, N = 1025, S = 8192 (profiles/r02/partial_sweep.txt): one warp per sample, 8 CTAs per SM = 0.179 ms;
            // 2 warps x 5 CTAs 0.219 ms; 4 warps x 2 CTAs 0.261 ms
            if (variant == 1) rc = ur::sr::launch_ring<128, 2, 4, 2, 2>(p, st);
            else if (variant == 3) rc = ur::sr::launch_ring<128, 2, 2, 3, 4>(p, st);
            else if (variant == 4) rc = ur::sr::launch_ring<128, 2, 1, 3, 7>(p, st);
            else if (variant == 5) rc = ur::sr::launch_ring<128, 2, 2, 2, 5>(p, st);
            if (rc == UR_ERR_UNSUPPORTED) rc = ur::sr::launch_ring<128, 2, 1, 2, 8>(p, st);
            if (rc == UR_ERR_UNSUPPORTED) rc = ur::sr::launch_ring<128, 2, 2, 2, 5>(p, st);
            if (rc == UR_ERR_UNSUPPORTED) rc = ur::sr::launch_ring<128, 2, 2, 2, 2>(p, st);
        } else if (d == 256) {
            if (variant == 1) rc = ur::sr::launch_ring<256, 4, 4, 2, 2>(p, st);
            else if (variant == 2) rc = ur::sr::launch_ring<256, 4, 1, 2, 6>(p, st);
            else if (variant == 3) rc = ur::sr::launch_ring<256, 4, 8, 2, 1>(p, st);
            else if (variant == 4) rc = ur::sr::launch_ring<256, 4, 2, 3, 2>(p, st);
            if (rc == UR_ERR_UNSUPPORTED) rc = ur::sr::launch_ring<256, 4, 2, 2, 3>(p, st);
            if (rc == UR_ERR_UNSUPPORTED) rc = ur::sr::launch_ring<256, 4, 2, 2, 1>(p, st);
        }
    }
    if (rc == UR_ERR_UNSUPPORTED) {
        ur::sr::score_partial_simple_kernel<<<(unsigned)((S + 7) / 8), 256, 0, st>>>(p, d / 4);
        rc = UR_OK;
    }
    if (rc != UR_OK) return rc;
    UR_RETURN_LAST_ERROR();
}

int ur_score_merge_f32(const float* states, int world, int64_t B, int d, float tau, const float* norm_dev, float* loss_vec,
                       float* lse_ny, float* grad_user, void* stream) {
    if (world < 1 || d <= 0 || tau == 0.f) return UR_ERR_BAD_ARG;
    if (B == 0) return UR_OK;
    ur::sr::score_merge_kernel<<<(unsigned)B, 128, 0, (cudaStream_t)stream>>>(states, world, B, d, 1.f / tau, norm_dev, loss_vec, lse_ny, grad_user);
    UR_RETURN_LAST_ERROR();
}

int ur_score_dscore_f32(const float* z, const int32_t* ids_packed, const float* lse_ny, int64_t S, int N, int world, int rank,
                        float tau, float score_clip, const float* norm_dev, float* dscore, void* stream) {
    if (S == 0) return UR_OK;
    ur::sr::score_dscore_kernel<<<ur::sr::grid_for(S * N), 256, 0, (cudaStream_t)stream>>>(z, ids_packed, lse_ny, S, N, world, rank, 1.f / tau,
                                                                                            score_clip, norm_dev, dscore);
    UR_RETURN_LAST_ERROR();
}

int ur_shard_scores_f32(const float* table_local, int d, const float* user_emb, const int32_t* ids_packed, int64_t S, int N,
                        const float* item_bias, const float* user_bias, const int64_t* user_id, float tau, int world, int rank,
                        float* z, void* stream) {
    if (d <= 0 || (d & 3) || N <= 0 || tau == 0.f || world < 1 || rank < 0 || rank >= world) return UR_ERR_BAD_ARG;
    if (user_bias && !user_id) return UR_ERR_BAD_ARG;
    if (S == 0) return UR_OK;
    ur::sr::shard_scores_kernel<<<ur::sr::grid_for(S * N, 8), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)table_local, d / 4, (const float4*)user_emb, ids_packed, S, N, item_bias, user_bias, user_id, 1.f / tau, world, rank, z);
    UR_RETURN_LAST_ERROR();
}

int ur_bpr_from_scores_f32(const float* z, int64_t B, int N, float tau, float score_clip, float norm, float* loss_vec, float* dscore,
                           float* scores, void* stream) {
    if (N < 2 || tau == 0.f || norm == 0.f) return UR_ERR_BAD_ARG;
    if (B == 0) return UR_OK;
    ur::sr::bpr_from_scores_kernel<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(z, B, N, 1.f / tau, score_clip, norm,
                                                                                                   loss_vec, dscore, scores);
    UR_RETURN_LAST_ERROR();
}

int ur_shard_grad_user_f32(const float* table_local, int d, const int32_t* ids_packed, const float* dscore, int64_t S, int N, int world,
                           int rank, float* out, void* stream) {
    if (d <= 0 || (d & 3) || N <= 0 || world < 1 || rank < 0 || rank >= world) return UR_ERR_BAD_ARG;
    if (S == 0) return UR_OK;
    ur::sr::shard_grad_user_kernel<<<ur::sr::grid_for(S, 8), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)table_local, d / 4, ids_packed, dscore, S, N, world, rank, (float4*)out);
    UR_RETURN_LAST_ERROR();
}

}  // extern "C"
