// K5: fused multi-head self-attention for short sequences (L <= 256), forward and backward.
// Reference: unirec/model/modules.py:284-312 (scores / sqrt(d_h) + additive mask -> softmax -> P V) with the mask of
// unirec/model/sequential/sasrec.py:40-57: 0 where the key is a real item (and key <= query when causal), else -10000.
// The [B,H,L,L] score tensor never reaches HBM: K and V of one (sample, head) live in shared memory, one warp owns a
// query row, softmax statistics by warp shuffles.  Backward recomputes P from the saved log-sum-exp.
// Input layout: packed QKV [T, 3d] (row = q | k | v, head h at columns h*dh), output ctx [T, d].
#include "common.cuh"
#include "dropout.cuh"

namespace ur {

constexpr float kMaskAdd = -10000.f;

// head dims below 4 (stock config/model/SASRec.yaml: n_heads 16 -> d_h = d/16, e.g. 2 at d = 32) use 64-bit accesses
template <int DH>
struct HeadVec {
    static constexpr int W = (DH % 4 == 0) ? 4 : 2;
};
template <int DH>
__device__ __forceinline__ void copy_head_vec(float* dst, const float* src) {
    if constexpr (DH % 4 == 0) *reinterpret_cast<float4*>(dst) = __ldg(reinterpret_cast<const float4*>(src));
    else *reinterpret_cast<float2*>(dst) = __ldg(reinterpret_cast<const float2*>(src));
}
template <int DH>
__device__ __forceinline__ float head_dot(const float* a, const float* b) {
    float dot = 0.f;
    if constexpr (DH % 4 == 0) {
#pragma unroll
        for (int c = 0; c < DH; c += 4) dot += f4_dot(*reinterpret_cast<const float4*>(a + c), *reinterpret_cast<const float4*>(b + c));
    } else {
#pragma unroll
        for (int c = 0; c < DH; c += 2) {
            const float2 x = *reinterpret_cast<const float2*>(a + c), y = *reinterpret_cast<const float2*>(b + c);
            dot += fmaf(x.x, y.x, x.y * y.y);
        }
    }
    return dot;
}

template <int DH>
__global__ void __launch_bounds__(256) attn_fwd_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ seq, int Lmax, int H,
                                                       float scale, int causal, int q_tile, int q_only_last,
                                                       float* __restrict__ ctx, float* __restrict__ lse,
                                                       const int32_t* __restrict__ offs, const int32_t* __restrict__ tok_src,
                                                       const float* __restrict__ q_last, const long long* __restrict__ rng,
                                                       float drop_p, int drop_site) {
    constexpr int KS = DH + 4;               // padded row stride: conflict-free float4 reads with lanes over keys
    constexpr int CPL = (DH + 31) / 32;      // output columns per lane
    constexpr int VW = HeadVec<DH>::W;
    extern __shared__ __align__(16) float smem[];
    float* Ks = smem;                        // [Lmax][KS]
    float* Vs = Ks + (size_t)Lmax * KS;      // [Lmax][KS]
    const int Lp = (Lmax + 3) & ~3;          // keeps every sub-buffer 16-byte aligned
    float* madd = Vs + (size_t)Lmax * KS;    // [Lp]
    float* pbuf = madd + Lp;                 // [8][Lp]
    float* qbuf = pbuf + 8 * Lp;             // [8][DH]
    int* posj = reinterpret_cast<int*>(qbuf + 8 * ((DH + 3) & ~3));   // [Lp] original position of live row j (dropout counter)
    const DropCfg dc = drop_cfg(rng, drop_p, drop_site);
    const int bh = blockIdx.x, b = bh / H, h = bh - b * H;
    const int d = H * DH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // packed mode (csrc/pack.cu): the sample owns rows [offs[b], offs[b+1]) of qkv / ctx, its live positions in order (the last
    // one is position L-1); row j is the position tok_src[row0 + j].  Unpacked: rows b*L .. b*L+L-1.
    const int64_t row0 = offs ? (int64_t)offs[b] : (int64_t)b * Lmax;
    const int L = offs ? offs[b + 1] - offs[b] : Lmax;       // live positions of this sample (>= 1)
    const float* base = qkv + row0 * 3 * d + h * DH;

    for (int i = threadIdx.x; i < L * (DH / VW); i += blockDim.x) {
        const int j = i / (DH / VW), c = (i - j * (DH / VW)) * VW;
        copy_head_vec<DH>(Ks + j * KS + c, base + (int64_t)j * 3 * d + d + c);
        copy_head_vec<DH>(Vs + j * KS + c, base + (int64_t)j * 3 * d + 2 * d + c);
    }
    __shared__ int s_jlo;
    if (threadIdx.x == 0) s_jlo = L;
    __syncthreads();
    for (int j = threadIdx.x; j < L; j += blockDim.x) {
        const int64_t src = tok_src ? (int64_t)tok_src[row0 + j] : row0 + j;
        const bool valid = seq[src] > 0;
        madd[j] = valid ? 0.f : kMaskAdd;
        posj[j] = (int)(src - (int64_t)b * Lmax);
        if (valid) atomicMin(&s_jlo, j);
    }
    __syncthreads();
    // Exact work skipping (results identical to the dense reference computation; the mask depends on KEYS only, sasrec.py:40-57):
    //  * a padded query row other than L-1 never reaches the loss (as a key it is masked in the next layer, and only row L-1 of
    //    the last layer feeds the scorer), so its context is written as zeros.  Row L-1 is always computed, padded or not;
    //  * when the query sees at least one unmasked key, every masked key has weight exp(-10000 + ...) == 0 in fp32: keys before
    //    the first real item and (causal) keys after the query are skipped;
    //  * a sequence without any real item (empty history): every key carries the same -10000, which cancels in the softmax ->
    //    all keys participate (j_lo = 0).
    const bool none_valid = s_jlo == L;        // then every position is live: all rows feed the next layer's keys
    const int j_lo = none_valid ? 0 : s_jlo;

    const int q_begin = q_only_last ? L - 1 : blockIdx.y * q_tile;
    const int q_end = q_only_last ? L : min(L, q_begin + q_tile);
    // q_only_last with q_last != null: the query comes from the compact [B, d] buffer and the context goes to row b of ctx
    const bool compact = q_only_last && q_last != nullptr;
    float* pw = pbuf + warp * Lp;
    float* qw = qbuf + warp * DH;
    for (int i = q_begin + warp; i < q_end; i += 8) {
        if (madd[i] != 0.f && i != L - 1 && !none_valid) {   // padded query row (not the output position): dead, keep it finite
#pragma unroll
            for (int r = 0; r < CPL; ++r) {
                const int c = lane + 32 * r;
                if (c < DH) ctx[(row0 + i) * d + h * DH + c] = 0.f;
            }
            if (lane == 0) lse[((int64_t)b * H + h) * Lmax + i] = 0.f;
            continue;
        }
        // (no real key anywhere: future keys carry the same -10000 as the padded ones, so the row attends to all L keys)
        const int j_hi = (causal && !none_valid) ? i : L - 1;        // inclusive
        for (int c = lane; c < DH; c += 32)
            qw[c] = compact ? __ldg(q_last + (int64_t)b * d + h * DH + c) : __ldg(base + (int64_t)i * 3 * d + c);
        __syncwarp();
        float mx = -INFINITY;
        for (int j = j_lo + lane; j <= j_hi; j += 32) {
            const float dot = head_dot<DH>(qw, Ks + j * KS);
            // two roundings like the reference (scores / sqrt(d_h), then + mask): with every key masked the sum lands on the
            // 1e-3 grid of fp32 near -10000 and a fused multiply-add would round differently
            const float s = __fadd_rn(__fmul_rn(dot, scale), madd[j]);
            pw[j] = s;
            mx = fmaxf(mx, s);
        }
        mx = warp_max(mx);
        float sum = 0.f;
        // attention-probability dropout (modules.py:307): element ((b*H + h)*L + query position)*L + key position
        const unsigned long long drow = ((unsigned long long)bh * Lmax + posj[i]) * Lmax;
        for (int j = j_lo + lane; j <= j_hi; j += 32) {
            const float p = __expf(pw[j] - mx);
            sum += p;
            pw[j] = dc.on ? p * drop_mask1(dc, drow + posj[j]) : p;
        }
        sum = warp_sum(sum);
        __syncwarp();
        float o[CPL];
#pragma unroll
        for (int r = 0; r < CPL; ++r) o[r] = 0.f;
        for (int j = j_lo; j <= j_hi; ++j) {
            const float p = pw[j];
#pragma unroll
            for (int r = 0; r < CPL; ++r) {
                const int c = lane + 32 * r;
                if (c < DH) o[r] = fmaf(p, Vs[j * KS + c], o[r]);
            }
        }
        const float inv = 1.f / sum;
#pragma unroll
        for (int r = 0; r < CPL; ++r) {
            const int c = lane + 32 * r;
            if (c < DH) ctx[(compact ? (int64_t)b : row0 + i) * d + h * DH + c] = o[r] * inv;
        }
        if (lane == 0) lse[((int64_t)b * H + h) * Lmax + i] = mx + __logf(sum);
        __syncwarp();
    }
}

// Backward: one CTA per (sample, head).  Query rows are processed in tiles of QT; phase A (warp per query row)
// recomputes P, forms dS and writes dQ; phase B (warp per key, exclusive ownership -> no atomics) accumulates
// dK and dV in registers across all tiles.
template <int DH, int KPW>   // KPW = max keys owned per warp = ceil(L / 8); register cap scales with it
__global__ void __launch_bounds__(256, (KPW <= 8 ? 3 : (KPW <= 16 ? 2 : 1))) attn_bwd_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ seq, int Lmax, int H,
                                                          float scale, int causal, int q_only_last, const float* __restrict__ ctx,
                                                          const float* __restrict__ lse, const float* __restrict__ dctx,
                                                          float* __restrict__ dqkv, const int32_t* __restrict__ offs,
                                                          const int32_t* __restrict__ tok_src, const float* __restrict__ q_last,
                                                          float* __restrict__ dq_last, const long long* __restrict__ rng,
                                                          float drop_p, int drop_site) {
    constexpr int KS = DH + 4;
    constexpr int CPL = (DH + 31) / 32;
    constexpr int QT = 32;
    constexpr int VW = HeadVec<DH>::W;
    constexpr int QS = (DH + 3) & ~3;        // row stride of the query / dO tiles (16-byte aligned rows)
    extern __shared__ __align__(16) float smem[];
    float* Ks = smem;                        // [Lmax][KS]
    float* Vs = Ks + (size_t)Lmax * KS;      // [Lmax][KS]
    const int Lp = (Lmax + 3) & ~3;
    float* madd = Vs + (size_t)Lmax * KS;    // [Lp]
    float* Ps = madd + Lp;                   // [QT][Lp]
    float* dSs = Ps + QT * Lp;               // [QT][Lp]  (already multiplied by scale)
    float* Qs = dSs + QT * Lp;               // [QT][QS]
    float* dOs = Qs + QT * QS;               // [QT][QS]
    int* posj = reinterpret_cast<int*>(dOs + QT * QS);   // [Lp]
    const DropCfg dc = drop_cfg(rng, drop_p, drop_site);
    const int bh = blockIdx.x, b = bh / H, h = bh - b * H;
    const int d = H * DH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row0 = offs ? (int64_t)offs[b] : (int64_t)b * Lmax;     // packed mode: see attn_fwd_kernel
    const int L = offs ? offs[b + 1] - offs[b] : Lmax;
    const bool compact = q_only_last && q_last != nullptr;               // q / ctx / dctx / dq in compact [B, d] buffers
    const float* base = qkv + row0 * 3 * d + h * DH;
    float* dbase = dqkv + row0 * 3 * d + h * DH;

    for (int i = threadIdx.x; i < L * (DH / VW); i += blockDim.x) {
        const int j = i / (DH / VW), c = (i - j * (DH / VW)) * VW;
        copy_head_vec<DH>(Ks + j * KS + c, base + (int64_t)j * 3 * d + d + c);
        copy_head_vec<DH>(Vs + j * KS + c, base + (int64_t)j * 3 * d + 2 * d + c);
    }
    __shared__ int s_jlo;
    if (threadIdx.x == 0) s_jlo = L;
    __syncthreads();
    for (int j = threadIdx.x; j < L; j += blockDim.x) {
        const int64_t src = tok_src ? (int64_t)tok_src[row0 + j] : row0 + j;
        const bool valid = seq[src] > 0;
        madd[j] = valid ? 0.f : kMaskAdd;
        posj[j] = (int)(src - (int64_t)b * Lmax);
        if (valid) atomicMin(&s_jlo, j);
    }

    float dKr[KPW][CPL], dVr[KPW][CPL];
#pragma unroll
    for (int k = 0; k < KPW; ++k)
#pragma unroll
        for (int r = 0; r < CPL; ++r) { dKr[k][r] = 0.f; dVr[k][r] = 0.f; }

    const int q_first = q_only_last ? L - 1 : 0;
    for (int q0 = q_first; q0 < L; q0 += QT) {
        const int nq = min(QT, L - q0);
        __syncthreads();   // previous tile fully consumed (and K/V/madd visible on the first pass)
        for (int i = threadIdx.x; i < nq * (DH / VW); i += blockDim.x) {
            const int ii = i / (DH / VW), c = (i - ii * (DH / VW)) * VW;
            const int64_t row = compact ? (int64_t)b : row0 + q0 + ii;
            copy_head_vec<DH>(Qs + ii * QS + c, compact ? q_last + (int64_t)b * d + h * DH + c : base + (int64_t)(q0 + ii) * 3 * d + c);
            copy_head_vec<DH>(dOs + ii * QS + c, dctx + row * d + h * DH + c);
        }
        __syncthreads();
        // ---- phase A ----  (same exact skipping as the forward kernel: padded query rows have dctx == 0 and contribute
        //                     nothing; masked keys have P == 0 for real query rows)
        const bool none_valid = s_jlo == L;        // then every position is live: all rows feed the next layer's keys
    const int j_lo = none_valid ? 0 : s_jlo;
        for (int ii = warp; ii < nq; ii += 8) {
            const int i = q0 + ii;
            if (madd[i] != 0.f && i != L - 1 && !none_valid) {   // dead padded query row: dQ = 0, no dK/dV contribution
#pragma unroll
                for (int r = 0; r < CPL; ++r) {
                    const int c = lane + 32 * r;
                    if (c < DH) dbase[(int64_t)i * 3 * d + c] = 0.f;
                }
                continue;
            }
            const int j_hi = (causal && !none_valid) ? i : L - 1;
            const int64_t row = compact ? (int64_t)b : row0 + i;
            float dsum = 0.f;
            for (int c = lane; c < DH; c += 32) dsum = fmaf(dOs[ii * QS + c], __ldg(ctx + row * d + h * DH + c), dsum);
            const float Di = warp_sum(dsum);
            const float lse_i = lse[((int64_t)b * H + h) * Lmax + i];
            // with dropout O = (P o M) V: Di = dO . O is unchanged, dP = M o (dO V^T), dV needs P o M (held in Ps)
            const unsigned long long drow = ((unsigned long long)bh * Lmax + posj[i]) * Lmax;
            for (int j = j_lo + lane; j <= j_hi; j += 32) {
                const float dot = head_dot<DH>(Qs + ii * QS, Ks + j * KS);
                float dp = head_dot<DH>(dOs + ii * QS, Vs + j * KS);
                const float p = __expf(__fadd_rn(__fmul_rn(dot, scale), madd[j]) - lse_i);
                float pm = p;
                if (dc.on) { const float m = drop_mask1(dc, drow + posj[j]); pm = p * m; dp *= m; }
                Ps[ii * Lp + j] = pm;
                dSs[ii * Lp + j] = p * (dp - Di) * scale;
            }
            __syncwarp();
            float dq[CPL];
#pragma unroll
            for (int r = 0; r < CPL; ++r) dq[r] = 0.f;
            for (int j = j_lo; j <= j_hi; ++j) {
                const float ds = dSs[ii * Lp + j];
#pragma unroll
                for (int r = 0; r < CPL; ++r) {
                    const int c = lane + 32 * r;
                    if (c < DH) dq[r] = fmaf(ds, Ks[j * KS + c], dq[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < CPL; ++r) {
                const int c = lane + 32 * r;
                if (c < DH) {
                    if (compact) dq_last[(int64_t)b * d + h * DH + c] = dq[r];
                    else dbase[(int64_t)i * 3 * d + c] = dq[r];
                }
            }
        }
        __syncthreads();
        // ---- phase B ----
#pragma unroll
        for (int k = 0; k < KPW; ++k) {
            const int j = warp + 8 * k;
            if (j < L && j >= j_lo) {
                // rows that wrote P/dS for this key: real query rows, and (causal) only rows i >= j
                for (int ii = (causal && !none_valid) ? max(0, j - q0) : 0; ii < nq; ++ii) {
                    if (madd[q0 + ii] != 0.f && q0 + ii != L - 1 && !none_valid) continue;
                    const float ds = dSs[ii * Lp + j], p = Ps[ii * Lp + j];
#pragma unroll
                    for (int r = 0; r < CPL; ++r) {
                        const int c = lane + 32 * r;
                        if (c < DH) {
                            dKr[k][r] = fmaf(ds, Qs[ii * QS + c], dKr[k][r]);
                            dVr[k][r] = fmaf(p, dOs[ii * QS + c], dVr[k][r]);
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < KPW; ++k) {
        const int j = warp + 8 * k;
        if (j < L) {
#pragma unroll
            for (int r = 0; r < CPL; ++r) {
                const int c = lane + 32 * r;
                if (c < DH) {
                    dbase[(int64_t)j * 3 * d + d + c] = dKr[k][r];
                    dbase[(int64_t)j * 3 * d + 2 * d + c] = dVr[k][r];
                }
            }
        }
    }
}

static size_t attn_fwd_smem(int L, int dh) {
    const size_t Lp = (L + 3) & ~3;
    return sizeof(float) * ((size_t)2 * L * (dh + 4) + Lp + 8 * Lp + 8 * ((dh + 3) & ~3) + Lp);
}
static size_t attn_bwd_smem(int L, int dh) {
    const size_t Lp = (L + 3) & ~3;
    return sizeof(float) * ((size_t)2 * L * (dh + 4) + Lp + 2 * 32 * Lp + 2 * 32 * ((dh + 3) & ~3) + Lp);
}

}  // namespace ur

extern "C" {

// q_only_last: compute only query row L-1 (the last encoder layer feeds only [:, -1, :] into the scorer,
// unirec/model/sequential/sasrec.py:74-75); all keys/values are still used.
int ur_attn_fwd_f32(const float* qkv, const int32_t* item_seq, int64_t B, int L, int H, int dh, int causal, int q_only_last,
                    float* ctx, float* lse, const int32_t* offs, const int32_t* tok_src, const float* q_last, const int64_t* rng,
                    float drop_p, int drop_site, void* stream) {
    if ((offs == nullptr) != (tok_src == nullptr)) return UR_ERR_BAD_ARG;
    if (drop_p < 0.f || drop_p >= 1.f) return UR_ERR_BAD_ARG;
    if (L <= 0 || L > 256 || H <= 0) return UR_ERR_BAD_ARG;
    if (B == 0) return UR_OK;
    const size_t smem = ur::attn_fwd_smem(L, dh);
    if (smem > 220 * 1024) return UR_ERR_UNSUPPORTED;
    const float scale = 1.f / sqrtf((float)dh);
    const int q_tile = L <= 64 ? L : 64;
    dim3 grid((unsigned)(B * H), q_only_last ? 1 : (L + q_tile - 1) / q_tile);
    cudaStream_t st = (cudaStream_t)stream;
#define UR_CASE(DH)                                                                                                    \
    case DH:                                                                                                           \
        if (smem > 48 * 1024)                                                                                          \
            cudaFuncSetAttribute(ur::attn_fwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
        ur::attn_fwd_kernel<DH><<<grid, 256, smem, st>>>(qkv, item_seq, L, H, scale, causal, q_tile, q_only_last, ctx, lse,  \
                                                         offs, tok_src, q_last, (const long long*)rng, drop_p, drop_site);  \
        break;
    switch (dh) {
        UR_CASE(2) UR_CASE(4) UR_CASE(8) UR_CASE(16) UR_CASE(32) UR_CASE(64) UR_CASE(128)
        default: return UR_ERR_UNSUPPORTED;
    }
#undef UR_CASE
    UR_RETURN_LAST_ERROR();
}

// dqkv must be zero-filled by the caller when q_only_last (only row L-1 of dQ is written).
int ur_attn_bwd_f32(const float* qkv, const int32_t* item_seq, int64_t B, int L, int H, int dh, int causal, int q_only_last,
                    const float* ctx, const float* lse, const float* dctx, float* dqkv, const int32_t* offs, const int32_t* tok_src,
                    const float* q_last, float* dq_last, const int64_t* rng, float drop_p, int drop_site, void* stream) {
    if ((offs == nullptr) != (tok_src == nullptr) || (q_last == nullptr) != (dq_last == nullptr)) return UR_ERR_BAD_ARG;
    if (drop_p < 0.f || drop_p >= 1.f) return UR_ERR_BAD_ARG;
    if (L <= 0 || L > 256 || H <= 0) return UR_ERR_BAD_ARG;
    if (B == 0) return UR_OK;
    const size_t smem = ur::attn_bwd_smem(L, dh);
    if (smem > 220 * 1024) return UR_ERR_UNSUPPORTED;
    const float scale = 1.f / sqrtf((float)dh);
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)(B * H);
#define UR_LAUNCH(DH, KPW)                                                                                                  \
    do {                                                                                                                    \
        if (smem > 48 * 1024)                                                                                               \
            cudaFuncSetAttribute(ur::attn_bwd_kernel<DH, KPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
        ur::attn_bwd_kernel<DH, KPW><<<grid, 256, smem, st>>>(qkv, item_seq, L, H, scale, causal, q_only_last, ctx, lse, dctx, dqkv, \
                                                              offs, tok_src, q_last, dq_last, (const long long*)rng, drop_p,        \
                                                              drop_site);                                                          \
    } while (0)
#define UR_CASE(DH)                                           \
    case DH:                                                  \
        if (L <= 64) UR_LAUNCH(DH, 8);                        \
        else if (L <= 128) UR_LAUNCH(DH, 16);                 \
        else UR_LAUNCH(DH, 32);                               \
        break;
    switch (dh) {
        UR_CASE(2) UR_CASE(4) UR_CASE(8) UR_CASE(16) UR_CASE(32) UR_CASE(64) UR_CASE(128)
        default: return UR_ERR_UNSUPPORTED;
    }
#undef UR_CASE
#undef UR_LAUNCH
    UR_RETURN_LAST_ERROR();
}

}  // extern "C"
