// LayerNorm-family kernels (K3, K6): warp-per-row, warp-shuffle statistics, 128-bit accesses.
//   seq_prep_ln: Y = LN(E[item_seq] + P[0..L-1])      unirec/model/sequential/sasrec.py:60-68
//   add_ln:      Y = LN(X + R)  (post-LN residual)    unirec/model/modules.py:313-314, 352-353
#include "common.cuh"
#include "dropout.cuh"

namespace ur {

template <int MAXV>
struct RowRegs {
    float4 v[MAXV];
};

// mean / rstd of a row held across a warp (lane owns float4 c = lane + 32*i < d4)
template <int MAXV>
__device__ __forceinline__ void row_stats(const RowRegs<MAXV>& x, int d4, int lane, float eps, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (lane + 32 * i < d4) s += (x.v[i].x + x.v[i].y) + (x.v[i].z + x.v[i].w);
    const float inv_d = 1.f / (float)(d4 * 4);
    mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (lane + 32 * i < d4) {
            float a = x.v[i].x - mean, b = x.v[i].y - mean, c = x.v[i].z - mean, e = x.v[i].w - mean;
            q += (a * a + b * b) + (c * c + e * e);
        }
    rstd = rsqrtf(warp_sum(q) * inv_d + eps);
}

template <int MAXV>
__device__ __forceinline__ void ln_apply_store(const RowRegs<MAXV>& x, int d4, int lane, float mean, float rstd,
                                               const float4* __restrict__ gamma, const float4* __restrict__ beta,
                                               float4* __restrict__ y, const DropCfg& dc = DropCfg{0, 0, 0, 0, 0, 1.f, false},
                                               unsigned long long drop_row = 0) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        int c = lane + 32 * i;
        if (c < d4) {
            float4 g = __ldg(gamma + c), b = __ldg(beta + c), o;
            o.x = (x.v[i].x - mean) * rstd * g.x + b.x;
            o.y = (x.v[i].y - mean) * rstd * g.y + b.y;
            o.z = (x.v[i].z - mean) * rstd * g.z + b.z;
            o.w = (x.v[i].w - mean) * rstd * g.w + b.w;
            if (dc.on) o = f4_mul(o, drop_mask4(dc, drop_row * (unsigned long long)d4 + c));   // dropout(LN(.)), sasrec.py:68-69
            y[c] = o;
        }
    }
}

// ------------------------------------------------------------------ seq prep forward
template <int MAXV>
__global__ void __launch_bounds__(256) seq_prep_ln_fwd_kernel(const float4* __restrict__ table, const float4* __restrict__ pos,
                                                              const float4* __restrict__ gamma, const float4* __restrict__ beta,
                                                              float eps, const int32_t* __restrict__ seq, int64_t rows, int L, int d4,
                                                              float4* __restrict__ Y, float* __restrict__ mean_out,
                                                              float* __restrict__ rstd_out, const int32_t* __restrict__ tok_src,
                                                              const int32_t* __restrict__ n_tok_dev,
                                                              const long long* __restrict__ shard_ptrs, int shard_world,
                                                              const long long* __restrict__ rng, float drop_p, int drop_site) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const DropCfg dc = drop_cfg(rng, drop_p, drop_site);
    // packed mode (csrc/pack.cu): output row t holds position tok_src[t]; rows [n_tok, roundup32(n_tok)) are written as zeros
    // (the token-reduction GEMMs consume whole 32-row blocks)
    const int64_t n_live = n_tok_dev ? min((int64_t)*n_tok_dev, rows) : rows;
    const int64_t n_pad = n_tok_dev ? min(rows, (n_live + 31) & ~(int64_t)31) : rows;
    for (int64_t r = warp0; r < n_pad; r += nwarps) {
        if (r >= n_live) {
#pragma unroll
            for (int i = 0; i < MAXV; ++i) {
                int c = lane + 32 * i;
                if (c < d4) Y[r * d4 + c] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            continue;
        }
        const int64_t src = tok_src ? (int64_t)__ldg(tok_src + r) : r;
        const int64_t id = __ldg(seq + src);
        const int l = (int)(src % L);
        // row-sharded table: the row lives on rank id % W at local index id / W, read straight from the peer's HBM over NVLink
        const float4* trow = shard_ptrs ? reinterpret_cast<const float4*>(shard_ptrs[id % shard_world]) + (id / shard_world) * d4
                                        : table + id * d4;
        RowRegs<MAXV> x;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            int c = lane + 32 * i;
            if (c < d4) {
                x.v[i] = ldg_stream(trow + c);
                if (pos) x.v[i] = f4_add(x.v[i], __ldg(pos + (int64_t)l * d4 + c));
            }
        }
        float mean, rstd;
        row_stats<MAXV>(x, d4, lane, eps, mean, rstd);
        ln_apply_store<MAXV>(x, d4, lane, mean, rstd, gamma, beta, Y + r * d4, dc, (unsigned long long)src);
        if (lane == 0) { mean_out[r] = mean; rstd_out[r] = rstd; }
    }
}

// LN backward for one row held in registers: dx = rstd * (g - mean(g) - xhat * mean(g*xhat)), g = dy*gamma
template <int MAXV>
__device__ __forceinline__ void ln_bwd_row(const RowRegs<MAXV>& x, const RowRegs<MAXV>& dy, int d4, int lane, float mean,
                                           float rstd, const float4* __restrict__ gamma, RowRegs<MAXV>& dx,
                                           RowRegs<MAXV>& dgam, RowRegs<MAXV>& dbet) {
    float s1 = 0.f, s2 = 0.f;
    RowRegs<MAXV> xh;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        int c = lane + 32 * i;
        if (c < d4) {
            float4 g = __ldg(gamma + c);
            xh.v[i].x = (x.v[i].x - mean) * rstd; xh.v[i].y = (x.v[i].y - mean) * rstd;
            xh.v[i].z = (x.v[i].z - mean) * rstd; xh.v[i].w = (x.v[i].w - mean) * rstd;
            float4 t = make_float4(dy.v[i].x * g.x, dy.v[i].y * g.y, dy.v[i].z * g.z, dy.v[i].w * g.w);
            s1 += (t.x + t.y) + (t.z + t.w);
            s2 += (t.x * xh.v[i].x + t.y * xh.v[i].y) + (t.z * xh.v[i].z + t.w * xh.v[i].w);
            dgam.v[i].x += dy.v[i].x * xh.v[i].x; dgam.v[i].y += dy.v[i].y * xh.v[i].y;
            dgam.v[i].z += dy.v[i].z * xh.v[i].z; dgam.v[i].w += dy.v[i].w * xh.v[i].w;
            dbet.v[i] = f4_add(dbet.v[i], dy.v[i]);
            dx.v[i] = t;   // holds g for now
        }
    }
    const float inv_d = 1.f / (float)(d4 * 4);
    s1 = warp_sum(s1) * inv_d;
    s2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        int c = lane + 32 * i;
        if (c < d4) {
            dx.v[i].x = rstd * (dx.v[i].x - s1 - xh.v[i].x * s2);
            dx.v[i].y = rstd * (dx.v[i].y - s1 - xh.v[i].y * s2);
            dx.v[i].z = rstd * (dx.v[i].z - s1 - xh.v[i].z * s2);
            dx.v[i].w = rstd * (dx.v[i].w - s1 - xh.v[i].w * s2);
        }
    }
}

// block-level accumulate of per-warp column partials into a global vector (one atomic per column per CTA)
template <int MAXV>
__device__ __forceinline__ void block_accumulate(const RowRegs<MAXV>& part, int d4, float* __restrict__ gout, float* smem /*[d4*4]*/) {
    const int lane = threadIdx.x & 31;
    for (int c = threadIdx.x; c < d4 * 4; c += blockDim.x) smem[c] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        int c = lane + 32 * i;
        if (c < d4) {
            atomicAdd(smem + c * 4 + 0, part.v[i].x); atomicAdd(smem + c * 4 + 1, part.v[i].y);
            atomicAdd(smem + c * 4 + 2, part.v[i].z); atomicAdd(smem + c * 4 + 3, part.v[i].w);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < d4 * 4; c += blockDim.x) atomicAdd(gout + c, smem[c]);
    __syncthreads();
}

template <int MAXV>
__device__ __forceinline__ void zero_regs(RowRegs<MAXV>& r) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i) r.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------ seq prep backward
// grid = (nbx, L): every CTA works on one position l so the position-table gradient is reduced in registers.
template <int MAXV>
__global__ void __launch_bounds__(256) seq_prep_ln_bwd_kernel(const float4* __restrict__ table, const float4* __restrict__ pos,
                                                              const float4* __restrict__ gamma, const int32_t* __restrict__ seq,
                                                              int64_t B, int L, int d4, const float* __restrict__ mean_in,
                                                              const float* __restrict__ rstd_in, const float4* __restrict__ dY,
                                                              float4* __restrict__ dX, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta, float* __restrict__ dpos,
                                                              const int32_t* __restrict__ tok_inv,
                                                              const long long* __restrict__ shard_ptrs, int shard_world,
                                                              const long long* __restrict__ rng, float drop_p, int drop_site) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31;
    const int l = blockIdx.y;
    const DropCfg dc = drop_cfg(rng, drop_p, drop_site);
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    RowRegs<MAXV> dgam, dbet, dp;
    zero_regs(dgam); zero_regs(dbet); zero_regs(dp);
    for (int64_t b = warp0; b < B; b += nwarps) {
        const int64_t r = b * L + l;
        // packed mode: dY / mean / rstd are indexed by the packed token, dX (the table-row gradient the optimizer reads by batch
        // entry) keeps the [B*L, d] layout; dead positions carry no gradient and are never read (their key is the padding id)
        const int64_t t = tok_inv ? (int64_t)__ldg(tok_inv + r) : r;
        if (t < 0) continue;
        const int64_t id = __ldg(seq + r);
        const float4* trow = shard_ptrs ? reinterpret_cast<const float4*>(shard_ptrs[id % shard_world]) + (id / shard_world) * d4
                                        : table + id * d4;
        RowRegs<MAXV> x, dy, dx;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            int c = lane + 32 * i;
            if (c < d4) {
                x.v[i] = ldg_stream(trow + c);
                if (pos) x.v[i] = f4_add(x.v[i], __ldg(pos + (int64_t)l * d4 + c));
                dy.v[i] = dY[t * d4 + c];
                if (dc.on) dy.v[i] = f4_mul(dy.v[i], drop_mask4(dc, (unsigned long long)r * d4 + c));   // the forward's mask of position r
            }
        }
        ln_bwd_row<MAXV>(x, dy, d4, lane, mean_in[t], rstd_in[t], gamma, dx, dgam, dbet);
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            int c = lane + 32 * i;
            if (c < d4) {
                dX[r * d4 + c] = dx.v[i];
                dp.v[i] = f4_add(dp.v[i], dx.v[i]);
            }
        }
    }
    block_accumulate<MAXV>(dgam, d4, dgamma, smem);
    block_accumulate<MAXV>(dbet, d4, dbeta, smem);
    if (dpos) block_accumulate<MAXV>(dp, d4, dpos + (int64_t)l * d4 * 4, smem);
}

// ------------------------------------------------------------------ residual add + LN
// Z = dropout(X) + R is written back over X (kept for backward); Y = LN(Z).  Dropout (modules.py:313, :352) is indexed by the
// ORIGINAL position of the row: row_pos[r] (packed token map) or r * pos_mul + pos_add (compact last-layer rows).
template <int MAXV>
__global__ void __launch_bounds__(256) add_ln_fwd_kernel(float* __restrict__ X, int64_t ldx, const float* __restrict__ R, int64_t ldr,
                                                         const float4* __restrict__ gamma, const float4* __restrict__ beta, float eps,
                                                         int64_t rows, int d4, float* __restrict__ Y, int64_t ldy,
                                                         float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                         const int32_t* __restrict__ rows_dev, const long long* __restrict__ rng,
                                                         float drop_p, int drop_site, const int32_t* __restrict__ row_pos,
                                                         int64_t pos_mul, int64_t pos_add) {
    const DropCfg dc = drop_cfg(rng, drop_p, drop_site);
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    // rows_dev: device-resident live row count (packed tokens); rows up to the next multiple of 32 are written as zeros
    const int64_t n_live = rows_dev ? min((int64_t)*rows_dev, rows) : rows;
    const int64_t n_pad = rows_dev ? min(rows, (n_live + 31) & ~(int64_t)31) : rows;
    for (int64_t r = warp0; r < n_pad; r += nwarps) {
        if (r >= n_live) {
            float4* yr = reinterpret_cast<float4*>(Y + r * ldy);
#pragma unroll
            for (int i = 0; i < MAXV; ++i) {
                int c = lane + 32 * i;
                if (c < d4) yr[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            continue;
        }
        float4* xr = reinterpret_cast<float4*>(X + r * ldx);
        const float4* rr = R ? reinterpret_cast<const float4*>(R + r * ldr) : nullptr;
        RowRegs<MAXV> x;
        const unsigned long long prow = dc.on ? (unsigned long long)(row_pos ? (int64_t)__ldg(row_pos + r) : r * pos_mul + pos_add) : 0ull;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            int c = lane + 32 * i;
            if (c < d4) {
                x.v[i] = xr[c];
                if (dc.on) x.v[i] = f4_mul(x.v[i], drop_mask4(dc, prow * d4 + c));
                if (rr) x.v[i] = f4_add(x.v[i], rr[c]);
                if (rr || dc.on) xr[c] = x.v[i];
            }
        }
        float mean, rstd;
        row_stats<MAXV>(x, d4, lane, eps, mean, rstd);
        ln_apply_store<MAXV>(x, d4, lane, mean, rstd, gamma, beta, reinterpret_cast<float4*>(Y + r * ldy));
        if (lane == 0) { mean_out[r] = mean; rstd_out[r] = rstd; }
    }
}

// dZ = LN'(Z) dY (may alias dY); optional dZ += dExtra (gradient arriving through the residual branch of the NEXT op).
// With dropout on the linear branch (Z = dropout(X) + R): dZdrop = dZ * mask is the gradient of X (feeds the weight / input
// gradients of the producing linear layer and its bias gradient dzsum); dZ itself flows down the residual branch.
template <int MAXV>
__global__ void __launch_bounds__(256) add_ln_bwd_kernel(const float* __restrict__ Z, int64_t ldz, const float4* __restrict__ gamma,
                                                         const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                         const float* dY, int64_t lddy, const float* __restrict__ dExtra, int64_t ldde,
                                                         int64_t rows, int d4, float* dZ, int64_t lddz,
                                                         float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                         float* __restrict__ dzsum, const int32_t* __restrict__ rows_dev,
                                                         float* __restrict__ dZdrop, int64_t lddzd, const long long* __restrict__ rng,
                                                         float drop_p, int drop_site, const int32_t* __restrict__ row_pos,
                                                         int64_t pos_mul, int64_t pos_add) {
    extern __shared__ float smem[];
    const DropCfg dc = drop_cfg(dZdrop ? rng : nullptr, drop_p, drop_site);
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    RowRegs<MAXV> dgam, dbet, dzs;
    zero_regs(dgam); zero_regs(dbet); zero_regs(dzs);
    const int64_t n_live = rows_dev ? min((int64_t)*rows_dev, rows) : rows;
    const int64_t n_pad = rows_dev ? min(rows, (n_live + 31) & ~(int64_t)31) : rows;
    for (int64_t r = warp0; r < n_pad; r += nwarps) {
        if (r >= n_live) {                         // zero tail: dZ is the token-reduction operand of the weight-gradient GEMM
            float4* dzt = reinterpret_cast<float4*>(dZ + r * lddz);
            float4* dzdt = dc.on ? reinterpret_cast<float4*>(dZdrop + r * lddzd) : nullptr;
#pragma unroll
            for (int i = 0; i < MAXV; ++i) {
                int c = lane + 32 * i;
                if (c < d4) {
                    dzt[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (dzdt) dzdt[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            continue;
        }
        const float4* zr = reinterpret_cast<const float4*>(Z + r * ldz);
        const float4* dyr = reinterpret_cast<const float4*>(dY + r * lddy);
        const float4* der = dExtra ? reinterpret_cast<const float4*>(dExtra + r * ldde) : nullptr;
        RowRegs<MAXV> x, dy, dx;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            int c = lane + 32 * i;
            if (c < d4) {
                x.v[i] = zr[c];
                dy.v[i] = dyr[c];
                if (der) dy.v[i] = f4_add(dy.v[i], der[c]);
            }
        }
        ln_bwd_row<MAXV>(x, dy, d4, lane, mean_in[r], rstd_in[r], gamma, dx, dgam, dbet);
        float4* dzr = reinterpret_cast<float4*>(dZ + r * lddz);
        float4* dzdr = dc.on ? reinterpret_cast<float4*>(dZdrop + r * lddzd) : nullptr;
        const unsigned long long prow = dc.on ? (unsigned long long)(row_pos ? (int64_t)__ldg(row_pos + r) : r * pos_mul + pos_add) : 0ull;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            int c = lane + 32 * i;
            if (c < d4) {
                dzr[c] = dx.v[i];
                float4 gx = dx.v[i];
                if (dc.on) { gx = f4_mul(gx, drop_mask4(dc, prow * d4 + c)); dzdr[c] = gx; }
                dzs.v[i] = f4_add(dzs.v[i], gx);
            }
        }
    }
    block_accumulate<MAXV>(dgam, d4, dgamma, smem);
    block_accumulate<MAXV>(dbet, d4, dbeta, smem);
    if (dzsum) block_accumulate<MAXV>(dzs, d4, dzsum, smem);     // bias gradient of the linear layer that produced Z
}

static inline int ln_grid(int64_t rows) {
    int64_t g = (rows + 7) / 8;
    int64_t cap = (int64_t)kNumSMs * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace ur

#define UR_LN_DISPATCH(d4, CALL)            \
    if ((d4) <= 32) { CALL(1) }             \
    else if ((d4) <= 64) { CALL(2) }        \
    else if ((d4) <= 128) { CALL(4) }       \
    else if ((d4) <= 256) { CALL(8) }       \
    else return UR_ERR_UNSUPPORTED;

extern "C" {

int ur_seq_prep_ln_fwd_f32(const float* table, const float* pos, const float* gamma, const float* beta, float eps,
                           const int32_t* item_seq, int64_t B, int L, int d, float* Y, float* mean, float* rstd,
                           const int32_t* tok_src, const int32_t* n_tok_dev, const void* shard_ptrs, int shard_world,
                           const int64_t* rng, float drop_p, int drop_site, void* stream) {
    if (shard_ptrs && shard_world < 1) return UR_ERR_BAD_ARG;
    if (drop_p < 0.f || drop_p >= 1.f) return UR_ERR_BAD_ARG;
    if (d <= 0 || (d & 3)) return UR_ERR_BAD_ARG;
    const int64_t rows = B * L;
    if (rows == 0) return UR_OK;
    const int d4 = d / 4;
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(MV)                                                                                                    \
    ur::seq_prep_ln_fwd_kernel<MV><<<ur::ln_grid(rows), 256, 0, st>>>(                                              \
        (const float4*)table, (const float4*)pos, (const float4*)gamma, (const float4*)beta, eps, item_seq, rows, L, d4, \
        (float4*)Y, mean, rstd, tok_src, n_tok_dev, (const long long*)shard_ptrs, shard_world, (const long long*)rng, drop_p, drop_site);
    UR_LN_DISPATCH(d4, CALL)
#undef CALL
    UR_RETURN_LAST_ERROR();
}

int ur_seq_prep_ln_bwd_f32(const float* table, const float* pos, const float* gamma, const int32_t* item_seq, int64_t B, int L,
                           int d, const float* mean, const float* rstd, const float* dY, float* dX, float* dgamma, float* dbeta,
                           float* dpos, const int32_t* tok_inv, const void* shard_ptrs, int shard_world, const int64_t* rng,
                           float drop_p, int drop_site, void* stream) {
    if (shard_ptrs && shard_world < 1) return UR_ERR_BAD_ARG;
    if (drop_p < 0.f || drop_p >= 1.f) return UR_ERR_BAD_ARG;
    if (d <= 0 || (d & 3)) return UR_ERR_BAD_ARG;
    if (B * L == 0) return UR_OK;
    const int d4 = d / 4;
    cudaStream_t st = (cudaStream_t)stream;
    int nbx = (int)((B + 63) / 64);          // 8 warps x 8 samples per CTA column
    if (nbx < 1) nbx = 1;
    if (nbx > 64) nbx = 64;
    dim3 grid(nbx, L);
#define CALL(MV)                                                                                               \
    ur::seq_prep_ln_bwd_kernel<MV><<<grid, 256, d * sizeof(float), st>>>(                                      \
        (const float4*)table, (const float4*)pos, (const float4*)gamma, item_seq, B, L, d4, mean, rstd,        \
        (const float4*)dY, (float4*)dX, dgamma, dbeta, dpos, tok_inv, (const long long*)shard_ptrs, shard_world,    \
        (const long long*)rng, drop_p, drop_site);
    UR_LN_DISPATCH(d4, CALL)
#undef CALL
    UR_RETURN_LAST_ERROR();
}

int ur_add_ln_fwd_f32(float* X, int64_t ldx, const float* R, int64_t ldr, const float* gamma, const float* beta, float eps,
                      int64_t rows, int d, float* Y, int64_t ldy, float* mean, float* rstd, const int32_t* rows_dev,
                      const int64_t* rng, float drop_p, int drop_site, const int32_t* row_pos, int64_t pos_mul, int64_t pos_add,
                      void* stream) {
    if (d <= 0 || (d & 3) || (ldx & 3) || (ldr & 3) || (ldy & 3)) return UR_ERR_BAD_ARG;
    if (drop_p < 0.f || drop_p >= 1.f) return UR_ERR_BAD_ARG;
    if (rows == 0) return UR_OK;
    const int d4 = d / 4;
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(MV)                                                                                                      \
    ur::add_ln_fwd_kernel<MV><<<ur::ln_grid(rows), 256, 0, st>>>(X, ldx, R, ldr, (const float4*)gamma, (const float4*)beta, \
                                                                 eps, rows, d4, Y, ldy, mean, rstd, rows_dev,               \
                                                                 (const long long*)rng, drop_p, drop_site, row_pos, pos_mul, pos_add);
    UR_LN_DISPATCH(d4, CALL)
#undef CALL
    UR_RETURN_LAST_ERROR();
}

int ur_add_ln_bwd_f32(const float* Z, int64_t ldz, const float* gamma, const float* mean, const float* rstd, const float* dY,
                      int64_t lddy, const float* dExtra, int64_t ldde, int64_t rows, int d, float* dZ, int64_t lddz,
                      float* dgamma, float* dbeta, float* dzsum, const int32_t* rows_dev, float* dZdrop, int64_t lddzd,
                      const int64_t* rng, float drop_p, int drop_site, const int32_t* row_pos, int64_t pos_mul, int64_t pos_add,
                      void* stream) {
    if (d <= 0 || (d & 3) || (ldz & 3) || (lddy & 3) || (ldde & 3) || (lddz & 3) || (lddzd & 3)) return UR_ERR_BAD_ARG;
    if (drop_p < 0.f || drop_p >= 1.f || (drop_p > 0.f && rng != nullptr && dZdrop == nullptr)) return UR_ERR_BAD_ARG;
    if (rows == 0) return UR_OK;
    const int d4 = d / 4;
    cudaStream_t st = (cudaStream_t)stream;
    int grid = ur::ln_grid(rows);
    if (grid > ur::kNumSMs * 2) grid = ur::kNumSMs * 2;   // fewer CTAs -> fewer column atomics
#define CALL(MV)                                                                                                    \
    ur::add_ln_bwd_kernel<MV><<<grid, 256, d * sizeof(float), st>>>(Z, ldz, (const float4*)gamma, mean, rstd, dY, lddy, \
                                                                    dExtra, ldde, rows, d4, dZ, lddz, dgamma, dbeta, dzsum, rows_dev, \
                                                                    dZdrop, lddzd, (const long long*)rng, drop_p, drop_site,    \
                                                                    row_pos, pos_mul, pos_add);
    UR_LN_DISPATCH(d4, CALL)
#undef CALL
    UR_RETURN_LAST_ERROR();
}

}  // extern "C"
