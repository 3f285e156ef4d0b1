// K9: GRU cell pointwise kernels.  The matrix products (x W_ih^T for all steps at once, h W_hh^T per step) run through
// ur_gemm_*; these kernels fuse the gate nonlinearities and the state update, forward and backward.
// Reference: nn.GRU(batch_first, 1 layer) as used by unirec/model/sequential/gru.py:17-30.  Gate order r, z, n:
//   r = sigmoid(gi_r + gh_r); z = sigmoid(gi_z + gh_z); n = tanh(gi_n + r * gh_n); h' = (1 - z) * n + z * h
// (gi = x W_ih^T + b_ih, gh = h W_hh^T + b_hh).
#include <stdlib.h>
#include "common.cuh"

namespace ur {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// gi: rows at stride ld_gi (batch-major [B, L, 3H] viewed at step t); gh [B,3H]; h_prev/h_out [B,H]; save [B,4H] = r|z|n|gh_n
__global__ void __launch_bounds__(256) gru_gate_fwd_kernel(const float* __restrict__ gi, int64_t ld_gi, const float* __restrict__ gh,
                                                           const float* __restrict__ h_prev, float* __restrict__ h_out,
                                                           float* __restrict__ save, int64_t B, int H) {
    const int64_t total = B * H;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / H;
        const int c = (int)(i - b * H);
        const float* gib = gi + b * ld_gi;
        const float* ghb = gh + b * 3 * H;
        const float r = sigmoidf_(gib[c] + ghb[c]);
        const float z = sigmoidf_(gib[H + c] + ghb[H + c]);
        const float hn = ghb[2 * H + c];
        const float n = tanhf(gib[2 * H + c] + r * hn);
        const float hp = h_prev[i];
        h_out[i] = (1.f - z) * n + z * hp;
        float* s = save + b * 4 * H;
        s[c] = r; s[H + c] = z; s[2 * H + c] = n; s[3 * H + c] = hn;
    }
}

// dh: gradient wrt h_t [B,H] (in); writes dgi (strided, [B,3H] at step t), dgh [B,3H], and dh_prev = dh * z
// (the recurrent part dgh W_hh is accumulated on top by a GEMM).
__global__ void __launch_bounds__(256) gru_gate_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ save,
                                                           const float* __restrict__ h_prev, float* __restrict__ dgi, int64_t ld_dgi,
                                                           float* __restrict__ dgh, float* __restrict__ dh_prev, int64_t B, int H) {
    const int64_t total = B * H;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / H;
        const int c = (int)(i - b * H);
        const float* s = save + b * 4 * H;
        const float r = s[c], z = s[H + c], n = s[2 * H + c], hn = s[3 * H + c];
        const float g = dh[i];
        const float dn_pre = g * (1.f - z) * (1.f - n * n);
        const float dz_pre = g * (h_prev[i] - n) * z * (1.f - z);
        const float dr_pre = dn_pre * hn * r * (1.f - r);
        float* a = dgi + b * ld_dgi;
        a[c] = dr_pre; a[H + c] = dz_pre; a[2 * H + c] = dn_pre;
        float* q = dgh + b * 3 * H;
        q[c] = dr_pre; q[H + c] = dz_pre; q[2 * H + c] = dn_pre * r;
        dh_prev[i] = g * z;
    }
}


// ---------------------------------------------------------------------------------------------------------------------------
// Persistent recurrence: ONE launch runs all L time steps.  Batch rows are independent along the recurrence, so a
// CTA owns RB rows for the whole sequence: its hidden-state tile lives in shared memory (double buffered), W_hh (3h x h fp32, too
// large for shared memory at h >= 128) streams from L2 every step -- read coalesced through a transposed copy W_hh^T [h, 3h] made
// once per call.  blockDim = h: thread c owns hidden unit c, i.e. the three gate columns (c, h + c, 2h + c) of gh for its RB
// rows, and goes straight from the dot products to h_t (no gh round trip, no per-step launches).
//   forward : hs[t+1] = GRUCell(gi[:, t], hs[t]);  save[t] = (r, z, n, W_hn h + b_hn)                        (gru.py:30)
//   backward: dgi[:, t], dgh[t] from dh_t and the saved gates; dh_{t-1} = dh_t * z + dgh[t] W_hh
// The weight / bias gradients are token reductions over all steps and stay on the GEMM / column-sum kernels.
__device__ __forceinline__ uint32_t gsm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "GRU_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra GRU_DONE;\n\t"
        "bra GRU_WAIT;\n\t"
        "GRU_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// one elected thread: arm the stage barrier and start the bulk copy of a contiguous weight slab into shared memory
__device__ __forceinline__ void slab_issue(float* dst, const float* src, uint32_t bytes, uint64_t* bar) {
    const uint32_t b = gsm_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(gsm_u32(dst)), "l"(src), "r"(bytes), "r"(b) : "memory");
}

constexpr int GRU_NS = 4;       // weight-slab ring depth
// KT (template): forward = k rows of W_hh^T per slab (KT x 3H floats); backward = 3 * KT rows of W_hh (same bytes).  8 for H <= 256
// (24 KB slabs), 4 above (H = 768: 36 KB slabs, ring + tiles = 196 / 221 KB of shared memory).

// RT rows per thread, NG thread groups per CTA (blockDim = NG * H: thread = (group, hidden unit)); the CTA owns RB = RT * NG rows.
template <int RT, int NG, int GRU_KT, int MAXT>
__global__ void __launch_bounds__(MAXT) gru_seq_fwd_kernel(const float* __restrict__ gi, int64_t L, const float* __restrict__ whh_t, const float* __restrict__ b_hh,
                                   float* __restrict__ hs, float* __restrict__ save, int64_t B, int H) {
    extern __shared__ __align__(128) float dsm[];
    constexpr int RB = RT * NG;
    float* wring = dsm;                                        // [NS][KT][3H]  weight slabs (cp.async.bulk)
    float* hsm = wring + (size_t)GRU_NS * GRU_KT * 3 * H;      // [2][RB][H]    hidden-state tile, double buffered
    uint64_t* bars = reinterpret_cast<uint64_t*>(hsm + 2 * RB * H);
    const int tid = threadIdx.x;
    const int c = tid % H, r0 = (tid / H) * RT;                // hidden unit, first row of this thread's group
    const int64_t b0 = (int64_t)blockIdx.x * RB;
    const int nb = (int)min((int64_t)RB, B - b0);
    const float br = b_hh[c], bz = b_hh[H + c], bn = b_hh[2 * H + c];
    const int slabs = H / GRU_KT;                              // per time step
    const int64_t total = (int64_t)slabs * L;
    const uint32_t slab_bytes = (uint32_t)(GRU_KT * 3 * H * sizeof(float));
    if (tid == 0) {
        for (int s = 0; s < GRU_NS; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(gsm_u32(bars + s)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int r = r0; r < r0 + RT; ++r) hsm[r * H + c] = r < nb ? hs[(b0 + r) * H + c] : 0.f;       // hs[0] = initial state
    __syncthreads();
    if (tid == 0)
        for (int g = 0; g < GRU_NS - 1 && g < total; ++g)
            slab_issue(wring + (size_t)g * GRU_KT * 3 * H, whh_t + (size_t)(g % slabs) * GRU_KT * 3 * H, slab_bytes, bars + g);
    // gi rows of a step are first touched in its gate phase: prefetch them one step ahead (one lane per 128-byte line)
    auto prefetch_gi = [&](int64_t t) {
        if ((c & 31) == 0 && t < L) {
#pragma unroll
            for (int r = 0; r < RT; ++r)
                if (r0 + r < nb) {
                    const float* gp = gi + ((b0 + r0 + r) * L + t) * 3 * H + c;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(gp));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(gp + H));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(gp + 2 * H));
                }
        }
    };
    prefetch_gi(0);
    int64_t g = 0;
    int cur = 0;
    for (int64_t t = 0; t < L; ++t) {
        const float* hp = hsm + cur * RB * H + r0 * H;
        // this step's gate inputs: issued now, consumed after the k loop (their latency hides behind it)
        float gir[RT], giz[RT], gin[RT];
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            gir[r] = giz[r] = gin[r] = 0.f;
            if (r0 + r < nb) {
                const float* gp = gi + ((b0 + r0 + r) * L + t) * 3 * H;
                gir[r] = __ldg(gp + c); giz[r] = __ldg(gp + H + c); gin[r] = __ldg(gp + 2 * H + c);
            }
        }
        prefetch_gi(t + 1);
        float ar[RT], az[RT], an[RT];
#pragma unroll
        for (int r = 0; r < RT; ++r) { ar[r] = 0.f; az[r] = 0.f; an[r] = 0.f; }
        for (int sb = 0; sb < slabs; ++sb, ++g) {
            const int st = (int)(g % GRU_NS);
            if (tid == 0 && g + GRU_NS - 1 < total) {          // refill the stage freed by the barrier at the end of slab g-1
                const int64_t gn = g + GRU_NS - 1;
                slab_issue(wring + (size_t)(gn % GRU_NS) * GRU_KT * 3 * H, whh_t + (size_t)(gn % slabs) * GRU_KT * 3 * H, slab_bytes,
                           bars + (gn % GRU_NS));
            }
            gbar_wait(gsm_u32(bars + st), (uint32_t)((g / GRU_NS) & 1));
            const float* w = wring + (size_t)st * GRU_KT * 3 * H;
            const int k0 = sb * GRU_KT;
#pragma unroll
            for (int kq = 0; kq < GRU_KT; kq += 4) {
                float wr[4], wz[4], wn[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float* wk = w + (kq + q) * 3 * H;
                    wr[q] = wk[c]; wz[q] = wk[H + c]; wn[q] = wk[2 * H + c];
                }
#pragma unroll
                for (int r = 0; r < RT; ++r) {
                    const float4 h4 = *reinterpret_cast<const float4*>(hp + r * H + k0 + kq);       // broadcast read
                    ar[r] = fmaf(h4.x, wr[0], ar[r]); ar[r] = fmaf(h4.y, wr[1], ar[r]); ar[r] = fmaf(h4.z, wr[2], ar[r]); ar[r] = fmaf(h4.w, wr[3], ar[r]);
                    az[r] = fmaf(h4.x, wz[0], az[r]); az[r] = fmaf(h4.y, wz[1], az[r]); az[r] = fmaf(h4.z, wz[2], az[r]); az[r] = fmaf(h4.w, wz[3], az[r]);
                    an[r] = fmaf(h4.x, wn[0], an[r]); an[r] = fmaf(h4.y, wn[1], an[r]); an[r] = fmaf(h4.z, wn[2], an[r]); an[r] = fmaf(h4.w, wn[3], an[r]);
                }
            }
            __syncthreads();                                   // every thread is done with stage st
        }
        float* hn_s = hsm + (cur ^ 1) * RB * H + r0 * H;
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            if (r0 + r < nb) {
                const int64_t b = b0 + r0 + r;
                const float rr = sigmoidf_(gir[r] + (ar[r] + br));
                const float zz = sigmoidf_(giz[r] + (az[r] + bz));
                const float hn = an[r] + bn;
                const float nn = tanhf(gin[r] + rr * hn);
                const float hprev = hp[r * H + c];
                const float hnew = (1.f - zz) * nn + zz * hprev;
                hn_s[r * H + c] = hnew;
                hs[((t + 1) * B + b) * H + c] = hnew;
                float* sv = save + (t * B + b) * 4 * H;
                sv[c] = rr; sv[H + c] = zz; sv[2 * H + c] = nn; sv[3 * H + c] = hn;
            }
        }
        __syncthreads();
        cur ^= 1;
    }
}

template <int RT, int NG, int GRU_KT, int MAXT>
__global__ void __launch_bounds__(MAXT) gru_seq_bwd_kernel(const float* __restrict__ dh_last, const float* __restrict__ save, const float* __restrict__ hs,
                                   const float* __restrict__ whh, float* __restrict__ dgi, int64_t L, float* __restrict__ dgh_all,
                                   int64_t B, int H) {
    extern __shared__ __align__(128) float dsm[];
    constexpr int RB = RT * NG;
    constexpr int JT = 3 * GRU_KT;                             // rows of W_hh [3H, H] per slab
    float* wring = dsm;                                        // [NS][JT][H]
    float* sm = wring + (size_t)GRU_NS * JT * H;               // dgh tile [RB][3H]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + RB * 3 * H);
    const int tid = threadIdx.x;
    const int c = tid % H, r0 = (tid / H) * RT;
    const int64_t b0 = (int64_t)blockIdx.x * RB;
    const int nb = (int)min((int64_t)RB, B - b0);
    const int slabs = 3 * H / JT;
    const int64_t total = (int64_t)slabs * L;
    const uint32_t slab_bytes = (uint32_t)(JT * H * sizeof(float));
    if (tid == 0) {
        for (int s = 0; s < GRU_NS; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(gsm_u32(bars + s)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int g = 0; g < GRU_NS - 1 && g < total; ++g)
            slab_issue(wring + (size_t)g * JT * H, whh + (size_t)(g % slabs) * JT * H, slab_bytes, bars + g);
    float dh[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) dh[r] = r0 + r < nb ? dh_last[(b0 + r0 + r) * H + c] : 0.f;
    // saved gates / previous hidden state of a step: loaded one step ahead (their HBM latency hides behind the previous step's k loop)
    float nr[RT], nz[RT], nn_[RT], nhn[RT], nhp[RT];
    auto load_saved = [&](int64_t t) {
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            nr[r] = nz[r] = nn_[r] = nhn[r] = nhp[r] = 0.f;
            if (t >= 0 && r0 + r < nb) {
                const int64_t b = b0 + r0 + r;
                const float* sv = save + (t * B + b) * 4 * H;
                nr[r] = __ldg(sv + c); nz[r] = __ldg(sv + H + c); nn_[r] = __ldg(sv + 2 * H + c); nhn[r] = __ldg(sv + 3 * H + c);
                nhp[r] = __ldg(hs + (t * B + b) * H + c);
            }
        }
    };
    load_saved(L - 1);
    int64_t g = 0;
    for (int64_t t = L - 1; t >= 0; --t) {
        float acc[RT];
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            float dr = 0.f, dz = 0.f, dnr = 0.f;
            acc[r] = 0.f;
            if (r0 + r < nb) {
                const int64_t b = b0 + r0 + r;
                const float rr = nr[r], zz = nz[r], nn = nn_[r], hn = nhn[r];
                const float gq = dh[r];
                const float dn_pre = gq * (1.f - zz) * (1.f - nn * nn);
                dz = gq * (nhp[r] - nn) * zz * (1.f - zz);
                dr = dn_pre * hn * rr * (1.f - rr);
                dnr = dn_pre * rr;
                float* a = dgi + (b * L + t) * 3 * H;
                a[c] = dr; a[H + c] = dz; a[2 * H + c] = dn_pre;
                float* q = dgh_all + (t * B + b) * 3 * H;
                q[c] = dr; q[H + c] = dz; q[2 * H + c] = dnr;
                acc[r] = gq * zz;
            }
            sm[(r0 + r) * 3 * H + c] = dr; sm[(r0 + r) * 3 * H + H + c] = dz; sm[(r0 + r) * 3 * H + 2 * H + c] = dnr;
        }
        load_saved(t - 1);
        __syncthreads();
        // dh_{t-1}[r][c] = dh_t * z + sum_j dgh[r][j] * W_hh[j][c]   (W_hh [3H, H] row-major: slabs of JT rows)
        const float* smr = sm + r0 * 3 * H;
        for (int sb = 0; sb < slabs; ++sb, ++g) {
            const int st = (int)(g % GRU_NS);
            if (tid == 0 && g + GRU_NS - 1 < total) {
                const int64_t gn = g + GRU_NS - 1;
                slab_issue(wring + (size_t)(gn % GRU_NS) * JT * H, whh + (size_t)(gn % slabs) * JT * H, slab_bytes, bars + (gn % GRU_NS));
            }
            gbar_wait(gsm_u32(bars + st), (uint32_t)((g / GRU_NS) & 1));
            const float* w = wring + (size_t)st * JT * H;
            const int j0 = sb * JT;
#pragma unroll
            for (int jq = 0; jq < JT; jq += 4) {
                float wv[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) wv[q] = w[(jq + q) * H + c];
#pragma unroll
                for (int r = 0; r < RT; ++r) {
                    const float4 g4 = *reinterpret_cast<const float4*>(smr + r * 3 * H + j0 + jq);
                    acc[r] = fmaf(g4.x, wv[0], acc[r]); acc[r] = fmaf(g4.y, wv[1], acc[r]);
                    acc[r] = fmaf(g4.z, wv[2], acc[r]); acc[r] = fmaf(g4.w, wv[3], acc[r]);
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int r = 0; r < RT; ++r) dh[r] = acc[r];
    }
}

}  // namespace ur

extern "C" {

int ur_gru_gate_fwd_f32(const float* gi, int64_t ld_gi, const float* gh, const float* h_prev, float* h_out, float* save, int64_t B,
                        int H, void* stream) {
    if (B == 0) return UR_OK;
    int64_t blocks = (B * H + 255) / 256;
    const int64_t cap = (int64_t)ur::kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    ur::gru_gate_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(gi, ld_gi, gh, h_prev, h_out, save, B, H);
    UR_RETURN_LAST_ERROR();
}

int ur_gru_gate_bwd_f32(const float* dh, const float* save, const float* h_prev, float* dgi, int64_t ld_dgi, float* dgh,
                        float* dh_prev, int64_t B, int H, void* stream) {
    if (B == 0) return UR_OK;
    int64_t blocks = (B * H + 255) / 256;
    const int64_t cap = (int64_t)ur::kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    ur::gru_gate_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dh, save, h_prev, dgi, ld_dgi, dgh, dh_prev, B, H);
    UR_RETURN_LAST_ERROR();
}

// Whole-sequence recurrence in one launch each way.  gi / dgi: [B, L, 3H] (batch-major, as produced by the input GEMM); hs: [L+1, B, H]
// with hs[0] = initial state (zeros); save: [L, B, 4H]; whh_t: W_hh^T [H, 3H] (ur_transpose_f32 of weight_hh_l0); H % 32 == 0, H <= 768.
int ur_gru_seq_fwd_f32(const float* gi, const float* whh_t, const float* b_hh, float* hs, float* save, int64_t B, int64_t L, int H,
                       void* stream) {
    if (H <= 0 || (H & 31) || H > 768 || L <= 0) return UR_ERR_UNSUPPORTED;      // (register file: 768 threads x <= 80 registers)
    if (B == 0) return UR_OK;
    cudaStream_t st = (cudaStream_t)stream;
#define UR_GRU_FWD(RT, NG, KT, MAXT)                                                                                            \
    do {                                                                                                                    \
        constexpr int RB = RT * NG;                                                                                         \
        const size_t sm = ((size_t)ur::GRU_NS * KT * 3 * H + (size_t)2 * RB * H) * sizeof(float) + ur::GRU_NS * 8 + 16;      \
        cudaFuncSetAttribute(ur::gru_seq_fwd_kernel<RT, NG, KT, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);     \
        ur::gru_seq_fwd_kernel<RT, NG, KT, MAXT><<<(unsigned)((B + RB - 1) / RB), NG * H, sm, st>>>(gi, L, whh_t, b_hh, hs, save, B, H); \
    } while (0)
    static const int variant = getenv("UR_GRU_VARIANT") ? atoi(getenv("UR_GRU_VARIANT")) : 0;
    if (H <= 256 && variant == 1) UR_GRU_FWD(8, 2, 8, 512);      // 16 rows per CTA, two thread groups of 8 rows (16 warps at H = 256)
    else if (H <= 256) UR_GRU_FWD(16, 1, 8, 256);               // 16 rows per thread: fewest shared-memory reads per FMA
    else if (H <= 512) UR_GRU_FWD(16, 1, 4, 512);
    else UR_GRU_FWD(8, 1, 4, 768);
#undef UR_GRU_FWD
    UR_RETURN_LAST_ERROR();
}

int ur_gru_seq_bwd_f32(const float* dh_last, const float* save, const float* hs, const float* whh, float* dgi, float* dgh_all, int64_t B,
                       int64_t L, int H, void* stream) {
    if (H <= 0 || (H & 31) || H > 768 || L <= 0) return UR_ERR_UNSUPPORTED;
    if (B == 0) return UR_OK;
    cudaStream_t st = (cudaStream_t)stream;
#define UR_GRU_BWD(RT, NG, KT, MAXT)                                                                                            \
    do {                                                                                                                    \
        constexpr int RB = RT * NG;                                                                                         \
        const size_t sm = ((size_t)ur::GRU_NS * KT * 3 * H + (size_t)RB * 3 * H) * sizeof(float) + ur::GRU_NS * 8 + 16;      \
        cudaFuncSetAttribute(ur::gru_seq_bwd_kernel<RT, NG, KT, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);     \
        ur::gru_seq_bwd_kernel<RT, NG, KT, MAXT><<<(unsigned)((B + RB - 1) / RB), NG * H, sm, st>>>(dh_last, save, hs, whh, dgi, L, dgh_all, B, H); \
    } while (0)
    static const int variant = getenv("UR_GRU_VARIANT") ? atoi(getenv("UR_GRU_VARIANT")) : 0;
    if (H <= 256 && variant == 1) UR_GRU_BWD(8, 2, 8, 512);
    else if (H <= 256) UR_GRU_BWD(16, 1, 8, 256);
    else if (H <= 512) UR_GRU_BWD(16, 1, 4, 512);
    else UR_GRU_BWD(8, 1, 4, 768);
#undef UR_GRU_BWD
    UR_RETURN_LAST_ERROR();
}

}  // extern "C"
