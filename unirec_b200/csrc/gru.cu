// K9: GRU cell pointwise kernels.  The matrix products (x W_ih^T for all steps at once, h W_hh^T per step) run through
// ur_gemm_*; these kernels fuse the gate nonlinearities and the state update, forward and backward.
// Reference: nn.GRU(batch_first, 1 layer) as used by unirec/model/sequential/gru.py:17-30.  Gate order r, z, n:
//   r = sigmoid(gi_r + gh_r); z = sigmoid(gi_z + gh_z); n = tanh(gi_n + r * gh_n); h' = (1 - z) * n + z * h
// (gi = x W_ih^T + b_ih, gh = h W_hh^T + b_hh).
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
template <int RB>
__global__ void gru_seq_fwd_kernel(const float* __restrict__ gi, int64_t L, const float* __restrict__ whh_t, const float* __restrict__ b_hh,
                                   float* __restrict__ hs, float* __restrict__ save, int64_t B, int H) {
    extern __shared__ __align__(16) float hsm[];          // [2][RB][H]
    const int c = threadIdx.x;
    const int64_t b0 = (int64_t)blockIdx.x * RB;
    const int nb = (int)min((int64_t)RB, B - b0);
    const float br = b_hh[c], bz = b_hh[H + c], bn = b_hh[2 * H + c];
    for (int r = 0; r < RB; ++r) hsm[r * H + c] = r < nb ? hs[(b0 + r) * H + c] : 0.f;       // hs[0] = initial state
    __syncthreads();
    int cur = 0;
    for (int64_t t = 0; t < L; ++t) {
        const float* hp = hsm + cur * RB * H;
        float ar[RB], az[RB], an[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) { ar[r] = 0.f; az[r] = 0.f; an[r] = 0.f; }
        for (int k = 0; k < H; k += 4) {
            float wr[4], wz[4], wn[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float* w = whh_t + (int64_t)(k + q) * 3 * H;
                wr[q] = __ldg(w + c); wz[q] = __ldg(w + H + c); wn[q] = __ldg(w + 2 * H + c);
            }
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const float4 h4 = *reinterpret_cast<const float4*>(hp + r * H + k);       // broadcast read
                ar[r] = fmaf(h4.x, wr[0], ar[r]); ar[r] = fmaf(h4.y, wr[1], ar[r]); ar[r] = fmaf(h4.z, wr[2], ar[r]); ar[r] = fmaf(h4.w, wr[3], ar[r]);
                az[r] = fmaf(h4.x, wz[0], az[r]); az[r] = fmaf(h4.y, wz[1], az[r]); az[r] = fmaf(h4.z, wz[2], az[r]); az[r] = fmaf(h4.w, wz[3], az[r]);
                an[r] = fmaf(h4.x, wn[0], an[r]); an[r] = fmaf(h4.y, wn[1], an[r]); an[r] = fmaf(h4.z, wn[2], an[r]); an[r] = fmaf(h4.w, wn[3], an[r]);
            }
        }
        float* hn_s = hsm + (cur ^ 1) * RB * H;
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            if (r < nb) {
                const int64_t b = b0 + r;
                const float* g = gi + (b * L + t) * 3 * H;
                const float rr = sigmoidf_(g[c] + (ar[r] + br));
                const float zz = sigmoidf_(g[H + c] + (az[r] + bz));
                const float hn = an[r] + bn;
                const float nn = tanhf(g[2 * H + c] + rr * hn);
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

template <int RB>
__global__ void gru_seq_bwd_kernel(const float* __restrict__ dh_last, const float* __restrict__ save, const float* __restrict__ hs,
                                   const float* __restrict__ whh, float* __restrict__ dgi, int64_t L, float* __restrict__ dgh_all,
                                   int64_t B, int H) {
    extern __shared__ __align__(16) float sm[];           // dgh tile [RB][3H]
    const int c = threadIdx.x;
    const int64_t b0 = (int64_t)blockIdx.x * RB;
    const int nb = (int)min((int64_t)RB, B - b0);
    float dh[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) dh[r] = r < nb ? dh_last[(b0 + r) * H + c] : 0.f;
    for (int64_t t = L - 1; t >= 0; --t) {
        float dhz[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            float dr = 0.f, dz = 0.f, dnr = 0.f;
            dhz[r] = 0.f;
            if (r < nb) {
                const int64_t b = b0 + r;
                const float* sv = save + (t * B + b) * 4 * H;
                const float rr = sv[c], zz = sv[H + c], nn = sv[2 * H + c], hn = sv[3 * H + c];
                const float g = dh[r];
                const float dn_pre = g * (1.f - zz) * (1.f - nn * nn);
                dz = g * (hs[(t * B + b) * H + c] - nn) * zz * (1.f - zz);
                dr = dn_pre * hn * rr * (1.f - rr);
                dnr = dn_pre * rr;
                float* a = dgi + (b * L + t) * 3 * H;
                a[c] = dr; a[H + c] = dz; a[2 * H + c] = dn_pre;
                float* q = dgh_all + (t * B + b) * 3 * H;
                q[c] = dr; q[H + c] = dz; q[2 * H + c] = dnr;
                dhz[r] = g * zz;
            }
            sm[r * 3 * H + c] = dr; sm[r * 3 * H + H + c] = dz; sm[r * 3 * H + 2 * H + c] = dnr;
        }
        __syncthreads();
        // dh_{t-1}[r][c] = dh_t * z + sum_j dgh[r][j] * W_hh[j][c]   (W_hh [3H, H] row-major: coalesced over c)
        float acc[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) acc[r] = dhz[r];
        for (int j = 0; j < 3 * H; j += 4) {
            float w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) w[q] = __ldg(whh + (int64_t)(j + q) * H + c);
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const float4 g4 = *reinterpret_cast<const float4*>(sm + r * 3 * H + j);
                acc[r] = fmaf(g4.x, w[0], acc[r]); acc[r] = fmaf(g4.y, w[1], acc[r]);
                acc[r] = fmaf(g4.z, w[2], acc[r]); acc[r] = fmaf(g4.w, w[3], acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) dh[r] = acc[r];
        __syncthreads();
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
    if (H <= 512) {
        constexpr int RB = 16;
        const size_t sm = (size_t)2 * RB * H * sizeof(float);
        if (sm > 48 * 1024) cudaFuncSetAttribute(ur::gru_seq_fwd_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        ur::gru_seq_fwd_kernel<RB><<<(unsigned)((B + RB - 1) / RB), H, sm, st>>>(gi, L, whh_t, b_hh, hs, save, B, H);
    } else {
        constexpr int RB = 8;
        const size_t sm = (size_t)2 * RB * H * sizeof(float);
        if (sm > 48 * 1024) cudaFuncSetAttribute(ur::gru_seq_fwd_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        ur::gru_seq_fwd_kernel<RB><<<(unsigned)((B + RB - 1) / RB), H, sm, st>>>(gi, L, whh_t, b_hh, hs, save, B, H);
    }
    UR_RETURN_LAST_ERROR();
}

int ur_gru_seq_bwd_f32(const float* dh_last, const float* save, const float* hs, const float* whh, float* dgi, float* dgh_all, int64_t B,
                       int64_t L, int H, void* stream) {
    if (H <= 0 || (H & 31) || H > 768 || L <= 0) return UR_ERR_UNSUPPORTED;      // (register file: 768 threads x <= 80 registers)
    if (B == 0) return UR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (H <= 512) {
        constexpr int RB = 16;
        const size_t sm = (size_t)RB * 3 * H * sizeof(float);
        if (sm > 48 * 1024) cudaFuncSetAttribute(ur::gru_seq_bwd_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        ur::gru_seq_bwd_kernel<RB><<<(unsigned)((B + RB - 1) / RB), H, sm, st>>>(dh_last, save, hs, whh, dgi, L, dgh_all, B, H);
    } else {
        constexpr int RB = 8;
        const size_t sm = (size_t)RB * 3 * H * sizeof(float);
        if (sm > 48 * 1024) cudaFuncSetAttribute(ur::gru_seq_bwd_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        ur::gru_seq_bwd_kernel<RB><<<(unsigned)((B + RB - 1) / RB), H, sm, st>>>(dh_last, save, hs, whh, dgi, L, dgh_all, B, H);
    }
    UR_RETURN_LAST_ERROR();
}

}  // extern "C"
