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

}  // extern "C"
