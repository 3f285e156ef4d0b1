// Token packing for the sequence towers: the encoder only has to process the positions whose result can reach the loss.
//
// Reference semantics (unirec/model/sequential/sasrec.py:40-57, modules.py:289-311): the additive mask depends on KEYS only
// (0 for a real item that is not in the future of the query, -10000 otherwise).  Hence, per sequence:
//   * a real position is live (query, key and value);
//   * position L-1 is always live as a QUERY (its output is the user embedding, sasrec.py:74-75), padded or not;
//   * every other padded position is dead: as a key it is masked wherever a real key exists, and its own output is never read;
//   * a sequence without any real item: all L keys carry the same -10000, which cancels in the softmax, so every position is
//     live and attends to all L positions.
// `ur_pack_tokens` lists the live positions sample by sample in their original order:
//   tok_src[t]  = b*L + l of packed token t          (t < n_tok)
//   tok_inv[b*L + l] = t, or -1 for a dead position
//   offs[b]     = first packed token of sample b     (offs[B] = n_tok)
//   last_tok[b] = packed index of position L-1
// keep_all = 1 gives the identity map (the unpacked computation, used for A/B tests and by towers that need every position).
#include "common.cuh"

namespace ur {

__global__ void __launch_bounds__(1024) pack_tokens_kernel(const int32_t* __restrict__ seq, int B, int L, int keep_all,
                                                           int32_t* __restrict__ offs, int32_t* __restrict__ tok_src,
                                                           int32_t* __restrict__ tok_inv, int32_t* __restrict__ last_tok,
                                                           int32_t* __restrict__ n_tok) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < B; b0 += blockDim.x) {
        const int b = b0 + tid;
        int cnt = 0;
        bool all = keep_all != 0;
        if (b < B) {
            const int32_t* s = seq + (int64_t)b * L;
            int real = 0;
            for (int l = 0; l < L; ++l) real += s[l] > 0;
            if (real == 0) all = true;                                   // empty history: every position is live
            cnt = all ? L : real + (s[L - 1] > 0 ? 0 : 1);               // position L-1 is always kept
        }
        // block-wide exclusive scan of cnt
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = lane < (int)(blockDim.x >> 5) ? warp_tot[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += v;
            }
            warp_tot[lane] = w;                                          // inclusive totals of the warps
        }
        __syncthreads();
        const int base = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + incl - cnt;
        if (b < B) {
            offs[b] = base;
            const int32_t* s = seq + (int64_t)b * L;
            int t = base;
            for (int l = 0; l < L; ++l) {
                const bool keep = all || s[l] > 0 || l == L - 1;
                tok_inv[(int64_t)b * L + l] = keep ? t : -1;
                if (keep) tok_src[t++] = b * L + l;
            }
            last_tok[b] = t - 1;
        }
        __syncthreads();
        if (tid == 0) carry += warp_tot[(blockDim.x >> 5) - 1];
        __syncthreads();
    }
    if (tid == 0) { offs[B] = carry; *n_tok = carry; }
}

// zero rows [*n_dev, min(rows_cap, roundup(*n_dev, 32))) of a [rows_cap, width] matrix (token-reduction GEMMs read whole 32-row k-blocks)
__global__ void __launch_bounds__(256) zero_tail_rows_kernel(float* __restrict__ X, int64_t ld, int width, const int32_t* __restrict__ n_dev,
                                                             int rows_cap) {
    const int n = *n_dev;
    const int end = min(rows_cap, (n + 31) & ~31);
    for (int i = threadIdx.x; i < (end - n) * width; i += blockDim.x) {
        const int r = n + i / width, c = i - (i / width) * width;
        X[(int64_t)r * ld + c] = 0.f;
    }
}

}  // namespace ur

extern "C" {

int ur_pack_tokens(const int32_t* item_seq, int64_t B, int L, int keep_all, int32_t* offs, int32_t* tok_src, int32_t* tok_inv,
                   int32_t* last_tok, int32_t* n_tok, void* stream) {
    if (B <= 0 || L <= 0 || B * L >= ((int64_t)1 << 31)) return UR_ERR_BAD_ARG;
    ur::pack_tokens_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(item_seq, (int)B, L, keep_all, offs, tok_src, tok_inv, last_tok, n_tok);
    UR_RETURN_LAST_ERROR();
}

int ur_zero_tail_rows_f32(float* X, int64_t ld, int width, const int32_t* n_dev, int64_t rows_cap, void* stream) {
    ur::zero_tail_rows_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(X, ld, width, n_dev, (int)rows_cap);
    UR_RETURN_LAST_ERROR();
}

}  // extern "C"
