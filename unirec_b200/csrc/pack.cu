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

// 1) one warp per sample: count the live positions (coalesced reads, ballot)
__global__ void __launch_bounds__(256) pack_count_kernel(const int32_t* __restrict__ seq, int B, int L, int keep_all,
                                                         int32_t* __restrict__ cnt) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const int32_t* s = seq + (int64_t)b * L;
    int real = 0, last_real = 0;
    for (int l0 = 0; l0 < L; l0 += 32) {
        const int l = l0 + lane;
        const bool r = l < L && s[l] > 0;
        real += __popc(__ballot_sync(0xffffffffu, r));
        if (l == L - 1) last_real = r;
    }
    last_real = __shfl_sync(0xffffffffu, last_real, (L - 1) & 31);
    if (lane == 0) {
        // bit 30 flags "keep every position" (identity map, or a sequence without any real item)
        const bool all = keep_all || real == 0;
        cnt[b] = all ? (L | (1 << 30)) : real + (last_real ? 0 : 1);          // position L-1 is always kept
    }
}

// 2) single CTA: exclusive scan of the B counts
__global__ void __launch_bounds__(1024) pack_scan_kernel(const int32_t* __restrict__ cnt, int B, int32_t* __restrict__ offs,
                                                         int32_t* __restrict__ n_tok) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < B; b0 += blockDim.x) {
        const int b = b0 + tid;
        const int c = b < B ? (cnt[b] & ~(1 << 30)) : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += v;
            }
            warp_tot[lane] = w;                                          // inclusive totals of the warps
        }
        __syncthreads();
        if (b < B) offs[b] = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + incl - c;
        __syncthreads();
        if (tid == 0) carry += warp_tot[31];
        __syncthreads();
    }
    if (tid == 0) { offs[B] = carry; *n_tok = carry; }
}

// 3) one warp per sample: write the maps
__global__ void __launch_bounds__(256) pack_fill_kernel(const int32_t* __restrict__ seq, int B, int L, const int32_t* __restrict__ cnt,
                                                        const int32_t* __restrict__ offs, int32_t* __restrict__ tok_src,
                                                        int32_t* __restrict__ tok_inv, int32_t* __restrict__ last_tok) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const int32_t* s = seq + (int64_t)b * L;
    const bool all = (cnt[b] >> 30) & 1;
    int t = offs[b];
    for (int l0 = 0; l0 < L; l0 += 32) {
        const int l = l0 + lane;
        const bool keep = l < L && (all || s[l] > 0 || l == L - 1);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (l < L) {
            const int mine = t + __popc(m & ((1u << lane) - 1));
            tok_inv[(int64_t)b * L + l] = keep ? mine : -1;
            if (keep) tok_src[mine] = b * L + l;
        }
        t += __popc(m);
    }
    if (lane == 0) last_tok[b] = t - 1;
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
    if (B <= 0 || L <= 0 || L >= (1 << 30) || B * L >= ((int64_t)1 << 31)) return UR_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)((B + 7) / 8);
    int32_t* cnt = last_tok;          // last_tok [B] doubles as the per-sample count between the three launches
    ur::pack_count_kernel<<<blocks, 256, 0, st>>>(item_seq, (int)B, L, keep_all, cnt);
    ur::pack_scan_kernel<<<1, 1024, 0, st>>>(cnt, (int)B, offs, n_tok);
    ur::pack_fill_kernel<<<blocks, 256, 0, st>>>(item_seq, (int)B, L, cnt, offs, tok_src, tok_inv, last_tok);
    UR_RETURN_LAST_ERROR();
}

int ur_zero_tail_rows_f32(float* X, int64_t ld, int width, const int32_t* n_dev, int64_t rows_cap, void* stream) {
    ur::zero_tail_rows_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(X, ld, width, n_dev, (int)rows_cap);
    UR_RETURN_LAST_ERROR();
}

}  // extern "C"
