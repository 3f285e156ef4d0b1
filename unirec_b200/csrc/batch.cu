// a15 / f1: device-side batch builder.  One kernel turns (user_id, positive item) pairs plus the CSR user history into the
// reference's batch contract -- item_id [B,1+K] (positive first), label [B,1+K], item_seq [B,L] left-padded, item_seq_len [B] --
// replacing the per-sample Python transforms that run in DataLoader workers:
//   AddNegSamples.__call__   unirec/data/transform/addnegsamples.py:90-115  (K draws, <=100 retries each, reject the positive and
//                            every item of the user's history, give up with id 0; uniform or popularity^alpha alias sampling)
//   AddUserHistory.__call__  unirec/data/transform/adduserhistory.py:32-73  ('unorder': zero the target inside the history;
//                            'autoregressive': cut the history before the (last | random) occurrence of the target)
//   SeqRecDataset._padding   unirec/data/dataset/seqrecdataset.py:60-68     (keep the last L items, left-pad with 0)
// The RNG is a counter-based hash (seed, step, sample, draw, try): reproducible, no state; parity with the reference is
// distributional (Python's Mersenne Twister stream cannot be matched) while the rejection rules are exact.
#include "common.cuh"

namespace ur {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {      // splitmix64 finalizer
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t rng(uint64_t seed, uint64_t step, uint64_t b, uint64_t k, uint64_t t) {
    return mix64(mix64(mix64(seed ^ (step * 0xD6E8FEB86659FD93ull)) ^ (b * 0xA24BAED4963EE407ull)) ^ (k * 0x9FB21C651E98DF25ull) ^ t);
}

// membership of `x` in the sorted slice s[0..n)
__device__ __forceinline__ bool contains_sorted(const int32_t* __restrict__ s, int64_t n, int32_t x) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const int32_t v = __ldg(s + mid);
        if (v < x) lo = mid + 1; else hi = mid;
    }
    return lo < n && __ldg(s + lo) == x;
}

struct BatchParams {
    const int64_t* user_id;        // [B]
    const int64_t* pos_item;       // [B]
    const int64_t* hist_ptr;       // [n_users + 1] CSR offsets (null: no history)
    const int32_t* hist_items;     // [nnz] chronological order
    const int32_t* hist_sorted;    // [nnz] per-user ascending copy (membership tests)
    int64_t n_users;
    int64_t n_items;
    const float* alias_prob;       // [n_items] or null -> uniform
    const int32_t* alias_idx;      // [n_items]
    int K, L;
    int mask_mode;                 // 0 none, 1 unorder, 2 autoregressive
    int seq_last;
    uint64_t seed, step;
    int64_t B;
    int64_t* item_id;              // [B, 1+K]
    int32_t* label;                // [B, 1+K]
    int32_t* item_seq;             // [B, L]   (null: no history columns)
    int64_t* item_seq_len;         // [B]
};

__global__ void __launch_bounds__(256) build_batch_kernel(const BatchParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= p.B) return;
    const int64_t u = p.user_id[b];
    const int32_t pos = (int32_t)p.pos_item[b];
    int64_t h0 = 0, hn = 0;
    if (p.hist_ptr && u >= 0 && u < p.n_users) { h0 = p.hist_ptr[u]; hn = p.hist_ptr[u + 1] - h0; }
    const int32_t* hist = p.hist_items + h0;

    // ---- history columns ----
    if (p.item_seq) {
        int64_t cut = hn;                                    // history = hist[0:cut]
        if (p.mask_mode == 2 && hn > 0) {                    // autoregressive: cut before an occurrence of the target
            int64_t count = 0, last = -1;
            for (int64_t i0 = 0; i0 < hn; i0 += 32) {
                const int64_t i = i0 + lane;
                const bool m = i < hn && __ldg(hist + i) == pos;
                const unsigned bal = __ballot_sync(0xffffffffu, m);
                if (bal) { count += __popc(bal); last = i0 + 31 - __clz(bal); }
            }
            if (count > 0) {
                if (p.seq_last) {
                    cut = last;
                } else {
                    int64_t want = (int64_t)(rng(p.seed, p.step, (uint64_t)b, 0xFFFFFFFFull, 0) % (uint64_t)count), seen = 0;
                    for (int64_t i0 = 0; i0 < hn; i0 += 32) {
                        const int64_t i = i0 + lane;
                        const bool m = i < hn && __ldg(hist + i) == pos;
                        const unsigned bal = __ballot_sync(0xffffffffu, m);
                        const int c = __popc(bal);
                        if (want < seen + c) { cut = i0 + (int64_t)__fns(bal, 0, (int)(want - seen) + 1); break; }
                        seen += c;
                    }
                }
            }
        }
        const int64_t n = cut < p.L ? cut : p.L;             // kept items
        const int64_t src0 = cut - n;                        // last n items of hist[0:cut]
        int32_t* out = p.item_seq + b * p.L;
        for (int l = lane; l < p.L; l += 32) {
            int32_t v = 0;
            const int64_t k = l - (p.L - n);
            if (k >= 0) {
                v = __ldg(hist + src0 + k);
                if (p.mask_mode == 1 && v == pos) v = 0;     // unorder: the target is blanked, length unchanged
            }
            out[l] = v;
        }
        if (lane == 0) p.item_seq_len[b] = n;
    }

    // ---- target + K negatives ----
    int64_t* ids = p.item_id + b * (1 + p.K);
    int32_t* lab = p.label + b * (1 + p.K);
    if (lane == 0) { ids[0] = pos; lab[0] = 1; }
    const int32_t* hs = p.hist_sorted ? p.hist_sorted + h0 : nullptr;
    for (int k = lane; k < p.K; k += 32) {
        int64_t chosen = 0;                                  // a failed draw yields the padding id (reference behaviour)
        for (int t = 0; t < 100; ++t) {
            const uint64_t r = rng(p.seed, p.step, (uint64_t)b, (uint64_t)k, (uint64_t)t + 1);
            int64_t cand;
            if (p.alias_prob) {                              // Walker alias table over popularity^alpha (id 0 has weight 0)
                const int64_t slot = (int64_t)((r >> 32) % (uint64_t)p.n_items);
                const float uu = (float)(r & 0xFFFFFF) * (1.f / 16777216.f);
                cand = uu < __ldg(p.alias_prob + slot) ? slot : (int64_t)__ldg(p.alias_idx + slot);
            } else {
                cand = 1 + (int64_t)(r % (uint64_t)(p.n_items - 1));
            }
            if (cand == 0 || cand == pos) continue;
            if (hs && hn > 0 && contains_sorted(hs, hn, (int32_t)cand)) continue;
            chosen = cand;
            break;
        }
        ids[1 + k] = chosen;
        lab[1 + k] = 0;
    }
}

}  // namespace ur

extern "C" {

int ur_build_batch(const int64_t* user_id, const int64_t* pos_item, int64_t B, const int64_t* hist_ptr, const int32_t* hist_items,
                   const int32_t* hist_sorted, int64_t n_users, int64_t n_items, const float* alias_prob, const int32_t* alias_idx,
                   int K, int L, int mask_mode, int seq_last, int64_t seed, int64_t step, int64_t* item_id, int32_t* label,
                   int32_t* item_seq, int64_t* item_seq_len, void* stream) {
    if (K < 0 || L < 0 || n_items < 2 || mask_mode < 0 || mask_mode > 2) return UR_ERR_BAD_ARG;
    if (B == 0) return UR_OK;
    ur::BatchParams p;
    p.user_id = user_id; p.pos_item = pos_item; p.hist_ptr = hist_ptr; p.hist_items = hist_items; p.hist_sorted = hist_sorted;
    p.n_users = n_users; p.n_items = n_items; p.alias_prob = alias_prob; p.alias_idx = alias_idx; p.K = K; p.L = L;
    p.mask_mode = mask_mode; p.seq_last = seq_last; p.seed = (uint64_t)seed; p.step = (uint64_t)step; p.B = B; p.item_id = item_id; p.label = label;
    p.item_seq = L > 0 ? item_seq : nullptr; p.item_seq_len = item_seq_len;
    ur::build_batch_kernel<<<(unsigned)((B + 7) / 8), 256, 0, (cudaStream_t)stream>>>(p);
    UR_RETURN_LAST_ERROR();
}

}  // extern "C"
