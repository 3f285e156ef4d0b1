// f3: one-vs-all ranking on the device.  For every user of an evaluation batch: the number of catalogue items whose score
// exceeds the target's, with the user's history, the padding id and the target itself excluded -- without ever forming the
// [B, V] score matrix and without moving table rows between GPUs (each rank counts over the rows it owns; counts add).
// Reference: Evaluator.evaluate_with_full_items, unirec/facility/evaluation/evaluator_abc.py:190-278
//   scores = (U E^T + item_bias + user_bias) / tau                                     (:232-241)
//   target_score = scores[idx, itemid]; scores[idx][history] = NINF (= -9999, :46)      (:249-253)
//   scores[idx][0] = target_score; scores[idx][itemid] = NINF                           (:255-257)
// and get_rank, onepos.py:20-31: rank = #{j >= 1 : S[j] > S[0]}.  The reference adds +-1e-8 tie-breaking noise
// (onepos.py:116-120); here exact ties are not counted.
// Every dot product is accumulated as acc = fmaf(u[k], e[k], acc) for k = 0..d-1 in ONE thread, in all three kernels, so
// the target's score and an item's score are bit-identical functions of the same rows (an item with the target's row ties).
#include "common.cuh"

namespace ur {

constexpr float kNINF = -9999.f;        // evaluator_abc.py:46
constexpr int EBM = 128, EBN = 128, EBK = 8;

__device__ __forceinline__ float seq_dot(const float* __restrict__ u, const float* __restrict__ e, int d) {
    float acc = 0.f;
    for (int k = 0; k < d; k += 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(u + k));
        const float4 b = __ldg(reinterpret_cast<const float4*>(e + k));
        acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    }
    return acc;
}

// tscore[s] = score of the target item of user s when this rank owns it, else 0 (summed across ranks by the caller)
__global__ void __launch_bounds__(128) rank_target_kernel(const float* __restrict__ table, int d, const float* __restrict__ user,
                                                          const int64_t* __restrict__ target, int64_t S,
                                                          const float* __restrict__ item_bias, const float* __restrict__ user_bias,
                                                          const int64_t* __restrict__ user_id, float inv_tau, int W, int r,
                                                          float* __restrict__ tscore) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int64_t g = target[s];
    float out = 0.f;
    if (g % W == r) {
        const float dot = seq_dot(user + s * d, table + (g / W) * d, d);
        const float ib = item_bias ? __ldg(item_bias + g) : 0.f;
        const float ub = user_bias ? __ldg(user_bias + __ldg(user_id + s)) : 0.f;
        out = (dot + ib + ub) * inv_tau;
    }
    tscore[s] = out;
}

// counts[s] += #{owned rows i : g = i*W + r, g != 0, g != target[s], score(s, g) > tscore[s]}.  128 users x 128 items per CTA,
// 8 x 8 register tile per thread, k in order (bit-identical to seq_dot).
__global__ void __launch_bounds__(256) rank_count_kernel(const float* __restrict__ table, int64_t n_local, int d,
                                                         const float* __restrict__ user, int64_t S,
                                                         const int64_t* __restrict__ target, const float* __restrict__ tscore,
                                                         const float* __restrict__ item_bias, const float* __restrict__ user_bias,
                                                         const int64_t* __restrict__ user_id, float inv_tau, int W, int r,
                                                         int32_t* __restrict__ counts) {
    __shared__ __align__(16) float As[2][EBK][EBM];     // users
    __shared__ __align__(16) float Bs[2][EBK][EBN];     // items
    const int t = threadIdx.x;
    const int64_t n0 = (int64_t)blockIdx.x * EBN;
    const int64_t m0 = (int64_t)blockIdx.y * EBM;
    const int tx = t & 15, ty = t >> 4;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    auto load = [&](const float* __restrict__ P, int64_t row0, int64_t nrows, int k0) {
        const int64_t row = row0 + (t >> 1);
        const int k = k0 + (t & 1) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < nrows && k < d) v = __ldg(reinterpret_cast<const float4*>(P + row * d + k));
        return v;
    };
    auto store = [&](float (*Sm)[EBM], const float4& v) {
        const int row = t >> 1, k = (t & 1) * 4;
        Sm[k + 0][row] = v.x; Sm[k + 1][row] = v.y; Sm[k + 2][row] = v.z; Sm[k + 3][row] = v.w;
    };
    float4 ra = load(user, m0, S, 0), rb = load(table, n0, n_local, 0);
    store(As[0], ra);
    store(Bs[0], rb);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < d; k0 += EBK) {
        const bool more = k0 + EBK < d;
        if (more) { ra = load(user, m0, S, k0 + EBK); rb = load(table, n0, n_local, k0 + EBK); }
#pragma unroll
        for (int kk = 0; kk < EBK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) {
            store(As[buf ^ 1], ra);
            store(Bs[buf ^ 1], rb);
            __syncthreads();
            buf ^= 1;
        }
    }
    // epilogue: compare against the target score, count per user
    int64_t gid[8];
    float ib[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int64_t row = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
        gid[j] = row < n_local ? row * W + r : 0;        // 0 = padding id: never counted
        ib[j] = (item_bias && gid[j] > 0) ? __ldg(item_bias + gid[j]) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        int c = 0;
        if (m < S) {
            const float ts = __ldg(tscore + m);
            const int64_t tg = __ldg(target + m);
            const float ub = user_bias ? __ldg(user_bias + __ldg(user_id + m)) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float sc = (acc[i][j] + ib[j] + ub) * inv_tau;
                c += (gid[j] > 0 && gid[j] != tg && sc > ts) ? 1 : 0;
            }
        }
        // the 16 threads of one ty (half a warp) hold the same user rows
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (tx == 0 && m < S && c) atomicAdd(counts + m, c);
    }
}

// History exclusion (evaluator_abc.py:249-257): every DISTINCT owned history item h (h != 0, h != target) contributes
// [NINF > t] instead of [score(h) > t]; the target's own slot contributes [NINF > t] (its owner adds it).  One warp per user,
// one lane per history item, sorted CSR slices (duplicates are adjacent).
__global__ void __launch_bounds__(256) rank_exclude_kernel(const float* __restrict__ table, int d, const float* __restrict__ user,
                                                           int64_t S, const int64_t* __restrict__ target,
                                                           const float* __restrict__ tscore, const float* __restrict__ item_bias,
                                                           const float* __restrict__ user_bias, const int64_t* __restrict__ user_id,
                                                           float inv_tau, int W, int r, const int64_t* __restrict__ hist_ptr,
                                                           const int32_t* __restrict__ hist_sorted, int64_t n_hist_users,
                                                           int32_t* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= S) return;
    const float ts = tscore[s];
    const int64_t tg = target[s];
    const int ninf_wins = kNINF > ts ? 1 : 0;
    const float ub = user_bias ? __ldg(user_bias + __ldg(user_id + s)) : 0.f;
    int delta = 0;
    if (lane == 0 && tg > 0 && tg % W == r) delta += ninf_wins;
    const int64_t u = user_id ? user_id[s] : -1;
    if (hist_ptr && u >= 0 && u < n_hist_users) {
        const int64_t beg = hist_ptr[u], end = hist_ptr[u + 1];
        for (int64_t k = beg + lane; k < end; k += 32) {
            const int64_t h = hist_sorted[k];
            if (h <= 0 || h == tg || h % W != r) continue;
            if (k > beg && hist_sorted[k - 1] == h) continue;            // duplicate: masked once
            const float dot = seq_dot(user + s * d, table + (h / W) * d, d);
            const float sc = (dot + (item_bias ? __ldg(item_bias + h) : 0.f) + ub) * inv_tau;
            delta += ninf_wins - (sc > ts ? 1 : 0);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) delta += __shfl_xor_sync(0xffffffffu, delta, o);
    if (lane == 0 && delta) atomicAdd(counts + s, delta);
}

}  // namespace ur

extern "C" {

int ur_rank_target_f32(const float* table_local, int d, const float* user_emb, const int64_t* target, int64_t S,
                       const float* item_bias, const float* user_bias, const int64_t* user_id, float tau, int world, int rank,
                       float* tscore, void* stream) {
    if (d <= 0 || (d & 3) || world < 1 || rank < 0 || rank >= world || tau == 0.f) return UR_ERR_BAD_ARG;
    if (user_bias && !user_id) return UR_ERR_BAD_ARG;
    if (S == 0) return UR_OK;
    ur::rank_target_kernel<<<(unsigned)((S + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        table_local, d, user_emb, target, S, item_bias, user_bias, user_id, 1.f / tau, world, rank, tscore);
    UR_RETURN_LAST_ERROR();
}

int ur_rank_count_f32(const float* table_local, int64_t n_local, int d, const float* user_emb, int64_t S, const int64_t* target,
                      const float* tscore, const float* item_bias, const float* user_bias, const int64_t* user_id, float tau,
                      int world, int rank, int32_t* counts, void* stream) {
    if (d <= 0 || (d & 3) || world < 1 || rank < 0 || rank >= world || tau == 0.f) return UR_ERR_BAD_ARG;
    if (user_bias && !user_id) return UR_ERR_BAD_ARG;
    if (S == 0 || n_local == 0) return UR_OK;
    const int64_t gx = (n_local + ur::EBN - 1) / ur::EBN, gy = (S + ur::EBM - 1) / ur::EBM;
    if (gy > 65535) return UR_ERR_UNSUPPORTED;
    dim3 grid((unsigned)gx, (unsigned)gy);
    ur::rank_count_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(table_local, n_local, d, user_emb, S, target, tscore, item_bias,
                                                                  user_bias, user_id, 1.f / tau, world, rank, counts);
    UR_RETURN_LAST_ERROR();
}

int ur_rank_exclude_f32(const float* table_local, int d, const float* user_emb, int64_t S, const int64_t* target,
                        const float* tscore, const float* item_bias, const float* user_bias, const int64_t* user_id, float tau,
                        int world, int rank, const int64_t* hist_ptr, const int32_t* hist_sorted, int64_t n_hist_users,
                        int32_t* counts, void* stream) {
    if (d <= 0 || (d & 3) || world < 1 || rank < 0 || rank >= world || tau == 0.f) return UR_ERR_BAD_ARG;
    if ((user_bias || hist_ptr) && !user_id) return UR_ERR_BAD_ARG;
    if (S == 0) return UR_OK;
    ur::rank_exclude_kernel<<<(unsigned)((S + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        table_local, d, user_emb, S, target, tscore, item_bias, user_bias, user_id, 1.f / tau, world, rank, hist_ptr, hist_sorted,
        n_hist_users, counts);
    UR_RETURN_LAST_ERROR();
}

}  // extern "C"
