// K11 + K12 + K13: row-sparse embedding-gradient reduce fused with the optimizer, plus the flat dense optimizer.
//
// The reference produces a DENSE [V,d] gradient per table (embedding_dense_backward) and runs dense Adam over the
// whole table every step (unirec/facility/trainer.py:136,346-349).  Here the per-row gradient is never stored:
//   1. `ur_rowlist_link`   threads every batch entry (b,j) onto a per-row linked list: next[e] = atomicExch(head[id], e);
//                          the entry that finds the list empty also appends `id` to the compact `uniq` list.
//   2. `ur_rowlist_apply`  one lane-group per unique row walks its list, accumulates  g = sum_e coef_e * src_row(e)
//                          in registers (src rows are the small L2-resident [B,d] / [B*L,d] activations-gradients),
//                          then updates (param, m, v) in place and resets head[id] = -1.
// HBM traffic per touched row: read+write of param/m/v (6 d floats) - no gradient buffer, no zero-fill, no atomics
// on floating-point data.  `mode` selects Adam / AdamW / SGD arithmetic, or a squared-norm pass (global-norm clip).
#include "common.cuh"

namespace ur {

struct RowSource {
    const float* src;     // source rows, row stride = d
    const float* coef;    // per-entry (coef_group == 1) or per-source-row coefficient, or null (= 1)
    int64_t src_group;    // entries per source row
    int64_t coef_group;   // entries per coefficient
    const long long* part_ptrs;   // optional: source rows live in `part_rows`-row buffers on different GPUs (peer pointers, p2p.cu):
    int64_t part_rows;            // source row q is row q % part_rows of buffer part_ptrs[q / part_rows]
};

__global__ void __launch_bounds__(256) rowlist_link_kernel(int32_t* __restrict__ head, const void* __restrict__ keys, int idx64,
                                                           int64_t n, int32_t entry_offset, int32_t* __restrict__ next,
                                                           int32_t* __restrict__ uniq, int32_t* __restrict__ n_uniq, int64_t pad_id,
                                                           int W, int r, int64_t key_mask) {
    // The unique-row list is appended through ONE global counter.  Same-address atomics retire at ~3 ns each, so a warp-level
    // append (33 K atomics for 1.05 M entries) cost ~100 us; the append is aggregated per CTA instead (one atomic per 256 entries).
    __shared__ int warp_cnt[8];
    __shared__ int cta_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < n; base += (int64_t)gridDim.x * blockDim.x) {   // uniform per CTA
        const int64_t e = base + threadIdx.x;
        bool first = false;
        int64_t id = pad_id;
        if (e < n) {
            // keys may carry flag bits (packed ids: label in bit 31) and, for a row-sharded table, be GLOBAL ids: only the entries
            // this rank owns are linked, under their local row id / W
            id = load_index(keys, idx64, e) & key_mask;
            bool live = id != pad_id;
            if (live && W > 1) { live = (id % W) == r; id /= W; }
            if (live) {
                const int32_t prev = atomicExch(head + id, entry_offset + (int32_t)e);
                next[entry_offset + e] = prev;
                first = prev < 0;
            } else {
                next[entry_offset + e] = -1;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, first);
        if (lane == 0) warp_cnt[warp] = __popc(m);
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) { const int c = warp_cnt[w]; warp_cnt[w] = tot; tot += c; }   // exclusive prefix in place
            cta_base = tot ? atomicAdd(n_uniq, tot) : 0;
        }
        __syncthreads();
        if (first) uniq[cta_base + warp_cnt[warp] + __popc(m & ((1u << lane) - 1))] = (int32_t)id;
        __syncthreads();                                   // warp_cnt / cta_base are rewritten by the next iteration
    }
}

enum OptMode { OPT_ADAM = 0, OPT_ADAMW = 1, OPT_SGD = 2, OPT_SQNORM = 3 };

struct OptHyper {
    float lr, beta1, beta2, eps, weight_decay;
    const int32_t* step_dev;        // 1-based step about to be applied
    const float* grad_scale_dev;    // global-norm clip coefficient or null
    const int32_t* skip_flag;       // nonzero -> leave parameters untouched (NaN-loss skip, trainer.py:344-352)
};

template <int D4>
__global__ void __launch_bounds__(256) rowlist_apply_kernel(float4* table, float4* mom, float4* var,
                                                            int32_t* __restrict__ head, const int32_t* __restrict__ next,
                                                            const int32_t* __restrict__ uniq, const int32_t* __restrict__ n_uniq,
                                                            RowSource s0, int32_t n0, RowSource s1, int mode, OptHyper h,
                                                            float* __restrict__ sqnorm_out, const int32_t* __restrict__ u_begin_dev,
                                                            const int32_t* __restrict__ u_end_dev) {
    constexpr int LPR = D4 < 32 ? D4 : 32;
    constexpr int VPL = D4 / LPR;
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane / LPR, col = lane % LPR;
    const int64_t group0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + sub;
    const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x >> 5) * RPW;
    // [u_begin, u_end) of the unique-row list (device-resident bounds; default: the whole list).  Used to update the rows that
    // only the scorer touches while the encoder backward is still producing the gradients of the history rows.
    const int ub = u_begin_dev ? *u_begin_dev : 0;
    const int nu = u_end_dev ? *u_end_dev : *n_uniq;
    const bool skip = h.skip_flag && *h.skip_flag != 0;
    const float gs = h.grad_scale_dev ? *h.grad_scale_dev : 1.f;
    const uint64_t pol = l2_evict_first_policy();        // parameter / moment rows are touched once per step: stream them through L2
    float bc1 = 1.f, bc2s = 1.f;
    if (mode == OPT_ADAM || mode == OPT_ADAMW) {
        const float t = (float)(*h.step_dev);
        bc1 = 1.f - powf(h.beta1, t);
        bc2s = sqrtf(1.f - powf(h.beta2, t));
    }
    float sq = 0.f;
    const bool is_adam = (mode == OPT_ADAM || mode == OPT_ADAMW);
    const bool update = (mode != OPT_SQNORM) && !skip;
    // (row id, list head) of the NEXT row are fetched one iteration ahead: the chain uniq -> head (a random 4-byte read of a
    // 40 MB array) is ~1.3 us of the ~2.3 us dependent chain per row, so taking it off the critical path lets the same number
    // of resident warps keep nearly twice as many row loads in flight.
    int64_t u = ub + group0;
    int64_t id = u < nu ? (int64_t)__ldg(uniq + u) : -1;
    int32_t e = u < nu ? head[id] : -1;
    for (; u < nu; u += ngroups) {
        const int64_t u_next = u + ngroups;
        const int64_t id_next = u_next < nu ? (int64_t)__ldg(uniq + u_next) : -1;
        // Issue the parameter / moment loads FIRST: they do not depend on the list walk, so their HBM latency overlaps
        // the dependent chain entry -> source row below.
        float4 P[VPL], M[VPL], W[VPL];
        if (update) {
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int64_t o = id * D4 + v * LPR + col;
                P[v] = ld_stream_rw_ef(table + o, pol);
                if (is_adam) { M[v] = ld_stream_rw_ef(mom + o, pol); W[v] = ld_stream_rw_ef(var + o, pol); }
            }
        }
        const int32_t e_next = id_next >= 0 ? head[id_next] : -1;       // distinct row: not the head reset below
        float4 g[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) g[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        while (e >= 0) {
            const bool first = e < n0;
            const RowSource& s = first ? s0 : s1;
            const int64_t el = first ? e : e - n0;
            const float c = s.coef ? __ldg(s.coef + el / s.coef_group) : 1.f;
            const int64_t q = el / s.src_group;
            const float4* row = s.part_ptrs ? reinterpret_cast<const float4*>(s.part_ptrs[q / s.part_rows]) + (q % s.part_rows) * D4
                                            : reinterpret_cast<const float4*>(s.src) + q * D4;
            const int32_t nx = __ldg(next + e);
#pragma unroll
            for (int v = 0; v < VPL; ++v) g[v] = f4_fma(c, __ldg(row + v * LPR + col), g[v]);
            e = nx;
        }
        if (mode == OPT_SQNORM) {
#pragma unroll
            for (int v = 0; v < VPL; ++v) sq += f4_dot(g[v], g[v]);
            id = id_next; e = e_next;
            continue;
        }
        if (col == 0) head[id] = -1;
        if (skip) { id = id_next; e = e_next; continue; }
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int64_t o = id * D4 + v * LPR + col;
            float4 p = P[v];
            float4 gg = f4_scale(g[v], gs);
            if (mode == OPT_SGD) {
                if (h.weight_decay != 0.f) gg = f4_fma(h.weight_decay, p, gg);
                stg_stream_ef(table + o, f4_fma(-h.lr, gg, p), pol);
            } else {
                if (mode == OPT_ADAM && h.weight_decay != 0.f) gg = f4_fma(h.weight_decay, p, gg);
                if (mode == OPT_ADAMW && h.weight_decay != 0.f) p = f4_scale(p, 1.f - h.lr * h.weight_decay);
                float4 m = M[v], w = W[v];
                const float o1 = 1.f - h.beta1, o2 = 1.f - h.beta2;
                m.x = h.beta1 * m.x + o1 * gg.x; m.y = h.beta1 * m.y + o1 * gg.y;
                m.z = h.beta1 * m.z + o1 * gg.z; m.w = h.beta1 * m.w + o1 * gg.w;
                w.x = h.beta2 * w.x + o2 * gg.x * gg.x; w.y = h.beta2 * w.y + o2 * gg.y * gg.y;
                w.z = h.beta2 * w.z + o2 * gg.z * gg.z; w.w = h.beta2 * w.w + o2 * gg.w * gg.w;
                const float a = h.lr / bc1;
                p.x -= a * m.x / (sqrtf(w.x) / bc2s + h.eps); p.y -= a * m.y / (sqrtf(w.y) / bc2s + h.eps);
                p.z -= a * m.z / (sqrtf(w.z) / bc2s + h.eps); p.w -= a * m.w / (sqrtf(w.w) / bc2s + h.eps);
                stg_stream_ef(mom + o, m, pol); stg_stream_ef(var + o, w, pol); stg_stream_ef(table + o, p, pol);
            }
        }
        id = id_next; e = e_next;
    }
    if (mode == OPT_SQNORM) {
        sq = warp_sum(sq);
        if (lane == 0 && sq != 0.f) atomicAdd(sqnorm_out, sq);
    }
}

// flat dense optimizer over one contiguous parameter buffer (the whole encoder, or a whole table in exact-dense mode)
__global__ void __launch_bounds__(256) dense_opt_kernel(float4* __restrict__ p4, const float4* __restrict__ g4, float4* __restrict__ m4,
                                                        float4* __restrict__ v4, int64_t n4, int mode, OptHyper h) {
    if (h.skip_flag && *h.skip_flag != 0) return;
    const float gs = h.grad_scale_dev ? *h.grad_scale_dev : 1.f;
    float bc1 = 1.f, bc2s = 1.f;
    if (mode != OPT_SGD) {
        const float t = (float)(*h.step_dev);
        bc1 = 1.f - powf(h.beta1, t);
        bc2s = sqrtf(1.f - powf(h.beta2, t));
    }
    const float a = h.lr / bc1, o1 = 1.f - h.beta1, o2 = 1.f - h.beta2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 p = p4[i];
        float4 g = f4_scale(g4[i], gs);
        if (mode == OPT_SGD) {
            if (h.weight_decay != 0.f) g = f4_fma(h.weight_decay, p, g);
            p4[i] = f4_fma(-h.lr, g, p);
            continue;
        }
        if (mode == OPT_ADAM && h.weight_decay != 0.f) g = f4_fma(h.weight_decay, p, g);
        if (mode == OPT_ADAMW && h.weight_decay != 0.f) p = f4_scale(p, 1.f - h.lr * h.weight_decay);
        float4 m = m4[i], w = v4[i];
        m.x = h.beta1 * m.x + o1 * g.x; m.y = h.beta1 * m.y + o1 * g.y; m.z = h.beta1 * m.z + o1 * g.z; m.w = h.beta1 * m.w + o1 * g.w;
        w.x = h.beta2 * w.x + o2 * g.x * g.x; w.y = h.beta2 * w.y + o2 * g.y * g.y;
        w.z = h.beta2 * w.z + o2 * g.z * g.z; w.w = h.beta2 * w.w + o2 * g.w * g.w;
        p.x -= a * m.x / (sqrtf(w.x) / bc2s + h.eps); p.y -= a * m.y / (sqrtf(w.y) / bc2s + h.eps);
        p.z -= a * m.z / (sqrtf(w.z) / bc2s + h.eps); p.w -= a * m.w / (sqrtf(w.w) / bc2s + h.eps);
        m4[i] = m; v4[i] = w; p4[i] = p;
    }
}

__global__ void __launch_bounds__(256) sqnorm_kernel(const float4* __restrict__ g4, int64_t n4, float* __restrict__ out) {
    float s = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 g = g4[i];
        s += f4_dot(g, g);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0 && s != 0.f) atomicAdd(out, s);
}

// clip coefficient = min(1, max_norm / (sqrt(sqnorm) + 1e-6))   (torch.nn.utils.clip_grad_norm_, trainer.py:347-348)
__global__ void clip_coef_kernel(const float* sqnorm, float max_norm, float* coef) {
    const float c = max_norm / (sqrtf(*sqnorm) + 1e-6f);
    *coef = c < 1.f ? c : 1.f;
}

__global__ void step_advance_kernel(int32_t* step, const int32_t* skip_flag) {
    if (!(skip_flag && *skip_flag != 0)) *step += 1;
}

}  // namespace ur

extern "C" {

int ur_rowlist_link(int32_t* head, const void* keys, int idx_bits, int64_t n, int64_t entry_offset, int32_t* next, int32_t* uniq,
                    int32_t* n_uniq, int64_t pad_id, int world, int rank, int64_t key_mask, void* stream) {
    if (idx_bits != 32 && idx_bits != 64) return UR_ERR_BAD_ARG;
    if (world < 1 || rank < 0 || rank >= world) return UR_ERR_BAD_ARG;
    if (entry_offset + n >= (int64_t)1 << 31) return UR_ERR_UNSUPPORTED;
    if (n == 0) return UR_OK;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)ur::kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    ur::rowlist_link_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(head, keys, idx_bits == 64, n, (int32_t)entry_offset,
                                                                                next, uniq, n_uniq, pad_id, world, rank, key_mask);
    UR_RETURN_LAST_ERROR();
}

int ur_rowlist_apply_f32(float* table, float* mom, float* var, int d, int32_t* head, const int32_t* next, const int32_t* uniq,
                         const int32_t* n_uniq, int64_t max_uniq, const float* src0, int64_t src0_group, const float* coef0,
                         int64_t coef0_group, int64_t n0, const float* src1, int64_t src1_group, const float* coef1,
                         int64_t coef1_group, int mode, float lr, float beta1, float beta2, float eps, float weight_decay,
                         const int32_t* step_dev, const float* grad_scale_dev, const int32_t* skip_flag, float* sqnorm_out,
                         const int32_t* u_begin_dev, const int32_t* u_end_dev, int small_ctas, const void* src1_part_ptrs,
                         int64_t src1_part_rows, void* stream) {
    if (src1_part_ptrs && src1_part_rows < 1) return UR_ERR_BAD_ARG;
    if (d <= 0 || (d & 3) || mode < 0 || mode > 3) return UR_ERR_BAD_ARG;
    if (max_uniq == 0) return UR_OK;
    ur::RowSource s0{src0, coef0, src0_group > 0 ? src0_group : 1, coef0_group > 0 ? coef0_group : 1, nullptr, 1};
    ur::RowSource s1{src1, coef1, src1_group > 0 ? src1_group : 1, coef1_group > 0 ? coef1_group : 1,
                     (const long long*)src1_part_ptrs, src1_part_rows > 0 ? src1_part_rows : 1};
    ur::OptHyper h{lr, beta1, beta2, eps, weight_decay, step_dev, grad_scale_dev, skip_flag};
    const int rpw = d >= 128 ? 1 : 128 / d;
    // small_ctas: 128-thread CTAs (6 K registers) that fit next to a resident persistent GEMM CTA when this launch runs on a
    // side stream concurrently with the encoder backward
    const int threads = small_ctas ? 128 : 256;
    const int wpc = threads / 32;
    int64_t blocks = (max_uniq + wpc * rpw - 1) / (wpc * rpw);
    const int64_t cap = (int64_t)ur::kNumSMs * (small_ctas ? 32 : 16);
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
#define UR_CASE(D)                                                                                                        \
    case D:                                                                                                               \
        ur::rowlist_apply_kernel<D / 4><<<(unsigned)blocks, threads, 0, st>>>((float4*)table, (float4*)mom, (float4*)var, head,   \
                                                                              next, uniq, n_uniq, s0, (int32_t)n0, s1, mode, h,    \
                                                                              sqnorm_out, u_begin_dev, u_end_dev);                 \
        break;
    switch (d) {
        UR_CASE(4) UR_CASE(16) UR_CASE(32) UR_CASE(64) UR_CASE(128) UR_CASE(256) UR_CASE(512)
        default: return UR_ERR_UNSUPPORTED;
    }
#undef UR_CASE
    UR_RETURN_LAST_ERROR();
}

int ur_dense_opt_f32(float* param, const float* grad, float* mom, float* var, int64_t n, int mode, float lr, float beta1,
                     float beta2, float eps, float weight_decay, const int32_t* step_dev, const float* grad_scale_dev,
                     const int32_t* skip_flag, void* stream) {
    if ((n & 3) || mode < 0 || mode > 2) return UR_ERR_BAD_ARG;
    if (n == 0) return UR_OK;
    ur::OptHyper h{lr, beta1, beta2, eps, weight_decay, step_dev, grad_scale_dev, skip_flag};
    int64_t blocks = (n / 4 + 255) / 256;
    const int64_t cap = (int64_t)ur::kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    ur::dense_opt_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((float4*)param, (const float4*)grad, (float4*)mom,
                                                                             (float4*)var, n / 4, mode, h);
    UR_RETURN_LAST_ERROR();
}

int ur_sqnorm_accum_f32(const float* grad, int64_t n, float* sqnorm, void* stream) {
    if (n & 3) return UR_ERR_BAD_ARG;
    if (n == 0) return UR_OK;
    int64_t blocks = (n / 4 + 255) / 256;
    const int64_t cap = (int64_t)ur::kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    ur::sqnorm_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)grad, n / 4, sqnorm);
    UR_RETURN_LAST_ERROR();
}

int ur_clip_coef_f32(const float* sqnorm, float max_norm, float* coef, void* stream) {
    ur::clip_coef_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sqnorm, max_norm, coef);
    UR_RETURN_LAST_ERROR();
}

int ur_step_advance(int32_t* step, const int32_t* skip_flag, void* stream) {
    ur::step_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step, skip_flag);
    UR_RETURN_LAST_ERROR();
}

}  // extern "C"
