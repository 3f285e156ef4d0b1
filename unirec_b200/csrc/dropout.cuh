// Counter-based dropout for the fused encoder: no mask tensors, the keep decision of element e of dropout site s in training
// step t is a pure function of (seed, t, s, e), so the backward kernels regenerate exactly the forward's mask.
// Reference sites: nn.Dropout at unirec/model/sequential/sasrec.py:69 (input), unirec/model/modules.py:307 (attention
// probabilities), :313 (attention output), :352 (FFN output), unirec/model/sequential/gru.py:29 (item embeddings).
// Generator: Philox4x32-10 (Salmon et al., SC'11), key = (seed lo, seed hi), counter = (e/4 lo, e/4 hi, step, site); element e
// takes word e%4 of the block.  keep <=> word >= floor(p * 2^32); kept values are scaled by 1/(1-p) like nn.Dropout.
// oracle/philox.py restates the same function in numpy (pinned by the Random123 known-answer vectors).
#pragma once
#include <stdint.h>

namespace ur {

struct DropCfg {
    uint32_t k0, k1;      // seed
    uint32_t step, site;  // counter words 2, 3
    uint32_t thresh;      // drop when word < thresh
    float scale;          // 1 / (1 - p)
    bool on;
};

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += W0;
        k1 += W1;
    }
    return make_uint4(c0, c1, c2, c3);
}

// rng: device int64[2] = (seed, step).  p <= 0 or rng == null: dropout off (identity, bit-exact with the p = 0 build).
__device__ __forceinline__ DropCfg drop_cfg(const long long* __restrict__ rng, float p, int site) {
    DropCfg c;
    c.on = rng != nullptr && p > 0.f;
    c.k0 = c.k1 = c.step = 0u;
    c.site = (uint32_t)site;
    c.thresh = 0u;
    c.scale = 1.f;
    if (c.on) {
        const unsigned long long seed = (unsigned long long)rng[0];
        c.k0 = (uint32_t)seed;
        c.k1 = (uint32_t)(seed >> 32);
        c.step = (uint32_t)(unsigned long long)rng[1];
        const double t = (double)p * 4294967296.0;
        c.thresh = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
        c.scale = 1.f / (1.f - p);
    }
    return c;
}

// multipliers (0 or scale) of the four elements 4*g .. 4*g+3
__device__ __forceinline__ float4 drop_mask4(const DropCfg& c, unsigned long long g) {
    const uint4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), c.step, c.site, c.k0, c.k1);
    return make_float4(r.x >= c.thresh ? c.scale : 0.f, r.y >= c.thresh ? c.scale : 0.f, r.z >= c.thresh ? c.scale : 0.f,
                       r.w >= c.thresh ? c.scale : 0.f);
}

// multiplier of the single element e
__device__ __forceinline__ float drop_mask1(const DropCfg& c, unsigned long long e) {
    const uint4 r = philox4x32_10((uint32_t)(e >> 2), (uint32_t)(e >> 34), c.step, c.site, c.k0, c.k1);
    const uint32_t sel = (uint32_t)e & 3u;
    const uint32_t w = sel == 0 ? r.x : (sel == 1 ? r.y : (sel == 2 ? r.z : r.w));
    return w >= c.thresh ? c.scale : 0.f;
}

__device__ __forceinline__ float4 f4_mul(const float4& a, const float4& b) {
    return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}

}  // namespace ur
