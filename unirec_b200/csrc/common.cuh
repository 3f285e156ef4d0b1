// Shared device helpers for the unirec_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define UR_OK 0
#define UR_ERR_BAD_ARG (-1)
#define UR_ERR_UNSUPPORTED (-2)

// Launch-error check that never synchronises: only launch-configuration errors are reported here,
// asynchronous faults surface at the caller's next sync (same contract as a torch op).
#define UR_RETURN_LAST_ERROR()                         \
    do {                                               \
        cudaError_t e__ = cudaGetLastError();          \
        return e__ == cudaSuccess ? UR_OK : -(1000 + (int)e__); \
    } while (0)

namespace ur {

constexpr int kNumSMs = 148;   // B200

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// sum over an aligned group of G lanes (G power of two <= 32)
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 128-bit load: table rows are read once per step, keep them out of L1
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
// same, for data that the kernel also writes (no read-only .nc path)
__device__ __forceinline__ float4 ld_stream_rw(const float4* p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// L2 evict-first policy for data that is touched once per step (table rows, Adam moments): streaming it with normal priority pushes
// the step's activations out of the 126 MB L2 and turns their dirty lines into HBM write-backs in the middle of the streaming kernel
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ld_stream_rw_ef(const float4* p, uint64_t pol) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol) : "memory");
    return r;
}
__device__ __forceinline__ void stg_stream_ef(float4* p, const float4& v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ float4 ldg_stream_ef(const float4* p, uint64_t pol) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    return r;
}
// vector reduction into global memory (no return value): one 16-byte RED per lane
__device__ __forceinline__ void red_add_v4(float* p, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ int64_t load_index(const void* idx, int idx64, int64_t i) {
    return idx64 ? reinterpret_cast<const int64_t*>(idx)[i]
                 : (int64_t)reinterpret_cast<const int32_t*>(idx)[i];
}

__device__ __forceinline__ float4 f4_fma(float a, const float4& x, const float4& y) {
    return make_float4(fmaf(a, x.x, y.x), fmaf(a, x.y, y.y), fmaf(a, x.z, y.z), fmaf(a, x.w, y.w));
}
__device__ __forceinline__ float f4_dot(const float4& a, const float4& b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ float4 f4_scale(const float4& a, float s) {
    return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}
__device__ __forceinline__ float4 f4_add(const float4& a, const float4& b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

enum Act { ACT_NONE = 0, ACT_SWISH = 1, ACT_GELU = 2, ACT_RELU = 3, ACT_TANH = 4, ACT_SIGMOID = 5 };

__device__ __forceinline__ float act_fwd(float x, int act) {
    switch (act) {
        case ACT_SWISH: return x / (1.f + expf(-x));
        case ACT_GELU: return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
        case ACT_RELU: return fmaxf(x, 0.f);
        case ACT_TANH: return tanhf(x);
        case ACT_SIGMOID: return 1.f / (1.f + expf(-x));
        default: return x;
    }
}
// fast-intrinsic variant for the reduced-precision (TF32/BF16) GEMM epilogues: ~6 instructions instead of ~40 per element
__device__ __forceinline__ float act_fwd_fast(float x, int act) {
    switch (act) {
        case ACT_SWISH: return __fdividef(x, 1.f + __expf(-x));
        case ACT_SIGMOID: return __fdividef(1.f, 1.f + __expf(-x));
        case ACT_RELU: return fmaxf(x, 0.f);
        default: return act_fwd(x, act);
    }
}
// derivative of the activation at pre-activation x
__device__ __forceinline__ float act_bwd(float x, int act) {
    switch (act) {
        case ACT_SWISH: { float s = 1.f / (1.f + expf(-x)); return s * (1.f + x * (1.f - s)); }
        case ACT_GELU: {
            float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
            float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
            return cdf + x * pdf;
        }
        case ACT_RELU: return x > 0.f ? 1.f : 0.f;
        case ACT_TANH: { float t = tanhf(x); return 1.f - t * t; }
        case ACT_SIGMOID: { float s = 1.f / (1.f + expf(-x)); return s * (1.f - s); }
        default: return 1.f;
    }
}

// fast-intrinsic derivative for the reduced-precision (TF32) GEMM epilogue
__device__ __forceinline__ float act_bwd_fast(float x, int act) {
    switch (act) {
        case ACT_SWISH: { float s = __fdividef(1.f, 1.f + __expf(-x)); return s * (1.f + x * (1.f - s)); }
        case ACT_SIGMOID: { float s = __fdividef(1.f, 1.f + __expf(-x)); return s * (1.f - s); }
        case ACT_RELU: return x > 0.f ? 1.f : 0.f;
        default: return act_bwd(x, act);
    }
}

}  // namespace ur
