// Tensor-core GEMM for the encoder's dense layers (K4/K7): tcgen05.mma (kind::tf32, fp32 operands read in place, fp32
// accumulate in TMEM), operands staged by TMA (cp.async.bulk.tensor, 128B swizzle) through an mbarrier ring.
//
//   C[M, N] (+)= act(op(A) * op(B) + bias)
//
// Operand layouts, selected by the caller's transposes (template flags kAmn / kBmn = operand is MN-major in memory):
//   NT  (forward, y = x W^T)        : A [M,K] and B [N,K] both K-major
//   NN  (input grads, dx = dy W)    : A [M,K] K-major, B stored [K,N] MN-major (no transposed weight copies)
//   TN  (weight grads, dW = dy^T x) : A stored [K,M], B stored [K,N], both MN-major, split-K over the token dimension
//
// Persistent kernel, one CTA per SM looping over (row stripe, column chunk, k split) work units.  Warp roles: warp 0 = TMA
// producer, warp 1 = TMEM allocator + single-thread MMA issuer, warps 2..9 = epilogue (tcgen05.ld -> bias / activation /
// activation-backward / column sums -> swizzled shared-memory chunk -> TMA store or reduce-add), warps 10..13 = operand
// splitter of the 3xTF32 mode.  Two TMEM accumulators: the epilogue of one unit overlaps the main loop of the next.
#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include <unordered_map>
#include "common.cuh"

namespace ur {
namespace tc {

constexpr int BM = 128;          // rows of C per CTA (UMMA M)
constexpr int BK = 32;           // fp32 elements per k-block = 128 bytes = one swizzle span
constexpr int MAX_NC = 256;      // columns of C per work unit: two accumulators of NC columns fill the 512 TMEM columns

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand read from tensor memory (lane = row of the tile, one 32-bit column per k element): D += A_tmem * B_smem
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_c, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_c), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// registers -> tensor memory: this thread's 32 values go to 32 consecutive columns of its lane (warp w owns lanes 32*(w%4)..)
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const float4 (&x)[8]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "f"(x[0].x), "f"(x[0].y), "f"(x[0].z), "f"(x[0].w), "f"(x[1].x), "f"(x[1].y), "f"(x[1].z), "f"(x[1].w),
          "f"(x[2].x), "f"(x[2].y), "f"(x[2].z), "f"(x[2].w), "f"(x[3].x), "f"(x[3].y), "f"(x[3].z), "f"(x[3].w),
          "f"(x[4].x), "f"(x[4].y), "f"(x[4].z), "f"(x[4].w), "f"(x[5].x), "f"(x[5].y), "f"(x[5].z), "f"(x[5].w),
          "f"(x[6].x), "f"(x[6].y), "f"(x[6].z), "f"(x[6].w), "f"(x[7].x), "f"(x[7].y), "f"(x[7].z), "f"(x[7].w)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor (SM100 UMMA): 128-byte swizzle, version 1
//   K-major : rows of 128 B (32 fp32 of K), 8-row groups 1024 B apart (SBO), LBO unused (=1)
//   MN-major: rows of 128 B (32 fp32 of M/N) per k, 8 k-rows = one 1024 B atom; LBO = distance between 32-wide MN chunks,
//             SBO = distance between 8-row k groups
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;      // descriptor version (Blackwell)
    d |= (uint64_t)layout_type << 61;      // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B (the only MN-major layout for tf32)
    return d;
}
// instruction descriptor: D=f32, A=B=tf32, M=128, N=n; major bits: 0 = K-major, 1 = MN-major
__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

struct Params {
    int M, N, K;            // logical GEMM sizes (K = reduction length)
    int NC;                 // columns handled by one work unit (128 or 256)
    int stages;
    float* C; int64_t ldc;
    const float* bias; int act;
    float* preact; int64_t ldp;
    const float* dact; int64_t ldd;   // optional: C = acc * act'(dact[row, col])  (fused activation backward)
    float* colsum;                    // optional: colsum[col] += sum over rows of the stored C values (fused bias gradient)
    int accumulate;         // 0 store, 1 atomic add (split-K partials), 2 read-add-store
    int kb_per_split;       // k-blocks handled by one work unit along the split dimension
    int epi_bufs;           // store buffers per epilogue warp (1 or 2)
    int lo_stages;          // 3xTF32: depth of the lo ring
    int split_warps;        // 3xTF32: number of operand-splitter warps (4 or 8)
    int b_lo_tma;           // 3xTF32 + a_tmem: the lo part of B (a weight) is read from its precomputed lo plane by TMA (tmBlo): no B split
    int dual_mma;           // 3xTF32: a second MMA-issuing warp (launch with 32 more threads)
    int a_tmem;             // 3xTF32, K-major A: hi/lo of the A tile live in tensor memory (columns 256..511), the lo ring holds B only
    int split_trunc;        // 3xTF32: 1 = hi operand is the raw tile (hardware truncation), only lo is written
    int prof;               // bring-up: accumulate the role clocks into g_tc_prof
    int dbg_skip;           // bring-up: bit0 skip TMA store issue, bit1 skip bias, bit2 skip smem staging
    int n_chunks, m_stripes, total_units, splits;
    const int* rows_dev; int rows_dim;   // optional device-resident token count: rows_dim 1 -> M = min(M, *rows_dev) (row-parallel GEMMs),
                                         // 2 -> K = min(K, *rows_dev) (token-reduction GEMMs; one operand has zero rows up to the next k-block)
    int dbg_lbo, dbg_sbo, dbg_kstep, dbg_major, dbg_layout;   // MN-major descriptor parameters (bytes / flags), tunable for bring-up
};

static int g_dbg_lbo = 32 * BK * 4, g_dbg_sbo = 512, g_dbg_kstep = 1024, g_dbg_major = 1, g_dbg_layout = 1, g_dbg_tma_swz = 4;

// Bring-up profile (UR_TC_PROF=1): cycles each warp role spends in its waits / its work, summed over CTAs (ur_gemm_tc_prof reads them).
// 0 producer wait empty | 1 MMA wait ready | 2 MMA wait tmem_empty | 3 MMA issue | 4 splitter wait full | 5 splitter wait lo_empty |
// 6 splitter work | 7 epilogue (warp 2) wait tmem_full | 8 epilogue (warp 2) work | 9 CTA lifetime | 10 k-blocks | 11 units |
// 12 MMA instructions of a k-block (part of 3) | 13 its commits (part of 3) | 14 splitter A tile -> TMEM | 15 splitter B tile (parts of 6)
__device__ unsigned long long g_tc_prof[16];
struct ProfClock {
    long long t; bool on;
    __device__ __forceinline__ ProfClock(bool enabled) : t(0), on(enabled) { if (on) t = clock64(); }
    __device__ __forceinline__ void lap(long long& acc) { if (on) { const long long n = clock64(); acc += n - t; t = n; } }
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// round-to-nearest (ties away) to the 10-bit TF32 mantissa with two full-rate integer ops (same result as cvt.rna.tf32.f32)
__device__ __forceinline__ float to_tf32_rna(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

// Activation stages of the epilogue.  Both precisions use the fast-intrinsic swish (ex2.approx + fast divide, ~2 ulp each, i.e.
// 1e-6-class against a 1e-3 parity bound): the IEEE expf + division variant made the epilogue warps the limiter of the two
// FFN products with a fused activation in 3xTF32 mode (91 -> 59 us and 121 -> 88 us; the whole GPU suite stays green).
// The common activations get tight unrolled loops; everything else goes through one
// out-of-line call per element.  (Inlining the 5-way switch with erff/tanhf into a 32-element unrolled loop produced ~100 KB
// of SASS per kernel and made the epilogue instruction-fetch bound: 140 us instead of 40 us for the FFN GEMMs.)
__device__ __noinline__ float act_fwd_generic(float x, int act) { return act_fwd(x, act); }
__device__ __noinline__ float act_bwd_generic(float x, int act) { return act_bwd(x, act); }

template <bool kPrecise>
__device__ __forceinline__ void epilogue_act_fwd(float (&v)[32], int act) {
    if (act == ACT_SWISH) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = kPrecise ? v[i] / (1.f + expf(-v[i])) : __fdividef(v[i], 1.f + __expf(-v[i]));
    } else if (act == ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = act_fwd_generic(v[i], act);
    }
}
template <bool kPrecise>
__device__ __forceinline__ float swish_bwd(float x) {
    const float s = kPrecise ? 1.f / (1.f + expf(-x)) : __fdividef(1.f, 1.f + __expf(-x));
    return s * (1.f + x * (1.f - s));
}
template <bool kPrecise>
__device__ __forceinline__ void epilogue_act_bwd(float (&v)[32], const float4 (&z)[8], int act) {
    if (act == ACT_SWISH) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[4 * i] *= swish_bwd<kPrecise>(z[i].x); v[4 * i + 1] *= swish_bwd<kPrecise>(z[i].y);
            v[4 * i + 2] *= swish_bwd<kPrecise>(z[i].z); v[4 * i + 3] *= swish_bwd<kPrecise>(z[i].w);
        }
    } else if (act == ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[4 * i] = z[i].x > 0.f ? v[4 * i] : 0.f; v[4 * i + 1] = z[i].y > 0.f ? v[4 * i + 1] : 0.f;
            v[4 * i + 2] = z[i].z > 0.f ? v[4 * i + 2] : 0.f; v[4 * i + 3] = z[i].w > 0.f ? v[4 * i + 3] : 0.f;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[4 * i] *= act_bwd_generic(z[i].x, act); v[4 * i + 1] *= act_bwd_generic(z[i].y, act);
            v[4 * i + 2] *= act_bwd_generic(z[i].z, act); v[4 * i + 3] *= act_bwd_generic(z[i].w, act);
        }
    }
}

constexpr int EPI_WARPS = 8;                  // two warps per TMEM lane block, interleaved over the 32-column chunks
constexpr int EPI_TILE_FLOATS = 32 * 32;      // per epilogue warp and buffer: 32 rows x 32 columns, 128-byte swizzled (TMA store source)

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}

// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue,
// warps 10..13 (kSplit only) = operand splitter.  Two TMEM accumulators (2 x NC columns): the epilogue of unit i overlaps the
// main loop of unit i+1.
//
// kSplit ("3xTF32"): every staged fp32 tile x is rewritten in shared memory as hi = tf32(x) (in place) and lo = tf32(x - hi)
// (second buffer, same swizzled layout -> the split is purely elementwise on the raw tile bytes), and each k-step issues
// lo*hi + hi*lo + hi*hi into the fp32 accumulator: the dropped terms are O(2^-22), i.e. fp32-class results from the tensor pipe.
template <bool kAmn, bool kBmn, bool kSplit>
__global__ void __launch_bounds__(kSplit ? 480 : 320, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                       const __grid_constant__ CUtensorMap tmB,
                                                                       const __grid_constant__ CUtensorMap tmC,
                                                                       const __grid_constant__ CUtensorMap tmP,
                                                                       const __grid_constant__ CUtensorMap tmBlo, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];       // no static shared memory in this kernel: the window starts 1024-aligned
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NC = p.NC, S = p.stages;
    const uint32_t a_bytes = BM * BK * 4;                 // 16 KB
    const uint32_t b_bytes = (uint32_t)NC * BK * 4;
    const uint32_t raw_bytes = a_bytes + b_bytes;
    const uint32_t stage_bytes = raw_bytes;               // raw ring: S stages of [A | B]
    const int SL = kSplit ? p.lo_stages : 1;              // lo ring (3xTF32): SL stages of [A_lo | B_lo], SL <= S: the TMA prefetch
                                                          // runs S deep (HBM latency), the split only SL deep (just ahead of the MMA)
    uint8_t* tiles = smem;
    const bool a_tmem = kSplit && !kAmn && p.a_tmem;
    const uint32_t lo_bytes = a_tmem ? b_bytes : raw_bytes;       // one lo-ring stage
    uint8_t* lo_tiles = smem + (size_t)S * stage_bytes;
    float* epi_stage = reinterpret_cast<float*>(lo_tiles + (kSplit ? (size_t)SL * lo_bytes : 0));   // [8 warps][epi_bufs][32 x 32] swizzled
    float* colsum_sm = epi_stage + EPI_WARPS * p.epi_bufs * EPI_TILE_FLOATS;          // [N] when p.colsum
    uint64_t* full = reinterpret_cast<uint64_t*>(colsum_sm + (p.colsum ? p.N : 0));
    uint64_t* empty = full + S;
    uint64_t* ready = empty + S;                   // kSplit: split finished (MMA waits on this instead of `full`)
    uint64_t* lo_empty = ready + S;                // [SL] kSplit: the MMAs that read lo stage j have completed
    uint64_t* tmem_full = lo_empty + SL;           // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint64_t* turn = tmem_empty + 2;               // [1 of 2] dual issuers: phase n completes when the MMAs of the CTA's n-th k-block are issued
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(turn + 2);
    const bool dual = kSplit && p.dual_mma;        // two MMA-issuing warps (1 and mma_b_warp) alternate over the k-blocks
    const int mma_b_warp = dual ? 2 + EPI_WARPS + p.split_warps : -1;
    int KB_all = (p.K + BK - 1) / BK;
    int M_eff = p.M, m_stripes = p.m_stripes, kb_per_split = p.kb_per_split, total_units = p.total_units;
    if (p.rows_dev) {                      // packed sequences: the live token count is only known on the device
        const int n = *p.rows_dev;
        if (p.rows_dim == 1) {
            M_eff = min(p.M, n);
            m_stripes = (M_eff + BM - 1) / BM;
            total_units = m_stripes * p.n_chunks * p.splits;
        } else {
            KB_all = (min(p.K, n) + BK - 1) / BK;
            kb_per_split = (KB_all + p.splits - 1) / p.splits;
        }
    }
    const uint32_t tmem_cols = a_tmem ? 512 : 2 * NC;       // A in tensor memory: columns 256 + 64 j hold [hi | lo] of lo-ring stage j
    constexpr uint32_t A_TMEM_COL = 256;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
        if (kSplit && p.b_lo_tma) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBlo) : "memory");
        if (p.preact) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmP) : "memory");
        for (int s = 0; s < S; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); mbar_init(ready + s, kSplit ? p.split_warps : 1); }
        for (int s = 0; s < SL; ++s) mbar_init(lo_empty + s, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(tmem_full + i, dual ? 2 : 1); mbar_init(tmem_empty + i, EPI_WARPS); mbar_init(turn + i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (p.colsum)
        for (int c = threadIdx.x; c < p.N; c += blockDim.x) colsum_sm[c] = 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // work unit u -> (column chunk, row stripe, k split); chunks of one A stripe are adjacent so they share it through L2
    auto decode = [&](int u, int& m0, int& n0, int& kb_begin, int& KB) {
        const int chunk = u % p.n_chunks;
        const int rest = u / p.n_chunks;
        m0 = (rest % m_stripes) * BM;
        n0 = chunk * NC;
        kb_begin = (rest / m_stripes) * kb_per_split;
        KB = min(KB_all - kb_begin, kb_per_split);          // <= 0: nothing to do for this unit (every role skips it)
    };

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            int s = 0;
            uint32_t ph = 1;                         // parity of the previous round of stage s (first round: no wait)
            ProfClock pc(p.prof != 0);
            long long w_empty = 0, w_rest = 0;
            const long long t_begin = pc.t;
            for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
                int m0, n0, kb_begin, KB;
                decode(u, m0, n0, kb_begin, KB);
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    pc.lap(w_rest);
                    if (it >= (uint32_t)S) mbar_wait(empty + s, ph);
                    pc.lap(w_empty);
                    uint8_t* sa = tiles + (size_t)s * stage_bytes;
                    uint8_t* sb = sa + a_bytes;
                    mbar_expect_tx(full + s, raw_bytes + (p.b_lo_tma ? b_bytes : 0u));
                    const int kg = (kb_begin + kb) * BK;                                     // global k offset
                    if (!kAmn) {
                        tma_load_2d(sa, &tmA, full + s, kg, m0);                            // box [32 k] x [128 rows]
                    } else {                                                                 // box [32 m] x [32 k-rows], 4 boxes
                        for (int mb = 0; mb < BM / 32; ++mb)
                            tma_load_2d(sa + (size_t)mb * 32 * BK * 4, &tmA, full + s, m0 + mb * 32, kg);
                    }
                    if (!kBmn) {
                        for (int nb = 0; nb < NC / 128; ++nb)
                            tma_load_2d(sb + (size_t)nb * 128 * BK * 4, &tmB, full + s, kg, n0 + nb * 128);
                    } else {
                        for (int nb = 0; nb < NC / 32; ++nb)
                            tma_load_2d(sb + (size_t)nb * 32 * BK * 4, &tmB, full + s, n0 + nb * 32, kg);
                    }
                    if (kSplit && p.b_lo_tma) {
                        // lo ring stage == raw ring stage (S == SL, both advance once per k-block): the MMA commit that frees raw stage s
                        // (`empty`) is issued together with the one that frees lo stage s
                        uint8_t* sbl = lo_tiles + (size_t)s * b_bytes;
                        if (!kBmn) {
                            for (int nb = 0; nb < NC / 128; ++nb)
                                tma_load_2d(sbl + (size_t)nb * 128 * BK * 4, &tmBlo, full + s, kg, n0 + nb * 128);
                        } else {
                            for (int nb = 0; nb < NC / 32; ++nb)
                                tma_load_2d(sbl + (size_t)nb * 32 * BK * 4, &tmBlo, full + s, n0 + nb * 32, kg);
                        }
                    }
                    if (++s == S) { s = 0; ph ^= 1; }
                }
            }
            if (p.prof) {
                atomicAdd(g_tc_prof + 0, (unsigned long long)w_empty);
                atomicAdd(g_tc_prof + 9, (unsigned long long)(clock64() - t_begin));
                atomicAdd(g_tc_prof + 10, (unsigned long long)it);
            }
        }
    } else if (warp == 1 || warp == mma_b_warp) {
        if (lane == 0) {
            // MMA issuer(s): one elected thread per issuing warp.  Everything that does not change per instruction is hoisted: the issue cadence of this one
            // thread (dependent integer instructions at ~4 cycles each) bounded the kernel at one MMA per 130-400 cycles while the
            // tensor pipe needs ~90 per 128x128x8 MMA (ncu r2: tensor pipe 15-27% active).  Descriptors are 64-bit values whose low
            // 14 bits hold (shared address >> 4): stage / k-step / lo-ring moves are plain additions on that field.
            // Two issuers (3xTF32): warp 1 takes the even k-blocks of the CTA's sequence, warp 14 the odd ones.  The issue of a k-block's
            // MMAs blocks at the tensor pipe's pace (~67 cycles per 128x128x8), and the ~570 cycles one thread spends per k-block
            // outside of it (barrier wait, fence, two commits: profiles/gemm_roles.py) left the pipe idle 40% of the time; with two
            // threads that overhead runs while the other one's MMAs execute.  MMAs execute in issue order; the `turn` barrier keeps the
            // two threads' k-blocks in sequence.
            const int my = warp == 1 ? 0 : 1;
            uint32_t it = 0, lt = 0;
            const uint32_t idesc = make_idesc(NC, kAmn ? p.dbg_major : 0, kBmn ? p.dbg_major : 0);
            const uint32_t tiles_u32 = smem_u32(tiles), lo_u32 = smem_u32(lo_tiles);
            const uint64_t da0 = !kAmn ? make_desc(tiles_u32, 16, 1024) : make_desc(tiles_u32, p.dbg_lbo, p.dbg_sbo, p.dbg_layout);
            const uint64_t db0 = !kBmn ? make_desc(tiles_u32 + a_bytes, 16, 1024)
                                       : make_desc(tiles_u32 + a_bytes, p.dbg_lbo, p.dbg_sbo, p.dbg_layout);
            const uint64_t ka = (uint64_t)((!kAmn ? 32u : (uint32_t)p.dbg_kstep) >> 4);      // descriptor step per k8 (16-byte units)
            const uint64_t kb_step = (uint64_t)((!kBmn ? 32u : (uint32_t)p.dbg_kstep) >> 4);
            const uint64_t stage_step = (uint64_t)(stage_bytes >> 4), lo_step = (uint64_t)(raw_bytes >> 4);
            const uint64_t lo_base = (uint64_t)((lo_u32 - tiles_u32) >> 4);                 // lo ring sits above the raw ring
            const uint64_t db0_lo = !kBmn ? make_desc(lo_u32, 16, 1024) : make_desc(lo_u32, p.dbg_lbo, p.dbg_sbo, p.dbg_layout);   // a_tmem: B_lo ring
            const bool use_split = kSplit && !(p.dbg_skip & 16);
            int s = 0, sl = 0;                       // raw / lo ring positions and their phase bits, advanced without divisions
            uint32_t ph = 0;
            ProfClock pc(p.prof != 0);
            long long w_ready = 0, w_tmem = 0, w_issue = 0, w_mma = 0, w_commit = 0;
            for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
                int m0, n0, kb_begin, KB;
                decode(u, m0, n0, kb_begin, KB);
                if (KB <= 0) continue;
                const uint32_t ab = lt & 1;
                pc.lap(w_issue);
                mbar_wait(tmem_empty + ab, ((lt >> 1) & 1) ^ 1);       // passes immediately for the first use of each buffer
                pc.lap(w_tmem);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_base + ab * (uint32_t)NC;
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    if (dual && (int)(it & 1) != my) {       // the other issuer's k-block: only advance the ring positions
                        if (++sl == SL) sl = 0;
                        if (++s == S) { s = 0; ph ^= 1; }
                        continue;
                    }
                    pc.lap(w_issue);
                    mbar_wait((kSplit ? ready : full) + s, ph);
                    // strict alternation: k-block `it` is issued after k-block `it - 1` (by the other warp) -- the accumulation order,
                    // and with it every result bit, is the single issuer's
                    if (dual && it > 0) mbar_wait(turn, (it - 1) & 1);
                    pc.lap(w_ready);
                    if (!(p.dbg_skip & 32)) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    pc.lap(w_issue);
                    uint64_t da = da0 + (uint64_t)s * stage_step;
                    uint64_t db = db0 + (uint64_t)s * stage_step;
                    if (use_split && a_tmem) {
                        const uint32_t ta = tmem_base + A_TMEM_COL + (uint32_t)sl * 64u;          // hi columns; lo at +32
                        uint64_t dbl = db0_lo + (uint64_t)sl * (uint64_t)(b_bytes >> 4);
#pragma unroll
                        for (int k8 = 0; k8 < BK / 8; ++k8) {
                            umma_tf32_ts(tacc, ta + 32u + 8u * k8, db, idesc, (kb | k8) != 0);   // a_lo * b_hi
                            umma_tf32_ts(tacc, ta + 8u * k8, dbl, idesc, 1);                      // a_hi * b_lo
                            umma_tf32_ts(tacc, ta + 8u * k8, db, idesc, 1);                       // a_hi * b_hi
                            db += kb_step; dbl += kb_step;
                        }
                    } else if (use_split) {
                        // same layout in the lo ring: hi stage s -> lo stage sl
                        const uint64_t lo_off = lo_base + (uint64_t)sl * lo_step - (uint64_t)s * stage_step;
#pragma unroll
                        for (int k8 = 0; k8 < BK / 8; ++k8) {
                            umma_tf32(tacc, da + lo_off, db, idesc, (kb | k8) != 0);          // a_lo * b_hi
                            umma_tf32(tacc, da, db + lo_off, idesc, 1);                       // a_hi * b_lo
                            umma_tf32(tacc, da, db, idesc, 1);                                // a_hi * b_hi
                            da += ka; db += kb_step;
                        }
                    } else {
#pragma unroll
                        for (int k8 = 0; k8 < BK / 8; ++k8) {
                            umma_tf32(tacc, da, db, idesc, (kb | k8) != 0);
                            da += ka; db += kb_step;
                        }
                    }
                    if (dual) mbar_arrive(turn);
                    pc.lap(w_mma);
                    umma_commit(empty + s);          // frees this smem stage once the MMAs above have read it
                    if (kSplit) { if (!(p.dbg_skip & 64)) umma_commit(lo_empty + sl); if (++sl == SL) sl = 0; }
                    pc.lap(w_commit);
                    if (++s == S) { s = 0; ph ^= 1; }
                }
                umma_commit(tmem_full + ab);         // accumulator complete
                ++lt;
            }
            if (p.prof) {
                pc.lap(w_issue);
                atomicAdd(g_tc_prof + 1, (unsigned long long)w_ready);
                atomicAdd(g_tc_prof + 2, (unsigned long long)w_tmem);
                atomicAdd(g_tc_prof + 3, (unsigned long long)(w_issue + w_mma + w_commit));
                atomicAdd(g_tc_prof + 12, (unsigned long long)w_mma);
                atomicAdd(g_tc_prof + 13, (unsigned long long)w_commit);
                if (my == 0) atomicAdd(g_tc_prof + 11, (unsigned long long)lt);
            }
        }
    } else if (warp < 2 + EPI_WARPS) {
        // ---------------- epilogue: warps 2..9, TMEM lane block = warp % 4, two warps per block ----------------
        // Each warp owns 32 rows: tcgen05.ld a 32x32 chunk (lane = row), bias / activation / act' in registers, write the chunk
        // into its own 128B-swizzled 4 KB buffer and hand it to the TMA engine (store, or reduce-add for the accumulate modes):
        // no global store instructions, no read-add-store round trip, two buffers per warp so the next chunk overlaps the copy.
        const int lb = warp & 3;                                   // lanes [32*lb, 32*lb+32)
        const int half = (warp - 2) >> 2;                          // which of the two warps of this lane block
        const int nbuf = p.epi_bufs;
        float* bufs = epi_stage + (size_t)(warp - 2) * nbuf * EPI_TILE_FLOATS;
        uint32_t lt = 0, nstore = 0;
        auto emit = [&](const CUtensorMap* map, const float* v, bool reduce, int col, int row0) -> const float* {
            float* buf = bufs + (nbuf == 2 ? (nstore & 1) : 0) * EPI_TILE_FLOATS;
            if (nstore >= (uint32_t)nbuf) {          // the copy engine must have finished reading this buffer
                if (lane == 0) {
                    if (nbuf == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
                __syncwarp();
            }
            if (!(p.dbg_skip & 4)) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                *reinterpret_cast<float4*>(buf + lane * 32 + ((i ^ (lane & 7)) << 2)) =
                    make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            __syncwarp();
            if (lane == 0 && !(p.dbg_skip & 1)) {
                if (reduce) tma_reduce_add_2d(map, buf, col, row0);
                else tma_store_2d(map, buf, col, row0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            ++nstore;
            return buf;
        };
        ProfClock epc(p.prof != 0 && warp == 2 && lane == 0);
        long long e_wait = 0, e_work = 0;
        for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
            int m0, n0, kb_begin, KB;
            decode(u, m0, n0, kb_begin, KB);
            if (KB <= 0) continue;
            const int row0 = m0 + lb * 32;
            const int64_t grow = min(row0 + lane, p.M - 1);       // clamp: out-of-range rows are clipped by the TMA store
            float4 dnext[8];
            auto load_dact = [&](int col) {
#pragma unroll
                for (int i = 0; i < 8; ++i) dnext[i] = __ldg(reinterpret_cast<const float4*>(p.dact + grow * p.ldd + col) + i);
            };
            if (p.dact) load_dact(n0 + half * 32);
            const uint32_t ab = lt & 1;
            epc.lap(e_work);
            mbar_wait(tmem_full + ab, (lt >> 1) & 1);
            epc.lap(e_wait);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tacc = tmem_base + ab * (uint32_t)NC + ((uint32_t)(lb * 32) << 16);
            for (int c0 = half * 32; c0 < NC; c0 += 64) {
                float v[32];
                tmem_ld_x32(tacc + (uint32_t)c0, v);
                if (c0 + 64 >= NC) {          // this warp's last read of the accumulator: hand it back to the MMA warp
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tmem_empty + ab);
                }
                const int ncol = n0 + c0;
                if (p.bias && kb_begin == 0 && !(p.dbg_skip & 2)) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + ncol + i));
                        v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
                    }
                }
                if (p.preact) emit(&tmP, v, false, ncol, row0);
                if (p.dact) {
                    epilogue_act_bwd<false>(v, dnext, p.act);
                    if (c0 + 64 < NC) load_dact(ncol + 64);
                } else if (p.act != ACT_NONE) {
                    epilogue_act_fwd<false>(v, p.act);
                }
                const float* buf = emit(&tmC, v, p.accumulate != 0, ncol, row0);
                if (p.colsum) {
                    // lane c sums column c of the staged chunk over this warp's (valid) rows: conflict-free transposed read
                    const int nrows = min(32, M_eff - row0);
                    float cs = 0.f;
                    for (int r = 0; r < nrows; ++r) cs += buf[r * 32 + ((((lane >> 2) ^ (r & 7)) << 2) | (lane & 3))];
                    atomicAdd(colsum_sm + ncol + lane, cs);
                }
            }
            ++lt;
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
        if (epc.on) {
            epc.lap(e_work);
            atomicAdd(g_tc_prof + 7, (unsigned long long)e_wait);
            atomicAdd(g_tc_prof + 8, (unsigned long long)e_work);
        }
    } else if (kSplit) {
        // ---------------- splitter: warps 10..13 ----------------
        const int tid = threadIdx.x - (2 + EPI_WARPS) * 32;
        const int nsplit = p.split_warps * 32;
        const int n4 = (int)(raw_bytes >> 4);
        uint32_t it = 0;
        int s = 0, sl = 0;
        uint32_t ph = 0, lph = 1;
        ProfClock pc(p.prof != 0 && tid == 0);
        long long w_full = 0, w_lo = 0, w_work = 0, w_a = 0, w_b = 0;
        for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
            int m0, n0, kb_begin, KB;
            decode(u, m0, n0, kb_begin, KB);
            for (int kb = 0; kb < KB; ++kb, ++it) {
                pc.lap(w_work);
                mbar_wait(full + s, ph);
                pc.lap(w_full);
                if (it >= (uint32_t)SL) mbar_wait(((p.dbg_skip & 64) ? empty : lo_empty) + sl, lph);
                pc.lap(w_lo);
                float4* hi = reinterpret_cast<float4*>(tiles + (size_t)s * stage_bytes);
                float4* lo = reinterpret_cast<float4*>(lo_tiles + (size_t)sl * lo_bytes);
                if (!(p.dbg_skip & 8)) {
                    if (a_tmem) {
                        // A tile -> tensor memory: thread = row (warp w reaches lanes 32 (w % 4) ..), its 32 k values are the eight
                        // 16-byte chunks of the 128-byte swizzled row; the raw words are the hi operand (the tensor pipe truncates),
                        // lo = tf32(x - trunc(x)).  The MMAs then read A from TMEM: no shared-memory A traffic per MMA (the kernel is
                        // bound by the 128 B/clk shared-memory pipe: TMA fill + split + 3 operand reads per k-step).
                        const int r = (warp & 3) * 32 + lane;
                        const uint8_t* arow = reinterpret_cast<const uint8_t*>(hi) + r * 128;
                        const uint32_t ta = tmem_base + A_TMEM_COL + (uint32_t)sl * 64u + ((uint32_t)((warp & 3) * 32) << 16);
                        float4 x[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) x[c] = *reinterpret_cast<const float4*>(arow + ((c ^ (r & 7)) << 4));
                        tmem_st_x32(ta, x);
                        // (+0x1000 then the pipe's own truncation of the low 13 bits = round-to-nearest: no second mask needed)
                        auto lo1 = [](float v) { return __uint_as_float(__float_as_uint(v - __uint_as_float(__float_as_uint(v) & 0xffffe000u)) + 0x1000u); };
#pragma unroll
                        for (int c = 0; c < 8; ++c) x[c] = make_float4(lo1(x[c].x), lo1(x[c].y), lo1(x[c].z), lo1(x[c].w));
                        tmem_st_x32(ta + 32u, x);
                        pc.lap(w_a);
                        // B tile: elementwise on the raw bytes (any layout), lo into the B-only lo ring -- unless the producer already
                        // fetched it from the weight's lo plane
                        const float4* bh = reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(hi) + a_bytes);
                        const int n4b = p.b_lo_tma ? 0 : (int)(b_bytes >> 4);
                        int i = tid;
                        for (; i + 3 * nsplit < n4b; i += 4 * nsplit) {
                            float4 y[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) y[j] = bh[i + j * nsplit];
#pragma unroll
                            for (int j = 0; j < 4; ++j) lo[i + j * nsplit] = make_float4(lo1(y[j].x), lo1(y[j].y), lo1(y[j].z), lo1(y[j].w));
                        }
                        for (; i < n4b; i += nsplit) { const float4 y = bh[i]; lo[i] = make_float4(lo1(y.x), lo1(y.y), lo1(y.z), lo1(y.w)); }
                        pc.lap(w_b);
                        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    } else if (p.split_trunc) {
                        // The tensor pipe ignores the low 13 mantissa bits of a kind::tf32 operand, i.e. it multiplies trunc(x): the raw
                        // tile already IS the hi operand, only lo = tf32(x - trunc(x)) has to be produced (x - trunc(x) is exact in
                        // fp32).  Halves the splitter's shared-memory writes -- the split was the limiter of the 3xTF32 mode
                        // (ablation r2: 59 us -> 35 us without it on the 51200 x 384 x 128 product).
                        // (one splitter warp per SM sub-partition: the loads of four chunks are issued together so their shared-memory
                        // latencies overlap instead of serialising load -> ALU -> store per chunk: 739 -> 640 us per layer)
                        auto lo_of = [](const float4 x) {
                            float4 l;
                            l.x = to_tf32_rna(x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u));
                            l.y = to_tf32_rna(x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u));
                            l.z = to_tf32_rna(x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u));
                            l.w = to_tf32_rna(x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u));
                            return l;
                        };
                        int i = tid;
                        for (; i + 3 * nsplit < n4; i += 4 * nsplit) {
                            float4 x[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) x[j] = hi[i + j * nsplit];
#pragma unroll
                            for (int j = 0; j < 4; ++j) lo[i + j * nsplit] = lo_of(x[j]);
                        }
                        for (; i < n4; i += nsplit) lo[i] = lo_of(hi[i]);
                    } else {
#pragma unroll 4
                        for (int i = tid; i < n4; i += nsplit) {
                            const float4 x = hi[i];
                            float4 h, l;
                            h.x = to_tf32_rna(x.x); h.y = to_tf32_rna(x.y); h.z = to_tf32_rna(x.z); h.w = to_tf32_rna(x.w);
                            l.x = to_tf32_rna(x.x - h.x); l.y = to_tf32_rna(x.y - h.y); l.z = to_tf32_rna(x.z - h.z); l.w = to_tf32_rna(x.w - h.w);
                            hi[i] = h;
                            lo[i] = l;
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to tcgen05.mma
                __syncwarp();
                if (lane == 0) mbar_arrive(ready + s);
                if (++s == S) { s = 0; ph ^= 1; }
                if (++sl == SL) { sl = 0; lph ^= 1; }
            }
        }
        if (pc.on) {
            pc.lap(w_work);
            atomicAdd(g_tc_prof + 4, (unsigned long long)w_full);
            atomicAdd(g_tc_prof + 5, (unsigned long long)w_lo);
            atomicAdd(g_tc_prof + 6, (unsigned long long)(w_work + w_a + w_b));
            atomicAdd(g_tc_prof + 14, (unsigned long long)w_a);
            atomicAdd(g_tc_prof + 15, (unsigned long long)w_b);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (p.colsum)                                        // one atomic per column per CTA
        for (int c = threadIdx.x; c < p.N; c += blockDim.x) atomicAdd(p.colsum + c, colsum_sm[c]);
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// small row-major transpose out[c][r] = in[r][c] (weights only; used to express dx = dy W as an NT product)
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, int rows, int cols, float* __restrict__ out) {
    __shared__ float t[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8)
        if (by + j < rows && bx + tx < cols) t[j][tx] = in[(int64_t)(by + j) * cols + bx + tx];
    __syncthreads();
    for (int j = ty; j < 32; j += 8)
        if (bx + j < cols && by + tx < rows) out[(int64_t)(bx + j) * rows + by + tx] = t[tx][j];
}

// lo plane of a weight buffer: lo[i] = tf32-rounded (x[i] - trunc_tf32(x[i])) -- the second operand term of the 3xTF32 product, computed
// once per forward pass for all encoder weights instead of once per staged tile in every GEMM CTA (ur_split_lo_f32)
__global__ void __launch_bounds__(256) split_lo_kernel(const float4* __restrict__ x, float4* __restrict__ lo, int64_t n4) {
    auto lo1 = [](float v) { return __uint_as_float(__float_as_uint(v - __uint_as_float(__float_as_uint(v) & 0xffffe000u)) + 0x1000u); };
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = x[i];
        lo[i] = make_float4(lo1(v.x), lo1(v.y), lo1(v.z), lo1(v.w));
    }
}

// registered (weights, lo plane) pair: a GEMM whose B operand lies inside [base, base + n) reads B's lo term from the plane
static const float* g_lo_base = nullptr;
static const float* g_lo_plane = nullptr;
static int64_t g_lo_n = 0;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && sym) fn = (EncodeTiledFn)sym;
    }
    return fn;
}

// 2-D fp32 tensor map over a row-major matrix [rows, cols] with leading dimension ld; box = [box_cols (inner), box_rows]
// Encoded tensor maps are cached per (base, shape, stride, box, swizzle): a training step re-issues the same ~30 products on the
// same buffers, and in eager mode the five cuTensorMapEncodeTiled calls per launch sat between the caller's timing event and
// the kernel.  (A descriptor only names addresses and shapes, so a cached one stays valid for as long as the key matches.)
struct MapKey {
    const void* base; int64_t rows, cols, ld; int box_cols, box_rows, swz;
    bool operator==(const MapKey& o) const {
        return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_cols == o.box_cols && box_rows == o.box_rows &&
               swz == o.swz;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.base) * 0x9E3779B97F4A7C15ull;
        auto mix = [&](uint64_t v) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); };
        mix((uint64_t)k.rows); mix((uint64_t)k.cols); mix((uint64_t)k.ld);
        mix(((uint64_t)k.box_cols << 32) | ((uint64_t)k.box_rows << 8) | (uint64_t)k.swz);
        return h;
    }
};

static bool make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
                     CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    static std::mutex mu;
    const MapKey key{base, rows, cols, ld, box_cols, box_rows, (int)swz};
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *map = it->second; return true; }
    }
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    if (enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() >= 4096) cache.clear();
    cache.emplace(key, *map);
    return true;
}

}  // namespace tc
}  // namespace ur

extern "C" {

// bring-up hook: MN-major descriptor parameters of the TN kernel (defaults are the validated values)
int ur_gemm_tc_debug_set(int lbo_bytes, int sbo_bytes, int kstep_bytes, int major, int layout_type, int tma_swizzle) {
    ur::tc::g_dbg_lbo = lbo_bytes; ur::tc::g_dbg_sbo = sbo_bytes; ur::tc::g_dbg_kstep = kstep_bytes; ur::tc::g_dbg_major = major;
    ur::tc::g_dbg_layout = layout_type; ur::tc::g_dbg_tma_swz = tma_swizzle;
    return UR_OK;
}

// bring-up hook: read (and optionally clear) the role clocks accumulated by launches made with UR_TC_PROF=1
int ur_gemm_tc_prof(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    if (out16 && cudaMemcpyFromSymbol(out16, ur::tc::g_tc_prof, 16 * sizeof(unsigned long long)) != cudaSuccess) return -1000;
    if (reset) {
        unsigned long long z[16] = {0};
        if (cudaMemcpyToSymbol(ur::tc::g_tc_prof, z, sizeof(z)) != cudaSuccess) return -1000;
    }
    return UR_OK;
}

int ur_gemm_set_lo_plane(const float* base, const float* lo_plane, int64_t n) {
    if (n < 0 || (n > 0 && (!base || !lo_plane))) return UR_ERR_BAD_ARG;
    ur::tc::g_lo_base = n ? base : nullptr; ur::tc::g_lo_plane = n ? lo_plane : nullptr; ur::tc::g_lo_n = n;
    return UR_OK;
}

int ur_split_lo_f32(const float* x, float* lo, int64_t n, void* stream) {
    if (n < 0 || (n & 3) || ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(lo)) & 15)) return UR_ERR_BAD_ARG;
    if (n == 0) return UR_OK;
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > ur::kNumSMs * 8) blocks = ur::kNumSMs * 8;
    ur::tc::split_lo_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)x, (float4*)lo, n / 4);
    UR_RETURN_LAST_ERROR();
}

int ur_transpose_f32(const float* in, int64_t rows, int64_t cols, float* out, void* stream) {
    if (rows == 0 || cols == 0) return UR_OK;
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    ur::tc::transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, (int)rows, (int)cols, out);
    UR_RETURN_LAST_ERROR();
}

// Extended tensor-core GEMM behind ur_gemm_f32 / ur_gemm_fused_f32.  precision: 1 = TF32 (operands truncated by the tensor
// pipe), 3 = 3xTF32 split (fp32-class).  Returns UR_ERR_UNSUPPORTED when the shape/layout is outside the kernel; the dispatcher
// then uses the SIMT kernel (plus separate act_bwd / colsum launches).
int ur_gemm_tc_ex(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb,
                  float* C, int64_t ldc, const float* bias, int act, float* preact, int64_t ldp, int accumulate, int precision,
                  const float* dact, int64_t ldd, float* colsum, const int32_t* rows_dev, int rows_dim, void* stream) {
    using namespace ur::tc;
    if (rows_dev && rows_dim != 1 && rows_dim != 2) return UR_ERR_BAD_ARG;
    if (precision != 1 && precision != 3) return UR_ERR_UNSUPPORTED;
    const bool split3 = precision == 3;
    const bool a_mn = transA != 0;           // A stored [K, M]
    const bool b_mn = transB == 0;           // B stored [K, N]
    if (a_mn && !b_mn) return UR_ERR_UNSUPPORTED;                           // TT is never needed by the path
    if (M < 1 || N < 128 || (N % 128) || K < BK || (K % BK)) return UR_ERR_UNSUPPORTED;
    if ((lda & 3) || (ldb & 3) || (ldc & 3) || (preact && (ldp & 3)) || (dact && (ldd & 3))) return UR_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) & 15) return UR_ERR_UNSUPPORTED;
    if (a_mn && (M % 32)) return UR_ERR_UNSUPPORTED;
    if (dact && (accumulate || preact)) return UR_ERR_UNSUPPORTED;
    // columns per work unit: two accumulators must fit the 512 TMEM columns; the split kernel keeps hi+lo tiles -> 128
    static const int env_max_nc = getenv("UR_TC_MAX_NC") ? atoi(getenv("UR_TC_MAX_NC")) : 0;         // tuning knobs (profiles/gemm_shapes.py)
    static const int env_bufs = getenv("UR_TC_EPI_BUFS") ? atoi(getenv("UR_TC_EPI_BUFS")) : 0;
    static const int env_stages = getenv("UR_TC_STAGES") ? atoi(getenv("UR_TC_STAGES")) : 0;
    static const int env_split_nc = getenv("UR_TC_SPLIT_NC") ? atoi(getenv("UR_TC_SPLIT_NC")) : 0;
    const int max_nc = split3 ? (env_split_nc ? env_split_nc : 128) : (env_max_nc ? env_max_nc : MAX_NC);
    int NC = 0;
    for (int c = max_nc; c >= 128; c -= 128)
        if (N % c == 0) { NC = c; break; }
    if (!NC) return UR_ERR_UNSUPPORTED;
    const uint32_t raw_bytes = BM * BK * 4 + (uint32_t)NC * BK * 4;
    const uint32_t stage_bytes = raw_bytes;
    // one persistent CTA per SM: 3 x 64 KB (split) / 3 x 48 KB / 5 x 32 KB of operand stages + 32 or 64 KB of epilogue store buffers
    const int epi_bufs = env_bufs ? env_bufs : (split3 ? 1 : 2);
    static const int env_lo = getenv("UR_TC_LO_STAGES") ? atoi(getenv("UR_TC_LO_STAGES")) : 0;
    static const int env_atmem = getenv("UR_TC_A_TMEM") ? atoi(getenv("UR_TC_A_TMEM")) : 1;
    static const int env_trunc = getenv("UR_TC_SPLIT_TRUNC") ? atoi(getenv("UR_TC_SPLIT_TRUNC")) : 1;
    // K-major A in 3xTF32 mode: hi/lo of the A tile go to tensor memory (needs the 256 columns next to two 128-column accumulators)
    const bool a_tmem = split3 && !a_mn && NC == 128 && env_atmem && env_trunc;
    const uint32_t lo_bytes = a_tmem ? (uint32_t)NC * BK * 4 : raw_bytes;
    const int lo_stages = split3 ? (env_lo ? env_lo : (a_tmem ? 4 : 2)) : 0;
    if (a_tmem && lo_stages > 4) return UR_ERR_UNSUPPORTED;
    int stages = (int)(((split3 ? 192 : 160) * 1024 - (size_t)lo_stages * lo_bytes) / stage_bytes);
    const int KB = (int)(K / BK);
    // split-K along the reduction (token) dimension when the output has too few tiles to fill the GPU (weight gradients)
    const int64_t m_stripes = (M + BM - 1) / BM;
    const int64_t tiles = m_stripes * (N / NC);
    int splits = 1;
    if (accumulate && !preact && act == 0 && !bias && !colsum && tiles < ur::kNumSMs) {
        splits = (int)((ur::kNumSMs + tiles - 1) / tiles);
        if (splits > KB / 8) splits = KB / 8 > 0 ? KB / 8 : 1;          // at least 8 k-blocks per unit
    }
    int kb_per_split = (KB + splits - 1) / splits;
    splits = (KB + kb_per_split - 1) / kb_per_split;
    if (stages > 6) stages = 6;
    if (env_stages && stages > env_stages) stages = env_stages;
    if (stages < 2) return UR_ERR_UNSUPPORTED;
    if (split3 && lo_stages > stages) return UR_ERR_UNSUPPORTED;
    const size_t smem = (size_t)stages * stage_bytes + (size_t)lo_stages * lo_bytes + (size_t)EPI_WARPS * epi_bufs * EPI_TILE_FLOATS * sizeof(float) + (colsum ? (size_t)N * 4 : 0) +
                        (3 * stages + lo_stages + 7) * sizeof(uint64_t) + 16;
    if (smem > 227 * 1024) return UR_ERR_UNSUPPORTED;
    CUtensorMap tmA, tmB, tmC, tmP, tmBlo;
    bool ok;
    const CUtensorMapSwizzle swz_mn = (CUtensorMapSwizzle)g_dbg_tma_swz;      // 4 = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
    if (!a_mn) ok = make_map(&tmA, A, M, K, lda, BK, BM);                      // A [M,K]: inner = k
    else ok = make_map(&tmA, A, K, M, lda, 32, BK, swz_mn);                    // A stored [K,M]: inner = m, rows = k
    if (!b_mn) ok = ok && make_map(&tmB, B, N, K, ldb, BK, 128);               // B [N,K]
    else ok = ok && make_map(&tmB, B, K, N, ldb, 32, BK, swz_mn);              // B stored [K,N]
    // B inside the registered weight buffer: its lo term comes from the lo plane (same offset, same layout) by TMA
    static const int env_blo = getenv("UR_TC_B_LO_PLANE") ? atoi(getenv("UR_TC_B_LO_PLANE")) : 1;
    const float* B_lo = nullptr;
    if (a_tmem && env_blo && stages == lo_stages && g_lo_n > 0 && B >= g_lo_base &&
        B + (b_mn ? (K - 1) * ldb + N : (N - 1) * ldb + K) <= g_lo_base + g_lo_n) B_lo = g_lo_plane + (B - g_lo_base);
    tmBlo = tmB;
    if (B_lo) {
        if (!b_mn) ok = ok && make_map(&tmBlo, B_lo, N, K, ldb, BK, 128);
        else ok = ok && make_map(&tmBlo, B_lo, K, N, ldb, 32, BK, swz_mn);
    }
    ok = ok && make_map(&tmC, C, M, N, ldc, 32, 32);                           // epilogue store boxes: 32 columns x 32 rows
    tmP = tmC;
    if (preact) ok = ok && make_map(&tmP, preact, M, N, ldp, 32, 32);
    if (!ok) return UR_ERR_UNSUPPORTED;
    Params p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K; p.NC = NC; p.stages = stages; p.C = C; p.ldc = ldc; p.bias = bias; p.act = act;
    p.preact = preact; p.ldp = ldp; p.dact = dact; p.ldd = ldd; p.colsum = colsum;
    p.accumulate = accumulate ? (splits > 1 ? 1 : 2) : 0;
    p.kb_per_split = kb_per_split; p.epi_bufs = epi_bufs; p.lo_stages = lo_stages;
    static const int env_sw = getenv("UR_TC_SPLIT_WARPS") ? atoi(getenv("UR_TC_SPLIT_WARPS")) : 0;
    p.split_trunc = env_trunc;
    p.a_tmem = a_tmem ? 1 : 0;
    p.b_lo_tma = B_lo ? 1 : 0;
    static const int env_dual = getenv("UR_TC_DUAL_MMA") ? atoi(getenv("UR_TC_DUAL_MMA")) : 1;
    p.dual_mma = (split3 && env_dual) ? 1 : 0;
    p.split_warps = 4;      // (8 splitter warps measured no faster: the split is not the limiter, profiles/r02/gemm_split_warps.txt)
    (void)env_sw;
    static const int env_prof = getenv("UR_TC_PROF") ? atoi(getenv("UR_TC_PROF")) : 0;
    p.prof = env_prof;
    static const int env_skip = getenv("UR_TC_SKIP") ? atoi(getenv("UR_TC_SKIP")) : 0;
    p.dbg_skip = env_skip;
    p.n_chunks = (int)(N / NC); p.m_stripes = (int)m_stripes; p.total_units = (int)(tiles * splits); p.splits = splits;
    p.rows_dev = rows_dev; p.rows_dim = rows_dev ? rows_dim : 0;
    p.dbg_lbo = g_dbg_lbo; p.dbg_sbo = g_dbg_sbo; p.dbg_kstep = g_dbg_kstep; p.dbg_major = g_dbg_major; p.dbg_layout = g_dbg_layout;
    const unsigned grid = (unsigned)(p.total_units < ur::kNumSMs ? p.total_units : ur::kNumSMs);
    cudaStream_t st = (cudaStream_t)stream;
#define UR_TC_LAUNCH(AMN, BMN, SPL)                                                                                         \
    do {                                                                                                                    \
        static size_t smem_set[64] = {0};                  /* per instantiation and device: the attribute is sticky */           \
        int dev_id = 0;                                                                                                     \
        cudaGetDevice(&dev_id);                                                                                             \
        if (dev_id < 0 || dev_id >= 64 || smem > smem_set[dev_id]) {                                                        \
            cudaFuncSetAttribute(gemm_tc_kernel<AMN, BMN, SPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
            if (dev_id >= 0 && dev_id < 64) smem_set[dev_id] = smem;                                                        \
        }                                                                                                                   \
        gemm_tc_kernel<AMN, BMN, SPL><<<grid, SPL ? (2 + EPI_WARPS + p.split_warps + p.dual_mma) * 32 : 320, smem, st>>>(tmA, tmB, tmC, tmP, tmBlo, p);                                    \
    } while (0)
    if (split3) {
        if (a_mn) UR_TC_LAUNCH(true, true, true);
        else if (b_mn) UR_TC_LAUNCH(false, true, true);
        else UR_TC_LAUNCH(false, false, true);
    } else {
        if (a_mn) UR_TC_LAUNCH(true, true, false);
        else if (b_mn) UR_TC_LAUNCH(false, true, false);
        else UR_TC_LAUNCH(false, false, false);
    }
#undef UR_TC_LAUNCH
    UR_RETURN_LAST_ERROR();
}

int ur_gemm_tc_f32(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb,
                   float* C, int64_t ldc, const float* bias, int act, float* preact, int64_t ldp, int accumulate, int precision,
                   void* stream) {
    return ur_gemm_tc_ex(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, act, preact, ldp, accumulate, precision, nullptr, 0,
                         nullptr, nullptr, 0, stream);
}

}  // extern "C"
