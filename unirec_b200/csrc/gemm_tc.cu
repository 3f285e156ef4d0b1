// Tensor-core GEMM for the encoder's dense layers (K4/K7): tcgen05.mma (kind::tf32, fp32 operands read in place, fp32
// accumulate in TMEM), operands staged by TMA (cp.async.bulk.tensor, 128B swizzle) through an mbarrier ring.
//
//   C[M, N] (+)= act(op(A) * op(B) + bias)
//
// Two operand layouts, selected by the caller's transposes:
//   NT  (forward, y = x W^T)   : A [M,K] and B [N,K] both K-major          -> kTN = false
//   TN  (weight grads, dW = dy^T x): A stored [K,M], B stored [K,N], both MN-major -> kTN = true
// (NN, dx = dy W, is routed through NT with a transposed copy of the small weight matrix, see ur_transpose_f32.)
//
// One CTA owns a 128-row stripe of C and up to 512 columns (the whole TMEM: 128 lanes x 512 fp32 columns), so the A stripe
// is read from HBM once.  Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld -> bias/activation -> shared-memory transpose -> coalesced 128-byte row stores).
#include <cuda.h>
#include "common.cuh"

namespace ur {
namespace tc {

constexpr int BM = 128;          // rows of C per CTA (UMMA M)
constexpr int BK = 32;           // fp32 elements per k-block = 128 bytes = one swizzle span
constexpr int MAX_NC = 256;      // columns of C per CTA: 256 TMEM columns and <= 100 KB smem -> 2 CTAs per SM, so one CTA's
                                 // epilogue overlaps the other's TMA/MMA main loop (the A stripe is re-read from L2 per chunk)
constexpr int NUM_THREADS = 192;

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
__device__ __forceinline__ uint32_t make_idesc(int n, int mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)mn_major << 15) | ((uint32_t)mn_major << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

struct Params {
    int M, N, K;            // logical GEMM sizes (K = reduction length)
    int NC;                 // columns handled by one CTA (multiple of 128, <= 512)
    int stages;
    float* C; int64_t ldc;
    const float* bias; int act;
    float* preact; int64_t ldp;
    int accumulate;         // 0 store, 1 atomic add (split-K partials), 2 read-add-store
    int tmem_cols;
    int kb_per_split;       // k-blocks handled by one CTA along grid.z
    int dbg_lbo, dbg_sbo, dbg_kstep, dbg_major, dbg_layout;   // MN-major descriptor parameters (bytes / flags), tunable for bring-up
};

static int g_dbg_lbo = 32 * BK * 4, g_dbg_sbo = 512, g_dbg_kstep = 1024, g_dbg_major = 1, g_dbg_layout = 1, g_dbg_tma_swz = 4;

template <bool kTN>
__global__ void __launch_bounds__(NUM_THREADS, 2) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmB, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NC = p.NC, S = p.stages;
    const uint32_t a_bytes = BM * BK * 4;                 // 16 KB
    const uint32_t b_bytes = (uint32_t)NC * BK * 4;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    uint8_t* tiles = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
    uint64_t* empty = full + S;
    uint64_t* tmem_full = empty + S;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * NC;      // column chunks of one A stripe are adjacent in launch order
    const int KB_all = (p.K + BK - 1) / BK;
    const int kb_begin = blockIdx.z * p.kb_per_split;
    const int KB = min(KB_all - kb_begin, p.kb_per_split);     // k-blocks of this CTA (>= 1 by construction)

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < S; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % S;
                if (kb >= S) mbar_wait(empty + s, ((kb / S) - 1) & 1);
                uint8_t* sa = tiles + (size_t)s * stage_bytes;
                uint8_t* sb = sa + a_bytes;
                mbar_expect_tx(full + s, stage_bytes);
                const int kg = (kb_begin + kb) * BK;                                     // global k offset
                if (!kTN) {
                    tma_load_2d(sa, &tmA, full + s, kg, m0);                            // box [32 k] x [128 rows]
                    for (int nb = 0; nb < NC / 128; ++nb)
                        tma_load_2d(sb + (size_t)nb * 128 * BK * 4, &tmB, full + s, kg, n0 + nb * 128);
                } else {
                    // MN-major: box [32 m] x [32 k-rows]; 4 boxes cover 128 m, NC/32 boxes cover the n columns
                    for (int mb = 0; mb < BM / 32; ++mb)
                        tma_load_2d(sa + (size_t)mb * 32 * BK * 4, &tmA, full + s, m0 + mb * 32, kg);
                    for (int nb = 0; nb < NC / 32; ++nb)
                        tma_load_2d(sb + (size_t)nb * 32 * BK * 4, &tmB, full + s, n0 + nb * 32, kg);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % S;
                mbar_wait(full + s, (kb / S) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(tiles + (size_t)s * stage_bytes);
                const uint32_t sb = sa + a_bytes;
#pragma unroll
                for (int k8 = 0; k8 < BK / 8; ++k8) {
                    for (int nc = 0; nc < NC; nc += 256) {
                        const int n_part = (NC - nc) < 256 ? (NC - nc) : 256;
                        uint64_t da, db;
                        if (!kTN) {
                            da = make_desc(sa + k8 * 32, 16, 1024);
                            db = make_desc(sb + (uint32_t)nc * BK * 4 + k8 * 32, 16, 1024);
                        } else {
                            // one instruction consumes 8 k-rows = one 1024-byte atom per 32-wide MN chunk
                            da = make_desc(sa + k8 * p.dbg_kstep, p.dbg_lbo, p.dbg_sbo, p.dbg_layout);
                            db = make_desc(sb + (uint32_t)nc * BK * 4 + k8 * p.dbg_kstep, p.dbg_lbo, p.dbg_sbo, p.dbg_layout);
                        }
                        umma_tf32(tmem_base + (uint32_t)nc, da, db, make_idesc(n_part, kTN ? p.dbg_major : 0), (kb | k8) != 0);
                    }
                }
                umma_commit(empty + s);          // frees this smem stage once the MMAs above have read it
            }
            umma_commit(tmem_full);              // accumulator complete
        }
    } else {
        // ---------------- epilogue: warps 2..5, TMEM lane block = warp % 4 ----------------
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int lb = warp & 3;                                   // lanes [32*lb, 32*lb+32)
        float* stage = reinterpret_cast<float*>(tiles) + (size_t)(warp - 2) * 32 * 36;   // pipeline smem is free now
        const int row = m0 + lb * 32 + lane;
        for (int c0 = 0; c0 < NC; c0 += 32) {
            float v[32];
            tmem_ld_x32(tmem_base + ((uint32_t)(lb * 32) << 16) + (uint32_t)c0, v);
            const int ncol = n0 + c0;
            if (p.bias && blockIdx.z == 0) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + ncol + i));
                    v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
                }
            }
            for (int pass = 0; pass < 2; ++pass) {
                float* out = pass == 0 ? p.preact : p.C;
                const int64_t ld = pass == 0 ? p.ldp : p.ldc;
                if (out == nullptr) continue;
                if (pass == 1 && p.act != ACT_NONE) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = act_fwd_fast(v[i], p.act);
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4*>(stage + lane * 36 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                __syncwarp();
                // coalesced: 8 lanes cover one 128-byte row segment, 4 rows per instruction
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int r = it * 4 + (lane >> 3), c = (lane & 7) * 4;
                    const int grow = m0 + lb * 32 + r;
                    if (grow < p.M) {
                        float4 x = *reinterpret_cast<const float4*>(stage + r * 36 + c);
                        float* dst = out + (int64_t)grow * ld + ncol + c;
                        if (pass == 1 && p.accumulate == 1) { red_add_v4(dst, x); continue; }
                        if (pass == 1 && p.accumulate == 2) x = f4_add(x, *reinterpret_cast<const float4*>(dst));
                        *reinterpret_cast<float4*>(dst) = x;
                    }
                }
            }
        }
        (void)row;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
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
static bool make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
                     CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
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

int ur_transpose_f32(const float* in, int64_t rows, int64_t cols, float* out, void* stream) {
    if (rows == 0 || cols == 0) return UR_OK;
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    ur::tc::transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, (int)rows, (int)cols, out);
    UR_RETURN_LAST_ERROR();
}

// Returns UR_ERR_UNSUPPORTED when the shape/layout is outside the tensor-core kernel; the dispatcher then uses the SIMT kernel.
int ur_gemm_tc_f32(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb,
                   float* C, int64_t ldc, const float* bias, int act, float* preact, int64_t ldp, int accumulate, int precision,
                   void* stream) {
    using namespace ur::tc;
    if (precision != 1) return UR_ERR_UNSUPPORTED;                       // TF32 only for now
    const bool nt = !transA && transB, tn = transA && !transB;
    if (!nt && !tn) return UR_ERR_UNSUPPORTED;
    if (M < 1 || N < 128 || (N % 128) || K < BK || (K % BK)) return UR_ERR_UNSUPPORTED;
    if ((lda & 3) || (ldb & 3) || (ldc & 3) || (preact && (ldp & 3))) return UR_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) & 15) return UR_ERR_UNSUPPORTED;
    if (tn && (M % 32)) return UR_ERR_UNSUPPORTED;
    // columns per CTA: largest multiple of 128 that divides N and fits TMEM
    int NC = 0;
    for (int c = MAX_NC; c >= 128; c -= 128)
        if (N % c == 0) { NC = c; break; }
    if (!NC) return UR_ERR_UNSUPPORTED;
    const uint32_t stage_bytes = BM * BK * 4 + (uint32_t)NC * BK * 4;
    int stages = (int)((100 * 1024) / stage_bytes);      // two CTAs per SM
    const int KB = (int)(K / BK);
    // split-K along the reduction (token) dimension when the output has too few tiles to fill the GPU (weight gradients)
    const int64_t tiles = ((M + BM - 1) / BM) * (N / NC);
    int splits = 1;
    if (accumulate && !preact && act == 0 && !bias && tiles < ur::kNumSMs) {
        splits = (int)((ur::kNumSMs + tiles - 1) / tiles);
        if (splits > KB / 8) splits = KB / 8 > 0 ? KB / 8 : 1;          // at least 8 k-blocks per CTA
    }
    int kb_per_split = (KB + splits - 1) / splits;
    splits = (KB + kb_per_split - 1) / kb_per_split;
    if (stages > 4) stages = 4;
    if (stages > kb_per_split) stages = kb_per_split;
    if (stages < 1) return UR_ERR_UNSUPPORTED;
    const size_t smem = (size_t)stages * stage_bytes + (2 * stages + 1) * sizeof(uint64_t) + 16 + 1024;
    CUtensorMap tmA, tmB;
    bool ok;
    if (nt) {   // A [M,K] (ld lda), B [N,K] (ld ldb): inner = k
        ok = make_map(&tmA, A, M, K, lda, BK, BM) && make_map(&tmB, B, N, K, ldb, BK, 128);
    } else {    // A stored [K,M] (ld lda), B stored [K,N] (ld ldb): inner = m / n, rows = k
        const CUtensorMapSwizzle swz = (CUtensorMapSwizzle)g_dbg_tma_swz;      // 4 = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
        ok = make_map(&tmA, A, K, M, lda, 32, BK, swz) && make_map(&tmB, B, K, N, ldb, 32, BK, swz);
    }
    if (!ok) return UR_ERR_UNSUPPORTED;
    Params p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K; p.NC = NC; p.stages = stages; p.C = C; p.ldc = ldc; p.bias = bias; p.act = act;
    p.preact = preact; p.ldp = ldp; p.accumulate = accumulate ? (splits > 1 ? 1 : 2) : 0;
    p.tmem_cols = NC <= 128 ? 128 : (NC <= 256 ? 256 : 512);
    p.kb_per_split = kb_per_split;
    p.dbg_lbo = g_dbg_lbo; p.dbg_sbo = g_dbg_sbo; p.dbg_kstep = g_dbg_kstep; p.dbg_major = g_dbg_major; p.dbg_layout = g_dbg_layout;
    dim3 grid((unsigned)(N / NC), (unsigned)((M + BM - 1) / BM), (unsigned)splits);
    cudaStream_t st = (cudaStream_t)stream;
    if (nt) {
        cudaFuncSetAttribute(gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        gemm_tc_kernel<false><<<grid, NUM_THREADS, smem, st>>>(tmA, tmB, p);
    } else {
        cudaFuncSetAttribute(gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        gemm_tc_kernel<true><<<grid, NUM_THREADS, smem, st>>>(tmA, tmB, p);
    }
    UR_RETURN_LAST_ERROR();
}

}  // extern "C"
