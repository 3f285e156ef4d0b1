// fp32 SIMT GEMM with fused epilogue (bias, activation, pre-activation stash, accumulate / split-K).
// C[M,N] (+)= act(op(A)[M,K] * op(B)[K,N] + bias[N]);  row-major operands with leading dimensions.
// This is the exact-fp32 path (parity bar 1e-3 is met with large margin); the tcgen05 tensor-core path for the
// encoder's QKV/FFN GEMMs lives in gemm_tc.cu and is selected by ur_gemm_f32 when the shape qualifies.
#include <stdlib.h>
#include "common.cuh"

namespace ur {

constexpr int BM = 128, BN = 128, BK = 8;

// Load a [BK x 128] k-major tile into shared memory from an operand that is either
//   KCONTIG = true : stored [rows(=m or n), K], contiguous along k  -> float4 along k, transposed into smem
//   KCONTIG = false: stored [K, rows],          contiguous along row -> float4 along the row dimension
template <bool KCONTIG>
__device__ __forceinline__ float4 tile_load(const float* __restrict__ P, int64_t ld, int row0, int nrows, int k0, int kend, int t) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (KCONTIG) {
        const int r = row0 + (t >> 1), k = k0 + (t & 1) * 4;
        if (r < nrows && k < kend) v = __ldg(reinterpret_cast<const float4*>(P + (int64_t)r * ld + k));   // K % 4 == 0
    } else {
        const int k = k0 + (t >> 5), r = row0 + (t & 31) * 4;
        if (k < kend && r < nrows) v = __ldg(reinterpret_cast<const float4*>(P + (int64_t)k * ld + r));   // rows % 4 == 0
    }
    return v;
}
template <bool KCONTIG>
__device__ __forceinline__ void tile_store(float (*S)[BM], const float4& v, int t) {
    if (KCONTIG) {
        const int r = t >> 1, k = (t & 1) * 4;
        S[k + 0][r] = v.x; S[k + 1][r] = v.y; S[k + 2][r] = v.z; S[k + 3][r] = v.w;
    } else {
        const int k = t >> 5, r = (t & 31) * 4;
        *reinterpret_cast<float4*>(&S[k][r]) = v;
    }
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_simt_kernel(int M, int N, int K, const float* __restrict__ A, int64_t lda,
                                                        const float* __restrict__ B, int64_t ldb, float* __restrict__ C, int64_t ldc,
                                                        const float* __restrict__ bias, int act, float* __restrict__ preact,
                                                        int64_t ldp, int accumulate, int k_chunk,
                                                        const int32_t* __restrict__ rows_dev, int rows_dim) {
    if (rows_dev) {                // packed sequences: device-resident token count bounds the rows (1) or the reduction (2)
        const int n = *rows_dev;
        if (rows_dim == 1) M = min(M, n); else K = min(K, n);
    }
    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Bs[2][BK][BN];
    const int t = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * k_chunk;
    const int kend = min(K, kbeg + k_chunk);
    if (m0 >= M) return;
    const int tx = t & 15, ty = t >> 4;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    // op(A)[m][k]: !TA -> A stored [M,K] (k contiguous); TA -> A stored [K,M] (m contiguous)
    // op(B)[k][n]: !TB -> B stored [K,N] (n contiguous); TB -> B stored [N,K] (k contiguous)
    float4 ra = tile_load<!TA>(A, lda, m0, M, kbeg, kend, t);
    float4 rb = tile_load<TB>(B, ldb, n0, N, kbeg, kend, t);
    tile_store<!TA>(As[0], ra, t);
    tile_store<TB>(Bs[0], rb, t);
    __syncthreads();
    int buf = 0;
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        const bool more = k0 + BK < kend;
        if (more) {
            ra = tile_load<!TA>(A, lda, m0, M, k0 + BK, kend, t);
            rb = tile_load<TB>(B, ldb, n0, N, k0 + BK, kend, t);
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
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
            tile_store<!TA>(As[buf ^ 1], ra, t);
            tile_store<TB>(Bs[buf ^ 1], rb, t);
            __syncthreads();
            buf ^= 1;
        }
    }

    const bool lead = blockIdx.z == 0;   // bias is added by the first K-split only
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int n = n0 + jh * 64 + tx * 4;
            if (n >= N) continue;           // N % 4 == 0 -> whole float4 in or out
            float4 v = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
            if (bias && lead) v = f4_add(v, __ldg(reinterpret_cast<const float4*>(bias + n)));
            if (preact) *reinterpret_cast<float4*>(preact + (int64_t)m * ldp + n) = v;
            if (act != ACT_NONE) {
                v.x = act_fwd(v.x, act); v.y = act_fwd(v.y, act); v.z = act_fwd(v.z, act); v.w = act_fwd(v.w, act);
            }
            float* c = C + (int64_t)m * ldc + n;
            if (accumulate == 1) red_add_v4(c, v);                      // split-K partials
            else if (accumulate == 2) *reinterpret_cast<float4*>(c) = f4_add(*reinterpret_cast<float4*>(c), v);
            else *reinterpret_cast<float4*>(c) = v;
        }
    }
}

// Small-problem variant: 32 x 32 output tiles, 64 threads, 4 x 4 per thread.  The B-row GEMMs of the trimmed last encoder layer
// (M = batch size, N, K in {128, 512}) give the 128 x 128-tile kernels (SIMT or tcgen05) only 4-32 CTAs and 13-30 us of pure latency
// each; with 32 x 32 tiles the same products spread over 128-512 small CTAs.  Exact fp32 FMA, same epilogue options.
constexpr int SBM = 32, SBN = 32, SBK = 8;

template <bool KCONTIG>
__device__ __forceinline__ float4 stile_load(const float* __restrict__ P, int64_t ld, int row0, int nrows, int k0, int kend, int t) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (KCONTIG) {
        const int r = row0 + (t >> 1), k = k0 + (t & 1) * 4;
        if (r < nrows && k < kend) v = __ldg(reinterpret_cast<const float4*>(P + (int64_t)r * ld + k));
    } else {
        const int k = k0 + (t >> 3), r = row0 + (t & 7) * 4;
        if (k < kend && r < nrows) v = __ldg(reinterpret_cast<const float4*>(P + (int64_t)k * ld + r));
    }
    return v;
}
template <bool KCONTIG>
__device__ __forceinline__ void stile_store(float (*S)[SBM], const float4& v, int t) {
    if (KCONTIG) {
        const int r = t >> 1, k = (t & 1) * 4;
        S[k + 0][r] = v.x; S[k + 1][r] = v.y; S[k + 2][r] = v.z; S[k + 3][r] = v.w;
    } else {
        const int k = t >> 3, r = (t & 7) * 4;
        *reinterpret_cast<float4*>(&S[k][r]) = v;
    }
}

constexpr int SG = 4;      // k-groups per CTA: 4 x 64 threads work on the same 32 x 32 tile, each on a quarter of the k range

// (The first version ran one 64-thread group per CTA: a K = 128..1024 loop of 8-deep stages with ONE global load in flight per
// thread made every B-row product of the trimmed last layer 14-20 us of exposed load latency -- 208 us of a 1.85 ms step, ncu r2.
// Four groups keep four stage loads in flight per tile and cut the serial chain to a quarter; partials are summed in shared
// memory in group order, so the result does not depend on timing.)
template <bool TA, bool TB>
__global__ void __launch_bounds__(64 * SG) gemm_simt_small_kernel(int M, int N, int K, const float* __restrict__ A, int64_t lda,
                                                                  const float* __restrict__ B, int64_t ldb, float* __restrict__ C,
                                                                  int64_t ldc, const float* __restrict__ bias, int act,
                                                                  float* __restrict__ preact, int64_t ldp, int accumulate, int k_chunk) {
    __shared__ __align__(16) float As[SG][2][SBK][SBM];
    __shared__ __align__(16) float Bs[SG][2][SBK][SBN];
    __shared__ __align__(16) float red[SG - 1][SBM][SBN];
    const int g = threadIdx.x >> 6, t = threadIdx.x & 63;
    const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
    const int cbeg = blockIdx.z * k_chunk, cend = min(K, cbeg + k_chunk);                 // this CTA's k range (split-K over blockIdx.z)
    const int kq = ((cend - cbeg + SG - 1) / SG + SBK - 1) / SBK * SBK;                   // ... and this group's quarter of it
    const int kbeg = min(cend, cbeg + g * kq), kend = min(cend, kbeg + kq);
    const int tx = t & 7, ty = t >> 3;
    auto group_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + g) : "memory"); };
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    if (kbeg < kend) {
        float4 ra = stile_load<!TA>(A, lda, m0, M, kbeg, kend, t);
        float4 rb = stile_load<TB>(B, ldb, n0, N, kbeg, kend, t);
        stile_store<!TA>(As[g][0], ra, t);
        stile_store<TB>(Bs[g][0], rb, t);
        group_sync();
        int buf = 0;
        for (int k0 = kbeg; k0 < kend; k0 += SBK) {
            const bool more = k0 + SBK < kend;
            if (more) {
                ra = stile_load<!TA>(A, lda, m0, M, k0 + SBK, kend, t);
                rb = stile_load<TB>(B, ldb, n0, N, k0 + SBK, kend, t);
            }
#pragma unroll
            for (int kk = 0; kk < SBK; ++kk) {
                const float4 a4 = *reinterpret_cast<const float4*>(&As[g][buf][kk][ty * 4]);
                const float4 b4 = *reinterpret_cast<const float4*>(&Bs[g][buf][kk][tx * 4]);
                const float a[4] = {a4.x, a4.y, a4.z, a4.w};
                const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            if (more) {
                stile_store<!TA>(As[g][buf ^ 1], ra, t);
                stile_store<TB>(Bs[g][buf ^ 1], rb, t);
                group_sync();
                buf ^= 1;
            }
        }
    }
    if (g > 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(&red[g - 1][ty * 4 + i][tx * 4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
    __syncthreads();
    if (g > 0) return;
#pragma unroll
    for (int q = 0; q < SG - 1; ++q)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 r = *reinterpret_cast<const float4*>(&red[q][ty * 4 + i][tx * 4]);
            acc[i][0] += r.x; acc[i][1] += r.y; acc[i][2] += r.z; acc[i][3] += r.w;
        }
    const bool lead = blockIdx.z == 0;
    const int n = n0 + tx * 4;
    if (n >= N) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
        float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        if (bias && lead) v = f4_add(v, __ldg(reinterpret_cast<const float4*>(bias + n)));
        if (preact) *reinterpret_cast<float4*>(preact + (int64_t)m * ldp + n) = v;
        if (act != ACT_NONE) {
            v.x = act_fwd(v.x, act); v.y = act_fwd(v.y, act); v.z = act_fwd(v.z, act); v.w = act_fwd(v.w, act);
        }
        float* c = C + (int64_t)m * ldc + n;
        if (accumulate == 1) red_add_v4(c, v);
        else if (accumulate == 2) *reinterpret_cast<float4*>(c) = f4_add(*reinterpret_cast<float4*>(c), v);
        else *reinterpret_cast<float4*>(c) = v;
    }
}

// dX *= act'(preact)   (backward of the fused activation epilogue)
__global__ void __launch_bounds__(256) act_bwd_kernel(float4* __restrict__ dY, const float4* __restrict__ pre, int64_t n4, int act) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 g = dY[i];
        const float4 x = pre[i];
        g.x *= act_bwd(x.x, act); g.y *= act_bwd(x.y, act); g.z *= act_bwd(x.z, act); g.w *= act_bwd(x.w, act);
        dY[i] = g;
    }
}

// out[n] += sum_m X[m, n]   (bias gradients).  grid.x tiles columns by 128 (float4 x 32 lanes), grid.y splits rows.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, int64_t ldx, int64_t M, int N, float* __restrict__ out,
                                                     const int32_t* __restrict__ rows_dev) {
    __shared__ float4 sh[8][32];
    if (rows_dev) M = min(M, (int64_t)*rows_dev);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.x * 128 + lane * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N)
        for (int64_t m = (int64_t)blockIdx.y * 8 + warp; m < M; m += (int64_t)gridDim.y * 8)
            s = f4_add(s, *reinterpret_cast<const float4*>(X + m * ldx + n));
    sh[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && n < N) {
#pragma unroll
        for (int w = 1; w < 8; ++w) s = f4_add(s, sh[w][lane]);
        red_add_v4(out + n, s);
    }
}

}  // namespace ur

extern "C" {

static int gemm_simt_launch(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                            int64_t ldb, float* C, int64_t ldc, const float* bias, int act, float* preact, int64_t ldp, int accumulate,
                            const int32_t* rows_dev, int rows_dim, void* stream) {
    if (M < 0 || N < 0 || K < 0 || (N & 3) || (lda & 3) || (ldb & 3) || (ldc & 3)) return UR_ERR_BAD_ARG;
    if (!transA && (K & 3)) return UR_ERR_BAD_ARG;      // A contiguous along k
    if (transA && (M & 3)) return UR_ERR_BAD_ARG;       // A contiguous along m
    if (transB && (K & 3)) return UR_ERR_BAD_ARG;       // B contiguous along k
    if (preact && (ldp & 3)) return UR_ERR_BAD_ARG;
    if (M == 0 || N == 0) return UR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // small problems (at most 32 tiles of 128 x 128, no device-side row bound): 32 x 32 tiles over many small CTAs
    static const int small_enabled = getenv("UR_SMALL_GEMM") ? atoi(getenv("UR_SMALL_GEMM")) : 1;
    if (small_enabled && !rows_dev && ((M + ur::BM - 1) / ur::BM) * ((N + ur::BN - 1) / ur::BN) <= 32 && K <= 4096 && M * N * K <= (int64_t)80 * 1000 * 1000) {
        const int sgx = (int)((N + ur::SBN - 1) / ur::SBN), sgy = (int)((M + ur::SBM - 1) / ur::SBM);
        int ssplits = 1;
        int64_t sk_chunk = K;
        if (accumulate && !preact && act == 0) {
            const int64_t tiles = (int64_t)sgx * sgy;
            const int64_t want = (2 * ur::kNumSMs + tiles - 1) / tiles;
            const int64_t maxs = (K + 127) / 128;
            ssplits = (int)(want < 1 ? 1 : (want > maxs ? maxs : want));
            sk_chunk = ((K + ssplits - 1) / ssplits + ur::SBK - 1) / ur::SBK * ur::SBK;
            ssplits = (int)((K + sk_chunk - 1) / sk_chunk);
            if (ssplits < 1) { ssplits = 1; sk_chunk = K; }
        }
        const int sacc = accumulate ? (ssplits > 1 ? 1 : 2) : 0;
        dim3 sgrid(sgx, sgy, ssplits);
#define UR_SGEMM(TA, TB)                                                                                                     \
    ur::gemm_simt_small_kernel<TA, TB><<<sgrid, 64 * ur::SG, 0, st>>>((int)M, (int)N, (int)K, A, lda, B, ldb, C, ldc, bias, act, preact, ldp, \
                                                             sacc, (int)sk_chunk)
        if (!transA && !transB) UR_SGEMM(false, false);
        else if (!transA && transB) UR_SGEMM(false, true);
        else if (transA && !transB) UR_SGEMM(true, false);
        else UR_SGEMM(true, true);
#undef UR_SGEMM
        UR_RETURN_LAST_ERROR();
    }
    const int gx = (int)((N + ur::BN - 1) / ur::BN), gy = (int)((M + ur::BM - 1) / ur::BM);
    int splits = 1;
    int64_t k_chunk = K;
    if (accumulate && !preact && act == 0) {            // split-K only where the epilogue is linear
        const int64_t tiles = (int64_t)gx * gy;
        const int64_t want = (2 * ur::kNumSMs + tiles - 1) / tiles;
        const int64_t maxs = (K + 255) / 256;
        splits = (int)(want < 1 ? 1 : (want > maxs ? maxs : want));
        if (splits < 1) splits = 1;
        k_chunk = ((K + splits - 1) / splits + ur::BK - 1) / ur::BK * ur::BK;
        splits = (int)((K + k_chunk - 1) / k_chunk);
        if (splits < 1) { splits = 1; k_chunk = ur::BK; }
    }
    if (accumulate) accumulate = splits > 1 ? 1 : 2;    // atomics only when several CTAs share an output tile
    dim3 grid(gx, gy, splits);
#define UR_GEMM(TA, TB)                                                                                                   \
    ur::gemm_simt_kernel<TA, TB><<<grid, 256, 0, st>>>((int)M, (int)N, (int)K, A, lda, B, ldb, C, ldc, bias, act, preact, ldp, \
                                                       accumulate, (int)k_chunk, rows_dev, rows_dim)
    if (!transA && !transB) UR_GEMM(false, false);
    else if (!transA && transB) UR_GEMM(false, true);
    else if (transA && !transB) UR_GEMM(true, false);
    else UR_GEMM(true, true);
#undef UR_GEMM
    UR_RETURN_LAST_ERROR();
}

int ur_gemm_simt_f32(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                     int64_t ldb, float* C, int64_t ldc, const float* bias, int act, float* preact, int64_t ldp, int accumulate,
                     void* stream) {
    return gemm_simt_launch(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, act, preact, ldp, accumulate, nullptr, 0, stream);
}

// same, with a device-resident token count (see ur_gemm_fused_f32)
int ur_gemm_simt_rows_f32(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                          int64_t ldb, float* C, int64_t ldc, const float* bias, int act, float* preact, int64_t ldp, int accumulate,
                          const int32_t* rows_dev, int rows_dim, void* stream) {
    if (rows_dev && rows_dim != 1 && rows_dim != 2) return UR_ERR_BAD_ARG;
    return gemm_simt_launch(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, act, preact, ldp, accumulate, rows_dev, rows_dim, stream);
}

int ur_act_bwd_f32(float* dY, const float* preact, int64_t n, int act, void* stream) {
    if (n & 3) return UR_ERR_BAD_ARG;
    if (n == 0 || act == 0) return UR_OK;
    int64_t blocks = (n / 4 + 255) / 256;
    const int64_t cap = (int64_t)ur::kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    ur::act_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((float4*)dY, (const float4*)preact, n / 4, act);
    UR_RETURN_LAST_ERROR();
}

int ur_colsum_accum_f32(const float* X, int64_t ldx, int64_t M, int64_t N, float* out, const int32_t* rows_dev, void* stream) {
    if ((N & 3) || (ldx & 3)) return UR_ERR_BAD_ARG;
    if (M == 0 || N == 0) return UR_OK;
    int gy = (int)((M + 255) / 256);
    if (gy > 256) gy = 256;
    if (gy < 1) gy = 1;
    dim3 grid((unsigned)((N + 127) / 128), gy);
    ur::colsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, ldx, M, (int)N, out, rows_dev);
    UR_RETURN_LAST_ERROR();
}

}  // extern "C"
