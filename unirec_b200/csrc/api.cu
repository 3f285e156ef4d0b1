// ABI glue: version query and the GEMM dispatcher (exact-fp32 SIMT kernel vs tcgen05 tensor-core kernel).
#include <stdlib.h>
#include "common.cuh"
#include "../../include/unirec_b200.h"

extern "C" int ur_gemm_tc_f32(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                              int64_t ldb, float* C, int64_t ldc, const float* bias, int act, float* preact, int64_t ldp,
                              int accumulate, int precision, void* stream) __attribute__((weak));

extern "C" int ur_gemm_tc_ex(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                             int64_t ldb, float* C, int64_t ldc, const float* bias, int act, float* preact, int64_t ldp,
                             int accumulate, int precision, const float* dact, int64_t ldd, float* colsum,
                             const int32_t* rows_dev, int rows_dim, void* stream) __attribute__((weak));

// Products with at most 32 output tiles of 128 x 128 (the B-row GEMMs of the trimmed last encoder layer, M = batch size) are
// latency-bound on the persistent tcgen05 kernel (4-32 CTAs, 13-30 us each): they go to the small-tile exact-fp32 kernel instead.
static inline bool ur_small_product(int64_t M, int64_t N, int64_t K) {
    static const int enabled = getenv("UR_SMALL_GEMM") ? atoi(getenv("UR_SMALL_GEMM")) : 1;      // A/B switch
    if (!enabled) return false;
    // (long token reductions and anything above ~160 MFLOP -- e.g. the per-step GRU products -- stay on the tensor cores)
    return ((M + 127) / 128) * ((N + 127) / 128) <= 32 && K <= 4096 && M * N * K <= (int64_t)80 * 1000 * 1000;
}

extern "C" {

int ur_version(void) { return UR_ABI_VERSION; }

int ur_has_tensor_core_gemm(void) { return ur_gemm_tc_f32 != nullptr; }

int ur_gemm_f32(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb,
                float* C, int64_t ldc, const float* bias, int act, float* preact, int64_t ldp, int accumulate, int precision,
                void* stream) {
    if (precision != 0 && ur_gemm_tc_f32 && !ur_small_product(M, N, K)) {
        const int rc = ur_gemm_tc_f32(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, act, preact, ldp, accumulate, precision,
                                      stream);
        if (rc != UR_ERR_UNSUPPORTED) return rc;     // shape not covered by the tensor-core kernel -> exact path
    }
    return ur_gemm_simt_f32(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, act, preact, ldp, accumulate, stream);
}

int ur_gemm_fused_f32(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                      int64_t ldb, float* C, int64_t ldc, const float* bias, int act, float* preact, int64_t ldp, int accumulate,
                      int precision, const float* dact, int64_t ldd, float* colsum, const int32_t* rows_dev, int rows_dim,
                      void* stream) {
    if (dact && (accumulate || preact)) return UR_ERR_BAD_ARG;
    if (precision != 0 && ur_gemm_tc_ex && !(ur_small_product(M, N, K) && !rows_dev)) {
        const int rc = ur_gemm_tc_ex(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, act, preact, ldp, accumulate, precision,
                                     dact, ldd, colsum, rows_dev, rows_dim, stream);
        if (rc != UR_ERR_UNSUPPORTED) return rc;
    }
    // unfused route: exact GEMM, then the activation derivative and the column sums as separate launches
    int rc = ur_gemm_simt_rows_f32(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, dact ? 0 : act, preact, ldp, accumulate,
                                   rows_dev, rows_dim, stream);
    if (rc != UR_OK) return rc;
    if (dact) {
        if (ldc != N || ldd != N) return UR_ERR_UNSUPPORTED;
        rc = ur_act_bwd_f32(C, dact, M * N, act, stream);
        if (rc != UR_OK) return rc;
    }
    if (colsum) rc = ur_colsum_accum_f32(C, ldc, M, N, colsum, rows_dim == 1 ? rows_dev : nullptr, stream);
    return rc;
}

}  // extern "C"
