// Fit of a SHORT matrix (fewer rows than genes: the pseudobulk matrix of scMerge2) that has to return more rows of fullalpha
// than the subspace iteration carries (R/scMerge2_utilis.R:60-65: with the default RandomParam the reference asks for
// svd_k = m_pseudo - ncol(M) singular vectors, the whole row space of Y0).  The n x n Gram would be rank deficient and n can be
// every gene; the m x m side is small.  Literally the reference's formula:
//   Y0 = my_residop(Y, M)                    residual_rows_kernel (group means from K1)
//   K  = Y0 Y0'            (m x m)           DMMA GEMM
//   K  = U diag(sigma^2) U'                  eig_full (whole spectrum, block Jacobi)
//   fullalpha = t(u[, 1:s]) %*% Y  (s x n)   DMMA GEMM        (R/scMerge2_utilis.R:84-86)
// Single rank, no batch standardisation (pseudoRUVIII has none).
#pragma once
#include "common.cuh"
#include "eig_topk.cuh"
#include "gemm_f64.cuh"
#include "seg_colsum.cuh"

namespace scm {

__global__ void residual_rows_kernel(const double* __restrict__ Y, int64_t m, int64_t n, int64_t ld, const int32_t* __restrict__ key,
                                     int32_t G, const double* __restrict__ gsums, const int64_t* __restrict__ gcounts,
                                     double* __restrict__ Y0, int64_t ld0, int64_t* __restrict__ meta, const int32_t* __restrict__ flag) {
    const int64_t r = blockIdx.y;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t k = key[r];
    const bool valid = k >= 0 && k < G;
    if (j < n) {
        const double mean = valid ? gsums[(int64_t)k * n + j] / (double)gcounts[k] : 0.0;
        Y0[r * ld0 + j] = Y[r * ld + j] - mean;
    }
    if (j == 0) {
        if (valid) atomicAdd(reinterpret_cast<unsigned long long*>(meta + 2), 1ull);
        if (r == 0) { meta[0] = flag[0]; meta[1] = m; }
    }
}

__global__ void sqrt_vals_kernel(const double* __restrict__ vals, int s, double* __restrict__ sigma) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < s) sigma[i] = sqrt(fmax(vals[i], 0.0));
}

inline bool fit_dual_applies(const scm_ctx* ctx, int64_t m_local, int64_t n, int32_t s, const int32_t* d_batch, const double* d_gram_centred,
                             const double* d_colmean) {
    if (const char* e = getenv("SCM_FIT_DUAL")) if (atoi(e) == 0) return false;  // diagnostics: force the n x n route
    return s > SCM_MAX_TOPK_ROWS && !d_batch && ctx->world == 1 && m_local < n && m_local <= FULL_EIG_MAX_N && !d_gram_centred && !d_colmean;
}

inline int ruv3_fit_dual(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, const int32_t* d_group, int32_t G, int32_t s,
                         double* d_fullalpha, double* d_sigma, int32_t* iters_out, double* resid_out) {
    if (s > m) return fail(SCM_ERR_INVALID, "scm_ruv3_fit: s=%d rows asked of a %lld-row matrix", s, (long long)m);
    cudaStream_t st = ctx->stream;
    Scratch ws(ctx);
    double *gsums, *Y0, *K, *vals, *U;
    int64_t *gcounts, *meta;
    int32_t* flag;
    const int64_t ld0 = (n + 1) & ~(int64_t)1;
    SCM_TRY(ws.alloc(&gsums, (int64_t)G * n)); SCM_TRY(ws.alloc(&gcounts, G)); SCM_TRY(ws.alloc(&meta, 4)); SCM_TRY(ws.alloc(&flag, 1));
    SCM_TRY(ws.alloc(&Y0, m * ld0)); SCM_TRY(ws.alloc(&K, m * m)); SCM_TRY(ws.alloc(&vals, s)); SCM_TRY(ws.alloc(&U, (int64_t)s * m));
    SCM_TRY(zero_async(ctx, flag, sizeof(int32_t)));
    SCM_TRY(zero_async(ctx, meta, sizeof(int64_t) * 4));
    ctx->ev_used[ST_PRESHIFT] = false;
    ctx->ev_used[ST_STD] = false;
    {
        StageTimer t(ctx, ST_SEG);
        SCM_TRY(seg_colsum(ctx, d_Y, m, n, ld, d_group, G, nullptr, gsums, gcounts, flag));
        residual_rows_kernel<<<dim3((unsigned)ceil_div(n, 256), (unsigned)m), 256, 0, st>>>(d_Y, m, n, ld, d_group, G, gsums, gcounts, Y0, ld0,
                                                                                            meta, flag);
        SCM_LAUNCH_CHECK(ctx);
    }
    {
        StageTimer t(ctx, ST_GRAM);
        SCM_TRY(gemm(ctx, (int)m, (int)m, (int)n, 1.0, rowmajor(Y0, ld0), transposed(Y0, ld0), 0.0, K, m));
    }
    int64_t* hmeta = reinterpret_cast<int64_t*>(ctx->h_pinned + 2048);
    SCM_CUDA(cudaMemcpyAsync(hmeta, meta, sizeof(int64_t) * 3, cudaMemcpyDeviceToHost, st));
    int rc;
    {
        StageTimer t(ctx, ST_EIG);
        rc = eig_full(ctx, K, m, s, vals, U, iters_out, resid_out, nullptr);  // its first read-back also lands hmeta
        if (hmeta[0] != 0) return fail(SCM_ERR_NONFINITE, "Y contains missing values or infinities.  This is not supported.");
        if (hmeta[2] != hmeta[1])
            return fail(SCM_ERR_INVALID, "%lld of %lld cells have a replicate key outside [0, G): M must assign every cell to a group",
                        (long long)(hmeta[1] - hmeta[2]), (long long)hmeta[1]);
        if (rc != 0 && rc != SCM_ERR_NOCONV) return rc;
        SCM_TRY(gemm(ctx, s, (int)n, (int)m, 1.0, rowmajor(U, m), rowmajor(d_Y, ld), 0.0, d_fullalpha, n));
        if (d_sigma) {
            sqrt_vals_kernel<<<(unsigned)ceil_div(s, 256), 256, 0, st>>>(vals, s, d_sigma);
            SCM_LAUNCH_CHECK(ctx);
        }
    }
    return rc;
}

}  // namespace scm
