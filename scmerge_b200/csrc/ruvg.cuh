// scRUVg (R/scRUVg.R:30-85) on the kernels of the RUV-III path.
//
// Reference (Z = 1, eta = NULL, what the default call does):
//     Y0 = Y - colMeans(Y)                        (my_residop(Y, matrix(1, m, 1)), R/scRUVg.R:61-63)
//     W  = first k left singular vectors of Y0[, ctl]   (svd of Y0c or of Y0c Y0c', R/scRUVg.R:66-73,77)
//     alpha = solve(W'W) W' Y0 = W' Y0            (R/scRUVg.R:78)
//     newY  = Y - W alpha                         (R/scRUVg.R:80)
// Reduced form used here: with G = Y0'Y0 (n x n, K3 with shift = column means) and (lambda_a, v_a) the top-k eigenpairs
// of its control block G[ctl, ctl] (K4),
//     W = Y0c V diag(1/sigma),   alpha = diag(1/sigma) V' G[ctl, :],   sigma_a = sqrt(lambda_a),
// i.e. newY = Y - ((Y - mean) B) A with B = 1[ctl] V diag(1/sigma) (n x k) and A = alpha (k x n): the fused rank-k update
// (K6) does the rest and also emits W.  The m x m cross product of the reference's >= 5:1 aspect-ratio branch is never
// formed.  Cross-rank state (column sums, Gram) is summed exactly like in scm_ruv3_fit (fit_core.cuh: the Gram runs on the
// pre-shifted copy of the cells, SHIFT = false kernel).
#pragma once
#include "adjust.cuh"
#include "common.cuh"
#include "eig_topk.cuh"
#include "fit_core.cuh"
#include "gemm_f64.cuh"
#include "gram_tma.cuh"
#include "seg_colsum.cuh"

namespace scm {

// cidx[j] = j-th control gene (ordered), *count = c.  Single CTA.
__global__ void ctl_index_kernel(const uint8_t* __restrict__ ctl, int64_t n, int32_t* __restrict__ cidx, int32_t* __restrict__ count) {
    __shared__ int s_flag[1024];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int64_t g0 = 0; g0 < n; g0 += 1024) {
        const int64_t g = g0 + threadIdx.x;
        s_flag[threadIdx.x] = (g < n && ctl[g]) ? 1 : 0;
        __syncthreads();
        if (threadIdx.x == 0) {
            int base = s_base;
            for (int t = 0; t < 1024 && g0 + t < n; t++) if (s_flag[t]) cidx[base++] = (int32_t)(g0 + t);
            s_base = base;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = s_base;
}

// Gc[i, :] = G[cidx[i], :] (c x n) and Gcc[i, j] = G[cidx[i], cidx[j]] (c x c)
__global__ void gather_ctl_rows_kernel(const double* __restrict__ G, int64_t n, const int32_t* __restrict__ cidx, int c,
                                       double* __restrict__ Gc, double* __restrict__ Gcc) {
    const int i = blockIdx.x;
    const double* src = G + (int64_t)cidx[i] * n;
    for (int64_t g = threadIdx.x; g < n; g += blockDim.x) Gc[(int64_t)i * n + g] = src[g];
    for (int j = threadIdx.x; j < c; j += blockDim.x) Gcc[(int64_t)i * c + j] = src[cidx[j]];
}

// B[g, a] = ctl gene g == cidx[j] ? vecs[a, j] / sigma_a : 0;  Vs[a, j] = vecs[a, j] / sigma_a;  sigma_a = sqrt(vals[a])
__global__ void ruvg_factors_kernel(const double* __restrict__ vals, const double* __restrict__ vecs, const int32_t* __restrict__ cidx,
                                    int c, int k, int64_t n, double* __restrict__ B, double* __restrict__ Vs, double* __restrict__ sigma,
                                    int* __restrict__ flag) {
    const int a = blockIdx.x;
    const double lam = vals[a];
    // a (numerically) zero eigenvalue means k exceeds the rank of the centred control block: W would be undefined
    if (!(lam > 0.0) || lam <= 1e-26 * vals[0]) { if (threadIdx.x == 0) *flag = 1; }
    const double inv = lam > 0.0 ? 1.0 / sqrt(lam) : 0.0;
    if (threadIdx.x == 0 && sigma) sigma[a] = sqrt(fmax(lam, 0.0));
    for (int j = threadIdx.x; j < c; j += blockDim.x) {
        const double v = vecs[(int64_t)a * c + j] * inv;
        Vs[(int64_t)a * c + j] = v;
        B[(int64_t)cidx[j] * k + a] = v;
    }
}

// The fit half of scRUVg.  Outputs: d_alpha (k x n), d_center (n, the column means), d_sigma (k, optional: singular values
// of Y0[, ctl]) and the coefficient object for scm_adjust (which also emits W).  LOCAL rows + all-reduce hook.
inline int ruvg_fit(scm_ctx* ctx, const double* d_Y, int64_t m_local, int64_t n, int64_t ld, const uint8_t* d_ctl, int32_t k, int mode,
                    double tol, int32_t maxit, uint64_t seed, double* d_alpha, double* d_center, double* d_sigma, scm_coef** coef_out,
                    int32_t* iters_out, double* resid_out) {
    if (!d_Y || !d_ctl || n <= 0 || k <= 0 || !d_alpha || !d_center || !coef_out) return fail(SCM_ERR_INVALID, "scm_ruvg_fit: bad arguments");
    if (k > SCM_MAX_K) return fail(SCM_ERR_UNSUPPORTED, "scm_ruvg_fit: k=%d > %d", k, SCM_MAX_K);
    cudaStream_t st = ctx->stream;
    Scratch ws(ctx);
    double *Gc, *Gcc, *vals, *vecs, *Bm, *Vs;
    int32_t *key, *flag, *cidx, *ccnt;
    SCM_TRY(ws.alloc(&flag, 2)); SCM_TRY(ws.alloc(&key, m_local > 0 ? m_local : 1));
    SCM_TRY(ws.alloc(&cidx, n)); SCM_TRY(ws.alloc(&ccnt, 1));
    SCM_TRY(zero_async(ctx, flag, 2 * sizeof(int32_t)));
    SCM_TRY(zero_async(ctx, key, sizeof(int32_t) * (m_local > 0 ? m_local : 1)));
    // column means (my_residop with Z = 1) + NA/Inf scan + centred Gram: the fit's cell passes with ONE replicate group
    FitState fs;
    SCM_TRY(fit_state(ctx, ws, d_Y, m_local, n, ld, key, 1, nullptr, 0, fs));
    double* Gm = fs.Gm;
    SCM_TRY(copy_async(ctx, d_center, fs.ybar, sizeof(double) * n));
    ctl_index_kernel<<<1, 1024, 0, st>>>(d_ctl, n, cidx, ccnt);
    SCM_LAUNCH_CHECK(ctx);
    int64_t* hmeta;
    SCM_TRY(fit_meta_to_host(ctx, fs, &hmeta));
    int32_t c = 0;
    SCM_CUDA(cudaMemcpyAsync(&c, ccnt, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SCM_CUDA(cudaStreamSynchronize(st));
    SCM_TRY(fit_meta_check(hmeta));
    if (k > c) return fail(SCM_ERR_INVALID, "k must not be larger than the number of controls");  // R/scRUVg.R:56-58
    SCM_TRY(ws.alloc(&Gc, (int64_t)c * n)); SCM_TRY(ws.alloc(&Gcc, (int64_t)c * c));
    SCM_TRY(ws.alloc(&vals, k)); SCM_TRY(ws.alloc(&vecs, (int64_t)k * c));
    SCM_TRY(ws.alloc(&Bm, n * k)); SCM_TRY(ws.alloc(&Vs, (int64_t)k * c));
    gather_ctl_rows_kernel<<<c, 256, 0, st>>>(Gm, n, cidx, c, Gc, Gcc);
    SCM_LAUNCH_CHECK(ctx);
    int rc;
    {
        StageTimer t(ctx, ST_EIG);
        rc = eig_topk(ctx, Gcc, c, k, k, mode, tol, maxit, seed, vals, vecs, iters_out, resid_out);
        if (rc != 0 && rc != SCM_ERR_NOCONV) return rc;
    }
    {
        StageTimer t(ctx, ST_COEF);
        SCM_CUDA(cudaMemsetAsync(Bm, 0, sizeof(double) * n * k, st));
        ruvg_factors_kernel<<<k, 256, 0, st>>>(vals, vecs, cidx, c, k, n, Bm, Vs, d_sigma, flag + 1);
        SCM_LAUNCH_CHECK(ctx);
        SCM_TRY(gemm(ctx, k, (int)n, c, 1.0, rowmajor(Vs, c), rowmajor(Gc, n), 0.0, d_alpha, n));  // alpha = diag(1/sigma) V' G[ctl, :]
        int32_t hflag[2];
        SCM_CUDA(cudaMemcpyAsync(hflag, flag, sizeof(hflag), cudaMemcpyDeviceToHost, st));
        SCM_CUDA(cudaStreamSynchronize(st));
        if (hflag[1]) return fail(SCM_ERR_SINGULAR, "scm_ruvg_fit: the centred control block has rank < k=%d (W'W is singular)", k);
        SCM_TRY(coef_from_factors(ctx, Bm, d_alpha, n, k, d_ctl, d_center, coef_out));
    }
    return rc;
}

// alpha = solve(W'W) W' (Y - 1 center')   (R/scRUVg.R:78 with a GIVEN W, e.g. a re-used `fullW[, 1:k]`; center = colMeans of the
// matrix the residual operator of Z = 1 is applied to, :61-63).  W: m_local x k row-major.  Sums over cells are all-reduced.
// One skinny DMMA GEMM over the cells (k x n, reduction length m, split over the SMs and reduced in order) + k x k work.
__global__ void ls_rank1_kernel(double* __restrict__ P, const double* __restrict__ s, const double* __restrict__ center, int k, int64_t n) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)k * n) return;
    const int a = (int)(e / n);
    const int64_t g = e - (int64_t)a * n;
    P[e] -= s[a] * center[g];
}
__global__ void ls_colsum_kernel(const double* __restrict__ W, int64_t m, int k, double* __restrict__ s) {  // s[a] = sum_i W[i, a]; one CTA per column
    __shared__ double wb[32];
    const int a = blockIdx.x;
    double t = 0.0;
    for (int64_t i = threadIdx.x; i < m; i += blockDim.x) t += W[i * k + a];
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) wb[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) { double r = 0.0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) r += wb[w]; s[a] = r; }
}
inline int ls_coef(scm_ctx* ctx, const double* d_Y, int64_t m_local, int64_t n, int64_t ld, const double* d_W, int32_t k,
                   const double* d_center, double* d_alpha) {
    if (!d_Y || !d_W || !d_alpha || n <= 0 || m_local < 0 || ld < n || k <= 0) return fail(SCM_ERR_INVALID, "ls_coef: bad arguments");
    if (k > SCM_MAX_K) return fail(SCM_ERR_UNSUPPORTED, "ls_coef: k=%d > %d", k, SCM_MAX_K);
    if (m_local >= ((int64_t)1 << 31)) return fail(SCM_ERR_UNSUPPORTED, "ls_coef: m >= 2^31");
    cudaStream_t st = ctx->stream;
    Scratch ws(ctx);
    const int64_t len = (int64_t)k * k + (int64_t)k * n + k;
    double *buf, *Linv, *Ninv;
    int* flag;
    SCM_TRY(ws.alloc(&buf, len)); SCM_TRY(ws.alloc(&Linv, (int64_t)k * k)); SCM_TRY(ws.alloc(&Ninv, (int64_t)k * k)); SCM_TRY(ws.alloc(&flag, 2));
    double *N = buf, *P = buf + (int64_t)k * k, *s = P + (int64_t)k * n;
    SCM_TRY(zero_async(ctx, buf, sizeof(double) * len));
    SCM_TRY(zero_async(ctx, flag, sizeof(int) * 2));
    if (m_local > 0) {
        SCM_TRY(gemm(ctx, k, k, (int)m_local, 1.0, transposed(d_W, k), rowmajor(d_W, k), 0.0, N, k));
        SCM_TRY(gemm(ctx, k, (int)n, (int)m_local, 1.0, transposed(d_W, k), rowmajor(d_Y, ld), 0.0, P, n));
        ls_colsum_kernel<<<k, 256, 0, st>>>(d_W, m_local, k, s);
        SCM_LAUNCH_CHECK(ctx);
    }
    SCM_TRY(allreduce(ctx, buf, len, 0));
    if (d_center) {
        ls_rank1_kernel<<<(unsigned)ceil_div((int64_t)k * n, 256), 256, 0, st>>>(P, s, d_center, k, n);
        SCM_LAUNCH_CHECK(ctx);
    }
    SCM_TRY(chol_inv_launch(ctx, N, k, 0.0, Linv, flag));
    SCM_TRY(gemm(ctx, k, k, k, 1.0, transposed(Linv, k), rowmajor(Linv, k), 0.0, Ninv, k));
    SCM_TRY(gemm(ctx, k, (int)n, k, 1.0, rowmajor(Ninv, k), rowmajor(P, n), 0.0, d_alpha, n));
    int h = 0;
    SCM_CUDA(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    SCM_CUDA(cudaStreamSynchronize(st));
    if (h) return fail(SCM_ERR_SINGULAR, "ls_coef: W'W is not positive definite (k=%d)", k);
    return 0;
}

}  // namespace scm
