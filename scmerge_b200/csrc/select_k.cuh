// Choice of ruvK in scRUVIII (R/scRUVIII.R:80-114,182-199):
//     for every candidate k:  PCA(newY_k, rank 10, centred + scaled) -> silhouette width of the cell types and of the batches
//     in PC space -> F-score; the k with the largest F-score wins.
//
// PCA without another pass over the cells.  The adjusted matrix is an affine image of Y,
//     newY - mean(newY) = (Y - ybar)(I - B A),
// so its centred Gram is C = (I - BA)' Gc (I - BA) with Gc the centred Gram of Y that the fit has already computed
// (scm_ruv3_fit, d_gram_centred).  C is expanded in the rank-k factors (O(n^2 k), never an n^3 product):
//     C = Gc - (Gc B) A - A' (Gc B)' + A' (B' Gc B) A.
// Scaling to unit variance (runPCA(scale = TRUE)) turns it into R = D C D, D = diag(1 / sd_j), sd_j^2 = C_jj / (m - 1); the
// per-gene affine de-standardisation of scRUVIII (R/scRUVIII.R:117-119) cancels in R, so the PCA of the final newY equals the
// PCA of the standardised newY the reference feeds to calculateSil.  With (theta_a, v_a) the top eigenpairs of R (K4) the
// scores are x = (newY - mean) D V = (Y - ybar) Bp, Bp = (I - BA) D V (n x rank): ONE W-only pass of the rank-k kernel (K6
// pass A) over the ORIGINAL matrix gives the PC scores of the adjusted one.
//
// Silhouette (cluster::silhouette on dist(x), exact, O(m^2) pair distances but O(m) memory -- the reference materialises
// the m x m `dist`).  Cells are ordered by the composite key (label_a, label_b); a thread owns SIL_IPT cells and walks all
// cells group by group (tiles broadcast from shared memory), flushing one sum of distances per (cell, composite group), so
// one distance pass serves both labelings.  Sums run over j in a fixed order: deterministic.  FP64 FMA-pipe bound:
// ~ (3 d + 12) flop per pair.
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include "adjust.cuh"
#include "common.cuh"
#include "eig_topk.cuh"
#include "gemm_f64.cuh"
#include "seg_colsum.cuh"

namespace scm {

// R[i, j] = d_i d_j (C[i, j] + C[j, i]) / 2,  d_j = 1 / sqrt(C[j, j] / (m - 1))   (zero-variance gene -> d_j = 0: the gene is
// dropped from the PCA instead of poisoning it with NaN; the reference's scale() would fail there)
__global__ void corr_scale_kernel(const double* __restrict__ Cm, int64_t n, double denom, double* __restrict__ R, double* __restrict__ dvec) {
    const int64_t i = blockIdx.y;
    const double cii = Cm[i * n + i];
    const double di = cii > 0.0 ? 1.0 / sqrt(cii / denom) : 0.0;
    if (blockIdx.x == 0 && threadIdx.x == 0) dvec[i] = di;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const double cjj = Cm[j * n + j];
        const double dj = cjj > 0.0 ? 1.0 / sqrt(cjj / denom) : 0.0;
        R[i * n + j] = di * dj * 0.5 * (Cm[i * n + j] + Cm[j * n + i]);
    }
}
// DVt[j, a] = d_j * vecs[a, j]
__global__ void scale_vecs_kernel(const double* __restrict__ vecs, const double* __restrict__ dvec, int64_t n, int r, double* __restrict__ DVt) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    for (int a = 0; a < r; a++) DVt[j * r + a] = dvec[j] * vecs[(int64_t)a * n + j];
}
__global__ void fill_u8_kernel(uint8_t* p, int64_t n, uint8_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// PC scores of the adjusted matrix (coef != NULL) or of Y itself (coef == NULL): d_scores[m_local x rank] (row-major),
// d_sdev[rank] = sqrt(theta / (m - 1)) (prcomp's sdev).  d_colmean = column means of Y over ALL cells.
inline int pca_scores(scm_ctx* ctx, const double* d_gram_centred, int64_t n, int64_t m_total, const scm_coef* coef,
                      const double* d_colmean, int32_t rank, const double* d_Y, int64_t m_local, int64_t ld, double* d_scores,
                      double* d_sdev, double tol, int32_t maxit, int32_t* iters_out, double* resid_out) {
    if (!d_gram_centred || !d_colmean || !d_scores || n <= 0 || rank <= 0 || m_total < 2) return fail(SCM_ERR_INVALID, "pca_scores: bad arguments");
    if (rank > n || rank > SCM_MAX_K) return fail(SCM_ERR_UNSUPPORTED, "pca_scores: rank=%d outside 1..min(n, %d)", rank, SCM_MAX_K);
    if (coef && coef->n != n) return fail(SCM_ERR_INVALID, "pca_scores: coefficient object does not match n");
    cudaStream_t st = ctx->stream;
    Scratch ws(ctx);
    double *Cm, *R, *dvec, *vals, *vecs, *DVt, *Bp;
    SCM_TRY(ws.alloc(&Cm, n * n)); SCM_TRY(ws.alloc(&R, n * n)); SCM_TRY(ws.alloc(&dvec, n));
    SCM_TRY(ws.alloc(&vals, rank)); SCM_TRY(ws.alloc(&vecs, (int64_t)rank * n));
    SCM_TRY(ws.alloc(&DVt, n * rank)); SCM_TRY(ws.alloc(&Bp, n * rank));
    SCM_CUDA(cudaMemcpyAsync(Cm, d_gram_centred, sizeof(double) * n * n, cudaMemcpyDeviceToDevice, st));
    if (coef) {
        const int k = coef->k;
        const double *B = coef->Bmat, *A = coef->Amat;
        double *GB, *T, *TA;
        SCM_TRY(ws.alloc(&GB, n * k)); SCM_TRY(ws.alloc(&T, (int64_t)k * k)); SCM_TRY(ws.alloc(&TA, n * k));
        SCM_TRY(gemm(ctx, (int)n, k, (int)n, 1.0, rowmajor(d_gram_centred, n), rowmajor(B, k), 0.0, GB, k));
        SCM_TRY(gemm(ctx, k, k, (int)n, 1.0, transposed(B, k), rowmajor(GB, k), 0.0, T, k));
        SCM_TRY(gemm(ctx, k, (int)n, k, 1.0, rowmajor(T, k), rowmajor(A, n), 0.0, TA, n));
        SCM_TRY(gemm(ctx, (int)n, (int)n, k, -1.0, rowmajor(GB, k), rowmajor(A, n), 1.0, Cm, n));
        SCM_TRY(gemm(ctx, (int)n, (int)n, k, -1.0, transposed(A, n), transposed(GB, k), 1.0, Cm, n));
        SCM_TRY(gemm(ctx, (int)n, (int)n, k, 1.0, transposed(A, n), rowmajor(TA, n), 1.0, Cm, n));
    }
    {
        dim3 grid((unsigned)(ceil_div(n, 256) < 8 ? ceil_div(n, 256) : 8), (unsigned)n);
        corr_scale_kernel<<<grid, 256, 0, st>>>(Cm, n, (double)(m_total - 1), R, dvec);
        SCM_LAUNCH_CHECK(ctx);
    }
    int rc;
    {
        StageTimer t(ctx, ST_EIG);
        rc = eig_topk(ctx, R, n, rank, rank, SCM_MODE_EXACT, tol, maxit, 1, vals, vecs, iters_out, resid_out);
        if (rc != 0 && rc != SCM_ERR_NOCONV) return rc;
    }
    scale_vecs_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(vecs, dvec, n, rank, DVt);
    SCM_LAUNCH_CHECK(ctx);
    SCM_CUDA(cudaMemcpyAsync(Bp, DVt, sizeof(double) * n * rank, cudaMemcpyDeviceToDevice, st));
    if (coef) {
        double* ADV;
        SCM_TRY(ws.alloc(&ADV, (int64_t)coef->k * rank));
        SCM_TRY(gemm(ctx, coef->k, rank, (int)n, 1.0, rowmajor(coef->Amat, n), rowmajor(DVt, rank), 0.0, ADV, rank));
        SCM_TRY(gemm(ctx, (int)n, rank, coef->k, -1.0, rowmajor(coef->Bmat, coef->k), rowmajor(ADV, rank), 1.0, Bp, rank));
    }
    if (d_sdev) {
        fullalpha_kernel<<<dim3(1, (unsigned)rank), 32, 0, st>>>(vals, vecs, rank, 0, nullptr, d_sdev);  // sqrt(theta): singular values
        SCM_LAUNCH_CHECK(ctx);
    }
    // scores = (Y - ybar) Bp: the W-only pass of the rank-k kernel
    uint8_t* all_rows;
    double* Azero;
    SCM_TRY(ws.alloc(&all_rows, n)); SCM_TRY(ws.alloc(&Azero, n * rank));
    fill_u8_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(all_rows, n, 1);
    SCM_LAUNCH_CHECK(ctx);
    SCM_CUDA(cudaMemsetAsync(Azero, 0, sizeof(double) * n * rank, st));
    scm_coef* pc = nullptr;
    SCM_TRY(coef_from_factors(ctx, Bp, Azero, n, rank, all_rows, d_colmean, &pc));
    int arc = 0;
    if (m_local > 0) {
        StageTimer t(ctx, ST_ADJUST);
        arc = adjust(ctx, d_Y, m_local, n, ld, pc, nullptr, 0, nullptr, 0, d_scores);
    }
    coef_destroy(ctx, pc);
    if (arc != 0) return arc;
    return rc;
}

// ---- silhouette ---------------------------------------------------------------------------------------------------------
constexpr int SIL_THREADS = 128, SIL_IPT = 2, SIL_JT = 128, SIL_MAXD = 16;

__global__ void sil_key_kernel(const int32_t* __restrict__ la, const int32_t* __restrict__ lb, int64_t m, int32_t Ca, int32_t Cb,
                               int32_t* __restrict__ key, int32_t* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int32_t a = la[i], b = lb ? lb[i] : 0;
    if (a < 0 || a >= Ca || b < 0 || b >= Cb) { *bad = 1; key[i] = 0; return; }
    key[i] = a * Cb + b;
}
// Xs[pos, :] = X[perm[pos], :] padded to D columns (zeros do not change distances)
template <int D>
__global__ void sil_gather_kernel(const double* __restrict__ X, int64_t ldx, int d, const int32_t* __restrict__ perm, int64_t m,
                                  double* __restrict__ Xs) {
    const int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= m) return;
    const double* src = X + (int64_t)perm[pos] * ldx;
#pragma unroll
    for (int c = 0; c < D; c++) Xs[pos * D + c] = c < d ? src[c] : 0.0;
}

// S[g, i - i0] = sum over the cells j of composite group g of ||x_i - x_j||   (sorted order, i in [i0, i1))
template <int D>
__global__ void __launch_bounds__(SIL_THREADS)
sil_pairsum_kernel(const double* __restrict__ Xs, int64_t m, const int64_t* __restrict__ gstart, int32_t Cc, int64_t i0, int64_t i1,
                   double* __restrict__ S) {
    __shared__ __align__(16) double tile[SIL_JT * D];
    const int64_t ni = i1 - i0;
    const int64_t ib = (int64_t)blockIdx.x * (SIL_THREADS * SIL_IPT) + threadIdx.x;  // local index of this thread's first cell
    double xi[SIL_IPT][D];
    bool ok[SIL_IPT];
#pragma unroll
    for (int u = 0; u < SIL_IPT; u++) {
        const int64_t il = ib + (int64_t)u * SIL_THREADS;
        ok[u] = il < ni;
        const double* p = Xs + (i0 + (ok[u] ? il : 0)) * D;
#pragma unroll
        for (int c = 0; c < D; c++) xi[u][c] = p[c];
    }
    for (int32_t g = 0; g < Cc; g++) {
        const int64_t j0 = gstart[g], j1 = gstart[g + 1];
        double acc[SIL_IPT];
#pragma unroll
        for (int u = 0; u < SIL_IPT; u++) acc[u] = 0.0;
        for (int64_t jt = j0; jt < j1; jt += SIL_JT) {
            const int len = (int)((j1 - jt) < SIL_JT ? (j1 - jt) : SIL_JT);
            __syncthreads();
            for (int e = threadIdx.x; e < len * D; e += SIL_THREADS) tile[e] = Xs[jt * D + e];
            __syncthreads();
#pragma unroll 4
            for (int j = 0; j < len; j++) {
                double xj[D];
#pragma unroll
                for (int c = 0; c < D; c += 2) {
                    const double2 v = *reinterpret_cast<const double2*>(tile + j * D + c);
                    xj[c] = v.x; xj[c + 1] = v.y;
                }
#pragma unroll
                for (int u = 0; u < SIL_IPT; u++) {
                    double s2 = 0.0;
#pragma unroll
                    for (int c = 0; c < D; c++) { const double t = xi[u][c] - xj[c]; s2 = fma(t, t, s2); }
                    acc[u] += sqrt(s2);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < SIL_IPT; u++) {
            const int64_t il = ib + (int64_t)u * SIL_THREADS;
            if (ok[u]) S[(int64_t)g * ni + il] = acc[u];
        }
    }
}

// s_i for both labelings from the per-group sums; per-CTA partial sums of s_i (fixed order inside the CTA).
__global__ void __launch_bounds__(256)
sil_final_kernel(const double* __restrict__ S, const int32_t* __restrict__ skey, const int64_t* __restrict__ gstart, int32_t Ca, int32_t Cb,
                 int64_t i0, int64_t i1, double* __restrict__ part_a, double* __restrict__ part_b) {
    __shared__ double red_a[256], red_b[256];
    const int64_t ni = i1 - i0;
    const int64_t il = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double sa = 0.0, sb = 0.0;
    if (il < ni) {
        const int32_t key = skey[i0 + il];
        const int32_t own_a = key / Cb, own_b = key % Cb;
        for (int pass = 0; pass < 2; pass++) {
            const int32_t Cx = pass == 0 ? Ca : Cb, own = pass == 0 ? own_a : own_b;
            if (pass == 1 && Cb <= 1) break;
            double a_sum = 0.0, b_best = INFINITY;
            int64_t own_cnt = 0;
            int nonempty = 0;
            for (int32_t c = 0; c < Cx; c++) {
                double t = 0.0;
                int64_t cnt = 0;
                const int32_t inner = pass == 0 ? Cb : Ca;
                for (int32_t q = 0; q < inner; q++) {
                    const int32_t g = pass == 0 ? c * Cb + q : q * Cb + c;
                    t += S[(int64_t)g * ni + il];
                    cnt += gstart[g + 1] - gstart[g];
                }
                if (cnt == 0) continue;
                nonempty++;
                if (c == own) { a_sum = t; own_cnt = cnt; }
                else { const double mean = t / (double)cnt; if (mean < b_best) b_best = mean; }
            }
            double s = 0.0;
            if (own_cnt > 1 && nonempty > 1) {
                const double a = a_sum / (double)(own_cnt - 1);
                const double mx = a > b_best ? a : b_best;
                s = mx > 0.0 ? (b_best - a) / mx : 0.0;
            }
            if (pass == 0) sa = s; else sb = s;
        }
    }
    red_a[threadIdx.x] = sa; red_b[threadIdx.x] = sb;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) { red_a[threadIdx.x] += red_a[threadIdx.x + off]; red_b[threadIdx.x] += red_b[threadIdx.x + off]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part_a[blockIdx.x] = red_a[0]; part_b[blockIdx.x] = red_b[0]; }
}

// Sum of the silhouette widths s_i over the cells at sorted positions [i0, i1) (order by (label_a, label_b); any partition of
// [0, m) over ranks covers every cell once) for labeling a and, when d_lab_b != NULL, labeling b.  avg.width = sum / m.
inline int silhouette(scm_ctx* ctx, const double* d_X, int64_t m, int32_t d, int64_t ldx, const int32_t* d_lab_a, int32_t Ca,
                      const int32_t* d_lab_b, int32_t Cb, int64_t i0, int64_t i1, double* h_sum_a, double* h_sum_b) {
    if (!d_X || !d_lab_a || m <= 0 || d <= 0 || ldx < d || Ca <= 0 || !h_sum_a) return fail(SCM_ERR_INVALID, "silhouette: bad arguments");
    if (d > SIL_MAXD) return fail(SCM_ERR_UNSUPPORTED, "silhouette: more than %d dimensions", SIL_MAXD);
    if (m >= (int64_t)1 << 31) return fail(SCM_ERR_UNSUPPORTED, "silhouette: m >= 2^31");
    if (!d_lab_b) Cb = 1;
    if (Cb <= 0 || (int64_t)Ca * Cb > (1 << 20)) return fail(SCM_ERR_UNSUPPORTED, "silhouette: too many label combinations");
    if (i0 < 0) i0 = 0;
    if (i1 > m) i1 = m;
    *h_sum_a = 0.0;
    if (h_sum_b) *h_sum_b = 0.0;
    if (i1 <= i0) return 0;
    const int32_t Cc = Ca * Cb;
    const int64_t ni = i1 - i0;
    if ((double)Cc * (double)ni * 8.0 > 64e9) return fail(SCM_ERR_UNSUPPORTED, "silhouette: %d label combinations x %lld cells exceed the 64 GB work buffer; split the cell range", Cc, (long long)ni);
    cudaStream_t st = ctx->stream;
    Scratch ws(ctx);
    int32_t *key, *bad, *sval_in, *perm;
    uint32_t *skey_in, *skey_out;
    int *cnt, *chunk_off, *pslot_off;
    int64_t* gstart;
    double *Xs, *S, *part_a, *part_b;
    const int D = d <= 2 ? 2 : (d <= 4 ? 4 : (d <= 8 ? 8 : (d <= 10 ? 10 : 16)));
    SCM_TRY(ws.alloc(&key, m)); SCM_TRY(ws.alloc(&bad, 1));
    SCM_TRY(ws.alloc(&skey_in, m)); SCM_TRY(ws.alloc(&skey_out, m)); SCM_TRY(ws.alloc(&sval_in, m)); SCM_TRY(ws.alloc(&perm, m));
    SCM_TRY(ws.alloc(&cnt, (int64_t)Cc + 1)); SCM_TRY(ws.alloc(&chunk_off, (int64_t)Cc + 1)); SCM_TRY(ws.alloc(&pslot_off, (int64_t)Cc + 1));
    SCM_TRY(ws.alloc(&gstart, (int64_t)Cc + 1));
    SCM_TRY(ws.alloc(&Xs, m * D)); SCM_TRY(ws.alloc(&S, (int64_t)Cc * ni));
    const unsigned fblocks = (unsigned)ceil_div(ni, 256);
    SCM_TRY(ws.alloc(&part_a, fblocks)); SCM_TRY(ws.alloc(&part_b, fblocks));
    SCM_CUDA(cudaMemsetAsync(bad, 0, sizeof(int32_t), st));
    SCM_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * ((int64_t)Cc + 1), st));
    sil_key_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, st>>>(d_lab_a, d_lab_b, m, Ca, Cb, key, bad);
    SCM_LAUNCH_CHECK(ctx);
    const int use_smem = Cc < 4096;
    seg_prepare_kernel<<<(unsigned)ceil_div(m, SEG_PREP_ROWS), 256, use_smem ? sizeof(int) * (Cc + 1) : 0, st>>>(key, m, Cc, skey_in, sval_in, cnt, use_smem);
    SCM_LAUNCH_CHECK(ctx);
    seg_scan_kernel<<<1, 1024, 0, st>>>(cnt, Cc, gstart, chunk_off, pslot_off, nullptr);
    SCM_LAUNCH_CHECK(ctx);
    int end_bit = 1;
    while (((int64_t)1 << end_bit) <= Cc) end_bit++;
    size_t tmp_bytes = 0;
    SCM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, skey_in, skey_out, sval_in, perm, (int)m, 0, end_bit, st));
    char* tmp;
    SCM_TRY(ws.alloc(&tmp, (int64_t)tmp_bytes));
    SCM_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, skey_in, skey_out, sval_in, perm, (int)m, 0, end_bit, st));
    ctx->launches += 3;
    int32_t hbad = 0;
    SCM_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SCM_CUDA(cudaStreamSynchronize(st));
    if (hbad) return fail(SCM_ERR_INVALID, "silhouette: a label is outside [0, C)");
    const unsigned gblocks = (unsigned)ceil_div(m, 256);
    const unsigned pblocks = (unsigned)ceil_div(ni, SIL_THREADS * SIL_IPT);
#define SIL_RUN(DD)                                                                                        \
    do {                                                                                                   \
        sil_gather_kernel<DD><<<gblocks, 256, 0, st>>>(d_X, ldx, d, perm, m, Xs);                           \
        SCM_LAUNCH_CHECK(ctx);                                                                             \
        sil_pairsum_kernel<DD><<<pblocks, SIL_THREADS, 0, st>>>(Xs, m, gstart, Cc, i0, i1, S);              \
        SCM_LAUNCH_CHECK(ctx);                                                                             \
    } while (0)
    switch (D) {
        case 2: SIL_RUN(2); break;
        case 4: SIL_RUN(4); break;
        case 8: SIL_RUN(8); break;
        case 10: SIL_RUN(10); break;
        default: SIL_RUN(16); break;
    }
#undef SIL_RUN
    // skey_out holds the sorted composite keys
    sil_final_kernel<<<fblocks, 256, 0, st>>>(S, reinterpret_cast<const int32_t*>(skey_out), gstart, Ca, Cb, i0, i1, part_a, part_b);
    SCM_LAUNCH_CHECK(ctx);
    std::vector<double> ha(fblocks), hb(fblocks);
    SCM_CUDA(cudaMemcpyAsync(ha.data(), part_a, sizeof(double) * fblocks, cudaMemcpyDeviceToHost, st));
    SCM_CUDA(cudaMemcpyAsync(hb.data(), part_b, sizeof(double) * fblocks, cudaMemcpyDeviceToHost, st));
    SCM_CUDA(cudaStreamSynchronize(st));
    double ta = 0.0, tb = 0.0;
    for (unsigned b = 0; b < fblocks; b++) { ta += ha[b]; tb += hb[b]; }
    *h_sum_a = ta;
    if (h_sum_b) *h_sum_b = tb;
    return 0;
}

}  // namespace scm
