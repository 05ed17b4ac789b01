// Numerics of scReplicate's replicate finding: what is arithmetic in R/scReplicate.R once the
// clustering / graph logic (kmeans, igraph: out of scope, they stay R) is taken away.
//
//   centroidDist        R/scReplicate.R:458-464   per-gene MEDIAN of a cluster's cells, squared distance of every cell to it
//   compute_dist_mat /  R/scReplicate.R:470-520   pairwise cell-cell distances between two batches (proxyC: correlation, cosine,
//   compute_dist_res                              Euclidean) and the MEDIAN of every (cluster of batch 1) x (cluster of batch 2) block
//   computePCA_byHVGMarker  :793-819              runPCA(rank 10, centred, scaled) of one batch on its HVG / marker genes
//                                                 (scm_gather + scm_centred_gram + scm_pca_scores)
//
// Medians are order statistics, not arithmetic: values are brought into contiguous segments and ordered with CUB's segmented
// radix sort (index plumbing, like the key sort of K1); everything that computes -- distances, the correlation GEMM on the DMMA
// kernel, the row statistics -- is this repository's code.
#pragma once
#include <cub/device/device_segmented_sort.cuh>

#include "common.cuh"
#include "gemm_f64.cuh"

namespace scm {

// The radix order of doubles puts a NaN with a clear sign bit after +Inf and one with the sign bit set before -Inf: every value
// that enters a segmented sort goes through this, so that "NaNs are the tail of the sorted segment" holds for any NaN payload.
__device__ __forceinline__ double canon_nan(double v) { return v != v ? __longlong_as_double(0x7ff8000000000000LL) : v; }

// out[i, j] = Y[rows[i], cols[j]]   (rows / cols NULL = identity): `exprs_mat[HVG, batch == b]` of R/scReplicate.R:398-404,808-811
__global__ void gather_kernel(const double* __restrict__ Y, int64_t ld, const int32_t* __restrict__ rows, int64_t msub,
                              const int32_t* __restrict__ cols, int64_t nsub, double* __restrict__ out, int64_t ldo) {
    for (int64_t i = blockIdx.x; i < msub; i += gridDim.x) {
        const double* src = Y + (int64_t)(rows ? rows[i] : i) * ld;
        for (int64_t j = threadIdx.x; j < nsub; j += blockDim.x) out[i * ldo + j] = src[cols ? cols[j] : j];
    }
}

// seg[(g * nb + jj) segment][position] = X[perm[start_g + position], j0 + jj]: the values of gene j0 + jj over the cells of
// group g, contiguous (the transposing gather in front of the segmented sort).  One CTA per (row of the sorted order).
__global__ void median_scatter_kernel(const double* __restrict__ X, int64_t ld, const int32_t* __restrict__ perm,
                                      const int64_t* __restrict__ seg_start, const int32_t* __restrict__ gkey, int32_t G, int64_t mk,
                                      int64_t j0, int nb, const int64_t* __restrict__ seg_off, double* __restrict__ seg) {
    for (int64_t pos = blockIdx.x; pos < mk; pos += gridDim.x) {
        const int32_t row = perm[pos];
        const int32_t g = gkey[row];
        const int64_t within = pos - seg_start[g];
        const int64_t cnt = seg_start[g + 1] - seg_start[g];
        const double* src = X + (int64_t)row * ld + j0;
        for (int jj = threadIdx.x; jj < nb; jj += blockDim.x) seg[seg_off[g] * nb + (int64_t)jj * cnt + within] = canon_nan(src[jj]);
    }
}

// med[g, j0 + jj] = median of the sorted segment (NaNs, sorted last, are ignored: median(..., na.rm = TRUE)); NaN for an empty one
__global__ void median_pick_kernel(const double* __restrict__ sorted, const int64_t* __restrict__ begin, const int64_t* __restrict__ end,
                                   int64_t nseg, double* __restrict__ out, int64_t out_stride_seg) {
    const int64_t sgi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sgi >= nseg) return;
    const double* p = sorted + begin[sgi];
    int64_t c = end[sgi] - begin[sgi];
    int64_t lo = 0, hi = c;  // first NaN (radix order puts NaN after +Inf)
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (p[mid] != p[mid]) hi = mid; else lo = mid + 1; }
    c = lo;
    double v = __longlong_as_double(0x7ff8000000000000LL);
    if (c > 0) v = (c & 1) ? p[c / 2] : 0.5 * (p[c / 2 - 1] + p[c / 2]);  // stats::median: mean of the two middle values
    out[sgi * out_stride_seg] = v;
}

// dist[i] = sum_j (X[i, j] - center[key_i, j])^2   (R/scReplicate.R:460); rows with key outside [0, G) get NaN.  Warp per row.
__global__ void __launch_bounds__(256)
sqdist_center_kernel(const double* __restrict__ X, int64_t m, int64_t n, int64_t ld, const int32_t* __restrict__ key, int32_t G,
                     const double* __restrict__ center, double* __restrict__ dist) {
    const int lane = threadIdx.x & 31;
    for (int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); i < m; i += (int64_t)gridDim.x * 8) {
        const int32_t g = key[i];
        double a = 0.0;
        if (g >= 0 && g < G) {
            const double* c = center + (int64_t)g * n;
            for (int64_t j = lane; j < n; j += 32) { const double d = X[i * ld + j] - c[j]; a += d * d; }
        }
        a = warp_sum(a);
        if (lane == 0) dist[i] = (g >= 0 && g < G) ? a : __longlong_as_double(0x7ff8000000000000LL);
    }
}

// Row preparation for the pairwise distances: out[i, :] = (X[i, :] - mean_i * centre) * scale_i and sq[i] = ||X[i, :]||^2.
//   metric 0 Euclidean  : out = X,                    D = sqrt(max(sq_a + sq_b - 2 a.b, 0))            (proxyC::dist)
//   metric 1 cosine     : out = X / ||X||,            D = 1 - a'.b'                                     (proxyC::simil "cosine")
//   metric 2 correlation: out = (X - mean) / ||X - mean||,  D = 1 - a'.b'                               (proxyC::simil "correlation")
// A zero row (cosine) / constant row (correlation) has no direction: its scaled row is NaN, like the 0 / 0 of the reference.
__global__ void __launch_bounds__(256)
rowprep_kernel(const double* __restrict__ X, int64_t m, int64_t n, int64_t ld, int metric, double* __restrict__ out, int64_t ldo,
               double* __restrict__ sq) {
    const int lane = threadIdx.x & 31;
    for (int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); i < m; i += (int64_t)gridDim.x * 8) {
        const double* x = X + i * ld;
        double s = 0.0;
        if (metric == 2) {
            for (int64_t j = lane; j < n; j += 32) s += x[j];
            s = warp_sum(s) / (double)n;
        }
        double q = 0.0;
        for (int64_t j = lane; j < n; j += 32) { const double d = x[j] - s; q += d * d; }
        q = warp_sum(q);
        const double sc = metric == 0 ? 1.0 : 1.0 / sqrt(q);  // q == 0 -> Inf, (x - s) * Inf = NaN: the undefined direction
        for (int64_t j = lane; j < n; j += 32) out[i * ldo + j] = metric == 0 ? x[j] : (x[j] - s) * sc;
        if (lane == 0 && sq) sq[i] = q;
    }
}

// D[i, j] <- distance from the inner product a_i . b_j left there by the GEMM
__global__ void dist_epilogue_kernel(double* __restrict__ D, int64_t m1, int64_t m2, int64_t ldd, int metric, const double* __restrict__ sqa,
                                     const double* __restrict__ sqb) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m2) return;
    for (int64_t i = blockIdx.y; i < m1; i += gridDim.y) {
        const double g = D[i * ldd + j];
        D[i * ldd + j] = metric == 0 ? sqrt(fmax(sqa[i] + sqb[j] - 2.0 * g, 0.0)) : 1.0 - g;
    }
}

// seg[off[bi * K2 + bj] + (i - r0) * (c1 - c0) + (j - c0)] = D[i, j] for the rectangle [r0, r1) x [c0, c1) of block (bi, bj)
__global__ void block_scatter_kernel(const double* __restrict__ D, int64_t ldd, const int64_t* __restrict__ row_off, int K1,
                                     const int64_t* __restrict__ col_off, int K2, const int64_t* __restrict__ seg_off,
                                     double* __restrict__ seg) {
    const int bi = blockIdx.y / K2, bj = blockIdx.y % K2;
    const int64_t r0 = row_off[bi], r1 = row_off[bi + 1], c0 = col_off[bj], c1 = col_off[bj + 1];
    const int64_t w = c1 - c0, total = (r1 - r0) * w;
    double* dst = seg + seg_off[bi * K2 + bj];
    (void)total;
    for (int64_t i = blockIdx.x; i < r1 - r0; i += gridDim.x)  // a CTA per row of the rectangle: coalesced, no index division
        for (int64_t j = threadIdx.x; j < w; j += blockDim.x) dst[i * w + j] = canon_nan(D[(r0 + i) * ldd + c0 + j]);
}

// Sorts `nseg` segments of d_in ([begin, end) offsets on the device) into d_out.  CUB limits one call to 2^31 - 1 items.
inline int segmented_sort(scm_ctx* ctx, Scratch& ws, const double* d_in, double* d_out, int64_t items, int64_t nseg,
                          const int64_t* d_begin, const int64_t* d_end) {
    if (items <= 0 || nseg <= 0) return 0;
    if (items >= ((int64_t)1 << 31) || nseg >= ((int64_t)1 << 31)) return fail(SCM_ERR_UNSUPPORTED, "segmented sort of %lld items", (long long)items);
    size_t tmp_bytes = 0;
    SCM_CUDA(cub::DeviceSegmentedSort::SortKeys(nullptr, tmp_bytes, d_in, d_out, (int)items, (int)nseg, d_begin, d_end, ctx->stream));
    char* tmp;
    SCM_TRY(ws.alloc(&tmp, (int64_t)tmp_bytes + 16));
    SCM_CUDA(cub::DeviceSegmentedSort::SortKeys(tmp, tmp_bytes, d_in, d_out, (int)items, (int)nseg, d_begin, d_end, ctx->stream));
    ctx->launches += 3;  // CUB's own kernels (partition + small / large segment sorts): counted from its structure, not observed
    return 0;
}

inline int gather(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, const int32_t* d_rows, int64_t msub,
                  const int32_t* d_cols, int64_t nsub, double* d_out, int64_t ldo) {
    if (!d_Y || !d_out || msub < 0 || nsub <= 0 || ldo < nsub || ld < n) return fail(SCM_ERR_INVALID, "scm_gather: bad arguments");
    if (msub == 0) return 0;
    gather_kernel<<<(unsigned)(msub < 65535 * 8 ? msub : 65535 * 8), 256, 0, ctx->stream>>>(d_Y, ld, d_rows, msub, d_cols, nsub, d_out, ldo);
    SCM_LAUNCH_CHECK(ctx);
    (void)m;
    return 0;
}

// med[g, j] = median over the rows with key g of X[:, j]  (G x n, row-major; NaN for an empty group).  h_key: the keys on the HOST
// (the segment layout is index plumbing and is built there), d_key the same on the device.
inline int group_median(scm_ctx* ctx, const double* d_X, int64_t m, int64_t n, int64_t ld, const int32_t* h_key, const int32_t* d_key,
                        int32_t G, double* d_med) {
    if (!d_X || !h_key || !d_key || !d_med || m <= 0 || n <= 0 || G <= 0 || ld < n) return fail(SCM_ERR_INVALID, "scm_group_median: bad arguments");
    if (m >= ((int64_t)1 << 31)) return fail(SCM_ERR_UNSUPPORTED, "scm_group_median: m >= 2^31");
    cudaStream_t st = ctx->stream;
    Scratch ws(ctx);
    // rows ordered by key (stable counting sort on the host: O(m) index work)
    std::vector<int64_t> start(G + 1, 0);
    for (int64_t i = 0; i < m; i++) if (h_key[i] >= 0 && h_key[i] < G) start[h_key[i] + 1]++;
    for (int g = 0; g < G; g++) start[g + 1] += start[g];
    const int64_t mk = start[G];
    std::vector<int32_t> perm(mk > 0 ? mk : 1);
    {
        std::vector<int64_t> cur(start.begin(), start.end() - 1);
        for (int64_t i = 0; i < m; i++) if (h_key[i] >= 0 && h_key[i] < G) perm[cur[h_key[i]]++] = (int32_t)i;
    }
    // gene blocks of nb columns so that the two segment buffers stay below ~2 GB (and one CUB call below 2^31 items)
    int64_t nb = mk > 0 ? ((int64_t)1 << 27) / mk : n;
    if (nb < 1) nb = 1;
    if (nb > n) nb = n;
    if (nb > 1024) nb = 1024;
    std::vector<int64_t> hbeg((size_t)G * nb), hend((size_t)G * nb);
    for (int g = 0; g < G; g++)
        for (int64_t jj = 0; jj < nb; jj++) {
            const int64_t cnt = start[g + 1] - start[g];
            hbeg[g * nb + jj] = start[g] * nb + jj * cnt;
            hend[g * nb + jj] = hbeg[g * nb + jj] + cnt;
        }
    int32_t* dperm; int64_t *dstart, *dbeg, *dend;
    double *seg_in, *seg_out;
    SCM_TRY(ws.alloc(&dperm, mk > 0 ? mk : 1)); SCM_TRY(ws.alloc(&dstart, G + 1));
    SCM_TRY(ws.alloc(&dbeg, (int64_t)G * nb)); SCM_TRY(ws.alloc(&dend, (int64_t)G * nb));
    SCM_TRY(ws.alloc(&seg_in, (mk > 0 ? mk : 1) * nb)); SCM_TRY(ws.alloc(&seg_out, (mk > 0 ? mk : 1) * nb));
    SCM_CUDA(cudaMemcpyAsync(dperm, perm.data(), sizeof(int32_t) * (mk > 0 ? mk : 1), cudaMemcpyHostToDevice, st));
    SCM_CUDA(cudaMemcpyAsync(dstart, start.data(), sizeof(int64_t) * (G + 1), cudaMemcpyHostToDevice, st));
    SCM_CUDA(cudaMemcpyAsync(dbeg, hbeg.data(), sizeof(int64_t) * G * nb, cudaMemcpyHostToDevice, st));
    SCM_CUDA(cudaMemcpyAsync(dend, hend.data(), sizeof(int64_t) * G * nb, cudaMemcpyHostToDevice, st));
    for (int64_t j0 = 0; j0 < n; j0 += nb) {
        const int nbc = (int)(n - j0 < nb ? n - j0 : nb);
        if (mk > 0) {
            median_scatter_kernel<<<(unsigned)(mk < 65535 * 16 ? mk : 65535 * 16), 128, 0, st>>>(d_X, ld, dperm, dstart, d_key, G, mk, j0, nbc,
                                                                                              dstart, seg_in);
            SCM_LAUNCH_CHECK(ctx);
            if (nbc == nb) {
                SCM_TRY(segmented_sort(ctx, ws, seg_in, seg_out, mk * nb, (int64_t)G * nb, dbeg, dend));
            } else {
                // last, narrower block: its own segment table
                std::vector<int64_t> b2((size_t)G * nbc), e2((size_t)G * nbc);
                for (int g = 0; g < G; g++)
                    for (int jj = 0; jj < nbc; jj++) {
                        const int64_t cnt = start[g + 1] - start[g];
                        b2[g * nbc + jj] = start[g] * nbc + jj * cnt;
                        e2[g * nbc + jj] = b2[g * nbc + jj] + cnt;
                    }
                SCM_CUDA(cudaMemcpyAsync(dbeg, b2.data(), sizeof(int64_t) * G * nbc, cudaMemcpyHostToDevice, st));
                SCM_CUDA(cudaMemcpyAsync(dend, e2.data(), sizeof(int64_t) * G * nbc, cudaMemcpyHostToDevice, st));
                SCM_CUDA(cudaStreamSynchronize(st));  // b2 / e2 are stack-scoped
                SCM_TRY(segmented_sort(ctx, ws, seg_in, seg_out, mk * nbc, (int64_t)G * nbc, dbeg, dend));
            }
        }
        for (int g = 0; g < G; g++) {
            // segment (g, jj) -> med[g, j0 + jj]: consecutive segments of a group are consecutive genes
            median_pick_kernel<<<(unsigned)ceil_div(nbc, 128), 128, 0, st>>>(seg_out, dbeg + (int64_t)g * nbc, dend + (int64_t)g * nbc, nbc,
                                                                            d_med + (int64_t)g * n + j0, 1);
            SCM_LAUNCH_CHECK(ctx);
        }
    }
    SCM_CUDA(cudaStreamSynchronize(st));  // the host vectors above back asynchronous copies
    return 0;
}

inline int sqdist_to_center(scm_ctx* ctx, const double* d_X, int64_t m, int64_t n, int64_t ld, const int32_t* d_key, int32_t G,
                            const double* d_center, double* d_dist) {
    if (!d_X || !d_key || !d_center || !d_dist || m < 0 || n <= 0 || ld < n) return fail(SCM_ERR_INVALID, "scm_sqdist_to_center: bad arguments");
    if (m == 0) return 0;
    const int64_t blocks = ceil_div(m, 8);
    sqdist_center_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, ctx->stream>>>(d_X, m, n, ld, d_key, G, d_center, d_dist);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

// D (m1 x m2, row-major, leading dimension ldd) = pairwise distances between the rows of A (m1 x n) and B (m2 x n).
inline int pair_dist(scm_ctx* ctx, const double* d_A, int64_t m1, int64_t lda, const double* d_B, int64_t m2, int64_t ldb, int64_t n,
                     int metric, double* d_D, int64_t ldd) {
    if (!d_A || !d_B || !d_D || m1 <= 0 || m2 <= 0 || n <= 0 || lda < n || ldb < n || ldd < m2 || metric < 0 || metric > 2)
        return fail(SCM_ERR_INVALID, "scm_pair_dist: bad arguments");
    if (m1 >= ((int64_t)1 << 31) || m2 >= ((int64_t)1 << 31) || n >= ((int64_t)1 << 31)) return fail(SCM_ERR_UNSUPPORTED, "scm_pair_dist: too large");
    cudaStream_t st = ctx->stream;
    Scratch ws(ctx);
    double *An, *Bn, *sqa, *sqb;
    SCM_TRY(ws.alloc(&An, m1 * n)); SCM_TRY(ws.alloc(&Bn, m2 * n)); SCM_TRY(ws.alloc(&sqa, m1)); SCM_TRY(ws.alloc(&sqb, m2));
    rowprep_kernel<<<(unsigned)(ceil_div(m1, 8) < 148 * 16 ? ceil_div(m1, 8) : 148 * 16), 256, 0, st>>>(d_A, m1, n, lda, metric, An, n, sqa);
    SCM_LAUNCH_CHECK(ctx);
    rowprep_kernel<<<(unsigned)(ceil_div(m2, 8) < 148 * 16 ? ceil_div(m2, 8) : 148 * 16), 256, 0, st>>>(d_B, m2, n, ldb, metric, Bn, n, sqb);
    SCM_LAUNCH_CHECK(ctx);
    SCM_TRY(gemm(ctx, (int)m1, (int)m2, (int)n, 1.0, rowmajor(An, n), transposed(Bn, n), 0.0, d_D, ldd));  // inner products on the DMMA GEMM
    dim3 grid((unsigned)ceil_div(m2, 256), (unsigned)(m1 < 65535 ? m1 : 65535));
    dist_epilogue_kernel<<<grid, 256, 0, st>>>(d_D, m1, m2, ldd, metric, sqa, sqb);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

// h_med[bi * K2 + bj] = median (NaNs ignored) of D[row_off[bi] : row_off[bi + 1], col_off[bj] : col_off[bj + 1]]: the rows / columns
// of D are ordered by cluster, so every (cluster, cluster) block is a rectangle (R/scReplicate.R:507-511).
inline int block_median(scm_ctx* ctx, const double* d_D, int64_t ldd, const int64_t* h_row_off, int32_t K1, const int64_t* h_col_off,
                        int32_t K2, double* h_med) {
    if (!d_D || !h_row_off || !h_col_off || !h_med || K1 <= 0 || K2 <= 0) return fail(SCM_ERR_INVALID, "scm_block_median: bad arguments");
    cudaStream_t st = ctx->stream;
    Scratch ws(ctx);
    const int64_t nseg = (int64_t)K1 * K2;
    std::vector<int64_t> beg(nseg), end(nseg);
    int64_t total = 0;
    for (int bi = 0; bi < K1; bi++)
        for (int bj = 0; bj < K2; bj++) {
            const int64_t c = (h_row_off[bi + 1] - h_row_off[bi]) * (h_col_off[bj + 1] - h_col_off[bj]);
            if (c < 0) return fail(SCM_ERR_INVALID, "scm_block_median: offsets must be non-decreasing");
            beg[bi * K2 + bj] = total; total += c; end[bi * K2 + bj] = total;
        }
    if (total >= ((int64_t)1 << 31)) return fail(SCM_ERR_UNSUPPORTED, "scm_block_median: %lld distances in one call (limit 2^31 - 1): split the batch pair", (long long)total);
    int64_t *drow, *dcol, *dbeg, *dend;
    double *seg_in, *seg_out, *dmed;
    SCM_TRY(ws.alloc(&drow, K1 + 1)); SCM_TRY(ws.alloc(&dcol, K2 + 1)); SCM_TRY(ws.alloc(&dbeg, nseg)); SCM_TRY(ws.alloc(&dend, nseg));
    SCM_TRY(ws.alloc(&seg_in, total > 0 ? total : 1)); SCM_TRY(ws.alloc(&seg_out, total > 0 ? total : 1)); SCM_TRY(ws.alloc(&dmed, nseg));
    SCM_CUDA(cudaMemcpyAsync(drow, h_row_off, sizeof(int64_t) * (K1 + 1), cudaMemcpyHostToDevice, st));
    SCM_CUDA(cudaMemcpyAsync(dcol, h_col_off, sizeof(int64_t) * (K2 + 1), cudaMemcpyHostToDevice, st));
    SCM_CUDA(cudaMemcpyAsync(dbeg, beg.data(), sizeof(int64_t) * nseg, cudaMemcpyHostToDevice, st));
    SCM_CUDA(cudaMemcpyAsync(dend, end.data(), sizeof(int64_t) * nseg, cudaMemcpyHostToDevice, st));
    if (total > 0) {
        dim3 grid(148 * 4, (unsigned)nseg);
        block_scatter_kernel<<<grid, 256, 0, st>>>(d_D, ldd, drow, K1, dcol, K2, dbeg, seg_in);
        SCM_LAUNCH_CHECK(ctx);
        SCM_TRY(segmented_sort(ctx, ws, seg_in, seg_out, total, nseg, dbeg, dend));
    }
    median_pick_kernel<<<(unsigned)ceil_div(nseg, 128), 128, 0, st>>>(seg_out, dbeg, dend, nseg, dmed, 1);
    SCM_LAUNCH_CHECK(ctx);
    SCM_CUDA(cudaMemcpyAsync(h_med, dmed, sizeof(double) * nseg, cudaMemcpyDeviceToHost, st));
    SCM_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // namespace scm
