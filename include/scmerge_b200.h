/*
 * scmerge_b200 -- C ABI of the B200-native RUV-III fit-and-adjust path.
 *
 * The reference (SydneyBioX/scMerge v1.19.0) is pure R and has NO FFI: the boundary it exposes is a
 * set of R closures.  This header is the C-ABI a `.Call` shim (r-package/src/rshim.c)
 * binds to replace the arithmetic inside those closures.  Each entry point cites the reference lines
 * whose work it takes over (paths relative to the reference repository root).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; `scm_last_error()` returns a thread-local message.
 *   - no exceptions, no C++/torch types cross the ABI: plain pointers and sizes only.
 *   - "d_" arguments are DEVICE pointers (owned by the caller, e.g. from scm_dev_alloc); "h_" are HOST pointers.
 *   - the canonical device layout of an expression matrix is ROW-MAJOR cells x genes (`Y[cell*ld + gene]`),
 *     which is byte-identical to R's column-major genes x cells (what scMerge()/scMerge2() hold,
 *     R/scMerge.R:91, R/scMerge2.R:298).  R's cells x genes column-major (a direct fastRUVIII(Y) call)
 *     is converted on the device by scm_transpose().
 *   - all arithmetic is IEEE float64; replicate / batch / pseudobulk keys are int32, 0-based, -1 = "skip row".
 *   - all work is enqueued on the context's stream; functions with host outputs synchronise that stream.
 *   - there is no CPU fallback: every entry point fails with SCM_ERR_CUDA if no sm_100 device is usable.
 *   - process model: either ONE process per GPU (one torchrun rank per device; cross-rank sums through the built-in NCCL
 *     communicator, scm_ctx_comm_init_nccl, or a host-supplied hook, scm_ctx_set_allreduce), or ONE process driving several GPUs
 *     (R's single thread: scm_multi_create + scm_multi_ruv3_host; ncclCommInitAll and one host thread per device inside the
 *     library).  A context belongs to one device; contexts on different devices may live in one process.  Calls on one context
 *     must not overlap (no internal locking); different contexts may be used from different threads.
 */
#ifndef SCMERGE_B200_H
#define SCMERGE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCM_OK 0
#define SCM_ERR_INVALID (-1)     /* bad argument */
#define SCM_ERR_CUDA (-2)        /* CUDA runtime / launch failure */
#define SCM_ERR_NONFINITE (-3)   /* NA / Inf in the input (R/fastRUVIII.R:52-60 stop()) */
#define SCM_ERR_SINGULAR (-4)    /* ac %*% t(ac) not invertible (R `solve()` error, R/fastRUVIII.R:99) */
#define SCM_ERR_UNSUPPORTED (-5) /* shape outside what the device path implements */
#define SCM_ERR_NOCONV (-6)      /* eigensolver did not reach the requested tolerance */
#define SCM_ERR_COMM (-7)        /* the host-supplied all-reduce hook failed */

#define SCM_MODE_EXACT 0      /* BSPARAM = ExactParam(): subspace iteration converged to `tol` */
#define SCM_MODE_RANDOMIZED 1 /* BSPARAM = RandomParam(): rsvd-style, oversampling p, q power iterations */

#define SCM_MAX_K 64     /* largest ruvK the fused adjust kernel is instantiated for */
#define SCM_MAX_BLOCK 112 /* largest subspace block (s + oversampling) of the single-CTA Rayleigh-Ritz solver */
#define SCM_MAX_TOPK_ROWS 88      /* fits returning more rows of fullalpha than this take the whole spectrum (block Jacobi) */
#define SCM_MAX_FULL_EIG_N 4096   /* largest Gram the whole-spectrum solver is instantiated for */

typedef struct scm_ctx scm_ctx;
typedef struct scm_coef scm_coef;
typedef struct scm_multi scm_multi;

/* Host-supplied in-place sum all-reduce over the ranks that share the cells (one process per GPU).
 * `dtype`: 0 = float64, 1 = int64.  Must be enqueued on (or ordered after) `stream`.  Return 0 on success. */
typedef int (*scm_allreduce_fn)(void* d_buf, int64_t count, int dtype, void* stream, void* user);

const char* scm_last_error(void);
int scm_version(void);

/* ---- context, memory, layout ------------------------------------------------------------------------ */
/* PCI bus id ("0000:3b:00.0") of a device: lets the host bind its threads / page-locked buffers to the NUMA node the GPU hangs
 * on (sysfs local_cpulist) before scm_host_alloc_pinned -- cross-socket staging halves the end-to-end rate on 2-socket hosts. */
int scm_device_pci_bus_id(int device, char* buf, int len);
int scm_ctx_create(int device, void* cuda_stream /* NULL = own stream; (void*)1 = the legacy default stream; (void*)2 = own stream of the LEAST priority (background context) */, scm_ctx** out);
int scm_ctx_destroy(scm_ctx* ctx);
int scm_ctx_set_allreduce(scm_ctx* ctx, scm_allreduce_fn fn, void* user, int world_size);
/* Built-in communicator: NCCL, resolved with dlopen("libnccl.so.2") at first use (no link-time dependency; SCM_ERR_COMM when it
 * cannot be loaded).  One process per GPU: rank 0 obtains the 128-byte unique id (scm_nccl_unique_id), the host hands it to the
 * other ranks by any means (a file, MPI, torch.distributed, ...), and every rank calls scm_ctx_comm_init_nccl (collective).  The
 * communicator takes precedence over a hook and is destroyed with the context.  world_size == 1 resets to "no communication".
 * These replace the BiocParallel process pool of the reference's chunked adjust (R/scMerge2_utilis.R:104-124) as the way the
 * cells are spread over workers; the reference has no collective of its own. */
int scm_nccl_unique_id(void* id128 /* 128 bytes */);
int scm_ctx_comm_init_nccl(scm_ctx* ctx, const void* id128, int rank, int world_size);
/* transport: 0 none (single rank), 1 built-in NCCL, 2 host hook */
int scm_ctx_comm_info(scm_ctx* ctx, int* rank, int* world_size, int* transport);
/* In-place sum of a device buffer over the ranks through whichever transport the context has (dtype 0 = float64, 1 = int64);
 * enqueued on the context's stream.  The host-side gathers of the ruvK selection use it. */
int scm_allreduce_sum(scm_ctx* ctx, void* d_buf, int64_t count, int dtype);
/* Facts about the last eigen-solve on this context: Rayleigh-Ritz steps, the subspace-angle estimate it stopped at, `conv_rank`
 * = the largest r such that the first r rows of the returned fullalpha span a converged invariant subspace (angle estimate
 * <= 1e-8): re-using fullalpha with a ruvK above it (fullalpha= of R/fastRUVIII.R:83, getAdjustedMat R/scMerge2.R:399) is not
 * backed by the fit; `products` = Gram x block products after the first Rayleigh-Ritz step. */
int scm_ctx_fit_info(scm_ctx* ctx, int32_t* iters, double* resid, int32_t* conv_rank, int32_t* products);

/* ---- one process, several GPUs (what a single R process needs) ---------------------------------------
 * scm_multi_create: a context + stream per listed device (devices == NULL: 0 .. ndev-1) and one NCCL rank per device
 * (ncclCommInitAll).  scm_multi_ruv3_host: scm_ruv3_host with the cells split into contiguous, balanced row blocks, one host
 * thread per device, every block streamed over its own PCIe link; fit results are reported by device 0 (after the Gram
 * all-reduce they are bit-identical on all devices).  Same arguments and error behaviour as scm_ruv3_host.
 * scm_multi_ctx hands out the per-device context (device-resident calls on one shard; owned by the handle). */
int scm_multi_create(const int* devices, int ndev, scm_multi** out);
int scm_multi_destroy(scm_multi* mg);
int scm_multi_size(const scm_multi* mg);
scm_ctx* scm_multi_ctx(scm_multi* mg, int i);
int scm_multi_ruv3_host(scm_multi* mg, const double* h_Y, int64_t m, int64_t n, int64_t ldh, const int32_t* h_group, int32_t G,
                        const int32_t* h_batch, int32_t Bt, const uint8_t* h_ctl, int32_t k, int32_t s, int mode, double tol,
                        int32_t maxit, uint64_t seed, const double* h_fullalpha_in, int32_t s_in, int center_mode,
                        const double* h_center, double* h_out, int64_t ldo, double* h_fullalpha, double* h_sigma, double* h_mean,
                        double* h_sd, int32_t* iters_out, double* resid_out);

/* Additional subspace dimensions the EXACT eigensolver has to converge in the following fits (besides `kconv`): scRUVIII with
 * several ruvK (R/scRUVIII.R:64-75) re-uses ONE fullalpha for every k, so the top-k subspace must be accurate for each of
 * them.  nk = 0 resets. */
int scm_ctx_set_conv_ranks(scm_ctx* ctx, const int32_t* ks, int32_t nk);
int scm_ctx_sync(scm_ctx* ctx);
int scm_ctx_launch_count(scm_ctx* ctx, int64_t* n_launches); /* kernels launched by this library so far */
/* Scratch (Gram partials, the pre-shifted copy of the cells the fit's Gram reads: m x n doubles, ...) comes from the device's
 * stream-ordered memory pool and stays cached there between calls (a fit re-allocates the same sizes; plain cudaFree next to a
 * populated pool costs up to a second).  This hands the cached, unused blocks back to the driver, e.g. before another library
 * needs the memory.  Synchronises the context's stream. */
int scm_ctx_release_scratch(scm_ctx* ctx);

int scm_dev_alloc(scm_ctx* ctx, int64_t bytes, void** d_ptr);
int scm_dev_free(scm_ctx* ctx, void* d_ptr);
int scm_host_alloc_pinned(int64_t bytes, void** h_ptr);
int scm_host_free_pinned(void* h_ptr);
int scm_h2d(scm_ctx* ctx, void* d_dst, const void* h_src, int64_t bytes);
int scm_d2h(scm_ctx* ctx, void* h_dst, const void* d_src, int64_t bytes);
/* Strided host<->device matrix copy (cudaMemcpy2DAsync): `rows` rows of `row_bytes`, pitches in bytes.
 * kind: 0 = host->device, 1 = device->host (synchronises), 2 = device->device. */
int scm_copy2d(scm_ctx* ctx, void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t row_bytes,
               int64_t rows, int kind);
/* out[c*ldo + r] = in[r*ldi + c]; converts R cells x genes column-major into the canonical layout
 * (the free `t()` of R/scRUVIII.R:48, R/scMerge2.R:298-299,313). */
int scm_transpose(scm_ctx* ctx, const double* d_in, int64_t rows, int64_t cols, int64_t ldi, double* d_out, int64_t ldo);

/* ---- K7: per-cell cosine normalisation ---------------------------------------------
 * out[i, :] = Y[i, :] / max(||Y[i, :]||_2, min_norm); in place allowed.  batchelor::cosineNorm as called by
 * scMerge2 (R/scMerge2.R:193-198; un-vendored, published behaviour: each cell divided by its L2 norm, norms
 * floored at 1e-8). */
int scm_cosine_rows(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, double* d_out, int64_t ldo, double min_norm);

/* ---- K8: sparse ingest -------------------------------------------------------------
 * scMerge2 holds its matrix as a dgCMatrix of genes x cells (R/scMerge2.R:164-166), i.e. compressed by CELL: in the
 * canonical cells x genes layout every cell is a sparse row with offsets p[m+1] (@p widened to int64), 0-based gene
 * indices i[nnz] (@i) and values x[nnz] (@x).  Produces the dense cell rows on the device,
 *     out[cell, :] = scatter(x_cell) / (cosine ? max(||x_cell||_2, min_norm) : 1),
 * fusing batchelor::cosineNorm (R/scMerge2.R:193-198) into the densification.  Indices need not be sorted; equal
 * indices inside a cell must be adjacent (they are summed).  d_inv_norm (optional, m) receives the applied scale,
 * d_nonfinite (optional int32[1]) is set to 1 when a stored value is NA/NaN/Inf.  Fails with SCM_ERR_INVALID on a
 * corrupt index structure (offsets decreasing / outside [0, nnz], gene index outside [0, n)). */
int scm_csc_to_dense(scm_ctx* ctx, const int64_t* d_indptr, const int32_t* d_idx, const double* d_val, int64_t m, int64_t n,
                     int64_t nnz, int cosine, double min_norm, double* d_out, int64_t ldo, double* d_inv_norm, int32_t* d_nonfinite);

/* ---- sparse-native scMerge2 passes (the dense normalised matrix is never an input) ------------------------------------
 * For cells-compressed input with CANONICAL cells (gene indices strictly increasing inside every cell: every dgCMatrix) the
 * three passes over the cells read 12 bytes per non-zero instead of 8 bytes per entry:
 *   scm_csc_check        validates the index structure (SCM_ERR_INVALID when corrupt) and reports `canonical` (0/1)
 *   scm_csc_rownorm      d_inv_norm[i] = cosine ? 1 / max(||x_i||_2, min_norm) : 1     (batchelor::cosineNorm, R/scMerge2.R:193-198)
 *   scm_csc_seg_colmean  means[g, :] = mean over the cells with key g of inv_i x_i      (create_pseudoBulk, R/scMerge2_utilis.R:461-486;
 *                        G x ldn row-major with ldn >= n even, deterministic)
 *   scm_csc_adjust       out[i, :] = inv_i x_i - ((inv_i x_i - center) B) A, dense m x n (pseudoRUVIII's adjust of every cell,
 *                        R/scMerge2_utilis.R:97-137): sparse gather for W, DMMA rank-k update on shared-memory tiles of the scaled
 *                        sparse rows, newY written once.  k <= 32.
 * Un-canonical input: use scm_csc_to_dense and the dense kernels.  scm_csc_rownorm / _seg_colmean / _adjust trust the index
 * structure: call scm_csc_check once per matrix first (a gene index outside [0, n) would make scm_csc_adjust read out of bounds). */
int scm_csc_check(scm_ctx* ctx, const int64_t* d_indptr, const int32_t* d_idx, int64_t m, int64_t n, int64_t nnz, int32_t* canonical);
int scm_csc_rownorm(scm_ctx* ctx, const int64_t* d_indptr, const double* d_val, int64_t m, int cosine, double min_norm, double* d_inv_norm,
                    int32_t* d_nonfinite);
int scm_csc_seg_colmean(scm_ctx* ctx, const int64_t* d_indptr, const int32_t* d_idx, const double* d_val, const double* d_inv_norm,
                        int64_t m, int64_t n, const int32_t* d_key, int32_t G, double* d_means, int64_t ldn, int64_t* d_counts);
int scm_csc_adjust(scm_ctx* ctx, const int64_t* d_indptr, const int32_t* d_idx, const double* d_val, const double* d_inv_norm, int64_t m,
                   int64_t n, const scm_coef* coef, double* d_out, int64_t ldo);
/* scm_csc_adjust with the result streamed to the (page-locked) host matrix h_out chunk by chunk through two device staging
 * buffers while the next chunk is computed: the dense matrix never exists on the device.  Synchronises. */
int scm_csc_adjust_to_host(scm_ctx* ctx, const int64_t* d_indptr, const int32_t* d_idx, const double* d_val, const double* d_inv_norm,
                           int64_t m, int64_t n, const scm_coef* coef, double* h_out, int64_t ldo);

/* ---- K1: segmented column sums ---------------------------------------------------------------------
 * sums[g, j] = sum over rows i with key[i] == g of f(Y[i, j]),  counts[g] = #rows,
 *   f(y) = y                       (d_center == NULL)
 *   f(y) = (y - center[g, j])^2    (d_center != NULL)
 * Replaces the dense-indicator products of my_residop (R/fastRUVIII.R:121-129), the batch means and
 * pooled variance of standardize2 (R/scRUVIII.R:158-175), aggregate.Matrix / create_pseudoBulk
 * (R/scMerge2_utilis.R:325-346,461-486) and colMeans (R/scMerge2_utilis.R:97) with G = 1.
 * `d_nonfinite` (optional int32[1]) is set to 1 if any visited element is NA/NaN/Inf (R/fastRUVIII.R:52-60).
 * Deterministic: the result is bit-identical run to run for the same inputs.  LOCAL to this rank. */
int scm_seg_colsum(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld,
                   const int32_t* d_key, int32_t G, const double* d_center,
                   double* d_sums, int64_t* d_counts, int32_t* d_nonfinite);

/* Same pass, followed by the division by the group sizes on the device: means[g, :] = sums[g, :] / counts[g]
 * (rows of empty groups are 0).  This is `res / cellTypes_n_mat` of create_pseudoBulk (R/scMerge2_utilis.R:474-477)
 * and colMeans (R/scMerge2_utilis.R:97, R/scMerge2.R:393).  GLOBAL over the ranks of the context (sums and counts are
 * all-reduced before the division: a pseudobulk's cells may live on any rank); so is scm_csc_seg_colmean. */
int scm_seg_colmean(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld,
                    const int32_t* d_key, int32_t G, double* d_means, int64_t* d_counts, int32_t* d_nonfinite);

/* Column means of ALL cells recovered from group means and counts: out[j] = sum_g counts[g] means[g, j] / sum_g counts[g].
 * scMerge2 needs both the pseudobulk means and colMeans(Y) (R/scMerge2_utilis.R:97); this saves the second pass over Y. */
int scm_group_colmean(scm_ctx* ctx, const double* d_means, const int64_t* d_counts, int32_t G, int64_t n, int64_t ldm, double* d_out);

/* ---- K3: Gram of the shifted matrix -----------------------------------------------------------------
 * gram[n x n] = sum_i (y_i - shift)(y_i - shift)'   (shift may be NULL), full symmetric output.
 * FP64 tensor-core (DMMA) SYRK over lower-triangular 128x128 tiles, split over cells, deterministic.
 * Together with scm_gram_residual() this replaces Y0 <- my_residop(Y, M) + the Gram that
 * BiocSingular::runSVD's dgesdd implicitly works on (R/fastRUVIII.R:89-91).  LOCAL to this rank. */
int scm_gram(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, const double* d_shift, double* d_gram);

/* gram <- diag(scale) * ( gram - sum_g cnt_g (mu_g - shift)(mu_g - shift)' ) * diag(scale), mu_g = sums_g / cnt_g.
 * Turns the shifted Gram of Y into the Gram of the replicate residual Y0 (optionally of the standardised
 * matrix, scale = 1/sd, R/scRUVIII.R:171).  d_shift / d_scale may be NULL. */
int scm_gram_residual(scm_ctx* ctx, double* d_gram, int64_t n, const double* d_sums, const int64_t* d_counts,
                      int32_t G, const double* d_shift, const double* d_scale);

/* ---- K4: top-s eigenpairs of a symmetric PSD matrix -------------------------------------------------
 * Block subspace iteration (DMMA GEMMs, CholeskyQR2, single-CTA Jacobi Rayleigh-Ritz).
 *   mode EXACT      : iterate until the Davis-Kahan estimate of the angle between the computed and the true
 *                     top-`kconv` invariant subspace, ||R_kconv||_F / (theta_kconv - theta_kconv+1), is <= tol
 *                     (default 1e-10), at most maxit iterations; SCM_ERR_NOCONV (results still written) otherwise
 *   mode RANDOMIZED : rsvd semantics in gene space: oversampling p = 10, q = 2 power iterations, no test
 * d_vecs: s x n row-major (row r = eigenvector r, unit norm, largest |entry| positive), d_vals: s (descending).
 * Replaces BiocSingular::runSVD (R/fastRUVIII.R:91, R/scMerge2_utilis.R:84): with u the left singular
 * vectors of Y0, t(u) %*% Y == diag(sqrt(vals)) %*% vecs  (R/fastRUVIII.R:93). */
int scm_eig_topk(scm_ctx* ctx, const double* d_gram, int64_t n, int32_t s, int32_t kconv, int mode,
                 double tol, int32_t maxit, uint64_t seed, double* d_vals, double* d_vecs,
                 int32_t* iters_out, double* resid_out);

/* fullalpha[r, :] = sqrt(max(vals[r], 0)) * vecs[r, :] * colscale   (s x n; R/fastRUVIII.R:93) */
int scm_fullalpha(scm_ctx* ctx, const double* d_vals, const double* d_vecs, int32_t s, int64_t n, double* d_fullalpha);

/* ---- K5: adjustment coefficients --------------------------------------------------------------------
 * From alpha = first k rows of d_alpha (k x n, row-major, any row scaling/rotation of the factor basis:
 * fullalpha or unit eigenvectors), the control mask and the optional column vectors builds
 *   B = diag(pre) * 1[ctl] * ac' (ac ac')^-1   (n x k),   A = alpha * diag(post)   (k x n)
 * so that  newY = Y - ((Y - center) B) A.  fastRUVIII: center/pre/post NULL (R/fastRUVIII.R:97-101);
 * pseudoRUVIII/getAdjustedMat: center = colMeans (R/scMerge2_utilis.R:97-137, R/scMerge2.R:392-406);
 * scRUVIII: center = stand_mean, pre = 1/sd, post = sd (standardise + de-standardise fused, R/scRUVIII.R:47-55,117-119).
 * Fails with SCM_ERR_SINGULAR when ac ac' is not positive definite. */
int scm_coef_create(scm_ctx* ctx, const double* d_alpha, int64_t n, int32_t k, const uint8_t* d_ctl,
                    const double* d_center, const double* d_pre, const double* d_post, scm_coef** out);
/* Coefficient object from explicit factors B (n x k, row-major; rows of genes with d_rows[g] == 0 must be zero) and
 * A (k x n): newY = Y - ((Y - center) B) A.  Used by scRUVg (B = 1[ctl] V diag(1/sigma), A = alpha, R/scRUVg.R:77-80). */
int scm_coef_from_factors(scm_ctx* ctx, const double* d_B, const double* d_A, int64_t n, int32_t k, const uint8_t* d_rows,
                          const double* d_center, scm_coef** out);
int scm_coef_destroy(scm_ctx* ctx, scm_coef* coef);

/* ---- K6: fused rank-k adjustment --------------------------------------------------------------------
 * out[i, :] = Y[i, :] - ((Y[i, :] - center) B) A      (in place allowed: d_out == d_Y)
 * d_subset (int32[nsub], optional): only these gene columns are produced, out is m x nsub
 * (return_subset_genes, R/scMerge2_utilis.R:117-121, R/scMerge2.R:402-406).
 * d_W (optional, m x k row-major): emits W = (Y - center) B (R/fastRUVIII.R:99).  d_out == NULL with d_W given runs
 * the projection only.  Rank-local. */
int scm_adjust(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, const scm_coef* coef,
               const int32_t* d_subset, int32_t nsub, double* d_out, int64_t ldo, double* d_W);

/* ---- drivers (what the R closures call) -------------------------------------------------------------
 * scm_ruv3_fit: the fit half of fastRUVIII (R/fastRUVIII.R:46-93) on a device-resident, cell-sharded Y.
 *   group : replicate id per local cell (from ruv::replicate.matrix(M), one-hot rows), G groups in total
 *   batch : NULL for fastRUVIII / pseudoRUVIII; batch id per cell (Bt batches) for scRUVIII, which adds
 *           standardize2 (R/scRUVIII.R:158-175): stand_mean / stand_sd are returned in d_mean / d_sd (n each)
 *   s     : rows of fullalpha to return (host applies the svd_k rule R/fastRUVIII.R:66-70), kconv = min(k, s).
 *           s <= SCM_MAX_TOPK_ROWS: block subspace iteration, the top kconv rows converged (scm_ctx_fit_info).
 *           s  > SCM_MAX_TOPK_ROWS (the non-exact BSPARAM rule, R/fastRUVIII.R:69, R/scMerge2_utilis.R:64): the WHOLE
 *           spectrum by one-sided block Jacobi -- of the n x n Gram (n <= SCM_MAX_FULL_EIG_N), or, for a short matrix
 *           on one rank without batch (m_local < n, m_local <= SCM_MAX_FULL_EIG_N: pseudoRUVIII), of Y0 Y0' with
 *           fullalpha = t(u) %*% Y as the reference writes it (R/scMerge2_utilis.R:84-86).  Every row exact;
 *           SCM_ERR_UNSUPPORTED beyond those sizes.
 * Every local cell must carry a key in [0, G) (M assigns each cell to a replicate group; SCM_ERR_INVALID otherwise).
 * Cross-rank state is summed with TWO all-reduces per fit: [group sums | batch sums | counts | NA flag, rows] and
 * [n x n centred Gram | within-batch squared deviations] (built-in NCCL communicator or the hook).
 * Outputs: d_fullalpha (s x n), d_sigma (s singular values of Y0). */
int scm_ruv3_fit(scm_ctx* ctx, const double* d_Y, int64_t m_local, int64_t n, int64_t ld,
                 const int32_t* d_group, int32_t G, const int32_t* d_batch, int32_t Bt,
                 int32_t s, int32_t kconv, int mode, double tol, int32_t maxit, uint64_t seed,
                 double* d_fullalpha, double* d_sigma, double* d_mean, double* d_sd,
                 int32_t* iters_out, double* resid_out);

/* scm_ruv3_fit with two more outputs (either may be NULL): d_gram_centred (n x n) = sum over ALL cells of
 * (y - ybar)(y - ybar)' and d_colmean (n) = ybar, the inputs of scm_pca_scores (no extra pass over the cells). */
int scm_ruv3_fit_gram(scm_ctx* ctx, const double* d_Y, int64_t m_local, int64_t n, int64_t ld,
                      const int32_t* d_group, int32_t G, const int32_t* d_batch, int32_t Bt,
                      int32_t s, int32_t kconv, int mode, double tol, int32_t maxit, uint64_t seed,
                      double* d_fullalpha, double* d_sigma, double* d_mean, double* d_sd,
                      int32_t* iters_out, double* resid_out, double* d_gram_centred, double* d_colmean);

/* ---- choice of ruvK (scRUVIII with several k, R/scRUVIII.R:80-114,182-199) --------------------------------------------
 * scm_pca_scores: BiocSingular::runPCA(newY, rank, center = TRUE, scale = TRUE)$x (R/scRUVIII.R:183-184) of the ADJUSTED
 * matrix newY = Y - ((Y - center) B) A described by `coef` (NULL: of Y itself) WITHOUT materialising newY: its centred Gram
 * is (I - BA)' Gc (I - BA), expanded in the rank-k factors; after scaling to unit variance (n - 1 denominator) the top
 * `rank` eigenpairs (K4, converged to `tol`) give the scores with one projection pass (K6 pass A) over the local rows of Y:
 * d_scores (m_local x rank, row-major; columns defined up to sign), d_sing (rank, optional: singular values of the centred,
 * scaled matrix).  m_total = number of cells over all ranks, d_colmean = their column means (scm_ruv3_fit_gram). */
int scm_pca_scores(scm_ctx* ctx, const double* d_gram_centred, int64_t n, int64_t m_total, const scm_coef* coef, const double* d_colmean,
                   int32_t rank, const double* d_Y, int64_t m_local, int64_t ld, double* d_scores, double* d_sing, double tol,
                   int32_t maxit, int32_t* iters_out, double* resid_out);

/* scm_silhouette: kBET_batch_sil (R/scRUVIII.R:192-199) = summary(cluster::silhouette(labels, dist(x)))$avg.width, exact
 * (all m^2 Euclidean distances, O(m) memory).  d_X: m x d scores of ALL cells (d <= 16), labels 0-based in [0, C).  One
 * distance pass serves two labelings (cell type and batch; d_lab_b may be NULL).  Returns in h_sum_a / h_sum_b the SUM of
 * the widths s_i over the cells at positions [i0, i1) of the order by (label_a, label_b): ranks take disjoint ranges and
 * add the sums; avg.width = total / m.  s_i = 0 for singleton clusters (cluster::silhouette) and when fewer than two
 * clusters are populated (the reference would return NA). */
int scm_silhouette(scm_ctx* ctx, const double* d_X, int64_t m, int32_t d, int64_t ldx, const int32_t* d_lab_a, int32_t Ca,
                   const int32_t* d_lab_b, int32_t Cb, int64_t i0, int64_t i1, double* h_sum_a, double* h_sum_b);

/* scm_ruvg_fit: the fit half of scRUVg (R/scRUVg.R:30-85 with Z = 1, eta = NULL) on a device-resident,
 * cell-sharded Y: column means (my_residop(Y, 1), :61-63), the top-k left singular vectors of the centred control block
 * (:66-73) through the Gram of the centred cells and its control block (K3 + K4; the m x m cross product of the >= 5:1
 * aspect-ratio branch is never formed), alpha = W'Y0 (:78).  Outputs d_alpha (k x n), d_center (n), d_sigma (k, optional:
 * singular values of Y0[, ctl]) and the coefficient object: scm_adjust(..., coef, ..., d_W) then gives newY (:80) and W (:77).
 * Errors: SCM_ERR_INVALID when k exceeds the number of controls (:56-58), SCM_ERR_SINGULAR when the control block has rank < k. */
int scm_ruvg_fit(scm_ctx* ctx, const double* d_Y, int64_t m_local, int64_t n, int64_t ld, const uint8_t* d_ctl, int32_t k, int mode,
                 double tol, int32_t maxit, uint64_t seed, double* d_alpha, double* d_center, double* d_sigma, scm_coef** coef_out,
                 int32_t* iters_out, double* resid_out);
/* scRUVg with a GIVEN factor matrix W (R/scRUVg.R:64,77-80: a re-used `fullW[, 1:k]`), and the last step of scRUVg with eta, where W
 * comes from RUV1(Y) and is subtracted from Y itself:
 *   scm_ls_coef         alpha = solve(W'W) W' (Y - 1 center')  (k x n; :78; center = colMeans of the matrix, Z = 1), sums over the
 *                       cells all-reduced across ranks.  d_W: m_local x k row-major
 *   scm_rankk_subtract  out = Y - W A for given W (m x k) and A (k x n): pass B of the fused rank-k update alone (:80) */
int scm_ls_coef(scm_ctx* ctx, const double* d_Y, int64_t m_local, int64_t n, int64_t ld, const double* d_W, int32_t k,
                const double* d_center, double* d_alpha);
int scm_rankk_subtract(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, const double* d_W, int32_t k, const double* d_A,
                       double* d_out, int64_t ldo);

/* ---- replicate-finding numerics of scReplicate ----------------------------------------------------
 * What is arithmetic in R/scReplicate.R once kmeans / igraph (out of scope, they stay R) are taken away.
 *   scm_gather            out[i, j] = Y[rows[i], cols[j]] (NULL = identity): exprs_mat[HVG_list[[j]], batch == batch_list[j]]
 *                         (R/scReplicate.R:398-404,808-811)
 *   scm_group_median      med[g, j] = median over the cells with key g of X[:, j] (G x n; NaN for an empty group, NaNs ignored):
 *                         DelayedMatrixStats::rowMedians per cluster in centroidDist (R/scReplicate.R:459).  Keys on the host AND
 *                         on the device (the segment layout is index plumbing built on the host)
 *   scm_sqdist_to_center  dist[i] = sum_j (X[i, j] - center[key_i, j])^2: colSums((exprs_mat - centroid_batch)^2) (R/scReplicate.R:460)
 *   scm_pair_dist         D[i, j] = distance between row i of A and row j of B; metric 0 Euclidean (proxyC::dist), 1 one minus
 *                         cosine, 2 one minus Pearson correlation (proxyC::simil): compute_dist_mat (R/scReplicate.R:470-490,497-505).
 *                         Inner products on the FP64 DMMA GEMM; a zero / constant row gives NaN like the reference's 0 / 0
 *   scm_block_median      h_med[bi * K2 + bj] = median(D[rows of cluster bi, columns of cluster bj], na.rm = TRUE) with the rows /
 *                         columns of D ordered by cluster (offsets on the host): compute_dist_res (R/scReplicate.R:507-513)
 *   scm_centred_gram      centred Gram sum (y - ybar)(y - ybar)' and column means of ALL cells (cross-rank like scm_ruv3_fit): with
 *                         scm_pca_scores(coef = NULL) this is BiocSingular::runPCA(x, rank = 10, scale = TRUE, center = TRUE)$x of
 *                         computePCA_byHVGMarker (R/scReplicate.R:793-819) and of the kmeans input (:392-395)
 *   scm_kmeans            stats::kmeans(x, centers = K, nstart, algorithm = "Hartigan-Wong") of identifyCluster
 *                         (R/scReplicate.R:394-395: kmeans(pca_current[, 1:10], centers = kmeansK[j], nstart = 1000)): every
 *                         start draws K distinct rows (counter-based hash of (seed, start): R's RNG stream is not reproduced),
 *                         runs Lloyd iterations to a fixed point (at most iter_max), the start with the smallest total
 *                         within-cluster sum of squares wins and is finished with Hartigan-Wong single-point transfers until
 *                         none lowers the objective (the reference's stopping rule).  d_X: m x d, d <= 16.  d_label: cluster
 *                         ids in [0, K) (R: + 1); d_centers K x d; h_withinss / h_size K; h_info: [winning start, its Lloyd
 *                         iterations, transfers, 1 if the transfer cap was hit].  Deterministic for a given seed */
int scm_gather(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, const int32_t* d_rows, int64_t msub,
               const int32_t* d_cols, int64_t nsub, double* d_out, int64_t ldo);
int scm_group_median(scm_ctx* ctx, const double* d_X, int64_t m, int64_t n, int64_t ld, const int32_t* h_key, const int32_t* d_key,
                     int32_t G, double* d_med);
int scm_sqdist_to_center(scm_ctx* ctx, const double* d_X, int64_t m, int64_t n, int64_t ld, const int32_t* d_key, int32_t G,
                         const double* d_center, double* d_dist);
int scm_pair_dist(scm_ctx* ctx, const double* d_A, int64_t m1, int64_t lda, const double* d_B, int64_t m2, int64_t ldb, int64_t n,
                  int metric, double* d_D, int64_t ldd);
int scm_block_median(scm_ctx* ctx, const double* d_D, int64_t ldd, const int64_t* h_row_off, int32_t K1, const int64_t* h_col_off,
                     int32_t K2, double* h_med);
int scm_centred_gram(scm_ctx* ctx, const double* d_Y, int64_t m_local, int64_t n, int64_t ld, double* d_gram, double* d_colmean,
                     int64_t* m_total_out);
int scm_kmeans(scm_ctx* ctx, const double* d_X, int64_t m, int32_t d, int64_t ldx, int32_t K, int32_t nstart, int32_t iter_max, uint64_t seed,
               int32_t* d_label, double* d_centers, double* h_withinss, int64_t* h_size, double* h_tot_withinss, int32_t* h_info);

/* scm_ruv3_host: the whole closure on HOST matrices (what `.Call` hands over), with the PCIe copies overlapped with
 * the kernels: Y is uploaded in row chunks on a copy stream while the replicate sums and the Gram accumulate behind
 * it, then the adjust kernel runs chunk by chunk in place and every finished chunk is downloaded while the next is
 * computed.  Use page-locked host buffers (scm_host_alloc_pinned) for full PCIe speed.
 *   h_Y, h_out        : row-major cells x genes (== R genes x cells column-major), leading dimensions ldh / ldo
 *   h_batch != NULL   : scRUVIII (standardise + de-standardise fused, R/scRUVIII.R:41-137); else fastRUVIII
 *   h_fullalpha_in    : non-NULL (s_in x n) = adjust only (fullalpha re-use R/fastRUVIII.R:83, pseudoRUVIII /
 *                       getAdjustedMat R/scMerge2_utilis.R:97-137, R/scMerge2.R:392-406) with
 *                       center_mode 0 none, 1 column means of Y (computed during the upload), 2 h_center
 *   outputs (optional): h_fullalpha s x n, h_sigma s, h_mean n (stand_mean or the column means), h_sd n */
int scm_ruv3_host(scm_ctx* ctx, const double* h_Y, int64_t m, int64_t n, int64_t ldh, const int32_t* h_group, int32_t G,
                  const int32_t* h_batch, int32_t Bt, const uint8_t* h_ctl, int32_t k, int32_t s, int mode, double tol,
                  int32_t maxit, uint64_t seed, const double* h_fullalpha_in, int32_t s_in, int center_mode,
                  const double* h_center, double* h_out, int64_t ldo, double* h_fullalpha, double* h_sigma, double* h_mean,
                  double* h_sd, int32_t* iters_out, double* resid_out);

/* The download half of the host drivers for a DEVICE-resident matrix: K6 runs chunk by chunk IN PLACE on d_Y and every
 * finished chunk is copied to the (page-locked) host matrix h_out on a second stream while the next chunk is adjusted
 * (the chunked `bplapply` adjust + rbind of pseudoRUVIII, R/scMerge2_utilis.R:104-137, with the PCIe copy overlapped).
 * d_Y holds newY afterwards.  Synchronises. */
int scm_adjust_to_host(scm_ctx* ctx, double* d_Y, int64_t m, int64_t n, int64_t ld, const scm_coef* coef, double* h_out, int64_t ldo);

/* Stage timers (ms, CUDA events on the context stream) of the last scm_ruv3_fit / scm_adjust:
 * 0 seg_colsum (the summing pass; it also writes the pre-shifted copy of the cells the fit's Gram reads), 1 gram (SYRK + split
 * reduce + re-centring), 2 allreduce (the LAST collective), 3 eig, 4 coef, 5 adjust, 6 standardise (last pass), 7 the gram_syrk
 * kernel alone, 8 ingest (scm_csc_to_dense / scm_cosine_rows), 9 the pre-shifted copy as a separate pass (only with
 * SCM_K1_FUSED_COPY=0; part of 1). */
int scm_timers(scm_ctx* ctx, double* ms, int n);

#ifdef __cplusplus
}
#endif
#endif
