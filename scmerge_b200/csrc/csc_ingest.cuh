// K8: sparse ingest.  scMerge2 holds its expression matrix as a dgCMatrix (genes x cells, compressed by CELL:
// R/scMerge2.R:164-166 coerces dense input to CsparseMatrix), i.e. every cell is a sparse row of the canonical
// cells x genes layout: p[m+1] offsets, i[nnz] gene indices, x[nnz] values.  Shipping that over PCIe instead of the
// dense matrix cuts the upload by the sparsity factor (~10x for single-cell data); the dense cell rows the fit and
// the rank-k update stream are produced here, on the device, fused with batchelor::cosineNorm
// (R/scMerge2.R:193-198): out[i, :] = scatter(x_i) / max(||x_i||_2, min_norm).
//
// One CTA per cell (grid-stride): the cell's non-zeros are read once (12 bytes each), the norm is reduced, the scaled
// values are scattered into a zeroed shared-memory image of the row, which then leaves with 16-byte streaming stores, so
// HBM sees 12*nnz bytes read + 8*m*n written -- the algorithmic traffic.  Rows longer than the shared image are produced
// slab by slab (generic kernel).  Indices need not be sorted; duplicate gene indices inside a cell are summed.
#pragma once
#include "common.cuh"

namespace scm {

constexpr int CSC_SLAB = 4096;  // genes per shared-memory row image of the generic kernel (32 KB: 7 CTAs per SM)
constexpr int CSC_THREADS = 256;
constexpr int CSC_REG = 2;      // non-zeros per thread the fast kernel keeps in registers (cells with more loop over the rest)

// Fast path (n <= slab, i.e. the whole row fits one shared-memory image): the cell's non-zeros are read from global memory
// ONCE into registers (the loads of the next cell are issued before the current row is written out), the row image is
// double-buffered, and a cell costs two __syncthreads.  Requires unique indices inside a cell (csc_check_kernel verifies
// "strictly increasing", which every dgCMatrix satisfies); anything else takes the generic kernel, whose scatter sums
// duplicates with shared-memory atomics (a CAS loop for f64: ATOMS.CAST.SPIN in SASS, hence not on the fast path).
__global__ void __launch_bounds__(CSC_THREADS)
csc_densify_fast_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val, int64_t m,
                        int n, int slab, int cosine, double min_norm, double* __restrict__ out, int64_t ldo, int vec_out,
                        double* __restrict__ inv_norm_out, int32_t* __restrict__ flag) {
    extern __shared__ __align__(16) double csc_smem[];
    double* red = csc_smem + 2 * slab;  // [2][8]
    const int tid = threadIdx.x;
    int32_t ci[CSC_REG];
    double cv[CSC_REG];
    int64_t p0 = 0, p1 = 0;
    auto prefetch = [&](int64_t cell) {
        if (cell < m) {
            p0 = indptr[cell]; p1 = indptr[cell + 1];
#pragma unroll
            for (int r = 0; r < CSC_REG; r++) {
                const int64_t p = p0 + (int64_t)r * CSC_THREADS + tid;
                const bool ok = p < p1;
                ci[r] = ok ? idx[p] : -1;
                cv[r] = ok ? val[p] : 0.0;
            }
        }
    };
    int buf = 0;
    bool bad = false;
    prefetch(blockIdx.x);
    for (int64_t cell = blockIdx.x; cell < m; cell += gridDim.x, buf ^= 1) {
        double* row = csc_smem + buf * slab;
        const int64_t q0 = p0, q1 = p1;  // this cell's range (prefetch overwrites p0 / p1)
        for (int j = tid * 2; j < n; j += CSC_THREADS * 2) *reinterpret_cast<double2*>(row + j) = make_double2(0.0, 0.0);
        double ss = 0.0;
#pragma unroll
        for (int r = 0; r < CSC_REG; r++) { bad |= nonfinite(cv[r]); ss += cv[r] * cv[r]; }
        for (int64_t p = q0 + (int64_t)CSC_REG * CSC_THREADS + tid; p < q1; p += CSC_THREADS) {  // rare: > 512 non-zeros
            const double v = val[p];
            bad |= nonfinite(v);
            ss += v * v;
        }
        ss = warp_sum(ss);
        if ((tid & 31) == 0) red[buf * 8 + (tid >> 5)] = ss;
        __syncthreads();  // row image zeroed, partial norms visible
        double inv = 1.0;
        if (cosine) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < CSC_THREADS / 32; w++) t += red[buf * 8 + w];  // fixed order: deterministic
            inv = 1.0 / fmax(sqrt(t), min_norm);
            if (tid == 0 && inv_norm_out) inv_norm_out[cell] = inv;
        }
#pragma unroll
        for (int r = 0; r < CSC_REG; r++)
            if (ci[r] >= 0) row[ci[r]] = cv[r] * inv;  // canonical input (checked by the caller): indices are unique
        for (int64_t p = q0 + (int64_t)CSC_REG * CSC_THREADS + tid; p < q1; p += CSC_THREADS) row[idx[p]] = val[p] * inv;
        __syncthreads();  // scatter complete
        prefetch(cell + gridDim.x);  // next cell's loads fly while this row is written out
        double* dst = out + cell * ldo;
        if (vec_out) {
            for (int j = tid * 2; j < n; j += CSC_THREADS * 2) {
                if (j + 1 < n) stg_stream2(dst + j, *reinterpret_cast<const double2*>(row + j));
                else dst[j] = row[j];
            }
        } else {
            for (int j = tid; j < n; j += CSC_THREADS) dst[j] = row[j];
        }
    }
    if (bad && flag) *flag = 1;
}

// Generic path (rows longer than one shared-memory image are produced slab by slab).
__global__ void __launch_bounds__(CSC_THREADS)
csc_densify_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val, int64_t m,
                   int64_t n, int cosine, double min_norm, double* __restrict__ out, int64_t ldo, int vec_out,
                   double* __restrict__ inv_norm_out, int32_t* __restrict__ flag) {
    __shared__ __align__(16) double row[CSC_SLAB];
    __shared__ double red[CSC_THREADS / 32];
    __shared__ double s_inv;
    const int tid = threadIdx.x;
    for (int64_t cell = blockIdx.x; cell < m; cell += gridDim.x) {
        const int64_t p0 = indptr[cell], p1 = indptr[cell + 1];
        double inv = 1.0;
        bool bad = false;
        if (cosine) {
            double ss = 0.0;
            for (int64_t p = p0 + tid; p < p1; p += CSC_THREADS) {
                const double v = val[p];
                bad |= nonfinite(v);
                ss += v * v;
            }
            ss = warp_sum(ss);
            if ((tid & 31) == 0) red[tid >> 5] = ss;
            __syncthreads();
            if (tid == 0) {
                double t = 0.0;
                for (int w = 0; w < CSC_THREADS / 32; w++) t += red[w];  // fixed order: deterministic
                s_inv = 1.0 / fmax(sqrt(t), min_norm);
                if (inv_norm_out) inv_norm_out[cell] = s_inv;
            }
            __syncthreads();
            inv = s_inv;
        }
        double* dst = out + cell * ldo;
        for (int64_t g0 = 0; g0 < n; g0 += CSC_SLAB) {
            const int width = (int)((n - g0) < CSC_SLAB ? (n - g0) : CSC_SLAB);
            for (int j = tid * 2; j < width; j += CSC_THREADS * 2) *reinterpret_cast<double2*>(row + j) = make_double2(0.0, 0.0);
            __syncthreads();
            for (int64_t p = p0 + tid; p < p1; p += CSC_THREADS) {
                const int64_t g = (int64_t)idx[p] - g0;
                if (g >= 0 && g < width) {
                    const double v = val[p];
                    if (!cosine) bad |= nonfinite(v);
                    atomicAdd(row + g, v * inv);  // duplicates are summed; unique indices are stored exactly (0 + x)
                }
            }
            __syncthreads();
            if (vec_out) {
                for (int j = tid * 2; j < width; j += CSC_THREADS * 2) {
                    if (j + 1 < width) stg_stream2(dst + g0 + j, *reinterpret_cast<const double2*>(row + j));
                    else dst[g0 + j] = row[j];
                }
            } else {
                for (int j = tid; j < width; j += CSC_THREADS) dst[g0 + j] = row[j];
            }
            __syncthreads();
        }
        if (bad && flag) *flag = 1;
    }
}

// Validates the index structure on the device (what a corrupt dgCMatrix would otherwise turn into silent garbage or an
// out-of-bounds shared-memory store): offsets non-decreasing and inside [0, nnz], gene indices inside [0, n).  One warp per
// cell; also reports whether every cell's indices are strictly increasing (canonical: sorted, no duplicates).
// flags[0] = 1 bad offsets / 2 bad index, flags[1] = 1 when some cell is not canonical.
__global__ void csc_check_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, int64_t m, int64_t n,
                                 int64_t nnz, int32_t* __restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp; c < m; c += nwarps) {
        const int64_t p0 = indptr[c], p1 = indptr[c + 1];
        if (p0 > p1 || p0 < 0 || p1 > nnz) { if (lane == 0) flags[0] = 1; continue; }
        for (int64_t p = p0 + lane; p < p1; p += 32) {
            const int32_t g = idx[p];
            if (g < 0 || g >= n) flags[0] = 2;
            if (p > p0 && idx[p - 1] >= g) flags[1] = 1;
        }
    }
}

inline int csc_to_dense(scm_ctx* ctx, const int64_t* d_indptr, const int32_t* d_idx, const double* d_val, int64_t m, int64_t n,
                        int64_t nnz, int cosine, double min_norm, double* d_out, int64_t ldo, double* d_inv_norm, int32_t* d_nonfinite) {
    if (m <= 0 || n <= 0) return 0;
    if (!d_indptr || !d_out || ldo < n || (nnz > 0 && (!d_idx || !d_val))) return fail(SCM_ERR_INVALID, "csc_to_dense: bad arguments");
    cudaStream_t st = ctx->stream;
    Scratch ws(ctx);
    int32_t* bad;
    SCM_TRY(ws.alloc(&bad, 2));
    SCM_TRY(zero_async(ctx, bad, 8));
    csc_check_kernel<<<ctx->num_sms * 8, 256, 0, st>>>(d_indptr, d_idx, m, n, nnz, bad);
    SCM_LAUNCH_CHECK(ctx);
    int32_t hbad[2] = {0, 0};
    SCM_CUDA(cudaMemcpyAsync(hbad, bad, sizeof(hbad), cudaMemcpyDeviceToHost, st));
    SCM_CUDA(cudaStreamSynchronize(st));
    if (hbad[0]) return fail(SCM_ERR_INVALID, hbad[0] == 1 ? "csc_to_dense: column pointers are not a non-decreasing sequence inside [0, nnz]"
                                                         : "csc_to_dense: a gene index is outside [0, n)");
    const bool canonical = hbad[1] == 0;
    const int vec_out = (ldo % 2 == 0) && (((uintptr_t)d_out) % 16 == 0);
    if (n <= 6144 && canonical && !getenv("SCM_CSC_GENERIC")) {
        const int slab = (int)((n + 1) & ~(int64_t)1);
        const int smem = (2 * slab + 16) * (int)sizeof(double);
        SCM_CUDA(cudaFuncSetAttribute(csc_densify_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (2 * 6144 + 16) * 8));  // per device
        int per_sm = (int)((220 * 1024) / (smem + 1024));
        if (per_sm > 8) per_sm = 8;
        if (per_sm < 1) per_sm = 1;
        const int64_t want = (int64_t)ctx->num_sms * per_sm;
        const unsigned grid = (unsigned)(m < want ? m : want);
        csc_densify_fast_kernel<<<grid, CSC_THREADS, smem, st>>>(d_indptr, d_idx, d_val, m, (int)n, slab, cosine, min_norm, d_out, ldo,
                                                                 vec_out, d_inv_norm, d_nonfinite);
        SCM_LAUNCH_CHECK(ctx);
        return 0;
    }
    const int64_t want = (int64_t)ctx->num_sms * 7;
    const unsigned grid = (unsigned)(m < want ? m : want);
    csc_densify_kernel<<<grid, CSC_THREADS, 0, st>>>(d_indptr, d_idx, d_val, m, n, cosine, min_norm, d_out, ldo, vec_out, d_inv_norm,
                                                     d_nonfinite);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace scm
