// The cell-sized half of every fit (fastRUVIII / scRUVIII / scRUVg): ONE summing pass and ONE Gram pass over the local cells,
// two all-reduces, and out comes the state every rank needs, bit-identical everywhere:
//
//   gsums / gcounts   replicate-group sums and sizes of ALL cells        (my_residop, R/fastRUVIII.R:121-129)
//   ybar              column means of all cells                          (stand_mean, R/scRUVIII.R:162; adjusted_means)
//   bmeans / ssq      batch means, within-batch squared deviations       (standardize2, R/scRUVIII.R:163-170; only with batches)
//   Gm                centred Gram  sum_i (y_i - ybar)(y_i - ybar)'      (what runSVD's dgesdd works on, after gram_residual)
//
// The Gram kernel without a shift vector has no DADD in its DMMA stream and is 10 % faster (gram_tma.cuh), so it reads a
// PRE-SHIFTED copy Y - a of the local cells.  `a` is the column mean of a strided sample of <= 4096 LOCAL cells: it is known
// before the first full pass, so the copy is written by the summing pass itself (K1 streams every cell anyway: 8 m n more bytes
// written, no second read), and it needs no agreement between ranks.  Exactness does not depend on `a`: with d = a - ybar and
// u = sum_local (y_i - a),
//       sum_local (y_i - ybar)(y_i - ybar)' = sum_local (y_i - a)(y_i - a)' + u d' + d u' + m_local d d'
// (a rank-2 update of n x n, applied before the Gram all-reduce).  `a` only has to be close enough to the mean that nothing
// cancels in the SYRK; |d| is the standard error of a 4096-cell mean.
// Collectives per fit: [gsums | bsums | gcounts | bcounts | flag, rows] as ONE float64 buffer (counts < 2^53 are exact in
// float64), then [Gm | ssq] as one buffer -- two all-reduces for fastRUVIII and for scRUVIII (round 1: six and nine host hooks).
#pragma once
#include "comm.cuh"
#include "common.cuh"
#include "gram_tma.cuh"
#include "seg_colsum.cuh"

namespace scm {

// out[j] = sums[j] / count[0]  (0 when there are no rows)
__global__ void vec_mean_kernel(const double* __restrict__ sums, const int64_t* __restrict__ count, int64_t n, double* __restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) out[j] = count[0] > 0 ? sums[j] / (double)count[0] : 0.0;
}

// out[i, j] = Y[i, j] - shift[j]: the pre-shifted copy as a separate pass (SCM_K1_FUSED_COPY=0; HBM-bound: 8 m n read + 8 m n
// written).  One CTA per 8-row group (grid-stride), 16-byte accesses when both leading dimensions are even and the bases aligned.
template <int VEC>
__global__ void __launch_bounds__(256)
shift_copy_kernel(const double* __restrict__ Y, int64_t m, int64_t n, int64_t ld, const double* __restrict__ shift,
                  double* __restrict__ out, int64_t ldo) {
    for (int64_t r0 = (int64_t)blockIdx.x * 8; r0 < m; r0 += (int64_t)gridDim.x * 8) {
        const int nr = (int)min((int64_t)8, m - r0);
        for (int64_t j = (int64_t)threadIdx.x * VEC; j < n; j += 256 * VEC) {
            if (VEC == 2 && j + 1 < n) {
                const double2 s = *reinterpret_cast<const double2*>(shift + j);
                double2 v[8];
#pragma unroll
                for (int u = 0; u < 8; u++)
                    if (u < nr) v[u] = ldg_stream2(Y + (r0 + u) * ld + j);
#pragma unroll
                for (int u = 0; u < 8; u++)
                    if (u < nr) *reinterpret_cast<double2*>(out + (r0 + u) * ldo + j) = make_double2(v[u].x - s.x, v[u].y - s.y);
            } else {
                for (int u = 0; u < nr; u++) out[(r0 + u) * ldo + j] = Y[(r0 + u) * ld + j] - shift[j];
            }
        }
    }
}

// Before the first all-reduce (local values): counts and the flag / row count join the float64 buffer, and
// u[j] = sum_local (y_ij - a_j) = sum_g gsums[g, j] - m_local a_j is taken while gsums still holds the LOCAL sums.
__global__ void fit_pack_kernel(const double* __restrict__ gsums, int32_t G, int64_t n, const int64_t* __restrict__ gcounts,
                                const int64_t* __restrict__ bcounts, int32_t Bt, const int32_t* __restrict__ flag, int64_t m_local,
                                const double* __restrict__ a, double* __restrict__ cnt_f64, double* __restrict__ u) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) {
        double s = 0.0;
        for (int g = 0; g < G; g++) s += gsums[(int64_t)g * n + j];
        u[j] = s - (double)m_local * a[j];
    }
    if (j < G) cnt_f64[j] = (double)gcounts[j];
    if (j < Bt) cnt_f64[G + j] = (double)bcounts[j];
    if (j == 0) { cnt_f64[G + Bt] = flag[0] ? 1.0 : 0.0; cnt_f64[G + Bt + 1] = (double)m_local; }
}

// After it (global values): counts back to int64, meta = {non-finite flag, cells in total, cells that carry a group key},
// ybar = column means of all cells, d = a - ybar.
__global__ void fit_unpack_kernel(const double* __restrict__ gsums, int32_t G, int64_t n, const double* __restrict__ cnt_f64, int32_t Bt,
                                  int64_t* __restrict__ gcounts, int64_t* __restrict__ bcounts, int64_t* __restrict__ meta,
                                  const double* __restrict__ a, double* __restrict__ ybar, double* __restrict__ d) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const double m_total = cnt_f64[G + Bt + 1];
    if (j < n) {
        double s = 0.0;
        for (int g = 0; g < G; g++) s += gsums[(int64_t)g * n + j];
        const double mean = m_total > 0 ? s / m_total : 0.0;
        ybar[j] = mean;
        d[j] = a[j] - mean;
    }
    if (j < G) gcounts[j] = (int64_t)llrint(cnt_f64[j]);
    if (j < Bt) bcounts[j] = (int64_t)llrint(cnt_f64[G + j]);
    if (j == 0) {
        double keyed = 0.0;
        for (int g = 0; g < G; g++) keyed += cnt_f64[g];
        meta[0] = cnt_f64[G + Bt] != 0.0 ? 1 : 0;
        meta[1] = (int64_t)llrint(m_total);
        meta[2] = (int64_t)llrint(keyed);
    }
}

// gram[i, j] += u_i d_j + d_i u_j + m_local d_i d_j   (shift of the LOCAL Gram from a to the global column mean)
__global__ void gram_recentre_kernel(double* __restrict__ gram, int64_t n, const double* __restrict__ u, const double* __restrict__ d,
                                     double m_local) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = blockIdx.y;
    if (j >= n) return;
    // explicit roundings, no FMA contraction: every term is symmetric in (i, j) bit for bit, so the centred Gram stays EXACTLY
    // symmetric (fma(u_i, d_j, d_i u_j) would round one product and not the other)
    const double di = d[i], dj = d[j];
    const double t = __dadd_rn(__dadd_rn(__dmul_rn(u[i], dj), __dmul_rn(di, u[j])), __dmul_rn(m_local, __dmul_rn(di, dj)));
    gram[i * n + j] = __dadd_rn(gram[i * n + j], t);
}

// sums[g, :] /= counts[g]  (empty groups -> 0)
__global__ void groupmean_kernel(double* __restrict__ sums, const int64_t* __restrict__ counts, int32_t G, int64_t n) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    // groups over gridDim.y (a 1-D launch walks all of them: fine for a handful of batches; the 4800 pseudobulks of scMerge2 took
    // 2.5 ms that way, 2000 threads with a 4800-long dependent loop each)
    for (int g = blockIdx.y; g < G; g += gridDim.y) {
        const int64_t c = counts[g];
        sums[(int64_t)g * n + j] = c > 0 ? sums[(int64_t)g * n + j] / (double)c : 0.0;
    }
}

struct FitState {
    double* gsums = nullptr;     // G x n (global)
    int64_t* gcounts = nullptr;  // G
    double* bmeans = nullptr;    // Bt x n (global batch means; only with batches)
    int64_t* bcounts = nullptr;  // Bt
    double* ssq = nullptr;       // Bt x n, sum over the batch's cells of (y - batch mean)^2 (global)
    double* ybar = nullptr;      // n
    int64_t* meta = nullptr;     // {non-finite flag, cells in total, cells with a group key}
    double* Gm = nullptr;        // n x n centred Gram of all cells
};

// Sample mean of <= 4096 strided local cells (zeros when the rank holds no cells).
inline int sample_mean(scm_ctx* ctx, Scratch& ws, const double* d_Y, int64_t m_local, int64_t n, int64_t ld, double* a) {
    double* psum;
    int64_t* pcount;
    int32_t* zkey;
    const int64_t ms = m_local < 4096 ? m_local : 4096, stride = ms > 0 ? m_local / ms : 1;
    SCM_TRY(ws.alloc(&psum, n)); SCM_TRY(ws.alloc(&pcount, 1)); SCM_TRY(ws.alloc(&zkey, ms > 0 ? ms : 1));
    SCM_TRY(zero_async(ctx, zkey, sizeof(int32_t) * (ms > 0 ? ms : 1)));
    SCM_TRY(seg_colsum(ctx, d_Y, ms, n, ld * stride, zkey, 1, nullptr, psum, pcount, nullptr));
    vec_mean_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, ctx->stream>>>(psum, pcount, n, a);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

// See the header comment.  All buffers of `fs` are allocated from `ws` (they live as long as the caller's Scratch).
inline int fit_state(scm_ctx* ctx, Scratch& ws, const double* d_Y, int64_t m_local, int64_t n, int64_t ld, const int32_t* d_group,
                     int32_t G, const int32_t* d_batch, int32_t Bt, FitState& fs) {
    cudaStream_t st = ctx->stream;
    if (!d_batch) Bt = 0;
    // buffer 1: [gsums G n | bsums Bt n | gcounts G | bcounts Bt | flag | rows]     buffer 2: [Gm n n | ssq Bt n]
    const int64_t len1 = (int64_t)(G + Bt) * n + G + Bt + 2, len2 = n * n + (int64_t)Bt * n;
    double *buf1, *buf2, *a, *u, *d;
    int64_t *gcnt_local, *bcnt_local;
    int32_t* flag;
    SCM_TRY(ws.alloc(&buf1, len1)); SCM_TRY(ws.alloc(&buf2, len2));
    SCM_TRY(ws.alloc(&a, n)); SCM_TRY(ws.alloc(&u, n)); SCM_TRY(ws.alloc(&d, n)); SCM_TRY(ws.alloc(&fs.ybar, n));
    SCM_TRY(ws.alloc(&fs.gcounts, G)); SCM_TRY(ws.alloc(&gcnt_local, G)); SCM_TRY(ws.alloc(&fs.meta, 4)); SCM_TRY(ws.alloc(&flag, 1));
    SCM_TRY(ws.alloc(&fs.bcounts, Bt > 0 ? Bt : 1)); SCM_TRY(ws.alloc(&bcnt_local, Bt > 0 ? Bt : 1));
    fs.gsums = buf1;
    double* bsums = buf1 + (int64_t)G * n;
    double* cnt_f64 = buf1 + (int64_t)(G + Bt) * n;
    fs.Gm = buf2;
    fs.ssq = Bt > 0 ? buf2 + n * n : nullptr;
    fs.bmeans = Bt > 0 ? bsums : nullptr;
    SCM_TRY(zero_async(ctx, flag, sizeof(int32_t)));
    ctx->ev_used[ST_PRESHIFT] = false;
    ctx->ev_used[ST_STD] = false;

    // ---- the shift and room for the shifted copy
    struct AsyncBuf {  // stream-ordered buffer that may fail to allocate (freed on every exit path)
        double* p = nullptr;
        cudaStream_t st;
        ~AsyncBuf() { if (p) cudaFreeAsync(p, st); }
    } ypre;
    ypre.st = st;
    const int64_t ldp = (n + 1) & ~(int64_t)1;
    SCM_TRY(sample_mean(ctx, ws, d_Y, m_local, n, ld, a));
    if (ctx->gram_preshift && ctx->use_tma && m_local > 0) {
        if (cudaMallocAsync((void**)&ypre.p, sizeof(double) * (size_t)m_local * (size_t)ldp, st) != cudaSuccess) {
            cudaGetLastError();  // no room for the copy: the Gram kernel subtracts `a` from its fragments instead
            ypre.p = nullptr;
        }
    }
    const bool fused_copy = ypre.p != nullptr && ctx->k1_fused_copy;
    {   // residual operator, part 1: replicate-group sums (my_residop, R/fastRUVIII.R:121-129) + NA/Inf scan (:52-60) + Y - a
        StageTimer t(ctx, ST_SEG);
        if (fused_copy) SCM_TRY(seg_colsum(ctx, d_Y, m_local, n, ld, d_group, G, nullptr, fs.gsums, gcnt_local, flag, a, ypre.p, ldp));
        else SCM_TRY(seg_colsum(ctx, d_Y, m_local, n, ld, d_group, G, nullptr, fs.gsums, gcnt_local, flag));
    }
    if (Bt > 0) {  // standardize2, pass 1 (R/scRUVIII.R:163-166): batch sums
        StageTimer t(ctx, ST_STD);
        SCM_TRY(seg_colsum(ctx, d_Y, m_local, n, ld, d_batch, Bt, nullptr, bsums, bcnt_local, nullptr));
    }
    {
        const int64_t span = n > G + Bt ? n : G + Bt;
        fit_pack_kernel<<<(unsigned)ceil_div(span, 256), 256, 0, st>>>(fs.gsums, G, n, gcnt_local, bcnt_local, Bt, flag, m_local, a, cnt_f64, u);
        SCM_LAUNCH_CHECK(ctx);
        SCM_TRY(allreduce(ctx, buf1, len1, 0));
        fit_unpack_kernel<<<(unsigned)ceil_div(span, 256), 256, 0, st>>>(fs.gsums, G, n, cnt_f64, Bt, fs.gcounts, fs.bcounts, fs.meta, a,
                                                                        fs.ybar, d);
        SCM_LAUNCH_CHECK(ctx);
    }
    if (Bt > 0) {  // standardize2, pass 2 (R/scRUVIII.R:167-170): squared deviations from the (global) batch means
        StageTimer t(ctx, ST_STD);
        groupmean_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(bsums, fs.bcounts, Bt, n);
        SCM_LAUNCH_CHECK(ctx);
        SCM_TRY(seg_colsum(ctx, d_Y, m_local, n, ld, d_batch, Bt, bsums, fs.ssq, nullptr, nullptr));
    }
    {   // Gram of the local cells about `a`, moved to the global mean
        StageTimer t(ctx, ST_GRAM);
        if (ypre.p && !fused_copy) {
            const bool vec2 = (ld % 2 == 0) && (((uintptr_t)d_Y) % 16 == 0);
            const int64_t groups = ceil_div(m_local, 8);
            const unsigned grid = (unsigned)(groups < (int64_t)ctx->num_sms * 16 ? groups : (int64_t)ctx->num_sms * 16);
            StageTimer tc(ctx, ST_PRESHIFT);
            if (vec2) shift_copy_kernel<2><<<grid, 256, 0, st>>>(d_Y, m_local, n, ld, a, ypre.p, ldp);
            else shift_copy_kernel<1><<<grid, 256, 0, st>>>(d_Y, m_local, n, ld, a, ypre.p, ldp);
            SCM_LAUNCH_CHECK(ctx);
        }
        if (ypre.p) SCM_TRY(gram(ctx, ypre.p, m_local, n, ldp, nullptr, fs.Gm, true));
        else SCM_TRY(gram(ctx, d_Y, m_local, n, ld, a, fs.Gm, true));
        gram_recentre_kernel<<<dim3((unsigned)ceil_div(n, 256), (unsigned)n), 256, 0, st>>>(fs.Gm, n, u, d, (double)m_local);
        SCM_LAUNCH_CHECK(ctx);
    }
    SCM_TRY(allreduce(ctx, buf2, len2, 0));
    return 0;
}

// meta -> host (pinned), to be read after the caller's next stream synchronisation
inline int fit_meta_to_host(scm_ctx* ctx, const FitState& fs, int64_t** h_meta) {
    int64_t* h = reinterpret_cast<int64_t*>(ctx->h_pinned + 2048);
    SCM_CUDA(cudaMemcpyAsync(h, fs.meta, sizeof(int64_t) * 3, cudaMemcpyDeviceToHost, ctx->stream));
    *h_meta = h;
    return 0;
}
inline int fit_meta_check(const int64_t* h_meta) {
    if (h_meta[0] != 0) return fail(SCM_ERR_NONFINITE, "Y contains missing values or infinities.  This is not supported.");
    if (h_meta[2] != h_meta[1])
        return fail(SCM_ERR_INVALID, "%lld of %lld cells have a replicate key outside [0, G): M must assign every cell to a group",
                    (long long)(h_meta[1] - h_meta[2]), (long long)h_meta[1]);
    return 0;
}

}  // namespace scm
