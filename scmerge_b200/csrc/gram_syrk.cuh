// K3: Gram matrix of the (column-shifted) cells x genes matrix, FP64 tensor cores.
//
//   gram = sum_i (y_i - a)(y_i - a)'      (n x n, symmetric)
//
// * only lower-triangular 128x128 tiles are computed (SYRK) and mirrored by the reduce kernel,
// * the cell dimension is cut into S splits; work unit u = split * T + tile is handled by CTA u % grid, so
//   all resident CTAs sweep the same range of cells at the same time and every cell tile is fetched from
//   HBM once and then served from L2 to the other tiles that need it,
// * each unit: 4-stage cp.async pipeline of [16 cells x 128 genes] operand tiles in shared memory
//   (row stride 132 doubles => conflict-free DMMA fragment loads), 8 warps x (64x32) warp tiles,
//   32 DMMA.8x8x4 per 12 shared-memory loads,
// * the optional column shift is applied in shared memory by the thread that copied the chunk,
// * unit results go to a partial buffer [S][T][128x128]; a second kernel adds the splits in order
//   (deterministic) and writes the full symmetric matrix.
// Credited work: m*n*(n+1) flop (SYRK), 8*m*n bytes.
#pragma once
#include "common.cuh"

namespace scm {

constexpr int GR_BT = 128;                // tile edge (genes)
constexpr int GR_KC = 16;                 // cells per pipeline stage
constexpr int GR_LDS = GR_BT + 4;         // smem row stride in doubles (== 4 mod 16: conflict-free fragments)
constexpr int GR_STAGES = 4;
constexpr int GR_THREADS = 256;
constexpr int GR_STAGE_DOUBLES = 2 * GR_KC * GR_LDS;  // A tile + B tile
constexpr int GR_SMEM_BYTES = GR_STAGES * GR_STAGE_DOUBLES * 8;

struct GramPlan {
    int ntile;     // tiles per dimension
    int T;         // lower-triangular tiles
    int S;         // splits over cells
    int64_t rows_per_split;  // multiple of GR_KC
    int grid;
};

inline GramPlan gram_plan(int64_t m, int64_t n, int num_sms) {
    GramPlan p;
    p.ntile = (int)ceil_div(n, GR_BT);
    p.T = p.ntile * (p.ntile + 1) / 2;
    const int64_t kchunks = ceil_div(m, GR_KC);
    // choose S (<= 64, each split >= 64 chunks) minimising the idle fraction of the last wave
    int bestS = 1;
    double bestWaste = 1e9;
    for (int S = 1; S <= 64; S++) {
        if (S > 1 && kchunks / S < 64) break;
        const int64_t units = (int64_t)p.T * S;
        const int64_t waves = ceil_div(units, num_sms);
        const double waste = (double)(waves * num_sms) / (double)units;
        if (waste < bestWaste - 1e-9) { bestWaste = waste; bestS = S; }
    }
    p.S = bestS;
    p.rows_per_split = ceil_div(kchunks, p.S) * GR_KC;
    const int64_t units = (int64_t)p.T * p.S;
    p.grid = (int)(units < num_sms ? units : num_sms);
    return p;
}

// tile index t (0..T-1) -> (ti, tj) with tj <= ti, row-major over the lower triangle
__device__ __forceinline__ void tri_decode(int t, int& ti, int& tj) {
    int i = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= t) i++;
    while (i * (i + 1) / 2 > t) i--;
    ti = i;
    tj = t - i * (i + 1) / 2;
}

template <int VEC>
__global__ void __launch_bounds__(GR_THREADS, 1)
gram_syrk_kernel(const double* __restrict__ Y, int64_t m, int64_t n, int64_t ld, const double* __restrict__ shift,
                 int T, int S, int64_t rows_per_split, double* __restrict__ partial) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps: warp tile = 64 (i) x 32 (j)

    // copy assignment: 16-byte chunk (2 genes) `cc` of rows cr, cr+4, cr+8, cr+12 of each operand tile
    const int cc = tid & 63, cr = tid >> 6;

    const int64_t units = (int64_t)T * S;
    for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
        const int split = (int)(u / T);
        int ti, tj;
        tri_decode((int)(u % T), ti, tj);
        const bool diag = (ti == tj);
        const int64_t row_begin = (int64_t)split * rows_per_split;
        const int64_t row_end = min(m, row_begin + rows_per_split);
        const int nchunk = row_end > row_begin ? (int)ceil_div(row_end - row_begin, GR_KC) : 0;
        const int64_t colA = (int64_t)ti * GR_BT + 2 * cc, colB = (int64_t)tj * GR_BT + 2 * cc;
        const bool vA0 = colA < n, vA1 = colA + 1 < n, vB0 = colB < n, vB1 = colB + 1 < n;
        double sA0 = 0, sA1 = 0, sB0 = 0, sB1 = 0;
        if (shift) {
            if (vA0) sA0 = shift[colA];
            if (vA1) sA1 = shift[colA + 1];
            if (vB0) sB0 = shift[colB];
            if (vB1) sB1 = shift[colB + 1];
        }

        auto issue = [&](int chunk) {
            double* sa = smem + (chunk % GR_STAGES) * GR_STAGE_DOUBLES;
            double* sb = sa + GR_KC * GR_LDS;
            const int64_t r0 = row_begin + (int64_t)chunk * GR_KC;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int r = cr + 4 * i;
                const int64_t row = r0 + r;
                const bool rv = row < row_end;
                const double* ga = Y + (rv ? row : 0) * ld + (vA0 ? colA : 0);
                const double* gb = Y + (rv ? row : 0) * ld + (vB0 ? colB : 0);
                double* da = sa + r * GR_LDS + 2 * cc;
                double* db = sb + r * GR_LDS + 2 * cc;
                if (VEC == 2) {
                    // a 16-byte chunk never straddles n when n is even; for odd n the last chunk is split
                    if (vA0 == vA1) cp_async16(da, ga, rv && vA0);
                    else { cp_async8(da, ga, rv); cp_async8(da + 1, ga, false); }
                    if (!diag) {
                        if (vB0 == vB1) cp_async16(db, gb, rv && vB0);
                        else { cp_async8(db, gb, rv); cp_async8(db + 1, gb, false); }
                    }
                } else {
                    cp_async8(da, ga, rv && vA0);
                    cp_async8(da + 1, ga + (vA1 ? 1 : 0), rv && vA1);
                    if (!diag) {
                        cp_async8(db, gb, rv && vB0);
                        cp_async8(db + 1, gb + (vB1 ? 1 : 0), rv && vB1);
                    }
                }
            }
        };

        double acc[8][4][2];
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;

        for (int c = 0; c < GR_STAGES - 1; c++) {
            if (c < nchunk) issue(c);
            cp_async_commit();
        }
        for (int c = 0; c < nchunk; c++) {
            cp_async_wait<GR_STAGES - 2>();
            double* sa = smem + (c % GR_STAGES) * GR_STAGE_DOUBLES;
            double* sb = sa + GR_KC * GR_LDS;
            if (shift) {  // subtract the column shift from the chunks this thread copied (valid entries only)
                const int64_t r0 = row_begin + (int64_t)c * GR_KC;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int r = cr + 4 * i;
                    if (r0 + r < row_end) {
                        double2* pa = reinterpret_cast<double2*>(sa + r * GR_LDS + 2 * cc);
                        double2 v = *pa;
                        if (vA0) v.x -= sA0;
                        if (vA1) v.y -= sA1;
                        *pa = v;
                        if (!diag) {
                            double2* pb = reinterpret_cast<double2*>(sb + r * GR_LDS + 2 * cc);
                            double2 w = *pb;
                            if (vB0) w.x -= sB0;
                            if (vB1) w.y -= sB1;
                            *pb = w;
                        }
                    }
                }
            }
            __syncthreads();
            if (c + GR_STAGES - 1 < nchunk) issue(c + GR_STAGES - 1);
            cp_async_commit();
            const double* ta = sa + wm * 64 + gid;
            const double* tb = (diag ? sa : sb) + wn * 32 + gid;
#pragma unroll
            for (int ks = 0; ks < GR_KC / 4; ks++) {
                double af[8], bf[4];
                const int koff = (ks * 4 + tig) * GR_LDS;
#pragma unroll
                for (int a = 0; a < 8; a++) af[a] = ta[koff + a * 8];
#pragma unroll
                for (int b = 0; b < 4; b++) bf[b] = tb[koff + b * 8];
#pragma unroll
                for (int a = 0; a < 8; a++)
#pragma unroll
                    for (int b = 0; b < 4; b++) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
            }
        }
        cp_async_wait<0>();
        __syncthreads();  // all warps done with smem before the next unit's prologue overwrites it

        // unit result -> partial[split][tile][128][128]
        double* out = partial + ((int64_t)split * T + (u % T)) * (GR_BT * GR_BT);
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int i = wm * 64 + a * 8 + gid, j = wn * 32 + b * 8 + 2 * tig;
                *reinterpret_cast<double2*>(out + i * GR_BT + j) = make_double2(acc[a][b][0], acc[a][b][1]);
            }
    }
}

// gram[i][j] = sum_s partial[s][tile(i,j)][..] for j-tile <= i-tile, mirrored for the upper triangle.
// One CTA per lower tile; writes both (ti,tj) and its transpose.  tri != 0: the partials of DIAGONAL tiles hold only the
// 8x8 blocks on or below the diagonal (lean TMA kernel); the blocks above are filled by mirroring.
__global__ void __launch_bounds__(256)
gram_reduce_kernel(const double* __restrict__ partial, int T, int S, int64_t n, double* __restrict__ gram, int accumulate, int tri) {
    __shared__ double tile[32][33];
    int ti, tj;
    tri_decode(blockIdx.x, ti, tj);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int bi = (blockIdx.y >> 2) * 32, bj = (blockIdx.y & 3) * 32;  // 32x32 sub-block of the tile
    const bool dtile = tri && ti == tj;
    if (dtile && bi < bj) return;  // entirely above the diagonal: written as the mirror image of sub-block (bj, bi)
    for (int r = ty; r < 32; r += 8) {
        const int i = bi + r, j = bj + tx;
        const bool valid = !dtile || (i >> 3) >= (j >> 3);
        const int64_t gi = (int64_t)ti * GR_BT + i, gj = (int64_t)tj * GR_BT + j;
        double acc = 0.0;
        if (valid) {
            if (accumulate && gi < n && gj < n) acc = gram[gi * n + gj];  // chunked fits add into the running Gram
            for (int s = 0; s < S; s++) acc += partial[((int64_t)s * T + blockIdx.x) * (GR_BT * GR_BT) + i * GR_BT + j];
            if (gi < n && gj < n) gram[gi * n + gj] = acc;
        }
        tile[r][tx] = acc;
    }
    __syncthreads();
    if (ti != tj || dtile) {
        for (int r = ty; r < 32; r += 8) {
            const int j = bj + r, i = bi + tx;  // transposed write: row = gene j, col = gene i
            if (dtile && (i >> 3) <= (j >> 3)) continue;  // only the strictly-lower blocks are mirrored
            const int64_t gi = (int64_t)ti * GR_BT + i, gj = (int64_t)tj * GR_BT + j;
            if (gi < n && gj < n) gram[gj * n + gi] = tile[tx][r];
        }
    }
}

// gram <- scale_i scale_j ( gram - sum_g cnt_g (mu_g - a)_i (mu_g - a)_j ) given corr = Gram of the G x n
// matrix C[g] = sqrt(cnt_g) (mu_g - a) (computed with the same SYRK kernel).
__global__ void gram_residual_kernel(double* __restrict__ gram, const double* __restrict__ corr, int64_t n,
                                     const double* __restrict__ scale) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = blockIdx.y;
    if (j >= n) return;
    double v = gram[i * n + j] - corr[i * n + j];
    if (scale) v *= scale[i] * scale[j];
    gram[i * n + j] = v;
}

// C[g, j] = sqrt(cnt_g) * (sums[g, j] / cnt_g - shift_j)   (0 for empty groups); C has row stride ldc
__global__ void group_dev_kernel(const double* __restrict__ sums, const int64_t* __restrict__ counts, int64_t n,
                                 int32_t G, const double* __restrict__ shift, double* __restrict__ C, int64_t ldc) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ldc) return;
    for (int64_t g = blockIdx.y; g < G; g += gridDim.y) {
        const int64_t c = counts[g];
        double v = 0.0;
        if (c > 0 && j < n) v = sqrt((double)c) * (sums[g * n + j] / (double)c - (shift ? shift[j] : 0.0));
        C[g * ldc + j] = v;
    }
}

// TMA + mbarrier variant (gram_tma.cuh): returns 1 when not applicable
inline int gram_tma_launch(scm_ctx* ctx, Scratch& ws, const double* d_Y, int64_t m, int64_t n, int64_t ld, const double* d_shift,
                           const GramPlan& p, double* partial);

inline int gram(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, const double* d_shift, double* d_gram,
                bool timed = false, bool accumulate = false) {
    if (m < 0 || n <= 0 || ld < n) return fail(SCM_ERR_INVALID, "gram: bad shape m=%lld n=%lld ld=%lld", (long long)m, (long long)n, (long long)ld);
    cudaStream_t st = ctx->stream;
    if (m == 0) {
        if (!accumulate) SCM_TRY(zero_async(ctx, d_gram, sizeof(double) * n * n));
        return 0;
    }
    // function attributes are per device: set on every call (microseconds), never cached in a process-wide flag
    SCM_CUDA(cudaFuncSetAttribute(gram_syrk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GR_SMEM_BYTES));
    SCM_CUDA(cudaFuncSetAttribute(gram_syrk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GR_SMEM_BYTES));
    const GramPlan p = gram_plan(m, n, ctx->num_sms);
    Scratch ws(ctx);
    double* partial;
    SCM_TRY(ws.alloc(&partial, (int64_t)p.S * p.T * GR_BT * GR_BT));
    const bool vec2 = (ld % 2 == 0) && (((uintptr_t)d_Y) % 16 == 0);
    if (timed) cudaEventRecord(ctx->ev0[ST_SYRK], st);
    const int tma_rc = ctx->use_tma ? gram_tma_launch(ctx, ws, d_Y, m, n, ld, d_shift, p, partial) : 1;
    if (tma_rc < 0) return tma_rc;
    if (tma_rc == 0) { /* launched on the TMA path */ }
    else if (vec2)
        gram_syrk_kernel<2><<<p.grid, GR_THREADS, GR_SMEM_BYTES, st>>>(d_Y, m, n, ld, d_shift, p.T, p.S, p.rows_per_split, partial);
    else
        gram_syrk_kernel<1><<<p.grid, GR_THREADS, GR_SMEM_BYTES, st>>>(d_Y, m, n, ld, d_shift, p.T, p.S, p.rows_per_split, partial);
    if (timed) { cudaEventRecord(ctx->ev1[ST_SYRK], st); ctx->ev_used[ST_SYRK] = true; }
    SCM_LAUNCH_CHECK(ctx);
    gram_reduce_kernel<<<dim3(p.T, 16), 256, 0, st>>>(partial, p.T, p.S, n, d_gram, accumulate ? 1 : 0,
                                                      (tma_rc == 0 && (ctx->gram_lean & 2)) ? 1 : 0);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

inline int gram_residual(scm_ctx* ctx, double* d_gram, int64_t n, const double* d_sums, const int64_t* d_counts,
                         int32_t G, const double* d_shift, const double* d_scale) {
    cudaStream_t st = ctx->stream;
    Scratch ws(ctx);
    double *C, *corr;
    const int64_t ldc = (n + 1) & ~(int64_t)1;  // even leading dimension => 16-byte copies
    SCM_TRY(ws.alloc(&C, (int64_t)G * ldc));
    SCM_TRY(ws.alloc(&corr, n * n));
    {
        dim3 grid((unsigned)ceil_div(ldc, 256), (unsigned)(G < 65535 ? G : 65535));
        group_dev_kernel<<<grid, 256, 0, st>>>(d_sums, d_counts, n, G, d_shift, C, ldc);
        SCM_LAUNCH_CHECK(ctx);
    }
    SCM_TRY(gram(ctx, C, G, n, ldc, nullptr, corr));
    dim3 grid((unsigned)ceil_div(n, 256), (unsigned)n);
    gram_residual_kernel<<<grid, 256, 0, st>>>(d_gram, corr, n, d_scale);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace scm
