// Sparse-native scMerge2 path: the cosine-normalised dense matrix is never materialised as an INPUT.
//
// For the dgCMatrix scMerge2 holds (cells compressed, ~10 % non-zero) the three passes over the cells
//     (a) cosine norms            batchelor::cosineNorm                       R/scMerge2.R:193-198
//     (b) pseudobulk means        create_pseudoBulk / aggregate.Matrix        R/scMerge2_utilis.R:325-346,461-486
//     (c) newY = Y - ((Y - c) B) A  pseudoRUVIII's chunked adjust             R/scMerge2_utilis.R:97-137
// read 12 bytes per NON-ZERO instead of 8 bytes per ENTRY; only (c) touches dense memory, and only to WRITE newY once.
// HBM traffic per cell x gene entry drops from 8 (densify, written) + 8 (K1) + 16 + 8 (K6 with all genes as controls:
// two reads, one write) = 40 bytes to 8 bytes + 3 x 12 x density.
//
//   csc_rownorm_kernel      one warp per cell: inv[i] = 1 / max(||x_i||_2, min_norm), NA/Inf scan
//   csc_seg_partial_kernel  K1 on sparse rows: cells ordered by key (same plumbing as the dense K1), one CTA per chunk of
//                           <= 128 cells, every warp sums ITS cells (fixed order) into a private shared-memory accumulator
//                           row (no atomics: indices inside a cell are unique), the warp rows are added in warp order and
//                           the chunk partial goes through the same ordered final reduction: deterministic
//   csc_project_kernel      pass A of K6 on sparse rows, W_i = s_i sum_nz x_ij B[j, :] - c'B, as its own kernel: the rows of B for
//                           a range of genes stay resident in shared memory for the whole launch (682 genes at k <= 24 in 128 KB,
//                           one launch per range), a quad of lanes per cell, eight non-zeros per trip.  History of this gather: lanes
//                           over the k factors with shuffles 26 ms; one lane per non-zero 14 ms; a quad per non-zero through L1TEX
//                           7.5 ms (6 sector look-ups per non-zero, L1TEX 73 % of peak); B staged per 32-cell tile in 256-gene
//                           slabs 5.9 ms (eight staging round trips per tile); this form 2.85 ms, within 1.7x of the shared-memory
//                           bandwidth bound (192 bytes of B per non-zero)
//   csc_adjust_kernel       pass B, 16 cells per CTA, 4 CTAs per SM: builds 16 x 256 tiles of the scaled sparse rows in shared
//                           memory (scatter through per-cell cursors, the first 32 candidates of all cells of a warp loaded
//                           before any is used; the tile is cleared behind the fragment reads, no zero-fill phase) and adds
//                           (-W) A with DMMA.8x8x4 exactly like the dense kernel, streaming newY out; the next block's A
//                           fragments are in flight, the next tile's non-zeros prefetched to L2 (5.9 ms; ablation in DESIGN.md)
// All three require canonical cells (indices strictly increasing, what csc_check_kernel verifies); otherwise the caller
// falls back to densify + the dense kernels.
#pragma once
#include "adjust.cuh"
#include "common.cuh"
#include "csc_ingest.cuh"
#include "seg_colsum.cuh"

namespace scm {

__global__ void __launch_bounds__(256)
csc_rownorm_kernel(const int64_t* __restrict__ indptr, const double* __restrict__ val, int64_t m, int cosine, double min_norm,
                   double* __restrict__ inv, int32_t* __restrict__ flag) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    bool bad = false;
    for (int64_t c = warp; c < m; c += nwarps) {
        const int64_t p0 = indptr[c], p1 = indptr[c + 1];
        double ss = 0.0;
        for (int64_t p = p0 + lane; p < p1; p += 32) {
            const double v = val[p];
            bad |= nonfinite(v);
            ss += v * v;
        }
        ss = warp_sum(ss);  // xor butterfly: every lane holds the same, order-fixed sum
        if (lane == 0) inv[c] = cosine ? 1.0 / fmax(sqrt(ss), min_norm) : 1.0;
    }
    if (bad && flag) *flag = 1;
}

constexpr int CSG_THREADS = 128;  // 4 warps, each with a private accumulator row of n doubles in shared memory

// One CTA per chunk (grid-stride).  acc: [nwarp][n] doubles (dynamic shared memory).
__global__ void __launch_bounds__(CSG_THREADS)
csc_seg_partial_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                       const double* __restrict__ inv, int n, const int32_t* __restrict__ perm, const int64_t* __restrict__ seg_start,
                       const int* __restrict__ chunk_off, const int* __restrict__ pslot_off, int32_t G, int nwarp,
                       double* __restrict__ partial, double* __restrict__ sums) {
    extern __shared__ __align__(16) double csg_acc[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nchunks = chunk_off[G];
    for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        int lo = 0, hi = G;  // largest g with chunk_off[g] <= chunk
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (chunk_off[mid] <= chunk) lo = mid; else hi = mid;
        }
        const int g = lo, j = chunk - chunk_off[g];
        const int64_t start = seg_start[g] + (int64_t)j * SEG_R;
        const int len = (int)min((int64_t)SEG_R, seg_start[g + 1] - start);
        for (int e = tid * 2; e < nwarp * n; e += CSG_THREADS * 2) {  // n is even-padded by the host
            *reinterpret_cast<double2*>(csg_acc + e) = make_double2(0.0, 0.0);
        }
        __syncthreads();
        if (warp < nwarp) {
            double* acc = csg_acc + (int64_t)warp * n;
            // warp w takes cells w, w + nwarp, ... of the chunk, one after the other: a fixed summation order per gene.  A cell costs
            // a chain of dependent loads (perm -> indptr -> idx / val): the chain of the NEXT cell is started before the current
            // cell's non-zeros are used, and a lane keeps eight index / value loads in flight (a cell of 200 non-zeros is one trip)
            int r = warp;
            int64_t cell = r < len ? perm[start + r] : 0;
            int64_t p0 = r < len ? indptr[cell] : 0, p1 = r < len ? indptr[cell + 1] : 0;
            double sc = r < len ? inv[cell] : 0.0;
            for (; r < len; r += nwarp) {
                const int rn = r + nwarp;
                int64_t celln = 0, p0n = 0, p1n = 0;
                double scn = 0.0;
                if (rn < len) { celln = perm[start + rn]; p0n = indptr[celln]; p1n = indptr[celln + 1]; scn = inv[celln]; }
                for (int64_t pb = p0 + lane; pb < p1; pb += 256) {
                    int32_t gi[8];
                    double gv[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int64_t p = pb + 32 * u;
                        gi[u] = p < p1 ? idx[p] : -1;
                        gv[u] = p < p1 ? val[p] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 8; u++)
                        if ((unsigned)gi[u] < (unsigned)n) acc[gi[u]] += gv[u] * sc;  // unique indices inside a cell: no conflicts (a corrupt
                                                                                       // index can never leave the accumulator row)
                }
                __syncwarp();
                cell = celln; p0 = p0n; p1 = p1n; sc = scn;
            }
        }
        __syncthreads();
        const int nch = chunk_off[g + 1] - chunk_off[g];
        double* dst = (nch == 1) ? (sums + (int64_t)g * n) : (partial + (int64_t)(pslot_off[g] + j) * n);
        for (int e = tid; e < n; e += CSG_THREADS) {
            double t = 0.0;
            for (int w = 0; w < nwarp; w++) t += csg_acc[(int64_t)w * n + e];
            dst[e] = t;
        }
        __syncthreads();
    }
}

// Segmented column sums of the scaled sparse rows: sums[g, :] = sum over cells i with key g of inv[i] * x_i (G x ldn, ldn = n
// rounded up to even), counts[g].  Deterministic.  Rows of empty groups are 0.
inline int csc_seg_colsum(scm_ctx* ctx, const int64_t* d_indptr, const int32_t* d_idx, const double* d_val, const double* d_inv,
                          int64_t m, int64_t n, const int32_t* d_key, int32_t G, double* d_sums, int64_t ldn, int64_t* d_counts) {
    if (m < 0 || n <= 0 || G <= 0 || ldn < n || (ldn & 1)) return fail(SCM_ERR_INVALID, "csc_seg_colsum: bad shape");
    if (m >= (int64_t)1 << 31) return fail(SCM_ERR_UNSUPPORTED, "csc_seg_colsum: m >= 2^31");
    if (ldn > 24576) return fail(SCM_ERR_UNSUPPORTED, "csc_seg_colsum: more than 24576 genes");
    cudaStream_t st = ctx->stream;
    Scratch ws(ctx);
    SCM_TRY(zero_async(ctx, d_sums, sizeof(double) * G * ldn));
    if (m == 0) {
        if (d_counts) SCM_TRY(zero_async(ctx, d_counts, sizeof(int64_t) * G));
        return 0;
    }
    SegPlan sp;
    SCM_TRY(seg_plan_build(ctx, ws, d_key, m, G, d_counts, sp));
    const int64_t max_slots = 2 * ceil_div(m, SEG_R) + 1;
    double* partial;
    SCM_TRY(ws.alloc(&partial, max_slots * ldn));
    int nwarp = (int)((200 * 1024) / (ldn * 8));
    if (nwarp > CSG_THREADS / 32) nwarp = CSG_THREADS / 32;
    if (nwarp < 1) nwarp = 1;
    const int smem = (int)(nwarp * ldn * 8);
    static int attr_smem = 0;
    if (smem > attr_smem) {
        SCM_CUDA(cudaFuncSetAttribute(csc_seg_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_smem = smem;
    }
    const int64_t max_chunks = ceil_div(m, SEG_R) + G;
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    int64_t grid = (int64_t)ctx->num_sms * per_sm;
    if (grid > max_chunks) grid = max_chunks;
    csc_seg_partial_kernel<<<(unsigned)grid, CSG_THREADS, smem, st>>>(d_indptr, d_idx, d_val, d_inv, (int)ldn, sp.perm, sp.seg_start,
                                                                     sp.chunk_off, sp.pslot_off, G, nwarp, partial, d_sums);
    SCM_LAUNCH_CHECK(ctx);
    dim3 fgrid((unsigned)G, (unsigned)ceil_div(ldn, 256));
    seg_final_kernel<<<fgrid, 256, 0, st>>>(partial, sp.chunk_off, sp.pslot_off, ldn, d_sums);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

// ---- pass A of K6 on sparse rows, as its own kernel: W[cell, :] = s_cell * sum_nz x B[idx, :] -----------------------------
// The gather of one 8 KT-double row of B per non-zero is what bounded the fused kernel (round 1: through L1TEX, 6 sector look-ups
// per non-zero, 7.5 of 12.5 ms; staged per 32-cell tile in 256-gene slabs it was still 5.9 ms: eight staging round trips per tile
// and a global-load latency per pair of non-zeros).  Here the rows of B for a RANGE OF GENES (as many as fit 200 KB of shared
// memory: 1066 genes at k <= 24) stay resident in shared memory for the whole launch; one launch per gene range.  A quad of
// lanes owns a cell: lane `sub` holds the 16-byte pieces sub, sub + 4, ... of the row (2 KT accumulators per lane), walks the
// cell's non-zeros of the range eight at a time (all index / value loads of a trip in flight together), and 32 warps x 8 quads
// per SM keep 256 cells in flight.  Indices are sorted inside a cell, so a range is a run: the position where it ended is kept
// in `cursor` for the next range's launch.  W accumulates over the launches in global memory (m x 8 KT doubles).
constexpr int CSP_THREADS = 1024;
template <int KT>
__global__ void __launch_bounds__(CSP_THREADS, 1)
csc_project_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                   const double* __restrict__ inv, int64_t m, const double* __restrict__ Bpad, int32_t g_lo, int32_t g_hi, int first,
                   int last, const double* __restrict__ w0, int k, int64_t* __restrict__ cursor, double* __restrict__ W) {
    extern __shared__ __align__(16) double csp_B[];  // [(g_hi - g_lo)][KP]
    constexpr int KP = KT * 8;
    {
        const double2* src = reinterpret_cast<const double2*>(Bpad + (int64_t)g_lo * KP);
        double2* dst = reinterpret_cast<double2*>(csp_B);
        const int cnt = (g_hi - g_lo) * (KP / 2);
        for (int e = threadIdx.x; e < cnt; e += CSP_THREADS) cp_async16(dst + e, src + e, true);
        cp_async_commit();
        cp_async_wait<0>();
    }
    __syncthreads();
    const int sub = threadIdx.x & 3;
    const int64_t quad = ((int64_t)blockIdx.x * CSP_THREADS + threadIdx.x) >> 2, nquad = ((int64_t)gridDim.x * CSP_THREADS) >> 2;
    for (int64_t cell = quad; cell < m; cell += nquad) {
        int64_t p = first ? indptr[cell] : cursor[cell];
        const int64_t p1 = indptr[cell + 1];
        double v[2 * KT];
#pragma unroll
        for (int i = 0; i < 2 * KT; i++) v[i] = 0.0;
        auto add_row = [&](int32_t g, double x) {
            const double2* r = reinterpret_cast<const double2*>(csp_B + (int64_t)(g - g_lo) * KP) + sub;
#pragma unroll
            for (int i = 0; i < KT; i++) { const double2 b = r[4 * i]; v[2 * i] = fma(x, b.x, v[2 * i]); v[2 * i + 1] = fma(x, b.y, v[2 * i + 1]); }
        };
        while (p + 7 < p1) {  // eight non-zeros per trip: all sixteen loads are issued before the range test (p + 7 < p1: in bounds)
            int32_t gg[8];
            double xx[8];
#pragma unroll
            for (int u = 0; u < 8; u++) { gg[u] = idx[p + u]; xx[u] = val[p + u]; }
            if (gg[7] >= g_hi) break;  // sorted: all eight are inside the range iff the last one is
#pragma unroll
            for (int u = 0; u < 8; u++) add_row(gg[u], xx[u]);
            p += 8;
        }
        while (p + 1 < p1) {  // the rest of the run two at a time
            const int32_t ga = idx[p], gb = idx[p + 1];
            const double xa = val[p], xb = val[p + 1];
            if (gb >= g_hi) break;
            add_row(ga, xa); add_row(gb, xb);
            p += 2;
        }
        while (p < p1) {
            const int32_t g = idx[p];
            if (g >= g_hi) break;
            add_row(g, val[p]);
            p++;
        }
        if (!last && sub == 0) cursor[cell] = p;
        const double sq = inv[cell];
        double* wr = W + cell * KP;
#pragma unroll
        for (int i = 0; i < KT; i++) {
            const int col = 2 * (sub + 4 * i);
            double2 acc = first ? make_double2(0.0, 0.0) : *reinterpret_cast<const double2*>(wr + col);
            acc.x += v[2 * i] * sq; acc.y += v[2 * i + 1] * sq;
            if (last) {  // finished: subtract center' B, zero the padding columns
                acc.x = col < k ? acc.x - w0[col] : 0.0;
                acc.y = col + 1 < k ? acc.y - w0[col + 1] : 0.0;
            }
            *reinterpret_cast<double2*>(wr + col) = acc;
        }
    }
}

template <int KT>
inline int csc_project_launch(scm_ctx* ctx, Scratch& ws, const int64_t* d_indptr, const int32_t* d_idx, const double* d_val, const double* d_inv,
                              int64_t m, int64_t n, const scm_coef* c, double* W) {
    constexpr int KP = KT * 8;
    static const int smem_kb = getenv("SCM_CSP_SMEM_KB") ? atoi(getenv("SCM_CSP_SMEM_KB")) : 128;  // shared-memory budget in KB.  Measured at 1M x 2000, k = 20 (projection + pass B): 200 KB (228 KB carve-out, 2 gene ranges) 6.90 ms,
                                                                                              // 192 KB 6.62, 160 KB 6.59, 128 KB (4 ranges) 6.55, 96 KB 6.61, 64 KB 6.77: the L1 that is left matters more than the launches
    int64_t per = ((int64_t)smem_kb * 1024) / (KP * 8);  // genes whose rows of B fit the shared memory of one CTA
    if (per > n) per = n;
    const int nr = (int)ceil_div(n, per);
    int64_t* cursor = nullptr;
    if (nr > 1) SCM_TRY(ws.alloc(&cursor, m));
    SCM_CUDA(cudaFuncSetAttribute(csc_project_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    for (int r = 0; r < nr; r++) {
        const int64_t lo = r * per, hi = lo + per < n ? lo + per : n;
        const size_t smem = (size_t)(hi - lo) * KP * 8;
        csc_project_kernel<KT><<<ctx->num_sms, CSP_THREADS, smem, ctx->stream>>>(d_indptr, d_idx, d_val, d_inv, m, c->Bpad, (int32_t)lo, (int32_t)hi,
                                                                              r == 0, r == nr - 1, c->w0, c->k, cursor, W);
        SCM_LAUNCH_CHECK(ctx);
    }
    return 0;
}

// ---- K6 on sparse rows ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// Cells per CTA tile = 8 * SCM_CSA_MT, CTAs per SM.  A/B on one box (pass B alone): 32 cells x 3 CTAs (156 registers) 6.19 ms,
// 16 cells x 4 CTAs (108 registers) 5.88, 16 x 5 (96 registers, 16 B spilled) 5.84, 16 x 6 (spills) 8.3, 8 x 8 9.5.
#ifndef SCM_CSA_MT
#define SCM_CSA_MT 2
#endif
#ifndef SCM_CSA_CTAS
#define SCM_CSA_CTAS 4
#endif
constexpr int CSA_WARPS = 4, CSA_MT = SCM_CSA_MT, CSA_ROWS = 8 * CSA_MT, CSA_SLAB = 256, CSA_CPW = CSA_ROWS / CSA_WARPS, CSA_CTAS = SCM_CSA_CTAS;  // tile: 16 cells x 256 genes x 8 B = 33 KB

template <int KT>
__global__ void __launch_bounds__(CSA_WARPS * 32, CSA_CTAS)
csc_adjust_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                  const double* __restrict__ inv, int64_t m, int64_t n, const double* __restrict__ Wg, const double* __restrict__ Afrag,
                  int k, int nblk, double* __restrict__ out, int64_t ldo, int vec_out, int dbg) {
    extern __shared__ __align__(16) uint8_t csa_smem[];
    double* tile = reinterpret_cast<double*>(csa_smem);                       // [32][CSA_SLAB + 2] (padded: conflict-free fragment reads)
    constexpr int TLD = CSA_SLAB + 2;
    double* Wsm = tile + CSA_ROWS * TLD;                                      // [32][KT * 8]
    int64_t* cur = reinterpret_cast<int64_t*>(Wsm + CSA_ROWS * KT * 8);       // [32] cursor into the cell's non-zeros
    int64_t* endp = cur + CSA_ROWS;                                           // [32]
    double* sc = reinterpret_cast<double*>(endp + CSA_ROWS);                  // [32] scale of the cell
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int64_t ntiles = ceil_div(m, CSA_ROWS);
    // the tile is all-zero at the start of every slab: zero-filled once here, and the DMMA phase clears what it reads
    for (int e = tid; e < CSA_ROWS * (TLD / 2); e += CSA_WARPS * 32) *reinterpret_cast<double2*>(tile + 2 * e) = make_double2(0.0, 0.0);
    (void)Wsm;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t row0 = t * CSA_ROWS;
        __syncthreads();  // previous tile fully consumed (its cursors are rewritten below)
        if (tid < CSA_ROWS) {
            const int64_t cell = row0 + tid;
            const bool ok = cell < m;
            cur[tid] = ok ? indptr[cell] : 0;
            endp[tid] = ok ? indptr[cell + 1] : 0;
            sc[tid] = ok ? inv[cell] : 0.0;
        }
        {   // the NEXT tile's non-zeros and rows of W into L2 while this one is processed: every later load of the tile is an L2
            // hit (~300 cycles) instead of an HBM access -- the kernel is bound by load latency, not by bandwidth (ncu: 42 % of
            // the warp samples on long_scoreboard at 12 warps per SM)
            const int64_t tn = t + gridDim.x;
            if (tn < ntiles) {
                const int64_t r0n = tn * CSA_ROWS, r1n = min(m, r0n + (int64_t)CSA_ROWS);
                const int64_t q0 = indptr[r0n], q1 = indptr[r1n];
                for (int64_t q = (q0 & ~(int64_t)31) + tid * 32; q < q1; q += CSA_WARPS * 32 * 32) prefetch_l2(idx + q);
                for (int64_t q = (q0 & ~(int64_t)15) + tid * 16; q < q1; q += CSA_WARPS * 32 * 16) prefetch_l2(val + q);
                const int64_t w0 = r0n * (KT * 8), w1 = r1n * (KT * 8);
                for (int64_t q = w0 + tid * 16; q < w1; q += CSA_WARPS * 32 * 16) prefetch_l2(Wg + q);
            }
        }
        // ---- W = (s x - center) B comes from csc_project_kernel (global memory, m x 8 KT): every lane reads its fragments directly
        double w[CSA_MT][KT][2];
#pragma unroll
        for (int mt = 0; mt < CSA_MT; mt++) {
            const int64_t cell = row0 + mt * 8 + gid;
#pragma unroll
            for (int nt = 0; nt < KT; nt++) {
                const double2 v = cell < m ? __ldg(reinterpret_cast<const double2*>(Wg + cell * (KT * 8) + 8 * nt + 2 * tig)) : make_double2(0.0, 0.0);
                w[mt][nt][0] = -v.x; w[mt][nt][1] = -v.y;  // pass B adds (-W) A
            }
        }
        // A fragments of this warp's first 16-gene block; inside the loop the next block's are always in flight (a warp's blocks
        // are warp, warp + 4, ... across the slabs: a slab holds 16 of them)
        double2 an[2 * KT];
        {
            const double2* af = reinterpret_cast<const double2*>(Afrag + (int64_t)warp * KT * 128) + lane;
#pragma unroll
            for (int i = 0; i < 2 * KT; i++) an[i] = warp < nblk ? __ldg(af + i * 32) : make_double2(0.0, 0.0);
        }
        __syncthreads();
        // ---- pass B, slab by slab
        for (int64_t g0 = 0; g0 < n && !(dbg & 2); g0 += CSA_SLAB) {
            const int width = (int)min((int64_t)CSA_SLAB, n - g0);
            // scatter: warp w walks the cursors of its 8 cells (indices are sorted: the slab's non-zeros are a contiguous run).  The
            // first 32 candidates of ALL eight cells are loaded before any of them is used (eight independent loads in flight per
            // lane instead of a chain of eight); a cell with more than 32 non-zeros in the slab finishes in the loop below.
            if (!(dbg & 64)) {
                int32_t gq[CSA_CPW];
                double xq[CSA_CPW];
#pragma unroll
                for (int r = 0; r < CSA_CPW; r++) {
                    const int c = warp * CSA_CPW + r;
                    const int64_t q = cur[c] + lane;
                    const bool ok = q < endp[c];
                    gq[r] = ok ? idx[q] : INT32_MAX;
                    xq[r] = ok ? val[q] : 0.0;
                }
#pragma unroll
                for (int r = 0; r < CSA_CPW; r++) {
                    const int c = warp * CSA_CPW + r;
                    const double s = sc[c];
                    const bool in = gq[r] < g0 + width;
                    if (in && gq[r] >= g0) tile[c * TLD + (gq[r] - (int32_t)g0)] = xq[r] * s;  // g < g0 only for un-sorted (unsupported) input
                    unsigned mask = __ballot_sync(0xffffffffu, in);
                    int64_t p = cur[c] + __popc(mask);
                    const int64_t p1 = endp[c];
                    while (mask == 0xffffffffu) {  // rare: more than 32 non-zeros of this cell in the slab
                        const int64_t q = p + lane;
                        const int32_t g = q < p1 ? idx[q] : INT32_MAX;
                        const bool in2 = g < g0 + width;
                        if (in2 && g >= g0) tile[c * TLD + (g - (int32_t)g0)] = val[q] * s;
                        mask = __ballot_sync(0xffffffffu, in2);
                        p += __popc(mask);
                    }
                    __syncwarp();
                    if (lane == 0) cur[c] = p;
                }
            }
            __syncthreads();
            // DMMA over this slab's 16-gene blocks, warps round-robin; the fragments read from the tile are cleared behind the read
            const int nb_slab = (width + 15) / 16;
            for (int b = warp; b < nb_slab; b += CSA_WARPS) {
                const int gb = (int)(g0 / 16) + b;
                double2 a[2 * KT];
#pragma unroll
                for (int i = 0; i < 2 * KT; i++) a[i] = an[i];
                if (gb + CSA_WARPS < nblk) {
                    const double2* af = reinterpret_cast<const double2*>(Afrag + (int64_t)((dbg & 8) ? 0 : gb + CSA_WARPS) * KT * 128) + lane;
#pragma unroll
                    for (int i = 0; i < 2 * KT; i++) an[i] = __ldg(af + i * 32);
                }
                double y[CSA_MT][4];
#pragma unroll
                for (int mt = 0; mt < CSA_MT; mt++) {
                    double* src = tile + (mt * 8 + gid) * TLD + b * 16 + 4 * tig;
                    const double2 a0 = *reinterpret_cast<const double2*>(src), c2 = *reinterpret_cast<const double2*>(src + 2);
                    *reinterpret_cast<double2*>(src) = make_double2(0.0, 0.0);
                    *reinterpret_cast<double2*>(src + 2) = make_double2(0.0, 0.0);
                    y[mt][0] = a0.x; y[mt][1] = a0.y; y[mt][2] = c2.x; y[mt][3] = c2.y;
                }
#pragma unroll
                for (int tau = 0; tau < 2; tau++)
#pragma unroll
                    for (int nt = 0; nt < KT; nt++)
#pragma unroll
                        for (int mt = 0; mt < CSA_MT; mt++) {
                            if (dbg & 32) { y[mt][2 * tau] += w[mt][nt][0] * a[tau * KT + nt].x; continue; }
                            dmma884(y[mt][2 * tau], y[mt][2 * tau + 1], w[mt][nt][0], a[tau * KT + nt].x);
                            dmma884(y[mt][2 * tau], y[mt][2 * tau + 1], w[mt][nt][1], a[tau * KT + nt].y);
                        }
                const int64_t gg = (int64_t)gb * 16 + 4 * tig;
#pragma unroll
                for (int mt = 0; mt < CSA_MT; mt++) {
                    const int64_t cell = row0 + mt * 8 + gid;
                    store4(out + (cell < m ? cell : 0) * ldo, gg, n, vec_out, cell < m && !(dbg & 16), y[mt]);
                }
            }
            __syncthreads();  // slab consumed and cleared before the next scatter
        }
    }
}

template <int KT>
inline int csc_adjust_red_launch(scm_ctx* ctx, Scratch& ws, const int64_t* d_indptr, const int32_t* d_idx, const double* d_val,
                                 const double* d_inv, int64_t m, int64_t n, const scm_coef* c, const double* W, double* d_out, int64_t ldo);

template <int KT>
inline int csc_adjust_launch(scm_ctx* ctx, const int64_t* d_indptr, const int32_t* d_idx, const double* d_val, const double* d_inv,
                             int64_t m, int64_t n, const scm_coef* c, double* d_out, int64_t ldo) {
    const int smem = (CSA_ROWS * (CSA_SLAB + 2) + CSA_ROWS * KT * 8) * 8 + CSA_ROWS * (8 + 8 + 8);
    SCM_CUDA(cudaFuncSetAttribute(csc_adjust_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t ntiles = ceil_div(m, CSA_ROWS);
    int64_t grid = (int64_t)ctx->num_sms * CSA_CTAS;
    if (grid > ntiles) grid = ntiles;
    const int vec_out = vec_flag(d_out, ldo);
    Scratch ws(ctx);
    double* W;
    SCM_TRY(ws.alloc(&W, m * KT * 8));
    const int dbg = getenv("SCM_CSC_DEBUG") ? atoi(getenv("SCM_CSC_DEBUG")) : 0;  // diagnostics (results are wrong with any bit set): 1 no projection, 2 no pass B, 8 A fragments from one block, 16 no stores, 32 no DMMA, 64 no scatter
    if (dbg & 1) SCM_TRY(zero_async(ctx, W, sizeof(double) * m * KT * 8));
    else SCM_TRY(csc_project_launch<KT>(ctx, ws, d_indptr, d_idx, d_val, d_inv, m, n, c, W));
    static const int use_red = getenv("SCM_CSC_RED") ? atoi(getenv("SCM_CSC_RED")) : 1;  // 0: the tile kernel (kept for A/B runs)
    if (use_red && !(dbg & ~1)) return csc_adjust_red_launch<KT>(ctx, ws, d_indptr, d_idx, d_val, d_inv, m, n, c, W, d_out, ldo);
    csc_adjust_kernel<KT><<<(unsigned)grid, CSA_WARPS * 32, smem, ctx->stream>>>(d_indptr, d_idx, d_val, d_inv, m, n, W, c->Afrag, c->k, c->nblk,
                                                                                d_out, ldo, vec_out, dbg);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

// ---- pass B without a tile: DMMA result stored, non-zeros added behind it with RED -------------------------------------------
// csc_adjust_kernel above is bound by load LATENCY: its tile round trip (scatter -> barrier -> fragment
// read -> clear -> barrier) serialises the candidate loads, the DMMAs and the stores of a CTA, and 12 warps per SM cannot fill the
// gaps.  This form has no tile, no barrier and no shared-memory traffic for the cells: a WARP owns 8 MT cells; per slab of
// genes it computes (-W) A with C = 0 and stores the result with ordinary (L2 write-back) stores, then walks the slab's
// non-zeros of its cells -- 32 consecutive candidates of a cell per load, coalesced -- and adds s x to the element just written
// with `red.global.add.f64` (fire and forget: no load latency anywhere on the scatter side; the line is still dirty in L2, so
// the RED costs no DRAM traffic).  Canonical cells have unique indices: every element receives at most one RED, the result
// is deterministic (fl(fl(-W A) + s x)).  The A fragments of a RANGE of gene blocks stay resident in shared memory for a
// whole launch (one persistent 16-warp CTA per SM, conflict-free LDS.128 in fragment order; one launch per range, the
// position where a cell's run ended is kept in `cursor` like in csc_project_kernel): through L1/L2 they would be 24 GB of
// re-reads at 16 cells per warp.
#ifndef SCM_CSR_MT
#define SCM_CSR_MT 1
#endif
#ifndef SCM_CSR_ORDER
#define SCM_CSR_ORDER 0
#endif
#ifndef SCM_CSR_PF
#define SCM_CSR_PF 1
#endif
#ifndef SCM_CSR_WARPS
#define SCM_CSR_WARPS 20
#endif
#ifndef SCM_CSR_SLAB
#define SCM_CSR_SLAB 128
#endif
constexpr int CSR_WARPS = SCM_CSR_WARPS, CSR_MT = SCM_CSR_MT, CSR_ROWS = 8 * CSR_MT, CSR_SLAB = SCM_CSR_SLAB;

template <int KT>
__global__ void __launch_bounds__(CSR_WARPS * 32, 1)
csc_adjust_red_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                      const double* __restrict__ inv, int64_t m, int64_t n, const double* __restrict__ Wg, const double* __restrict__ Afrag,
                      int blk_lo, int blk_hi, int first, int last, int64_t* __restrict__ cursor, double* __restrict__ out, int64_t ldo,
                      int vec_out) {
    extern __shared__ __align__(16) double csr_A[];  // [(blk_hi - blk_lo)][2 KT][32] double2, the layout of Afrag
    {
        const double2* src = reinterpret_cast<const double2*>(Afrag + (int64_t)blk_lo * KT * 128);
        double2* dst = reinterpret_cast<double2*>(csr_A);
        const int cnt = (blk_hi - blk_lo) * KT * 64;
        for (int e = threadIdx.x; e < cnt; e += CSR_WARPS * 32) cp_async16(dst + e, src + e, true);
        cp_async_commit();
        cp_async_wait<0>();
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int64_t nunits = ceil_div(m, CSR_ROWS);
    const int64_t gwarp = (int64_t)warp * gridDim.x + blockIdx.x, nwarps = (int64_t)gridDim.x * CSR_WARPS;  // neighbouring units on different SMs
    const int32_t g_lo = blk_lo * 16, g_hi = (int32_t)min((int64_t)blk_hi * 16, n);
    for (int64_t unit = gwarp; unit < nunits; unit += nwarps) {
        const int64_t row0 = unit * CSR_ROWS;
        // lane r < CSR_ROWS keeps the cursor, the end and the scale of cell row0 + r
        const int64_t cell_l = row0 + lane;
        const bool okc = lane < CSR_ROWS && cell_l < m;
        int64_t mycur = okc ? (first ? indptr[cell_l] : cursor[cell_l]) : 0;
        const int64_t myend = okc ? indptr[cell_l + 1] : 0;
        const double mysc = okc ? inv[cell_l] : 0.0;
#if SCM_CSR_PF
        {   // the next unit's rows of W, row pointers and scales into L2 while this one is processed (its start is two dependent
            // DRAM latencies otherwise, against ~8 slabs of work)
            const int64_t nr0 = (unit + nwarps) * CSR_ROWS;
            if (nr0 < m) {
                const int64_t wb = nr0 * (KT * 8), we = min(m, nr0 + CSR_ROWS) * (KT * 8);
                if (wb + lane * 16 < we) prefetch_l2(Wg + wb + lane * 16);
                if (lane == 31) { prefetch_l2(indptr + nr0); prefetch_l2(inv + nr0); if (!first) prefetch_l2(cursor + nr0); }
                if (lane == 30) prefetch_l2(indptr + min(m, nr0 + CSR_ROWS));
            }
        }
#endif
        double w[CSR_MT][KT][2];
#pragma unroll
        for (int mt = 0; mt < CSR_MT; mt++) {
            const int64_t cell = row0 + mt * 8 + gid;
#pragma unroll
            for (int nt = 0; nt < KT; nt++) {
                const double2 v = cell < m ? __ldg(reinterpret_cast<const double2*>(Wg + cell * (KT * 8) + 8 * nt + 2 * tig)) : make_double2(0.0, 0.0);
                w[mt][nt][0] = -v.x; w[mt][nt][1] = -v.y;
            }
        }
        // 32 consecutive candidates of every cell of the unit, from the cells' cursors (coalesced)
        auto load_cand = [&](int32_t (&gq)[CSR_ROWS], double (&xq)[CSR_ROWS]) {
#pragma unroll
            for (int r = 0; r < CSR_ROWS; r++) {
                const int64_t q = __shfl_sync(0xffffffffu, mycur, r) + lane;
                const bool ok = q < __shfl_sync(0xffffffffu, myend, r);
                gq[r] = ok ? idx[q] : INT32_MAX;
                xq[r] = ok ? val[q] : 0.0;
            }
        };
        // candidates inside [g0, gend) stay (scaled), the others become -1; cursors advance; a cell with more than 32 non-zeros in
        // the slab adds the rest at once
        auto classify = [&](int32_t (&gq)[CSR_ROWS], double (&xq)[CSR_ROWS], int32_t g0, int32_t gend) {
#pragma unroll
            for (int r = 0; r < CSR_ROWS; r++) {
                const double s = __shfl_sync(0xffffffffu, mysc, r);
                const bool in = gq[r] < gend;
                unsigned mask = __ballot_sync(0xffffffffu, in);
                int adv = __popc(mask);
                xq[r] *= s;
                gq[r] = (in && gq[r] >= g0) ? gq[r] : -1;
                if (mask == 0xffffffffu) {
                    double* rowp = out + (row0 + r) * ldo;
                    int64_t p = __shfl_sync(0xffffffffu, mycur, r) + 32;
                    const int64_t p1 = __shfl_sync(0xffffffffu, myend, r);
                    while (mask == 0xffffffffu) {
                        const int64_t q = p + lane;
                        const int32_t g = q < p1 ? idx[q] : INT32_MAX;
                        const bool in2 = g < gend;
                        if (in2 && g >= g0) atomicAdd(rowp + g, val[q] * s);
                        mask = __ballot_sync(0xffffffffu, in2);
                        const int c2 = __popc(mask);
                        p += c2; adv += c2;
                    }
                }
                if (lane == r) mycur += adv;
            }
        };
        auto red_cand = [&](const int32_t (&gq)[CSR_ROWS], const double (&xq)[CSR_ROWS]) {
#pragma unroll
            for (int r = 0; r < CSR_ROWS; r++)
                if (gq[r] >= 0) atomicAdd(out + (row0 + r) * ldo + gq[r], xq[r]);  // result unused: RED
        };
        // (-W) A for the slab's 16-gene blocks, stored with plain stores (the lines must stay in L2 for the REDs)
        auto dmma_slab = [&](int32_t g0, int32_t gend) {
            for (int gb = g0 >> 4; gb * 16 < gend; gb++) {
                const double2* af = reinterpret_cast<const double2*>(csr_A) + (int64_t)(gb - blk_lo) * KT * 64 + lane;
                double2 a[2 * KT];
#pragma unroll
                for (int i = 0; i < 2 * KT; i++) a[i] = af[i * 32];
                double y[CSR_MT][4];
#pragma unroll
                for (int mt = 0; mt < CSR_MT; mt++) { y[mt][0] = y[mt][1] = y[mt][2] = y[mt][3] = 0.0; }
#pragma unroll
                for (int tau = 0; tau < 2; tau++)
#pragma unroll
                    for (int nt = 0; nt < KT; nt++)
#pragma unroll
                        for (int mt = 0; mt < CSR_MT; mt++) {
                            dmma884(y[mt][2 * tau], y[mt][2 * tau + 1], w[mt][nt][0], a[tau * KT + nt].x);
                            dmma884(y[mt][2 * tau], y[mt][2 * tau + 1], w[mt][nt][1], a[tau * KT + nt].y);
                        }
                const int64_t gg = (int64_t)gb * 16 + 4 * tig;
#pragma unroll
                for (int mt = 0; mt < CSR_MT; mt++) {
                    const int64_t cell = row0 + mt * 8 + gid;
                    if (cell < m) {
                        double* rowp = out + cell * ldo;
                        if (vec_out == 2 && gg + 3 < n) {
                            stg256(rowp + gg, y[mt][0], y[mt][1], y[mt][2], y[mt][3]);
                        } else if (vec_out && gg + 3 < n) {
                            *reinterpret_cast<double2*>(rowp + gg) = make_double2(y[mt][0], y[mt][1]);
                            *reinterpret_cast<double2*>(rowp + gg + 2) = make_double2(y[mt][2], y[mt][3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; j++) if (gg + j < n) rowp[gg + j] = y[mt][j];
                        }
                    }
                }
            }
            __syncwarp();  // orders this warp's stores before its REDs to the same elements
        };
#if SCM_CSR_ORDER == 2
        // pipelined: the candidates of slab j are loaded one DMMA phase before they are classified, and added one DMMA phase after
        // the slab was stored (a RED that follows the store of its line at once is slow: measured)
        int32_t gc[CSR_ROWS], gs[CSR_ROWS];
        double xc[CSR_ROWS], xs[CSR_ROWS];
#pragma unroll
        for (int r = 0; r < CSR_ROWS; r++) { gs[r] = -1; xs[r] = 0.0; }
        load_cand(gc, xc);
        for (int32_t g0 = g_lo; g0 < g_hi; g0 += CSR_SLAB) {
            const int32_t gend = min(g_hi, g0 + CSR_SLAB);
            dmma_slab(g0, gend);
            classify(gc, xc, g0, gend);
            red_cand(gs, xs);
#pragma unroll
            for (int r = 0; r < CSR_ROWS; r++) { gs[r] = gc[r]; xs[r] = xc[r]; }
            if (g0 + CSR_SLAB < g_hi) load_cand(gc, xc);
        }
        red_cand(gs, xs);
#else
        for (int32_t g0 = g_lo; g0 < g_hi; g0 += CSR_SLAB) {
            const int32_t gend = min(g_hi, g0 + CSR_SLAB);
            int32_t gc[CSR_ROWS];
            double xc[CSR_ROWS];
            if (SCM_CSR_ORDER == 1) load_cand(gc, xc);
            dmma_slab(g0, gend);
            if (SCM_CSR_ORDER == 0) load_cand(gc, xc);
            classify(gc, xc, g0, gend);
            red_cand(gc, xc);
        }
#endif
        if (!last && okc) cursor[cell_l] = mycur;
    }
}

template <int KT>
inline int csc_adjust_red_launch(scm_ctx* ctx, Scratch& ws, const int64_t* d_indptr, const int32_t* d_idx, const double* d_val,
                                 const double* d_inv, int64_t m, int64_t n, const scm_coef* c, const double* W, double* d_out, int64_t ldo) {
    int per = (192 * 1024) / (KT * 1024);  // 16-gene blocks whose A fragments fit 192 KB: one byte more and the driver takes the 228 KB
                                           // carve-out, and with the L1 that is left the kernel runs 26 % slower per gene (measured)
    per -= per % (CSR_SLAB / 16);          // ranges start on a slab boundary
    if (per > c->nblk) per = c->nblk;
    const int nr = (int)ceil_div((int64_t)c->nblk, (int64_t)per);
    int64_t* cursor = nullptr;
    if (nr > 1) SCM_TRY(ws.alloc(&cursor, m));
    SCM_CUDA(cudaFuncSetAttribute(csc_adjust_red_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    const int vec_out = vec_flag(d_out, ldo);  // 2: one 32-byte store per lane and block
    const int64_t nunits = ceil_div(m, CSR_ROWS);
    int64_t grid = ctx->num_sms;
    if (grid * CSR_WARPS > nunits) grid = ceil_div(nunits, CSR_WARPS);
    for (int r = 0; r < nr; r++) {
        const int lo = r * per, hi = lo + per < c->nblk ? lo + per : c->nblk;
        const size_t smem = (size_t)(hi - lo) * KT * 1024;
        csc_adjust_red_kernel<KT><<<(unsigned)grid, CSR_WARPS * 32, smem, ctx->stream>>>(d_indptr, d_idx, d_val, d_inv, m, n, W, c->Afrag, lo, hi,
                                                                                         r == 0, r == nr - 1, cursor, d_out, ldo, vec_out);
        SCM_LAUNCH_CHECK(ctx);
    }
    return 0;
}

// out[i, :] = s_i x_i - ((s_i x_i - center) B) A for cells-compressed sparse input (s = d_inv).  out is dense m x n.
inline int csc_adjust(scm_ctx* ctx, const int64_t* d_indptr, const int32_t* d_idx, const double* d_val, const double* d_inv, int64_t m,
                      int64_t n, const scm_coef* c, double* d_out, int64_t ldo) {
    if (!c || c->n != n) return fail(SCM_ERR_INVALID, "csc_adjust: coefficient object does not match n=%lld", (long long)n);
    if (!d_indptr || !d_inv || !d_out || ldo < n) return fail(SCM_ERR_INVALID, "csc_adjust: bad arguments");
    if (m <= 0) return 0;
    if (n >= ((int64_t)1 << 31) - CSA_SLAB) return fail(SCM_ERR_UNSUPPORTED, "csc_adjust: too many genes");
    switch (c->KT) {
        case 1: return csc_adjust_launch<1>(ctx, d_indptr, d_idx, d_val, d_inv, m, n, c, d_out, ldo);
        case 2: return csc_adjust_launch<2>(ctx, d_indptr, d_idx, d_val, d_inv, m, n, c, d_out, ldo);
        case 3: return csc_adjust_launch<3>(ctx, d_indptr, d_idx, d_val, d_inv, m, n, c, d_out, ldo);
        case 4: return csc_adjust_launch<4>(ctx, d_indptr, d_idx, d_val, d_inv, m, n, c, d_out, ldo);
        default: return fail(SCM_ERR_UNSUPPORTED, "csc_adjust: k=%d > 32 is not supported on the sparse path (densify first)", c->k);
    }
}

// Sparse rows in, adjusted dense rows straight to the host: the rank-k update runs chunk by chunk into two staging buffers and
// every finished chunk goes down the copy stream while the next one is computed -- the dense matrix never exists on the device.
inline int64_t host_chunk_rows(int64_t ld);
inline int csc_adjust_download(scm_ctx* ctx, const int64_t* d_indptr, const int32_t* d_idx, const double* d_val, const double* d_inv,
                               int64_t m, int64_t n, const scm_coef* c, double* h_out, int64_t ldo) {
    if (!h_out || ldo < n) return fail(SCM_ERR_INVALID, "csc_adjust_download: bad arguments");
    if (m <= 0) return 0;
    cudaStream_t st = ctx->stream;
    if (!ctx->copy_stream_d2h) SCM_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream_d2h, cudaStreamNonBlocking));
    cudaStream_t cs = ctx->copy_stream_d2h;
    const int64_t ld = (n + 1) & ~(int64_t)1;
    int64_t rows_per_chunk = host_chunk_rows(ld);
    if (rows_per_chunk > m) rows_per_chunk = m;
    const int nchunks = (int)ceil_div(m, rows_per_chunk);
    Scratch ws(ctx);
    double* stage[2];
    SCM_TRY(ws.alloc(&stage[0], rows_per_chunk * ld));
    SCM_TRY(ws.alloc(&stage[1], rows_per_chunk * ld));
    cudaEvent_t evk[2], evc[2];
    for (int i = 0; i < 2; i++) {
        SCM_CUDA(cudaEventCreateWithFlags(&evk[i], cudaEventDisableTiming));
        SCM_CUDA(cudaEventCreateWithFlags(&evc[i], cudaEventDisableTiming));
    }
    struct Guard { cudaEvent_t *a, *b; ~Guard() { for (int i = 0; i < 2; i++) { cudaEventDestroy(a[i]); cudaEventDestroy(b[i]); } } } guard{evk, evc};
    SCM_CUDA(cudaStreamSynchronize(st));  // staging buffers allocated (stream-ordered) before the copy stream touches them
    for (int ch = 0; ch < nchunks; ch++) {
        const int b = ch & 1;
        const int64_t r0 = (int64_t)ch * rows_per_chunk, rows = (r0 + rows_per_chunk <= m ? rows_per_chunk : m - r0);
        if (ch >= 2) SCM_CUDA(cudaEventSynchronize(evc[b]));  // host-mediated: the staging buffer has been downloaded
        {
            StageTimer t(ctx, ST_ADJUST);
            SCM_TRY(csc_adjust(ctx, d_indptr + r0, d_idx, d_val, d_inv + r0, rows, n, c, stage[b], ld));
        }
        SCM_CUDA(cudaEventRecord(evk[b], st));
        SCM_CUDA(cudaStreamWaitEvent(cs, evk[b], 0));
        if (ldo == n && ld == n)
            SCM_CUDA(cudaMemcpyAsync(h_out + r0 * ldo, stage[b], (size_t)rows * n * 8, cudaMemcpyDeviceToHost, cs));
        else
            SCM_CUDA(cudaMemcpy2DAsync(h_out + r0 * ldo, ldo * 8, stage[b], ld * 8, n * 8, rows, cudaMemcpyDeviceToHost, cs));
        SCM_CUDA(cudaEventRecord(evc[b], cs));
    }
    SCM_CUDA(cudaStreamSynchronize(cs));
    SCM_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // namespace scm
