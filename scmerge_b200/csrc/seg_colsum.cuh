// K1: segmented column sums by int32 key (replicate group / batch / pseudobulk id).
//
// HBM-bound streaming reduction, deterministic:
//   1. rows are ordered by key with a stable radix sort of (key, row) pairs (CUB, index plumbing only),
//   2. every group is cut into chunks of <= SEG_R rows; one CTA sums one chunk x 512-column slab with
//      16-byte streaming loads of whole row segments (rows are gathered through the permutation, but each
//      gathered piece is 4 KB contiguous, so DRAM efficiency is that of a plain stream),
//   3. chunk partials of a group are added in chunk order.
// Algorithmic traffic: 8*m*n bytes read (+4*m keys); partials add <= 2/SEG_R of that.
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace scm {

constexpr int SEG_R = 128;        // rows per chunk
constexpr int SEG_THREADS = 256;  // threads per CTA, each owns VEC adjacent columns

// Maps keys to sort keys (skipped rows -> G), emits the identity payload and the group histogram.
// Each CTA covers SEG_PREP_ROWS rows; with G < 4096 the histogram is first built in shared memory.
constexpr int SEG_PREP_ROWS = 4096;
__global__ void seg_prepare_kernel(const int32_t* __restrict__ key, int64_t m, int32_t G,
                                   uint32_t* __restrict__ skey, int32_t* __restrict__ sval, int* __restrict__ cnt,
                                   int use_smem) {
    extern __shared__ int s_hist[];
    if (use_smem) {
        for (int i = threadIdx.x; i <= G; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
    }
    const int64_t base = (int64_t)blockIdx.x * SEG_PREP_ROWS;
    for (int r = threadIdx.x; r < SEG_PREP_ROWS; r += blockDim.x) {
        const int64_t i = base + r;
        if (i >= m) break;
        const int32_t k = key[i];
        const uint32_t kk = (k < 0 || k >= G) ? (uint32_t)G : (uint32_t)k;  // out-of-range keys are skipped rows
        skey[i] = kk;
        sval[i] = (int32_t)i;
        atomicAdd(use_smem ? &s_hist[kk] : &cnt[kk], 1);
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i <= G; i += blockDim.x)
            if (s_hist[i]) atomicAdd(&cnt[i], s_hist[i]);
    }
}

// Single CTA: exclusive scans over the G group counts.
//   seg_start[g]  : first position of group g in the sorted order           (G+1 entries)
//   chunk_off[g]  : first chunk id of group g                               (G+1 entries)
//   pslot_off[g]  : first partial slot of group g (only groups with >1 chunk use slots) (G+1 entries)
__global__ void seg_scan_kernel(const int* __restrict__ cnt, int32_t G, int64_t* __restrict__ seg_start,
                                int* __restrict__ chunk_off, int* __restrict__ pslot_off, int64_t* __restrict__ counts_out) {
    __shared__ int64_t s_start[1024];
    __shared__ int s_chunk[1024], s_slot[1024];
    __shared__ int64_t base_start;
    __shared__ int base_chunk, base_slot;
    if (threadIdx.x == 0) { base_start = 0; base_chunk = 0; base_slot = 0; }
    __syncthreads();
    for (int32_t g0 = 0; g0 < G; g0 += 1024) {
        int32_t g = g0 + threadIdx.x;
        int c = (g < G) ? cnt[g] : 0;
        int nch = (c + SEG_R - 1) / SEG_R;
        int nsl = nch > 1 ? nch : 0;
        s_start[threadIdx.x] = c; s_chunk[threadIdx.x] = nch; s_slot[threadIdx.x] = nsl;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele inclusive scan
            int64_t a = 0; int b = 0, d = 0;
            if ((int)threadIdx.x >= off) { a = s_start[threadIdx.x - off]; b = s_chunk[threadIdx.x - off]; d = s_slot[threadIdx.x - off]; }
            __syncthreads();
            s_start[threadIdx.x] += a; s_chunk[threadIdx.x] += b; s_slot[threadIdx.x] += d;
            __syncthreads();
        }
        if (g < G) {
            seg_start[g] = base_start + s_start[threadIdx.x] - c;
            chunk_off[g] = base_chunk + s_chunk[threadIdx.x] - nch;
            pslot_off[g] = base_slot + s_slot[threadIdx.x] - nsl;
            if (counts_out) counts_out[g] = c;
        }
        __syncthreads();
        if (threadIdx.x == 1023) { base_start += s_start[1023]; base_chunk += s_chunk[1023]; base_slot += s_slot[1023]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { seg_start[G] = base_start; chunk_off[G] = base_chunk; pslot_off[G] = base_slot; }
}

// COPY: every visited element is also written, shifted, to copy_out[row, col] = Y[row, col] - copy_shift[col] (the pre-shifted
// copy of the cells the fit's Gram reads): the stream that is being summed anyway pays for the copy's read.
// VEC columns per thread: 4 = one 32-byte access per row (LDG / STG.E.256 on sm_100: half the L1TEX requests of VEC = 2), 2 = 16 bytes,
// 1 = scalar (odd leading dimension).  A column is summed by ONE thread in row order whatever VEC is: the sums do not depend on it.
template <int VEC, bool SQUARE, bool COPY = false>
__global__ void __launch_bounds__(SEG_THREADS, VEC == 4 ? 2 : 1)
seg_partial_kernel(const double* __restrict__ Y, int64_t n, int64_t ld, const int32_t* __restrict__ perm,
                   const int64_t* __restrict__ seg_start, const int* __restrict__ chunk_off,
                   const int* __restrict__ pslot_off, int32_t G, const double* __restrict__ center,
                   double* __restrict__ partial, double* __restrict__ sums, int* __restrict__ flag,
                   const double* __restrict__ copy_shift = nullptr, double* __restrict__ copy_out = nullptr, int64_t ldc = 0) {
    const int nchunks = chunk_off[G];
    const int64_t col = ((int64_t)blockIdx.x * SEG_THREADS + threadIdx.x) * VEC;
    const bool v0 = col < n, full = col + VEC <= n;  // `full`: all VEC columns exist (vector access); else column by column
    bool bad = false;
    double h[VEC];
#pragma unroll
    for (int e = 0; e < VEC; e++) h[e] = (COPY && col + e < n) ? copy_shift[col + e] : 0.0;
    for (int chunk = blockIdx.y; chunk < nchunks; chunk += gridDim.y) {
        int lo = 0, hi = G;  // largest g with chunk_off[g] <= chunk (empty groups share offsets: take the last)
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (chunk_off[mid] <= chunk) lo = mid; else hi = mid;
        }
        const int g = lo;
        const int j = chunk - chunk_off[g];
        const int64_t start = seg_start[g] + (int64_t)j * SEG_R;
        const int len = (int)min((int64_t)SEG_R, seg_start[g + 1] - start);
        double c[VEC], acc[VEC];
#pragma unroll
        for (int e = 0; e < VEC; e++) { c[e] = (SQUARE && col + e < n) ? center[(int64_t)g * n + col + e] : 0.0; acc[e] = 0.0; }
        if (v0) {
            for (int r = 0; r < len; r += 8) {
                double x[8][VEC];
                int rows[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const bool ok = r + u < len;
                    const int64_t row = ok ? (int64_t)perm[start + r + u] : 0;
                    rows[u] = (int)row;
                    const double* p = Y + row * ld + col;
#pragma unroll
                    for (int e = 0; e < VEC; e++) x[u][e] = 0.0;
                    if (ok) {
                        if (VEC == 4 && full) ldg256_stream(p, x[u][0], x[u][1], x[u][2], x[u][3]);
                        else if (VEC == 2 && full) { const double2 t = ldg_stream2(p); x[u][0] = t.x; x[u][1] = t.y; }
                        else {
#pragma unroll
                            for (int e = 0; e < VEC; e++) if (col + e < n) x[u][e] = __ldg(p + e);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    if (r + u < len) {
#pragma unroll
                        for (int e = 0; e < VEC; e++) {
                            bad |= nonfinite(x[u][e]);
                            if (SQUARE) { const double d = x[u][e] - c[e]; acc[e] += d * d; }
                            else acc[e] += x[u][e];
                        }
                        if (COPY) {
                            double* q = copy_out + (int64_t)rows[u] * ldc + col;
                            if (VEC == 4 && full) stg256_stream(q, x[u][0] - h[0], x[u][1] - h[1], x[u][2] - h[2], x[u][3] - h[3]);
                            else if (VEC == 2 && full) __stcs(reinterpret_cast<double2*>(q), make_double2(x[u][0] - h[0], x[u][1] - h[1]));
                            else {
#pragma unroll
                                for (int e = 0; e < VEC; e++) if (col + e < n) __stcs(q + e, x[u][e] - h[e]);
                            }
                        }
                    }
                }
            }
            const int nch = chunk_off[g + 1] - chunk_off[g];
            double* dst = (nch == 1) ? (sums + (int64_t)g * n) : (partial + (int64_t)(pslot_off[g] + j) * n);
#pragma unroll
            for (int e = 0; e < VEC; e++) if (col + e < n) dst[col + e] = acc[e];
        }
    }
    if (bad && flag) *flag = 1;
}

// Rows whose key is outside [0, G) are not visited by seg_partial_kernel: they sit behind seg_start[G] in the sorted order and
// still belong in the shifted copy (the Gram is over ALL cells).
__global__ void seg_copy_skipped_kernel(const double* __restrict__ Y, int64_t m, int64_t n, int64_t ld, const int32_t* __restrict__ perm,
                                        const int64_t* __restrict__ seg_start, int32_t G, const double* __restrict__ shift,
                                        double* __restrict__ out, int64_t ldo) {
    for (int64_t pos = seg_start[G] + blockIdx.x; pos < m; pos += gridDim.x) {
        const int64_t row = perm[pos];
        for (int64_t j = threadIdx.x; j < n; j += blockDim.x) out[row * ldo + j] = Y[row * ld + j] - shift[j];
    }
}

// sums[g, col] = sum of the group's chunk partials in chunk order (groups with 0 rows -> 0, 1 chunk -> already written)
__global__ void seg_final_kernel(const double* __restrict__ partial, const int* __restrict__ chunk_off,
                                 const int* __restrict__ pslot_off, int64_t n, double* __restrict__ sums) {
    const int g = blockIdx.x;
    const int nch = chunk_off[g + 1] - chunk_off[g];
    if (nch == 1) return;
    const int64_t col = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (col >= n) return;
    double acc = 0.0;
    const double* p = partial + (int64_t)pslot_off[g] * n + col;
    for (int c = 0; c < nch; c++) acc += p[(int64_t)c * n];
    sums[(int64_t)g * n + col] = acc;
}

// Index plumbing shared by the dense and the sparse segmented sums: rows ordered by key (stable radix sort), group starts,
// chunk / partial-slot offsets.  All buffers live in the caller's Scratch.
struct SegPlan {
    int32_t* perm = nullptr;       // row order by key (m)
    int64_t* seg_start = nullptr;  // G + 1
    int* chunk_off = nullptr;      // G + 1
    int* pslot_off = nullptr;      // G + 1
};

inline int seg_plan_build(scm_ctx* ctx, Scratch& ws, const int32_t* d_key, int64_t m, int32_t G, int64_t* d_counts, SegPlan& sp) {
    cudaStream_t st = ctx->stream;
    uint32_t *skey_in, *skey_out;
    int32_t *sval_in, *perm;
    int *cnt, *chunk_off, *pslot_off;
    int64_t* seg_start;
    SCM_TRY(ws.alloc(&skey_in, m)); SCM_TRY(ws.alloc(&skey_out, m));
    SCM_TRY(ws.alloc(&sval_in, m)); SCM_TRY(ws.alloc(&perm, m));
    SCM_TRY(ws.alloc(&cnt, (int64_t)G + 1)); SCM_TRY(ws.alloc(&chunk_off, (int64_t)G + 1));
    SCM_TRY(ws.alloc(&pslot_off, (int64_t)G + 1)); SCM_TRY(ws.alloc(&seg_start, (int64_t)G + 1));
    SCM_TRY(zero_async(ctx, cnt, sizeof(int) * ((int64_t)G + 1)));
    const int use_smem = G < 4096;
    seg_prepare_kernel<<<(unsigned)ceil_div(m, SEG_PREP_ROWS), 256, use_smem ? sizeof(int) * (G + 1) : 0, st>>>(
        d_key, m, G, skey_in, sval_in, cnt, use_smem);
    SCM_LAUNCH_CHECK(ctx);
    seg_scan_kernel<<<1, 1024, 0, st>>>(cnt, G, seg_start, chunk_off, pslot_off, d_counts);
    SCM_LAUNCH_CHECK(ctx);
    int end_bit = 1;
    while (((int64_t)1 << end_bit) <= G) end_bit++;
    size_t tmp_bytes = 0;
    SCM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, skey_in, skey_out, sval_in, perm, (int)m, 0, end_bit, st));
    char* tmp;
    SCM_TRY(ws.alloc(&tmp, (int64_t)tmp_bytes));
    SCM_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, skey_in, skey_out, sval_in, perm, (int)m, 0, end_bit, st));
    // CUB's launches are counted from its documented structure, not observed: histogram + exclusive sum + one onesweep pass per
    // 8 key bits (the launch lists under profiles/ show exactly these three kernels for G < 256)
    ctx->launches += 2 + (end_bit + 7) / 8;
    sp.perm = perm; sp.seg_start = seg_start; sp.chunk_off = chunk_off; sp.pslot_off = pslot_off;
    return 0;
}

// d_copy_out (optional, m x ldc, ldc even when ld is): receives Y - d_copy_shift (n) for EVERY row, written from the summing pass.
inline int seg_colsum(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, const int32_t* d_key,
                      int32_t G, const double* d_center, double* d_sums, int64_t* d_counts, int32_t* d_nonfinite,
                      const double* d_copy_shift = nullptr, double* d_copy_out = nullptr, int64_t ldc = 0) {
    if (m < 0 || n <= 0 || G <= 0 || ld < n) return fail(SCM_ERR_INVALID, "seg_colsum: bad shape m=%lld n=%lld ld=%lld G=%d", (long long)m, (long long)n, (long long)ld, G);
    if (m >= (int64_t)1 << 31) return fail(SCM_ERR_UNSUPPORTED, "seg_colsum: m >= 2^31 rows per rank");
    cudaStream_t st = ctx->stream;
    Scratch ws(ctx);
    if (m == 0) {
        SCM_TRY(zero_async(ctx, d_sums, sizeof(double) * G * n));
        if (d_counts) SCM_TRY(zero_async(ctx, d_counts, sizeof(int64_t) * G));
        return 0;
    }
    SegPlan sp;
    SCM_TRY(seg_plan_build(ctx, ws, d_key, m, G, d_counts, sp));
    int32_t* perm = sp.perm;
    int64_t* seg_start = sp.seg_start;
    int *chunk_off = sp.chunk_off, *pslot_off = sp.pslot_off;

    const int64_t max_slots = 2 * ceil_div(m, SEG_R) + 1;
    double* partial;
    SCM_TRY(ws.alloc(&partial, max_slots * n));
    SCM_TRY(zero_async(ctx, d_sums, sizeof(double) * G * n));  // empty groups stay 0
    static const int no256 = getenv("SCM_NO_256") ? atoi(getenv("SCM_NO_256")) : 0;  // A/B switch: 16-byte accesses only
    const bool vec2 = (ld % 2 == 0) && (((uintptr_t)d_Y) % 16 == 0);
    bool vec4 = !no256 && (ld % 4 == 0) && (((uintptr_t)d_Y) % 32 == 0);
    if (d_copy_out) vec4 = vec4 && (ldc % 4 == 0) && (((uintptr_t)d_copy_out) % 32 == 0);
    const bool vec2c = vec2 && (!d_copy_out || ((ldc % 2 == 0) && (((uintptr_t)d_copy_out) % 16 == 0)));
    const int vec = vec4 ? 4 : (vec2c ? 2 : 1);
    dim3 grid((unsigned)ceil_div(n, (int64_t)SEG_THREADS * vec), 1);
    const int64_t max_chunks = ceil_div(m, SEG_R) + G;
    int64_t gy = ((int64_t)ctx->num_sms * 8) / grid.x;
    if (gy < 1) gy = 1;
    if (gy > max_chunks) gy = max_chunks;
    if (gy > 65535) gy = 65535;
    grid.y = (unsigned)gy;
    const bool sq = d_center != nullptr;
#define SEG_LAUNCH(V, S, C) seg_partial_kernel<V, S, C><<<grid, SEG_THREADS, 0, st>>>(d_Y, n, ld, perm, seg_start, chunk_off, pslot_off, G, d_center, partial, d_sums, d_nonfinite, d_copy_shift, d_copy_out, ldc)
    if (d_copy_out) {
        if (sq || !d_copy_shift || ldc < n) return fail(SCM_ERR_INVALID, "seg_colsum: the shifted copy needs a shift vector, ldc >= n and the plain-sum mode");
        if (vec == 4) SEG_LAUNCH(4, false, true); else if (vec == 2) SEG_LAUNCH(2, false, true); else SEG_LAUNCH(1, false, true);
        SCM_LAUNCH_CHECK(ctx);
        seg_copy_skipped_kernel<<<ctx->num_sms * 4, 256, 0, st>>>(d_Y, m, n, ld, perm, seg_start, G, d_copy_shift, d_copy_out, ldc);
    }
    else if (vec == 4) { if (sq) SEG_LAUNCH(4, true, false); else SEG_LAUNCH(4, false, false); }
    else if (vec == 2) { if (sq) SEG_LAUNCH(2, true, false); else SEG_LAUNCH(2, false, false); }
    else               { if (sq) SEG_LAUNCH(1, true, false); else SEG_LAUNCH(1, false, false); }
#undef SEG_LAUNCH
    SCM_LAUNCH_CHECK(ctx);
    dim3 fgrid((unsigned)G, (unsigned)ceil_div(n, 256));
    seg_final_kernel<<<fgrid, 256, 0, st>>>(partial, chunk_off, pslot_off, n, d_sums);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace scm
