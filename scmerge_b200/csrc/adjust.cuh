// K5 + K6: adjustment coefficients and the fused rank-k update  newY = Y - ((Y - center) B) A.
//
// K6 is a streaming kernel whose arithmetic (2*k*(c+n) flop per cell, ~3 flop/byte at k = 20) is too much
// for the FP64 FMA pipe at HBM speed unless it is issued as tensor-core DMMA, so both products are
// DMMA.8x8x4 with the big operand (Y) taken STRAIGHT from global memory in fragment layout:
//   * each warp owns 8*MT cells and is fully autonomous (no shared memory, no barriers),
//   * a lane loads 32 contiguous bytes (4 genes) of a cell row; the 4 lanes of a quad cover 128 bytes, so a
//     warp-level load is 8 full cache lines.  The same 4 doubles serve 4 k-steps of pass A (W += Y_blk * B_blk:
//     the k index of an MMA may be permuted freely) or the C operands of 2 n-tiles of pass B (out = Y - W*A:
//     the n index may be permuted freely), so no shuffles or transposes are needed anywhere,
//   * the small operands (B: n x k, A: k x n) are pre-arranged in per-lane fragment order by coef_create, each
//     lane reads its 4*KT doubles contiguously (L1/L2 resident),
//   * pass A visits only the 16-gene blocks that contain a control gene; pass B re-reads the cell rows
//     (L2 hits for the part pass A touched), adds -W*A with W kept in registers, and streams the result out.
// Algorithmic traffic 8*m*n read + 8*m*n_out written; flops 2*m*k*(c + n_out).
#pragma once
#include "common.cuh"
#include "eig_topk.cuh"
#include "gemm_f64.cuh"

struct scm_coef {
    int64_t n = 0;
    int32_t k = 0, KT = 0;
    int32_t nblk = 0;        // 16-gene blocks
    int32_t nblk_ctl = 0;    // blocks containing a control gene
    double* Bmat = nullptr;  // n x k   (row-major)
    double* Amat = nullptr;  // k x n   (row-major, unit rows * post)
    double* w0 = nullptr;    // k       center' B
    double* Bfrag = nullptr; // [nblk][32][KT][4]
    double* Afrag = nullptr; // [nblk][32][2][KT][2]
    double* Bpad = nullptr;  // n x (KT*8), rows of Bmat zero-padded to a multiple of 8 (16-byte vector loads, sparse pass A)
    int32_t* ctl_blocks = nullptr;
};

namespace scm {

// AU[a, :] = alpha[a, :] / ||alpha[a, :]||,  ACU = AU masked to control genes.  One CTA per row.
__global__ void __launch_bounds__(256)
coef_rows_kernel(const double* __restrict__ alpha, int64_t n, const uint8_t* __restrict__ ctl,
                 double* __restrict__ AU, double* __restrict__ ACU) {
    __shared__ double red[8];
    __shared__ double s_inv;
    const int a = blockIdx.x;
    double s = 0.0;
    for (int64_t g = threadIdx.x; g < n; g += blockDim.x) { const double v = alpha[a * n + g]; s += v * v; }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0; for (int w = 0; w < 8; w++) t += red[w]; s_inv = t > 0 ? 1.0 / sqrt(t) : 0.0; }
    __syncthreads();
    const double inv = s_inv;
    for (int64_t g = threadIdx.x; g < n; g += blockDim.x) {
        const double v = alpha[a * n + g] * inv;
        AU[a * n + g] = v;
        ACU[a * n + g] = ctl[g] ? v : 0.0;
    }
}

// Bmat[g, :] *= pre[g];  Amat[a, g] = AU[a, g] * post[g]
__global__ void coef_scale_kernel(double* __restrict__ Bmat, double* __restrict__ Amat, const double* __restrict__ AU,
                                  int64_t n, int k, const double* __restrict__ pre, const double* __restrict__ post) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const double pr = pre ? pre[g] : 1.0, po = post ? post[g] : 1.0;
    for (int a = 0; a < k; a++) {
        Bmat[g * k + a] *= pr;
        Amat[a * n + g] = AU[a * n + g] * po;
    }
}

// w0[a] = sum_g center[g] Bmat[g, a]   (single CTA, one warp per column)
__global__ void __launch_bounds__(1024)
coef_w0_kernel(const double* __restrict__ Bmat, const double* __restrict__ center, int64_t n, int k, double* __restrict__ w0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int a = warp; a < k; a += 32) {
        double s = 0.0;
        if (center) for (int64_t g = lane; g < n; g += 32) s += center[g] * Bmat[g * k + a];
        s = warp_sum(s);
        if (lane == 0) w0[a] = s;
    }
}

// fragment-ordered copies (see header comment) + list of blocks that contain a control gene
__global__ void coef_frag_kernel(const double* __restrict__ Bmat, const double* __restrict__ Amat, int64_t n, int k, int KT,
                                 int nblk, double* __restrict__ Bfrag, double* __restrict__ Afrag) {
    const int gb = blockIdx.x, lane = threadIdx.x;
    if (lane >= 32) return;
    const int gid = lane >> 2, tig = lane & 3;
    // layout [block][q][lane] in 16-byte units (q < 2*KT): a warp-wide 16-byte load of unit q is 512 contiguous bytes
    // in global memory and conflict-free in shared memory
    double* bf = Bfrag + (int64_t)gb * KT * 128;
    for (int nt = 0; nt < KT; nt++)
        for (int j = 0; j < 4; j++) {
            const int64_t gene = (int64_t)gb * 16 + 4 * tig + j;
            const int kk = 8 * nt + gid;
            bf[((nt * 2 + (j >> 1)) * 32 + lane) * 2 + (j & 1)] = (gene < n && kk < k) ? Bmat[gene * k + kk] : 0.0;
        }
    double* af = Afrag + (int64_t)gb * KT * 128;
    for (int tau = 0; tau < 2; tau++)
        for (int nt = 0; nt < KT; nt++)
            for (int e = 0; e < 2; e++) {
                const int64_t gene = (int64_t)gb * 16 + 4 * (gid >> 1) + 2 * tau + (gid & 1);
                const int kk = 8 * nt + 2 * tig + e;
                af[((tau * KT + nt) * 32 + lane) * 2 + e] = (gene < n && kk < k) ? Amat[(int64_t)kk * n + gene] : 0.0;
            }
}
__global__ void coef_ctlblocks_kernel(const uint8_t* __restrict__ ctl, int64_t n, int nblk, int32_t* __restrict__ blocks,
                                      int32_t* __restrict__ count) {
    // single thread block, ordered compaction (nblk is small: n/16)
    __shared__ int s_flag[1024];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nblk; b0 += 1024) {
        const int b = b0 + threadIdx.x;
        int f = 0;
        if (b < nblk)
            for (int j = 0; j < 16; j++) { const int64_t g = (int64_t)b * 16 + j; if (g < n && ctl[g]) f = 1; }
        s_flag[threadIdx.x] = f;
        __syncthreads();
        if (threadIdx.x == 0) {
            int base = s_base;
            for (int t = 0; t < 1024 && b0 + t < nblk; t++) if (s_flag[t]) blocks[base++] = b0 + t;
            s_base = base;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = s_base;
}

// Bpad[g, a] = a < k ? Bmat[g, a] : 0
__global__ void coef_bpad_kernel(const double* __restrict__ Bmat, int64_t n, int k, int KP, double* __restrict__ Bpad) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * KP) return;
    const int64_t g = e / KP;
    const int a = (int)(e % KP);
    Bpad[e] = a < k ? Bmat[g * k + a] : 0.0;
}

inline int coef_destroy(scm_ctx* ctx, scm_coef* c) {
    if (!c) return 0;
    cudaStream_t st = ctx->stream;
    void* ptrs[] = {c->Bmat, c->Amat, c->w0, c->Bfrag, c->Afrag, c->Bpad, c->ctl_blocks};
    for (void* p : ptrs) if (p) cudaFreeAsync(p, st);
    delete c;
    return 0;
}

// Allocates the coefficient object (Bmat / Amat uninitialised).
inline int coef_alloc(scm_ctx* ctx, int64_t n, int32_t k, scm_coef** out) {
    if (k <= 0 || k > SCM_MAX_K) return fail(SCM_ERR_UNSUPPORTED, "coef: k=%d outside 1..%d", k, SCM_MAX_K);
    if (n <= 0) return fail(SCM_ERR_INVALID, "coef: bad n");
    cudaStream_t st = ctx->stream;
    scm_coef* c = new scm_coef();
    c->n = n; c->k = k; c->KT = (k + 7) / 8; c->nblk = (int32_t)ceil_div(n, 16);
    auto dalloc = [&](void** p, int64_t bytes) -> int {
        cudaError_t e = cudaMallocAsync(p, (size_t)bytes, st);
        return e == cudaSuccess ? 0 : fail(SCM_ERR_CUDA, "coef alloc failed: %s", cudaGetErrorString(e));
    };
    int rc = 0;
    rc |= dalloc((void**)&c->Bmat, sizeof(double) * n * k);
    rc |= dalloc((void**)&c->Amat, sizeof(double) * n * k);
    rc |= dalloc((void**)&c->w0, sizeof(double) * k);
    rc |= dalloc((void**)&c->Bfrag, sizeof(double) * (int64_t)c->nblk * 32 * c->KT * 4);
    rc |= dalloc((void**)&c->Afrag, sizeof(double) * (int64_t)c->nblk * 32 * c->KT * 4);
    rc |= dalloc((void**)&c->Bpad, sizeof(double) * n * c->KT * 8);
    rc |= dalloc((void**)&c->ctl_blocks, sizeof(int32_t) * (c->nblk + 1));
    if (rc) { coef_destroy(ctx, c); return SCM_ERR_CUDA; }
    *out = c;
    return 0;
}

// From Bmat / Amat: w0 = center' B, the fragment-ordered copies and the list of 16-gene blocks pass A must visit
// (`d_rows` marks the genes whose row of B may be non-zero).  Synchronises the stream (reads back the block count).
inline int coef_finalize(scm_ctx* ctx, scm_coef* c, const uint8_t* d_rows, const double* d_center) {
    cudaStream_t st = ctx->stream;
    coef_w0_kernel<<<1, 1024, 0, st>>>(c->Bmat, d_center, c->n, c->k, c->w0);
    SCM_LAUNCH_CHECK(ctx);
    coef_frag_kernel<<<c->nblk, 32, 0, st>>>(c->Bmat, c->Amat, c->n, c->k, c->KT, c->nblk, c->Bfrag, c->Afrag);
    SCM_LAUNCH_CHECK(ctx);
    coef_bpad_kernel<<<(unsigned)ceil_div(c->n * c->KT * 8, 256), 256, 0, st>>>(c->Bmat, c->n, c->k, c->KT * 8, c->Bpad);
    SCM_LAUNCH_CHECK(ctx);
    coef_ctlblocks_kernel<<<1, 1024, 0, st>>>(d_rows, c->n, c->nblk, c->ctl_blocks, c->ctl_blocks + c->nblk);
    SCM_LAUNCH_CHECK(ctx);
    int h = 0;
    SCM_CUDA(cudaMemcpyAsync(&h, c->ctl_blocks + c->nblk, sizeof(int), cudaMemcpyDeviceToHost, st));
    SCM_CUDA(cudaStreamSynchronize(st));
    c->nblk_ctl = h;
    return 0;
}

inline int coef_create(scm_ctx* ctx, const double* d_alpha, int64_t n, int32_t k, const uint8_t* d_ctl,
                       const double* d_center, const double* d_pre, const double* d_post, scm_coef** out) {
    if (n <= 0 || !d_alpha || !d_ctl) return fail(SCM_ERR_INVALID, "coef_create: bad arguments");
    cudaStream_t st = ctx->stream;
    scm_coef* c = nullptr;
    SCM_TRY(coef_alloc(ctx, n, k, &c));
    struct Guard { scm_ctx* ctx; scm_coef*& c; bool ok = false; ~Guard() { if (!ok) coef_destroy(ctx, c); } } guard{ctx, c};
    Scratch ws(ctx);
    double *AU, *ACU, *N, *Linv, *Ninv;
    int* flag;
    SCM_TRY(ws.alloc(&AU, n * k)); SCM_TRY(ws.alloc(&ACU, n * k));
    SCM_TRY(ws.alloc(&N, (int64_t)k * k)); SCM_TRY(ws.alloc(&Linv, (int64_t)k * k)); SCM_TRY(ws.alloc(&Ninv, (int64_t)k * k));
    SCM_TRY(ws.alloc(&flag, 2));
    SCM_CUDA(cudaMemsetAsync(flag, 0, sizeof(int) * 2, st));
    coef_rows_kernel<<<k, 256, 0, st>>>(d_alpha, n, d_ctl, AU, ACU);
    SCM_LAUNCH_CHECK(ctx);
    // N = ac ac'  (k x k), Cholesky, N^-1 = Linv' Linv, B = ac' N^-1
    SCM_TRY(gemm(ctx, k, k, (int)n, 1.0, rowmajor(ACU, n), transposed(ACU, n), 0.0, N, k));
    SCM_TRY(chol_inv_launch(ctx, N, k, 0.0, Linv, flag));
    SCM_TRY(gemm(ctx, k, k, k, 1.0, transposed(Linv, k), rowmajor(Linv, k), 0.0, Ninv, k));
    SCM_TRY(gemm(ctx, (int)n, k, k, 1.0, transposed(ACU, n), rowmajor(Ninv, k), 0.0, c->Bmat, k));
    coef_scale_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(c->Bmat, c->Amat, AU, n, k, d_pre, d_post);
    SCM_LAUNCH_CHECK(ctx);
    int h = 0;
    SCM_CUDA(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    SCM_TRY(coef_finalize(ctx, c, d_ctl, d_center));  // synchronises: h is valid afterwards
    if (h)
        return fail(SCM_ERR_SINGULAR, "coef_create: ac %%*%% t(ac) is not positive definite (k=%d exceeds the rank of the control-gene block?)", k);
    guard.ok = true;
    *out = c;
    return 0;
}

// Coefficient object from explicit factors: newY = Y - ((Y - center) B) A with B (n x k, rows outside `d_rows` must be
// zero) and A (k x n) given by the caller (scRUVg: B = 1[ctl] V diag(1/sigma), A = alpha).
inline int coef_from_factors(scm_ctx* ctx, const double* d_B, const double* d_A, int64_t n, int32_t k, const uint8_t* d_rows,
                             const double* d_center, scm_coef** out) {
    if (!d_B || !d_A || !d_rows || !out) return fail(SCM_ERR_INVALID, "coef_from_factors: bad arguments");
    scm_coef* c = nullptr;
    SCM_TRY(coef_alloc(ctx, n, k, &c));
    cudaError_t e1 = cudaMemcpyAsync(c->Bmat, d_B, sizeof(double) * n * k, cudaMemcpyDeviceToDevice, ctx->stream);
    cudaError_t e2 = cudaMemcpyAsync(c->Amat, d_A, sizeof(double) * n * k, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e1 != cudaSuccess || e2 != cudaSuccess) { coef_destroy(ctx, c); return fail(SCM_ERR_CUDA, "coef_from_factors: copy failed"); }
    const int rc = coef_finalize(ctx, c, d_rows, d_center);
    if (rc != 0) { coef_destroy(ctx, c); return rc; }
    *out = c;
    return 0;
}

// ---- K6 ---------------------------------------------------------------------------------------------
// `nc`: the matrix is not written by this kernel, so it may be read through the non-coherent path (ld.global.nc, streaming).  An
// IN-PLACE adjust (out == Y) must use ordinary loads: PTX leaves ld.global.nc undefined for locations the kernel also writes.
// `vec`: 0 scalar accesses, 1 two 16-byte accesses, 2 one 32-byte access (row base and leading dimension 32-byte aligned)
__device__ __forceinline__ void load4(const double* rowp, int64_t g0, int64_t n, int vec, bool rowok, bool nc, double (&v)[4]) {
    if (rowok && vec && g0 + 3 < n) {
        if (vec == 2 && nc) { ldg256_stream(rowp + g0, v[0], v[1], v[2], v[3]); return; }
        double2 a, b;
        if (nc) { a = ldg_stream2(rowp + g0); b = ldg_stream2(rowp + g0 + 2); }
        else { a = *reinterpret_cast<const double2*>(rowp + g0); b = *reinterpret_cast<const double2*>(rowp + g0 + 2); }
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = (rowok && g0 + j < n) ? (nc ? __ldg(rowp + g0 + j) : rowp[g0 + j]) : 0.0;
    }
}
__device__ __forceinline__ void store4(double* rowp, int64_t g0, int64_t n, int vec, bool rowok, const double (&v)[4]) {
    if (!rowok) return;
    if (vec == 2 && g0 + 3 < n) {
        stg256_stream(rowp + g0, v[0], v[1], v[2], v[3]);
    } else if (vec && g0 + 3 < n) {
        stg_stream2(rowp + g0, make_double2(v[0], v[1]));
        stg_stream2(rowp + g0 + 2, make_double2(v[2], v[3]));
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) if (g0 + j < n) rowp[g0 + j] = v[j];
    }
}
// 16-byte vector flag -> 2 when the 32-byte forms apply (SCM_NO_256=1 keeps the 16-byte pairs: A/B switch)
inline int vec_flag(const void* base, int64_t ld) {
    static const int no256 = getenv("SCM_NO_256") ? atoi(getenv("SCM_NO_256")) : 0;
    if ((ld % 2 != 0) || (((uintptr_t)base) % 16 != 0)) return 0;
    return (!no256 && ld % 4 == 0 && ((uintptr_t)base) % 32 == 0) ? 2 : 1;
}

constexpr int ADJ_WARPS = 4;

// One CTA = 8*MT cells.  Its 4 warps split the GENE blocks round-robin (warp w takes blocks w, w+4, ...), so the
// four warps touch adjacent 128-byte lines of the same rows, the part of the rows pass A has read is at most
// 8*MT rows x c genes per CTA (it is still in L2 when pass B re-reads it), and every warp keeps one block of
// loads in flight while it computes on the previous one (explicit double buffering: ya / yb).
template <int KT, int MT, bool PASS_B, bool GIVEN_W = false>  // GIVEN_W: no pass A, the rows of W (m x k) are read from `Wout`
__global__ void __launch_bounds__(ADJ_WARPS * 32, (KT <= 3) ? 3 : 2)
adjust_kernel(const double* Y /* may alias out */, int64_t m, int64_t n, int64_t ld, const double* __restrict__ Bfrag,
              const double* __restrict__ Afrag, const double* __restrict__ w0, const int32_t* __restrict__ ctl_blocks,
              int nblk_ctl, int nblk, int k, double* out, int64_t ldo, double* __restrict__ Wout, int vec_in, int vec_out, int nc) {
    __shared__ double s_part[GIVEN_W ? 1 : ADJ_WARPS][MT * KT * 2][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int64_t row0 = (int64_t)blockIdx.x * (8 * MT);
    const double* rowp[MT];
    bool rowok[MT];
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
        const int64_t r = row0 + mt * 8 + gid;
        rowok[mt] = r < m;
        rowp[mt] = Y + (rowok[mt] ? r : 0) * ld;
    }
    auto load_block = [&](double (&y)[MT][4], int gb) {
        const int64_t g0 = (int64_t)gb * 16 + 4 * tig;
#pragma unroll
        for (int mt = 0; mt < MT; mt++) load4(rowp[mt], g0, n, vec_in, rowok[mt], nc != 0, y[mt]);
    };
    // ---- pass A: partial W over this warp's control-gene blocks
    double w[MT][KT][2];
#pragma unroll
    for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < KT; nt++) w[mt][nt][0] = w[mt][nt][1] = 0.0;
    if constexpr (GIVEN_W) {
        (void)s_part;
#pragma unroll
        for (int nt = 0; nt < KT; nt++) {
            const int kk = 8 * nt + 2 * tig;
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                const double* wr = Wout + (row0 + mt * 8 + gid) * k;
                w[mt][nt][0] = (rowok[mt] && kk < k) ? -wr[kk] : 0.0;
                w[mt][nt][1] = (rowok[mt] && kk + 1 < k) ? -wr[kk + 1] : 0.0;
            }
        }
    } else {
        auto pass_a_block = [&](const double (&y)[MT][4], int gb) {
            const double2* bf = reinterpret_cast<const double2*>(Bfrag + (int64_t)gb * KT * 128) + lane;
    #pragma unroll
            for (int nt = 0; nt < KT; nt++) {
                const double2 b01 = __ldg(bf + (nt * 2) * 32), b23 = __ldg(bf + (nt * 2 + 1) * 32);
                const double b[4] = {b01.x, b01.y, b23.x, b23.y};
    #pragma unroll
                for (int j = 0; j < 4; j++)
    #pragma unroll
                    for (int mt = 0; mt < MT; mt++) dmma884(w[mt][nt][0], w[mt][nt][1], y[mt][j], b[j]);
            }
        };
        {
            double ya[MT][4], yb[MT][4];
            int bi = warp;
            if (bi < nblk_ctl) load_block(ya, ctl_blocks[bi]);
            while (bi < nblk_ctl) {
                int nb = bi + ADJ_WARPS;
                if (nb < nblk_ctl) load_block(yb, ctl_blocks[nb]);
                pass_a_block(ya, ctl_blocks[bi]);
                bi = nb;
                if (bi >= nblk_ctl) break;
                nb = bi + ADJ_WARPS;
                if (nb < nblk_ctl) load_block(ya, ctl_blocks[nb]);
                pass_a_block(yb, ctl_blocks[bi]);
                bi = nb;
            }
        }
        // ---- combine the four partial W tiles (fixed order => deterministic), subtract center' B
    #pragma unroll
        for (int mt = 0; mt < MT; mt++)
    #pragma unroll
            for (int nt = 0; nt < KT; nt++)
    #pragma unroll
                for (int e = 0; e < 2; e++) s_part[warp][(mt * KT + nt) * 2 + e][lane] = w[mt][nt][e];
        __syncthreads();
    #pragma unroll
        for (int nt = 0; nt < KT; nt++) {
            const int kk = 8 * nt + 2 * tig;
            const double c0 = kk < k ? w0[kk] : 0.0, c1 = kk + 1 < k ? w0[kk + 1] : 0.0;
    #pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                double t0 = 0.0, t1 = 0.0;
    #pragma unroll
                for (int ww = 0; ww < ADJ_WARPS; ww++) {
                    t0 += s_part[ww][(mt * KT + nt) * 2 + 0][lane];
                    t1 += s_part[ww][(mt * KT + nt) * 2 + 1][lane];
                }
                t0 -= c0; t1 -= c1;
                if (Wout && warp == 0 && rowok[mt]) {
                    double* wr = Wout + (row0 + mt * 8 + gid) * k;
                    if (kk < k) wr[kk] = t0;
                    if (kk + 1 < k) wr[kk + 1] = t1;
                }
                w[mt][nt][0] = -t0; w[mt][nt][1] = -t1;  // pass B adds (-W) A
            }
        }
    }
    if constexpr (PASS_B) {
        // ---- pass B: out = Y + (-W) A over this warp's gene blocks, 16 genes (two n-tiles) per block
        double* outp[MT];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) outp[mt] = out + (rowok[mt] ? row0 + mt * 8 + gid : 0) * ldo;
        auto pass_b_block = [&](double (&y)[MT][4], int gb) {
            const double2* af = reinterpret_cast<const double2*>(Afrag + (int64_t)gb * KT * 128) + lane;
            double2 a[2 * KT];
#pragma unroll
            for (int i = 0; i < 2 * KT; i++) a[i] = __ldg(af + i * 32);
#pragma unroll
            for (int tau = 0; tau < 2; tau++)
#pragma unroll
                for (int nt = 0; nt < KT; nt++)
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        dmma884(y[mt][2 * tau], y[mt][2 * tau + 1], w[mt][nt][0], a[tau * KT + nt].x);
                        dmma884(y[mt][2 * tau], y[mt][2 * tau + 1], w[mt][nt][1], a[tau * KT + nt].y);
                    }
            const int64_t g0 = (int64_t)gb * 16 + 4 * tig;
#pragma unroll
            for (int mt = 0; mt < MT; mt++) store4(outp[mt], g0, n, vec_out, rowok[mt], y[mt]);
        };
        double ya[MT][4], yb[MT][4];
        int gb = warp;
        if (gb < nblk) load_block(ya, gb);
        while (gb < nblk) {
            int nb = gb + ADJ_WARPS;
            if (nb < nblk) load_block(yb, nb);
            pass_b_block(ya, gb);
            gb = nb;
            if (gb >= nblk) break;
            nb = gb + ADJ_WARPS;
            if (nb < nblk) load_block(ya, nb);
            pass_b_block(yb, gb);
            gb = nb;
        }
    }
}

// out[i, o] = Y[i, sub[o]] - sum_a W[i, a] Amat[a, sub[o]]   (return_subset_genes path; W from a PASS_B=false run)
__global__ void __launch_bounds__(256)
adjust_subset_kernel(const double* __restrict__ Y, int64_t m, int64_t n, int64_t ld, const double* __restrict__ W, int k,
                     const double* __restrict__ Amat, const int32_t* __restrict__ sub, int nsub, double* __restrict__ out, int64_t ldo) {
    __shared__ double sw[SCM_MAX_K];
    const int64_t i = blockIdx.x;
    if (threadIdx.x < k) sw[threadIdx.x] = W[i * k + threadIdx.x];
    __syncthreads();
    for (int o = threadIdx.x; o < nsub; o += blockDim.x) {
        const int g = sub[o];
        // a gene index outside [0, n) never touches memory: the column comes out as NaN (hosts validate the subset before
        // uploading it; this is the guard for callers that do not)
        if ((unsigned)g >= (unsigned)n) { out[i * ldo + o] = __longlong_as_double(0x7ff8000000000000LL); continue; }
        double v = Y[i * ld + g];
        for (int a = 0; a < k; a++) v -= sw[a] * Amat[(int64_t)a * n + g];
        out[i * ldo + o] = v;
    }
}

template <int KT, int MT>
inline int adjust_launch(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, const scm_coef* c, bool pass_b,
                         double* d_out, int64_t ldo, double* d_W) {
    const int vec_in = vec_flag(d_Y, ld);
    const int vec_out = pass_b ? vec_flag(d_out, ldo) : 0;
    const int64_t rows_per_cta = 8 * MT;
    const unsigned grid = (unsigned)ceil_div(m, rows_per_cta);
    const int nc = (d_out == nullptr || (const double*)d_out != d_Y) ? 1 : 0;  // in place: ordinary loads (see load4)
    if (pass_b)
        adjust_kernel<KT, MT, true><<<grid, ADJ_WARPS * 32, 0, ctx->stream>>>(d_Y, m, n, ld, c->Bfrag, c->Afrag, c->w0, c->ctl_blocks,
                                                                             c->nblk_ctl, c->nblk, c->k, d_out, ldo, d_W, vec_in, vec_out, nc);
    else
        adjust_kernel<KT, MT, false><<<grid, ADJ_WARPS * 32, 0, ctx->stream>>>(d_Y, m, n, ld, c->Bfrag, c->Afrag, c->w0, c->ctl_blocks,
                                                                              c->nblk_ctl, c->nblk, c->k, d_out, ldo, d_W, vec_in, vec_out, nc);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

// out = Y - W A for GIVEN factors W (m x k, row-major) and A (k x n): pass B of K6 alone.  scRUVg with a re-used `fullW`, and
// scRUVg with eta, where W comes from RUV1(Y) but is subtracted from Y itself (R/scRUVg.R:59,77-80).
template <int KT, int MT>
inline int rankk_subtract_launch(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, const scm_coef* c, const double* d_W,
                                 double* d_out, int64_t ldo) {
    const int nc = ((const double*)d_out != d_Y) ? 1 : 0;
    adjust_kernel<KT, MT, true, true><<<(unsigned)ceil_div(m, (int64_t)8 * MT), ADJ_WARPS * 32, 0, ctx->stream>>>(
        d_Y, m, n, ld, c->Bfrag, c->Afrag, c->w0, c->ctl_blocks, 0, c->nblk, c->k, d_out, ldo, const_cast<double*>(d_W), vec_flag(d_Y, ld),
        vec_flag(d_out, ldo), nc);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}
inline int rankk_subtract(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, const double* d_W, int32_t k, const double* d_A,
                          double* d_out, int64_t ldo) {
    if (!d_Y || !d_W || !d_A || !d_out || n <= 0 || m < 0 || ld < n || ldo < n || k <= 0) return fail(SCM_ERR_INVALID, "rankk_subtract: bad arguments");
    if (k > SCM_MAX_K) return fail(SCM_ERR_UNSUPPORTED, "rankk_subtract: k=%d > %d", k, SCM_MAX_K);
    if (m == 0) return 0;
    scm_coef* c = nullptr;
    SCM_TRY(coef_alloc(ctx, n, k, &c));
    struct Guard { scm_ctx* ctx; scm_coef* c; ~Guard() { coef_destroy(ctx, c); } } guard{ctx, c};
    Scratch ws(ctx);
    uint8_t* rows;
    SCM_TRY(ws.alloc(&rows, n + 4));
    SCM_TRY(zero_async(ctx, rows, (n + 3) & ~(int64_t)3));
    SCM_TRY(zero_async(ctx, c->Bmat, sizeof(double) * n * k));
    SCM_CUDA(cudaMemcpyAsync(c->Amat, d_A, sizeof(double) * n * k, cudaMemcpyDeviceToDevice, ctx->stream));
    SCM_TRY(coef_finalize(ctx, c, rows, nullptr));
    int rc;
    switch (c->KT) {
        case 1: rc = rankk_subtract_launch<1, 4>(ctx, d_Y, m, n, ld, c, d_W, d_out, ldo); break;
        case 2: rc = rankk_subtract_launch<2, 4>(ctx, d_Y, m, n, ld, c, d_W, d_out, ldo); break;
        case 3: rc = rankk_subtract_launch<3, 4>(ctx, d_Y, m, n, ld, c, d_W, d_out, ldo); break;
        case 4: rc = rankk_subtract_launch<4, 4>(ctx, d_Y, m, n, ld, c, d_W, d_out, ldo); break;
        case 5: rc = rankk_subtract_launch<5, 2>(ctx, d_Y, m, n, ld, c, d_W, d_out, ldo); break;
        case 6: rc = rankk_subtract_launch<6, 2>(ctx, d_Y, m, n, ld, c, d_W, d_out, ldo); break;
        case 7: rc = rankk_subtract_launch<7, 2>(ctx, d_Y, m, n, ld, c, d_W, d_out, ldo); break;
        case 8: rc = rankk_subtract_launch<8, 2>(ctx, d_Y, m, n, ld, c, d_W, d_out, ldo); break;
        default: return fail(SCM_ERR_UNSUPPORTED, "rankk_subtract: k=%d not supported", k);
    }
    return rc;  // the coefficient object is released in stream order
}

// TMA-staged variant (adjust_tma.cuh): returns 1 when not applicable
inline int adjust_tma(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, const scm_coef* c, double* d_out, int64_t ldo);

inline int adjust(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, const scm_coef* c,
                  const int32_t* d_subset, int32_t nsub, double* d_out, int64_t ldo, double* d_W) {
    if (!c || c->n != n) return fail(SCM_ERR_INVALID, "adjust: coefficient object does not match n=%lld", (long long)n);
    if (m <= 0) return 0;
    if (!d_out && (d_subset || !d_W)) return fail(SCM_ERR_INVALID, "adjust: no output requested");
    if (!d_subset && !d_W) {
        const int trc = adjust_tma(ctx, d_Y, m, n, ld, c, d_out, ldo);
        if (trc <= 0) return trc;
    }
    Scratch ws(ctx);
    const bool subset = d_subset != nullptr;
    double* W = d_W;
    if (subset && !W) SCM_TRY(ws.alloc(&W, m * c->k));
    const bool pass_b = !subset && d_out != nullptr;  // d_out == NULL: projection only (W = (Y - center) B)
    int rc;
    switch (c->KT) {
        case 1: rc = adjust_launch<1, 4>(ctx, d_Y, m, n, ld, c, pass_b, d_out, ldo, W); break;
        case 2: rc = adjust_launch<2, 4>(ctx, d_Y, m, n, ld, c, pass_b, d_out, ldo, W); break;
        case 3: rc = adjust_launch<3, 4>(ctx, d_Y, m, n, ld, c, pass_b, d_out, ldo, W); break;
        case 4: rc = adjust_launch<4, 4>(ctx, d_Y, m, n, ld, c, pass_b, d_out, ldo, W); break;
        case 5: rc = adjust_launch<5, 2>(ctx, d_Y, m, n, ld, c, pass_b, d_out, ldo, W); break;
        case 6: rc = adjust_launch<6, 2>(ctx, d_Y, m, n, ld, c, pass_b, d_out, ldo, W); break;
        case 7: rc = adjust_launch<7, 2>(ctx, d_Y, m, n, ld, c, pass_b, d_out, ldo, W); break;
        case 8: rc = adjust_launch<8, 2>(ctx, d_Y, m, n, ld, c, pass_b, d_out, ldo, W); break;
        default: return fail(SCM_ERR_UNSUPPORTED, "adjust: k=%d not supported", c->k);
    }
    SCM_TRY(rc);
    if (subset) {
        if (m > 2147483647) return fail(SCM_ERR_UNSUPPORTED, "adjust: subset path limited to 2^31 rows");
        adjust_subset_kernel<<<(unsigned)m, 256, 0, ctx->stream>>>(d_Y, m, n, ld, W, c->k, c->Amat, d_subset, nsub, d_out, ldo);
        SCM_LAUNCH_CHECK(ctx);
    }
    return 0;
}

}  // namespace scm
