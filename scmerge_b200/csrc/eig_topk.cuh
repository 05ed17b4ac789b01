// K4: top-s eigenpairs of the n x n symmetric PSD Gram by block subspace iteration.
//
//   X <- orth(random n x l)                                   l = s + oversampling (<= SCM_MAX_BLOCK)
//   X <- orth(G X), twice                                     two products before the first Rayleigh-Ritz step (both modes)
//   repeat:  Z = G X ; H = X'Z ; (Q, theta) = eig(H)          Rayleigh-Ritz, single-CTA one-sided Jacobi
//            W = Z Q ; V = X Q ; r_i = ||W_i - theta_i V_i||   true residuals of the Ritz pairs
//            stop when ||R_kconv||_F / (theta_kconv - theta_kconv+1) <= tol (EXACT; Davis-Kahan bound on the
//            subspace angle); RANDOMIZED stops after its q + 1 = 3 products (rsvd: p = 10, q = 2)
//            W <- (G Theta^-1)^(pw-1) W Theta^-1 ; X <- CholeskyQR2(W)
// All n-sized products run on the DMMA GEMM; the l x l factorizations are single-CTA shared-memory kernels.
//
// Control flow lives on the DEVICE: a one-thread kernel (eig_check_kernel) evaluates the stopping rule and the number of extra
// products of the next iteration into a control block, and every later kernel returns at once when the block says "done" (or
// "this extra product is not needed").  The host therefore enqueues iterations blindly, two at a time, and reads the control
// block back once per batch: a converged fit costs ONE host synchronisation instead of one per iteration, and a skipped
// iteration costs a few microseconds of empty launches (strong scaling: this loop is replicated on every rank).
#pragma once
#include "common.cuh"
#include "gemm_f64.cuh"

namespace scm {

struct EigCtl {
    int done;       // 0 running, 1 converged, 2 stagnated, 3 non-finite input, 4 iteration limit      (first two ints = the skip
    int pw;         // products of the current outer iteration (1..4)                                    words of gemm_f64.cuh)
    int it;         // Rayleigh-Ritz steps finished
    int conv_rank;  // largest r <= s whose top-r subspace passed the rank test (set when done)
    int jac_sweeps;
    int chol_flag;  // a Cholesky pivot was not positive
    int products;   // G * block products so far
    int pass1_only; // CholeskyQR passes that were enough on their own (diagnostics)
    int skip2[2];   // skip words of the SECOND CholeskyQR pass: {1 = not needed (or the solve is done), large}
    int pad2[2];
    double worst, gap, rmax, top;
    double hist[16];
};

struct EigRanks {
    int n;
    int k[17];
};

__device__ __forceinline__ bool eig_skip(const int* skip, int need) { return skip && (skip[0] != 0 || skip[1] < need); }

// ---- random start block ---------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__global__ void randn_kernel(double* __restrict__ X, int64_t count, uint64_t seed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint64_t a = splitmix64(seed ^ (uint64_t)(2 * i)), b = splitmix64(seed ^ (uint64_t)(2 * i + 1));
    const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740993.0);  // (0, 1)
    const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
    X[i] = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// ---- single-CTA Cholesky + inverse of the factor ----------------------------------------------------
// S (l x l, symmetric, row-major ld = l) -> Linv = inverse of the lower Cholesky factor of S + shift*maxdiag*I.
// A non-positive pivot is replaced by tiny*maxdiag and reported through `flag`.
// Right-looking elimination WITHOUT scaling the pivot column (A[i][k] -= A[i][j] A[k][j] / d_j; L[i][j] = A[i][j] / sqrt(d_j) is
// applied once at the end): column j and d_j are read-only during step j, so a step needs ONE barrier (the scaled form needs
// three and a serial sqrt on one thread).  16 x 16 threads tile the trailing block, no index divisions.
// 120 us -> see DESIGN.md K4 table for the measured time at l = 80.
constexpr int CHOL_THREADS = 256;
__global__ void __launch_bounds__(CHOL_THREADS)
chol_inv_kernel(const double* __restrict__ S, int l, double shift_rel, double* __restrict__ Linv, int* __restrict__ flag,
                const int* __restrict__ skip, int need) {
    if (eig_skip(skip, need)) return;
    extern __shared__ double sm[];
    const int ld = l + 1;
    double* A = sm;            // l x ld   (becomes L, lower)
    double* X = sm + l * ld;   // l x ld   (L^-1, lower)
    __shared__ double s_maxdiag;
    __shared__ double s_rd[SCM_MAX_BLOCK + 2];  // 1 / sqrt(d_j)
    const int tid = threadIdx.x, nt = blockDim.x;
    const int tx = tid & 15, ty = tid >> 4;
    for (int idx = tid; idx < l * l; idx += nt) {
        const int i = idx / l, j = idx % l;
        A[i * ld + j] = 0.5 * (S[i * l + j] + S[j * l + i]);
        X[i * ld + j] = 0.0;
    }
    __syncthreads();
    if (tid < 32) {
        double md = 0.0;
        for (int i = tid; i < l; i += 32) md = fmax(md, A[i * ld + i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) md = fmax(md, __shfl_xor_sync(0xffffffffu, md, o));
        if (tid == 0) s_maxdiag = md;
    }
    __syncthreads();
    const double maxdiag = s_maxdiag;
    const double floor_d = 1e-280 + 1e-30 * maxdiag;
    for (int i = tid; i < l; i += nt) A[i * ld + i] += shift_rel * maxdiag;
    for (int j = 0; j < l; j++) {
        __syncthreads();
        double d = A[j * ld + j];
        if (!(d > floor_d)) { d = floor_d; if (tid == 0 && flag) *flag = 1; }
        if (tid == 0) s_rd[j] = rsqrt(d);
        const double inv_d = 1.0 / d;
        for (int i = j + 1 + ty; i < l; i += 16) {
            const double aij = A[i * ld + j] * inv_d;
            for (int k = j + 1 + tx; k <= i; k += 16) A[i * ld + k] -= aij * A[k * ld + j];
        }
    }
    __syncthreads();
    // L[i][j] = A[i][j] / sqrt(d_j);  L[j][j] = sqrt(d_j) = d_j * rd_j (the diagonal entry may have been floored: use 1 / rd_j)
    for (int idx = tid; idx < l * l; idx += nt) {
        const int i = idx / l, j = idx % l;
        if (j < i) A[i * ld + j] *= s_rd[j];
    }
    __syncthreads();
    // forward substitution, thread c owns column c of L^-1: X[i][c] = (delta_ic - sum_{k=c}^{i-1} L[i][k] X[k][c]) / L[i][i],
    // 1 / L[i][i] = rd_i
    for (int i = 0; i < l; i++) {
        if (tid < l && tid <= i) {
            const int c = tid;
            double s0 = (i == c) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;  // four partial sums break the FMA dependency chain
            int k = c;
            for (; k + 3 < i; k += 4) {
                s0 -= A[i * ld + k] * X[k * ld + c];
                s1 -= A[i * ld + k + 1] * X[(k + 1) * ld + c];
                s2 -= A[i * ld + k + 2] * X[(k + 2) * ld + c];
                s3 -= A[i * ld + k + 3] * X[(k + 3) * ld + c];
            }
            for (; k < i; k++) s0 -= A[i * ld + k] * X[k * ld + c];
            X[i * ld + c] = ((s0 + s1) + (s2 + s3)) * s_rd[i];
        }
        __syncthreads();
    }
    for (int idx = tid; idx < l * l; idx += nt) {
        const int i = idx / l, j = idx % l;
        Linv[idx] = (j <= i) ? X[i * ld + j] : 0.0;
    }
}

inline int chol_inv_launch(scm_ctx* ctx, const double* S, int l, double shift_rel, double* Linv, int* flag, const int* skip = nullptr,
                           int need = 0) {
    if (l > SCM_MAX_BLOCK) return fail(SCM_ERR_UNSUPPORTED, "chol_inv: l=%d > %d", l, SCM_MAX_BLOCK);
    const size_t smem = sizeof(double) * 2 * l * (l + 1);
    SCM_CUDA(cudaFuncSetAttribute(chol_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));  // per device, cheap
    chol_inv_kernel<<<1, CHOL_THREADS, smem, ctx->stream>>>(S, l, shift_rel, Linv, flag, skip, need);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

// ---- CholeskyQR building blocks of the eigensolver: register-tiled factorization + row-parallel triangular solve -------
// chol_inv_kernel above keeps its whole matrix in shared memory and walks it with dependent load / FMA / store chains
// (170 us at l = 80, eight of them per fit: more than a third of K4).  Here the 16 x 16 threads of the CTA own the elements
// (i, k) = (ty + 16 a, tx + 16 b) of the symmetric matrix in REGISTERS; a step broadcasts the pivot column through a
// double-buffered shared-memory vector (one barrier) and updates the trailing block with TL x TL register FMAs.  The factor
// goes out as L (row-major, lower); X L^-T is then a triangular SOLVE per row of X (trsm_rows_kernel, one thread per row, the
// row in shared memory), so the inverse of L is never formed.
template <int TL>
__global__ void __launch_bounds__(256)
chol_factor_kernel(const double* __restrict__ S, int l, double shift_rel, double* __restrict__ Lout, int* __restrict__ flag,
                   const int* __restrict__ skip, int* __restrict__ pass2_words, int* __restrict__ pass1_count) {
    if (eig_skip(skip, 0)) {
        if (pass2_words && threadIdx.x == 0) { pass2_words[0] = 1; pass2_words[1] = 1 << 30; }  // the whole solve is over: no second pass either
        return;
    }
    __shared__ double colbuf[2][SCM_MAX_BLOCK + 16];
    __shared__ double s_red[8];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    double a[TL][TL];
    double md = 0.0;
#pragma unroll
    for (int ii = 0; ii < TL; ii++)
#pragma unroll
        for (int kk = 0; kk < TL; kk++) {
            const int i = ty + 16 * ii, k = tx + 16 * kk;
            a[ii][kk] = (i < l && k < l) ? 0.5 * (S[i * l + k] + S[k * l + i]) : 0.0;
            if (i == k && i < l) md = fmax(md, a[ii][kk]);
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) md = fmax(md, __shfl_xor_sync(0xffffffffu, md, o));
    if ((tid & 31) == 0) s_red[tid >> 5] = md;
    __syncthreads();
    double maxdiag = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) maxdiag = fmax(maxdiag, s_red[w]);
    const double floor_d = 1e-280 + 1e-30 * maxdiag;
    double dmax = 0.0, dmin = 1e300;  // extreme pivots: d_max / d_min bounds the condition number of S from below
#pragma unroll
    for (int ii = 0; ii < TL; ii++)
#pragma unroll
        for (int kk = 0; kk < TL; kk++)
            if (ty + 16 * ii == tx + 16 * kk) a[ii][kk] += shift_rel * maxdiag;
    // The owner of column j holds it in a[.][j >> 4]: the block index of the register must be a compile-time constant (a
    // run-time index sends the whole array to local memory: 80 LDL/STL in the first version, 83 us at l = 80), so the column
    // loop is split into an unrolled loop over 16-column blocks and a run-time loop inside the block.
#pragma unroll
    for (int kb = 0; kb < TL; kb++) {
        for (int jj = 0; jj < 16; jj++) {
            const int j = 16 * kb + jj;
            if (j >= l) break;
            double* cb = colbuf[j & 1];
            if (tx == jj) {  // the threads that own column j publish it
#pragma unroll
                for (int ii = 0; ii < TL; ii++) cb[ty + 16 * ii] = a[ii][kb];
            }
            __syncthreads();
            double d = cb[j];
            if (!(d > floor_d)) { d = floor_d; if (tid == 0 && flag) *flag = 1; }
            dmax = fmax(dmax, d); dmin = fmin(dmin, d);
            const double inv_d = 1.0 / d, rd = rsqrt(d);
            double ci[TL], ck[TL];
#pragma unroll
            for (int t = 0; t < TL; t++) { ci[t] = cb[ty + 16 * t]; ck[t] = cb[tx + 16 * t]; }
#pragma unroll
            for (int ii = 0; ii < TL; ii++) {
                const double f = (ty + 16 * ii > j) ? ci[ii] * inv_d : 0.0;  // rows <= j are final
#pragma unroll
                for (int kk = 0; kk < TL; kk++)
                    if (kk >= kb && tx + 16 * kk > j) a[ii][kk] -= f * ck[kk];  // column blocks left of the pivot's are final
            }
            if (tid < l) Lout[tid * l + j] = tid > j ? cb[tid] * rd : (tid == j ? d * rd : 0.0);  // column j of L (0 above the diagonal)
        }
    }
    // One CholeskyQR pass leaves an orthogonality error of about eps * cond(S).  With cond(S) <~ 1e3 (a random start block, a
    // Ritz-scaled block) that is rounding level and the second pass would change nothing: it is switched off on the device.
    if (pass2_words && tid == 0) {
        const int enough = dmax <= 1e3 * dmin;
        pass2_words[0] = enough; pass2_words[1] = 1 << 30;
        if (enough && pass1_count) *pass1_count += 1;
    }
}

inline int chol_factor_launch(scm_ctx* ctx, const double* S, int l, double shift_rel, double* Lout, int* flag, const int* skip,
                              int* pass2_words = nullptr, int* pass1_count = nullptr) {
    if (l > SCM_MAX_BLOCK) return fail(SCM_ERR_UNSUPPORTED, "chol_factor: l=%d > %d", l, SCM_MAX_BLOCK);
    cudaStream_t st = ctx->stream;
    switch ((l + 15) / 16) {
        case 1: chol_factor_kernel<1><<<1, 256, 0, st>>>(S, l, shift_rel, Lout, flag, skip, pass2_words, pass1_count); break;
        case 2: chol_factor_kernel<2><<<1, 256, 0, st>>>(S, l, shift_rel, Lout, flag, skip, pass2_words, pass1_count); break;
        case 3: chol_factor_kernel<3><<<1, 256, 0, st>>>(S, l, shift_rel, Lout, flag, skip, pass2_words, pass1_count); break;
        case 4: chol_factor_kernel<4><<<1, 256, 0, st>>>(S, l, shift_rel, Lout, flag, skip, pass2_words, pass1_count); break;
        case 5: chol_factor_kernel<5><<<1, 256, 0, st>>>(S, l, shift_rel, Lout, flag, skip, pass2_words, pass1_count); break;
        case 6: chol_factor_kernel<6><<<1, 256, 0, st>>>(S, l, shift_rel, Lout, flag, skip, pass2_words, pass1_count); break;
        default: chol_factor_kernel<7><<<1, 256, 0, st>>>(S, l, shift_rel, Lout, flag, skip, pass2_words, pass1_count); break;
    }
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

// dst[r, :] = src[r, :] L^-T  (rows of an n x l row-major block; L lower triangular l x l row-major; in place allowed).
// Forward substitution per row, x_j = (b_j - sum_{k<j} x_k L[j][k]) / L[j][j]: a latency chain of l steps, so the kernel is
// built for many short chains in flight: FOUR lanes share a row (lane q sums the k = q mod 4 terms, two shuffles combine them),
// a CTA of 128 threads owns 32 rows (63 CTAs at n = 2000, ~3 per SM), the rows sit in shared memory column-major over the rows
// (conflict-free), L is read as broadcasts, and the loads of four terms are issued before their FMAs.  First version: one
// thread per row, 111 us at l = 80 (the compiler reused one register pair for consecutive loads: one term in flight).
constexpr int TRSM_ROWS = 32;
__global__ void __launch_bounds__(4 * TRSM_ROWS)
trsm_rows_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t n, int l, const double* __restrict__ L,
                 const int* __restrict__ skip) {
    if (eig_skip(skip, 0)) return;
    extern __shared__ double sm[];
    constexpr int RP = TRSM_ROWS + 1;  // odd row stride: the transposing loads / stores spread over the banks
    const int tid = threadIdx.x, nt = blockDim.x;
    double* Ls = sm;                   // l x l, row-major
    double* rd = sm + l * l;           // 1 / L[j][j]
    double* xs = rd + ((l + 1) & ~1);  // [l][RP]
    for (int e = tid; e < l * l; e += nt) cp_async8(Ls + e, L + e, true);  // asynchronous: every element in flight at once (a plain
    cp_async_commit();                                                      // load -> store loop waits a DRAM latency per element)
    const int64_t r0 = (int64_t)blockIdx.x * TRSM_ROWS;
    const int rows = (int)min((int64_t)TRSM_ROWS, n - r0);
    const int lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    // a warp per row, lanes along the row: coalesced, and no integer division (a 64-bit e / l per element costs more than the solve)
    for (int rw0 = warp; rw0 < rows; rw0 += 4 * nwarp) {  // four rows per warp and trip: their loads are issued together
        double v[4][4];
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int rw = rw0 + u * nwarp, col = lane + 32 * c;
                v[u][c] = (rw < rows && col < l) ? src[(r0 + rw) * l + col] : 0.0;
            }
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int rw = rw0 + u * nwarp, col = lane + 32 * c;
                if (rw < rows && col < l) xs[col * RP + rw] = v[u][c];
            }
    }
    cp_async_wait<0>();
    __syncthreads();
    for (int j = tid; j < l; j += nt) rd[j] = 1.0 / Ls[j * l + j];
    __syncthreads();
    const int row = tid >> 2, q = tid & 3;  // the four lanes of a row are adjacent lanes of one warp
    {   // every lane runs the loop (full-mask shuffles); rows past the end of the block work on unused shared memory
        for (int j = 0; j < l; j++) {
            const double* Lj = Ls + j * l;
            double s0 = 0.0, s1 = 0.0;
            int k = q;
            for (; k + 12 < j; k += 16) {  // four terms per lane and trip: all eight loads first
                const double x0 = xs[k * RP + row], x1 = xs[(k + 4) * RP + row], x2 = xs[(k + 8) * RP + row], x3 = xs[(k + 12) * RP + row];
                const double l0 = Lj[k], l1 = Lj[k + 4], l2 = Lj[k + 8], l3 = Lj[k + 12];
                s0 = fma(x0, l0, s0); s1 = fma(x1, l1, s1); s0 = fma(x2, l2, s0); s1 = fma(x3, l3, s1);
            }
            for (; k < j; k += 4) s0 = fma(xs[k * RP + row], Lj[k], s0);
            double sum = s0 + s1;
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            if (q == 0) xs[j * RP + row] = (xs[j * RP + row] - sum) * rd[j];
            __syncwarp();  // x_j is read by the other three lanes of the row from the next step on
        }
    }
    __syncthreads();
    for (int rw = warp; rw < rows; rw += nwarp)
        for (int col = lane; col < l; col += 32) dst[(r0 + rw) * l + col] = xs[col * RP + rw];
}

inline int trsm_rows_launch(scm_ctx* ctx, const double* src, double* dst, int64_t n, int l, const double* L, const int* skip) {
    const size_t smem = sizeof(double) * ((size_t)l * l + ((l + 1) & ~1) + (size_t)l * (TRSM_ROWS + 1));
    SCM_CUDA(cudaFuncSetAttribute(trsm_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));  // per device, cheap
    trsm_rows_kernel<<<(unsigned)ceil_div(n, TRSM_ROWS), 4 * TRSM_ROWS, smem, ctx->stream>>>(src, dst, n, l, L, skip);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

// ---- single-CTA symmetric eigensolver: one-sided (Hestenes) Jacobi ---------------------------------
// H (l x l symmetric PSD, row-major) -> theta (descending), Q (row-major l x l, column j = eigenvector j).
// One half-warp (16 lanes) per column pair: all <= 56 disjoint pairs of a round run concurrently on the 64 half-warps.
// Squared column norms are cached (recomputed once per sweep, updated by the Rutishauser formulas ||a_p'||^2 = alpha - t*gamma,
// ||a_q'||^2 = beta + t*gamma), so a pair costs one dot product + the rotation.
// A round is a pure latency chain (dot -> shuffle tree -> rotation parameters -> rotation -> barrier), ~1000 of them per solve,
// so everything in it is shortened (ncu r02: 3100 cycles per round before):
//   * the (p, q) tournament schedule is a shared-memory table built once (no integer modulo in the loop),
//   * NT = ceil(l / 16) is a template parameter: the dot product and the rotation are fully unrolled, all loads of a phase
//     are in flight together,
//   * the rotation parameters use two reciprocal square roots and no division / sqrt (below).
// `ktop` / `jtol_bulk`: pairs of two columns >= ktop only have to be orthogonal to jtol_bulk (columns >= ktop = the unconverged
// noise-bulk directions of a warm-started step).  MEASURED AND SWITCHED OFF (ktop = l by default): a looser bulk tolerance makes
// the solve SLOWER -- 6 sweeps at 1e-5 (same as uniform), 11 at 1e-3, 34 at 1e-2 (tools/probe/eig_probe.py, SCM_EIG_JTOL_BULK).
// Rotating a leading column against a bulk column q1 mixes q1 into it, and q1's leftover non-orthogonality to the other bulk
// columns comes along: the leading pairs then converge linearly at the bulk tolerance instead of quadratically.
template <int NT>
__global__ void __launch_bounds__(1024)
jacobi_eig_kernel(const double* __restrict__ H, int l, double* __restrict__ Q, double* __restrict__ theta,
                  int max_sweeps, double jtol, int ktop, double jtol_bulk, int* __restrict__ sweeps_out, const int* __restrict__ skip) {
    if (eig_skip(skip, 0)) return;
    extern __shared__ double sm[];
    const int lp = (l + 1) & ~1;  // even number of columns (a zero column pads odd l)
    double* A = sm;               // column-major: column j at A + j*cs
    const int cs = lp + 1;        // column stride: odd, so the two half-warps of a warp hit different banks
    double* V = sm + lp * cs;
    unsigned* pairs = reinterpret_cast<unsigned*>(V + lp * cs);  // [(lp - 1) rounds][npair]: p | q << 16
    __shared__ int s_rot;
    __shared__ int s_order[SCM_MAX_BLOCK + 2];
    __shared__ double s_norm[SCM_MAX_BLOCK + 2];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    const int npair = lp / 2;
    for (int idx = tid; idx < lp * lp; idx += nt) {
        const int j = idx / lp, i = idx % lp;
        A[j * cs + i] = (i < l && j < l) ? 0.5 * (H[i * l + j] + H[j * l + i]) : 0.0;
        V[j * cs + i] = (i == j) ? 1.0 : 0.0;
    }
    for (int idx = tid; idx < (lp - 1) * npair; idx += nt) {  // round-robin tournament: every pair exactly once per sweep
        const int r = idx / npair, h = idx % npair;
        int p, q;
        if (h == 0) { p = lp - 1; q = r; }
        else { p = (r + h) % (lp - 1); q = (r - h + (lp - 1)) % (lp - 1); }
        pairs[idx] = (unsigned)p | ((unsigned)q << 16);
    }
    __syncthreads();
    const int hw = tid >> 4, hl = tid & 15;
    const bool act = hw < npair;
    double* nrm2 = s_norm;  // reused as the squared-norm cache during the sweeps
    int sweep = 0;
    for (; sweep < max_sweeps; sweep++) {
        if (tid == 0) s_rot = 0;
        for (int j = warp; j < lp; j += nwarp) {
            double a = 0.0;
            for (int i = lane; i < lp; i += 32) a += A[j * cs + i] * A[j * cs + i];
            a = warp_sum(a);
            if (lane == 0) nrm2[j] = a;
        }
        __syncthreads();
        for (int r = 0; r < lp - 1; r++) {
            const unsigned pq = act ? pairs[r * npair + hw] : 0u;
            const int p = (int)(pq & 0xffffu), q = (int)(pq >> 16);
            double* ap = A + p * cs; double* aq = A + q * cs;
            double x[NT], y[NT];
#pragma unroll
            for (int t = 0; t < NT; t++) {
                const int i = hl + 16 * t;
                const bool ok = act && i < lp;
                x[t] = ok ? ap[i] : 0.0;
                y[t] = ok ? aq[i] : 0.0;
            }
            double ga = 0.0, gb = 0.0;
#pragma unroll
            for (int t = 0; t < NT; t += 2) { ga += x[t] * y[t]; if (t + 1 < NT) gb += x[t + 1] * y[t + 1]; }
            ga += gb;
            const double al = nrm2[p], be = nrm2[q];  // read before the shuffles: they order it against lane 0's update
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) ga += __shfl_xor_sync(0xffffffffu, ga, o);  // all lanes: warp stays converged
            // Rotation parameters by ONE lane per pair, broadcast.  The scalar chain is the critical path of a round, so it is
            // written with two reciprocal square roots and no division / sqrt:
            //   rho = sqrt(tau^2 + 4 ga^2), cos 2x = |tau| / rho, cos^2 x = (1 + cos 2x) / 2 =: u, z = rsqrt(u)
            //   c = cos x = u z,  1 / c = z,  |s| = sin 2x / (2 c) = |ga| z / rho,  t = tan x = s z
            // (the small rotation angle |x| <= pi/4, sign as in t = sign(tau) 2 ga / (|tau| + rho)); c^2 + s^2 = u z^2 = 1.
            double c = 1.0, s = 0.0;
            if (act && hl == 0) {
                // rotate unless the columns are orthogonal to the requested level (jtol >= l * eps ~ 2e-14: a tighter
                // threshold is below the rounding noise of the dot product and never lets a sweep come out clean)
                const double tl = (p >= ktop && q >= ktop) ? jtol_bulk : jtol;
                if (ga * ga > tl * tl * al * be && fabs(ga) > 1e-300) {
                    const double tau = be - al;
                    const double yy = rsqrt(tau * tau + 4.0 * ga * ga);
                    const double u = 0.5 + 0.5 * fabs(tau) * yy;
                    const double z = rsqrt(u);
                    c = u * z;
                    const double sa = fabs(ga) * yy * z;
                    s = ((tau < 0.0) != (ga < 0.0)) ? -sa : sa;
                    const double t = s * z;
                    nrm2[p] = al - t * ga; nrm2[q] = be + t * ga; s_rot = 1;
                }
            }
            c = __shfl_sync(0xffffffffu, c, lane & 16);
            s = __shfl_sync(0xffffffffu, s, lane & 16);
            if (act && s != 0.0) {
                double* vp = V + p * cs; double* vq = V + q * cs;
                double vx[NT], vy[NT];
#pragma unroll
                for (int t = 0; t < NT; t++) {
                    const int i = hl + 16 * t;
                    vx[t] = i < lp ? vp[i] : 0.0;
                    vy[t] = i < lp ? vq[i] : 0.0;
                }
#pragma unroll
                for (int t = 0; t < NT; t++) {
                    const int i = hl + 16 * t;
                    if (i < lp) {
                        ap[i] = c * x[t] - s * y[t]; aq[i] = s * x[t] + c * y[t];
                        vp[i] = c * vx[t] - s * vy[t]; vq[i] = s * vx[t] + c * vy[t];
                    }
                }
            }
            __syncthreads();
        }
        const int rot = s_rot;
        __syncthreads();
        if (!rot) break;
    }
    // column norms = eigenvalues (H is PSD); rank them descending
    for (int j = warp; j < lp; j += nwarp) {
        double a = 0.0;
        for (int i = lane; i < lp; i += 32) a += A[j * cs + i] * A[j * cs + i];
        a = warp_sum(a);
        // NaN-safe: the ranking below indexes shared memory with its result, a NaN norm (poisoned input) must still get a rank
        if (lane == 0) s_norm[j] = (j < l) ? ((a == a) ? sqrt(a) : 0.0) : -1.0;
    }
    __syncthreads();
    for (int j = tid; j < lp; j += nt) {
        int rank = 0;
        for (int k = 0; k < lp; k++) rank += (s_norm[k] > s_norm[j]) || (s_norm[k] == s_norm[j] && k < j);
        s_order[rank] = j;
    }
    __syncthreads();
    for (int idx = tid; idx < l * l; idx += nt) {
        const int i = idx / l, jj = idx % l;
        Q[idx] = V[s_order[jj] * cs + i];
    }
    for (int j = tid; j < l; j += nt) theta[j] = s_norm[s_order[j]];
    if (tid == 0 && sweeps_out) *sweeps_out = sweep;
}

inline int jacobi_launch(scm_ctx* ctx, const double* H, int l, double* Q, double* theta, int max_sweeps, double jtol, int ktop,
                         double jtol_bulk, int* sweeps_out, const int* skip) {
    if (l > SCM_MAX_BLOCK) return fail(SCM_ERR_UNSUPPORTED, "jacobi: l=%d > %d", l, SCM_MAX_BLOCK);
    const int lp = (l + 1) & ~1;
    const size_t smem = sizeof(double) * 2 * lp * (lp + 1) + sizeof(unsigned) * (size_t)(lp - 1) * (lp / 2) + 16;
    // one half-warp per column pair: l / 2 pairs need 8 l threads; idle warps would only lengthen every barrier
    int threads = ((lp / 2) * 16 + 31) / 32 * 32;
    if (threads < 256) threads = 256;
    cudaStream_t st = ctx->stream;
#define JAC_CASE(N)                                                                                                                 \
    case N:                                                                                                                         \
        SCM_CUDA(cudaFuncSetAttribute(jacobi_eig_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));               \
        jacobi_eig_kernel<N><<<1, threads, smem, st>>>(H, l, Q, theta, max_sweeps, jtol, ktop, jtol_bulk, sweeps_out, skip);         \
        break;
    switch ((lp + 15) / 16) {
        JAC_CASE(1) JAC_CASE(2) JAC_CASE(3) JAC_CASE(4) JAC_CASE(5) JAC_CASE(6)
        default:
            SCM_CUDA(cudaFuncSetAttribute(jacobi_eig_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
            jacobi_eig_kernel<7><<<1, threads, smem, st>>>(H, l, Q, theta, max_sweeps, jtol, ktop, jtol_bulk, sweeps_out, skip);
    }
#undef JAC_CASE
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

// ---- misc column kernels -----------------------------------------------------------------------------
// res[j] = || W[:, j] - theta[j] V[:, j] ||  (n x l row-major, ld = l); one CTA per column
__global__ void __launch_bounds__(256)
resid_kernel(const double* __restrict__ W, const double* __restrict__ V, const double* __restrict__ theta, int64_t n,
             int l, double* __restrict__ res, const int* __restrict__ skip) {
    if (eig_skip(skip, 0)) return;
    __shared__ double red[8];
    const int j = blockIdx.x;
    const double th = theta[j];
    double a = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { const double d = W[i * l + j] - th * V[i * l + j]; a += d * d; }
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0; for (int w = 0; w < 8; w++) t += red[w]; res[j] = sqrt(t); }
}

// dst[:, j] = src[:, j] / max(theta[j], tiny * theta[0])   (dst may alias src)
__global__ void colscale_kernel(double* dst, const double* src, const double* __restrict__ theta, int64_t n, int l,
                                const int* __restrict__ skip, int need) {
    if (eig_skip(skip, need)) return;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * l) return;
    const int j = (int)(idx % l);
    const double th = fmax(theta[j], 1e-14 * theta[0] + 1e-300);
    dst[idx] = src[idx] / th;
}

// The stopping rule, evaluated by one thread after every Rayleigh-Ritz step (see the header comment).
//   EXACT: Davis-Kahan estimate of the angle between span(V[:, :kq]) and the true invariant subspace, ||R_kq||_F /
//   (theta_kq - theta_kq+1), for the requested rank and every additional one (scRUVIII with several ruvK re-uses one fit); a
//   residual at rounding level also counts as converged.  Hopeless case (theta_k ~ theta_k+1: the k-th factor is not separated,
//   the reference's own answer is ill-conditioned): stop when 15 more iterations gained less than a factor 2.
//   `nonfinite` (optional): the NA / Inf flag of the fit's first pass over the cells -- a poisoned Gram never converges, so the
//   loop is cut at once and the host reports the reference's stop() (R/fastRUVIII.R:52-60).
__global__ void eig_check_kernel(EigCtl* __restrict__ ctl, const double* __restrict__ res, const double* __restrict__ theta, int l, int s,
                                 EigRanks ranks, double tol, int maxit, int fixed_iters, const int* __restrict__ jac_sweeps,
                                 const int* __restrict__ chol_flag, const int64_t* __restrict__ nonfinite) {
    if (threadIdx.x != 0 || ctl->done) return;
    const int it = ++ctl->it;
    ctl->products += 1 + (ctl->pw > 1 ? ctl->pw - 1 : 0);  // this step's product + the extra ones of the previous step
    ctl->jac_sweeps = jac_sweeps ? *jac_sweeps : 0;
    ctl->chol_flag = chol_flag ? *chol_flag : 0;
    if (nonfinite && nonfinite[0] != 0) { ctl->done = 3; return; }
    const double top = theta[0] > 0 ? theta[0] : 1.0;
    double worst = 0.0, gap = 0.0, rmax = 0.0;
    auto estimate = [&](int kq, double& g_out, double& rm_out) -> double {
        double r2 = 0.0, rm = 0.0;
        for (int i = 0; i < kq; i++) { r2 += res[i] * res[i]; rm = rm > res[i] ? rm : res[i]; }
        double g = (kq < l) ? theta[kq - 1] - theta[kq] : theta[kq - 1];
        if (!(g > 1e-300)) g = 1e-300;
        g_out = g; rm_out = rm;
        return rm <= 1e-14 * top ? 0.0 : sqrt(r2) / g;
    };
    for (int q = 0; q < ranks.n; q++) {
        const int kq = ranks.k[q] < s ? ranks.k[q] : s;
        if (kq <= 0) continue;
        double g, rm;
        const double est = estimate(kq, g, rm);
        if (q == 0 || !(est <= worst)) { worst = est; gap = g; }
        rmax = rmax > rm ? rmax : rm;
    }
    ctl->worst = worst; ctl->gap = gap; ctl->rmax = rmax; ctl->top = top;
    if (!(worst == worst) || !(top == top)) ctl->done = 3;  // NaN
    else if (fixed_iters) { if (it >= fixed_iters) ctl->done = 1; }
    else if (worst <= tol) ctl->done = 1;
    else {
        ctl->hist[it % 16] = worst;
        if (it > 40 && worst > 0.5 * ctl->hist[(it + 1) % 16]) ctl->done = 2;
        else if (it >= maxit) ctl->done = 4;
        else if (it >= 8) {
            // rate over the last four steps: stop when, at that rate, the tolerance is out of reach within maxit (flat noise bulk:
            // k exceeds the number of separated factors; 60 Rayleigh-Ritz steps of stagnation cost 66 ms at l = 80)
            const double r4 = worst / ctl->hist[(it - 4) % 16];
            if (!(r4 < 1.0)) { if (it >= 12) ctl->done = 2; }
            else if (4.0 * log(tol / worst) / log(r4) > (double)(maxit - it)) ctl->done = 2;
        }
    }
    if (ctl->done) {
        // how far down the returned rows are trustworthy as a BASIS: the largest r whose top-r subspace passes a 1e-8 angle test
        // (fullalpha re-used with a larger ruvK than it was fitted for, R/fastRUVIII.R:83 / R/scMerge2.R:399)
        int cr = 0;
        for (int r = 1; r <= s; r++) {
            double g, rm;
            if (estimate(r, g, rm) <= (tol > 1e-8 ? tol : 1e-8)) cr = r;
        }
        ctl->conv_rank = cr;
        return;
    }
    // Extra products of the next outer iteration: W <- G (W Theta^-1) costs one skinny GEMM (~40 us) while a Rayleigh-Ritz
    // step costs ~1 ms, so the convergence factor per outer iteration becomes rho^pw.  pw is limited by the spread of the Ritz
    // values inside the block: the scaled columns stay well conditioned for CholeskyQR2 as long as (theta_1 / theta_l)^(pw-1) <= ~1e5.
    const double spread = (theta[0] > 0 && theta[l - 1] > 0) ? theta[0] / theta[l - 1] : 1e300;
    int pw = spread > 1.0 ? (int)(log(1e5) / log(spread)) + 1 : 4;
    ctl->pw = fixed_iters ? 1 : (pw < 1 ? 1 : (pw > 4 ? 4 : pw));
}

// Arms the control block before the first product: a non-finite input (the NA / Inf flag of the fit's pass over the cells) ends
// the solve before any kernel touches the poisoned Gram.
__global__ void eig_prime_kernel(EigCtl* __restrict__ ctl, const int64_t* __restrict__ nonfinite) {
    if (threadIdx.x == 0 && nonfinite && nonfinite[0] != 0) ctl->done = 3;
}

// vecs[r, :] = sign * V[:, r] with the largest-|.| entry made positive; vals[r] = theta[r].  One CTA per row r.
__global__ void __launch_bounds__(256)
eig_output_kernel(const double* __restrict__ V, const double* __restrict__ theta, int64_t n, int l,
                  double* __restrict__ vecs, double* __restrict__ vals) {
    __shared__ double s_best[256];
    __shared__ double s_val[256];
    const int r = blockIdx.x;
    double best = -1.0, bval = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = V[i * l + r];
        if (fabs(v) > best) { best = fabs(v); bval = v; }  // first maximum wins within a thread (ascending i)
    }
    s_best[threadIdx.x] = best; s_val[threadIdx.x] = bval;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int t = 1; t < 256; t++) if (s_best[t] > s_best[0]) { s_best[0] = s_best[t]; s_val[0] = s_val[t]; }
    }
    __syncthreads();
    const double sg = s_val[0] < 0.0 ? -1.0 : 1.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) vecs[(int64_t)r * n + i] = sg * V[i * l + r];
    if (threadIdx.x == 0) vals[r] = theta[r];
}

// dst <- orthonormal basis of span(src) by CholeskyQR2 (n x l row-major; dst may be src).  `Lfac` receives the Cholesky
// factors (l x l).  Per pass: Gram (DMMA GEMM, split over n), register-tiled Cholesky, row solve (in place on dst for the second
// pass).  `pass2` (device, optional: &ctl->skip2[0]): the first factorization decides on the device whether the second pass runs.
inline int cholqr2(scm_ctx* ctx, const double* src, double* dst, int64_t n, int l, double* S, double* Lfac, int* flag,
                   const int* skip = nullptr, int* pass2 = nullptr, int* pass1_count = nullptr) {
    SCM_TRY(gemm(ctx, l, l, (int)n, 1.0, transposed(src, l), rowmajor(src, l), 0.0, S, l, skip));
    SCM_TRY(chol_factor_launch(ctx, S, l, 1e-14, Lfac, flag, skip, pass2, pass1_count));
    SCM_TRY(trsm_rows_launch(ctx, src, dst, n, l, Lfac, skip));
    const int* skip_b = pass2 ? pass2 : skip;
    SCM_TRY(gemm(ctx, l, l, (int)n, 1.0, transposed(dst, l), rowmajor(dst, l), 0.0, S, l, skip_b));
    SCM_TRY(chol_factor_launch(ctx, S, l, 0.0, Lfac, flag, skip_b));
    SCM_TRY(trsm_rows_launch(ctx, dst, dst, n, l, Lfac, skip_b));
    return 0;
}

}  // namespace scm
#include "eig_full.cuh"  // whole-spectrum solver (uses chol_factor_launch / trsm_rows_launch / gemm above)
namespace scm {

// d_nonfinite (optional, device int64): see eig_check_kernel.  Results: ctx->fit_iters / fit_resid / fit_conv_rank / fit_products.
inline int eig_topk(scm_ctx* ctx, const double* d_gram, int64_t n, int32_t s, int32_t kconv, int mode, double tol,
                    int32_t maxit, uint64_t seed, double* d_vals, double* d_vecs, int32_t* iters_out, double* resid_out,
                    const int64_t* d_nonfinite = nullptr) {
    if (n <= 0 || s <= 0 || s > n) return fail(SCM_ERR_INVALID, "eig_topk: bad n=%lld s=%d", (long long)n, s);
    if (n >= (int64_t)1 << 30) return fail(SCM_ERR_UNSUPPORTED, "eig_topk: n too large");
    if (s > SCM_MAX_TOPK_ROWS)  // more rows than a subspace block carries: the whole spectrum (eig_full.cuh)
        return eig_full(ctx, d_gram, n, s, d_vals, d_vecs, iters_out, resid_out, d_nonfinite);
    if (kconv <= 0 || kconv > s) kconv = s;
    int kmax = kconv;  // the largest rank that has to CONVERGE (scRUVIII with several ruvK lists more)
    for (int q = 0; q < ctx->n_conv_ranks; q++) { const int kq = ctx->conv_ranks[q] < s ? ctx->conv_ranks[q] : s; kmax = kmax > kq ? kmax : kq; }
    int l;
    if (mode == SCM_MODE_RANDOMIZED) l = s + 10;  // rsvd oversampling p = 10
    else {
        // Oversampling serves the ranks that must converge (rate lambda_{l+1} / lambda_kmax), not the s rows that are returned:
        // the block is kmax + max(kmax / 2, 16) columns, at least s + 8.  For s = 50, k = 20 that is 64 columns instead of the
        // 80 of round 1: every l x l kernel of the iteration (Jacobi above all) shrinks with it.
        const int over = kmax + (kmax / 2 > 16 ? kmax / 2 : 16);
        l = over > s + 8 ? over : s + 8;
    }
    l = (l + 7) & ~7;
    if (l > n) l = (int)n;
    if (l > SCM_MAX_BLOCK) {
        if (s + 8 <= SCM_MAX_BLOCK) l = SCM_MAX_BLOCK;
        else return fail(SCM_ERR_UNSUPPORTED, "eig_topk: s=%d needs a subspace block > %d (not implemented on the device path)", s, SCM_MAX_BLOCK);
    }
    if (tol <= 0) tol = 1e-10;
    if (maxit <= 0) maxit = 300;
    cudaStream_t st = ctx->stream;
    Scratch ws(ctx);
    double *X, *Z, *W, *V, *H, *Q, *theta, *S, *Linv, *res;
    int* flag;
    EigCtl* ctl;
    const int64_t nl = n * l;
    SCM_TRY(ws.alloc(&X, nl)); SCM_TRY(ws.alloc(&Z, nl)); SCM_TRY(ws.alloc(&W, nl)); SCM_TRY(ws.alloc(&V, nl));
    SCM_TRY(ws.alloc(&H, (int64_t)l * l)); SCM_TRY(ws.alloc(&Q, (int64_t)l * l)); SCM_TRY(ws.alloc(&S, (int64_t)l * l));
    SCM_TRY(ws.alloc(&Linv, (int64_t)l * l)); SCM_TRY(ws.alloc(&theta, l)); SCM_TRY(ws.alloc(&res, l));
    SCM_TRY(ws.alloc(&flag, 4)); SCM_TRY(ws.alloc(&ctl, 1));
    EigRanks ranks;
    ranks.n = 1; ranks.k[0] = kconv;
    for (int q = 0; q < ctx->n_conv_ranks && ranks.n < 17; q++) {
        const int kq = ctx->conv_ranks[q] < s ? ctx->conv_ranks[q] : s;
        if (kq > 0 && kq != kconv) ranks.k[ranks.n++] = kq;
    }
    const int fixed_iters = (mode == SCM_MODE_RANDOMIZED) ? 1 : 0;  // rsvd: the three products are done before the single Rayleigh-Ritz step
    const int* skip = reinterpret_cast<const int*>(ctl);
    const int ktop = kmax + 8 < l ? kmax + 8 : l;  // leading columns of the experimental bulk tolerance (see jacobi_eig_kernel)
    EigCtl* h = reinterpret_cast<EigCtl*>(ctx->h_pinned);
    static_assert(sizeof(EigCtl) <= sizeof(double) * 4096, "control block must fit the pinned staging area");
    const bool debug = getenv("SCM_DEBUG_EIG") != nullptr;
    // Jacobi tolerances of the first (cold) Rayleigh-Ritz step and of bulk-bulk pairs in the warm ones (see jacobi_eig_kernel)
    double jtol_cold = 1e-6, jtol_bulk = 0.0;  // cold: 1e-4 .. 1e-2 cost a third Rayleigh-Ritz step (3.3 ms instead of 2.5)
    if (const char* e = getenv("SCM_EIG_JTOL_COLD")) jtol_cold = atof(e);
    if (const char* e = getenv("SCM_EIG_JTOL_BULK")) jtol_bulk = atof(e);

    // `safe`: no product before the first Rayleigh-Ritz step.  G X for a random X has the condition number of the block's
    // spectrum (theta_1 / theta_l); CholeskyQR2 takes ~1e7 of it.  Beyond that the first attempt reports a failed pivot and the
    // solve is repeated the careful way (Ritz scaling before every orthonormalisation).
    for (int safe = 0; safe < 2; safe++) {
        SCM_TRY(zero_async(ctx, flag, sizeof(int) * 4));
        SCM_TRY(zero_async(ctx, ctl, sizeof(EigCtl)));
        if (d_nonfinite) {
            eig_prime_kernel<<<1, 32, 0, st>>>(ctl, d_nonfinite);
            SCM_LAUNCH_CHECK(ctx);
        }
        randn_kernel<<<(unsigned)ceil_div(nl, 256), 256, 0, st>>>(X, nl, seed * 0x9E3779B97F4A7C15ull + 12345);
        SCM_LAUNCH_CHECK(ctx);
        SCM_TRY(cholqr2(ctx, X, X, n, l, S, Linv, flag, nullptr, ctl->skip2, &ctl->pass1_only));
        // products before the first Rayleigh-Ritz step: a product + CholeskyQR2 costs ~0.2 ms at n = 2000, l = 64, a Rayleigh-Ritz
        // step ~0.65 ms (Jacobi 0.45).  With two of them the 1M x 2000 bench problem converges at the SECOND step (one gave three
        // steps with the 64-column block; the 80-column block of round 1 needed two steps but pays 25 % more per l x l kernel).
        int pre = safe ? 0 : 2;
        if (mode == SCM_MODE_RANDOMIZED && safe) pre = 2;  // rsvd's semantics are three products, whatever the conditioning
        for (int j = 0; j < pre; j++) {
            SCM_TRY(gemm(ctx, (int)n, l, (int)n, 1.0, rowmajor(d_gram, n), rowmajor(X, l), 0.0, Z, l, skip));
            SCM_TRY(cholqr2(ctx, Z, X, n, l, S, Linv, flag, skip, ctl->skip2, &ctl->pass1_only));
        }
        int enq = 0;       // Rayleigh-Ritz steps enqueued so far
        bool finished = false;
        while (!finished) {
            const int batch = fixed_iters ? 1 : 2;
            for (int b = 0; b < batch && enq < maxit; b++, enq++) {
                SCM_TRY(gemm(ctx, (int)n, l, (int)n, 1.0, rowmajor(d_gram, n), rowmajor(X, l), 0.0, Z, l, skip));
                SCM_TRY(gemm(ctx, l, l, (int)n, 1.0, transposed(X, l), rowmajor(Z, l), 0.0, H, l, skip));
                // The first Rayleigh-Ritz problem only has to rotate the basis roughly and to provide the Ritz scaling: a loose
                // Jacobi tolerance saves about half of its sweeps.  Later ones are solved to rounding level, otherwise the
                // residual test measures the Jacobi tolerance instead of the subspace error (the test itself is rigorous for any
                // orthonormal Q).
                const double jtol = (!fixed_iters && enq == 0) ? jtol_cold : 2e-14;
                // a looser tolerance for bulk-bulk pairs of warm-started steps is available for experiments only (see the kernel)
                const bool warm = !fixed_iters && enq > 0 && jtol_bulk > 2e-14;
                SCM_TRY(jacobi_launch(ctx, H, l, Q, theta, 40, jtol, warm ? ktop : l, warm ? jtol_bulk : jtol, flag + 1, skip));
                SCM_TRY(gemm(ctx, (int)n, l, l, 1.0, rowmajor(Z, l), rowmajor(Q, l), 0.0, W, l, skip));
                SCM_TRY(gemm(ctx, (int)n, l, l, 1.0, rowmajor(X, l), rowmajor(Q, l), 0.0, V, l, skip));
                resid_kernel<<<l, 256, 0, st>>>(W, V, theta, n, l, res, skip);
                SCM_LAUNCH_CHECK(ctx);
                eig_check_kernel<<<1, 32, 0, st>>>(ctl, res, theta, l, s, ranks, tol, maxit, fixed_iters, flag + 1, flag, d_nonfinite);
                SCM_LAUNCH_CHECK(ctx);
                if (fixed_iters) break;
                colscale_kernel<<<(unsigned)ceil_div(nl, 256), 256, 0, st>>>(W, W, theta, n, l, skip, 0);
                SCM_LAUNCH_CHECK(ctx);
                for (int j = 2; j <= 4; j++) {  // extra products, each skipped on the device unless ctl->pw >= j
                    SCM_TRY(gemm(ctx, (int)n, l, (int)n, 1.0, rowmajor(d_gram, n), rowmajor(W, l), 0.0, Z, l, skip, j));
                    colscale_kernel<<<(unsigned)ceil_div(nl, 256), 256, 0, st>>>(W, Z, theta, n, l, skip, j);
                    SCM_LAUNCH_CHECK(ctx);
                }
                SCM_TRY(cholqr2(ctx, W, X, n, l, S, Linv, flag, skip, ctl->skip2, &ctl->pass1_only));
            }
            SCM_CUDA(cudaMemcpyAsync(h, ctl, sizeof(EigCtl), cudaMemcpyDeviceToHost, st));
            SCM_CUDA(cudaStreamSynchronize(st));
            if (debug)
                fprintf(stderr, "[eig] safe %d l %d it %d done %d pw %d products %d jacobi sweeps %d chol_flag %d single-pass QR %d angle_est %.3e rmax/top %.3e gap/top %.3e\n",
                        safe, l, h->it, h->done, h->pw, h->products, h->jac_sweeps, h->chol_flag, h->pass1_only, h->worst, h->rmax / h->top, h->gap / h->top);
            finished = h->done != 0 || enq >= maxit;
        }
        if (h->done == 1 || h->done == 3 || safe == 1 || !(h->chol_flag)) break;
    }
    const int done = h->done;
    const double worst = h->worst;
    const int it = h->it;
    ctx->fit_iters = it; ctx->fit_resid = worst; ctx->fit_conv_rank = h->conv_rank; ctx->fit_products = h->products;  // excluding the products before the first Rayleigh-Ritz step
    if (iters_out) *iters_out = it;
    if (resid_out) *resid_out = worst;
    if (done == 3) return fail(SCM_ERR_NONFINITE, "Y contains missing values or infinities.  This is not supported.");
    eig_output_kernel<<<s, 256, 0, st>>>(V, theta, n, l, d_vecs, d_vals);
    SCM_LAUNCH_CHECK(ctx);
    if (done != 1)
        return fail(SCM_ERR_NOCONV, "eig_topk: subspace angle estimate %.3e > tol %.3e after %d iterations "
                    "(factor %d is not separated from factor %d)", worst, tol, it, kconv, kconv + 1);
    return 0;
}

}  // namespace scm
