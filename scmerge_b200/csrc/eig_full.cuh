// K4': the WHOLE spectrum of the n x n Gram, for fits that must hand back more rows of fullalpha than the subspace iteration of
// eig_topk.cuh can carry (s > SCM_MAX_TOPK_ROWS).
//
// Where it is needed.  With a randomized / IRLBA BSPARAM the reference asks runSVD for min(m - ncol(M), sum(ctl)) singular
// vectors (R/fastRUVIII.R:66-70, R/scMerge2_utilis.R:60-65 -- the user's svd_k is ignored there), typically hundreds:
// scMerge2's default reaches this on every call.  rsvd at (nearly) full rank is the exact SVD, and the rows beyond the top few
// dozen sit in the noise bulk where no subspace iteration converges, so the device equivalent is a direct solver.
//
// Method: one-sided block Jacobi on the rows of the symmetric positive semi-definite G = V diag(lambda) V'.  Orthogonal row
// mixing Q G converges to diag(lambda) V': the rows become mutually orthogonal, row i = lambda_i v_i', so an eigenvalue is a row
// norm and an eigenvector the normalised row -- no accumulation of Q.  Rows are grouped in blocks of RB; one tournament step
// pairs all blocks (circle method, nb - 1 steps per sweep), one CTA per block pair:
//   1. the 2 RB rows are read once into REGISTERS (coalesced, element e of a thread is column e * THREADS + tid),
//   2. their (2 RB) x (2 RB) Gram is formed (per-thread partial sums, a butterfly transpose-reduction per warp -- 16 shuffles
//      for 16 sums instead of 80 -- and one pass over the warps' partials in shared memory),
//   3. warp 0 runs ONE parallel-ordered two-sided Jacobi sweep on that small Gram (4 disjoint rotations per step, 7 steps;
//      a complete inner solve costs more per step and saves no outer sweep -- measured),
//   4. every thread applies the small orthogonal factor to its columns in registers and writes the rows back.
// A step is one kernel launch (the kernel boundary is the global barrier); the matrix lives in L2 (32 MB at n = 2000), so a step
// costs about one L2 read + write of it.  Steps whose small Gram is already diagonal to the tolerance write nothing.  The host
// reads one word per sweep (largest |cos| seen between two rows) and stops at 1e-13 or when a sweep no longer improves it.
//
// Start matrix: the Cholesky factor R of G (G = R'R, chol_start_rows below) when G is numerically positive definite -- its rows
// end as sqrt(lambda_i) v_i' (the rows of fullalpha), need fewer sweeps and see the spectrum of G; otherwise G itself, whose row
// Gram is G^2: there the relative gap information below ~1e-8 lambda_1 is lost (sigma_i / sigma_1 < 1e-4 of Y0), the limit the
// n x n Gram formulation has everywhere in this library.  Rows whose norm is below 1e-12 of the largest are treated as the null
// space (a P-row pseudobulk matrix has rank <= P - G) and left alone; the rows a fit returns (s <= m - G) are never among them
// unless Y0 itself is rank deficient.
#pragma once
// included by eig_topk.cuh after its Cholesky / triangular-solve / GEMM helpers (do not include on its own)
#include "common.cuh"
#include "gemm_f64.cuh"

namespace scm {

constexpr int FULL_EIG_MAX_N = SCM_MAX_FULL_EIG_N;

struct FullEigStat {
    unsigned long long maxcos_bits;  // bits of the largest |cos| between two live rows seen in this sweep (non-negative doubles order as integers)
    unsigned long long maxnorm_bits; // bits of the largest squared row norm seen
    int rotated;                     // CTAs that applied a rotation in this sweep
    int pad;
};

// sum over the warp of NV per-lane values with NV - 1 + log2(32 / NV) shuffles; on return v[0] of lane L holds the warp sum of
// value L % NV
template <int NV>
__device__ __forceinline__ void butterfly_reduce(double (&v)[NV], int lane) {
#pragma unroll
    for (int h = NV / 2; h >= 1; h >>= 1) {
        const bool up = (lane & h) != 0;
#pragma unroll
        for (int i = 0; i < h; i++) {
            const double send = up ? v[i] : v[i + h];
            const double keep = up ? v[i + h] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
#pragma unroll
    for (int h = NV; h < 32; h <<= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], h);
}

__host__ __device__ constexpr int tri_index(int i, int j, int R) { return i * R - i * (i - 1) / 2 + (j - i); }  // i <= j

// circle method: players 0 .. nb-2 on a circle, nb-1 fixed; round `step` (0 .. nb-2), pair k (0 .. nb/2-1)
__device__ __forceinline__ void tournament_pair(int nb, int step, int k, int& a, int& b) {
    const int c = nb - 1;
    if (k == 0) { a = c; b = step % c; }
    else { a = (step + k) % c; b = (step - k + 2 * c) % c; }
}

template <int ROWS, int THREADS, int EPT>
__global__ void __launch_bounds__(THREADS, 1)
full_eig_step_kernel(double* __restrict__ Wm, int N, int64_t ld, int nb, int step, double tol, double floor2,
                     FullEigStat* __restrict__ stat, int inner_sweeps) {
    constexpr int RB = ROWS / 2;
    constexpr int NE = ROWS * (ROWS + 1) / 2;
    constexpr int NG = (NE + 15) / 16;
    constexpr int WARPS = THREADS / 32;
    __shared__ double s_part[WARPS][NG * 16];
    __shared__ double s_C[ROWS][ROWS + 1];
    __shared__ double s_B[ROWS][ROWS + 1];
    __shared__ double s_Q[ROWS][ROWS + 1];
    __shared__ int s_apply;
    __shared__ unsigned char s_pair[ROWS - 1][ROWS / 2][2];  // inner tournament: (p, q) of pair pr in step st
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int ba, bb;
    tournament_pair(nb, step, blockIdx.x, ba, bb);
    int64_t rowoff[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++) rowoff[r] = (int64_t)(r < RB ? ba * RB + r : bb * RB + (r - RB)) * ld;

    // programmatic dependent launch: the next step's CTAs may be placed as soon as every CTA of this one has got here, and wait
    // at griddepcontrol.wait for this grid to finish -- the launch latency between two steps is hidden
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    double x[ROWS][EPT];
#pragma unroll
    for (int r = 0; r < ROWS; r++)
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            const int c = e * THREADS + tid;
            x[r][e] = c < N ? Wm[rowoff[r] + c] : 0.0;
        }

    // small Gram: 16 entries at a time
#pragma unroll
    for (int g = 0; g < NG; g++) {
        double v[16];
#pragma unroll
        for (int q = 0; q < 16; q++) v[q] = 0.0;
#pragma unroll
        for (int i = 0; i < ROWS; i++)
#pragma unroll
            for (int j = i; j < ROWS; j++) {
                const int idx = tri_index(i, j, ROWS);
                if (idx / 16 == g) {
                    double acc = 0.0;
#pragma unroll
                    for (int e = 0; e < EPT; e++) acc = fma(x[i][e], x[j][e], acc);
                    v[idx % 16] = acc;
                }
            }
        butterfly_reduce<16>(v, lane);
        if (lane < 16) s_part[warp][g * 16 + lane] = v[0];
    }
    __syncthreads();
    if (tid < NE) {
        double acc = 0.0;
#pragma unroll
        for (int w = 0; w < WARPS; w++) acc += s_part[w][tid];
        // invert tri_index
        int i = 0, base = 0;
        while (tid >= base + (ROWS - i)) { base += ROWS - i; i++; }
        const int j = i + (tid - base);
        s_C[i][j] = acc;
        s_C[j][i] = acc;
    }
    if (tid < ROWS * ROWS) s_Q[tid / ROWS][tid % ROWS] = (tid / ROWS == tid % ROWS) ? 1.0 : 0.0;
    if (tid < (ROWS - 1) * (ROWS / 2)) {
        int p, q;
        tournament_pair(ROWS, tid / (ROWS / 2), tid % (ROWS / 2), p, q);
        s_pair[tid / (ROWS / 2)][tid % (ROWS / 2)][0] = (unsigned char)p;
        s_pair[tid / (ROWS / 2)][tid % (ROWS / 2)][1] = (unsigned char)q;
    }
    __syncthreads();

    if (warp == 0) {
        // convergence measure of this block pair before it is touched
        double mx = 0.0;
        for (int e = lane; e < NE - ROWS; e += 32) {  // off-diagonal pair e -> (i, j), i < j
            int i = 0, base = 0;
            while (e >= base + (ROWS - 1 - i)) { base += ROWS - 1 - i; i++; }
            const int j = i + 1 + (e - base);
            const double d = s_C[i][i] * s_C[j][j];
            if (s_C[i][i] > floor2 && s_C[j][j] > floor2 && d > 0.0) mx = fmax(mx, fabs(s_C[i][j]) * rsqrt(d));
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const bool work = mx > tol;
        if (lane == 0) {
            atomicMax(&stat->maxcos_bits, (unsigned long long)__double_as_longlong(mx));
            s_apply = work ? 1 : 0;
            if (work) atomicAdd(&stat->rotated, 1);
        }
        if (work) {
            constexpr int PAIRS = ROWS / 2;
            const int pr = lane / ROWS, t = lane % ROWS;  // lane = (pair, index); lanes beyond PAIRS * ROWS idle
            const bool active = pr < PAIRS;
            for (int sweep = 0; sweep < inner_sweeps; sweep++) {
                unsigned any = 0;
                for (int st = 0; st < ROWS - 1; st++) {
                    int p = 0, q = 1;
                    double c = 1.0, s = 0.0;
                    bool rot = false;
                    if (active) {
                        p = s_pair[st][pr][0]; q = s_pair[st][pr][1];
                        const double al = s_C[p][p], be = s_C[q][q], ga = s_C[p][q];
                        const double d = al * be;
                        rot = al > floor2 && be > floor2 && d > 0.0 && ga * ga > 1e-30 * d;
                        if (rot) {
                            // small-angle rotation with two reciprocal square roots and no division (see jacobi_eig_kernel):
                            // rho = sqrt(tau^2 + 4 ga^2), u = cos^2 x = (1 + |tau| / rho) / 2, c = u rsqrt(u), |s| = |ga| rsqrt(u) / rho
                            const double tau = be - al;
                            const double yy = rsqrt(tau * tau + 4.0 * ga * ga);
                            const double u = 0.5 + 0.5 * fabs(tau) * yy;
                            double z = rsqrt(u);
                            z = z * (1.5 - 0.5 * u * z * z);  // one Newton step: c^2 + s^2 = u z^2 must be 1 to the last bit or the row norms (the eigenvalues) drift over ~1e4 rotations
                            c = u * z;
                            const double sa = fabs(ga) * yy * z;
                            s = ((tau < 0.0) != (ga < 0.0)) ? -sa : sa;
                        }
                    }
                    any |= __ballot_sync(0xffffffffu, rot);
                    __syncwarp();
                    // B = C J (columns p, q of row t), Q <- Q J
                    if (active) {
                        const double cp = s_C[t][p], cq = s_C[t][q];
                        s_B[t][p] = c * cp - s * cq;
                        s_B[t][q] = s * cp + c * cq;
                        const double qp = s_Q[t][p], qq = s_Q[t][q];
                        s_Q[t][p] = c * qp - s * qq;
                        s_Q[t][q] = s * qp + c * qq;
                    }
                    __syncwarp();
                    // C = J' B (rows p, q of column t)
                    if (active) {
                        const double bp = s_B[p][t], bq = s_B[q][t];
                        s_C[p][t] = c * bp - s * bq;
                        s_C[q][t] = s * bp + c * bq;
                    }
                    __syncwarp();
                    if (active && rot && t == 0) { s_C[p][q] = 0.0; s_C[q][p] = 0.0; }
                    __syncwarp();
                }
                if (!any) break;
            }
        }
    }
    __syncthreads();
    if (!s_apply) return;
    // rows <- Q' rows, one output row at a time (its 2 RB coefficients in registers), stored as it is formed
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        double qc[ROWS];
#pragma unroll
        for (int c0 = 0; c0 < ROWS; c0++) qc[c0] = s_Q[c0][r];
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            const int c = e * THREADS + tid;
            double acc = 0.0;
#pragma unroll
            for (int c0 = 0; c0 < ROWS; c0++) acc = fma(qc[c0], x[c0][e], acc);
            if (c < N) Wm[rowoff[r] + c] = acc;
        }
    }
}

// squared row norms
__global__ void full_eig_norms_kernel(const double* __restrict__ Wm, int N, int rows, int64_t ld, double* __restrict__ norm2) {
    const int r = blockIdx.x;
    if (r >= rows) return;
    double acc = 0.0;
    for (int c = threadIdx.x; c < N; c += blockDim.x) { const double v = Wm[(int64_t)r * ld + c]; acc = fma(v, v, acc); }
    __shared__ double s_red[32];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s_red[w];
        norm2[r] = t;
    }
}

// order[0 .. rows) = row indices by descending norm (ties: ascending index), single CTA bitonic sort in shared memory
__global__ void full_eig_order_kernel(const double* __restrict__ norm2, int rows, int P2, int* __restrict__ order) {
    extern __shared__ unsigned char s_raw[];
    double* key = reinterpret_cast<double*>(s_raw);
    int* idx = reinterpret_cast<int*>(key + P2);
    for (int i = threadIdx.x; i < P2; i += blockDim.x) {
        const double v = i < rows ? norm2[i] : -1.0;
        key[i] = (v == v) ? v : -1.0;  // NaN sorts last
        idx[i] = i;
    }
    __syncthreads();
    for (int k = 2; k <= P2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P2; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const bool desc = (i & k) == 0;
                    const bool before = key[i] > key[p] || (key[i] == key[p] && idx[i] < idx[p]);  // i belongs before p in the final order
                    if (desc ? !before : before) {
                        const double tk = key[i]; key[i] = key[p]; key[p] = tk;
                        const int ti = idx[i]; idx[i] = idx[p]; idx[p] = ti;
                    }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < rows; i += blockDim.x) order[i] = idx[i];
}

// vals[r] = lambda_r, vecs[r, :] = unit eigenvector with its largest-magnitude entry positive (the convention of eig_topk)
__global__ void full_eig_extract_kernel(const double* __restrict__ Wm, int N, int64_t ld, const int* __restrict__ order,
                                        const double* __restrict__ norm2, double* __restrict__ vals, double* __restrict__ vecs, int rows_are_sqrt) {
    const int r = blockIdx.x;
    const int row = order[r];
    const double nrm = sqrt(fmax(norm2[row], 0.0));               // row = lambda v' (start G) or sqrt(lambda) v' (start R, G = R'R)
    const double lam = rows_are_sqrt ? fmax(norm2[row], 0.0) : nrm;
    __shared__ double s_best[256], s_val[256];
    double best = -1.0, bval = 0.0;
    for (int c = threadIdx.x; c < N; c += blockDim.x) {
        const double v = Wm[(int64_t)row * ld + c];
        if (fabs(v) > best) { best = fabs(v); bval = v; }
    }
    s_best[threadIdx.x] = best; s_val[threadIdx.x] = bval;
    __syncthreads();
    if (threadIdx.x == 0)
        for (int t = 1; t < (int)blockDim.x; t++) if (s_best[t] > s_best[0]) { s_best[0] = s_best[t]; s_val[0] = s_val[t]; }
    __syncthreads();
    const double sc = nrm > 0.0 ? (s_val[0] < 0.0 ? -1.0 : 1.0) / nrm : 0.0;
    for (int c = threadIdx.x; c < N; c += blockDim.x) vecs[(int64_t)r * N + c] = sc * Wm[(int64_t)row * ld + c];
    if (threadIdx.x == 0) vals[r] = lam;
}

__global__ void full_eig_copy_kernel(const double* __restrict__ G, int N, int rows_pad, int64_t ld, double* __restrict__ Wm) {
    const int r = blockIdx.y;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ld; c += gridDim.x * blockDim.x)
        Wm[(int64_t)r * ld + c] = (r < N && c < N) ? G[(int64_t)r * N + c] : 0.0;
    (void)rows_pad;
}

// ---- start matrix R (upper triangular, G = R'R) by a blocked right-looking Cholesky ----------------------------------------
// One-sided Jacobi on the rows of R converges in about a third fewer sweeps than on the rows of G (emulated: 10-11 against
// 13-15; measured below) and works with the spectrum of G instead of that of G^2; its rows end as sqrt(lambda_i) v_i', i.e. the
// rows of fullalpha themselves.  Panels of 64 columns: diagonal block by the register-tiled single-CTA factorization of the
// subspace solver, the panel below it by its row solve, the trailing update by the DMMA GEMM.  A Gram that is not numerically
// positive definite (smallest pivot below 1e-12 of the largest diagonal entry: rank-deficient pseudobulk Grams) keeps the
// start matrix G.
__global__ void block_copy_kernel(const double* __restrict__ src, int64_t lds, double* __restrict__ dst, int64_t ldd, int rows, int cols) {
    const int r = blockIdx.y;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += gridDim.x * blockDim.x) dst[(int64_t)r * ldd + c] = src[(int64_t)r * lds + c];
    (void)rows;
}
// Wm[i][j] = L[j][i] for j >= i (rows of R = columns of L), 0 elsewhere and in the padding; pivots[0] = min L[i][i]^2, pivots[1] = max
__global__ void chol_rows_kernel(const double* __restrict__ A, int N, int rows_pad, int64_t ld, double* __restrict__ Wm,
                                 unsigned long long* __restrict__ pivots) {
    __shared__ double tile[32][33];
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;  // output tile rows i0.., cols j0..
    const int tx = threadIdx.x, ty = threadIdx.y;          // 32 x 8
    for (int k = ty; k < 32; k += 8) {
        const int jr = j0 + k, ic = i0 + tx;               // A[jr][ic]
        tile[k][tx] = (jr < N && ic < N && jr >= ic) ? A[(int64_t)jr * N + ic] : 0.0;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int i = i0 + k, j = j0 + tx;
        if (i < rows_pad && j < ld) Wm[(int64_t)i * ld + j] = tile[tx][k];
        if (i < N && i == j) {
            const double d = tile[tx][k] * tile[tx][k];
            atomicMin(&pivots[0], (unsigned long long)__double_as_longlong(d));
            atomicMax(&pivots[1], (unsigned long long)__double_as_longlong(d));
        }
    }
}

// Returns 0 and *used = 1 when Wm holds R; *used = 0 when the Gram is not positive definite enough (Wm untouched).
inline int chol_start_rows(scm_ctx* ctx, const double* d_gram, int N, int rows_pad, int64_t ld, double* Wm, int* used) {
    *used = 0;
    constexpr int NB = 64;
    cudaStream_t st = ctx->stream;
    Scratch ws(ctx);
    double *A, *Sblk, *Lblk, *P;
    int* flag;
    unsigned long long* piv;
    SCM_TRY(ws.alloc(&A, (int64_t)N * N)); SCM_TRY(ws.alloc(&Sblk, NB * NB)); SCM_TRY(ws.alloc(&Lblk, NB * NB));
    SCM_TRY(ws.alloc(&P, (int64_t)N * NB)); SCM_TRY(ws.alloc(&flag, 4)); SCM_TRY(ws.alloc(&piv, 2));
    SCM_TRY(zero_async(ctx, flag, sizeof(int) * 4));
    SCM_TRY(copy_async(ctx, A, d_gram, sizeof(double) * (size_t)N * N));
    for (int k0 = 0; k0 < N; k0 += NB) {
        const int nbk = N - k0 < NB ? N - k0 : NB;
        const int r = N - k0 - nbk;
        double* Akk = A + (int64_t)k0 * N + k0;
        block_copy_kernel<<<dim3(1, (unsigned)nbk), 64, 0, st>>>(Akk, N, Sblk, nbk, nbk, nbk);
        SCM_LAUNCH_CHECK(ctx);
        SCM_TRY(chol_factor_launch(ctx, Sblk, nbk, 0.0, Lblk, flag, nullptr));
        block_copy_kernel<<<dim3(1, (unsigned)nbk), 64, 0, st>>>(Lblk, nbk, Akk, N, nbk, nbk);
        SCM_LAUNCH_CHECK(ctx);
        if (r > 0) {
            double* A21 = A + (int64_t)(k0 + nbk) * N + k0;
            block_copy_kernel<<<dim3(1, (unsigned)r), 64, 0, st>>>(A21, N, P, nbk, r, nbk);
            SCM_LAUNCH_CHECK(ctx);
            SCM_TRY(trsm_rows_launch(ctx, P, P, r, nbk, Lblk, nullptr));
            block_copy_kernel<<<dim3(1, (unsigned)r), 64, 0, st>>>(P, nbk, A21, N, r, nbk);
            SCM_LAUNCH_CHECK(ctx);
            SCM_TRY(gemm(ctx, r, r, nbk, -1.0, rowmajor(P, nbk), transposed(P, nbk), 1.0, A + (int64_t)(k0 + nbk) * N + (k0 + nbk), N));
        }
    }
    unsigned long long* hp = reinterpret_cast<unsigned long long*>(ctx->h_pinned + 1024);
    hp[0] = ~0ull; hp[1] = 0ull;
    SCM_CUDA(cudaMemcpyAsync(piv, hp, sizeof(unsigned long long) * 2, cudaMemcpyHostToDevice, st));
    // into a scratch copy first: Wm keeps the Gram rows if the factor is rejected
    double* R;
    SCM_TRY(ws.alloc(&R, (int64_t)rows_pad * ld));
    chol_rows_kernel<<<dim3((unsigned)ceil_div(ld, 32), (unsigned)ceil_div(rows_pad, 32)), dim3(32, 8), 0, st>>>(A, N, rows_pad, ld, R, piv);
    SCM_LAUNCH_CHECK(ctx);
    int* hf = reinterpret_cast<int*>(ctx->h_pinned + 1024 + 64);
    SCM_CUDA(cudaMemcpyAsync(hp, piv, sizeof(unsigned long long) * 2, cudaMemcpyDeviceToHost, st));
    SCM_CUDA(cudaMemcpyAsync(hf, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    SCM_CUDA(cudaStreamSynchronize(st));
    double pmin, pmax;
    memcpy(&pmin, &hp[0], 8); memcpy(&pmax, &hp[1], 8);
    const bool ok = hf[0] == 0 && pmin == pmin && pmax == pmax && pmax > 0.0 && pmin >= 1e-12 * pmax;
    if (getenv("SCM_DEBUG_EIG")) fprintf(stderr, "[eig_full] Cholesky start: pivots %.3e .. %.3e, flag %d -> %s\n", pmin, pmax, hf[0], ok ? "R" : "G");
    if (!ok) return 0;
    SCM_TRY(copy_async(ctx, Wm, R, sizeof(double) * (size_t)rows_pad * (size_t)ld));
    SCM_CUDA(cudaStreamSynchronize(st));  // R is scratch of this function
    *used = 1;
    return 0;
}

template <int ROWS, int THREADS, int EPT>
inline int full_eig_sweep(scm_ctx* ctx, double* Wm, int N, int64_t ld, int nb, double tol, double floor2, FullEigStat* stat, int inner, bool pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nb / 2));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    for (int step = 0; step < nb - 1; step++) {
        SCM_CUDA(cudaLaunchKernelEx(&cfg, full_eig_step_kernel<ROWS, THREADS, EPT>, Wm, N, ld, nb, step, tol, floor2, stat, inner));
        ctx->launches++;
    }
    return 0;
}

// Top-s eigenpairs of the symmetric positive semi-definite d_gram (n x n, row-major, dense ld = n) from its whole spectrum.
// Same outputs and conventions as eig_topk.  ctx->fit_iters = sweeps, fit_conv_rank = s, fit_resid = last max |cos|.
inline int eig_full(scm_ctx* ctx, const double* d_gram, int64_t n, int32_t s, double* d_vals, double* d_vecs, int32_t* iters_out,
                    double* resid_out, const int64_t* d_nonfinite = nullptr) {
    if (n <= 0 || s <= 0 || s > n) return fail(SCM_ERR_INVALID, "eig_full: bad n=%lld s=%d", (long long)n, s);
    if (n > FULL_EIG_MAX_N)
        return fail(SCM_ERR_UNSUPPORTED, "eig_full: s=%d rows need the whole spectrum, implemented for n <= %d (n=%lld)", s, FULL_EIG_MAX_N, (long long)n);
    const int N = (int)n;
    cudaStream_t st = ctx->stream;
    if (d_nonfinite) {  // a poisoned Gram never converges: leave it to the caller's NA / Inf report
        int64_t* h = reinterpret_cast<int64_t*>(ctx->h_pinned);
        SCM_CUDA(cudaMemcpyAsync(h, d_nonfinite, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        SCM_CUDA(cudaStreamSynchronize(st));
        if (h[0] != 0) return SCM_ERR_NONFINITE;
    }
    // geometry: N <= 2048 -> 8 rows per CTA, 4 elements per thread; above, 4 rows x 8 elements (both 32 doubles of rows per thread)
    int rows_cta, threads;
    if (N <= 512) { rows_cta = 8; threads = 128; }
    else if (N <= 1024) { rows_cta = 8; threads = 256; }
    else if (N <= 2048) { rows_cta = 8; threads = 512; }
    else { rows_cta = 4; threads = 512; }
    const int rb = rows_cta / 2;
    int nb = (N + rb - 1) / rb;
    if (nb & 1) nb++;
    if (nb < 2) nb = 2;
    const int rows_pad = nb * rb;
    const int64_t ld = (N + 1) & ~1;
    Scratch ws(ctx);
    double *Wm, *norm2;
    int* order;
    FullEigStat* stat;
    SCM_TRY(ws.alloc(&Wm, (int64_t)rows_pad * ld));
    SCM_TRY(ws.alloc(&norm2, rows_pad));
    SCM_TRY(ws.alloc(&order, rows_pad));
    SCM_TRY(ws.alloc(&stat, 1));
    {
        dim3 grid((unsigned)ceil_div(ld, 256), (unsigned)rows_pad);
        full_eig_copy_kernel<<<grid, 256, 0, st>>>(d_gram, N, rows_pad, ld, Wm);
        SCM_LAUNCH_CHECK(ctx);
    }
    int rows_are_sqrt = 0;  // 1: the start matrix is the Cholesky factor R (rows end as sqrt(lambda) v'), 0: the Gram itself
    if (!(getenv("SCM_EIGF_CHOL") && atoi(getenv("SCM_EIGF_CHOL")) == 0)) SCM_TRY(chol_start_rows(ctx, d_gram, N, rows_pad, ld, Wm, &rows_are_sqrt));
    // null-space floor: squared row norms below (1e-12)^2 of the largest possible one.  Row norms only redistribute under the
    // rotations: the largest final one (lambda_1^2) is at most the sum and at least the largest of the start matrix's.
    const double tol = 1e-13;
    full_eig_norms_kernel<<<rows_pad, 256, 0, st>>>(Wm, N, rows_pad, ld, norm2);
    SCM_LAUNCH_CHECK(ctx);
    double floor2 = 0.0;
    {
        std::vector<double> hn((size_t)rows_pad);
        SCM_CUDA(cudaMemcpyAsync(hn.data(), norm2, sizeof(double) * rows_pad, cudaMemcpyDeviceToHost, st));
        SCM_CUDA(cudaStreamSynchronize(st));
        double mxn = 0.0;
        for (double v : hn) {
            if (!(v == v) || v > 1e300) return fail(SCM_ERR_NONFINITE, "eig_full: the Gram contains NA / NaN / Inf");
            mxn = v > mxn ? v : mxn;
        }
        floor2 = 1e-24 * mxn;
    }
    FullEigStat* h = reinterpret_cast<FullEigStat*>(ctx->h_pinned);
    const bool debug = getenv("SCM_DEBUG_EIG") != nullptr;
    int inner = 1;  // one inner sweep per block pair: the number of outer sweeps is the same as with a complete inner solve (measured)
    if (const char* e = getenv("SCM_EIGF_INNER")) inner = atoi(e);
    bool pdl = true;
    if (const char* e = getenv("SCM_EIGF_PDL")) pdl = atoi(e) != 0;
    int sweeps = 0;
    double last = 1.0, prev = 2.0;
    int stalled = 0;
    const int max_sweeps = 40;
    for (; sweeps < max_sweeps;) {
        {
            FullEigStat z; z.maxcos_bits = 0; z.maxnorm_bits = 0; z.rotated = 0; z.pad = 0;
            *h = z;
            SCM_CUDA(cudaMemcpyAsync(stat, h, sizeof(FullEigStat), cudaMemcpyHostToDevice, st));
        }
        int rc;
        if (rows_cta == 8 && threads == 128) rc = full_eig_sweep<8, 128, 4>(ctx, Wm, N, ld, nb, tol, floor2, stat, inner, pdl);
        else if (rows_cta == 8 && threads == 256) rc = full_eig_sweep<8, 256, 4>(ctx, Wm, N, ld, nb, tol, floor2, stat, inner, pdl);
        else if (rows_cta == 8) rc = full_eig_sweep<8, 512, 4>(ctx, Wm, N, ld, nb, tol, floor2, stat, inner, pdl);
        else rc = full_eig_sweep<4, 512, 8>(ctx, Wm, N, ld, nb, tol, floor2, stat, inner, pdl);
        if (rc) return rc;
        SCM_CUDA(cudaMemcpyAsync(h, stat, sizeof(FullEigStat), cudaMemcpyDeviceToHost, st));
        SCM_CUDA(cudaStreamSynchronize(st));
        sweeps++;
        prev = last;
        { const unsigned long long bits = h->maxcos_bits; memcpy(&last, &bits, sizeof(double)); }
        if (debug) fprintf(stderr, "[eig_full] n=%d sweep %d max|cos| %.3e rotated CTAs %d of %d\n", N, sweeps, last, h->rotated, (nb / 2) * (nb - 1));
        if (!(last == last)) return fail(SCM_ERR_NONFINITE, "eig_full: the Gram contains NA / NaN / Inf");
        if (last <= tol) break;
        // rounding floor: once the measure stops falling by at least 2x per sweep below 1e-10 the remaining angles are noise
        if (last < 1e-10 && last > 0.5 * prev) { if (++stalled >= 2) break; } else stalled = 0;
    }
    const bool converged = last <= 1e-10;
    full_eig_norms_kernel<<<rows_pad, 256, 0, st>>>(Wm, N, rows_pad, ld, norm2);
    SCM_LAUNCH_CHECK(ctx);
    int P2 = 1;
    while (P2 < rows_pad) P2 <<= 1;
    full_eig_order_kernel<<<1, 1024, (size_t)P2 * 12, st>>>(norm2, rows_pad, P2, order);
    SCM_LAUNCH_CHECK(ctx);
    full_eig_extract_kernel<<<s, 256, 0, st>>>(Wm, N, ld, order, norm2, d_vals, d_vecs, rows_are_sqrt);
    SCM_LAUNCH_CHECK(ctx);
    ctx->launches += 4;
    ctx->fit_iters = sweeps;
    ctx->fit_conv_rank = converged ? s : 0;
    ctx->fit_products = 0;
    ctx->fit_resid = last;
    if (iters_out) *iters_out = sweeps;
    if (resid_out) *resid_out = last;
    if (!converged) {
        fail(SCM_ERR_NOCONV, "eig_full: %d sweeps, max |cos| between rows %.3e", sweeps, last);
        return SCM_ERR_NOCONV;
    }
    return 0;
}

}  // namespace scm
