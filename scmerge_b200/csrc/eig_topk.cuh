// K4: top-s eigenpairs of the n x n symmetric PSD Gram by block subspace iteration.
//
//   X <- orth(random n x l)                                   l = s + oversampling (<= SCM_MAX_BLOCK)
//   repeat:  Z = G X ; H = X'Z ; (Q, theta) = eig(H)          Rayleigh-Ritz, single-CTA one-sided Jacobi
//            W = Z Q ; V = X Q ; r_i = ||W_i - theta_i V_i||   true residuals of the Ritz pairs
//            stop when ||R_kconv||_F / (theta_kconv - theta_kconv+1) <= tol (EXACT; Davis-Kahan bound on the
//            subspace angle) or after q+1 = 3 products (RANDOMIZED)
//            X <- CholeskyQR2(W diag(1/theta))                 columns ~ orthonormal already => well conditioned
// All n-sized products run on the DMMA GEMM; the l x l factorizations are single-CTA shared-memory kernels.
#pragma once
#include "common.cuh"
#include "gemm_f64.cuh"

namespace scm {

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
__global__ void __launch_bounds__(1024)
chol_inv_kernel(const double* __restrict__ S, int l, double shift_rel, double* __restrict__ Linv, int* __restrict__ flag) {
    extern __shared__ double sm[];
    const int ld = l + 1;
    double* A = sm;            // l x ld   (becomes L, lower)
    double* X = sm + l * ld;   // l x ld   (L^-1, lower)
    __shared__ double s_maxdiag;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int idx = tid; idx < l * l; idx += nt) {
        const int i = idx / l, j = idx % l;
        A[i * ld + j] = 0.5 * (S[i * l + j] + S[j * l + i]);
        X[i * ld + j] = 0.0;
    }
    __syncthreads();
    if (tid == 0) {
        double md = 0.0;
        for (int i = 0; i < l; i++) md = fmax(md, A[i * ld + i]);
        s_maxdiag = md;
    }
    __syncthreads();
    const double maxdiag = s_maxdiag;
    for (int i = tid; i < l; i += nt) A[i * ld + i] += shift_rel * maxdiag;
    __syncthreads();
    for (int j = 0; j < l; j++) {
        if (tid == 0) {
            double d = A[j * ld + j];
            if (!(d > 1e-280 + 1e-30 * maxdiag)) { d = 1e-30 * maxdiag + 1e-280; if (flag) *flag = 1; }
            A[j * ld + j] = sqrt(d);
        }
        __syncthreads();
        const double djj = A[j * ld + j];
        for (int i = j + 1 + tid; i < l; i += nt) A[i * ld + j] /= djj;
        __syncthreads();
        const int rem = l - j - 1;
        for (int idx = tid; idx < rem * rem; idx += nt) {
            const int i = j + 1 + idx / rem, k = j + 1 + idx % rem;
            if (k <= i) A[i * ld + k] -= A[i * ld + j] * A[k * ld + j];
        }
        __syncthreads();
    }
    // forward substitution, thread c owns column c of L^-1
    for (int i = 0; i < l; i++) {
        if (tid < l && tid <= i) {
            const int c = tid;
            // X[k][c] = 0 for k < c; four independent partial sums break the FMA dependency chain
            double s0 = (i == c) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int k = c;
            for (; k + 3 < i; k += 4) {
                s0 -= A[i * ld + k] * X[k * ld + c];
                s1 -= A[i * ld + k + 1] * X[(k + 1) * ld + c];
                s2 -= A[i * ld + k + 2] * X[(k + 2) * ld + c];
                s3 -= A[i * ld + k + 3] * X[(k + 3) * ld + c];
            }
            for (; k < i; k++) s0 -= A[i * ld + k] * X[k * ld + c];
            X[i * ld + c] = ((s0 + s1) + (s2 + s3)) / A[i * ld + i];
        }
        __syncthreads();
    }
    for (int idx = tid; idx < l * l; idx += nt) {
        const int i = idx / l, j = idx % l;
        Linv[idx] = (j <= i) ? X[i * ld + j] : 0.0;
    }
}

// ---- single-CTA symmetric eigensolver: one-sided (Hestenes) Jacobi ---------------------------------
// H (l x l symmetric PSD, row-major) -> theta (descending), Q (row-major l x l, column j = eigenvector j).
__global__ void __launch_bounds__(1024)
jacobi_eig_kernel(const double* __restrict__ H, int l, double* __restrict__ Q, double* __restrict__ theta,
                  int max_sweeps, double jtol, int* __restrict__ sweeps_out) {
    extern __shared__ double sm[];
    const int lp = (l + 1) & ~1;  // even number of columns (a zero column pads odd l)
    double* A = sm;               // column-major: column j at A + j*lp
    const int cs = lp + 1;        // column stride: odd, so the two half-warps of a warp hit different banks
    double* V = sm + lp * cs;
    __shared__ int s_rot;
    __shared__ int s_order[SCM_MAX_BLOCK + 2];
    __shared__ double s_norm[SCM_MAX_BLOCK + 2];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    for (int idx = tid; idx < lp * lp; idx += nt) {
        const int j = idx / lp, i = idx % lp;
        A[j * cs + i] = (i < l && j < l) ? 0.5 * (H[i * l + j] + H[j * l + i]) : 0.0;
        V[j * cs + i] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    // One half-warp (16 lanes) per column pair: all <= 56 disjoint pairs of a round run concurrently on the 64
    // half-warps.  Squared column norms are cached (recomputed once per sweep, updated by the Rutishauser formulas
    // ||a_p'||^2 = alpha - t*gamma, ||a_q'||^2 = beta + t*gamma), so a pair costs one dot product + the rotation.
    const int npair = lp / 2;
    const int hw = tid >> 4, hl = tid & 15;
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
            const bool act = hw < npair;
            int p = 0, q = 0;
            if (act) {
                if (hw == 0) { p = lp - 1; q = r; }
                else { p = (r + hw) % (lp - 1); q = (r - hw + (lp - 1)) % (lp - 1); }
            }
            double* ap = A + p * cs; double* aq = A + q * cs;
            double ga = 0.0;
            if (act) for (int i = hl; i < lp; i += 16) ga += ap[i] * aq[i];
            const double al = nrm2[p], be = nrm2[q];  // read before the shuffles: they order it against lane 0's update
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) ga += __shfl_xor_sync(0xffffffffu, ga, o);  // all lanes: warp stays converged
            // The rotation parameters (two FP64 divisions, a sqrt and an rsqrt: ~150 FP64 pipe slots) are computed by
            // ONE lane per pair and broadcast; doing it redundantly in all 16 lanes made this scalar math the bulk of a round.
            double c = 1.0, s = 0.0;
            if (act && hl == 0) {
                // rotate unless the columns are orthogonal to the requested level (jtol >= l * eps ~ 2e-14: a tighter
                // threshold is below the rounding noise of the dot product and never lets a sweep come out clean)
                if (ga * ga > jtol * jtol * al * be && fabs(ga) > 1e-300) {
                    // t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = (be - al) / (2 ga), with the inner division folded
                    // away: one sqrt + one division + one rsqrt on the critical path of a round instead of 2 + 2 + 1
                    const double tau = be - al;
                    const double t = (tau < 0.0 ? -2.0 * ga : 2.0 * ga) / (fabs(tau) + sqrt(tau * tau + 4.0 * ga * ga));
                    c = rsqrt(1.0 + t * t);
                    s = c * t;
                    nrm2[p] = al - t * ga; nrm2[q] = be + t * ga; s_rot = 1;
                }
            }
            c = __shfl_sync(0xffffffffu, c, lane & 16);
            s = __shfl_sync(0xffffffffu, s, lane & 16);
            if (act && s != 0.0) {
                double* vp = V + p * cs; double* vq = V + q * cs;
                for (int i = hl; i < lp; i += 16) {
                    const double x = ap[i], y = aq[i];
                    ap[i] = c * x - s * y; aq[i] = s * x + c * y;
                    const double vx = vp[i], vy = vq[i];
                    vp[i] = c * vx - s * vy; vq[i] = s * vx + c * vy;
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
        if (lane == 0) s_norm[j] = (j < l) ? sqrt(a) : -1.0;
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

// ---- misc column kernels -----------------------------------------------------------------------------
// res[j] = || W[:, j] - theta[j] V[:, j] ||  (n x l row-major, ld = l); one CTA per column
__global__ void __launch_bounds__(256)
resid_kernel(const double* __restrict__ W, const double* __restrict__ V, const double* __restrict__ theta, int64_t n,
             int l, double* __restrict__ res) {
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

// W[:, j] *= 1 / max(theta[j], tiny * theta[0])
__global__ void colscale_kernel(double* __restrict__ W, const double* __restrict__ theta, int64_t n, int l) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * l) return;
    const int j = (int)(idx % l);
    const double th = fmax(theta[j], 1e-14 * theta[0] + 1e-300);
    W[idx] /= th;
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

// X <- orthonormal basis of span(X) by CholeskyQR2 (X: n x l row-major, T: n x l scratch).
inline int cholqr2(scm_ctx* ctx, double* X, double* T, int64_t n, int l, double* S, double* Linv, int* flag) {
    cudaStream_t st = ctx->stream;
    const size_t smem = sizeof(double) * 2 * l * (l + 1);
    double* src = X; double* dst = T;
    for (int pass = 0; pass < 2; pass++) {
        SCM_TRY(gemm(ctx, l, l, (int)n, 1.0, transposed(src, l), rowmajor(src, l), 0.0, S, l));
        chol_inv_kernel<<<1, 1024, smem, st>>>(S, l, pass == 0 ? 1e-14 : 0.0, Linv, flag);
        SCM_LAUNCH_CHECK(ctx);
        SCM_TRY(gemm(ctx, (int)n, l, l, 1.0, rowmajor(src, l), transposed(Linv, l), 0.0, dst, l));
        double* t = src; src = dst; dst = t;
    }
    return 0;  // two passes: the result is back in X
}

inline int eig_topk(scm_ctx* ctx, const double* d_gram, int64_t n, int32_t s, int32_t kconv, int mode, double tol,
                    int32_t maxit, uint64_t seed, double* d_vals, double* d_vecs, int32_t* iters_out, double* resid_out) {
    if (n <= 0 || s <= 0 || s > n) return fail(SCM_ERR_INVALID, "eig_topk: bad n=%lld s=%d", (long long)n, s);
    if (n >= (int64_t)1 << 30) return fail(SCM_ERR_UNSUPPORTED, "eig_topk: n too large");
    int l;
    if (mode == SCM_MODE_RANDOMIZED) l = s + 10;  // rsvd oversampling p = 10
    else l = s + (s / 2 > 16 ? s / 2 : 16);
    l = (l + 7) & ~7;
    if (l > n) l = (int)n;
    if (l > SCM_MAX_BLOCK) {
        if (s + 8 <= SCM_MAX_BLOCK) l = SCM_MAX_BLOCK;
        else return fail(SCM_ERR_UNSUPPORTED, "eig_topk: s=%d needs a subspace block > %d (not implemented on the device path)", s, SCM_MAX_BLOCK);
    }
    if (kconv <= 0 || kconv > s) kconv = s;
    if (tol <= 0) tol = 1e-10;
    if (maxit <= 0) maxit = 300;
    cudaStream_t st = ctx->stream;
    static bool attr_set = false;
    if (!attr_set) {
        SCM_CUDA(cudaFuncSetAttribute(chol_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        SCM_CUDA(cudaFuncSetAttribute(jacobi_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_set = true;
    }
    Scratch ws(ctx);
    double *X, *Z, *W, *V, *H, *Q, *theta, *S, *Linv, *res;
    int* flag;
    const int64_t nl = n * l;
    SCM_TRY(ws.alloc(&X, nl)); SCM_TRY(ws.alloc(&Z, nl)); SCM_TRY(ws.alloc(&W, nl)); SCM_TRY(ws.alloc(&V, nl));
    SCM_TRY(ws.alloc(&H, (int64_t)l * l)); SCM_TRY(ws.alloc(&Q, (int64_t)l * l)); SCM_TRY(ws.alloc(&S, (int64_t)l * l));
    SCM_TRY(ws.alloc(&Linv, (int64_t)l * l)); SCM_TRY(ws.alloc(&theta, l)); SCM_TRY(ws.alloc(&res, l));
    SCM_TRY(ws.alloc(&flag, 4));
    SCM_CUDA(cudaMemsetAsync(flag, 0, sizeof(int) * 4, st));
    randn_kernel<<<(unsigned)ceil_div(nl, 256), 256, 0, st>>>(X, nl, seed * 0x9E3779B97F4A7C15ull + 12345);
    SCM_LAUNCH_CHECK(ctx);
    SCM_TRY(cholqr2(ctx, X, Z, n, l, S, Linv, flag));
    const size_t jsmem = sizeof(double) * 2 * ((l + 1) & ~1) * (((l + 1) & ~1) + 1);
    const int fixed_iters = (mode == SCM_MODE_RANDOMIZED) ? 3 : 0;  // q = 2 power iterations + the final product
    double* h = ctx->h_pinned;
    int it = 0;
    double worst = 0.0;
    double hist[16];
    for (int i = 0; i < 16; i++) hist[i] = 1e300;
    bool converged = false;

    for (it = 1; it <= maxit; it++) {
        SCM_TRY(gemm(ctx, (int)n, l, (int)n, 1.0, rowmajor(d_gram, n), rowmajor(X, l), 0.0, Z, l));
        SCM_TRY(gemm(ctx, l, l, (int)n, 1.0, transposed(X, l), rowmajor(Z, l), 0.0, H, l));
        // The first Rayleigh-Ritz problem (random start block) only needs to rotate the basis roughly: a loose Jacobi
        // tolerance saves ~8 of its 13 sweeps.  Later ones are solved to rounding level, otherwise the residual test below
        // measures the Jacobi tolerance instead of the subspace error (it stays rigorous for any orthonormal Q).
        const double jtol = (!fixed_iters && it == 1) ? 1e-6 : 2e-14;

        // one half-warp per column pair: l / 2 pairs need 8 l threads; idle warps would only lengthen every barrier
        const int jthreads = ((((l + 1) / 2) * 16 + 31) / 32 * 32) < 256 ? 256 : ((((l + 1) / 2) * 16 + 31) / 32 * 32);
        jacobi_eig_kernel<<<1, jthreads, jsmem, st>>>(H, l, Q, theta, 40, jtol, flag + 1);
        SCM_LAUNCH_CHECK(ctx);
        SCM_TRY(gemm(ctx, (int)n, l, l, 1.0, rowmajor(Z, l), rowmajor(Q, l), 0.0, W, l));
        SCM_TRY(gemm(ctx, (int)n, l, l, 1.0, rowmajor(X, l), rowmajor(Q, l), 0.0, V, l));
        const bool check = fixed_iters ? (it == fixed_iters) : true;
        if (check) {
            resid_kernel<<<l, 256, 0, st>>>(W, V, theta, n, l, res);
            SCM_LAUNCH_CHECK(ctx);
            SCM_CUDA(cudaMemcpyAsync(h, res, sizeof(double) * l, cudaMemcpyDeviceToHost, st));
            SCM_CUDA(cudaMemcpyAsync(h + l, theta, sizeof(double) * l, cudaMemcpyDeviceToHost, st));
            SCM_CUDA(cudaMemcpyAsync(h + 2 * l, flag, sizeof(int) * 4, cudaMemcpyDeviceToHost, st));
            SCM_CUDA(cudaStreamSynchronize(st));
            if (getenv("SCM_DEBUG_EIG")) fprintf(stderr, "[eig] it %d l %d jacobi sweeps %d\n", it, l, reinterpret_cast<int*>(h + 2 * l)[1]);
            // Davis-Kahan estimate of the angle between span(V[:, :kconv]) and the true invariant subspace:
            // ||R_k||_F / (theta_k - theta_{k+1}); a residual at rounding level also counts as converged.
            const double* th = h + l;
            const double top = th[0] > 0 ? th[0] : 1.0;
            double rmax = 0.0, gap = 0.0;
            worst = 0.0;
            // the requested rank and every additional one (scRUVIII with several ruvK re-uses one fit for all of them)
            for (int q = -1; q < ctx->n_conv_ranks; q++) {
                const int kq = q < 0 ? kconv : (ctx->conv_ranks[q] < s ? ctx->conv_ranks[q] : s);
                if (kq <= 0 || (q >= 0 && kq == kconv)) continue;
                double r2 = 0.0, rm = 0.0;
                for (int i = 0; i < kq; i++) { r2 += h[i] * h[i]; rm = rm > h[i] ? rm : h[i]; }
                double g = (kq < l) ? th[kq - 1] - th[kq] : th[kq - 1];
                if (!(g > 1e-300)) g = 1e-300;
                const double est = rm <= 1e-14 * top ? 0.0 : sqrt(r2) / g;  // a residual at rounding level counts as converged
                if (q < 0 || est > worst) { if (est >= worst) { worst = est; gap = g; } }
                rmax = rmax > rm ? rmax : rm;
            }
            if (getenv("SCM_DEBUG_EIG"))
                fprintf(stderr, "[eig] it %d angle_est %.3e rmax/top %.3e gap/top %.3e theta1 %.3e theta_k %.3e theta_l %.3e\n", it, worst,
                        rmax / top, gap / top, th[0], th[kconv - 1], th[l - 1]);
            if (fixed_iters || worst <= tol) { converged = true; break; }

            hist[it % 16] = worst;
            // hopeless case (theta_k ~ theta_{k+1}: the k-th factor is not separated, the reference's own answer
            // is ill-conditioned): stop when 15 more iterations gained less than a factor 2
            if (it > 40 && worst > 0.5 * hist[(it + 1) % 16]) break;
        }
        if (it == maxit) break;
        colscale_kernel<<<(unsigned)ceil_div(nl, 256), 256, 0, st>>>(W, theta, n, l);
        SCM_LAUNCH_CHECK(ctx);
        if (!fixed_iters) {
            // Extra products per outer iteration: W <- G (W Theta^-1) costs one skinny GEMM (~40 us) while a Rayleigh-Ritz
            // step costs ~1 ms, so the convergence factor per outer iteration becomes rho^p.  p is limited by the spread
            // of the Ritz values inside the block: the scaled columns stay well conditioned for CholeskyQR2 as long as
            // (theta_1 / theta_l)^(p-1) <= ~1e5.
            const double spread = (h[l] > 0 && h[2 * l - 1] > 0) ? h[l] / h[2 * l - 1] : 1e300;
            int pw = spread > 1.0 ? (int)(log(1e5) / log(spread)) + 1 : 4;
            pw = pw < 1 ? 1 : (pw > 4 ? 4 : pw);
            for (int j = 1; j < pw; j++) {
                SCM_TRY(gemm(ctx, (int)n, l, (int)n, 1.0, rowmajor(d_gram, n), rowmajor(W, l), 0.0, Z, l));
                colscale_kernel<<<(unsigned)ceil_div(nl, 256), 256, 0, st>>>(Z, theta, n, l);
                SCM_LAUNCH_CHECK(ctx);
                double* t2 = W; W = Z; Z = t2;
            }
        }
        SCM_TRY(cholqr2(ctx, W, Z, n, l, S, Linv, flag));
        double* t = X; X = W; W = t;
    }
    eig_output_kernel<<<s, 256, 0, st>>>(V, theta, n, l, d_vecs, d_vals);
    SCM_LAUNCH_CHECK(ctx);
    if (iters_out) *iters_out = it > maxit ? maxit : it;
    if (resid_out) *resid_out = worst;
    if (!converged)
        return fail(SCM_ERR_NOCONV, "eig_topk: subspace angle estimate %.3e > tol %.3e after %d iterations "
                    "(factor %d is not separated from factor %d)", worst, tol, it > maxit ? maxit : it, kconv, kconv + 1);
    return 0;
}

}  // namespace scm
