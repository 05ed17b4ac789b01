// General FP64 DMMA GEMM for the small / skinny products of the eigensolver (G*X, X'Z, Z*Q, X*Linv').
//   C[M x N] = alpha * A[M x K] * B[K x N] + beta * C,   A and B given as strided views (any transposition).
// 64x64 CTA tile, 4 warps x (32x32), K chunks of 16 staged through shared memory with register prefetch,
// optional split over K with an ordered (deterministic) reduction.
#pragma once
#include "common.cuh"

namespace scm {

struct MatView {
    const double* p;
    int64_t rs, cs;  // element (r, c) = p[r * rs + c * cs]
};
inline MatView rowmajor(const double* p, int64_t ld) { return MatView{p, ld, 1}; }
inline MatView transposed(const double* p, int64_t ld) { return MatView{p, 1, ld}; }  // view of P' for row-major P

constexpr int GM_BT = 64, GM_KC = 16, GM_LDS = GM_BT + 4, GM_THREADS = 128;

__global__ void __launch_bounds__(GM_THREADS)
gemm_kernel(int M, int N, int K, MatView A, MatView B, double* __restrict__ out, int64_t ldo, int64_t split_stride,
            int kchunks_per_split, double alpha, double beta, int direct, const int* __restrict__ skip, int need) {
    if (skip && (skip[0] != 0 || skip[1] < need)) return;  // device-side control flow of the eigensolver (eig_topk.cuh)
    __shared__ double As[2][GM_KC][GM_LDS];
    __shared__ double Bs[2][GM_KC][GM_LDS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    const int i0 = blockIdx.y * GM_BT, j0 = blockIdx.x * GM_BT;
    const int kc_begin = blockIdx.z * kchunks_per_split;
    const int kc_total = (K + GM_KC - 1) / GM_KC;
    const int kc_end = min(kc_total, kc_begin + kchunks_per_split);

    // element -> thread mapping for the global loads: vary the memory-contiguous index fastest
    const bool a_kfast = (A.cs == 1);  // A(i,k) contiguous in k
    const bool b_kfast = (B.rs == 1);  // B(k,j) contiguous in k
    double ra[8], rb[8];
    auto gload = [&](int kc) {
        const int k0 = kc * GM_KC;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int idx = tid + e * GM_THREADS;  // 0..1023
            int ka, ia, kb, jb;
            if (a_kfast) { ka = idx & 15; ia = idx >> 4; } else { ia = idx & 63; ka = idx >> 6; }
            if (b_kfast) { kb = idx & 15; jb = idx >> 4; } else { jb = idx & 63; kb = idx >> 6; }
            const int gi = i0 + ia, gj = j0 + jb;
            ra[e] = (gi < M && k0 + ka < K) ? A.p[(int64_t)gi * A.rs + (int64_t)(k0 + ka) * A.cs] : 0.0;
            rb[e] = (gj < N && k0 + kb < K) ? B.p[(int64_t)(k0 + kb) * B.rs + (int64_t)gj * B.cs] : 0.0;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int idx = tid + e * GM_THREADS;
            int ka, ia, kb, jb;
            if (a_kfast) { ka = idx & 15; ia = idx >> 4; } else { ia = idx & 63; ka = idx >> 6; }
            if (b_kfast) { kb = idx & 15; jb = idx >> 4; } else { jb = idx & 63; kb = idx >> 6; }
            As[buf][ka][ia] = ra[e];
            Bs[buf][kb][jb] = rb[e];
        }
    };

    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;

    if (kc_begin < kc_end) {
        gload(kc_begin);
        sstore(0);
        __syncthreads();
        for (int kc = kc_begin; kc < kc_end; kc++) {
            const int buf = (kc - kc_begin) & 1;
            if (kc + 1 < kc_end) gload(kc + 1);
#pragma unroll
            for (int ks = 0; ks < GM_KC / 4; ks++) {
                double af[4], bf[4];
#pragma unroll
                for (int a = 0; a < 4; a++) af[a] = As[buf][ks * 4 + tig][wm * 32 + a * 8 + gid];
#pragma unroll
                for (int b = 0; b < 4; b++) bf[b] = Bs[buf][ks * 4 + tig][wn * 32 + b * 8 + gid];
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 4; b++) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
            }
            if (kc + 1 < kc_end) sstore(buf ^ 1);
            __syncthreads();
        }
    }
    double* o = out + (int64_t)blockIdx.z * split_stride;
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int i = i0 + wm * 32 + a * 8 + gid, j = j0 + wn * 32 + b * 8 + 2 * tig + e;
                if (i < M && j < N) {
                    double v = acc[a][b][e];
                    if (direct) {
                        v *= alpha;
                        if (beta != 0.0) v += beta * o[(int64_t)i * ldo + j];
                    }
                    o[(int64_t)i * ldo + j] = v;
                }
            }
}

__global__ void gemm_splitk_reduce_kernel(const double* __restrict__ part, int splits, int64_t split_stride, int M, int N,
                                          double alpha, double beta, double* __restrict__ C, int64_t ldc,
                                          const int* __restrict__ skip, int need) {
    if (skip && (skip[0] != 0 || skip[1] < need)) return;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)M * N) return;
    const int i = (int)(idx / N), j = (int)(idx % N);
    double acc = 0.0;
    for (int s = 0; s < splits; s++) acc += part[(int64_t)s * split_stride + idx];
    double v = alpha * acc;
    if (beta != 0.0) v += beta * C[(int64_t)i * ldc + j];
    C[(int64_t)i * ldc + j] = v;
}

// `skip` (optional, device): the launch returns at once when skip[0] != 0 or skip[1] < need (see EigCtl in eig_topk.cuh).
inline int gemm(scm_ctx* ctx, int M, int N, int K, double alpha, MatView A, MatView B, double beta, double* C, int64_t ldc,
                const int* skip = nullptr, int need = 0) {
    if (M <= 0 || N <= 0) return 0;
    cudaStream_t st = ctx->stream;
    const int tiles = (int)(ceil_div(M, GM_BT) * ceil_div(N, GM_BT));
    const int kc_total = (int)ceil_div(K > 0 ? K : 1, GM_KC);
    int splits = 1;
    if (tiles < ctx->num_sms) {
        splits = (2 * ctx->num_sms) / tiles;
        const int max_splits = kc_total / 4 > 0 ? kc_total / 4 : 1;  // at least 64 k per split
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
    }
    const int per = (int)ceil_div(kc_total, splits);
    splits = (int)ceil_div(kc_total, per);
    dim3 grid((unsigned)ceil_div(N, GM_BT), (unsigned)ceil_div(M, GM_BT), (unsigned)splits);
    if (splits == 1) {
        gemm_kernel<<<grid, GM_THREADS, 0, st>>>(M, N, K, A, B, C, ldc, 0, per, alpha, beta, 1, skip, need);
        SCM_LAUNCH_CHECK(ctx);
        return 0;
    }
    Scratch ws(ctx);
    double* part;
    const int64_t stride = (int64_t)M * N;
    SCM_TRY(ws.alloc(&part, stride * splits));
    gemm_kernel<<<grid, GM_THREADS, 0, st>>>(M, N, K, A, B, part, N, stride, per, 1.0, 0.0, 0, skip, need);
    SCM_LAUNCH_CHECK(ctx);
    gemm_splitk_reduce_kernel<<<(unsigned)ceil_div(stride, 256), 256, 0, st>>>(part, splits, stride, M, N, alpha, beta, C, ldc, skip, need);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace scm
