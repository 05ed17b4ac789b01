// k-means with many random starts on the device: `stats::kmeans(pca[, 1:10], centers = kmeansK[j], nstart = 1000)` of
// scReplicate's identifyCluster (R/scReplicate.R:391-395) -- once RUV-III takes 0.13 s this call (1000 sequential Hartigan-Wong
// runs per batch in R) is most of scMerge()'s wall time.
//
// What the reference computes: for every start K distinct rows of x as centres, Hartigan-Wong (AS 136) to a partition that no
// single-point transfer improves, and the start with the smallest total within-cluster sum of squares wins.  R's answer depends on
// its RNG stream; what is reproducible is the OPTIMUM the starts agree on.  Here:
//   km_restart_kernel   one CTA per start (grid-stride over the starts): K distinct rows from a counter-based hash of
//                       (seed, start), Lloyd iterations until no centre changes, objective; every CTA keeps its best start.
//                       Points are read from a transposed copy (coalesced), every thread adds its points into PRIVATE
//                       shared-memory accumulators (no atomics), which are summed in a fixed order: a start's result depends
//                       on (seed, start) only, not on the grid.
//   km_finish_kernel    one CTA: winner over the CTAs (ties: lowest start), labels, then Hartigan-Wong's own stopping rule --
//                       single transfers with the n/(n -+ 1) weights (AS 136, optimal-transfer step) until none reduces the
//                       objective -- so the partition returned is stable in the reference's sense, not only in Lloyd's; exact
//                       centres, sizes and within-cluster sums of squares of the final partition.
// Limits: d <= 16 coordinates (scReplicate uses 10), K (d + 1) <= 800 accumulators.
#pragma once
#include "common.cuh"

namespace scm {

constexpr int KM_DMAX = 16;

__device__ __forceinline__ uint64_t km_mix(uint64_t x) {  // splitmix64 finaliser
    x += 0x9e3779b97f4a7c15ULL;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
}

__global__ void km_transpose_kernel(const double* __restrict__ X, int64_t m, int d, int64_t ld, double* __restrict__ Xt) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < m * d; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / d;
        const int j = (int)(e - i * d);
        Xt[(int64_t)j * m + i] = X[i * ld + j];
    }
}

// index of the nearest centre (ties: the lowest index) and the squared distance to it
template <int DM>
__device__ __forceinline__ int km_nearest(const double (&x)[DM], const double* __restrict__ cen, int K, int d, double& dbest) {
    int cb = 0;
    dbest = INFINITY;
    for (int c = 0; c < K; c++) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < DM; j++)
            if (j < d) { const double t = x[j] - cen[c * d + j]; s = fma(t, t, s); }
        if (s < dbest) { dbest = s; cb = c; }
    }
    return cb;
}

template <int DM>
__device__ __forceinline__ void km_load(const double* __restrict__ Xt, int64_t m, int d, int64_t i, double (&x)[DM]) {
#pragma unroll
    for (int j = 0; j < DM; j++) x[j] = j < d ? Xt[(int64_t)j * m + i] : 0.0;
}

// red[s] = sum over the T threads of acc[s * T + t], s < nslots: a warp per slot, lanes take t = lane, lane + 32, ... in order,
// then a fixed xor tree -- the same association for every launch geometry
template <int T>
__device__ __forceinline__ void km_reduce_slots(const double* __restrict__ acc, int nslots, double* __restrict__ red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int s = warp; s < nslots; s += T / 32) {
        double v = 0.0;
#pragma unroll
        for (int t = 0; t < T / 32; t++) v += acc[s * T + lane + 32 * t];
        v = warp_sum(v);
        if (lane == 0) red[s] = v;
    }
}

// deterministic block sum of one double per thread (shuffle tree inside the warps, warps added in order by thread 0)
template <int T>
__device__ __forceinline__ double km_block_sum(double v, double* __restrict__ wbuf) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) wbuf[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < T / 32; w++) s += wbuf[w];
    __syncthreads();
    return s;
}

// Shared memory: cen[K d] | best[K d] | red[K (d + 1)] | wbuf[32] | pick[K] (int) | acc[K (d + 1) T]
template <int T, int DD>  // DD > 0: exactly DD coordinates (no predicated-off work); DD = 0: any d <= KM_DMAX
__global__ void __launch_bounds__(T)
km_restart_kernel(const double* __restrict__ Xt, int64_t m, int d_rt, int K, int nstart, int iter_max, uint64_t seed,
                  double* __restrict__ cta_centers, double* __restrict__ cta_obj, int32_t* __restrict__ cta_info) {
    extern __shared__ __align__(16) double km_sm[];
    constexpr int DM = DD ? DD : KM_DMAX;
    const int d = DD ? DD : d_rt;
    const int nslots = K * (d + 1);
    double* cen = km_sm;
    double* best = cen + K * d;
    double* red = best + K * d;
    double* wbuf = red + nslots;
    int* pick = reinterpret_cast<int*>(wbuf + 32);
    double* acc = wbuf + 32 + (K + 1) / 2 + 1;
    const int tid = threadIdx.x;
    double best_obj = INFINITY;
    int best_r = -1, best_it = 0;
    for (int r = blockIdx.x; r < nstart; r += gridDim.x) {
        if (tid == 0) {  // K rows of x that differ from each other (kmeans draws from unique(x), stats::kmeans with nstart >= 2)
            for (int c = 0; c < K; c++) {
                int64_t idx = 0;
                for (int att = 0; att < 64; att++) {
                    idx = (int64_t)(km_mix(km_mix(seed ^ ((uint64_t)r << 20)) + (uint64_t)c * 64 + att) % (uint64_t)m);
                    bool dup = false;
                    for (int p = 0; p < c && !dup; p++) {
                        bool same = true;
                        for (int j = 0; j < d && same; j++) same = Xt[(int64_t)j * m + idx] == Xt[(int64_t)j * m + pick[p]];
                        dup = same;
                    }
                    if (!dup) break;
                }
                pick[c] = (int)idx;
                for (int j = 0; j < d; j++) cen[c * d + j] = Xt[(int64_t)j * m + idx];
            }
        }
        __syncthreads();
        int it = 0;
        while (true) {
            for (int s = 0; s < nslots; s++) acc[s * T + tid] = 0.0;
            for (int64_t i = tid; i < m; i += T) {
                double x[DM], db;
                km_load(Xt, m, d, i, x);
                const int c = km_nearest(x, cen, K, d, db);
                double* a = acc + (int64_t)c * (d + 1) * T + tid;
#pragma unroll
                for (int j = 0; j < DM; j++)
                    if (j < d) a[j * T] += x[j];
                a[d * T] += 1.0;
            }
            __syncthreads();
            km_reduce_slots<T>(acc, nslots, red);
            __syncthreads();
            int changed = 0;
            for (int e = tid; e < K * d; e += T) {
                const int c = e / d, j = e - c * d;
                const double cnt = red[c * (d + 1) + d];
                const double nv = cnt > 0.0 ? red[c * (d + 1) + j] / cnt : cen[e];  // an emptied cluster keeps its centre
                changed |= (nv != cen[e]);
                cen[e] = nv;
            }
            it++;
            changed = __syncthreads_or(changed);
            if (!changed || it >= iter_max) break;
        }
        double part = 0.0;
        for (int64_t i = tid; i < m; i += T) {
            double x[DM], db;
            km_load(Xt, m, d, i, x);
            km_nearest(x, cen, K, d, db);
            part += db;
        }
        const double obj = km_block_sum<T>(part, wbuf);
        if (obj < best_obj) {  // starts come in increasing order: ties keep the lowest start
            best_obj = obj; best_r = r; best_it = it;
            for (int e = tid; e < K * d; e += T) best[e] = cen[e];
        }
        __syncthreads();
    }
    for (int e = tid; e < K * d; e += T) cta_centers[(int64_t)blockIdx.x * K * d + e] = best[e];
    if (tid == 0) {
        cta_obj[blockIdx.x] = best_obj;
        cta_info[2 * blockIdx.x] = best_r;
        cta_info[2 * blockIdx.x + 1] = best_it;
    }
}

// out_stats: [tot.withinss | withinss[K] | size[K]];  out_info: [winning start, its Lloyd iterations, transfers, flag: 1 = transfer cap hit]
template <int T, int DD>
__global__ void __launch_bounds__(T)
km_finish_kernel(const double* __restrict__ Xt, int64_t m, int d_rt, int K, int ncta, const double* __restrict__ cta_centers,
                 const double* __restrict__ cta_obj, const int32_t* __restrict__ cta_info, int max_transfers, int32_t* __restrict__ label,
                 double* __restrict__ centers, double* __restrict__ out_stats, int32_t* __restrict__ out_info) {
    extern __shared__ __align__(16) double km_sm[];
    constexpr int DM = DD ? DD : KM_DMAX;
    const int d = DD ? DD : d_rt;
    const int nslots = K * (d + 1);
    double* cen = km_sm;
    double* cnt = cen + K * d;           // [K]
    double* red = cnt + K;               // [nslots]
    double* wbuf = red + nslots;         // [32]
    double* gbuf = wbuf + 32;            // [T / 32] best gain of a warp
    long long* ibuf = reinterpret_cast<long long*>(gbuf + 32);  // [T / 32] its point, [32 + w] its target cluster
    double* acc = gbuf + 32 + 64;
    __shared__ int s_win, s_move_to;
    __shared__ long long s_move_i;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        int w = 0;
        for (int c = 1; c < ncta; c++)
            if (cta_info[2 * c] >= 0 && (cta_info[2 * w] < 0 || cta_obj[c] < cta_obj[w] || (cta_obj[c] == cta_obj[w] && cta_info[2 * c] < cta_info[2 * w]))) w = c;
        s_win = w;
    }
    __syncthreads();
    for (int e = tid; e < K * d; e += T) cen[e] = cta_centers[(int64_t)s_win * K * d + e];
    __syncthreads();
    for (int64_t i = tid; i < m; i += T) {
        double x[DM], db;
        km_load(Xt, m, d, i, x);
        label[i] = km_nearest(x, cen, K, d, db);
    }
    __syncthreads();
    // exact centres and sizes of the partition given by `label`
    auto recompute = [&]() {
        for (int s = 0; s < nslots; s++) acc[s * T + tid] = 0.0;
        for (int64_t i = tid; i < m; i += T) {
            double x[DM];
            km_load(Xt, m, d, i, x);
            double* a = acc + (int64_t)label[i] * (d + 1) * T + tid;
#pragma unroll
            for (int j = 0; j < DM; j++)
                if (j < d) a[j * T] += x[j];
            a[d * T] += 1.0;
        }
        __syncthreads();
        km_reduce_slots<T>(acc, nslots, red);
        __syncthreads();
        for (int e = tid; e < K * d; e += T) {
            const int c = e / d, j = e - c * d;
            const double n = red[c * (d + 1) + d];
            if (n > 0.0) cen[e] = red[c * (d + 1) + j] / n;
        }
        for (int c = tid; c < K; c += T) cnt[c] = red[c * (d + 1) + d];
        __syncthreads();
    };
    recompute();
    // Hartigan-Wong's optimal-transfer criterion (AS 136): moving point i from its cluster a (n_a > 1) to b lowers the objective
    // iff n_b / (n_b + 1) |x - c_b|^2 < n_a / (n_a - 1) |x - c_a|^2.  The single best transfer is applied (centres updated
    // incrementally like AS 136 does), until none is left.
    int transfers = 0, capped = 0;
    while (true) {
        double gbest = 0.0;
        long long ibest = -1;
        int tbest = -1;
        for (int64_t i = tid; i < m; i += T) {
            double x[DM];
            km_load(Xt, m, d, i, x);
            const int a = label[i];
            const double na = cnt[a];
            if (na <= 1.0) continue;
            double r1 = 0.0;
#pragma unroll
            for (int j = 0; j < DM; j++)
                if (j < d) { const double t = x[j] - cen[a * d + j]; r1 = fma(t, t, r1); }
            r1 *= na / (na - 1.0);
            for (int b = 0; b < K; b++) {
                if (b == a) continue;
                double r2 = 0.0;
#pragma unroll
                for (int j = 0; j < DM; j++)
                    if (j < d) { const double t = x[j] - cen[b * d + j]; r2 = fma(t, t, r2); }
                r2 *= cnt[b] / (cnt[b] + 1.0);
                const double gain = r1 - r2;
                if (gain > 1e-13 * r1 && gain > gbest) { gbest = gain; ibest = i; tbest = b; }
            }
        }
        // arg max over the CTA; ties: the lowest point index
        for (int o = 16; o > 0; o >>= 1) {
            const double g2 = __shfl_xor_sync(0xffffffffu, gbest, o);
            const long long i2 = __shfl_xor_sync(0xffffffffu, ibest, o);
            const int t2 = __shfl_xor_sync(0xffffffffu, tbest, o);
            if (i2 >= 0 && (ibest < 0 || g2 > gbest || (g2 == gbest && i2 < ibest))) { gbest = g2; ibest = i2; tbest = t2; }
        }
        if (lane == 0) { gbuf[warp] = gbest; ibuf[warp] = ibest; ibuf[32 + warp] = tbest; }
        __syncthreads();
        if (tid == 0) {
            double g = 0.0;
            long long bi = -1;
            int bt = -1;
            for (int w = 0; w < T / 32; w++)
                if (ibuf[w] >= 0 && (bi < 0 || gbuf[w] > g || (gbuf[w] == g && ibuf[w] < bi))) { g = gbuf[w]; bi = ibuf[w]; bt = (int)ibuf[32 + w]; }
            s_move_i = bi;
            s_move_to = bt;
        }
        __syncthreads();
        if (s_move_i < 0) break;
        if (transfers >= max_transfers) { capped = 1; break; }
        if (tid < d) {
            const int a = label[s_move_i], b = s_move_to;
            const double xv = Xt[(int64_t)tid * m + s_move_i], na = cnt[a], nb = cnt[b];
            cen[a * d + tid] = (cen[a * d + tid] * na - xv) / (na - 1.0);
            cen[b * d + tid] = (cen[b * d + tid] * nb + xv) / (nb + 1.0);
        }
        __syncthreads();
        if (tid == 0) {
            const int a = label[s_move_i];
            cnt[a] -= 1.0;
            cnt[s_move_to] += 1.0;
            label[s_move_i] = s_move_to;
        }
        transfers++;
        __syncthreads();
    }
    recompute();
    // within-cluster sums of squares about the exact centres
    for (int c = 0; c < K; c++) acc[c * T + tid] = 0.0;
    for (int64_t i = tid; i < m; i += T) {
        double x[DM];
        km_load(Xt, m, d, i, x);
        const int a = label[i];
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < DM; j++)
            if (j < d) { const double t = x[j] - cen[a * d + j]; s = fma(t, t, s); }
        acc[a * T + tid] += s;
    }
    __syncthreads();
    km_reduce_slots<T>(acc, K, red);
    __syncthreads();
    for (int e = tid; e < K * d; e += T) centers[e] = cen[e];
    if (tid == 0) {
        double tot = 0.0;
        for (int c = 0; c < K; c++) { tot += red[c]; out_stats[1 + c] = red[c]; out_stats[1 + K + c] = cnt[c]; }
        out_stats[0] = tot;
        out_info[0] = cta_info[2 * s_win];
        out_info[1] = cta_info[2 * s_win + 1];
        out_info[2] = transfers;
        out_info[3] = capped;
    }
}

template <int T, int DD>
inline int kmeans_launch(scm_ctx* ctx, const double* Xt, int64_t m, int d, int K, int nstart, int iter_max, uint64_t seed, int max_transfers,
                         int32_t* d_label, double* d_centers, double* d_stats, int32_t* d_info, Scratch& ws) {
    const int nslots = K * (d + 1);
    const size_t smem_r = sizeof(double) * ((size_t)2 * K * d + nslots + 32 + (K + 1) / 2 + 1 + (size_t)nslots * T);
    const size_t smem_f = sizeof(double) * ((size_t)K * d + K + nslots + 32 + 32 + 64 + (size_t)nslots * T);
    SCM_CUDA(cudaFuncSetAttribute(km_restart_kernel<T, DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r));
    SCM_CUDA(cudaFuncSetAttribute(km_finish_kernel<T, DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
    int per_sm = (int)((220 * 1024) / (smem_r + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2048 / T) per_sm = 2048 / T;
    int grid = ctx->num_sms * per_sm;
    if (grid > nstart) grid = nstart;
    double *cta_centers, *cta_obj;
    int32_t* cta_info;
    SCM_TRY(ws.alloc(&cta_centers, (int64_t)grid * K * d));
    SCM_TRY(ws.alloc(&cta_obj, grid));
    SCM_TRY(ws.alloc(&cta_info, 2 * (int64_t)grid));
    km_restart_kernel<T, DD><<<grid, T, smem_r, ctx->stream>>>(Xt, m, d, K, nstart, iter_max, seed, cta_centers, cta_obj, cta_info);
    SCM_LAUNCH_CHECK(ctx);
    km_finish_kernel<T, DD><<<1, T, smem_f, ctx->stream>>>(Xt, m, d, K, grid, cta_centers, cta_obj, cta_info, max_transfers, d_label, d_centers,
                                                       d_stats, d_info);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

// stats::kmeans(x, centers = K, nstart, algorithm = "Hartigan-Wong") up to the random stream (see the header).  d_X: m x d.
// d_label: m cluster ids in [0, K); d_centers: K x d; h_withinss, h_size: K; h_info: [winning start, Lloyd iterations, transfers, cap flag].
inline int kmeans(scm_ctx* ctx, const double* d_X, int64_t m, int32_t d, int64_t ldx, int32_t K, int32_t nstart, int32_t iter_max, uint64_t seed,
                  int32_t* d_label, double* d_centers, double* h_withinss, int64_t* h_size, double* h_tot_withinss, int32_t* h_info) {
    if (!d_X || !d_label || !d_centers || m <= 0 || d <= 0 || ldx < d || K <= 0 || nstart <= 0 || iter_max <= 0)
        return fail(SCM_ERR_INVALID, "kmeans: bad arguments");
    if (d > KM_DMAX) return fail(SCM_ERR_UNSUPPORTED, "kmeans: d = %d > %d coordinates", d, KM_DMAX);
    if (K > m) return fail(SCM_ERR_INVALID, "kmeans: more cluster centers than data points");
    if (m >= ((int64_t)1 << 31)) return fail(SCM_ERR_UNSUPPORTED, "kmeans: m >= 2^31");
    const int nslots = K * (d + 1);
    if (nslots > 800) return fail(SCM_ERR_UNSUPPORTED, "kmeans: K (d + 1) = %d > 800", nslots);
    Scratch ws(ctx);
    double *Xt, *d_stats;
    int32_t* d_info;
    SCM_TRY(ws.alloc(&Xt, m * d));
    SCM_TRY(ws.alloc(&d_stats, 1 + 2 * (int64_t)K));
    SCM_TRY(ws.alloc(&d_info, 4));
    km_transpose_kernel<<<(unsigned)std::min<int64_t>(ceil_div(m * d, 256), 4096), 256, 0, ctx->stream>>>(d_X, m, d, ldx, Xt);
    SCM_LAUNCH_CHECK(ctx);
    const int max_transfers = 5000;  // every transfer is one pass of a single CTA over the points; after a converged Lloyd run a handful are needed (the cap is reported: h_info[3])
    int rc;
#define SCM_KM_GO(T_, DD_) rc = kmeans_launch<T_, DD_>(ctx, Xt, m, d, K, nstart, iter_max, seed, max_transfers, d_label, d_centers, d_stats, d_info, ws)
    if (nslots * 128 * 8 <= 96 * 1024) { if (d == 10) SCM_KM_GO(128, 10); else SCM_KM_GO(128, 0); }  // d = 10: scReplicate's first 10 PCs
    else if (nslots * 64 * 8 <= 200 * 1024) { if (d == 10) SCM_KM_GO(64, 10); else SCM_KM_GO(64, 0); }
    else SCM_KM_GO(32, 0);
#undef SCM_KM_GO
    if (rc != 0) return rc;
    std::vector<double> st(1 + 2 * (size_t)K);
    int32_t info[4];
    SCM_CUDA(cudaMemcpyAsync(st.data(), d_stats, sizeof(double) * st.size(), cudaMemcpyDeviceToHost, ctx->stream));
    SCM_CUDA(cudaMemcpyAsync(info, d_info, sizeof(info), cudaMemcpyDeviceToHost, ctx->stream));
    SCM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h_tot_withinss) *h_tot_withinss = st[0];
    for (int c = 0; c < K; c++) {
        if (h_withinss) h_withinss[c] = st[1 + c];
        if (h_size) h_size[c] = (int64_t)st[1 + K + c];
    }
    if (h_info) for (int i = 0; i < 4; i++) h_info[i] = info[i];
    return 0;
}

}  // namespace scm
