// C ABI of libscmerge_b200 (see include/scmerge_b200.h).  Single translation unit: all kernels are header-only.
#include "adjust.cuh"
#include "adjust_tma.cuh"
#include "comm.cuh"
#include "common.cuh"
#include "csc_ingest.cuh"
#include "eig_topk.cuh"
#include "fit_core.cuh"
#include "fit_dual.cuh"
#include "gemm_f64.cuh"
#include "gram_syrk.cuh"
#include "gram_tma.cuh"
#include "seg_colsum.cuh"

namespace scm {

__global__ void transpose_kernel(const double* __restrict__ in, int64_t rows, int64_t cols, int64_t ldi,
                                 double* __restrict__ out, int64_t ldo) {
    __shared__ double tile[32][33];
    const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int64_t r = r0 + j, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[j][threadIdx.x] = in[r * ldi + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int64_t c = c0 + j, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[c * ldo + r] = tile[threadIdx.x][j];
    }
}

// meta[0] = nonfinite flag, meta[1] = local rows (summed across ranks by the hook)
__global__ void meta_kernel(const int32_t* __restrict__ flag, int64_t m_local, int64_t* __restrict__ meta) {
    meta[0] = flag[0] ? 1 : 0;
    meta[1] = m_local;
}

// mean[j] = sum_g sums[g, j] / total
__global__ void colmean_kernel(const double* __restrict__ sums, int32_t G, int64_t n, const int64_t* __restrict__ meta,
                               double* __restrict__ mean) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double s = 0.0;
    for (int g = 0; g < G; g++) s += sums[(int64_t)g * n + j];
    mean[j] = s / (double)meta[1];
}

// tot[j] = sum_g sums[g, j]
__global__ void group_total_kernel(const double* __restrict__ sums, int32_t G, int64_t n, double* __restrict__ tot) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double s = 0.0;
    for (int g = 0; g < G; g++) s += sums[(int64_t)g * n + j];
    tot[j] = s;
}

// out[j] = sum_g counts[g] * means[g, j] / sum_g counts[g]: the column means of all cells from the group means.  A CTA takes 32
// genes; its 32 thread rows split the groups (g = ty, ty + 32, ...: with the 4800 pseudobulks of scMerge2 one thread per gene was a
// 4800-long dependent chain, ~1 ms) and are added in row order: deterministic, independent of the grid.
__global__ void __launch_bounds__(1024)
weighted_colmean_kernel(const double* __restrict__ means, const int64_t* __restrict__ counts, int32_t G, int64_t n, int64_t ldm,
                        double* __restrict__ out) {
    __shared__ double ps[32][33];
    __shared__ double pc[32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t j = (int64_t)blockIdx.x * 32 + tx;
    double s = 0.0, tot = 0.0;  // counts < 2^53 are exact in float64
    for (int g = ty; g < G; g += 32) {
        const double c = (double)counts[g];
        tot += c;
        if (j < n) s += c * means[(int64_t)g * ldm + j];
    }
    ps[ty][tx] = s;
    if (tx == 0) pc[ty] = tot;
    __syncthreads();
    if (ty == 0 && j < n) {
        double a = 0.0, t = 0.0;
        for (int r = 0; r < 32; r++) { a += ps[r][tx]; t += pc[r]; }
        out[j] = t > 0.0 ? a / t : 0.0;
    }
}

// pooled within-batch sd (R/scRUVIII.R:170): sd[j] = sqrt( sum_b ssq[b, j] / (m - #non-empty batches) )
__global__ void pooled_sd_kernel(const double* __restrict__ ssq, const int64_t* __restrict__ counts, int32_t Bt, int64_t n,
                                 const int64_t* __restrict__ meta, double* __restrict__ sd, double* __restrict__ inv_sd) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int nb = 0;
    double s = 0.0;
    for (int b = 0; b < Bt; b++) { nb += counts[b] > 0; s += ssq[(int64_t)b * n + j]; }
    const double v = s / (double)(meta[1] - nb);
    sd[j] = sqrt(v);
    inv_sd[j] = 1.0 / sqrt(v);  // var == 0 -> Inf, like the reference (no zero-variance guard, R/scRUVIII.R:171)
}

__global__ void fullalpha_kernel(const double* __restrict__ vals, const double* __restrict__ vecs, int32_t s, int64_t n,
                                 double* __restrict__ fa, double* __restrict__ sigma) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    const double sg = sqrt(fmax(vals[r], 0.0));
    if (j < n && fa) fa[(int64_t)r * n + j] = sg * vecs[(int64_t)r * n + j];
    if (j == 0 && sigma) sigma[r] = sg;
}

// K7: per-cell cosine normalisation  out[i, :] = Y[i, :] / max(||Y[i, :]||_2, min_norm)   (batchelor::cosineNorm,
// call site R/scMerge2.R:193-198).  One CTA per cell: the row is held in registers between the norm and the scale, so
// the traffic is the algorithmic 8*m*n read + 8*m*n written.
template <int PER>
__global__ void __launch_bounds__(256)
cosine_rows_kernel(const double* __restrict__ Y, int64_t m, int64_t n, int64_t ld, double* __restrict__ out, int64_t ldo,
                   double min_norm) {
    __shared__ double red[8];
    __shared__ double s_inv;
    for (int64_t row = blockIdx.x; row < m; row += gridDim.x) {
        const double* src = Y + row * ld;
        double v[PER];
        double ss = 0.0;
#pragma unroll
        for (int j = 0; j < PER; j++) {
            const int64_t c = (int64_t)j * 256 + threadIdx.x;
            v[j] = c < n ? src[c] : 0.0;
            ss += v[j] * v[j];
        }
        ss = warp_sum(ss);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; w++) t += red[w];
            s_inv = 1.0 / fmax(sqrt(t), min_norm);
        }
        __syncthreads();
        const double inv = s_inv;
        double* dst = out + row * ldo;
#pragma unroll
        for (int j = 0; j < PER; j++) {
            const int64_t c = (int64_t)j * 256 + threadIdx.x;
            if (c < n) dst[c] = v[j] * inv;
        }
        __syncthreads();
    }
}

inline int cosine_rows(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, double* d_out, int64_t ldo, double min_norm) {
    if (m <= 0 || n <= 0) return 0;
    const int per = (int)ceil_div(n, 256);
    const unsigned grid = (unsigned)(m < (int64_t)ctx->num_sms * 16 ? m : (int64_t)ctx->num_sms * 16);
#define COS_LAUNCH(P) cosine_rows_kernel<P><<<grid, 256, 0, ctx->stream>>>(d_Y, m, n, ld, d_out, ldo, min_norm)
    if (per <= 4) COS_LAUNCH(4);
    else if (per <= 8) COS_LAUNCH(8);
    else if (per <= 16) COS_LAUNCH(16);
    else if (per <= 32) COS_LAUNCH(32);
    else if (per <= 96) COS_LAUNCH(96);
    else return fail(SCM_ERR_UNSUPPORTED, "cosine_rows: more than 24576 genes per cell");
#undef COS_LAUNCH
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace scm

#include "host_pipeline.cuh"
#include "ruvg.cuh"
#include "select_k.cuh"
#include "csc_sparse.cuh"
#include "replicate.cuh"
#include "kmeans.cuh"

using namespace scm;

#define SCM_REQUIRE_CTX(ctx)                                             \
    do {                                                                 \
        if (!(ctx)) return scm::fail(SCM_ERR_INVALID, "null context");   \
        SCM_CUDA(cudaSetDevice((ctx)->device));                          \
    } while (0)

extern "C" {

const char* scm_last_error(void) { return scm::g_err; }
int scm_version(void) { return 100; }

int scm_device_pci_bus_id(int device, char* buf, int len) {
    if (!buf || len < 16) return fail(SCM_ERR_INVALID, "scm_device_pci_bus_id: buffer too small");
    SCM_CUDA(cudaDeviceGetPCIBusId(buf, len, device));
    return 0;
}

int scm_ctx_create(int device, void* cuda_stream, scm_ctx** out) {
    if (!out) return fail(SCM_ERR_INVALID, "scm_ctx_create: out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(SCM_ERR_CUDA, "no CUDA device available (%s): libscmerge_b200 has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(SCM_ERR_INVALID, "device %d out of range (have %d)", device, ndev);
    SCM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SCM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(SCM_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    struct Guard {  // the half-built context must not leak when one of the CUDA calls below fails
        scm_ctx* c;
        ~Guard() { if (c) scm_ctx_destroy(c); }
    } guard{new scm_ctx()};
    scm_ctx* c = guard.c;
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    if (const char* e = getenv("SCM_GRAM_TMA")) c->use_tma = (e[0] != '0');
    if (const char* e = getenv("SCM_GRAM_LEAN")) c->gram_lean = atoi(e) & 3;
    if (const char* e = getenv("SCM_GRAM_ORDER")) c->gram_class_order = (e[0] != '0');
    if (const char* e = getenv("SCM_GRAM_PRESHIFT")) c->gram_preshift = (e[0] != '0');
    if (const char* e = getenv("SCM_K1_FUSED_COPY")) c->k1_fused_copy = (e[0] != '0');
    if (cuda_stream && cuda_stream != (void*)2) { c->stream = (cudaStream_t)cuda_stream; c->own_stream = false; }
    else {
        // own stream: the greatest priority, or -- handle 2, a BACKGROUND context -- the least one: the block scheduler then gives
        // free SMs to the foreground context's kernels first (scMerge2: the whole-spectrum solve beside the adjust + download)
        int least = 0, greatest = 0;
        SCM_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        SCM_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, cuda_stream ? least : greatest));
        c->own_stream = true;
    }
    for (int i = 0; i < ST_COUNT; i++) {
        SCM_CUDA(cudaEventCreate(&c->ev0[i]));
        SCM_CUDA(cudaEventCreate(&c->ev1[i]));
        c->ev_used[i] = false;
    }
    SCM_CUDA(cudaMallocHost((void**)&c->h_pinned, sizeof(double) * 4096));
    // keep freed scratch in the pool: the fit re-allocates the same sizes every call
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thresh = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);
    }
    guard.c = nullptr;
    *out = c;
    return 0;
}

int scm_ctx_destroy(scm_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    comm_release(ctx);
    for (int i = 0; i < ST_COUNT; i++) {
        if (ctx->ev0[i]) cudaEventDestroy(ctx->ev0[i]);
        if (ctx->ev1[i]) cudaEventDestroy(ctx->ev1[i]);
    }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->copy_stream_d2h) cudaStreamDestroy(ctx->copy_stream_d2h);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    for (void* b : ctx->stage_buf) if (b) cudaFreeHost(b);
    delete ctx;
    return 0;
}

int scm_ctx_set_allreduce(scm_ctx* ctx, scm_allreduce_fn fn, void* user, int world_size) {
    if (!ctx) return fail(SCM_ERR_INVALID, "null context");
    ctx->allreduce = fn; ctx->allreduce_user = user; ctx->world = world_size;
    return 0;
}

int scm_nccl_unique_id(void* id128) {
    if (!id128) return fail(SCM_ERR_INVALID, "scm_nccl_unique_id: null buffer");
    const NcclApi& api = nccl_api();
    if (!api.ok) return fail(SCM_ERR_COMM, "NCCL is not available: %s", api.why.c_str());
    nccl_unique_id id;
    SCM_NCCL(api.GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return 0;
}

int scm_ctx_comm_init_nccl(scm_ctx* ctx, const void* id128, int rank, int world_size) {
    SCM_REQUIRE_CTX(ctx);
    if (world_size < 1 || rank < 0 || rank >= world_size || (world_size > 1 && !id128))
        return fail(SCM_ERR_INVALID, "scm_ctx_comm_init_nccl: bad rank %d / world %d", rank, world_size);
    comm_release(ctx);
    ctx->world = world_size; ctx->rank = rank;
    if (world_size == 1) return 0;
    const NcclApi& api = nccl_api();
    if (!api.ok) return fail(SCM_ERR_COMM, "NCCL is not available: %s", api.why.c_str());
    nccl_unique_id id;
    memcpy(id.internal, id128, 128);
    nccl_comm_t comm = nullptr;
    SCM_NCCL(api.CommInitRank(&comm, world_size, id, rank));
    ctx->nccl_comm = comm; ctx->comm_owned = true;
    return 0;
}

int scm_ctx_comm_info(scm_ctx* ctx, int* rank, int* world_size, int* transport) {
    if (!ctx) return fail(SCM_ERR_INVALID, "null context");
    if (rank) *rank = ctx->rank;
    if (world_size) *world_size = ctx->world;
    if (transport) *transport = ctx->world <= 1 ? 0 : (ctx->nccl_comm ? 1 : (ctx->allreduce ? 2 : 0));
    return 0;
}

int scm_allreduce_sum(scm_ctx* ctx, void* d_buf, int64_t count, int dtype) {
    SCM_REQUIRE_CTX(ctx);
    if (!d_buf || count < 0 || (dtype != 0 && dtype != 1)) return fail(SCM_ERR_INVALID, "scm_allreduce_sum: bad arguments");
    return allreduce(ctx, d_buf, count, dtype);
}

int scm_ctx_fit_info(scm_ctx* ctx, int32_t* iters, double* resid, int32_t* conv_rank, int32_t* products) {
    if (!ctx) return fail(SCM_ERR_INVALID, "null context");
    if (iters) *iters = ctx->fit_iters;
    if (resid) *resid = ctx->fit_resid;
    if (conv_rank) *conv_rank = ctx->fit_conv_rank;
    if (products) *products = ctx->fit_products;
    return 0;
}

// ---- one process, several GPUs -------------------------------------------------------------------------------------
int scm_multi_create(const int* devices, int ndev, scm_multi** out) {
    if (!out || ndev < 1 || ndev > 64) return fail(SCM_ERR_INVALID, "scm_multi_create: bad arguments");
    std::vector<int> dev(ndev);
    for (int i = 0; i < ndev; i++) dev[i] = devices ? devices[i] : i;
    for (int i = 0; i < ndev; i++)
        for (int j = 0; j < i; j++)
            if (dev[i] == dev[j]) return fail(SCM_ERR_INVALID, "scm_multi_create: device %d listed twice", dev[i]);
    scm_multi* mg = new scm_multi();
    struct Guard { scm_multi* m; bool ok; ~Guard() { if (!ok) scm_multi_destroy(m); } } guard{mg, false};
    for (int i = 0; i < ndev; i++) {
        scm_ctx* c = nullptr;
        SCM_TRY(scm_ctx_create(dev[i], nullptr, &c));
        c->rank = i; c->world = ndev;
        mg->ctx.push_back(c);
    }
    if (ndev > 1) {
        const NcclApi& api = nccl_api();
        if (!api.ok) return fail(SCM_ERR_COMM, "NCCL is not available: %s", api.why.c_str());
        mg->comms.assign(ndev, nullptr);
        SCM_NCCL(api.CommInitAll(mg->comms.data(), ndev, dev.data()));
        for (int i = 0; i < ndev; i++) { mg->ctx[i]->nccl_comm = mg->comms[i]; mg->ctx[i]->comm_owned = false; }
    }
    guard.ok = true;
    *out = mg;
    return 0;
}

int scm_multi_destroy(scm_multi* mg) {
    if (!mg) return 0;
    for (scm_ctx* c : mg->ctx) {
        if (!c) continue;
        cudaSetDevice(c->device);
        if (c->stream) cudaStreamSynchronize(c->stream);
    }
    for (auto comm : mg->comms) if (comm) nccl_api().CommDestroy(comm);
    for (scm_ctx* c : mg->ctx) { if (c) { c->nccl_comm = nullptr; scm_ctx_destroy(c); } }
    delete mg;
    return 0;
}

int scm_multi_size(const scm_multi* mg) { return mg ? (int)mg->ctx.size() : 0; }

scm_ctx* scm_multi_ctx(scm_multi* mg, int i) { return (mg && i >= 0 && i < (int)mg->ctx.size()) ? mg->ctx[i] : nullptr; }

int scm_ctx_set_conv_ranks(scm_ctx* ctx, const int32_t* ks, int32_t nk) {
    if (!ctx || nk < 0 || nk > 16 || (nk > 0 && !ks)) return fail(SCM_ERR_INVALID, "scm_ctx_set_conv_ranks: bad arguments (at most 16 ranks)");
    for (int i = 0; i < nk; i++) ctx->conv_ranks[i] = ks[i];
    ctx->n_conv_ranks = nk;
    return 0;
}

int scm_ctx_sync(scm_ctx* ctx) {
    SCM_REQUIRE_CTX(ctx);
    SCM_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int scm_ctx_release_scratch(scm_ctx* ctx) {
    SCM_REQUIRE_CTX(ctx);
    SCM_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaMemPool_t pool;
    SCM_CUDA(cudaDeviceGetDefaultMemPool(&pool, ctx->device));
    SCM_CUDA(cudaMemPoolTrimTo(pool, 0));
    return 0;
}

int scm_ctx_launch_count(scm_ctx* ctx, int64_t* n) {
    if (!ctx || !n) return fail(SCM_ERR_INVALID, "null argument");
    *n = ctx->launches;
    return 0;
}

int scm_dev_alloc(scm_ctx* ctx, int64_t bytes, void** d_ptr) {
    SCM_REQUIRE_CTX(ctx);
    if (bytes < 0 || !d_ptr) return fail(SCM_ERR_INVALID, "scm_dev_alloc: bad arguments");
    // stream-ordered pool allocation: cudaMalloc/cudaFree serialise the device and cost up to ~1 s per call
    // next to a populated async pool (measured), the pool calls are microseconds
    SCM_CUDA(cudaMallocAsync(d_ptr, (size_t)(bytes > 0 ? bytes : 16), ctx->stream));
    return 0;
}
int scm_dev_free(scm_ctx* ctx, void* d_ptr) {
    SCM_REQUIRE_CTX(ctx);
    if (!d_ptr) return 0;
    SCM_CUDA(cudaFreeAsync(d_ptr, ctx->stream));
    return 0;
}
int scm_host_alloc_pinned(int64_t bytes, void** h_ptr) {
    if (bytes < 0 || !h_ptr) return fail(SCM_ERR_INVALID, "scm_host_alloc_pinned: bad arguments");
    SCM_CUDA(cudaMallocHost(h_ptr, (size_t)(bytes > 0 ? bytes : 16)));
    return 0;
}
int scm_host_free_pinned(void* h_ptr) {
    SCM_CUDA(cudaFreeHost(h_ptr));
    return 0;
}
int scm_h2d(scm_ctx* ctx, void* d_dst, const void* h_src, int64_t bytes) {
    SCM_REQUIRE_CTX(ctx);
    SCM_CUDA(cudaMemcpyAsync(d_dst, h_src, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}
int scm_d2h(scm_ctx* ctx, void* h_dst, const void* d_src, int64_t bytes) {
    SCM_REQUIRE_CTX(ctx);
    SCM_CUDA(cudaMemcpyAsync(h_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SCM_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int scm_copy2d(scm_ctx* ctx, void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t row_bytes,
               int64_t rows, int kind) {
    SCM_REQUIRE_CTX(ctx);
    if (rows <= 0 || row_bytes <= 0) return 0;
    const cudaMemcpyKind kk = kind == 0 ? cudaMemcpyHostToDevice : (kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
    SCM_CUDA(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)row_bytes, (size_t)rows, kk, ctx->stream));
    if (kind == 1) SCM_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int scm_transpose(scm_ctx* ctx, const double* d_in, int64_t rows, int64_t cols, int64_t ldi, double* d_out, int64_t ldo) {
    SCM_REQUIRE_CTX(ctx);
    if (rows <= 0 || cols <= 0) return 0;
    if (ceil_div(rows, 32) > 65535) return fail(SCM_ERR_UNSUPPORTED, "scm_transpose: more than 2M rows; transpose by column blocks");
    dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32));
    transpose_kernel<<<grid, dim3(32, 8), 0, ctx->stream>>>(d_in, rows, cols, ldi, d_out, ldo);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

int scm_cosine_rows(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, double* d_out, int64_t ldo, double min_norm) {
    SCM_REQUIRE_CTX(ctx);
    StageTimer t(ctx, ST_INGEST);
    return cosine_rows(ctx, d_Y, m, n, ld, d_out, ldo, min_norm);
}

int scm_csc_to_dense(scm_ctx* ctx, const int64_t* d_indptr, const int32_t* d_idx, const double* d_val, int64_t m, int64_t n,
                     int64_t nnz, int cosine, double min_norm, double* d_out, int64_t ldo, double* d_inv_norm, int32_t* d_nonfinite) {
    SCM_REQUIRE_CTX(ctx);
    StageTimer t(ctx, ST_INGEST);
    return csc_to_dense(ctx, d_indptr, d_idx, d_val, m, n, nnz, cosine, min_norm, d_out, ldo, d_inv_norm, d_nonfinite);
}

int scm_csc_check(scm_ctx* ctx, const int64_t* d_indptr, const int32_t* d_idx, int64_t m, int64_t n, int64_t nnz, int32_t* canonical) {
    SCM_REQUIRE_CTX(ctx);
    if (m < 0 || n <= 0 || !d_indptr || (nnz > 0 && !d_idx)) return fail(SCM_ERR_INVALID, "scm_csc_check: bad arguments");
    Scratch ws(ctx);
    int32_t* flags;
    SCM_TRY(ws.alloc(&flags, 2));
    SCM_TRY(zero_async(ctx, flags, 8));
    if (m > 0) {
        csc_check_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(d_indptr, d_idx, m, n, nnz, flags);
        SCM_LAUNCH_CHECK(ctx);
    }
    int32_t h[2] = {0, 0};
    SCM_CUDA(cudaMemcpyAsync(h, flags, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    SCM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h[0]) return fail(SCM_ERR_INVALID, h[0] == 1 ? "scm_csc_check: column pointers are not a non-decreasing sequence inside [0, nnz]"
                                                     : "scm_csc_check: a gene index is outside [0, n)");
    if (canonical) *canonical = h[1] == 0;
    return 0;
}

int scm_csc_rownorm(scm_ctx* ctx, const int64_t* d_indptr, const double* d_val, int64_t m, int cosine, double min_norm, double* d_inv_norm,
                    int32_t* d_nonfinite) {
    SCM_REQUIRE_CTX(ctx);
    if (m <= 0) return 0;
    if (!d_indptr || !d_inv_norm) return fail(SCM_ERR_INVALID, "scm_csc_rownorm: bad arguments");
    StageTimer t(ctx, ST_INGEST);
    csc_rownorm_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(d_indptr, d_val, m, cosine, min_norm, d_inv_norm, d_nonfinite);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

int scm_csc_seg_colmean(scm_ctx* ctx, const int64_t* d_indptr, const int32_t* d_idx, const double* d_val, const double* d_inv_norm,
                        int64_t m, int64_t n, const int32_t* d_key, int32_t G, double* d_means, int64_t ldn, int64_t* d_counts) {
    SCM_REQUIRE_CTX(ctx);
    if (!d_counts) return fail(SCM_ERR_INVALID, "scm_csc_seg_colmean: d_counts is required");
    StageTimer t(ctx, ST_SEG);
    SCM_TRY(csc_seg_colsum(ctx, d_indptr, d_idx, d_val, d_inv_norm, m, n, d_key, G, d_means, ldn, d_counts));
    SCM_TRY(allreduce(ctx, d_means, (int64_t)G * ldn, 0));  // global over the ranks, like scm_seg_colmean
    SCM_TRY(allreduce(ctx, d_counts, G, 1));
    groupmean_kernel<<<dim3((unsigned)ceil_div(ldn, 256), (unsigned)(G < 2048 ? G : 2048)), 256, 0, ctx->stream>>>(d_means, d_counts, G, ldn);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

int scm_csc_adjust(scm_ctx* ctx, const int64_t* d_indptr, const int32_t* d_idx, const double* d_val, const double* d_inv_norm, int64_t m,
                   int64_t n, const scm_coef* coef, double* d_out, int64_t ldo) {
    SCM_REQUIRE_CTX(ctx);
    StageTimer t(ctx, ST_ADJUST);
    return csc_adjust(ctx, d_indptr, d_idx, d_val, d_inv_norm, m, n, coef, d_out, ldo);
}

int scm_csc_adjust_to_host(scm_ctx* ctx, const int64_t* d_indptr, const int32_t* d_idx, const double* d_val, const double* d_inv_norm,
                           int64_t m, int64_t n, const scm_coef* coef, double* h_out, int64_t ldo) {
    SCM_REQUIRE_CTX(ctx);
    if (!coef || coef->n != n) return fail(SCM_ERR_INVALID, "scm_csc_adjust_to_host: coefficient object does not match n");
    return csc_adjust_download(ctx, d_indptr, d_idx, d_val, d_inv_norm, m, n, coef, h_out, ldo);
}

int scm_adjust_to_host(scm_ctx* ctx, double* d_Y, int64_t m, int64_t n, int64_t ld, const scm_coef* coef, double* h_out, int64_t ldo) {
    SCM_REQUIRE_CTX(ctx);
    if (!coef || coef->n != n) return fail(SCM_ERR_INVALID, "scm_adjust_to_host: coefficient object does not match n");
    if (m <= 0) return 0;
    return adjust_download(ctx, d_Y, m, n, ld, coef, h_out, ldo, 0);
}

int scm_seg_colsum(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, const int32_t* d_key, int32_t G,
                   const double* d_center, double* d_sums, int64_t* d_counts, int32_t* d_nonfinite) {
    SCM_REQUIRE_CTX(ctx);
    StageTimer t(ctx, ST_SEG);
    return seg_colsum(ctx, d_Y, m, n, ld, d_key, G, d_center, d_sums, d_counts, d_nonfinite);
}

int scm_seg_colmean(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, const int32_t* d_key, int32_t G,
                    double* d_means, int64_t* d_counts, int32_t* d_nonfinite) {
    SCM_REQUIRE_CTX(ctx);
    if (!d_counts) return fail(SCM_ERR_INVALID, "scm_seg_colmean: d_counts is required");
    StageTimer t(ctx, ST_SEG);
    SCM_TRY(seg_colsum(ctx, d_Y, m, n, ld, d_key, G, nullptr, d_means, d_counts, d_nonfinite));
    // several ranks: a group's cells may live anywhere (pseudobulks of scMerge2, colMeans over ALL cells) -- sums and counts are
    // summed over the ranks before the division, so every rank gets the same, global means
    SCM_TRY(allreduce(ctx, d_means, (int64_t)G * n, 0));
    SCM_TRY(allreduce(ctx, d_counts, G, 1));
    groupmean_kernel<<<dim3((unsigned)ceil_div(n, 256), (unsigned)(G < 2048 ? G : 2048)), 256, 0, ctx->stream>>>(d_means, d_counts, G, n);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

int scm_group_colmean(scm_ctx* ctx, const double* d_means, const int64_t* d_counts, int32_t G, int64_t n, int64_t ldm, double* d_out) {
    SCM_REQUIRE_CTX(ctx);
    weighted_colmean_kernel<<<(unsigned)ceil_div(n, 32), 1024, 0, ctx->stream>>>(d_means, d_counts, G, n, ldm, d_out);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

int scm_gram(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, const double* d_shift, double* d_gram) {
    SCM_REQUIRE_CTX(ctx);
    StageTimer t(ctx, ST_GRAM);
    return gram(ctx, d_Y, m, n, ld, d_shift, d_gram, true);
}

int scm_gram_residual(scm_ctx* ctx, double* d_gram, int64_t n, const double* d_sums, const int64_t* d_counts, int32_t G,
                      const double* d_shift, const double* d_scale) {
    SCM_REQUIRE_CTX(ctx);
    return gram_residual(ctx, d_gram, n, d_sums, d_counts, G, d_shift, d_scale);
}

int scm_eig_topk(scm_ctx* ctx, const double* d_gram, int64_t n, int32_t s, int32_t kconv, int mode, double tol, int32_t maxit,
                 uint64_t seed, double* d_vals, double* d_vecs, int32_t* iters_out, double* resid_out) {
    SCM_REQUIRE_CTX(ctx);
    StageTimer t(ctx, ST_EIG);
    return eig_topk(ctx, d_gram, n, s, kconv, mode, tol, maxit, seed, d_vals, d_vecs, iters_out, resid_out);
}

int scm_fullalpha(scm_ctx* ctx, const double* d_vals, const double* d_vecs, int32_t s, int64_t n, double* d_fullalpha) {
    SCM_REQUIRE_CTX(ctx);
    dim3 grid((unsigned)ceil_div(n, 256), (unsigned)s);
    fullalpha_kernel<<<grid, 256, 0, ctx->stream>>>(d_vals, d_vecs, s, n, d_fullalpha, nullptr);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

int scm_coef_create(scm_ctx* ctx, const double* d_alpha, int64_t n, int32_t k, const uint8_t* d_ctl, const double* d_center,
                    const double* d_pre, const double* d_post, scm_coef** out) {
    SCM_REQUIRE_CTX(ctx);
    if (!out) return fail(SCM_ERR_INVALID, "scm_coef_create: out is NULL");
    StageTimer t(ctx, ST_COEF);
    return coef_create(ctx, d_alpha, n, k, d_ctl, d_center, d_pre, d_post, out);
}
int scm_coef_from_factors(scm_ctx* ctx, const double* d_B, const double* d_A, int64_t n, int32_t k, const uint8_t* d_rows,
                          const double* d_center, scm_coef** out) {
    SCM_REQUIRE_CTX(ctx);
    StageTimer t(ctx, ST_COEF);
    return coef_from_factors(ctx, d_B, d_A, n, k, d_rows, d_center, out);
}
int scm_ruvg_fit(scm_ctx* ctx, const double* d_Y, int64_t m_local, int64_t n, int64_t ld, const uint8_t* d_ctl, int32_t k, int mode,
                 double tol, int32_t maxit, uint64_t seed, double* d_alpha, double* d_center, double* d_sigma, scm_coef** coef_out,
                 int32_t* iters_out, double* resid_out) {
    SCM_REQUIRE_CTX(ctx);
    return ruvg_fit(ctx, d_Y, m_local, n, ld, d_ctl, k, mode, tol, maxit, seed, d_alpha, d_center, d_sigma, coef_out, iters_out, resid_out);
}
int scm_coef_destroy(scm_ctx* ctx, scm_coef* coef) {
    SCM_REQUIRE_CTX(ctx);
    return coef_destroy(ctx, coef);
}

int scm_adjust(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, const scm_coef* coef, const int32_t* d_subset,
               int32_t nsub, double* d_out, int64_t ldo, double* d_W) {
    SCM_REQUIRE_CTX(ctx);
    StageTimer t(ctx, ST_ADJUST);
    return adjust(ctx, d_Y, m, n, ld, coef, d_subset, nsub, d_out, ldo, d_W);
}

static int ruv3_fit_impl(scm_ctx* ctx, const double* d_Y, int64_t m_local, int64_t n, int64_t ld, const int32_t* d_group, int32_t G,
                         const int32_t* d_batch, int32_t Bt, int32_t s, int32_t kconv, int mode, double tol, int32_t maxit, uint64_t seed,
                         double* d_fullalpha, double* d_sigma, double* d_mean, double* d_sd, int32_t* iters_out, double* resid_out,
                         double* d_gram_centred, double* d_colmean) {
    SCM_REQUIRE_CTX(ctx);
    if (!d_Y || !d_group || G <= 0 || n <= 0 || s <= 0 || !d_fullalpha)
        return fail(SCM_ERR_INVALID, "scm_ruv3_fit: bad arguments");
    if (d_batch && (Bt <= 0 || !d_mean || !d_sd)) return fail(SCM_ERR_INVALID, "scm_ruv3_fit: batch given without Bt / d_mean / d_sd");
    if (fit_dual_applies(ctx, m_local, n, s, d_batch, d_gram_centred, d_colmean))  // short matrix, whole row space: the m x m side
        return ruv3_fit_dual(ctx, d_Y, m_local, n, ld, d_group, G, s, d_fullalpha, d_sigma, iters_out, resid_out);
    cudaStream_t st = ctx->stream;
    Scratch ws(ctx);
    double *vals, *vecs, *inv_sd = nullptr;
    SCM_TRY(ws.alloc(&vals, s)); SCM_TRY(ws.alloc(&vecs, (int64_t)s * n));
    // One summing pass + one Gram pass over the local cells, two all-reduces (fit_core.cuh): replicate-group sums (my_residop,
    // R/fastRUVIII.R:121-129), NA/Inf scan (:52-60), standardize2's statistics (R/scRUVIII.R:158-175) and the centred Gram.
    FitState fs;
    SCM_TRY(fit_state(ctx, ws, d_Y, m_local, n, ld, d_group, G, d_batch, Bt, fs));
    if (d_batch) {  // global mean + pooled within-batch sd
        StageTimer t(ctx, ST_STD);
        SCM_TRY(ws.alloc(&inv_sd, n));
        pooled_sd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(fs.ssq, fs.bcounts, Bt, n, fs.meta, d_sd, inv_sd);
        SCM_LAUNCH_CHECK(ctx);
        SCM_TRY(copy_async(ctx, d_mean, fs.ybar, sizeof(double) * n));
    }
    // centred Gram of ALL cells (PCA of Y / of any newY, the ruvK selection) and their column means
    if (d_gram_centred) SCM_TRY(copy_async(ctx, d_gram_centred, fs.Gm, sizeof(double) * n * n));
    if (d_colmean) SCM_TRY(copy_async(ctx, d_colmean, fs.ybar, sizeof(double) * n));
    // residual operator, part 2: Gram of Y0 = centred Gram - sum_g n_g (mu_g - ybar)(mu_g - ybar)', standardised when scRUVIII
    SCM_TRY(gram_residual(ctx, fs.Gm, n, fs.gsums, fs.gcounts, G, fs.ybar, inv_sd));
    int64_t* hmeta;
    SCM_TRY(fit_meta_to_host(ctx, fs, &hmeta));  // lands with the eigensolver's first read-back: no synchronisation of its own
    int rc;
    {   // runSVD (R/fastRUVIII.R:91) + t(u) %*% Y (:93) == sqrt(lambda) v'
        StageTimer t(ctx, ST_EIG);
        rc = eig_topk(ctx, fs.Gm, n, s, kconv, mode, tol, maxit, seed, vals, vecs, iters_out, resid_out, fs.meta);
        if (rc == SCM_ERR_NONFINITE || rc == 0 || rc == SCM_ERR_NOCONV) SCM_TRY(fit_meta_check(hmeta));
        if (rc != 0 && rc != SCM_ERR_NOCONV) return rc;
        dim3 grid((unsigned)ceil_div(n, 256), (unsigned)s);
        fullalpha_kernel<<<grid, 256, 0, st>>>(vals, vecs, s, n, d_fullalpha, d_sigma);
        SCM_LAUNCH_CHECK(ctx);
    }
    return rc;
}

int scm_ruv3_fit(scm_ctx* ctx, const double* d_Y, int64_t m_local, int64_t n, int64_t ld, const int32_t* d_group, int32_t G,
                 const int32_t* d_batch, int32_t Bt, int32_t s, int32_t kconv, int mode, double tol, int32_t maxit, uint64_t seed,
                 double* d_fullalpha, double* d_sigma, double* d_mean, double* d_sd, int32_t* iters_out, double* resid_out) {
    return ruv3_fit_impl(ctx, d_Y, m_local, n, ld, d_group, G, d_batch, Bt, s, kconv, mode, tol, maxit, seed, d_fullalpha, d_sigma,
                         d_mean, d_sd, iters_out, resid_out, nullptr, nullptr);
}
int scm_ruv3_fit_gram(scm_ctx* ctx, const double* d_Y, int64_t m_local, int64_t n, int64_t ld, const int32_t* d_group, int32_t G,
                      const int32_t* d_batch, int32_t Bt, int32_t s, int32_t kconv, int mode, double tol, int32_t maxit, uint64_t seed,
                      double* d_fullalpha, double* d_sigma, double* d_mean, double* d_sd, int32_t* iters_out, double* resid_out,
                      double* d_gram_centred, double* d_colmean) {
    return ruv3_fit_impl(ctx, d_Y, m_local, n, ld, d_group, G, d_batch, Bt, s, kconv, mode, tol, maxit, seed, d_fullalpha, d_sigma,
                         d_mean, d_sd, iters_out, resid_out, d_gram_centred, d_colmean);
}

int scm_pca_scores(scm_ctx* ctx, const double* d_gram_centred, int64_t n, int64_t m_total, const scm_coef* coef, const double* d_colmean,
                   int32_t rank, const double* d_Y, int64_t m_local, int64_t ld, double* d_scores, double* d_sing, double tol,
                   int32_t maxit, int32_t* iters_out, double* resid_out) {
    SCM_REQUIRE_CTX(ctx);
    return pca_scores(ctx, d_gram_centred, n, m_total, coef, d_colmean, rank, d_Y, m_local, ld, d_scores, d_sing, tol, maxit, iters_out,
                      resid_out);
}

int scm_silhouette(scm_ctx* ctx, const double* d_X, int64_t m, int32_t d, int64_t ldx, const int32_t* d_lab_a, int32_t Ca,
                   const int32_t* d_lab_b, int32_t Cb, int64_t i0, int64_t i1, double* h_sum_a, double* h_sum_b) {
    SCM_REQUIRE_CTX(ctx);
    return silhouette(ctx, d_X, m, d, ldx, d_lab_a, Ca, d_lab_b, Cb, i0, i1, h_sum_a, h_sum_b);
}

int scm_ruv3_host(scm_ctx* ctx, const double* h_Y, int64_t m, int64_t n, int64_t ldh, const int32_t* h_group, int32_t G,
                  const int32_t* h_batch, int32_t Bt, const uint8_t* h_ctl, int32_t k, int32_t s, int mode, double tol, int32_t maxit,
                  uint64_t seed, const double* h_fullalpha_in, int32_t s_in, int center_mode, const double* h_center, double* h_out,
                  int64_t ldo, double* h_fullalpha, double* h_sigma, double* h_mean, double* h_sd, int32_t* iters_out,
                  double* resid_out) {
    SCM_REQUIRE_CTX(ctx);
    HostArgs a{h_Y, m, n, ldh, h_group, G, h_batch, Bt, h_ctl, k, s, mode, tol, maxit, seed, h_fullalpha_in, s_in, center_mode,
               h_center, h_out, ldo, h_fullalpha, h_sigma, h_mean, h_sd, iters_out, resid_out};
    return ruv3_host(ctx, a);
}

// ---- replicate-finding numerics ------------------------------------------------------------------
int scm_gather(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, const int32_t* d_rows, int64_t msub,
               const int32_t* d_cols, int64_t nsub, double* d_out, int64_t ldo) {
    SCM_REQUIRE_CTX(ctx);
    return gather(ctx, d_Y, m, n, ld, d_rows, msub, d_cols, nsub, d_out, ldo);
}
int scm_group_median(scm_ctx* ctx, const double* d_X, int64_t m, int64_t n, int64_t ld, const int32_t* h_key, const int32_t* d_key,
                     int32_t G, double* d_med) {
    SCM_REQUIRE_CTX(ctx);
    return group_median(ctx, d_X, m, n, ld, h_key, d_key, G, d_med);
}
int scm_sqdist_to_center(scm_ctx* ctx, const double* d_X, int64_t m, int64_t n, int64_t ld, const int32_t* d_key, int32_t G,
                         const double* d_center, double* d_dist) {
    SCM_REQUIRE_CTX(ctx);
    return sqdist_to_center(ctx, d_X, m, n, ld, d_key, G, d_center, d_dist);
}
int scm_pair_dist(scm_ctx* ctx, const double* d_A, int64_t m1, int64_t lda, const double* d_B, int64_t m2, int64_t ldb, int64_t n,
                  int metric, double* d_D, int64_t ldd) {
    SCM_REQUIRE_CTX(ctx);
    return pair_dist(ctx, d_A, m1, lda, d_B, m2, ldb, n, metric, d_D, ldd);
}
int scm_block_median(scm_ctx* ctx, const double* d_D, int64_t ldd, const int64_t* h_row_off, int32_t K1, const int64_t* h_col_off,
                     int32_t K2, double* h_med) {
    SCM_REQUIRE_CTX(ctx);
    return block_median(ctx, d_D, ldd, h_row_off, K1, h_col_off, K2, h_med);
}
int scm_ls_coef(scm_ctx* ctx, const double* d_Y, int64_t m_local, int64_t n, int64_t ld, const double* d_W, int32_t k,
                const double* d_center, double* d_alpha) {
    SCM_REQUIRE_CTX(ctx);
    return ls_coef(ctx, d_Y, m_local, n, ld, d_W, k, d_center, d_alpha);
}
int scm_rankk_subtract(scm_ctx* ctx, const double* d_Y, int64_t m, int64_t n, int64_t ld, const double* d_W, int32_t k, const double* d_A,
                       double* d_out, int64_t ldo) {
    SCM_REQUIRE_CTX(ctx);
    return rankk_subtract(ctx, d_Y, m, n, ld, d_W, k, d_A, d_out, ldo);
}
int scm_kmeans(scm_ctx* ctx, const double* d_X, int64_t m, int32_t d, int64_t ldx, int32_t K, int32_t nstart, int32_t iter_max, uint64_t seed,
               int32_t* d_label, double* d_centers, double* h_withinss, int64_t* h_size, double* h_tot_withinss, int32_t* h_info) {
    SCM_REQUIRE_CTX(ctx);
    return kmeans(ctx, d_X, m, d, ldx, K, nstart, iter_max, seed, d_label, d_centers, h_withinss, h_size, h_tot_withinss, h_info);
}
int scm_centred_gram(scm_ctx* ctx, const double* d_Y, int64_t m_local, int64_t n, int64_t ld, double* d_gram, double* d_colmean,
                     int64_t* m_total_out) {
    SCM_REQUIRE_CTX(ctx);
    if (!d_Y || !d_gram || !d_colmean || n <= 0 || m_local < 0 || ld < n) return fail(SCM_ERR_INVALID, "scm_centred_gram: bad arguments");
    Scratch ws(ctx);
    int32_t* key;
    SCM_TRY(ws.alloc(&key, m_local > 0 ? m_local : 1));
    SCM_TRY(zero_async(ctx, key, sizeof(int32_t) * (m_local > 0 ? m_local : 1)));
    FitState fs;
    SCM_TRY(fit_state(ctx, ws, d_Y, m_local, n, ld, key, 1, nullptr, 0, fs));
    SCM_TRY(copy_async(ctx, d_gram, fs.Gm, sizeof(double) * n * n));
    SCM_TRY(copy_async(ctx, d_colmean, fs.ybar, sizeof(double) * n));
    int64_t* hmeta;
    SCM_TRY(fit_meta_to_host(ctx, fs, &hmeta));
    SCM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (hmeta[0] != 0) return fail(SCM_ERR_NONFINITE, "Y contains missing values or infinities.  This is not supported.");
    if (m_total_out) *m_total_out = hmeta[1];
    return 0;
}

int scm_multi_ruv3_host(scm_multi* mg, const double* h_Y, int64_t m, int64_t n, int64_t ldh, const int32_t* h_group, int32_t G,
                        const int32_t* h_batch, int32_t Bt, const uint8_t* h_ctl, int32_t k, int32_t s, int mode, double tol, int32_t maxit,
                        uint64_t seed, const double* h_fullalpha_in, int32_t s_in, int center_mode, const double* h_center, double* h_out,
                        int64_t ldo, double* h_fullalpha, double* h_sigma, double* h_mean, double* h_sd, int32_t* iters_out,
                        double* resid_out) {
    if (!mg || mg->ctx.empty()) return fail(SCM_ERR_INVALID, "scm_multi_ruv3_host: null handle");
    const int nd = (int)mg->ctx.size();
    if (m < nd) return fail(SCM_ERR_INVALID, "scm_multi_ruv3_host: fewer cells (%lld) than devices (%d)", (long long)m, nd);
    const int64_t base = m / nd, rem = m % nd;
    return multi_run(mg, [&](int r) -> int {
        scm_ctx* ctx = mg->ctx[r];
        SCM_CUDA(cudaSetDevice(ctx->device));
        const int64_t r0 = r * base + (r < rem ? r : rem), rows = base + (r < rem ? 1 : 0);  // contiguous, balanced, never empty
        const bool lead = r == 0;  // after the Gram all-reduce every device holds bit-identical fit results: device 0 reports them
        HostArgs a{h_Y + r0 * ldh, rows, n, ldh, h_group ? h_group + r0 : nullptr, G, h_batch ? h_batch + r0 : nullptr, Bt, h_ctl, k, s, mode,
                   tol, maxit, seed, h_fullalpha_in, s_in, center_mode, h_center, h_out + r0 * ldo, ldo, lead ? h_fullalpha : nullptr,
                   lead ? h_sigma : nullptr, lead ? h_mean : nullptr, lead ? h_sd : nullptr, lead ? iters_out : nullptr,
                   lead ? resid_out : nullptr};
        return ruv3_host(ctx, a);
    });
}

int scm_timers(scm_ctx* ctx, double* ms, int n) {
    SCM_REQUIRE_CTX(ctx);
    SCM_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; i++) {
        ms[i] = 0.0;
        if (i < ST_COUNT && ctx->ev_used[i]) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, ctx->ev0[i], ctx->ev1[i]) == cudaSuccess) ms[i] = t;
        }
    }
    return 0;
}

}  // extern "C"
