// End-to-end driver for HOST matrices: RUV-III with the PCIe transfers overlapped with the kernels.
//
//   upload  : Y arrives in row chunks on a copy stream; as soon as a chunk has landed the compute stream adds its
//             replicate-group sums (K1) and its shifted Gram (K3, accumulate) -- the fit streams behind the H2D copy
//             (the Gram identity of DESIGN.md section 2 holds for ANY fixed shift, so the shift is the column mean of
//             the first chunk and no second pass over Y is needed),
//   solve   : Gram correction, top-s eigenpairs, coefficients (small, latency bound),
//   download: the fused adjust kernel (K6) runs chunk by chunk IN PLACE and each finished chunk goes back over
//             PCIe while the next one is being adjusted.
//   A matrix that does not fit the device is streamed TWICE through three chunk buffers instead (out-of-core, see ruv3_host).
// Two things keep the overlap real (both found with SCM_HOST_TRACE=1, which prints the per-chunk device timeline):
//   * nothing that a copy engine might execute (cudaMemsetAsync, small cudaMemcpyAsync) is issued on the compute stream
//     while bulk copies are in flight, and the compute stream is moved back to the compute engine by a trivial kernel
//     before the first bulk copy: otherwise its next operation is arbitrated behind ALL queued bulk copies,
//   * uploads and downloads use separate copy streams, and chunks are handed over by the host (event wait) with at most
//     two bulk copies queued.
// Included by scm_api.cu after the small helper kernels (meta_kernel, colmean_kernel, ...).
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>

#include "adjust.cuh"
#include "common.cuh"
#include "eig_topk.cuh"
#include "fit_core.cuh"
#include "gram_tma.cuh"
#include "seg_colsum.cuh"

namespace scm {

__global__ void add_f64_kernel(double* __restrict__ dst, const double* __restrict__ src, int64_t count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dst[i] += src[i];
}
__global__ void add_i64_kernel(int64_t* __restrict__ dst, const int64_t* __restrict__ src, int64_t count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dst[i] += src[i];
}
// shift[j] = sum_g sums[g, j] / rows: column mean of the first chunk (every cell of a fit carries a group key)
__global__ void first_chunk_mean_kernel(const double* __restrict__ sums, int32_t G, int64_t n, double rows, double* __restrict__ shift) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double s = 0.0;
    for (int g = 0; g < G; g++) s += sums[(int64_t)g * n + j];
    shift[j] = rows > 0 ? s / rows : 0.0;
}

// Rows per PCIe chunk: ~512 MB, a multiple of 1024 rows (test hook SCM_HOST_CHUNK_ROWS forces several chunks on small inputs).
inline int64_t host_chunk_rows(int64_t ld) {
    int64_t rows_per_chunk = ((int64_t)512 << 20) / (ld * 8);
    rows_per_chunk = rows_per_chunk / 1024 * 1024;
    if (rows_per_chunk < 4096) rows_per_chunk = 4096;
    if (const char* e = getenv("SCM_HOST_CHUNK_ROWS")) {
        const long v = atol(e);
        if (v >= 16) rows_per_chunk = v;
    }
    return rows_per_chunk;
}

// ---- pageable host matrices ------------------------------------------------------------------------------------------
// cudaMemcpyAsync from / to ordinary (pageable) memory is staged by the driver through ONE internal buffer, synchronously and
// single-threaded: 7-12 GB/s instead of the 55 GB/s of the link, and nothing overlaps (measured: 2.2 s instead of 0.58 s per
// 1M x 2000 fastRUVIII call).  An R session's REAL(Y) is such memory.  The host drivers therefore detect pageable arguments and
// stage them themselves: a helper thread copies chunk c + 1 into a page-locked bounce buffer with several memcpy threads while
// chunk c crosses PCIe and chunk c - 1 is being summed; downloads mirror that.  Page-locked callers take the direct path.
constexpr int STAGE_NB = 4;  // bounce buffers: 2 for the upload ring + 2 for the download ring (used one direction at a time: all 4)

inline bool host_is_pageable(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

inline int stage_reserve(scm_ctx* ctx, size_t bytes) {
    if (ctx->stage_bytes >= bytes) return 0;
    for (void*& b : ctx->stage_buf) { if (b) cudaFreeHost(b); b = nullptr; }
    ctx->stage_bytes = 0;
    for (void*& b : ctx->stage_buf) SCM_CUDA(cudaMallocHost(&b, bytes));
    ctx->stage_bytes = bytes;
    return 0;
}

inline int host_copy_threads() {
    int t = 8;
    if (const char* e = getenv("SCM_HOST_COPY_THREADS")) t = atoi(e);
    const int hw = (int)std::thread::hardware_concurrency();
    if (hw > 0 && t > hw) t = hw;
    return t < 1 ? 1 : t;
}

// rows x row_bytes strided host copy on `nt` threads (memcpy saturates one core long before it saturates the memory system)
inline void parallel_copy2d(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t row_bytes, int64_t rows, int nt) {
    if (rows <= 0) return;
    auto body = [=](int64_t r0, int64_t r1) {
        if (dst_pitch == row_bytes && src_pitch == row_bytes) {
            memcpy((char*)dst + r0 * row_bytes, (const char*)src + r0 * row_bytes, (size_t)(r1 - r0) * row_bytes);
        } else {
            for (int64_t r = r0; r < r1; r++) memcpy((char*)dst + r * dst_pitch, (const char*)src + r * src_pitch, row_bytes);
        }
    };
    if (nt <= 1 || rows < 64) { body(0, rows); return; }
    std::vector<std::thread> th;
    const int64_t per = (rows + nt - 1) / nt;
    for (int t = 0; t < nt; t++) {
        const int64_t r0 = t * per, r1 = r0 + per < rows ? r0 + per : rows;
        if (r0 >= r1) break;
        th.emplace_back(body, r0, r1);
    }
    for (auto& x : th) x.join();
}

// Hand-over counter between the staging helper thread and the driver (chunks published in order).
struct ChunkGate {
    std::mutex mu;
    std::condition_variable cv;
    int ready = 0;
    int error = 0;
    char msg[256] = "";
    void publish(int upto) { { std::lock_guard<std::mutex> g(mu); ready = upto; } cv.notify_all(); }
    void fail_with(int rc, const char* why) {
        { std::lock_guard<std::mutex> g(mu); if (!error) { error = rc; snprintf(msg, sizeof(msg), "%.250s", why); } }
        cv.notify_all();
    }
    int wait_for(int c) {  // returns 0 once chunk c has been published, the helper's status if it died
        std::unique_lock<std::mutex> g(mu);
        cv.wait(g, [&] { return ready > c || error != 0; });
        return error;
    }
};
// joins the helper on every exit path; `gate` is failed first so that a helper blocked on it wakes up and leaves
struct HelperJoin {
    std::thread& t;
    ChunkGate* wake;
    ~HelperJoin() {
        if (wake) wake->fail_with(SCM_ERR_INVALID, "driver left early");
        if (t.joinable()) t.join();
    }
};
#define SCM_HELPER_CUDA(expr, gate)                                                                    \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) { (gate).fail_with(SCM_ERR_CUDA, cudaGetErrorString(_e)); return; }     \
    } while (0)

// ---- chunk movers ---------------------------------------------------------------------------------------------------------
// Both directions exist in a direct form (page-locked host memory: cudaMemcpyAsync straight from / to the caller's buffer, issued by
// the driver thread) and a staged form (pageable memory: a helper thread + memcpy threads + the bounce buffers).  A "slot" is
// where a chunk lives on the device: its own place inside the resident matrix (in-core) or one of a small ring of chunk buffers
// that are re-used (out-of-core); a re-used slot is guarded by the event the driver records when its previous tenant is done.
struct ChunkPlan {
    int64_t m, n, ld, rows_per_chunk;
    int nchunks;
    int64_t row0(int c) const { return (int64_t)c * rows_per_chunk; }
    int64_t rows(int c) const { return row0(c) + rows_per_chunk <= m ? rows_per_chunk : m - row0(c); }
};

struct Uploader {
    scm_ctx* ctx;
    ChunkPlan pl;
    const double* h_src;
    int64_t ldh;
    cudaStream_t cs;
    bool staged;
    int slots_in_flight;            // ring size of the device slots (0: every chunk has its own place)
    int stage_lo, stage_n;          // bounce buffers this direction may use: stage_buf[stage_lo .. stage_lo + stage_n)
    std::vector<double*> dst;       // device address of chunk c
    std::vector<cudaEvent_t> ev;    // chunk c has landed
    std::vector<cudaEvent_t> freed; // recorded by the driver when the tenant of chunk c's slot is done (out-of-core only)
    ChunkGate fed, released;        // staged: chunks published by the helper / slots released by the driver
    std::thread feeder;
    int issued = 0;                 // direct mode: copies issued so far

    ~Uploader() {
        fed.fail_with(SCM_ERR_INVALID, "driver left early");
        released.fail_with(SCM_ERR_INVALID, "driver left early");
        if (feeder.joinable()) feeder.join();
        cudaStreamSynchronize(cs);
        for (auto e : ev) if (e) cudaEventDestroy(e);
        for (auto e : freed) if (e) cudaEventDestroy(e);
    }
    int init() {
        ev.assign(pl.nchunks, nullptr);
        for (auto& e : ev) SCM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        if (slots_in_flight > 0) {
            freed.assign(pl.nchunks, nullptr);
            for (auto& e : freed) SCM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        return 0;
    }
    int copy_chunk(int c, const double* src, int64_t src_ld, cudaStream_t stream) {
        const int64_t rows = pl.rows(c), n = pl.n, ld = pl.ld;
        if (slots_in_flight > 0 && c >= slots_in_flight) SCM_CUDA(cudaStreamWaitEvent(stream, freed[c - slots_in_flight], 0));
        if (src_ld == n && ld == n) SCM_CUDA(cudaMemcpyAsync(dst[c], src, (size_t)rows * n * 8, cudaMemcpyHostToDevice, stream));
        else SCM_CUDA(cudaMemcpy2DAsync(dst[c], ld * 8, src, src_ld * 8, n * 8, rows, cudaMemcpyHostToDevice, stream));
        SCM_CUDA(cudaEventRecord(ev[c], stream));
        return 0;
    }
    // direct mode: chunk c may be issued once the driver has recorded the release of chunk c - slots_in_flight
    int issue_direct_upto(int last) {
        for (; issued <= last && issued < pl.nchunks; issued++)
            SCM_TRY(copy_chunk(issued, h_src + pl.row0(issued) * ldh, ldh, cs));
        return 0;
    }
    int start() {
        if (!staged) return issue_direct_upto(slots_in_flight > 0 ? (slots_in_flight < 2 ? slots_in_flight : 2) - 1 : 1);
        const int nt = host_copy_threads();
        const int dev = ctx->device;
        feeder = std::thread([this, nt, dev] {
            cudaSetDevice(dev);
            for (int c = 0; c < pl.nchunks; c++) {
                if (fed.error) return;
                if (c >= stage_n) SCM_HELPER_CUDA(cudaEventSynchronize(ev[c - stage_n]), fed);  // the bounce buffer's previous H2D has finished
                double* sb = (double*)ctx->stage_buf[stage_lo + c % stage_n];
                parallel_copy2d(sb, (size_t)pl.n * 8, h_src + pl.row0(c) * ldh, (size_t)ldh * 8, (size_t)pl.n * 8, pl.rows(c), nt);
                if (slots_in_flight > 0 && c >= slots_in_flight && released.wait_for(c - slots_in_flight) != 0) return;  // event recorded?
                if (copy_chunk(c, sb, pl.n, cs) != 0) { fed.fail_with(SCM_ERR_CUDA, g_err); return; }
                fed.publish(c + 1);
            }
        });
        return 0;
    }
    // Returns when chunk c is in HBM.  Direct mode keeps two copies queued behind it.
    int wait(int c) {
        if (staged && fed.wait_for(c) != 0) return fail(fed.error, "upload helper: %s", fed.msg);
        SCM_CUDA(cudaEventSynchronize(ev[c]));
        return 0;
    }
    // The driver is done with the device slot of chunk c (everything that reads or writes it has been enqueued on `stream`).
    int release(int c, cudaStream_t stream) {
        if (slots_in_flight > 0) {
            SCM_CUDA(cudaEventRecord(freed[c], stream));
            if (staged) released.publish(c + 1);
        }
        if (!staged) SCM_TRY(issue_direct_upto(c + (slots_in_flight > 0 ? slots_in_flight : 2)));
        return 0;
    }
    int finish() {
        if (staged && feeder.joinable()) {
            if (fed.wait_for(pl.nchunks - 1) != 0) return fail(fed.error, "upload helper: %s", fed.msg);
            feeder.join();
        }
        return 0;
    }
};

struct Downloader {
    scm_ctx* ctx;
    ChunkPlan pl;
    double* h_dst;
    int64_t ldo;
    cudaStream_t cs;
    bool staged;
    int stage_lo, stage_n;
    std::vector<cudaEvent_t> evd;   // chunk c has left the device
    ChunkGate issued, drained;
    std::thread drainer;

    ~Downloader() {
        issued.fail_with(SCM_ERR_INVALID, "driver left early");
        if (drainer.joinable()) drainer.join();
        cudaStreamSynchronize(cs);
        for (auto e : evd) if (e) cudaEventDestroy(e);
    }
    int init() {
        evd.assign(pl.nchunks, nullptr);
        for (auto& e : evd) SCM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        if (!staged) return 0;
        const int nt = host_copy_threads();
        const int dev = ctx->device;
        drainer = std::thread([this, nt, dev] {
            cudaSetDevice(dev);
            for (int c = 0; c < pl.nchunks; c++) {
                if (issued.wait_for(c) != 0) return;
                SCM_HELPER_CUDA(cudaEventSynchronize(evd[c]), drained);
                parallel_copy2d(h_dst + pl.row0(c) * ldo, (size_t)ldo * 8, ctx->stage_buf[stage_lo + c % stage_n], (size_t)pl.n * 8,
                                (size_t)pl.n * 8, pl.rows(c), nt);
                drained.publish(c + 1);
            }
        });
        return 0;
    }
    // Chunk c (at d_src, leading dimension ld) goes home once `ready` (recorded on the compute stream) has fired.
    int push(int c, const double* d_src, cudaEvent_t ready) {
        const int64_t rows = pl.rows(c), n = pl.n, ld = pl.ld;
        SCM_CUDA(cudaStreamWaitEvent(cs, ready, 0));
        if (staged) {
            if (c >= stage_n && drained.wait_for(c - stage_n) != 0) return fail(drained.error, "download helper: %s", drained.msg);
            double* sb = (double*)ctx->stage_buf[stage_lo + c % stage_n];
            if (ld == n) SCM_CUDA(cudaMemcpyAsync(sb, d_src, (size_t)rows * n * 8, cudaMemcpyDeviceToHost, cs));
            else SCM_CUDA(cudaMemcpy2DAsync(sb, n * 8, d_src, ld * 8, n * 8, rows, cudaMemcpyDeviceToHost, cs));
        } else if (ldo == n && ld == n) {
            SCM_CUDA(cudaMemcpyAsync(h_dst + pl.row0(c) * ldo, d_src, (size_t)rows * n * 8, cudaMemcpyDeviceToHost, cs));
        } else {
            SCM_CUDA(cudaMemcpy2DAsync(h_dst + pl.row0(c) * ldo, ldo * 8, d_src, ld * 8, n * 8, rows, cudaMemcpyDeviceToHost, cs));
        }
        SCM_CUDA(cudaEventRecord(evd[c], cs));
        if (staged) issued.publish(c + 1);
        return 0;
    }
    int finish() {
        if (staged) {
            if (drained.wait_for(pl.nchunks - 1) != 0) return fail(drained.error, "download helper: %s", drained.msg);
            drainer.join();
        }
        SCM_CUDA(cudaStreamSynchronize(cs));
        return 0;
    }
};

// The download half of the host drivers for a RESIDENT matrix: the fused adjust kernel (K6) runs chunk by chunk IN PLACE and every
// finished chunk goes back over PCIe (copy stream) while the next one is being adjusted.  d_Y holds newY afterwards.  Returns
// after the last byte has landed in h_out.
inline int adjust_download(scm_ctx* ctx, double* dY, int64_t m, int64_t n, int64_t ld, const scm_coef* coef, double* h_out,
                           int64_t ldo, int64_t rows_per_chunk) {
    if (!dY || !h_out || !coef || ldo < n) return fail(SCM_ERR_INVALID, "adjust_download: bad arguments");
    cudaStream_t st = ctx->stream;
    if (!ctx->copy_stream_d2h) SCM_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream_d2h, cudaStreamNonBlocking));
    if (rows_per_chunk <= 0) rows_per_chunk = host_chunk_rows(ld);
    ChunkPlan pl{m, n, ld, rows_per_chunk, (int)ceil_div(m, rows_per_chunk)};
    const bool staged = host_is_pageable(h_out);
    if (staged) SCM_TRY(stage_reserve(ctx, (size_t)rows_per_chunk * n * 8));
    std::vector<cudaEvent_t> ev(pl.nchunks);
    for (auto& e : ev) SCM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    struct EvGuard {
        std::vector<cudaEvent_t>& v;
        ~EvGuard() { for (auto e : v) cudaEventDestroy(e); }
    } guard{ev};
    Downloader down{ctx, pl, h_out, ldo, ctx->copy_stream_d2h, staged, 0, STAGE_NB};
    SCM_TRY(down.init());
    for (int c = 0; c < pl.nchunks; c++) {
        double* dc = dY + pl.row0(c) * ld;
        {
            StageTimer t(ctx, ST_ADJUST);
            SCM_TRY(adjust(ctx, dc, pl.rows(c), n, ld, coef, nullptr, 0, dc, ld, nullptr));
        }
        SCM_CUDA(cudaEventRecord(ev[c], st));
        SCM_TRY(down.push(c, dc, ev[c]));
    }
    // the host waits for the downloads: stream-ordered frees of d_Y issued after this function returns are safe
    SCM_TRY(down.finish());
    SCM_CUDA(cudaStreamSynchronize(st));
    return 0;
}

struct HostArgs {
    const double* h_Y; int64_t m, n, ldh;
    const int32_t* h_group; int32_t G;
    const int32_t* h_batch; int32_t Bt;
    const uint8_t* h_ctl; int32_t k, s; int mode; double tol; int32_t maxit; uint64_t seed;
    const double* h_fullalpha_in; int32_t s_in;  // adjust-only when given
    int center_mode;                             // adjust-only: 0 none, 1 column means of Y, 2 h_center
    const double* h_center;
    double* h_out; int64_t ldo;
    double *h_fullalpha, *h_sigma, *h_mean, *h_sd;
    int32_t* iters_out; double* resid_out;
};

// SCM_HOST_TRACE=1: wall-clock phase marks of the host driver on stderr (diagnostics only)
struct HostTrace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    HostTrace() : on(getenv("SCM_HOST_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* what) const {
        if (!on) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "[scm_ruv3_host] %8.2f ms  %s\n", ms, what);
    }
};

// Out-of-core standardize2: the squared deviations have to be summed in the same pass that sums the batch totals, i.e. before the
// (global) batch means mu_b exist, so they are taken about the rank's shift a and moved afterwards.  For the LOCAL cells of batch b,
//     sum (y - mu_b)^2 = sum (y - a)^2 - 2 (mu_b - a) (s_loc - n_loc a) + n_loc (mu_b - a)^2,
// s_loc / n_loc the local batch sum / size (copies taken before the all-reduce).  |mu_b - a| is a batch effect, not a mean: nothing
// cancels beyond the ratio of between-batch to within-batch variance.
__global__ void ssq_finish_kernel(double* __restrict__ ssq, const double* __restrict__ bmeans, const double* __restrict__ s_loc,
                                  const int64_t* __restrict__ n_loc, int32_t Bt, int64_t n, const double* __restrict__ a) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    for (int b = 0; b < Bt; b++) {
        const double nl = (double)n_loc[b], dlt = bmeans[(int64_t)b * n + j] - a[j];
        ssq[(int64_t)b * n + j] += -2.0 * dlt * (s_loc[(int64_t)b * n + j] - nl * a[j]) + nl * dlt * dlt;
    }
}
// dst[b, :] = a for every batch b (the per-group centre argument of the squared-deviation K1 pass)
__global__ void tile_rows_kernel(double* __restrict__ dst, const double* __restrict__ a, int32_t Bt, int64_t n) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    for (int b = 0; b < Bt; b++) dst[(int64_t)b * n + j] = a[j];
}

// Device memory the resident (in-core) form may take: SCM_HOST_DEVICE_BUDGET bytes, else 85 % of what is free right now.
inline int64_t host_device_budget() {
    if (const char* e = getenv("SCM_HOST_DEVICE_BUDGET")) return atoll(e);
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); return (int64_t)1 << 62; }
    return (int64_t)((double)fr * 0.85);
}

// The whole closure on a host matrix.  Two forms, same kernels and the same results up to summation order:
//   in-core     : Y is uploaded once into a resident device matrix (chunks arrive on a copy stream while the fit accumulates behind
//                 them), adjusted in place and downloaded chunk by chunk.  PCIe: 8 m n up + 8 m n down, one after the other
//                 (the fit needs every cell before the first adjusted byte exists).
//   out-of-core : when the matrix does not fit the device budget, Y is STREAMED TWICE through a ring of three chunk buffers:
//                 pass 1 sums and accumulates the Gram, pass 2 re-uploads every chunk, adjusts it and sends it home, upload and
//                 download running concurrently (full duplex).  PCIe: 16 m n up + 8 m n down, but the second upload hides
//                 under the download, so the wall time matches the in-core form; device memory: 3 chunks (1.5 GB) + n^2.
//                 This is the chunked `bplapply` adjust of pseudoRUVIII (R/scMerge2_utilis.R:104-124) and the HDF5 / DelayedArray
//                 use case of the reference's vignette (vignettes/scMerge.Rmd:401-441) for a matrix that lives in host memory.
inline int ruv3_host(scm_ctx* ctx, const HostArgs& a) {
    const HostTrace trace;
    const int64_t m = a.m, n = a.n;
    if (!a.h_Y || !a.h_out || m <= 0 || n <= 0 || a.ldh < n || a.ldo < n || !a.h_ctl) return fail(SCM_ERR_INVALID, "scm_ruv3_host: bad arguments");
    const bool do_fit = a.h_fullalpha_in == nullptr;
    const bool sc = do_fit && a.h_batch != nullptr;
    if (do_fit && (!a.h_group || a.G <= 0 || a.s <= 0)) return fail(SCM_ERR_INVALID, "scm_ruv3_host: fit requested without groups / s");
    cudaStream_t st = ctx->stream;
    if (!ctx->copy_stream) SCM_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!ctx->copy_stream_d2h) SCM_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream_d2h, cudaStreamNonBlocking));
    const int64_t ld = (n + 1) & ~(int64_t)1;
    const int64_t rows_per_chunk = host_chunk_rows(ld);
    ChunkPlan pl{m, n, ld, rows_per_chunk, (int)ceil_div(m, rows_per_chunk)};
    const int nchunks = pl.nchunks;
    const bool need_sums = do_fit || a.center_mode == 1;
    constexpr int RING = 3;
    // an adjust-only call that needs no pass over the cells first (fullalpha given, centre none or given: getAdjustedMat with the
    // stored adjusted_means, fastRUVIII(fullalpha = )) ALWAYS streams: one pass, every chunk goes up, is adjusted and comes home
    // while the next ones are in flight in both directions -- half the PCIe wall time of upload-then-download
    const bool stream_mode = nchunks > RING && (getenv("SCM_HOST_STREAM") ? atoi(getenv("SCM_HOST_STREAM")) != 0
                                                                           : (!need_sums || m * ld * 8 + n * n * 32 > host_device_budget()));
    const bool staged_in = host_is_pageable(a.h_Y), staged_out = host_is_pageable(a.h_out);
    if (staged_in || staged_out) SCM_TRY(stage_reserve(ctx, (size_t)rows_per_chunk * n * 8));

    Scratch ws(ctx);
    // declared after `ws`, so destroyed before it: on EVERY exit path (errors included) the bulk copies still queued on the copy
    // streams have finished before Scratch hands the device matrix back to the pool
    struct CopyFence {
        scm_ctx* c;
        ~CopyFence() { if (c->copy_stream) cudaStreamSynchronize(c->copy_stream); if (c->copy_stream_d2h) cudaStreamSynchronize(c->copy_stream_d2h); }
    } fence{ctx};
    double *dY, *gsums = nullptr, *tsums = nullptr, *shift = nullptr, *Gm = nullptr, *vals = nullptr, *vecs = nullptr, *fa = nullptr;
    double *bsums = nullptr, *ssq = nullptr, *sd = nullptr, *inv_sd = nullptr, *center = nullptr, *sig = nullptr;
    double *buf1 = nullptr, *buf2 = nullptr, *cnt_f64 = nullptr, *uvec = nullptr, *dvec = nullptr, *tb = nullptr, *atile = nullptr;
    int64_t *gcounts = nullptr, *gcnt_local = nullptr, *tcounts = nullptr, *bcounts = nullptr, *bcnt_local = nullptr, *meta = nullptr;
    int32_t *dgroup = nullptr, *dbatch = nullptr, *flag = nullptr;
    uint8_t* dctl;
    const int32_t s = do_fit ? a.s : a.s_in;
    const int32_t Gk = do_fit ? a.G : 1;  // adjust-only runs use one group (column means)
    const int32_t Btk = sc ? a.Bt : 0;
    // the two all-reduce buffers of fit_core.cuh: [gsums | bsums | counts | flag, rows] and [Gm | ssq]
    const int64_t len1 = (int64_t)(Gk + Btk) * n + Gk + Btk + 2, len2 = (do_fit ? n * n : 0) + (int64_t)Btk * n;
    SCM_TRY(ws.alloc(&dY, stream_mode ? (int64_t)RING * rows_per_chunk * ld : m * ld));
    SCM_TRY(ws.alloc(&dctl, n));
    SCM_TRY(ws.alloc(&fa, (int64_t)s * n));
    SCM_TRY(ws.alloc(&meta, 4)); SCM_TRY(ws.alloc(&flag, 1));
    SCM_TRY(ws.alloc(&shift, n)); SCM_TRY(ws.alloc(&center, n)); SCM_TRY(ws.alloc(&uvec, n)); SCM_TRY(ws.alloc(&dvec, n));
    SCM_TRY(ws.alloc(&buf1, len1)); SCM_TRY(ws.alloc(&buf2, len2 > 0 ? len2 : 1));
    gsums = buf1; bsums = buf1 + (int64_t)Gk * n; cnt_f64 = buf1 + (int64_t)(Gk + Btk) * n;
    Gm = buf2; ssq = buf2 + (do_fit ? n * n : 0);
    SCM_CUDA(cudaMemsetAsync(flag, 0, 4, st));
    SCM_CUDA(cudaMemsetAsync(shift, 0, sizeof(double) * n, st));
    SCM_CUDA(cudaMemsetAsync(buf1, 0, sizeof(double) * len1, st));
    if (len2 > 0 && stream_mode) SCM_CUDA(cudaMemsetAsync(buf2, 0, sizeof(double) * len2, st));
    SCM_CUDA(cudaMemcpyAsync(dctl, a.h_ctl, n, cudaMemcpyHostToDevice, st));
    SCM_TRY(ws.alloc(&tsums, (int64_t)(Gk > Btk ? Gk : Btk) * n));
    SCM_TRY(ws.alloc(&gcounts, Gk)); SCM_TRY(ws.alloc(&gcnt_local, Gk)); SCM_TRY(ws.alloc(&tcounts, Gk > Btk ? Gk : Btk));
    SCM_TRY(ws.alloc(&bcounts, Btk > 0 ? Btk : 1)); SCM_TRY(ws.alloc(&bcnt_local, Btk > 0 ? Btk : 1));
    SCM_TRY(ws.alloc(&dgroup, m));
    SCM_CUDA(cudaMemsetAsync(gcnt_local, 0, sizeof(int64_t) * Gk, st));
    SCM_CUDA(cudaMemsetAsync(bcnt_local, 0, sizeof(int64_t) * (Btk > 0 ? Btk : 1), st));
    if (do_fit) {
        SCM_TRY(ws.alloc(&vals, s)); SCM_TRY(ws.alloc(&vecs, (int64_t)s * n)); SCM_TRY(ws.alloc(&sig, s));
        SCM_CUDA(cudaMemcpyAsync(dgroup, a.h_group, sizeof(int32_t) * m, cudaMemcpyHostToDevice, st));
    } else {
        SCM_CUDA(cudaMemsetAsync(dgroup, 0, sizeof(int32_t) * m, st));
        SCM_CUDA(cudaMemcpyAsync(fa, a.h_fullalpha_in, sizeof(double) * s * n, cudaMemcpyHostToDevice, st));
    }
    if (sc) {
        SCM_TRY(ws.alloc(&dbatch, m)); SCM_TRY(ws.alloc(&sd, n)); SCM_TRY(ws.alloc(&inv_sd, n));
        SCM_CUDA(cudaMemcpyAsync(dbatch, a.h_batch, sizeof(int32_t) * m, cudaMemcpyHostToDevice, st));
        if (stream_mode) { SCM_TRY(ws.alloc(&tb, (int64_t)Btk * n)); SCM_TRY(ws.alloc(&atile, (int64_t)Btk * n)); }
    }
    // The small uploads / memsets above are copy-engine work of the COMPUTE stream.  While its channel still targets the copy
    // engine, its next operation (even an event record) is arbitrated behind every bulk copy of the copy stream: measured with
    // SCM_HOST_TRACE=1, the whole fit then only started when the last H2D chunk had landed (+144 ms at 1M x 2000).  A trivial
    // kernel moves the channel back to the compute engine, and the host waits for it before the first bulk copy is issued
    // (this also makes the pool allocations visible to the copy stream).
    SCM_TRY(zero_async(ctx, flag, 4));
    SCM_CUDA(cudaStreamSynchronize(st));
    trace.mark(stream_mode ? "allocated (out-of-core: ring of 3 chunk buffers, two passes)" : "allocated (in-core)");

    auto slot = [&](int c) -> double* { return stream_mode ? dY + (int64_t)(c % RING) * rows_per_chunk * ld : dY + pl.row0(c) * ld; };
    // ---- pass 1: upload, with the sums and the Gram accumulating behind it
    if (need_sums || !stream_mode) {
        Uploader up{ctx, pl, a.h_Y, a.ldh, ctx->copy_stream, staged_in, stream_mode ? RING : 0, 0, STAGE_NB};
        up.dst.resize(nchunks);
        for (int c = 0; c < nchunks; c++) up.dst[c] = slot(c);
        SCM_TRY(up.init());
        SCM_TRY(up.start());
        for (int c = 0; c < nchunks; c++) {
            const int64_t r0 = pl.row0(c), rows = pl.rows(c);
            double* dc = slot(c);
            // host-mediated hand-over: chunk c is in HBM when wait() returns, its fit work is enqueued right away and the next
            // copies join the queue (the copy engine never idles, the queue never grows)
            SCM_TRY(up.wait(c));
            if (need_sums) {
                StageTimer t(ctx, ST_SEG);
                SCM_TRY(seg_colsum(ctx, dc, rows, n, ld, dgroup + r0, Gk, nullptr, tsums, tcounts, flag));
                add_f64_kernel<<<(unsigned)ceil_div((int64_t)Gk * n, 256), 256, 0, st>>>(gsums, tsums, (int64_t)Gk * n);
                SCM_LAUNCH_CHECK(ctx);
                add_i64_kernel<<<(unsigned)ceil_div(Gk, 256), 256, 0, st>>>(gcnt_local, tcounts, Gk);
                SCM_LAUNCH_CHECK(ctx);
            }
            if (do_fit) {
                if (c == 0) {
                    // Gram shift = column mean of this rank's FIRST chunk: known before the second chunk lands, no agreement
                    // between ranks needed (the local Gram is moved to the global mean by a rank-2 update before its all-reduce)
                    first_chunk_mean_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(tsums, Gk, n, (double)rows, shift);
                    SCM_LAUNCH_CHECK(ctx);
                    if (sc && stream_mode) {
                        tile_rows_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(atile, shift, Btk, n);
                        SCM_LAUNCH_CHECK(ctx);
                    }
                }
                if (sc && stream_mode) {
                    // standardize2 out of core: batch sums AND squared deviations about the shift in this one pass (the batch means
                    // they should be taken about do not exist yet; ssq_recentre_kernel moves them afterwards)
                    StageTimer t(ctx, ST_STD);
                    SCM_TRY(seg_colsum(ctx, dc, rows, n, ld, dbatch + r0, Btk, nullptr, tb, tcounts, nullptr));
                    add_f64_kernel<<<(unsigned)ceil_div((int64_t)Btk * n, 256), 256, 0, st>>>(bsums, tb, (int64_t)Btk * n);
                    SCM_LAUNCH_CHECK(ctx);
                    add_i64_kernel<<<(unsigned)ceil_div(Btk, 256), 256, 0, st>>>(bcnt_local, tcounts, Btk);
                    SCM_LAUNCH_CHECK(ctx);
                    SCM_TRY(seg_colsum(ctx, dc, rows, n, ld, dbatch + r0, Btk, atile, tb, nullptr, nullptr));
                    add_f64_kernel<<<(unsigned)ceil_div((int64_t)Btk * n, 256), 256, 0, st>>>(ssq, tb, (int64_t)Btk * n);
                    SCM_LAUNCH_CHECK(ctx);
                }
                StageTimer t(ctx, ST_GRAM);
                SCM_TRY(gram(ctx, dc, rows, n, ld, shift, Gm, c == nchunks - 1, c > 0));
            }
            SCM_TRY(up.release(c, st));
        }
        SCM_TRY(up.finish());
    }
    trace.mark("pass 1 enqueued (upload + sums + Gram)");
    int rc = 0;
    scm_coef* coef = nullptr;
    if (need_sums) {
        // all-reduce 1 (one float64 buffer): group sums, batch sums, counts, NA/Inf flag, rows -> global column means
        if (sc && !stream_mode) {  // standardize2, pass 1 over the resident cells (R/scRUVIII.R:163-166)
            StageTimer t(ctx, ST_STD);
            SCM_TRY(seg_colsum(ctx, dY, m, n, ld, dbatch, Btk, nullptr, bsums, bcnt_local, nullptr));
        }
        if (sc && stream_mode) SCM_TRY(copy_async(ctx, tb, bsums, sizeof(double) * Btk * n));  // local batch sums, for ssq_finish_kernel
        const int64_t span = n > Gk + Btk ? n : Gk + Btk;
        fit_pack_kernel<<<(unsigned)ceil_div(span, 256), 256, 0, st>>>(gsums, Gk, n, gcnt_local, bcnt_local, Btk, flag, m, shift, cnt_f64, uvec);
        SCM_LAUNCH_CHECK(ctx);
        SCM_TRY(allreduce(ctx, buf1, len1, 0));
        fit_unpack_kernel<<<(unsigned)ceil_div(span, 256), 256, 0, st>>>(gsums, Gk, n, cnt_f64, Btk, gcounts, bcounts, meta, shift, center, dvec);
        SCM_LAUNCH_CHECK(ctx);
    }
    if (do_fit) {
        if (sc) {  // standardize2, pass 2: squared deviations from the global batch means (R/scRUVIII.R:167-170)
            StageTimer t(ctx, ST_STD);
            groupmean_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(bsums, bcounts, Btk, n);
            SCM_LAUNCH_CHECK(ctx);
            if (stream_mode) {
                ssq_finish_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(ssq, bsums, tb, bcnt_local, Btk, n, shift);
                SCM_LAUNCH_CHECK(ctx);
            } else {
                SCM_TRY(seg_colsum(ctx, dY, m, n, ld, dbatch, Btk, bsums, ssq, nullptr, nullptr));
            }
        }
        gram_recentre_kernel<<<dim3((unsigned)ceil_div(n, 256), (unsigned)n), 256, 0, st>>>(Gm, n, uvec, dvec, (double)m);
        SCM_LAUNCH_CHECK(ctx);
        SCM_TRY(allreduce(ctx, buf2, len2, 0));  // all-reduce 2: [centred Gram | ssq]
        if (sc) {
            pooled_sd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(ssq, bcounts, Btk, n, meta, sd, inv_sd);
            SCM_LAUNCH_CHECK(ctx);
        }
        SCM_TRY(gram_residual(ctx, Gm, n, gsums, gcounts, a.G, center, inv_sd));
        int64_t* hmeta = reinterpret_cast<int64_t*>(ctx->h_pinned + 2048);
        SCM_CUDA(cudaMemcpyAsync(hmeta, meta, sizeof(int64_t) * 3, cudaMemcpyDeviceToHost, st));  // lands with the eigensolver's first read-back
        trace.mark("Gram residual enqueued");
        {
            StageTimer t(ctx, ST_EIG);
            const int32_t kconv = a.k < s ? a.k : s;
            rc = eig_topk(ctx, Gm, n, s, kconv, a.mode, a.tol, a.maxit, a.seed, vals, vecs, a.iters_out, a.resid_out, meta);
            if (rc == SCM_ERR_NONFINITE || rc == 0 || rc == SCM_ERR_NOCONV) SCM_TRY(fit_meta_check(hmeta));
            if (rc != 0 && rc != SCM_ERR_NOCONV) return rc;
            dim3 grid((unsigned)ceil_div(n, 256), (unsigned)s);
            fullalpha_kernel<<<grid, 256, 0, st>>>(vals, vecs, s, n, fa, sig);
            SCM_LAUNCH_CHECK(ctx);
        }
        if (a.h_fullalpha) SCM_CUDA(cudaMemcpyAsync(a.h_fullalpha, fa, sizeof(double) * s * n, cudaMemcpyDeviceToHost, st));
        if (a.h_sigma) SCM_CUDA(cudaMemcpyAsync(a.h_sigma, sig, sizeof(double) * s, cudaMemcpyDeviceToHost, st));
        if (sc && a.h_mean) SCM_CUDA(cudaMemcpyAsync(a.h_mean, center, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        if (sc && a.h_sd) SCM_CUDA(cudaMemcpyAsync(a.h_sd, sd, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    } else if (a.center_mode == 1) {
        if (a.h_mean) SCM_CUDA(cudaMemcpyAsync(a.h_mean, center, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    } else if (a.center_mode == 2) {
        if (!a.h_center) return fail(SCM_ERR_INVALID, "scm_ruv3_host: center_mode 2 without h_center");
        SCM_CUDA(cudaMemcpyAsync(center, a.h_center, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    }
    {
        StageTimer t(ctx, ST_COEF);
        const int32_t kk = a.k < s ? a.k : s;
        const double* c_center = sc ? center : (do_fit ? nullptr : (a.center_mode ? center : nullptr));
        int crc = coef_create(ctx, fa, n, kk, dctl, c_center, sc ? inv_sd : nullptr, sc ? sd : nullptr, &coef);
        if (crc != 0) return crc;
    }
    struct CoefGuard { scm_ctx* c; scm_coef* k; ~CoefGuard() { if (k) coef_destroy(c, k); } } coef_guard{ctx, coef};
    trace.mark("eig + coefficients done");
    if (!stream_mode) {
        // ---- in-core: adjust in place chunk by chunk, each finished chunk goes back to the host while the next is computed
        SCM_TRY(adjust_download(ctx, dY, m, n, ld, coef, a.h_out, a.ldo, rows_per_chunk));
        trace.mark("adjust + D2H done");
        return rc;
    }
    // ---- out-of-core, pass 2: every chunk comes up again, is adjusted in its ring slot and goes home; uploads and downloads run
    // concurrently on their own streams (and, for pageable buffers, their own halves of the bounce buffers)
    {
        Uploader up{ctx, pl, a.h_Y, a.ldh, ctx->copy_stream, staged_in, RING, 0, STAGE_NB / 2};
        up.dst.resize(nchunks);
        for (int c = 0; c < nchunks; c++) up.dst[c] = slot(c);
        Downloader down{ctx, pl, a.h_out, a.ldo, ctx->copy_stream_d2h, staged_out, STAGE_NB / 2, STAGE_NB / 2};
        std::vector<cudaEvent_t> ev(nchunks);
        for (auto& e : ev) SCM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        struct EvGuard {
            std::vector<cudaEvent_t>& v;
            ~EvGuard() { for (auto e : v) cudaEventDestroy(e); }
        } guard{ev};
        SCM_TRY(up.init());
        SCM_TRY(down.init());
        SCM_TRY(up.start());
        for (int c = 0; c < nchunks; c++) {
            double* dc = slot(c);
            SCM_TRY(up.wait(c));
            {
                StageTimer t(ctx, ST_ADJUST);
                SCM_TRY(adjust(ctx, dc, pl.rows(c), n, ld, coef, nullptr, 0, dc, ld, nullptr));
            }
            SCM_CUDA(cudaEventRecord(ev[c], st));
            SCM_TRY(down.push(c, dc, ev[c]));
            SCM_TRY(up.release(c, ctx->copy_stream_d2h));  // the slot is free once its chunk has left the device
        }
        SCM_TRY(up.finish());
        SCM_TRY(down.finish());
        SCM_CUDA(cudaStreamSynchronize(st));
    }
    trace.mark("pass 2 done (upload + adjust + download)");
    return rc;
}

}  // namespace scm
