// End-to-end driver for HOST matrices: RUV-III with the PCIe transfers overlapped with the kernels.
//
//   upload  : Y arrives in row chunks on a copy stream; as soon as a chunk has landed the compute stream adds its
//             replicate-group sums (K1) and its shifted Gram (K3, accumulate) -- the fit streams behind the H2D copy
//             (the Gram identity of DESIGN.md section 2 holds for ANY fixed shift, so the shift is the column mean of
//             the first chunk and no second pass over Y is needed),
//   solve   : Gram correction, top-s eigenpairs, coefficients (small, latency bound),
//   download: the fused adjust kernel (K6) runs chunk by chunk IN PLACE and each finished chunk goes back over
//             PCIe while the next one is being adjusted.
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
        { std::lock_guard<std::mutex> g(mu); if (!error) { error = rc; snprintf(msg, sizeof(msg), "%s", why); } }
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

// The download half of the host drivers: the fused adjust kernel (K6) runs chunk by chunk IN PLACE on the device-resident
// matrix and every finished chunk goes back over PCIe (copy stream) while the next one is being adjusted.  d_Y holds newY
// afterwards.  Returns after the last byte has landed in h_out.
inline int adjust_download(scm_ctx* ctx, double* dY, int64_t m, int64_t n, int64_t ld, const scm_coef* coef, double* h_out,
                           int64_t ldo, int64_t rows_per_chunk) {
    if (!dY || !h_out || !coef || ldo < n) return fail(SCM_ERR_INVALID, "adjust_download: bad arguments");
    cudaStream_t st = ctx->stream;
    if (!ctx->copy_stream_d2h) SCM_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream_d2h, cudaStreamNonBlocking));
    cudaStream_t cs = ctx->copy_stream_d2h;
    if (rows_per_chunk <= 0) rows_per_chunk = host_chunk_rows(ld);
    const int nchunks = (int)ceil_div(m, rows_per_chunk);
    std::vector<cudaEvent_t> ev(nchunks + 1);
    for (auto& e : ev) SCM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    struct EvGuard {
        std::vector<cudaEvent_t>& v;
        ~EvGuard() { for (auto e : v) cudaEventDestroy(e); }
    } guard{ev};
    const bool trace = getenv("SCM_HOST_TRACE") != nullptr;
    std::vector<cudaEvent_t> tl_adj, tl_copy;
    cudaEvent_t tl_zero = nullptr;
    if (trace) {
        tl_adj.resize(nchunks); tl_copy.resize(nchunks);
        for (int c = 0; c < nchunks; c++) { cudaEventCreate(&tl_adj[c]); cudaEventCreate(&tl_copy[c]); }
        cudaEventCreate(&tl_zero);
        cudaEventRecord(tl_zero, st);
    }
    // pageable destination: D2H lands in page-locked bounce buffers, a helper thread copies them out with several memcpy threads
    const bool staged = host_is_pageable(h_out);
    ChunkGate issued, drained;
    std::thread drainer;
    HelperJoin joiner{drainer, staged ? &issued : nullptr};
    std::vector<cudaEvent_t> evd;
    struct EvGuard2 {
        std::vector<cudaEvent_t>& v;
        ~EvGuard2() { for (auto e : v) cudaEventDestroy(e); }
    } guard2{evd};
    if (staged) {
        SCM_TRY(stage_reserve(ctx, (size_t)rows_per_chunk * n * 8));
        evd.resize(nchunks);
        for (auto& e : evd) SCM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        const int nt = host_copy_threads();
        const int dev = ctx->device;
        drainer = std::thread([&, nt, dev] {
            cudaSetDevice(dev);
            for (int c = 0; c < nchunks; c++) {
                if (issued.wait_for(c) != 0) return;
                SCM_HELPER_CUDA(cudaEventSynchronize(evd[c]), drained);
                const int64_t r0 = (int64_t)c * rows_per_chunk, rows = (r0 + rows_per_chunk <= m ? rows_per_chunk : m - r0);
                parallel_copy2d(h_out + r0 * ldo, (size_t)ldo * 8, ctx->stage_buf[c % STAGE_NB], (size_t)n * 8, (size_t)n * 8, rows, nt);
                drained.publish(c + 1);
            }
        });
    }
    for (int c = 0; c < nchunks; c++) {
        const int64_t r0 = (int64_t)c * rows_per_chunk, rows = (r0 + rows_per_chunk <= m ? rows_per_chunk : m - r0);
        {
            StageTimer t(ctx, ST_ADJUST);
            SCM_TRY(adjust(ctx, dY + r0 * ld, rows, n, ld, coef, nullptr, 0, dY + r0 * ld, ld, nullptr));
        }
        if (trace) cudaEventRecord(tl_adj[c], st);
        SCM_CUDA(cudaEventRecord(ev[c], st));
        SCM_CUDA(cudaStreamWaitEvent(cs, ev[c], 0));
        if (staged) {
            if (c >= STAGE_NB && drained.wait_for(c - STAGE_NB) != 0) return fail(drained.error, "download helper: %s", drained.msg);
            double* sb = (double*)ctx->stage_buf[c % STAGE_NB];
            if (ld == n) SCM_CUDA(cudaMemcpyAsync(sb, dY + r0 * ld, (size_t)rows * n * 8, cudaMemcpyDeviceToHost, cs));
            else SCM_CUDA(cudaMemcpy2DAsync(sb, n * 8, dY + r0 * ld, ld * 8, n * 8, rows, cudaMemcpyDeviceToHost, cs));
            SCM_CUDA(cudaEventRecord(evd[c], cs));
            issued.publish(c + 1);
        } else if (ldo == n && ld == n)
            SCM_CUDA(cudaMemcpyAsync(h_out + r0 * ldo, dY + r0 * ld, (size_t)rows * n * 8, cudaMemcpyDeviceToHost, cs));
        else
            SCM_CUDA(cudaMemcpy2DAsync(h_out + r0 * ldo, ldo * 8, dY + r0 * ld, ld * 8, n * 8, rows, cudaMemcpyDeviceToHost, cs));
        if (trace) cudaEventRecord(tl_copy[c], cs);
    }
    if (staged) {
        if (drained.wait_for(nchunks - 1) != 0) return fail(drained.error, "download helper: %s", drained.msg);
        joiner.wake = nullptr;  // finished normally
        drainer.join();
    }
    if (trace) {
        cudaStreamSynchronize(cs);
        cudaStreamSynchronize(st);
        for (int c = 0; c < nchunks; c++) {
            float ta = 0.f, tc = 0.f;
            cudaEventElapsedTime(&ta, tl_zero, tl_adj[c]);
            cudaEventElapsedTime(&tc, tl_zero, tl_copy[c]);
            if (c < 4 || c >= nchunks - 2) fprintf(stderr, "[adjust_download]   chunk %2d: adjusted %8.2f ms, D2H landed %8.2f ms\n", c, ta, tc);
            cudaEventDestroy(tl_adj[c]); cudaEventDestroy(tl_copy[c]);
        }
        cudaEventDestroy(tl_zero);
    }
    // the host waits for the downloads: stream-ordered frees of d_Y issued after this function returns are safe
    SCM_CUDA(cudaStreamSynchronize(cs));
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

inline int ruv3_host(scm_ctx* ctx, const HostArgs& a) {
    const HostTrace trace;
    const int64_t m = a.m, n = a.n;
    if (!a.h_Y || !a.h_out || m <= 0 || n <= 0 || a.ldh < n || a.ldo < n || !a.h_ctl) return fail(SCM_ERR_INVALID, "scm_ruv3_host: bad arguments");
    const bool do_fit = a.h_fullalpha_in == nullptr;
    const bool sc = do_fit && a.h_batch != nullptr;
    if (do_fit && (!a.h_group || a.G <= 0 || a.s <= 0)) return fail(SCM_ERR_INVALID, "scm_ruv3_host: fit requested without groups / s");
    cudaStream_t st = ctx->stream;
    if (!ctx->copy_stream) SCM_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    cudaStream_t cs = ctx->copy_stream;
    const int64_t ld = (n + 1) & ~(int64_t)1;
    const int64_t rows_per_chunk = host_chunk_rows(ld);
    const int nchunks = (int)ceil_div(m, rows_per_chunk);
    std::vector<cudaEvent_t> ev(nchunks);
    for (auto& e : ev) SCM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    struct EvGuard {
        std::vector<cudaEvent_t>& v;
        ~EvGuard() { for (auto e : v) cudaEventDestroy(e); }
    } guard{ev};

    Scratch ws(ctx);
    // declared after `ws`, so destroyed before it: on EVERY exit path (errors included) the bulk copies still queued on the copy
    // streams have finished before Scratch hands dY back to the pool
    struct CopyFence {
        scm_ctx* c;
        ~CopyFence() { if (c->copy_stream) cudaStreamSynchronize(c->copy_stream); if (c->copy_stream_d2h) cudaStreamSynchronize(c->copy_stream_d2h); }
    } fence{ctx};
    double *dY, *gsums = nullptr, *tsums = nullptr, *shift = nullptr, *Gm = nullptr, *vals = nullptr, *vecs = nullptr, *fa = nullptr;
    double *bsums = nullptr, *ssq = nullptr, *sd = nullptr, *inv_sd = nullptr, *center = nullptr, *sig = nullptr;
    double *buf1 = nullptr, *buf2 = nullptr, *cnt_f64 = nullptr, *uvec = nullptr, *dvec = nullptr;
    int64_t *gcounts = nullptr, *gcnt_local = nullptr, *tcounts = nullptr, *bcounts = nullptr, *bcnt_local = nullptr, *meta = nullptr;
    int32_t *dgroup = nullptr, *dbatch = nullptr, *flag = nullptr;
    uint8_t* dctl;
    const int32_t s = do_fit ? a.s : a.s_in;
    const int32_t Gk = do_fit ? a.G : 1;  // adjust-only runs use one group (column means)
    const int32_t Btk = sc ? a.Bt : 0;
    // the two all-reduce buffers of fit_core.cuh: [gsums | bsums | counts | flag, rows] and [Gm | ssq]
    const int64_t len1 = (int64_t)(Gk + Btk) * n + Gk + Btk + 2, len2 = (do_fit ? n * n : 0) + (int64_t)Btk * n;
    SCM_TRY(ws.alloc(&dY, m * ld));
    SCM_TRY(ws.alloc(&dctl, n));
    SCM_TRY(ws.alloc(&fa, (int64_t)s * n));
    SCM_TRY(ws.alloc(&meta, 4)); SCM_TRY(ws.alloc(&flag, 1));
    SCM_TRY(ws.alloc(&shift, n)); SCM_TRY(ws.alloc(&center, n)); SCM_TRY(ws.alloc(&uvec, n)); SCM_TRY(ws.alloc(&dvec, n));
    SCM_TRY(ws.alloc(&buf1, len1)); SCM_TRY(ws.alloc(&buf2, len2 > 0 ? len2 : 1));
    gsums = buf1; bsums = buf1 + (int64_t)Gk * n; cnt_f64 = buf1 + (int64_t)(Gk + Btk) * n;
    Gm = buf2; ssq = buf2 + (do_fit ? n * n : 0);
    SCM_CUDA(cudaMemsetAsync(flag, 0, 4, st));
    SCM_CUDA(cudaMemsetAsync(shift, 0, sizeof(double) * n, st));
    SCM_CUDA(cudaMemcpyAsync(dctl, a.h_ctl, n, cudaMemcpyHostToDevice, st));
    SCM_TRY(ws.alloc(&tsums, (int64_t)Gk * n));
    SCM_TRY(ws.alloc(&gcounts, Gk)); SCM_TRY(ws.alloc(&gcnt_local, Gk)); SCM_TRY(ws.alloc(&tcounts, Gk));
    SCM_TRY(ws.alloc(&bcounts, Btk > 0 ? Btk : 1)); SCM_TRY(ws.alloc(&bcnt_local, Btk > 0 ? Btk : 1));
    SCM_TRY(ws.alloc(&dgroup, m));
    SCM_CUDA(cudaMemsetAsync(gsums, 0, sizeof(double) * Gk * n, st));
    SCM_CUDA(cudaMemsetAsync(gcnt_local, 0, sizeof(int64_t) * Gk, st));
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
    }
    // The small uploads / memsets above are copy-engine work of the COMPUTE stream.  While its channel still targets the copy
    // engine, its next operation (even an event record) is arbitrated behind every bulk copy of the copy stream: measured with
    // SCM_HOST_TRACE=1, the whole fit then only started when the last H2D chunk had landed (+144 ms at 1M x 2000).  A trivial
    // kernel moves the channel back to the compute engine, and the host waits for it before the first bulk copy is issued
    // (this also makes the pool allocations visible to the copy stream).
    SCM_TRY(zero_async(ctx, flag, 4));
    SCM_CUDA(cudaStreamSynchronize(st));

    trace.mark("allocated");
    // trace mode: device timeline of every chunk (H2D landed / fit work done), relative to the first enqueue
    std::vector<cudaEvent_t> tl_copy, tl_comp, tl_wait, tl_seg;
    cudaEvent_t tl_zero = nullptr;
    if (trace.on) {
        tl_copy.resize(nchunks); tl_comp.resize(nchunks); tl_wait.resize(nchunks); tl_seg.resize(nchunks);
        for (int c = 0; c < nchunks; c++) { cudaEventCreate(&tl_copy[c]); cudaEventCreate(&tl_comp[c]); cudaEventCreate(&tl_wait[c]); cudaEventCreate(&tl_seg[c]); }
        cudaEventCreate(&tl_zero);
        cudaEventRecord(tl_zero, st);
    }
    // ---- upload, with the fit accumulating behind it
    const bool need_sums = do_fit || a.center_mode == 1;
    const bool copy_1d = a.ldh == n && ld == n;
    auto issue_copy = [&](int c) -> int {
        const int64_t r0 = (int64_t)c * rows_per_chunk, rows = (r0 + rows_per_chunk <= m ? rows_per_chunk : m - r0);
        if (copy_1d) SCM_CUDA(cudaMemcpyAsync(dY + r0 * ld, a.h_Y + r0 * a.ldh, (size_t)rows * n * 8, cudaMemcpyHostToDevice, cs));
        else SCM_CUDA(cudaMemcpy2DAsync(dY + r0 * ld, ld * 8, a.h_Y + r0 * a.ldh, a.ldh * 8, n * 8, rows, cudaMemcpyHostToDevice, cs));
        if (trace.on) cudaEventRecord(tl_copy[c], cs);
        SCM_CUDA(cudaEventRecord(ev[c], cs));
        return 0;
    };
    // pageable source: a helper thread stages chunk by chunk through the page-locked bounce buffers (see above)
    const bool staged_in = host_is_pageable(a.h_Y);
    ChunkGate fed;
    std::thread feeder;
    HelperJoin feeder_join{feeder, nullptr};
    if (staged_in) {
        SCM_TRY(stage_reserve(ctx, (size_t)rows_per_chunk * n * 8));
        const int nt = host_copy_threads();
        const int dev = ctx->device;
        feeder = std::thread([&, nt, dev] {
            cudaSetDevice(dev);
            for (int c = 0; c < nchunks; c++) {
                if (fed.error) return;
                if (c >= STAGE_NB) SCM_HELPER_CUDA(cudaEventSynchronize(ev[c - STAGE_NB]), fed);  // the slot's previous H2D has finished
                const int64_t r0 = (int64_t)c * rows_per_chunk, rows = (r0 + rows_per_chunk <= m ? rows_per_chunk : m - r0);
                double* sb = (double*)ctx->stage_buf[c % STAGE_NB];
                parallel_copy2d(sb, (size_t)n * 8, a.h_Y + r0 * a.ldh, (size_t)a.ldh * 8, (size_t)n * 8, rows, nt);
                if (ld == n) SCM_HELPER_CUDA(cudaMemcpyAsync(dY + r0 * ld, sb, (size_t)rows * n * 8, cudaMemcpyHostToDevice, cs), fed);
                else SCM_HELPER_CUDA(cudaMemcpy2DAsync(dY + r0 * ld, ld * 8, sb, n * 8, n * 8, rows, cudaMemcpyHostToDevice, cs), fed);
                SCM_HELPER_CUDA(cudaEventRecord(ev[c], cs), fed);
                fed.publish(c + 1);
            }
        });
        feeder_join.wake = &fed;
    } else {
        SCM_TRY(issue_copy(0));
        if (nchunks > 1) SCM_TRY(issue_copy(1));
    }
    for (int c = 0; c < nchunks; c++) {
        const int64_t r0 = (int64_t)c * rows_per_chunk, rows = (r0 + rows_per_chunk <= m ? rows_per_chunk : m - r0);
        if (trace.on && c > 0) cudaEventRecord(tl_comp[c - 1], st);
        // host-mediated hand-over with two bulk copies in flight: chunk c is in HBM when the event has fired, its fit work
        // is enqueued right away and copy c + 2 joins the queue (the copy engine never idles, the queue never grows)
        if (staged_in && fed.wait_for(c) != 0) return fail(fed.error, "upload helper: %s", fed.msg);
        SCM_CUDA(cudaEventSynchronize(ev[c]));
        if (!staged_in && c + 2 < nchunks) SCM_TRY(issue_copy(c + 2));
        if (trace.on) cudaEventRecord(tl_wait[c], st);
        if (!need_sums) continue;
        {
            StageTimer t(ctx, ST_SEG);
            SCM_TRY(seg_colsum(ctx, dY + r0 * ld, rows, n, ld, dgroup + r0, Gk, nullptr, tsums, tcounts, flag));
            if (trace.on) cudaEventRecord(tl_seg[c], st);
            add_f64_kernel<<<(unsigned)ceil_div((int64_t)Gk * n, 256), 256, 0, st>>>(gsums, tsums, (int64_t)Gk * n);
            SCM_LAUNCH_CHECK(ctx);
            add_i64_kernel<<<(unsigned)ceil_div(Gk, 256), 256, 0, st>>>(gcnt_local, tcounts, Gk);
            SCM_LAUNCH_CHECK(ctx);
        }
        if (!do_fit) continue;
        if (c == 0) {
            // Gram shift = column mean of this rank's FIRST chunk: known before the second chunk lands, no agreement between
            // ranks needed (the local Gram is moved to the global mean by a rank-2 update before its all-reduce, fit_core.cuh)
            first_chunk_mean_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(tsums, Gk, n, (double)rows, shift);
            SCM_LAUNCH_CHECK(ctx);
        }
        StageTimer t(ctx, ST_GRAM);
        SCM_TRY(gram(ctx, dY + r0 * ld, rows, n, ld, shift, Gm, c == nchunks - 1, c > 0));
    }
    if (staged_in) {  // every chunk has been published: the helper is done
        feeder_join.wake = nullptr;
        feeder.join();
    }
    trace.mark("upload + fit chunks enqueued");
    if (trace.on) {
        cudaEventRecord(tl_comp[nchunks - 1], st);
        cudaStreamSynchronize(cs); trace.mark("H2D done"); cudaStreamSynchronize(st); trace.mark("chunk Gram / sums done");
        for (int c = 0; c < nchunks; c++) {
            float t_copy = 0.f, t_comp = 0.f, t_wait = 0.f, t_seg = 0.f;
            cudaEventElapsedTime(&t_copy, tl_zero, tl_copy[c]);
            cudaEventElapsedTime(&t_comp, tl_zero, tl_comp[c]);
            cudaEventElapsedTime(&t_wait, tl_zero, tl_wait[c]);
            if (need_sums) cudaEventElapsedTime(&t_seg, tl_zero, tl_seg[c]);
            if (c < 6 || c >= nchunks - 3)
                fprintf(stderr, "[scm_ruv3_host]   chunk %2d: H2D landed %8.2f ms, wait released %8.2f, sums done %8.2f, fit work done %8.2f ms\n", c,
                        t_copy, t_wait, t_seg, t_comp);
            cudaEventDestroy(tl_copy[c]); cudaEventDestroy(tl_comp[c]); cudaEventDestroy(tl_wait[c]); cudaEventDestroy(tl_seg[c]);
        }
        cudaEventDestroy(tl_zero);
    }
    int rc = 0;
    scm_coef* coef = nullptr;
    if (need_sums) {
        // all-reduce 1 (one float64 buffer): group sums, batch sums, counts, NA/Inf flag, rows -> global column means
        if (sc) {  // standardize2, pass 1 over the resident cells (R/scRUVIII.R:163-166)
            StageTimer t(ctx, ST_STD);
            SCM_TRY(seg_colsum(ctx, dY, m, n, ld, dbatch, Btk, nullptr, bsums, bcnt_local, nullptr));
        }
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
            SCM_TRY(seg_colsum(ctx, dY, m, n, ld, dbatch, Btk, bsums, ssq, nullptr, nullptr));
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
    trace.mark("eig + coefficients done");
    // ---- adjust in place chunk by chunk, each finished chunk goes back to the host while the next is computed
    {
        const int arc = adjust_download(ctx, dY, m, n, ld, coef, a.h_out, a.ldo, rows_per_chunk);
        coef_destroy(ctx, coef);
        if (arc != 0) return arc;
    }
    trace.mark("adjust + D2H done");
    return rc;
}

}  // namespace scm
