// Shared infrastructure of libscmerge_b200: error handling, context, scratch memory, PTX helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/scmerge_b200.h"

namespace scm {

static thread_local char g_err[512] = "";

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define SCM_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return scm::fail(SCM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                             __FILE__, __LINE__);                                                   \
    } while (0)

#define SCM_TRY(expr)           \
    do {                        \
        int _rc = (expr);       \
        if (_rc != 0) return _rc; \
    } while (0)

#define SCM_LAUNCH_CHECK(ctx)   \
    do {                        \
        (ctx)->launches++;      \
        SCM_CUDA(cudaGetLastError()); \
    } while (0)

enum Stage { ST_SEG = 0, ST_GRAM = 1, ST_COMM = 2, ST_EIG = 3, ST_COEF = 4, ST_ADJUST = 5, ST_STD = 6, ST_SYRK = 7, ST_INGEST = 8, ST_PRESHIFT = 9, ST_COUNT = 10 };

}  // namespace scm

struct scm_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // host -> device bulk transfers of the host-matrix drivers (created on first use)
    cudaStream_t copy_stream_d2h = nullptr;  // device -> host bulk transfers (a separate stream per direction, see host_pipeline.cuh)
    bool own_stream = false;
    int64_t launches = 0;
    scm_allreduce_fn allreduce = nullptr;
    void* allreduce_user = nullptr;
    int world = 1;
    cudaEvent_t ev0[scm::ST_COUNT] = {}, ev1[scm::ST_COUNT] = {};
    bool ev_used[scm::ST_COUNT];
    bool use_tma = true;         // Gram: TMA + mbarrier kernel (false: cp.async kernel; env SCM_GRAM_TMA=0)
    bool gram_preshift = true;   // fits run the Gram on a pre-shifted copy of the cells when there is room (env SCM_GRAM_PRESHIFT=0: off)
    int gram_lean = 3;           // Gram TMA kernel: bit 0 thin last tile row, bit 1 triangular diagonal units (env SCM_GRAM_LEAN)
    bool gram_class_order = true;  // Gram TMA kernel: class-ordered unit list (env SCM_GRAM_ORDER=0: split-major list)
    bool k1_fused_copy = true;   // the pre-shifted copy is written by K1's summing pass (env SCM_K1_FUSED_COPY=0: separate copy kernel)
    double* h_pinned = nullptr;  // small pinned staging area (scalars read back by the eigensolver)
    // page-locked bounce buffers of the host drivers for callers whose matrices are ordinary pageable memory (R's REAL(),
    // plain numpy): allocated on first use, one PCIe chunk each, kept until the context dies (host_pipeline.cuh)
    void* stage_buf[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t stage_bytes = 0;
    int conv_ranks[16];          // additional subspace dimensions the EXACT eigensolver must converge (scm_ctx_set_conv_ranks)
    int n_conv_ranks = 0;
    // built-in communicator (comm.cuh): NCCL resolved with dlopen at run time, one rank per context.  Takes precedence over the
    // host-supplied hook.  `comm_owned` = created by scm_ctx_comm_init_nccl (destroyed with the context); a scm_multi owns its own.
    void* nccl_comm = nullptr;
    bool comm_owned = false;
    int rank = 0;
    // facts about the last fit on this context (scm_ctx_fit_info)
    int32_t fit_iters = 0, fit_conv_rank = 0, fit_products = 0;
    double fit_resid = 0.0;
};

namespace scm {

// Stream-ordered scratch allocation that frees itself on scope exit.
struct Scratch {
    scm_ctx* ctx;
    std::vector<void*> ptrs;
    explicit Scratch(scm_ctx* c) : ctx(c) {}
    template <typename T>
    int alloc(T** out, int64_t count) {
        void* p = nullptr;
        int64_t bytes = count * (int64_t)sizeof(T);
        if (bytes <= 0) bytes = 16;
        cudaError_t e = cudaMallocAsync(&p, (size_t)bytes, ctx->stream);
        if (e != cudaSuccess) return fail(SCM_ERR_CUDA, "scratch alloc of %lld bytes failed: %s", (long long)bytes, cudaGetErrorString(e));
        ptrs.push_back(p);
        *out = (T*)p;
        return 0;
    }
    ~Scratch() {
        for (void* p : ptrs) cudaFreeAsync(p, ctx->stream);
    }
};

struct StageTimer {
    scm_ctx* ctx;
    int st;
    StageTimer(scm_ctx* c, int s) : ctx(c), st(s) {
        cudaEventRecord(ctx->ev0[st], ctx->stream);
    }
    ~StageTimer() {
        cudaEventRecord(ctx->ev1[st], ctx->stream);
        ctx->ev_used[st] = true;
    }
};

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Small device-side zero fill / copy as KERNELS.  cudaMemsetAsync / device-to-device cudaMemcpyAsync may be executed by a
// copy engine, where they queue behind every bulk PCIe transfer already submitted by another stream: in the chunked host
// drivers one such call stalled the whole fit until the last H2D chunk had landed (measured with SCM_HOST_TRACE=1:
// 144 ms at 1M x 2000).  Anything issued on the compute stream while bulk copies are in flight uses these instead.
__global__ void zero_u32_kernel(uint32_t* __restrict__ p, int64_t nwords) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0u;
}
__global__ void copy_u32_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, int64_t nwords) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
inline int zero_async(scm_ctx* ctx, void* p, int64_t bytes) {  // bytes must be a multiple of 4
    if (bytes <= 0) return 0;
    const int64_t nw = bytes / 4;
    const int64_t blocks = ceil_div(nw, 256);
    zero_u32_kernel<<<(unsigned)(blocks < 1184 ? blocks : 1184), 256, 0, ctx->stream>>>((uint32_t*)p, nw);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}
inline int copy_async(scm_ctx* ctx, void* dst, const void* src, int64_t bytes) {  // bytes must be a multiple of 4
    if (bytes <= 0) return 0;
    const int64_t nw = bytes / 4;
    const int64_t blocks = ceil_div(nw, 256);
    copy_u32_kernel<<<(unsigned)(blocks < 1184 ? blocks : 1184), 256, 0, ctx->stream>>>((uint32_t*)dst, (const uint32_t*)src, nw);
    SCM_LAUNCH_CHECK(ctx);
    return 0;
}

// ---- device helpers ---------------------------------------------------------------------------------

// D(8x8) += A(8x4, row) * B(4x8, col), FP64 tensor core (SASS: DMMA.8x8x4).
// lane = 4*gid + tig: a = A[gid][tig], b = B[tig][gid], c0/c1 = C[gid][2*tig], C[gid][2*tig+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
// Ordered variant for the mbarrier-pipelined kernels: being `volatile` it cannot be moved past the (volatile)
// mbarrier.arrive that hands a shared-memory slot back to the TMA producer, and because a DMMA only issues once its
// source registers have arrived, "all DMMAs of the slot issued" implies "all shared-memory loads of the slot done".
__device__ __forceinline__ void dmma884_ordered(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// streaming 16-byte global load / store (read-once / write-once data: keep it out of L1)
__device__ __forceinline__ double2 ldg_stream2(const double* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream2(double* p, double2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};\n" ::"l"(p), "d"(v.x), "d"(v.y));
}

// 32-byte global load / store (sm_100: LDG.E.256 / STG.E.256): one L1TEX request per lane where two 16-byte ones were needed;
// the address must be 32-byte aligned
__device__ __forceinline__ void stg256(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void stg256_stream(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void ldg256_stream(const double* p, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];\n" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

__device__ __forceinline__ bool nonfinite(double v) {
    return (__double2hiint(v) & 0x7ff00000) == 0x7ff00000;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace scm
