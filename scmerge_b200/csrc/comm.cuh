// Cross-rank sums of the fit: group sums + counts, batch sums and the n x n Gram partial.
//
// Three transports, one entry point (`allreduce`):
//   1. built-in NCCL, one rank per context (scm_ctx_comm_init_nccl: one process per GPU, the host only distributes the
//      128-byte unique id) or one process driving several GPUs (scm_multi_create: ncclCommInitAll, one host thread per device),
//   2. the host-supplied hook (scm_ctx_set_allreduce; `torch.distributed` in the Python host, gloo in the CPU tests),
//   3. nothing (a single rank).
// NCCL is resolved with dlopen("libnccl.so.2") at first use, so the library has no link-time dependency on it: inside a torch
// process the loader hands back torch's bundled NCCL (same SONAME), inside R the system one.  Collectives are enqueued on the
// context's stream; nothing here synchronises the host.
#pragma once
#include <dlfcn.h>

#include <mutex>
#include <string>
#include <thread>

#include "common.cuh"

namespace scm {

// the part of nccl.h this file needs (stable ABI since NCCL 2.0)
typedef struct ncclComm* nccl_comm_t;
struct nccl_unique_id { char internal[128]; };
enum { NCCL_SUM = 0, NCCL_INT64 = 4, NCCL_FLOAT64 = 8 };

struct NcclApi {
    int (*GetUniqueId)(nccl_unique_id*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
    int (*CommInitAll)(nccl_comm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string why;
};

inline const NcclApi& nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        void* h = nullptr;
        const char* names[] = {getenv("SCM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) { a.why = std::string("dlopen(libnccl.so.2) failed: ") + (dlerror() ? dlerror() : "?"); return a; }
        auto sym = [&](const char* s) { return dlsym(h, s); };
        a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
        a.CommInitAll = (decltype(a.CommInitAll))sym("ncclCommInitAll");
        a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
        a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
        a.GetVersion = (decltype(a.GetVersion))sym("ncclGetVersion");
        a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommInitAll && a.CommDestroy && a.AllReduce && a.GetErrorString;
        if (!a.ok) a.why = "libnccl.so.2 lacks an expected symbol";
        return a;
    }();
    return api;
}

#define SCM_NCCL(expr)                                                                                              \
    do {                                                                                                            \
        int _r = (expr);                                                                                            \
        if (_r != 0) return scm::fail(SCM_ERR_COMM, "%s failed: %s", #expr, scm::nccl_api().GetErrorString(_r));    \
    } while (0)

// In-place sum over the ranks that share the cells.  dtype: 0 = float64, 1 = int64.
inline int allreduce(scm_ctx* ctx, void* d_buf, int64_t count, int dtype) {
    if (ctx->world <= 1 || count <= 0) return 0;
    StageTimer t(ctx, ST_COMM);
    if (ctx->nccl_comm) {
        const NcclApi& api = nccl_api();
        SCM_NCCL(api.AllReduce(d_buf, d_buf, (size_t)count, dtype == 0 ? NCCL_FLOAT64 : NCCL_INT64, NCCL_SUM, (nccl_comm_t)ctx->nccl_comm,
                               ctx->stream));
        return 0;
    }
    if (!ctx->allreduce) return fail(SCM_ERR_COMM, "world size %d but neither a communicator nor an all-reduce hook is set", ctx->world);
    const int rc = ctx->allreduce(d_buf, count, dtype, (void*)ctx->stream, ctx->allreduce_user);
    if (rc != 0) return fail(SCM_ERR_COMM, "all-reduce hook returned %d", rc);
    return 0;
}

inline int comm_release(scm_ctx* ctx) {
    if (ctx->nccl_comm && ctx->comm_owned) nccl_api().CommDestroy((nccl_comm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->comm_owned = false;
    return 0;
}

}  // namespace scm

// One process, several GPUs: a context, a stream and an NCCL rank per device.  Calls fan out to one host thread per device
// (each thread only ever touches its own context), which is what lets the unchanged per-rank drivers run side by side.
struct scm_multi {
    std::vector<scm_ctx*> ctx;
    std::vector<scm::nccl_comm_t> comms;
};

namespace scm {

// Runs fn(rank) on one thread per device and returns the first non-zero status (with that thread's message).
template <typename F>
inline int multi_run(scm_multi* mg, F fn) {
    const int nd = (int)mg->ctx.size();
    std::vector<int> rc(nd, 0);
    std::vector<std::string> msg(nd);
    std::vector<std::thread> th;
    th.reserve(nd);
    for (int r = 0; r < nd; r++)
        th.emplace_back([&, r] {
            g_err[0] = 0;
            rc[r] = fn(r);
            if (rc[r] != 0) msg[r] = g_err;
        });
    for (auto& t : th) t.join();
    int worst = 0;
    for (int r = 0; r < nd; r++) {
        if (rc[r] == 0) continue;
        // a real failure outranks "not converged", which every rank reports identically
        if (worst == 0 || (worst == SCM_ERR_NOCONV && rc[r] != SCM_ERR_NOCONV)) {
            worst = rc[r];
            snprintf(g_err, sizeof(g_err), "[device %d] %s", mg->ctx[r]->device, msg[r].c_str());
        }
    }
    return worst;
}

}  // namespace scm
