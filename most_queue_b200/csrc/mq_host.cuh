// mq_host.cuh -- host-side plumbing shared by the C-ABI translation units: the context, error
// reporting, scratch buffers, mq_dist validation and its fp32 device form.
#pragma once

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include "mq_device.cuh"

namespace mq {

extern thread_local char g_err[512];

inline int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(expr)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return ::mq::fail(MQ_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                              __LINE__);                                                      \
    } while (0)

struct Buf {
    void *p = nullptr;
    size_t cap = 0;
};

}  // namespace mq

struct mq_ctx {
    int device = 0;
    int sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // start, sim done, reduce done, end
    mq::Buf rec, agg, ring, in_a, in_s, desc, lookback;
    double *h_agg = nullptr;  // pinned staging
    size_t h_agg_cap = 0;
    double sim_ms = 0, reduce_ms = 0, total_ms = 0;
    int launches = 0;
    // mq_scan_gg1_summary(MQ_SCAN_KEEP) leaves the range's intermediate results (tile prefixes, segment prefixes, the
    // fp32 draw stash) in the scratch buffers; the next call on the context, if it is the matching mq_scan_gg1_apply,
    // walks from them instead of forming them again.  Any other entry point clears the flag (use_device).
    bool scan_valid = false;
    alignas(16) unsigned char scan_cache[1024];
    unsigned char scan_key[256];
};

namespace mq {

// every entry point that is about to use the context's scratch buffers starts with this
inline cudaError_t use_device(mq_ctx *ctx)
{
    ctx->scan_valid = false;
    return cudaSetDevice(ctx->device);
}

inline int ensure(Buf &b, size_t bytes)
{
    if (bytes <= b.cap) return MQ_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    CU(cudaMalloc(&b.p, bytes));
    b.cap = bytes;
    return MQ_OK;
}

inline int ensure_host_agg(mq_ctx *ctx, size_t bytes)
{
    if (int rc = ensure(ctx->agg, bytes)) return rc;
    if (ctx->h_agg_cap < bytes) {
        if (ctx->h_agg) cudaFreeHost(ctx->h_agg);
        ctx->h_agg = nullptr;
        ctx->h_agg_cap = 0;
        CU(cudaMallocHost(&ctx->h_agg, bytes));
        ctx->h_agg_cap = bytes;
    }
    return MQ_OK;
}

inline int finish_timing(mq_ctx *ctx)
{
    float a = 0, b = 0, c = 0;
    CU(cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]));
    CU(cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]));
    CU(cudaEventElapsedTime(&c, ctx->ev[0], ctx->ev[3]));
    ctx->sim_ms = a;
    ctx->reduce_ms = b;
    ctx->total_ms = c;
    return MQ_OK;
}

struct RatioDesc {
    int num, den;
};
cudaError_t launch_reduce_records(const double *rec, long long R, int p_bins, double *agg, cudaStream_t st);
cudaError_t launch_reduce_ratio(const double *rec, long long R, const RatioDesc *desc, int nq, double *out,
                                cudaStream_t st);

int check_dist(const mq_dist &d, const char *what);
DistF make_distf(const mq_dist &d);
DistD make_distd(const mq_dist &d);
uint32_t seed_to_key(uint64_t s);

}  // namespace mq
