// mq_api.cu -- the C ABI of libmqb200.so (include/mqb200.h): context, argument checking,
// scratch management, kernel dispatch, timing.  No CPU fallback exists: every entry point
// that computes launches CUDA kernels and fails with MQ_E_NODEVICE / MQ_E_CUDA otherwise.
#include "mq_host.cuh"
#include "k1_fifo_replay.cuh"
#include "k2_fifo_gen.cuh"

#include <string>
#include <thread>
#include <vector>

namespace mq {

#define MQ_DECL_K2(n) \
    cudaError_t k2_dispatch_ct##n(const K2Params &, int, bool, bool, int, size_t, cudaStream_t, int *);
#define MQ_DECL_K1(n) cudaError_t k1_dispatch_ct##n(const K1Params &, bool, cudaStream_t);
MQ_DECL_K2(1) MQ_DECL_K2(2) MQ_DECL_K2(3) MQ_DECL_K2(4) MQ_DECL_K2(5) MQ_DECL_K2(6) MQ_DECL_K2(7)
MQ_DECL_K2(8) MQ_DECL_K2(16) MQ_DECL_K2(32)
MQ_DECL_K1(1) MQ_DECL_K1(2) MQ_DECL_K1(3) MQ_DECL_K1(4) MQ_DECL_K1(5) MQ_DECL_K1(6) MQ_DECL_K1(7)
MQ_DECL_K1(8) MQ_DECL_K1(16) MQ_DECL_K1(32)

thread_local char g_err[512] = "";

static int ct_for(int c)
{
    if (c >= 1 && c <= 8) return c;
    if (c <= 16) return 16;
    if (c <= 32) return 32;
    return 0;
}

int check_dist(const mq_dist &d, const char *what)
{
    const double *p = d.p;
    auto pos = [](double x) { return std::isfinite(x) && x > 0.0; };
    switch (d.kind) {
    case MQ_M:
        if (!pos(p[0])) return fail(MQ_E_INVALID, "%s: M needs mu > 0", what);
        break;
    case MQ_H2:
    case MQ_C2:
        if (!(p[0] >= 0.0 && p[0] <= 1.0) || !pos(p[1]) || !pos(p[2]))
            return fail(MQ_E_INVALID, "%s: H/C need 0 <= p1 <= 1 and real mu1, mu2 > 0", what);
        break;
    case MQ_E:
        if (!(p[0] >= 1.0 && p[0] <= 64.0) || std::floor(p[0]) != p[0] || !pos(p[1]))
            return fail(MQ_E_INVALID, "%s: E needs integer 1 <= r <= 64 and mu > 0", what);
        break;
    case MQ_GAMMA:
        if (!pos(p[0]) || !pos(p[1])) return fail(MQ_E_INVALID, "%s: Gamma needs mu, alpha > 0", what);
        break;
    case MQ_PA:
        if (!pos(p[0]) || !pos(p[1])) return fail(MQ_E_INVALID, "%s: Pa needs alpha, K > 0", what);
        break;
    case MQ_D:
        if (!(std::isfinite(p[0]) && p[0] >= 0.0)) return fail(MQ_E_INVALID, "%s: D needs b >= 0", what);
        break;
    case MQ_UNIFORM:
        if (!(std::isfinite(p[0]) && p[1] >= 0.0 && p[0] - p[1] >= 0.0))
            return fail(MQ_E_INVALID, "%s: Uniform needs mean - half_interval >= 0", what);
        break;
    case MQ_NORM:
        if (!(std::isfinite(p[0]) && p[1] >= 0.0)) return fail(MQ_E_INVALID, "%s: Norm needs std_dev >= 0", what);
        break;
    case MQ_WEIBULL:
        if (!pos(p[0]) || !pos(p[1])) return fail(MQ_E_INVALID, "%s: Weibull needs k, W > 0", what);
        break;
    default:
        return fail(MQ_E_INVALID, "%s: unknown distribution kind %d", what, d.kind);
    }
    return MQ_OK;
}

static uint32_t branch_threshold(double p1)
{
    const double t = std::floor(p1 * 4294967296.0);
    return t >= 4294967295.0 ? 0xFFFFFFFFu : (t <= 0.0 ? 0u : (uint32_t)t);
}

DistF make_distf(const mq_dist &d)
{
    const double LN2 = 0.6931471805599453;
    DistF f;
    std::memset(&f, 0, sizeof(f));
    f.kind = d.kind;
    const double *p = d.p;
    switch (d.kind) {
    case MQ_M:
        f.a = (float)(-LN2 / p[0]);
        break;
    case MQ_H2:
        f.T = branch_threshold(p[0]);
        f.a = f.T ? (float)(1.0 / (double)f.T) : 0.0f;
        f.b = (float)(1.0 / (4294967296.0 - (double)f.T));
        f.c = (float)(-LN2 / p[1]);
        f.d = (float)(-LN2 / p[2]);
        break;
    case MQ_E:
        f.r = (int)p[0];
        f.a = (float)(-LN2 / p[1]);
        break;
    case MQ_GAMMA: {
        const double alpha = p[1], a = alpha >= 1.0 ? alpha : alpha + 1.0;
        const double dd = a - 1.0 / 3.0;
        f.a = (float)dd;
        f.b = (float)(1.0 / std::sqrt(9.0 * dd));
        f.c = (float)(1.0 / p[0]);
        f.d = (float)(1.0 / alpha);
        f.r = alpha < 1.0 ? 1 : 0;
        break;
    }
    case MQ_PA:
        f.a = (float)(-1.0 / p[0]);
        f.b = (float)p[1];
        break;
    case MQ_D:
        f.a = (float)p[0];
        break;
    case MQ_C2:
        f.T = branch_threshold(p[0]);
        f.a = (float)(-LN2 / p[1]);
        f.b = (float)(-LN2 / p[2]);
        break;
    case MQ_UNIFORM:
        f.a = (float)(p[0] - p[1]);
        f.b = (float)(2.0 * p[1]);
        break;
    case MQ_NORM:
        f.a = (float)p[0];
        f.b = (float)p[1];
        break;
    case MQ_WEIBULL:
        f.a = (float)p[1];
        f.b = (float)(-1.0 / p[0]);
        break;
    }
    return f;
}

DistD make_distd(const mq_dist &d)
{
    DistD f;
    std::memset(&f, 0, sizeof(f));
    f.kind = d.kind;
    const double *p = d.p;
    switch (d.kind) {
    case MQ_M:
        f.a = 1.0 / p[0];
        break;
    case MQ_H2:
    case MQ_C2:
        f.T = branch_threshold(p[0]);
        f.a = 1.0 / p[1];
        f.b = 1.0 / p[2];
        break;
    case MQ_E:
        f.r = (int)p[0];
        f.a = 1.0 / p[1];
        break;
    case MQ_GAMMA: {
        const double alpha = p[1], a = alpha >= 1.0 ? alpha : alpha + 1.0;
        f.a = a - 1.0 / 3.0;
        f.b = 1.0 / std::sqrt(9.0 * f.a);
        f.c = 1.0 / p[0];
        f.d = 1.0 / alpha;
        f.r = alpha < 1.0 ? 1 : 0;
        break;
    }
    case MQ_PA:
        f.a = -1.0 / p[0];
        f.b = p[1];
        break;
    case MQ_D:
        f.a = p[0];
        break;
    case MQ_UNIFORM:
        f.a = p[0] - p[1];
        f.b = 2.0 * p[1];
        break;
    case MQ_NORM:
        f.a = p[0];
        f.b = p[1];
        break;
    case MQ_WEIBULL:
        f.a = p[1];
        f.b = -1.0 / p[0];
        break;
    }
    return f;
}

// 64-bit user seed -> 32-bit Philox key (splitmix64 finaliser)
uint32_t seed_to_key(uint64_t s)
{
    s += 0x9E3779B97F4A7C15ull;
    s = (s ^ (s >> 30)) * 0xBF58476D1CE4E5B9ull;
    s = (s ^ (s >> 27)) * 0x94D049BB133111EBull;
    s ^= s >> 31;
    return (uint32_t)(s ^ (s >> 32));
}

static cudaError_t k2_dispatch(int ct, const K2Params &P, int pc, bool fin, bool busy, int grid, size_t smem,
                               cudaStream_t st, int *occ)
{
    switch (ct) {
#define MQ_CASE(n) \
    case n:        \
        return k2_dispatch_ct##n(P, pc, fin, busy, grid, smem, st, occ);
        MQ_CASE(1) MQ_CASE(2) MQ_CASE(3) MQ_CASE(4) MQ_CASE(5) MQ_CASE(6) MQ_CASE(7) MQ_CASE(8)
        MQ_CASE(16) MQ_CASE(32)
#undef MQ_CASE
    default:
        return cudaErrorInvalidValue;
    }
}

static cudaError_t k1_dispatch(int ct, const K1Params &P, bool fin, cudaStream_t st)
{
    switch (ct) {
#define MQ_CASE(n) \
    case n:        \
        return k1_dispatch_ct##n(P, fin, st);
        MQ_CASE(1) MQ_CASE(2) MQ_CASE(3) MQ_CASE(4) MQ_CASE(5) MQ_CASE(6) MQ_CASE(7) MQ_CASE(8)
        MQ_CASE(16) MQ_CASE(32)
#undef MQ_CASE
    default:
        return cudaErrorInvalidValue;
    }
}

__global__ void debug_philox_kernel(uint32_t c0, uint32_t c1, uint32_t key, uint32_t *out)
{
    uint2 w = philox2x32_10(c0, c1, key);
    out[0] = w.x;
    out[1] = w.y;
}

__global__ void debug_variates_kernel(DistF d, uint32_t key0, uint32_t rep, int which, long long n, float *out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint2 w = philox2x32_10((uint32_t)i, rep, key0);
    out[i] = sample<-1>(d, which ? w.y : w.x, (uint32_t)i, rep, key0, which);
}

}  // namespace mq

using namespace mq;

extern "C" {

int mq_version(void) { return MQ_ABI_VERSION; }

const char *mq_last_error(void) { return g_err; }

int mq_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int mq_create(mq_ctx **out, int device_id)
{
    if (!out) return fail(MQ_E_INVALID, "mq_create: out is NULL");
    *out = nullptr;
    int n = mq_device_count();
    if (n <= 0) return fail(MQ_E_NODEVICE, "mq_create: no CUDA device visible (this library has no CPU path)");
    if (device_id < 0 || device_id >= n) return fail(MQ_E_INVALID, "mq_create: device %d of %d", device_id, n);
    CU(cudaSetDevice(device_id));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device_id));
    if (prop.major != 10)
        return fail(MQ_E_NODEVICE, "mq_create: device %d is sm_%d%d; this library is built for sm_100a only",
                    device_id, prop.major, prop.minor);
    mq_ctx *ctx = new (std::nothrow) mq_ctx();
    if (!ctx) return fail(MQ_E_INVALID, "mq_create: out of host memory");
    ctx->device = device_id;
    ctx->sms = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    for (auto &e : ctx->ev) CU(cudaEventCreate(&e));
    *out = ctx;
    return MQ_OK;
}

int mq_destroy(mq_ctx *ctx)
{
    if (!ctx) return MQ_OK;
    cudaSetDevice(ctx->device);
    for (Buf *b : {&ctx->rec, &ctx->agg, &ctx->ring, &ctx->in_a, &ctx->in_s, &ctx->desc, &ctx->lookback})
        if (b->p) cudaFree(b->p);
    if (ctx->h_agg) cudaFreeHost(ctx->h_agg);
    for (auto &e : ctx->ev)
        if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return MQ_OK;
}

static int stage_agg(mq_ctx *ctx, int p_bins)
{
    return ensure_host_agg(ctx, sizeof(double) * MQ_AGG_LEN(p_bins));
}

int mq_sim_fifo(mq_ctx *ctx, const mq_fifo_cfg *cfg, double *agg, double *rep_out)
{
    if (!ctx || !cfg || !agg) return fail(MQ_E_INVALID, "mq_sim_fifo: NULL argument");
    if (cfg->c < 1) return fail(MQ_E_INVALID, "mq_sim_fifo: num_of_channels must be >= 1");
    const int ct = ct_for(cfg->c);
    if (!ct) return fail(MQ_E_UNSUPPORTED, "mq_sim_fifo: num_of_channels %d > %d", cfg->c, MQ_MAX_CHANNELS);
    if (cfg->p_bins < 1 || cfg->p_bins > MQ_MAX_PBINS)
        return fail(MQ_E_INVALID, "mq_sim_fifo: p_bins must be in 1..%d", MQ_MAX_PBINS);
    if (cfg->buffer < -1) return fail(MQ_E_INVALID, "mq_sim_fifo: buffer must be >= 0 or -1 (infinite)");
    if (cfg->total_served < 0 || cfg->total_served >= (1ll << 31))
        return fail(MQ_E_INVALID, "mq_sim_fifo: total_served must be in 0..2^31-1");
    if (cfg->replications < 1 || cfg->rep0 < 0 || cfg->rep0 + cfg->replications > (1ll << 32))
        return fail(MQ_E_INVALID, "mq_sim_fifo: replication ids must fit 32 bits");
    if (cfg->warmup < 0 || (cfg->warmup > 0 && cfg->warmup >= cfg->total_served))
        return fail(MQ_E_INVALID, "mq_sim_fifo: warmup must be in 0..total_served-1");
    if (int rc = check_dist(cfg->src, "set_sources")) return rc;
    if (int rc = check_dist(cfg->srv, "set_servers")) return rc;
    CU(use_device(ctx));

    const long long R = cfg->replications;
    const bool fp64 = (cfg->flags & MQ_FIFO_FP64) != 0;
    const bool busy = (cfg->flags & MQ_FIFO_BUSY) != 0;
    const int rows = MQ_REC_ROWS(cfg->p_bins) + (busy ? MQ_FIFO_BUSY_ROWS : 0);
    const int agg_len = MQ_AGG_LEN(cfg->p_bins) + (busy ? MQ_AGG_BUSY_LEN : 0);
    if (int rc = ensure(ctx->rec, sizeof(double) * (size_t)rows * (size_t)R)) return rc;
    if (int rc = ensure_host_agg(ctx, sizeof(double) * (size_t)agg_len)) return rc;

    K2Params P;
    std::memset(&P, 0, sizeof(P));
    P.c = cfg->c;
    P.p_bins = cfg->p_bins;
    P.buffer = cfg->buffer;
    P.jobs = cfg->total_served;
    P.reps = R;
    P.rep0 = (uint32_t)cfg->rep0;
    P.key0 = seed_to_key(cfg->seed);
    for (int r = 0; r < 10; ++r) P.rk[r] = P.key0 + (uint32_t)r * MQ_PHILOX_W;
    P.warm = (uint32_t)cfg->warmup;
    P.v_at_start = (cfg->flags & MQ_FIFO_V_AT_START) ? 1 : 0;
    P.src = make_distf(cfg->src);
    P.srv = make_distf(cfg->srv);
    P.srcd = make_distd(cfg->src);
    P.srvd = make_distd(cfg->srv);
    P.rec = (double *)ctx->rec.p;

    const bool finite = cfg->buffer >= 0;
    int pc = 0;
    if (cfg->src.kind == MQ_M && cfg->srv.kind == MQ_M) pc = 1;
    if (cfg->src.kind == MQ_M && cfg->srv.kind == MQ_H2) pc = 2;
    if (busy) pc = 0;  // the busy-period variant exists for the generic class only
    if (fp64) pc = 3;
    const size_t real_b = fp64 ? sizeof(double) : sizeof(float);
    const size_t smem = real_b * (size_t)(MQ_BF + ct) * MQ_BLOCK;
    int occ = 0;
    CU(k2_dispatch(ct, P, pc, finite, busy, 0, smem, nullptr, &occ));
    if (occ < 1) return fail(MQ_E_CUDA, "mq_sim_fifo: kernel does not fit on an SM");
    long long blocks_needed = (R + MQ_BLOCK - 1) / MQ_BLOCK;
    long long grid = (long long)occ * ctx->sms;
    if (grid > blocks_needed) grid = blocks_needed;
    P.lanes = grid * MQ_BLOCK;
    if (finite && cfg->buffer > 0) {
        const size_t bytes = 2 * real_b * (size_t)cfg->buffer * (size_t)P.lanes;
        if (bytes > (size_t)48 << 30)
            return fail(MQ_E_UNSUPPORTED, "mq_sim_fifo: buffer=%lld needs %.1f GiB of queue scratch",
                        (long long)cfg->buffer, bytes / 1073741824.0);
        if (int rc = ensure(ctx->ring, bytes)) return rc;
        P.ring = ctx->ring.p;
    }
    if (busy) {
        const int B = MQ_REC_ROWS(cfg->p_bins);
        const RatioDesc d[4] = {{B + MQ_FB_BUSY, B + MQ_FB_BUSYN}, {B + MQ_FB_BUSY + 1, B + MQ_FB_BUSYN},
                                {B + MQ_FB_BUSY + 2, B + MQ_FB_BUSYN}, {B + MQ_FB_BUSYN, -1}};
        if (int rc = ensure(ctx->desc, sizeof(d))) return rc;
        CU(cudaMemcpyAsync(ctx->desc.p, d, sizeof(d), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));  // d is a stack array
    }

    ctx->launches = 0;
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    CU(cudaMemsetAsync(ctx->rec.p, 0, sizeof(double) * (size_t)rows * (size_t)R, ctx->stream));
    CU(k2_dispatch(ct, P, pc, finite, busy, (int)grid, smem, ctx->stream, nullptr));
    ctx->launches++;
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    CU(launch_reduce_records((const double *)ctx->rec.p, R, cfg->p_bins, (double *)ctx->agg.p, ctx->stream));
    ctx->launches++;
    if (busy) {
        CU(launch_reduce_ratio((const double *)ctx->rec.p, R, (const RatioDesc *)ctx->desc.p, 4,
                               (double *)ctx->agg.p + MQ_AGG_LEN(cfg->p_bins), ctx->stream));
        ctx->launches++;
    }
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    CU(cudaMemcpyAsync(ctx->h_agg, ctx->agg.p, sizeof(double) * (size_t)agg_len, cudaMemcpyDeviceToHost, ctx->stream));
    if (rep_out)
        CU(cudaMemcpyAsync(rep_out, ctx->rec.p, sizeof(double) * (size_t)rows * (size_t)R,
                           cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaEventRecord(ctx->ev[3], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    std::memcpy(agg, ctx->h_agg, sizeof(double) * (size_t)agg_len);
    return finish_timing(ctx);
}

int mq_sim_fifo_multi(mq_ctx *const *ctxs, int n_ctx, const mq_fifo_cfg *cfg, double *agg)
{
    if (!ctxs || n_ctx < 1 || !cfg || !agg) return fail(MQ_E_INVALID, "mq_sim_fifo_multi: NULL argument");
    for (int i = 0; i < n_ctx; ++i)
        if (!ctxs[i]) return fail(MQ_E_INVALID, "mq_sim_fifo_multi: context %d is NULL", i);
    if (n_ctx == 1) return mq_sim_fifo(ctxs[0], cfg, agg, nullptr);
    if (cfg->p_bins < 1 || cfg->p_bins > MQ_MAX_PBINS)
        return fail(MQ_E_INVALID, "mq_sim_fifo_multi: p_bins must be in 1..%d", MQ_MAX_PBINS);
    const int len = MQ_AGG_LEN(cfg->p_bins) + ((cfg->flags & MQ_FIFO_BUSY) ? MQ_AGG_BUSY_LEN : 0);
    std::vector<std::vector<double>> part(n_ctx, std::vector<double>(len, 0.0));
    std::vector<int> rc(n_ctx, MQ_OK);
    std::vector<std::string> msg(n_ctx);
    std::vector<std::thread> th;
    const long long R = cfg->replications;
    for (int i = 0; i < n_ctx; ++i) {
        const long long lo = R * i / n_ctx, hi = R * (i + 1) / n_ctx;
        if (hi <= lo) continue;
        th.emplace_back([&, i, lo, hi]() {
            mq_fifo_cfg c = *cfg;
            c.rep0 = cfg->rep0 + lo;
            c.replications = hi - lo;
            rc[i] = mq_sim_fifo(ctxs[i], &c, part[i].data(), nullptr);
            if (rc[i] != MQ_OK) msg[i] = g_err;  // the message is thread-local
        });
    }
    for (auto &t : th) t.join();
    for (int i = 0; i < n_ctx; ++i)
        if (rc[i] != MQ_OK) return fail(rc[i], "mq_sim_fifo_multi: device %d: %s", ctxs[i]->device, msg[i].c_str());
    for (int q = 0; q < len; ++q) {
        double s = 0.0;
        for (int i = 0; i < n_ctx; ++i) s += part[i][q];  // fixed order: the sum does not depend on thread timing
        agg[q] = s;
    }
    // the timing of the call = the slowest device
    double sim = 0, red = 0, tot = 0;
    int launches = 0;
    for (int i = 0; i < n_ctx; ++i) {
        sim = std::fmax(sim, ctxs[i]->sim_ms);
        red = std::fmax(red, ctxs[i]->reduce_ms);
        tot = std::fmax(tot, ctxs[i]->total_ms);
        launches += ctxs[i]->launches;
    }
    ctxs[0]->sim_ms = sim;
    ctxs[0]->reduce_ms = red;
    ctxs[0]->total_ms = tot;
    ctxs[0]->launches = launches;
    return MQ_OK;
}

int mq_replay_fifo(mq_ctx *ctx, const mq_replay_cfg *cfg, const double *A, const double *S, double *agg,
                   double *rep_out)
{
    if (!ctx || !cfg || !A || !S || !rep_out) return fail(MQ_E_INVALID, "mq_replay_fifo: NULL argument");
    if (cfg->c < 1) return fail(MQ_E_INVALID, "mq_replay_fifo: num_of_channels must be >= 1");
    const int ct = ct_for(cfg->c);
    if (!ct) return fail(MQ_E_UNSUPPORTED, "mq_replay_fifo: num_of_channels %d > %d", cfg->c, MQ_MAX_CHANNELS);
    if (cfg->p_bins < 1 || cfg->p_bins > MQ_MAX_PBINS)
        return fail(MQ_E_INVALID, "mq_replay_fifo: p_bins must be in 1..%d", MQ_MAX_PBINS);
    if (cfg->buffer < -1) return fail(MQ_E_INVALID, "mq_replay_fifo: buffer must be >= 0 or -1 (infinite)");
    if (cfg->total_served < 0 || cfg->replications < 1 || cfg->n_a < 1 || cfg->n_s < 0)
        return fail(MQ_E_INVALID, "mq_replay_fifo: bad sizes");
    if (cfg->total_served >= (1ll << 31) || cfg->n_a >= (1ll << 31) || cfg->n_s >= (1ll << 31) ||
        cfg->buffer >= (1ll << 31))
        return fail(MQ_E_UNSUPPORTED, "mq_replay_fifo: total_served, n_a, n_s and buffer must be < 2^31");
    CU(use_device(ctx));

    const long long R = cfg->replications;
    const int rows = MQ_REPLAY_ROWS(cfg->p_bins);
    const size_t rec_bytes = sizeof(double) * (size_t)rows * (size_t)R;
    const size_t a_bytes = sizeof(double) * (size_t)cfg->n_a * (size_t)R;
    const size_t s_bytes = sizeof(double) * (size_t)cfg->n_s * (size_t)R;
    if (int rc = ensure(ctx->rec, rec_bytes)) return rc;
    if (int rc = stage_agg(ctx, cfg->p_bins)) return rc;

    ctx->launches = 0;
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    const double *dA = A, *dS = S;
    if (!cfg->device_ptrs) {
        if (int rc = ensure(ctx->in_a, a_bytes)) return rc;
        if (int rc = ensure(ctx->in_s, s_bytes ? s_bytes : 8)) return rc;
        CU(cudaMemcpyAsync(ctx->in_a.p, A, a_bytes, cudaMemcpyHostToDevice, ctx->stream));
        if (s_bytes) CU(cudaMemcpyAsync(ctx->in_s.p, S, s_bytes, cudaMemcpyHostToDevice, ctx->stream));
        dA = (const double *)ctx->in_a.p;
        dS = (const double *)ctx->in_s.p;
    }

    K1Params P;
    std::memset(&P, 0, sizeof(P));
    P.c = cfg->c;
    P.p_bins = cfg->p_bins;
    P.buffer = cfg->buffer;
    P.jobs = cfg->total_served;
    P.reps = R;
    P.n_a = cfg->n_a;
    P.n_s = cfg->n_s;
    P.A = dA;
    P.S = dS;
    P.rec = (double *)ctx->rec.p;
    const bool finite = cfg->buffer >= 0;
    if (finite && cfg->buffer > 0) {
        if (int rc = ensure(ctx->ring, sizeof(double) * (size_t)cfg->buffer * (size_t)R)) return rc;
        P.ring = (double *)ctx->ring.p;
    }
    CU(cudaMemsetAsync(ctx->rec.p, 0, rec_bytes, ctx->stream));
    // the timed "sim" span starts after the input copies
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    CU(k1_dispatch(ct, P, finite, ctx->stream));
    ctx->launches++;
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    if (agg) {
        CU(launch_reduce_records((const double *)ctx->rec.p, R, cfg->p_bins, (double *)ctx->agg.p, ctx->stream));
        ctx->launches++;
    }
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    if (agg)
        CU(cudaMemcpyAsync(ctx->h_agg, ctx->agg.p, sizeof(double) * MQ_AGG_LEN(cfg->p_bins),
                           cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(rep_out, ctx->rec.p, rec_bytes,
                       cfg->device_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaEventRecord(ctx->ev[3], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (agg) std::memcpy(agg, ctx->h_agg, sizeof(double) * MQ_AGG_LEN(cfg->p_bins));
    return finish_timing(ctx);
}

int mq_last_timing(mq_ctx *ctx, double *sim_ms, double *reduce_ms, double *total_ms)
{
    if (!ctx) return fail(MQ_E_INVALID, "mq_last_timing: NULL context");
    if (sim_ms) *sim_ms = ctx->sim_ms;
    if (reduce_ms) *reduce_ms = ctx->reduce_ms;
    if (total_ms) *total_ms = ctx->total_ms;
    return MQ_OK;
}

int mq_last_launches(mq_ctx *ctx) { return ctx ? ctx->launches : 0; }

int mq_debug_philox2x32(mq_ctx *ctx, uint32_t c0, uint32_t c1, uint32_t key, uint32_t out[2])
{
    if (!ctx || !out) return fail(MQ_E_INVALID, "mq_debug_philox2x32: NULL argument");
    CU(use_device(ctx));
    uint32_t *d = nullptr;
    CU(cudaMalloc(&d, 8));
    debug_philox_kernel<<<1, 1, 0, ctx->stream>>>(c0, c1, key, d);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    CU(e);
    return MQ_OK;
}

int mq_debug_variates(mq_ctx *ctx, const mq_dist *d, uint64_t seed, uint32_t rep, int which, int64_t n,
                      float *out)
{
    if (!ctx || !d || !out || n < 1) return fail(MQ_E_INVALID, "mq_debug_variates: bad argument");
    if (int rc = check_dist(*d, "mq_debug_variates")) return rc;
    CU(use_device(ctx));
    float *dev = nullptr;
    CU(cudaMalloc(&dev, sizeof(float) * (size_t)n));
    debug_variates_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(make_distf(*d), seed_to_key(seed),
                                                                               rep, which ? 1 : 0, n, dev);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dev, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dev);
    CU(e);
    return MQ_OK;
}

}  // extern "C"
