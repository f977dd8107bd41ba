// mq_api_scan.cu -- C ABI entry points of the Lindley scan (mq_scan_gg1*).
#include "k3_lindley.cuh"
#include "mq_host.cuh"

#include <cstdlib>

using namespace mq;

namespace {

struct ScanRun {
    K3Params P;
    int grid_sum = 0, grid_walk = 0, grid_rep = 0;
    bool staged_inputs = false;
    bool spec = false;     // generator form, speculative walk inside the summary pass + k3_repair (no stash, no walk pass)
    bool onepass = false;  // replay form in one pass (k3_onepass.cu): no summary launches, the walk is the fused kernel
};

// gaps past the range that a host-pointer replay stages for the time-in-state level walk (a job's sojourn would have to
// span more than this many later arrivals to notice)
constexpr long long SCAN_AHEAD_STAGED = 1ll << 22;

struct ScanKey {
    long long jobs, job0, jobs_ahead;
    unsigned long long seed;
    unsigned chain;
    int device_ptrs, flags;
    mq_dist src, srv;
    const double *A, *S;
};
static_assert(sizeof(ScanKey) <= 256, "ScanKey must fit mq_ctx::scan_key");

ScanKey make_key(const mq_scan_cfg *cfg, const double *A, const double *S)
{
    ScanKey k;
    std::memset(&k, 0, sizeof(k));
    k.jobs = cfg->jobs;
    k.job0 = cfg->job0;
    k.jobs_ahead = cfg->jobs_ahead;
    k.seed = cfg->seed;
    k.chain = cfg->chain;
    k.device_ptrs = cfg->device_ptrs;
    k.flags = cfg->flags & ~MQ_SCAN_KEEP;
    if (!A && !S) {
        k.src = cfg->src;
        k.srv = cfg->srv;
    }
    k.A = A;
    k.S = S;
    return k;
}

int scan_prepare(mq_ctx *ctx, const mq_scan_cfg *cfg, const double *A, const double *S, ScanRun &run, const char *who,
                 bool will_walk = false, bool allow_spec = false)
{
    if (!ctx || !cfg) return fail(MQ_E_INVALID, "%s: NULL argument", who);
    if (cfg->jobs < 1 || cfg->job0 < 0) return fail(MQ_E_INVALID, "%s: jobs must be >= 1 and job0 >= 0", who);
    const bool replay = A != nullptr || S != nullptr;
    if (replay && (!A || !S)) return fail(MQ_E_INVALID, "%s: both A and S are needed in replay form", who);
    if (!replay) {
        if (int rc = check_dist(cfg->src, "set_sources")) return rc;
        if (int rc = check_dist(cfg->srv, "set_servers")) return rc;
    }
    CU(use_device(ctx));
    K3Params &P = run.P;
    std::memset(&P, 0, sizeof(P));
    P.n = cfg->jobs;
    P.job0 = cfg->job0;
    P.replay = replay ? 1 : 0;
    P.key = seed_to_key(cfg->seed) + MQ_KEY_STRIDE * (64u * cfg->chain);
    for (int r = 0; r < 10; ++r) P.rk[r] = P.key + (uint32_t)r * MQ_PHILOX_W;
    const long long tile = (long long)K3_THREADS * K3_L;
    P.ntiles = (P.n + tile - 1) / tile;

    const bool want_p = (cfg->flags & MQ_SCAN_WANT_P) != 0;
    P.n_ahead = cfg->jobs_ahead > 0 ? cfg->jobs_ahead : 0;
    if (replay) {
        if (cfg->device_ptrs) {
            P.A = A;
            P.S = S;
        } else {
            // host arrays: the walk reads gaps past A[jobs] only for the time-in-state levels; stage what it may read and
            // tell the kernels how much that is (reading past the staged buffer was an out-of-bounds read)
            const long long ahead = want_p ? std::min<long long>(P.n_ahead, SCAN_AHEAD_STAGED) : 0;
            P.n_ahead = ahead;
            const size_t na = (size_t)(P.n + 1 + ahead);
            if (int rc = ensure(ctx->in_a, sizeof(double) * na)) return rc;
            if (int rc = ensure(ctx->in_s, sizeof(double) * (size_t)P.n)) return rc;
            CU(cudaMemcpyAsync(ctx->in_a.p, A, sizeof(double) * na, cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaMemcpyAsync(ctx->in_s.p, S, sizeof(double) * (size_t)P.n, cudaMemcpyHostToDevice, ctx->stream));
            P.A = (const double *)ctx->in_a.p;
            P.S = (const double *)ctx->in_s.p;
        }
    } else {
        P.src = make_distf(cfg->src);
        P.srv = make_distf(cfg->srv);
    }
    // replay form followed by a walk: one pass over A and S when the arrays can be fetched by TMA (16-byte aligned) and
    // no time-in-state levels are wanted; MQ_SCAN_TWO_PASS=1 keeps the three-launch path (the checker of the fused one)
    if (replay && will_walk && !want_p && P.ntiles >= 2 && !std::getenv("MQ_SCAN_TWO_PASS")) {
        K3Params probe = P;
        probe.level_partial = nullptr;
        int occ1 = 0;
        run.onepass = k3_onepass_usable(probe) && k3_onepass_occupancy(&occ1) == cudaSuccess && occ1 >= 1;
        cudaGetLastError();
    }

    // generator form followed by a walk: stash the fp32 draws of the summary pass (8 B per job) when they fit, so the
    // walk does not draw everything a second time.  MQ_SCAN_NO_STASH=1 turns it off.
    bool stash = false;
    if (!replay && will_walk && !std::getenv("MQ_SCAN_NO_STASH")) {
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        const size_t need = sizeof(float) * (2 * (size_t)P.n + 64);
        const size_t have = ctx->in_a.cap + ctx->in_s.cap;  // buffers that would be reused
        stash = need <= have || need < (free_b + have) / 2;
    }
    // Speculative mode: the walk is done inside the summary pass (every segment from its in-tile prefix) and k3_repair
    // re-draws the few segments that depended on the carry -- no stash, no walk pass, 1 B/job of HBM instead of 8.  On a
    // chain whose stash fits it is 5 % SLOWER than summary + walk (the summary pass is issue bound and the fp32 -> fp64
    // conversions of the walk share the MUFU pipe with the samplers), so it is used when the stash does not fit (chains
    // beyond ~1e10 jobs per GPU, which would otherwise be drawn twice) or when MQ_SCAN_SPEC=1 asks for it.
    const bool spec_wanted = !replay && will_walk && allow_spec && !want_p && !std::getenv("MQ_SCAN_NO_SPEC") &&
                             (!stash || std::getenv("MQ_SCAN_SPEC"));
    if (spec_wanted) {
        stash = false;
        // a stash left behind by an earlier, shorter chain is dead weight now: give its memory to the segment prefixes
        if (ctx->in_a.cap + ctx->in_s.cap > ((size_t)1 << 30)) {
            CU(cudaStreamSynchronize(ctx->stream));
            for (Buf *b : {&ctx->in_a, &ctx->in_s}) {
                if (b->p) cudaFree(b->p);
                b->p = nullptr;
                b->cap = 0;
            }
        }
    }
    if (stash) {
        // an allocation failure only costs the optimisation
        if (ensure(ctx->in_a, sizeof(float) * ((size_t)P.n + 32)) != MQ_OK ||
            ensure(ctx->in_s, sizeof(float) * ((size_t)P.n + 32)) != MQ_OK) {
            cudaGetLastError();
            stash = false;
        } else {
            P.stG = (float *)ctx->in_a.p;
            P.stV = (float *)ctx->in_s.p;
        }
    }
    // the summary pass can keep every segment's exclusive prefix (4 KB per tile) so that the walk skips its own fold and
    // CTA scan; MQ_SCAN_NO_PREFIX=1 turns it off
    if (will_walk && !run.onepass && (spec_wanted || !std::getenv("MQ_SCAN_NO_PREFIX"))) {
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        const size_t need = sizeof(MP) * (size_t)K3_THREADS * (size_t)P.ntiles;
        // (the speculative mode lives on these prefixes and keeps nothing else of the chain: it may take most of the HBM)
        if (need <= ctx->agg.cap || need < (free_b + ctx->agg.cap) / 4 * (spec_wanted ? 3 : 1)) {
            if (ensure(ctx->agg, need) == MQ_OK)
                P.thread_ex = (MP *)ctx->agg.p;
            else
                cudaGetLastError();
        }
    }
    run.spec = spec_wanted && P.thread_ex != nullptr;  // (without room for the segment prefixes: draw twice, as before)
    P.spec = run.spec ? 1 : 0;
    int occ_s = 0, occ_w = 0;
    CU(k3_occupancy(replay, run.spec ? 2 : (stash ? 1 : 0), want_p, P.src.kind, P.srv.kind, &occ_s, &occ_w));
    if (run.onepass) {
        occ_w = 1;  // one CTA per SM by design
        if (int rc = ensure(ctx->lookback, k3_onepass_scratch_bytes(P.ntiles))) return rc;
    }
    if (occ_s < 1 || occ_w < 1) return fail(MQ_E_CUDA, "%s: kernel does not fit on an SM", who);
    run.grid_sum = (int)std::min<long long>(P.ntiles, (long long)occ_s * ctx->sms);
    run.grid_walk = (int)std::min<long long>(P.ntiles, (long long)occ_w * ctx->sms);
    run.grid_rep = run.spec ? k3_repair_grid(P.ntiles, ctx->sms) : 0;

    // scratch: tile_sum, tile_ex, total, tail, partial (speculative mode: the summary's blocks, then the repair's), stats
    const size_t nblocks = run.spec ? (size_t)run.grid_sum + (size_t)run.grid_rep : (size_t)run.grid_walk;
    const size_t bytes = sizeof(MP) * (2 * (size_t)P.ntiles + 1 + 2 * K3_TOP) +
                         sizeof(double) * (2 + nblocks * (K3_NSTAT + K3_NLEVEL) + K3_NSTAT + K3_NLEVEL);
    if (int rc = ensure(ctx->rec, bytes)) return rc;
    char *base = (char *)ctx->rec.p;
    P.tile_sum = (MP *)base;
    P.tile_ex = P.tile_sum + P.ntiles;
    P.total = P.tile_ex + P.ntiles;
    P.chunk_f = P.total + 1;
    P.chunk_ex = P.chunk_f + K3_TOP;
    P.tail = (double *)(P.chunk_ex + K3_TOP);
    P.partial = P.tail + 2;
    P.spec_partial = P.partial;
    P.level_partial = want_p ? P.partial + (size_t)run.grid_walk * K3_NSTAT + K3_NSTAT + K3_NLEVEL : nullptr;
    if (int rc = ensure_host_agg(ctx, sizeof(double) * 64)) return rc;
    return MQ_OK;
}

int scan_summary(mq_ctx *ctx, ScanRun &run)
{
    if (run.onepass) return MQ_OK;  // the fused kernel forms and consumes the summaries itself
    CU(k3_launch_summary(run.P, run.grid_sum, ctx->stream));
    CU(k3_launch_top(run.P, ctx->stream));
    ctx->launches += 4;
    return MQ_OK;
}

int scan_apply(mq_ctx *ctx, const mq_scan_cfg *cfg, ScanRun &run, double w_in, double *stats, double *W)
{
    K3Params &P = run.P;
    P.w_in = w_in;
    double *dW = nullptr;
    if (W) {
        if (cfg->device_ptrs) {
            dW = W;
        } else {
            if (int rc = ensure(ctx->ring, sizeof(double) * (size_t)P.n)) return rc;
            dW = (double *)ctx->ring.p;
        }
    }
    P.W = dW;
    const int nblocks = run.spec ? run.grid_sum + run.grid_rep : run.grid_walk;
    double *d_stats = P.partial + (size_t)nblocks * K3_NSTAT;
    if (run.spec) {
        if (W) return fail(MQ_E_INVALID, "mq_scan_gg1: per-job waits need the walk pass (internal: speculative run with W)");
        K3Params R = P;  // the repair's blocks write their sums behind the summary's
        R.partial = P.partial + (size_t)run.grid_sum * K3_NSTAT;
        CU(k3_launch_repair(R, run.grid_rep, ctx->stream));
    } else if (run.onepass) {
        CU(k3_launch_onepass(P, ctx->lookback.p, run.grid_walk, ctx->stream));
    } else {
        CU(k3_launch_walk(P, run.grid_walk, ctx->stream));
    }
    CU(k3_launch_final(P.partial, P.level_partial, nblocks, d_stats, ctx->stream));
    ctx->launches += 2;
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    double *h = ctx->h_agg;
    CU(cudaMemcpyAsync(h, d_stats, sizeof(double) * (K3_NSTAT + K3_NLEVEL), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(h + 52, P.total, sizeof(MP), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(h + 54, P.tail, sizeof(double) * 2, cudaMemcpyDeviceToHost, ctx->stream));
    if (W && !cfg->device_ptrs)
        CU(cudaMemcpyAsync(W, dW, sizeof(double) * (size_t)P.n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaEventRecord(ctx->ev[3], ctx->stream));
    if (cudaError_t e_sync = cudaStreamSynchronize(ctx->stream)) {
        if (run.onepass && k3_onepass_error())
            return fail(MQ_E_CUDA, "mq_scan_gg1: the one-pass scan kernel gave up on a wait (code 0x%x): %s; the CUDA context is "
                        "lost, MQ_SCAN_TWO_PASS=1 selects the three-launch path", k3_onepass_error(), cudaGetErrorString(e_sync));
        return fail(MQ_E_CUDA, "cudaStreamSynchronize failed: %s (%s:%d)", cudaGetErrorString(e_sync), __FILE__, __LINE__);
    }

    if (run.onepass) k3_onepass_error();  // (debug builds print their stage checks here)
    for (int i = 0; i < MQ_SCAN_LEN; ++i) stats[i] = 0.0;
    for (int k = 0; k < 4; ++k) {
        stats[MQ_S_W + k] = h[k];
        stats[MQ_S_V + k] = h[4 + k];
    }
    stats[MQ_S_SUMA] = h[8];
    stats[MQ_S_SUMS] = h[9];
    stats[MQ_S_TAKED] = (double)P.n;
    stats[MQ_S_SERVED] = (double)P.n;
    stats[MQ_S_A] = h[52];
    stats[MQ_S_B] = h[53];
    stats[MQ_S_WIN] = w_in;
    stats[MQ_S_WOUT] = std::fmax(h[52], w_in + h[53]);
    stats[MQ_S_TAILRAW] = h[54];
    stats[MQ_S_TAILV] = h[55];
    for (int m = 0; m < K3_NLEVEL; ++m) stats[MQ_S_LEVELS + m] = P.level_partial ? h[K3_NSTAT + m] : 0.0;
    if (cfg->last_range && h[54] >= 0.0) {
        // job N arrived by the N-th departure: popped there and sampled (base.py:300-305)
        double pw = h[54];
        for (int k = 0; k < 4; ++k) {
            stats[MQ_S_W + k] += pw;
            pw *= h[54];
        }
        stats[MQ_S_TAKED] += 1.0;
    }
    return finish_timing(ctx);
}

}  // namespace

extern "C" {

int mq_scan_gg1_summary(mq_ctx *ctx, const mq_scan_cfg *cfg, const double *A, const double *S, double ab[2])
{
    if (!ab) return fail(MQ_E_INVALID, "mq_scan_gg1_summary: NULL output");
    ScanRun run;
    // generator form only: a replayed range is read once by the summary and once by the (one-pass) apply either way
    const bool keep = cfg && (cfg->flags & MQ_SCAN_KEEP) != 0 && !A && !S;
    if (int rc = scan_prepare(ctx, cfg, A, S, run, "mq_scan_gg1_summary", keep, keep)) return rc;
    ctx->launches = 0;
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (int rc = scan_summary(ctx, run)) return rc;
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    CU(cudaMemcpyAsync(ctx->h_agg, run.P.total, sizeof(MP), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaEventRecord(ctx->ev[3], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ab[0] = ctx->h_agg[0];
    ab[1] = ctx->h_agg[1];
    if (keep) {
        static_assert(sizeof(ScanRun) <= sizeof(ctx->scan_cache), "ScanRun must fit mq_ctx::scan_cache");
        const ScanKey key = make_key(cfg, A, S);
        std::memcpy(ctx->scan_cache, &run, sizeof(run));
        std::memcpy(ctx->scan_key, &key, sizeof(key));
        ctx->scan_valid = true;
    }
    return finish_timing(ctx);
}

int mq_scan_gg1_apply(mq_ctx *ctx, const mq_scan_cfg *cfg, const double *A, const double *S, double w_in,
                      double *stats, double *W)
{
    if (!stats) return fail(MQ_E_INVALID, "mq_scan_gg1_apply: NULL output");
    ScanRun run;
    if (ctx && cfg && ctx->scan_valid) {
        // the previous call on this context was the summary of this very range with MQ_SCAN_KEEP: walk from what it left
        const ScanKey key = make_key(cfg, A, S);
        ScanRun cached;
        std::memcpy(&cached, ctx->scan_cache, sizeof(cached));
        // (a kept speculative run has no walk pass to take per-job waits from: redo the range then)
        if (std::memcmp(ctx->scan_key, &key, sizeof(key)) == 0 && !(cached.spec && W)) {
            run = cached;
            ctx->scan_valid = false;
            CU(cudaSetDevice(ctx->device));
            ctx->launches = 0;
            CU(cudaEventRecord(ctx->ev[0], ctx->stream));
            return scan_apply(ctx, cfg, run, w_in, stats, W);
        }
    }
    if (int rc = scan_prepare(ctx, cfg, A, S, run, "mq_scan_gg1_apply", true, W == nullptr)) return rc;
    ctx->launches = 0;
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    // the tile prefixes are a function of the range only; recompute them (scratch is not kept between calls)
    if (int rc = scan_summary(ctx, run)) return rc;
    return scan_apply(ctx, cfg, run, w_in, stats, W);
}

int mq_scan_gg1(mq_ctx *ctx, const mq_scan_cfg *cfg, const double *A, const double *S, double *stats, double *W)
{
    if (!stats) return fail(MQ_E_INVALID, "mq_scan_gg1: NULL output");
    ScanRun run;
    if (int rc = scan_prepare(ctx, cfg, A, S, run, "mq_scan_gg1", true, W == nullptr)) return rc;
    ctx->launches = 0;
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (int rc = scan_summary(ctx, run)) return rc;
    mq_scan_cfg c2 = *cfg;
    c2.last_range = 1;
    return scan_apply(ctx, &c2, run, 0.0, stats, W);
}

}  // extern "C"
