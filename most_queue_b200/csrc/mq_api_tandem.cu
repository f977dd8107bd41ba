// mq_api_tandem.cu -- C ABI entry points of the tandem-with-blocking kernel.
#include "k6_tandem.cuh"
#include "mq_host.cuh"

#include <cstdlib>
#include <vector>

using namespace mq;

static int tandem_common(mq_ctx *ctx, const mq_tandem_cfg *cfg, bool replay, const double *A, int64_t n_a,
                         const double *const *S, const int64_t *n_s, double *agg, double *rep_out, const char *who)
{
    if (!ctx || !cfg || !agg) return fail(MQ_E_INVALID, "%s: NULL argument", who);
    const int m = cfg->m;
    if (m < 1 || m > K6_MAXM)
        return fail(m < 1 ? MQ_E_INVALID : MQ_E_UNSUPPORTED, "%s: number of nodes must be in 1..%d", who, K6_MAXM);
    if (cfg->total_served < 0 || cfg->replications < 1 || cfg->rep0 < 0 || cfg->rep0 + cfg->replications > (1ll << 32))
        return fail(MQ_E_INVALID, "%s: bad sizes", who);
    if (!(cfg->warmup_fraction >= 0.0) || cfg->warmup_fraction > 100.0)
        return fail(MQ_E_INVALID, "%s: warmup_fraction must be >= 0", who);
    if (!replay) {
        if (int rc = check_dist(cfg->src, "set_sources")) return rc;
        for (int i = 0; i < m; ++i)
            if (int rc = check_dist(cfg->srv[i], "set_nodes")) return rc;
    } else if (!A || !S || !n_s || n_a < 1) {
        return fail(MQ_E_INVALID, "%s: NULL or empty arrays", who);
    }
    CU(use_device(ctx));

    const long long R = cfg->replications;
    const int rows = MQ_TANDEM_REC_ROWS(m);
    const size_t rec_bytes = sizeof(double) * (size_t)rows * (size_t)R;
    const int nq = MQ_TANDEM_AGG_LEN(m);
    if (int rc = ensure(ctx->rec, rec_bytes)) return rc;
    if (int rc = ensure_host_agg(ctx, sizeof(double) * 2 * (size_t)nq)) return rc;

    K6Params P;
    std::memset(&P, 0, sizeof(P));
    P.m = m;
    P.replay = replay ? 1 : 0;
    for (int i = 0; i < m; ++i) P.capacity[i] = cfg->capacity[i];
    P.jobs = cfg->total_served;
    P.warm = (long long)((double)cfg->total_served * cfg->warmup_fraction);  // int(total_served * warmup_fraction)
    P.reps = R;
    P.rep0 = (uint32_t)cfg->rep0;
    P.key0 = seed_to_key(cfg->seed);
    P.rec = (double *)ctx->rec.p;

    ctx->launches = 0;
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (replay) {
        size_t ts = 0;
        for (int i = 0; i < m; ++i) {
            if (n_s[i] < 0 || (n_s[i] > 0 && !S[i])) return fail(MQ_E_INVALID, "%s: node %d service array missing", who, i);
            P.soff[i] = (long long)ts;
            P.n_s[i] = n_s[i];
            ts += (size_t)n_s[i] * (size_t)R;
        }
        if (int rc = ensure(ctx->in_a, sizeof(double) * (size_t)n_a * R)) return rc;
        if (int rc = ensure(ctx->in_s, sizeof(double) * (ts ? ts : 1))) return rc;
        CU(cudaMemcpyAsync(ctx->in_a.p, A, sizeof(double) * (size_t)n_a * R, cudaMemcpyHostToDevice, ctx->stream));
        for (int i = 0; i < m; ++i)
            if (n_s[i])
                CU(cudaMemcpyAsync((double *)ctx->in_s.p + P.soff[i], S[i], sizeof(double) * (size_t)n_s[i] * R,
                                   cudaMemcpyHostToDevice, ctx->stream));
        P.A = (const double *)ctx->in_a.p;
        P.S = (const double *)ctx->in_s.p;
        P.n_a = n_a;
    } else {
        P.src = make_distf(cfg->src);
        for (int i = 0; i < m; ++i) P.srv[i] = make_distf(cfg->srv[i]);
    }
    // lines of at most 16 nodes run the predicated, on-chip-state form (k6v2_tandem.cu); MQ_K6_V1=1 forces the general kernel
    const bool v2 = k6v2_supported(m) && !std::getenv("MQ_K6_V1") && P.jobs + P.warm < (1ll << 30) &&
                    (!replay || [&] {
                        for (int i = 0; i < m; ++i)
                            if (n_s[i] >= (1ll << 31)) return false;
                        return true;
                    }());
    int occ = 0;
    if (v2)
        CU(k6v2_occupancy(m, replay, &occ));
    else
        CU(k6_occupancy(&occ));
    if (occ < 1) return fail(MQ_E_CUDA, "%s: kernel does not fit on an SM", who);
    long long grid = (long long)occ * ctx->sms;
    const long long need = (R + K6_BLOCK - 1) / K6_BLOCK;
    if (grid > need) grid = need;
    P.lanes = grid * K6_BLOCK;

    std::vector<RatioDesc> desc((size_t)nq, RatioDesc{MQ_TG_FLAGS, MQ_TG_FLAGS});
    desc[1] = {MQ_TG_FLAGS, MQ_TG_FLAGS};
    desc[2] = {MQ_TG_DEPARTURES, MQ_TG_ELAPSED};
    desc[3] = desc[2];
    desc[4] = {MQ_TG_LOST, MQ_TG_ARRIVED};
    desc[5] = desc[4];
    desc[6] = {MQ_TG_AREASUM, MQ_TG_DEPARTURES};
    desc[7] = desc[6];
    for (int i = 0; i < m; ++i) {
        desc[8 + 2 * i] = {MQ_TG_ROWS + i * MQ_TN_ROWS + MQ_TN_AREA, MQ_TG_ELAPSED};
        desc[9 + 2 * i] = desc[8 + 2 * i];
    }
    if (int rc = ensure(ctx->desc, sizeof(RatioDesc) * (size_t)nq)) return rc;
    CU(cudaMemcpyAsync(ctx->desc.p, desc.data(), sizeof(RatioDesc) * (size_t)nq, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->rec.p, 0, rec_bytes, ctx->stream));
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (v2)
        CU(k6v2_launch(P, (int)grid, ctx->stream));
    else
        CU(k6_launch(P, (int)grid, ctx->stream));
    ctx->launches++;
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    CU(launch_reduce_ratio((const double *)ctx->rec.p, R, (const RatioDesc *)ctx->desc.p, nq, (double *)ctx->agg.p,
                           ctx->stream));
    ctx->launches++;
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    CU(cudaMemcpyAsync(ctx->h_agg, ctx->agg.p, sizeof(double) * 2 * (size_t)nq, cudaMemcpyDeviceToHost, ctx->stream));
    if (rep_out) CU(cudaMemcpyAsync(rep_out, ctx->rec.p, rec_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaEventRecord(ctx->ev[3], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    const double *sum = ctx->h_agg, *sq = ctx->h_agg + nq;
    agg[0] = (double)R;
    agg[1] = sum[1];
    for (int q = 2; q < nq; q += 2) {
        agg[q] = sum[q];
        agg[q + 1] = sq[q];
    }
    return finish_timing(ctx);
}

extern "C" {

int mq_sim_tandem(mq_ctx *ctx, const mq_tandem_cfg *cfg, double *agg, double *rep_out)
{
    return tandem_common(ctx, cfg, false, nullptr, 0, nullptr, nullptr, agg, rep_out, "mq_sim_tandem");
}

int mq_replay_tandem(mq_ctx *ctx, const mq_tandem_cfg *cfg, const double *A, int64_t n_a, const double *const *S,
                     const int64_t *n_s, double *agg, double *rep_out)
{
    return tandem_common(ctx, cfg, true, A, n_a, S, n_s, agg, rep_out, "mq_replay_tandem");
}

}  // extern "C"
