// mq_api_negatives.cu -- C ABI entry points of the negative-arrivals queue kernel.
#include "k9_negatives.cuh"
#include "mq_host.cuh"

#include <vector>

using namespace mq;

static int negatives_common(mq_ctx *ctx, const mq_neg_cfg *cfg, bool replay, const double *const *X, const int64_t *n,
                            double *agg, double *rep_out, const char *who)
{
    if (!ctx || !cfg || !agg) return fail(MQ_E_INVALID, "%s: NULL argument", who);
    if (cfg->c < 1) return fail(MQ_E_INVALID, "%s: num_of_channels must be >= 1", who);
    if (cfg->c > K9_MAXC) return fail(MQ_E_UNSUPPORTED, "%s: num_of_channels %d > %d", who, cfg->c, K9_MAXC);
    if (cfg->type != MQ_NEG_DISASTER && cfg->type != MQ_NEG_RCS)
        return fail(MQ_E_UNSUPPORTED, "%s: only the DISASTER and RCS types of negatives are implemented", who);
    if (cfg->scenario < 0 || cfg->scenario > MQ_NEG_REQUEUE_FIXED) return fail(MQ_E_INVALID, "%s: unknown scenario", who);
    if (cfg->p_bins < 1 || cfg->p_bins > MQ_MAX_PBINS) return fail(MQ_E_INVALID, "%s: p_bins must be in 1..%d", who, MQ_MAX_PBINS);
    if (cfg->total_served < 0 || cfg->replications < 1 || cfg->rep0 < 0 || cfg->rep0 + cfg->replications > (1ll << 32))
        return fail(MQ_E_INVALID, "%s: bad sizes", who);
    if (!replay) {
        if (cfg->total_served >= (1ll << 31)) return fail(MQ_E_INVALID, "%s: total_served must be < 2^31", who);
        if (int rc = check_dist(cfg->positive, "set_positive_sources")) return rc;
        if (int rc = check_dist(cfg->negative, "set_negative_sources")) return rc;
        if (int rc = check_dist(cfg->service, "set_servers")) return rc;
    } else if (!X || !n) {
        return fail(MQ_E_INVALID, "%s: NULL arrays", who);
    }
    CU(use_device(ctx));

    const long long R = cfg->replications;
    const int pb = cfg->p_bins;
    const size_t rec_bytes = sizeof(double) * (size_t)MQ_N_ROWS(pb) * (size_t)R;
    const int nq = 24 + pb;
    if (int rc = ensure(ctx->rec, rec_bytes)) return rc;
    if (int rc = ensure_host_agg(ctx, sizeof(double) * 2 * (size_t)nq)) return rc;

    K9Params P;
    std::memset(&P, 0, sizeof(P));
    P.c = cfg->c;
    P.p_bins = pb;
    P.replay = replay ? 1 : 0;
    P.type = cfg->type;
    P.scenario = cfg->scenario == 0 ? MQ_NEG_REMOVE : cfg->scenario;
    P.buffer = cfg->buffer < 0 ? -1 : cfg->buffer;
    P.jobs = cfg->total_served;
    P.reps = R;
    P.rep0 = (uint32_t)cfg->rep0;
    P.key0 = seed_to_key(cfg->seed);
    P.rec = (double *)ctx->rec.p;
    long long qcap = cfg->queue_cap > 0 ? cfg->queue_cap : 1024;
    if (cfg->buffer >= 0) qcap = cfg->buffer > 0 ? cfg->buffer : 1;
    if (qcap > (1 << 20)) return fail(MQ_E_UNSUPPORTED, "%s: waiting line of %lld jobs per replication", who, qcap);
    P.qcap = (int)qcap;

    ctx->launches = 0;
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (replay) {
        size_t tot = 0;
        for (int s = 0; s < MQ_NEG_STREAMS; ++s) {
            const long long ns = X[s] ? n[s] : 0;
            if (ns < 0) return fail(MQ_E_INVALID, "%s: stream %d has a negative length", who, s);
            if (s != MQ_NS_CHOICES && ns < 1) return fail(MQ_E_INVALID, "%s: stream %d is empty", who, s);
            P.off[s] = (long long)tot;
            P.n[s] = ns;
            tot += (size_t)ns * (size_t)R;
        }
        if (int rc = ensure(ctx->in_a, sizeof(double) * tot)) return rc;
        for (int s = 0; s < MQ_NEG_STREAMS; ++s)
            if (P.n[s])
                CU(cudaMemcpyAsync((double *)ctx->in_a.p + P.off[s], X[s], sizeof(double) * (size_t)P.n[s] * R,
                                   cudaMemcpyHostToDevice, ctx->stream));
        P.X = (const double *)ctx->in_a.p;
    } else {
        P.positive = make_distf(cfg->positive);
        P.negative = make_distf(cfg->negative);
        P.service = make_distf(cfg->service);
    }
    int occ = 0;
    CU(k9_occupancy(cfg->c, &occ));
    if (occ < 1) return fail(MQ_E_CUDA, "%s: kernel does not fit on an SM", who);
    long long grid = (long long)occ * ctx->sms;
    const long long want = (R + K9_BLOCK - 1) / K9_BLOCK;
    if (grid > want) grid = want;
    P.lanes = grid * K9_BLOCK;
    if (int rc = ensure(ctx->ring, sizeof(double) * (size_t)P.qcap * (size_t)P.lanes)) return rc;
    P.queue = (double *)ctx->ring.p;

    std::vector<RatioDesc> desc((size_t)nq);
    desc[0] = {MQ_N_TTEK, -1};
    desc[1] = {MQ_N_FLAGS, MQ_N_FLAGS};
    desc[2] = {MQ_N_SERVED, -1};
    desc[3] = {MQ_N_BROKEN, -1};
    desc[4] = {MQ_N_TOTAL, -1};
    desc[5] = {MQ_N_POS_ARRIVED, -1};
    desc[6] = {MQ_N_DROPPED, -1};
    desc[7] = {MQ_N_SERVED, MQ_N_TOTAL};
    for (int i = 0; i < 4; ++i) {
        desc[8 + i] = {MQ_N_WSUM + i, MQ_N_TAKED};
        desc[12 + i] = {MQ_N_VSUM + i, MQ_N_TOTAL};
        desc[16 + i] = {MQ_N_VSSUM + i, MQ_N_SERVED};
        desc[20 + i] = {MQ_N_VBSUM + i, MQ_N_BROKEN};
    }
    for (int b = 0; b < pb; ++b) desc[24 + b] = {MQ_N_P + b, MQ_N_TTEK};
    if (int rc = ensure(ctx->desc, sizeof(RatioDesc) * (size_t)nq)) return rc;
    CU(cudaMemcpyAsync(ctx->desc.p, desc.data(), sizeof(RatioDesc) * (size_t)nq, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->rec.p, 0, rec_bytes, ctx->stream));
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    CU(k9_launch(P, (int)grid, ctx->stream));
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
    for (int q = 0; q < MQ_N_AGG_LEN(pb); ++q) agg[q] = 0.0;
    agg[0] = (double)R;
    agg[1] = sum[0];
    agg[2] = sum[1];
    for (int i = 0; i < 5; ++i) agg[3 + i] = sum[2 + i];
    for (int g = 0; g < 4; ++g)
        for (int i = 0; i < 4; ++i) {
            agg[8 + 8 * g + i] = sum[8 + 4 * g + i];
            agg[12 + 8 * g + i] = sq[8 + 4 * g + i];
        }
    agg[40] = sum[7];
    agg[41] = sq[7];
    for (int b = 0; b < pb; ++b) {
        agg[48 + b] = sum[24 + b];
        agg[48 + pb + b] = sq[24 + b];
    }
    return finish_timing(ctx);
}

extern "C" {

int mq_sim_negatives(mq_ctx *ctx, const mq_neg_cfg *cfg, double *agg, double *rep_out)
{
    return negatives_common(ctx, cfg, false, nullptr, nullptr, agg, rep_out, "mq_sim_negatives");
}

int mq_replay_negatives(mq_ctx *ctx, const mq_neg_cfg *cfg, const double *const *X, const int64_t *n, double *agg,
                        double *rep_out)
{
    return negatives_common(ctx, cfg, true, X, n, agg, rep_out, "mq_replay_negatives");
}

}  // extern "C"
