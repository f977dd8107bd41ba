// mq_api_vacation.cu -- C ABI entry points of the vacation / N-policy queue kernel.
#include "k8_vacation.cuh"
#include "mq_host.cuh"

#include <vector>

using namespace mq;

static int vacation_common(mq_ctx *ctx, const mq_vac_cfg *cfg, bool replay, const double *const *X, const int64_t *n,
                           double *agg, double *rep_out, const char *who)
{
    if (!ctx || !cfg || !agg) return fail(MQ_E_INVALID, "%s: NULL argument", who);
    if (cfg->c < 1) return fail(MQ_E_INVALID, "%s: num_of_channels must be >= 1", who);
    if (cfg->c > K8_MAXC) return fail(MQ_E_UNSUPPORTED, "%s: num_of_channels %d > %d", who, cfg->c, K8_MAXC);
    if (cfg->p_bins < 1 || cfg->p_bins > MQ_MAX_PBINS) return fail(MQ_E_INVALID, "%s: p_bins must be in 1..%d", who, MQ_MAX_PBINS);
    if (cfg->total_served < 0 || cfg->replications < 1 || cfg->rep0 < 0 || cfg->rep0 + cfg->replications > (1ll << 32))
        return fail(MQ_E_INVALID, "%s: bad sizes", who);
    if (cfg->big_n < 0) return fail(MQ_E_INVALID, "%s: N must be >= 1", who);  // vacations.py:340
    // which streams the configuration can consume.  set_warm() puts the law on the system-level warm-up period and,
    // with is_service_on_warm_up, on every server as well (vacations.py:64-76)
    const bool warm_any = cfg->set[MQ_VS_WARM] != 0 || cfg->set[MQ_VS_WARM_SERVICES] != 0;
    const bool cold_any = cfg->set[MQ_VS_COLD] != 0 || cfg->big_n > 0;
    bool need[MQ_VAC_STREAMS];
    need[MQ_VS_ARRIVALS] = need[MQ_VS_SERVICES] = true;
    need[MQ_VS_WARM_SERVICES] = warm_any && cfg->service_on_warm_up != 0;
    need[MQ_VS_WARM] = warm_any && (!cfg->service_on_warm_up || cold_any);  // after a cold period: vacations.py:272
    need[MQ_VS_COLD] = cfg->set[MQ_VS_COLD] != 0 && cfg->big_n == 0;
    need[MQ_VS_DELAY] = cfg->set[MQ_VS_DELAY] != 0;
    if (cfg->set[MQ_VS_DELAY] != 0 && !cold_any)
        return fail(MQ_E_INVALID, "You must first set the cooling time. Use the set_cold() method.");  // vacations.py:99
    if (!replay) {
        if (cfg->total_served >= (1ll << 31)) return fail(MQ_E_INVALID, "%s: total_served must be < 2^31", who);
        static const char *where[MQ_VAC_STREAMS] = {"set_sources", "set_servers", "set_warm", "set_warm", "set_cold",
                                                    "set_cold_delay"};
        for (int s = 0; s < MQ_VAC_STREAMS; ++s)
            if (need[s])
                if (int rc = check_dist(cfg->dist[s], where[s])) return rc;
    } else if (!X || !n) {
        return fail(MQ_E_INVALID, "%s: NULL arrays", who);
    }
    CU(use_device(ctx));

    const long long R = cfg->replications;
    const int pb = cfg->p_bins;
    const int rows = MQ_V_ROWS(pb);
    const size_t rec_bytes = sizeof(double) * (size_t)rows * (size_t)R;
    const int nq = 17 + pb;
    if (int rc = ensure(ctx->rec, rec_bytes)) return rc;
    if (int rc = ensure_host_agg(ctx, sizeof(double) * 2 * (size_t)nq)) return rc;

    K8Params P;
    std::memset(&P, 0, sizeof(P));
    P.c = cfg->c;
    P.p_bins = pb;
    P.replay = replay ? 1 : 0;
    for (int s = 0; s < MQ_VAC_STREAMS; ++s) P.set[s] = cfg->set[s];
    P.service_on_warm_up = cfg->service_on_warm_up;
    P.multiple_vacations = cfg->multiple_vacations;
    P.buffer = cfg->buffer < 0 ? -1 : cfg->buffer;
    P.big_n = cfg->big_n;
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
        for (int s = 0; s < MQ_VAC_STREAMS; ++s) {
            const long long ns = X[s] ? n[s] : 0;  // a stream the run never consumes may be absent
            if (ns < 0 || (ns > 0 && !X[s])) return fail(MQ_E_INVALID, "%s: stream %d array missing", who, s);
            if (s <= MQ_VS_SERVICES && ns < 1) return fail(MQ_E_INVALID, "%s: empty arrival / service arrays", who);
            P.off[s] = (long long)tot;
            P.n[s] = ns;
            tot += (size_t)ns * (size_t)R;
        }
        if (int rc = ensure(ctx->in_a, sizeof(double) * tot)) return rc;
        for (int s = 0; s < MQ_VAC_STREAMS; ++s)
            if (P.n[s])
                CU(cudaMemcpyAsync((double *)ctx->in_a.p + P.off[s], X[s], sizeof(double) * (size_t)P.n[s] * R,
                                   cudaMemcpyHostToDevice, ctx->stream));
        P.X = (const double *)ctx->in_a.p;
    } else {
        for (int s = 0; s < MQ_VAC_STREAMS; ++s)
            if (need[s]) P.dist[s] = make_distf(cfg->dist[s]);
    }
    int occ = 0;
    CU(k8_occupancy(cfg->c, &occ));
    if (occ < 1) return fail(MQ_E_CUDA, "%s: kernel does not fit on an SM", who);
    long long grid = (long long)occ * ctx->sms;
    const long long want = (R + K8_BLOCK - 1) / K8_BLOCK;
    if (grid > want) grid = want;
    P.lanes = grid * K8_BLOCK;
    if (int rc = ensure(ctx->ring, sizeof(double) * (size_t)P.qcap * (size_t)P.lanes)) return rc;
    P.queue = (double *)ctx->ring.p;

    std::vector<RatioDesc> desc((size_t)nq);
    desc[0] = {MQ_V_TTEK, -1};
    desc[1] = {MQ_V_FLAGS, MQ_V_FLAGS};
    desc[2] = {MQ_V_SERVED, -1};
    desc[3] = {MQ_V_ARRIVED, -1};
    desc[4] = {MQ_V_DROPPED, -1};
    desc[5] = {MQ_V_TAKED, -1};
    for (int i = 0; i < 4; ++i) {
        desc[6 + i] = {MQ_V_WSUM + i, MQ_V_TAKED};
        desc[10 + i] = {MQ_V_VSUM + i, MQ_V_SERVED};
    }
    desc[14] = {MQ_V_WARMPROB, MQ_V_TTEK};
    desc[15] = {MQ_V_COLDPROB, MQ_V_TTEK};
    desc[16] = {MQ_V_DELAYPROB, MQ_V_TTEK};
    for (int b = 0; b < pb; ++b) desc[17 + b] = {MQ_V_P + b, MQ_V_TTEK};
    if (int rc = ensure(ctx->desc, sizeof(RatioDesc) * (size_t)nq)) return rc;
    CU(cudaMemcpyAsync(ctx->desc.p, desc.data(), sizeof(RatioDesc) * (size_t)nq, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->rec.p, 0, rec_bytes, ctx->stream));
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    CU(k8_launch(P, (int)grid, ctx->stream));
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
    for (int q = 0; q < MQ_V_AGG_LEN(pb); ++q) agg[q] = 0.0;
    agg[0] = (double)R;
    agg[1] = sum[0];
    agg[2] = sum[1];
    agg[3] = sum[2];
    agg[4] = sum[3];
    agg[5] = sum[4];
    agg[6] = sum[5];
    for (int i = 0; i < 4; ++i) {
        agg[8 + i] = sum[6 + i];
        agg[12 + i] = sq[6 + i];
        agg[16 + i] = sum[10 + i];
        agg[20 + i] = sq[10 + i];
    }
    for (int i = 0; i < 3; ++i) {
        agg[24 + i] = sum[14 + i];
        agg[27 + i] = sq[14 + i];
    }
    for (int b = 0; b < pb; ++b) {
        agg[32 + b] = sum[17 + b];
        agg[32 + pb + b] = sq[17 + b];
    }
    return finish_timing(ctx);
}

extern "C" {

int mq_sim_vacation(mq_ctx *ctx, const mq_vac_cfg *cfg, double *agg, double *rep_out)
{
    return vacation_common(ctx, cfg, false, nullptr, nullptr, agg, rep_out, "mq_sim_vacation");
}

int mq_replay_vacation(mq_ctx *ctx, const mq_vac_cfg *cfg, const double *const *X, const int64_t *n, double *agg,
                       double *rep_out)
{
    return vacation_common(ctx, cfg, true, X, n, agg, rep_out, "mq_replay_vacation");
}

}  // extern "C"
