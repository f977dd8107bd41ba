// mq_api_prio.cu -- C ABI entry points of the priority kernel (mq_sim_prio, mq_replay_prio).
#include "k4_prio.cuh"
#include "mq_host.cuh"

#include <cstdlib>
#include <vector>

using namespace mq;

static_assert(K4_ROWS(0) == MQ_PRIO_ROWS(0) && K4_G_ROWS == MQ_PG_ROWS, "record layout out of sync");
static_assert(K4_MAXK == MQ_MAX_CLASSES && K4_MAXC == MQ_PRIO_MAX_CHANNELS, "limits out of sync");

static int prio_common(mq_ctx *ctx, const mq_prio_cfg *cfg, bool replay, const double *const *A,
                       const int64_t *n_a, const double *const *S, const int64_t *n_s, double *agg,
                       double *rep_out, const char *who)
{
    if (!ctx || !cfg || !agg) return fail(MQ_E_INVALID, "%s: NULL argument", who);
    if (cfg->c < 1 || cfg->c > K4_MAXC)
        return fail(cfg->c < 1 ? MQ_E_INVALID : MQ_E_UNSUPPORTED, "%s: num_of_channels must be in 1..%d", who, K4_MAXC);
    if (cfg->K < 1 || cfg->K > K4_MAXK)
        return fail(cfg->K < 1 ? MQ_E_INVALID : MQ_E_UNSUPPORTED, "%s: num_of_classes must be in 1..%d", who, K4_MAXK);
    if (cfg->prty_type < MQ_PRTY_NP || cfg->prty_type > MQ_PRTY_RW)
        return fail(MQ_E_INVALID, "%s: prty_type must be NP, PR, RS or RW", who);
    if (cfg->p_bins < 1 || cfg->p_bins > MQ_MAX_PBINS) return fail(MQ_E_INVALID, "%s: p_bins out of range", who);
    if (cfg->total_served < 0 || cfg->replications < 1 || cfg->rep0 < 0 ||
        cfg->rep0 + cfg->replications > (1ll << 32))
        return fail(MQ_E_INVALID, "%s: bad sizes", who);
    const int K = cfg->K;
    if (!replay) {
        if (cfg->total_served >= (1ll << 31)) return fail(MQ_E_INVALID, "%s: total_served must be < 2^31", who);
        for (int k = 0; k < K; ++k) {
            if (int rc = check_dist(cfg->src[k], "set_sources")) return rc;
            if (int rc = check_dist(cfg->srv[k], "set_servers")) return rc;
        }
    } else if (!A || !S || !n_a || !n_s) {
        return fail(MQ_E_INVALID, "%s: NULL arrays", who);
    }
    CU(use_device(ctx));

    const long long R = cfg->replications;
    const int rows = MQ_PRIO_REC_ROWS(K, cfg->p_bins);
    const size_t rec_bytes = sizeof(double) * (size_t)rows * (size_t)R;
    const int nq = MQ_PRIO_AGG_LEN(K, cfg->p_bins);
    if (int rc = ensure(ctx->rec, rec_bytes)) return rc;
    if (int rc = ensure_host_agg(ctx, sizeof(double) * 2 * (size_t)nq)) return rc;

    K4Params P;
    std::memset(&P, 0, sizeof(P));
    P.c = cfg->c;
    P.K = K;
    P.prty = cfg->prty_type;
    P.p_bins = cfg->p_bins;
    P.replay = replay ? 1 : 0;
    P.qcap = cfg->queue_cap > 0 ? cfg->queue_cap : 512;
    P.buffer = cfg->buffer;
    P.jobs = cfg->total_served;
    P.reps = R;
    P.rep0 = (uint32_t)cfg->rep0;
    P.key0 = seed_to_key(cfg->seed);
    P.rec = (double *)ctx->rec.p;

    ctx->launches = 0;
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (replay) {
        size_t ta = 0, ts = 0;
        for (int k = 0; k < K; ++k) {
            if (n_a[k] < 1 || n_s[k] < 0 || !A[k] || (n_s[k] > 0 && !S[k]))
                return fail(MQ_E_INVALID, "%s: class %d arrays missing", who, k);
            P.aoff[k] = (long long)ta;
            P.soff[k] = (long long)ts;
            P.n_a[k] = n_a[k];
            P.n_s[k] = n_s[k];
            ta += (size_t)n_a[k] * (size_t)R;
            ts += (size_t)n_s[k] * (size_t)R;
        }
        if (int rc = ensure(ctx->in_a, sizeof(double) * ta)) return rc;
        if (int rc = ensure(ctx->in_s, sizeof(double) * (ts ? ts : 1))) return rc;
        for (int k = 0; k < K; ++k) {
            CU(cudaMemcpyAsync((double *)ctx->in_a.p + P.aoff[k], A[k], sizeof(double) * (size_t)n_a[k] * R,
                               cudaMemcpyHostToDevice, ctx->stream));
            if (n_s[k])
                CU(cudaMemcpyAsync((double *)ctx->in_s.p + P.soff[k], S[k], sizeof(double) * (size_t)n_s[k] * R,
                                   cudaMemcpyHostToDevice, ctx->stream));
        }
        P.A = (const double *)ctx->in_a.p;
        P.S = (const double *)ctx->in_s.p;
    } else {
        for (int k = 0; k < K; ++k) {
            P.src[k] = make_distf(cfg->src[k]);
            P.srv[k] = make_distf(cfg->srv[k]);
        }
    }

    // c <= 8 and K <= 4 run the predicated, on-chip-state form (k4v2_prio.cu); MQ_K4_V1=1 forces the general kernel
    bool v2 = k4v2_supported(cfg->c, K) && !std::getenv("MQ_K4_V1") && cfg->total_served < (1ll << 30);
    if (replay)  // its counters and cursors are 32 bit
        for (int k = 0; k < K; ++k) v2 = v2 && n_a[k] < (1ll << 31) && n_s[k] < (1ll << 31);
    int occ = 0;
    if (v2)
        CU(k4v2_occupancy(cfg->c, K, replay, &occ));
    else
        CU(k4_occupancy(&occ));
    if (occ < 1) return fail(MQ_E_CUDA, "%s: kernel does not fit on an SM", who);
    long long grid = (long long)occ * ctx->sms;
    const long long need = (R + K4_BLOCK - 1) / K4_BLOCK;
    if (grid > need) grid = need;
    P.lanes = grid * K4_BLOCK;
    const size_t qbytes = sizeof(double) * 4 * (size_t)K * (size_t)P.qcap * (size_t)P.lanes;
    if (qbytes > (size_t)64 << 30) return fail(MQ_E_UNSUPPORTED, "%s: queue_cap needs %.1f GiB", who, qbytes / 1073741824.0);
    if (int rc = ensure(ctx->ring, qbytes)) return rc;
    P.queue = (double *)ctx->ring.p;

    // reduction table
    std::vector<RatioDesc> desc((size_t)nq, RatioDesc{MQ_PG_FLAGS, -1});
    const int crow = MQ_PRIO_ROWS(cfg->p_bins), cagg = MQ_PRIO_AGG(cfg->p_bins);
    desc[0] = {MQ_PG_FLAGS, MQ_PG_FLAGS};  // x/x = 1 for flagged reps only; entry 0 is overwritten on the host
    desc[1] = {MQ_PG_TTEK, -1};
    desc[2] = {MQ_PG_FLAGS, MQ_PG_FLAGS};
    desc[3] = {MQ_PG_TTEK, MQ_PG_TTEK};
    for (int k = 0; k < K; ++k) {
        const int rb = MQ_PG_ROWS + k * crow, ab = 4 + k * cagg;
        for (int i = 0; i < 4; ++i) {
            desc[ab + i] = {rb + MQ_PF_WSUM + i, rb + MQ_PF_SERVED};
            desc[ab + 4 + i] = desc[ab + i];  // squares come from out[nq + q]; filled on the host
            desc[ab + 8 + i] = {rb + MQ_PF_VSUM + i, rb + MQ_PF_SERVED};
            desc[ab + 12 + i] = desc[ab + 8 + i];
        }
        desc[ab + 16] = {rb + MQ_PF_ARRIVED, -1};
        desc[ab + 17] = {rb + MQ_PF_TAKED, -1};
        desc[ab + 18] = {rb + MQ_PF_SERVED, -1};
        desc[ab + 19] = {rb + MQ_PF_DROPPED, -1};
        for (int i = 20; i < 24; ++i) desc[ab + i] = {MQ_PG_TTEK, -1};
        for (int j = 0; j < cfg->p_bins; ++j) {
            desc[ab + 24 + j] = {rb + MQ_PF_P + j, MQ_PG_TTEK};
            desc[ab + 24 + cfg->p_bins + j] = desc[ab + 24 + j];
        }
    }
    if (int rc = ensure(ctx->desc, sizeof(RatioDesc) * (size_t)nq)) return rc;
    CU(cudaMemcpyAsync(ctx->desc.p, desc.data(), sizeof(RatioDesc) * (size_t)nq, cudaMemcpyHostToDevice, ctx->stream));

    CU(cudaMemsetAsync(ctx->rec.p, 0, rec_bytes, ctx->stream));
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (v2)
        CU(k4v2_launch(P, (int)grid, ctx->stream));
    else
        CU(k4_launch(P, (int)grid, ctx->stream));
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

    // sums from out[q], squares from out[nq + q]
    const double *sum = ctx->h_agg, *sq = ctx->h_agg + nq;
    for (int q = 0; q < nq; ++q) agg[q] = sum[q];
    agg[0] = (double)R;
    agg[3] = 0.0;
    for (int k = 0; k < K; ++k) {
        const int ab = 4 + k * cagg;
        for (int i = 0; i < 4; ++i) {
            agg[ab + 4 + i] = sq[ab + i];
            agg[ab + 12 + i] = sq[ab + 8 + i];
        }
        for (int i = 20; i < 24; ++i) agg[ab + i] = 0.0;
        for (int j = 0; j < cfg->p_bins; ++j) agg[ab + 24 + cfg->p_bins + j] = sq[ab + 24 + j];
    }
    return finish_timing(ctx);
}

extern "C" {

int mq_sim_prio(mq_ctx *ctx, const mq_prio_cfg *cfg, double *agg, double *rep_out)
{
    return prio_common(ctx, cfg, false, nullptr, nullptr, nullptr, nullptr, agg, rep_out, "mq_sim_prio");
}

int mq_replay_prio(mq_ctx *ctx, const mq_prio_cfg *cfg, const double *const *A, const int64_t *n_a,
                   const double *const *S, const int64_t *n_s, double *agg, double *rep_out)
{
    return prio_common(ctx, cfg, true, A, n_a, S, n_s, agg, rep_out, "mq_replay_prio");
}

}  // extern "C"
