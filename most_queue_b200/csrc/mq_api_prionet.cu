// mq_api_prionet.cu -- C ABI entry points of the priority-network kernel.
#include "k7_prionet.cuh"
#include "mq_host.cuh"

#include <cstdlib>
#include <vector>

using namespace mq;

static int prionet_common(mq_ctx *ctx, const mq_prionet_cfg *cfg, bool replay, const double *const *A,
                          const int64_t *n_a, const double *U, int64_t n_u, const double *const *S,
                          const int64_t *n_s, double *agg, double *rep_out, const char *who)
{
    if (!ctx || !cfg || !agg) return fail(MQ_E_INVALID, "%s: NULL argument", who);
    const int m = cfg->m, K = cfg->K;
    if (m < 1 || m > K7_MAXN) return fail(m < 1 ? MQ_E_INVALID : MQ_E_UNSUPPORTED, "%s: nodes must be in 1..%d", who, K7_MAXN);
    if (K < 1 || K > K7_MAXK) return fail(K < 1 ? MQ_E_INVALID : MQ_E_UNSUPPORTED, "%s: classes must be in 1..%d", who, K7_MAXK);
    if (cfg->job_served < 0 || cfg->replications < 1 || cfg->rep0 < 0 || cfg->rep0 + cfg->replications > (1ll << 32))
        return fail(MQ_E_INVALID, "%s: bad sizes", who);

    K7Params P;
    std::memset(&P, 0, sizeof(P));
    P.m = m;
    P.K = K;
    int total = 0;
    for (int i = 0; i < m; ++i) {
        if (cfg->channels[i] < 1) return fail(MQ_E_INVALID, "%s: node %d needs >= 1 channel", who, i);
        if (cfg->prty[i] < MQ_PRTY_NP || cfg->prty[i] > MQ_PRTY_RW)
            return fail(MQ_E_INVALID, "%s: node %d: prty must be NP, PR, RS or RW", who, i);
        P.channels[i] = cfg->channels[i];
        P.prty[i] = cfg->prty[i];
        P.base[i] = total;
        total += cfg->channels[i];
        for (int k = 0; k < K; ++k) {
            const int nc = cfg->nodes_prty[i * K + k];
            if (nc < 0 || nc >= K) return fail(MQ_E_INVALID, "%s: nodes_prty[%d][%d] out of range", who, i, k);
            P.nodes_prty[i * K + k] = nc;
        }
    }
    P.base[m] = total;
    if (total > K7_MAXS) return fail(MQ_E_UNSUPPORTED, "%s: %d channels in total > %d", who, total, K7_MAXS);
    P.nserv = total;
    const int RS = (m + 1) * (m + 1);
    for (int k = 0; k < K; ++k) {
        for (int q = 0; q < RS; ++q) {
            const double x = cfg->R[k * RS + q];
            if (!(x >= 0.0 && x <= 1.0 + 1e-12)) return fail(MQ_E_INVALID, "%s: R[%d] entry %d is not a probability", who, k, q);
            P.R[k * RS + q] = x;
        }
        if (cfg->R[k * RS + m] > 0.0) return fail(MQ_E_INVALID, "%s: R[%d][0][m] > 0 routes the source straight out", who, k);
    }
    if (!replay) {
        if (cfg->job_served >= (1ll << 31)) return fail(MQ_E_INVALID, "%s: job_served must be < 2^31", who);
        for (int k = 0; k < K; ++k)
            if (int rc = check_dist(cfg->src[k], "set_sources")) return rc;
        for (int i = 0; i < m * K; ++i)
            if (int rc = check_dist(cfg->srv[i], "set_nodes")) return rc;
    } else if (!A || !n_a || !U || n_u < 1 || !S || !n_s) {
        return fail(MQ_E_INVALID, "%s: NULL or empty arrays", who);
    }
    CU(use_device(ctx));

    const long long R = cfg->replications;
    const int rows = MQ_PN_REC_ROWS(m, K);
    const size_t rec_bytes = sizeof(double) * (size_t)rows * (size_t)R;
    const int nq = MQ_PN_AGG_LEN(K);
    if (int rc = ensure(ctx->rec, rec_bytes)) return rc;
    if (int rc = ensure_host_agg(ctx, sizeof(double) * 2 * (size_t)nq)) return rc;
    P.replay = replay ? 1 : 0;
    P.qcap = cfg->queue_cap > 0 ? cfg->queue_cap : 128;
    P.jobs = cfg->job_served;
    P.reps = R;
    P.rep0 = (uint32_t)cfg->rep0;
    P.key0 = seed_to_key(cfg->seed);
    P.rec = (double *)ctx->rec.p;

    ctx->launches = 0;
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (replay) {
        size_t ta = 0, ts = 0;
        for (int k = 0; k < K; ++k) {
            if (n_a[k] < 1 || !A[k]) return fail(MQ_E_INVALID, "%s: class %d gap array missing", who, k);
            P.aoff[k] = (long long)ta;
            P.n_a[k] = n_a[k];
            ta += (size_t)n_a[k] * R;
        }
        for (int i = 0; i < m * K; ++i) {
            if (n_s[i] < 0 || (n_s[i] > 0 && !S[i])) return fail(MQ_E_INVALID, "%s: service array %d missing", who, i);
            P.soff[i] = (long long)ts;
            P.n_s[i] = n_s[i];
            ts += (size_t)n_s[i] * R;
        }
        const size_t nu = (size_t)n_u * R;
        if (int rc = ensure(ctx->in_a, sizeof(double) * (ta + nu))) return rc;
        if (int rc = ensure(ctx->in_s, sizeof(double) * (ts ? ts : 1))) return rc;
        for (int k = 0; k < K; ++k)
            CU(cudaMemcpyAsync((double *)ctx->in_a.p + P.aoff[k], A[k], sizeof(double) * (size_t)n_a[k] * R,
                               cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync((double *)ctx->in_a.p + ta, U, sizeof(double) * nu, cudaMemcpyHostToDevice, ctx->stream));
        for (int i = 0; i < m * K; ++i)
            if (n_s[i])
                CU(cudaMemcpyAsync((double *)ctx->in_s.p + P.soff[i], S[i], sizeof(double) * (size_t)n_s[i] * R,
                                   cudaMemcpyHostToDevice, ctx->stream));
        P.A = (const double *)ctx->in_a.p;
        P.U = (const double *)ctx->in_a.p + ta;
        P.S = (const double *)ctx->in_s.p;
        P.n_u = n_u;
    } else {
        for (int k = 0; k < K; ++k) P.src[k] = make_distf(cfg->src[k]);
        for (int i = 0; i < m * K; ++i) P.srv[i] = make_distf(cfg->srv[i]);
    }
    // small networks run the predicated, on-chip-state form (k7v2_prionet.cu); MQ_K7_V1=1 forces the general kernel.
    // Its cursors are 32 bit.
    bool v2 = k7v2_supported(m, K, P.nserv) && !std::getenv("MQ_K7_V1") && cfg->job_served < (1ll << 28);
    if (replay) {
        v2 = v2 && n_u < (1ll << 31);
        for (int k = 0; k < K; ++k) v2 = v2 && n_a[k] < (1ll << 31);
        for (int i = 0; i < m * K; ++i) v2 = v2 && n_s[i] < (1ll << 31);
    }
    int occ = 0;
    if (v2)
        CU(k7v2_occupancy(P, &occ));
    else
        CU(k7_occupancy(&occ));
    if (occ < 1) return fail(MQ_E_CUDA, "%s: kernel does not fit on an SM", who);
    if (occ > 8) occ = 8;  // bounds the waiting-line scratch
    long long grid = (long long)occ * ctx->sms;
    const long long need = (R + K7_BLOCK - 1) / K7_BLOCK;
    if (grid > need) grid = need;
    P.lanes = grid * K7_BLOCK;
    const size_t qbytes = sizeof(double) * K7_JOBF * (size_t)m * K * (size_t)P.qcap * (size_t)P.lanes;
    if (qbytes > (size_t)64 << 30) return fail(MQ_E_UNSUPPORTED, "%s: queue_cap needs %.1f GiB", who, qbytes / 1073741824.0);
    if (int rc = ensure(ctx->ring, qbytes)) return rc;
    P.queue = (double *)ctx->ring.p;

    std::vector<RatioDesc> desc((size_t)nq, RatioDesc{MQ_PN_G_TTEK, -1});
    desc[1] = {MQ_PN_G_TTEK, -1};
    desc[2] = {MQ_PN_G_FLAGS, MQ_PN_G_FLAGS};
    for (int k = 0; k < K; ++k) {
        const int rb = MQ_PN_G_ROWS + k * MQ_PNC_ROWS, ab = 4 + 16 * k;
        for (int q = 0; q < 3; ++q) {
            desc[ab + q] = {rb + MQ_PNC_VSUM + q, rb + MQ_PNC_SERVED};
            desc[ab + 3 + q] = desc[ab + q];
            desc[ab + 6 + q] = {rb + MQ_PNC_WSUM + q, rb + MQ_PNC_SERVED};
            desc[ab + 9 + q] = desc[ab + 6 + q];
        }
        desc[ab + 12] = {rb + MQ_PNC_SERVED, -1};
        desc[ab + 13] = {rb + MQ_PNC_ARRIVED, -1};
    }
    if (int rc = ensure(ctx->desc, sizeof(RatioDesc) * (size_t)nq)) return rc;
    CU(cudaMemcpyAsync(ctx->desc.p, desc.data(), sizeof(RatioDesc) * (size_t)nq, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->rec.p, 0, rec_bytes, ctx->stream));
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (v2)
        CU(k7v2_launch(P, (int)grid, ctx->stream));
    else
        CU(k7_launch(P, (int)grid, ctx->stream));
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
    for (int q = 0; q < nq; ++q) agg[q] = sum[q];
    agg[0] = (double)R;
    agg[3] = 0.0;
    for (int k = 0; k < K; ++k) {
        const int ab = 4 + 16 * k;
        for (int q = 0; q < 3; ++q) {
            agg[ab + 3 + q] = sq[ab + q];
            agg[ab + 9 + q] = sq[ab + 6 + q];
        }
        agg[ab + 14] = agg[ab + 15] = 0.0;
    }
    return finish_timing(ctx);
}

extern "C" {

int mq_sim_prionet(mq_ctx *ctx, const mq_prionet_cfg *cfg, double *agg, double *rep_out)
{
    return prionet_common(ctx, cfg, false, nullptr, nullptr, nullptr, 0, nullptr, nullptr, agg, rep_out, "mq_sim_prionet");
}

int mq_replay_prionet(mq_ctx *ctx, const mq_prionet_cfg *cfg, const double *const *A, const int64_t *n_a,
                      const double *U, int64_t n_u, const double *const *S, const int64_t *n_s, double *agg,
                      double *rep_out)
{
    return prionet_common(ctx, cfg, true, A, n_a, U, n_u, S, n_s, agg, rep_out, "mq_replay_prionet");
}

}  // extern "C"
