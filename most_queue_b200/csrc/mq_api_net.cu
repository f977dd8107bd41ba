// mq_api_net.cu -- C ABI entry points of the open-network kernel (mq_sim_net, mq_replay_net).
#include "k5_net.cuh"
#include "mq_host.cuh"

#include <cstdlib>
#include <vector>

using namespace mq;

static int net_common(mq_ctx *ctx, const mq_net_cfg *cfg, bool replay, const double *A, int64_t n_a,
                      const double *U, int64_t n_u, const double *const *S, const int64_t *n_s, double *agg,
                      double *rep_out, const char *who)
{
    if (!ctx || !cfg || !agg) return fail(MQ_E_INVALID, "%s: NULL argument", who);
    const int m = cfg->m;
    if (m < 1 || m > K5_MAXN)
        return fail(m < 1 ? MQ_E_INVALID : MQ_E_UNSUPPORTED, "%s: number of nodes must be in 1..%d", who, K5_MAXN);
    if (cfg->job_served < 0 || cfg->replications < 1 || cfg->rep0 < 0 || cfg->rep0 + cfg->replications > (1ll << 32))
        return fail(MQ_E_INVALID, "%s: bad sizes", who);

    K5Params P;
    std::memset(&P, 0, sizeof(P));
    P.m = m;
    int total = 0;
    for (int i = 0; i < m; ++i) {
        if (cfg->channels[i] < 1) return fail(MQ_E_INVALID, "%s: node %d needs >= 1 channel", who, i);
        P.channels[i] = cfg->channels[i];
        P.base[i] = total;
        total += cfg->channels[i];
    }
    P.base[m] = total;
    if (total > K5_MAXS) return fail(MQ_E_UNSUPPORTED, "%s: %d channels in total > %d", who, total, K5_MAXS);
    P.nserv = total;
    for (int r = 0; r < m + 1; ++r) {
        for (int q = 0; q < m + 1; ++q) {
            const double x = cfg->R[r * (m + 1) + q];
            if (!(x >= 0.0 && x <= 1.0 + 1e-12)) return fail(MQ_E_INVALID, "%s: R[%d][%d] is not a probability", who, r, q);
            P.R[r * (m + 1) + q] = x;
        }
    }
    if (cfg->R[m] > 0.0) return fail(MQ_E_INVALID, "%s: R[0][m] > 0 routes the source straight out of the network", who);
    if (!replay) {
        if (cfg->job_served >= (1ll << 31)) return fail(MQ_E_INVALID, "%s: job_served must be < 2^31", who);
        if (int rc = check_dist(cfg->src, "set_sources")) return rc;
        for (int i = 0; i < m; ++i)
            if (int rc = check_dist(cfg->srv[i], "set_nodes")) return rc;
    } else if (!A || !U || !S || !n_s || n_a < 1 || n_u < 1) {
        return fail(MQ_E_INVALID, "%s: NULL or empty arrays", who);
    }
    CU(use_device(ctx));

    const long long R = cfg->replications;
    const int rows = MQ_NET_REC_ROWS(m);
    const size_t rec_bytes = sizeof(double) * (size_t)rows * (size_t)R;
    const int nq = MQ_NET_AGG_LEN(m);
    if (int rc = ensure(ctx->rec, rec_bytes)) return rc;
    if (int rc = ensure_host_agg(ctx, sizeof(double) * 2 * (size_t)nq)) return rc;

    P.replay = replay ? 1 : 0;
    P.qcap = cfg->queue_cap > 0 ? cfg->queue_cap : 512;
    P.jobs = cfg->job_served;
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
        const size_t na = (size_t)n_a * R, nu = (size_t)n_u * R;
        if (int rc = ensure(ctx->in_a, sizeof(double) * (na + nu))) return rc;
        if (int rc = ensure(ctx->in_s, sizeof(double) * (ts ? ts : 1))) return rc;
        CU(cudaMemcpyAsync(ctx->in_a.p, A, sizeof(double) * na, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync((double *)ctx->in_a.p + na, U, sizeof(double) * nu, cudaMemcpyHostToDevice, ctx->stream));
        for (int i = 0; i < m; ++i)
            if (n_s[i])
                CU(cudaMemcpyAsync((double *)ctx->in_s.p + P.soff[i], S[i], sizeof(double) * (size_t)n_s[i] * R,
                                   cudaMemcpyHostToDevice, ctx->stream));
        P.A = (const double *)ctx->in_a.p;
        P.U = (const double *)ctx->in_a.p + na;
        P.S = (const double *)ctx->in_s.p;
        P.n_a = n_a;
        P.n_u = n_u;
    } else {
        P.src = make_distf(cfg->src);
        for (int i = 0; i < m; ++i) P.srv[i] = make_distf(cfg->srv[i]);
    }

    // small networks run the predicated, on-chip-state form (k5v2_net.cu); MQ_K5_V1=1 forces the general kernel.
    // Its cursors are 32 bit.
    bool v2 = k5v2_supported(m, P.nserv) && !std::getenv("MQ_K5_V1") && cfg->job_served < (1ll << 28);
    if (replay) {
        v2 = v2 && n_a < (1ll << 31) && n_u < (1ll << 31);
        for (int i = 0; i < m; ++i) v2 = v2 && n_s[i] < (1ll << 31);
    }
    int occ = 0;
    if (v2)
        CU(k5v2_occupancy(P, &occ));
    else
        CU(k5_occupancy(&occ));
    if (occ < 1) return fail(MQ_E_CUDA, "%s: kernel does not fit on an SM", who);
    if (const char *cap = std::getenv("MQ_K5_MAX_CTAS")) {  // measurement knob: resident CTAs per SM
        const int v = std::atoi(cap);
        if (v >= 1 && v < occ) occ = v;
    }
    long long grid = (long long)occ * ctx->sms;
    const long long need = (R + K5_BLOCK - 1) / K5_BLOCK;
    if (grid > need) grid = need;
    P.lanes = grid * K5_BLOCK;
    const size_t qbytes = sizeof(double) * 3 * (size_t)m * (size_t)P.qcap * (size_t)P.lanes;
    if (qbytes > (size_t)64 << 30) return fail(MQ_E_UNSUPPORTED, "%s: queue_cap needs %.1f GiB", who, qbytes / 1073741824.0);
    if (int rc = ensure(ctx->ring, qbytes)) return rc;
    P.queue = (double *)ctx->ring.p;

    std::vector<RatioDesc> desc((size_t)nq, RatioDesc{MQ_NG_TTEK, -1});
    desc[1] = {MQ_NG_TTEK, -1};
    desc[2] = {MQ_NG_FLAGS, MQ_NG_FLAGS};
    desc[3] = {MQ_NG_SERVED, -1};
    desc[4] = {MQ_NG_ARRIVED, -1};
    for (int k = 0; k < 3; ++k) {
        desc[5 + k] = {MQ_NG_VSUM + k, MQ_NG_SERVED};
        desc[8 + k] = desc[5 + k];
        desc[11 + k] = {MQ_NG_WSUM + k, MQ_NG_SERVED};
        desc[14 + k] = desc[11 + k];
    }
    for (int i = 0; i < m; ++i) {
        const int rb = MQ_NG_ROWS + i * MQ_NN_ROWS, ab = 20 + 16 * i;
        for (int k = 0; k < 4; ++k) {
            desc[ab + k] = {rb + MQ_NN_VSUM + k, rb + MQ_NN_SERVED};
            desc[ab + 4 + k] = {rb + MQ_NN_WSUM + k, rb + MQ_NN_TAKED};
        }
        desc[ab + 8] = {rb + MQ_NN_SERVED, -1};
        desc[ab + 9] = {rb + MQ_NN_TAKED, -1};
        desc[ab + 10] = {rb + MQ_NN_ARRIVED, -1};
        desc[ab + 11] = desc[ab + 0];
        desc[ab + 12] = desc[ab + 4];
    }
    if (int rc = ensure(ctx->desc, sizeof(RatioDesc) * (size_t)nq)) return rc;
    CU(cudaMemcpyAsync(ctx->desc.p, desc.data(), sizeof(RatioDesc) * (size_t)nq, cudaMemcpyHostToDevice, ctx->stream));

    CU(cudaMemsetAsync(ctx->rec.p, 0, rec_bytes, ctx->stream));
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (v2)
        CU(k5v2_launch(P, (int)grid, ctx->stream));
    else
        CU(k5_launch(P, (int)grid, ctx->stream));
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
    for (int k = 0; k < 3; ++k) {
        agg[8 + k] = sq[5 + k];
        agg[14 + k] = sq[11 + k];
    }
    for (int q = 17; q < 20; ++q) agg[q] = 0.0;
    for (int i = 0; i < m; ++i) {
        const int ab = 20 + 16 * i;
        agg[ab + 11] = sq[ab + 0];
        agg[ab + 12] = sq[ab + 4];
        for (int q = 13; q < 16; ++q) agg[ab + q] = 0.0;
    }
    return finish_timing(ctx);
}

extern "C" {

int mq_sim_net(mq_ctx *ctx, const mq_net_cfg *cfg, double *agg, double *rep_out)
{
    return net_common(ctx, cfg, false, nullptr, 0, nullptr, 0, nullptr, nullptr, agg, rep_out, "mq_sim_net");
}

int mq_replay_net(mq_ctx *ctx, const mq_net_cfg *cfg, const double *A, int64_t n_a, const double *U, int64_t n_u,
                  const double *const *S, const int64_t *n_s, double *agg, double *rep_out)
{
    return net_common(ctx, cfg, true, A, n_a, U, n_u, S, n_s, agg, rep_out, "mq_replay_net");
}

}  // extern "C"
