// mq_api_size.cu -- C ABI entry points of the size-based single-server kernel.
#include "k10_size.cuh"
#include "mq_host.cuh"

#include <vector>

using namespace mq;

static int size_common(mq_ctx *ctx, const mq_size_cfg *cfg, bool replay, const double *A, int64_t n_a, const double *S,
                       int64_t n_s, const double *Y, int64_t n_y, double *agg, double *rep_out, const char *who)
{
    if (!ctx || !cfg || !agg) return fail(MQ_E_INVALID, "%s: NULL argument", who);
    if (cfg->discipline < MQ_SZ_FCFS || cfg->discipline > MQ_SZ_SPRPT)
        return fail(MQ_E_INVALID, "%s: unknown discipline %d", who, cfg->discipline);
    if (cfg->p_bins < 1 || cfg->p_bins > MQ_MAX_PBINS) return fail(MQ_E_INVALID, "%s: p_bins must be in 1..%d", who, MQ_MAX_PBINS);
    if (cfg->total_served < 0 || cfg->replications < 1 || cfg->rep0 < 0 || cfg->rep0 + cfg->replications > (1ll << 32))
        return fail(MQ_E_INVALID, "%s: bad sizes", who);
    if (!replay) {
        if (cfg->total_served >= (1ll << 31)) return fail(MQ_E_INVALID, "%s: total_served must be < 2^31", who);
        if (int rc = check_dist(cfg->src, "set_sources")) return rc;
        if (int rc = check_dist(cfg->srv, "set_servers")) return rc;
        if (cfg->pred_kind < MQ_PRED_PERFECT || cfg->pred_kind > MQ_PRED_LOGNORMAL)
            return fail(MQ_E_INVALID, "%s: unknown predictor kind %d", who, cfg->pred_kind);
        if (cfg->pred_kind == MQ_PRED_LOGNORMAL && !(cfg->pred_sigma > 0.0))
            return fail(MQ_E_INVALID, "sigma must be positive");  // predictor.py:37
    } else if (!A || !S || n_a < 1 || n_s < 1) {
        return fail(MQ_E_INVALID, "%s: NULL or empty arrays", who);
    }
    CU(use_device(ctx));

    const long long R = cfg->replications;
    const int pb = cfg->p_bins;
    const size_t rec_bytes = sizeof(double) * (size_t)MQ_Z_ROWS(pb) * (size_t)R;
    const int nq = 17 + pb;
    if (int rc = ensure(ctx->rec, rec_bytes)) return rc;
    if (int rc = ensure_host_agg(ctx, sizeof(double) * 2 * (size_t)nq)) return rc;

    K10Params P;
    std::memset(&P, 0, sizeof(P));
    P.discipline = cfg->discipline;
    P.p_bins = pb;
    P.replay = replay ? 1 : 0;
    P.buffer = cfg->buffer < 0 ? -1 : cfg->buffer;
    P.pred_kind = replay ? (Y ? MQ_PRED_GIVEN : MQ_PRED_PERFECT) : cfg->pred_kind;
    P.pred_sigma = cfg->pred_sigma;
    P.jobs = cfg->total_served;
    P.reps = R;
    P.rep0 = (uint32_t)cfg->rep0;
    P.key0 = seed_to_key(cfg->seed);
    P.rec = (double *)ctx->rec.p;
    long long qcap = cfg->queue_cap > 0 ? cfg->queue_cap : 256;
    // finite buffer: pre-empted jobs join the line even when it is full (size_based.py:199), so leave headroom
    if (cfg->buffer >= 0) qcap = cfg->buffer + (cfg->queue_cap > 0 ? cfg->queue_cap : 64);
    if (qcap > (1 << 16)) return fail(MQ_E_UNSUPPORTED, "%s: waiting line of %lld jobs per replication", who, qcap);
    P.qcap = (int)qcap;

    ctx->launches = 0;
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (replay) {
        const size_t na = (size_t)n_a * R, ns = (size_t)n_s * R;
        if (int rc = ensure(ctx->in_a, sizeof(double) * na)) return rc;
        const size_t ny = Y ? (size_t)n_y * R : 0;
        if (Y && n_y < 1) return fail(MQ_E_INVALID, "%s: empty prediction array", who);
        if (int rc = ensure(ctx->in_s, sizeof(double) * (ns + ny))) return rc;
        CU(cudaMemcpyAsync(ctx->in_a.p, A, sizeof(double) * na, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->in_s.p, S, sizeof(double) * ns, cudaMemcpyHostToDevice, ctx->stream));
        if (Y) {
            CU(cudaMemcpyAsync((double *)ctx->in_s.p + ns, Y, sizeof(double) * ny, cudaMemcpyHostToDevice, ctx->stream));
            P.Y = (const double *)ctx->in_s.p + ns;
            P.n_y = n_y;
        }
        P.A = (const double *)ctx->in_a.p;
        P.S = (const double *)ctx->in_s.p;
        P.n_a = n_a;
        P.n_s = n_s;
    } else {
        P.src = make_distf(cfg->src);
        P.srv = make_distf(cfg->srv);
        mq_dist noise;
        std::memset(&noise, 0, sizeof(noise));
        if (cfg->pred_kind == MQ_PRED_EXP) {
            noise.kind = MQ_M;
            noise.p[0] = 1.0;
            P.pred_law = make_distf(noise);
        } else if (cfg->pred_kind == MQ_PRED_LOGNORMAL) {
            noise.kind = MQ_NORM;
            noise.p[0] = 0.0;
            noise.p[1] = 1.0;
            P.pred_law = make_distf(noise);
        }
    }
    int occ = 0;
    CU(k10_occupancy(&occ));
    if (occ < 1) return fail(MQ_E_CUDA, "%s: kernel does not fit on an SM", who);
    long long grid = (long long)occ * ctx->sms;
    const long long want = (R + K10_BLOCK - 1) / K10_BLOCK;
    if (grid > want) grid = want;
    P.lanes = grid * K10_BLOCK;
    if (int rc = ensure(ctx->ring, sizeof(double) * K10_JOBF * (size_t)P.qcap * (size_t)P.lanes)) return rc;
    P.queue = (double *)ctx->ring.p;

    std::vector<RatioDesc> desc((size_t)nq);
    desc[0] = {MQ_Z_TTEK, -1};
    desc[1] = {MQ_Z_FLAGS, MQ_Z_FLAGS};
    desc[2] = {MQ_Z_SERVED, -1};
    desc[3] = {MQ_Z_ARRIVED, -1};
    desc[4] = {MQ_Z_DROPPED, -1};
    desc[5] = {MQ_Z_TAKED, -1};
    desc[6] = {MQ_Z_PREEMPTIONS, -1};
    desc[7] = {MQ_Z_SLOWDOWN, MQ_Z_SERVED};
    for (int i = 0; i < 4; ++i) {
        desc[8 + i] = {MQ_Z_WSUM + i, MQ_Z_TAKED};
        desc[12 + i] = {MQ_Z_VSUM + i, MQ_Z_SERVED};
    }
    desc[16] = {MQ_Z_TTEK, -1};
    for (int b = 0; b < pb; ++b) desc[17 + b] = {MQ_Z_P + b, MQ_Z_TTEK};
    if (int rc = ensure(ctx->desc, sizeof(RatioDesc) * (size_t)nq)) return rc;
    CU(cudaMemcpyAsync(ctx->desc.p, desc.data(), sizeof(RatioDesc) * (size_t)nq, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->rec.p, 0, rec_bytes, ctx->stream));
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    CU(k10_launch(P, (int)grid, ctx->stream));
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
    for (int q = 0; q < MQ_Z_AGG_LEN(pb); ++q) agg[q] = 0.0;
    agg[0] = (double)R;
    agg[1] = sum[0];
    agg[2] = sum[1];
    for (int i = 0; i < 5; ++i) agg[3 + i] = sum[2 + i];
    for (int i = 0; i < 4; ++i) {
        agg[8 + i] = sum[8 + i];
        agg[12 + i] = sq[8 + i];
        agg[16 + i] = sum[12 + i];
        agg[20 + i] = sq[12 + i];
    }
    agg[24] = sum[7];
    agg[25] = sq[7];
    for (int b = 0; b < pb; ++b) {
        agg[32 + b] = sum[17 + b];
        agg[32 + pb + b] = sq[17 + b];
    }
    return finish_timing(ctx);
}

extern "C" {

int mq_sim_size(mq_ctx *ctx, const mq_size_cfg *cfg, double *agg, double *rep_out)
{
    return size_common(ctx, cfg, false, nullptr, 0, nullptr, 0, nullptr, 0, agg, rep_out, "mq_sim_size");
}

int mq_replay_size(mq_ctx *ctx, const mq_size_cfg *cfg, const double *A, int64_t n_a, const double *S, int64_t n_s,
                   const double *Y, int64_t n_y, double *agg, double *rep_out)
{
    return size_common(ctx, cfg, true, A, n_a, S, n_s, Y, n_y, agg, rep_out, "mq_replay_size");
}

}  // extern "C"
