// k1b_fifo_jobs.cu -- K1b: job-indexed replay of GI/G/c FIFO with an infinite buffer (fp64, HBM-roofline tier).
//
// QsSim.run() (most_queue/sim/base.py:488-528) is event driven: ~320 instructions per bit-exact fp64 event in K1, whose
// lanes also drift apart in the arrays they read.  With FIFO and an unbounded waiting line the same sample values follow
// from the Kiefer-Wolfowitz recursion over JOBS, in arrival order, for a lane-uniform number of steps:
//     t_n = t_{n-1} + A[n]                 arrival time          (base.py:230-236: arrival_time = ttek + gap)
//     m   = min of the c completion times  the server that frees first (free servers hold a time <= t_n)
//     start = max(t_n, m)                  service starts at the arrival or at that departure (base.py:155-184, 288-310)
//     w_n = start - t_n                    wait sample           (base.py:300-303: ttek - start_waiting_time; 0 when free)
//     d   = start + S[n]                   completion time       (servers.py:88-90: time_to_end_service = ttek + service)
//     v_n = d - t_n                        sojourn sample        (base.py:269-272: ttek - arr_time)
// Each of those is one correctly rounded fp64 operation, the same ones the reference performs on the same operands, so
// every per-job w_n and v_n equals the reference's sample BIT FOR BIT (tests: the reference's own fixtures, and the
// oracle's job-indexed mode).  What differs from K1 is the ESTIMATOR, not the samples: the sums run over jobs 0..N-1 in
// arrival order (the reference stops at the N-th departure and so drops the <= c-1 jobs still in service), and no
// time-in-state histogram is kept (K1 / K2 / K3 provide it).
//
// Every lane consumes row n of A and S at step n: a warp reads 256 contiguous bytes per row and array, every 32-byte
// sector exactly once (K1's event-driven lanes re-fetched sectors: 1.6x over-read), rows are prefetched MQ_K1B_AHEAD
// steps ahead.  The c completion times stay sorted in registers (a 2(c-1)-comparator insertion per job).
#include "mq_host.cuh"

#include <cstdlib>

namespace mq {

constexpr int MQ_K1B_BLOCK = 128;
constexpr int MQ_K1B_AHEAD = 8;

struct K1bParams {
    int c;
    long long jobs;     // N: jobs 0..N-1 of every replication
    long long reps;
    const double *A;    // [jobs][reps]  A[n] = gap before job n
    const double *S;    // [jobs][reps]
    double *rec;        // [MQ_JB_ROWS][reps]
    double *W;          // optional [jobs][reps] per-job waits
    double *V;          // optional [jobs][reps] per-job sojourn times
};

// min / max of POSITIVE doubles.  INT = false: one fp64 compare and two 32-bit selects (fmin / fmax add NaN handling:
// four instructions); INT = true: on the bit patterns with 64-bit integer compares, which keeps the half-rate fp64 pipe
// free but costs more instructions.  Both are exact; MQ_K1B_INT=1 selects the second at run time (A/B measurement).
template <bool INT>
__device__ __forceinline__ double dmin_pos(double a, double b)
{
    if (INT) return __longlong_as_double(min(__double_as_longlong(a), __double_as_longlong(b)));
    return a < b ? a : b;
}
template <bool INT>
__device__ __forceinline__ double dmax_pos(double a, double b)
{
    if (INT) return __longlong_as_double(max(__double_as_longlong(a), __double_as_longlong(b)));
    return a < b ? b : a;
}

// L[0..CT-1] ascending; new L = sorted(L[1..CT-1] + {y}): the minimum (the server that takes the job) is replaced
template <int CT, bool INT>
__device__ __forceinline__ void kw_replace_min(double (&L)[CT], double y)
{
    if (CT == 1) {
        L[0] = y;
        return;
    }
    double prev = L[1];
    L[0] = dmin_pos<INT>(prev, y);
#pragma unroll
    for (int j = 1; j < CT - 1; ++j) {
        const double cur = L[j + 1];
        L[j] = dmin_pos<INT>(cur, dmax_pos<INT>(prev, y));
        prev = cur;
    }
    L[CT - 1] = dmax_pos<INT>(prev, y);
}

template <int CT, bool SAMPLES, bool INT>
__global__ void __launch_bounds__(MQ_K1B_BLOCK) k1b_fifo_jobs(const K1bParams P)
{
    const long long rep = (long long)blockIdx.x * MQ_K1B_BLOCK + threadIdx.x;
    if (rep >= P.reps) return;
    const long long R = P.reps, N = P.jobs;
    const double *A = P.A + rep, *S = P.S + rep;
    double L[CT];
#pragma unroll
    for (int j = 0; j < CT; ++j) L[j] = j < P.c ? 0.0 : __longlong_as_double(0x7ff0000000000000ll);  // free since 0 / unused
    double t = 0.0, d_max = 0.0;
    double sw[4] = {0.0, 0.0, 0.0, 0.0}, sv[4] = {0.0, 0.0, 0.0, 0.0};
    unsigned long long waited = 0;
    // rows n .. n + AHEAD - 1 are in registers
    double a[MQ_K1B_AHEAD], s[MQ_K1B_AHEAD];
#pragma unroll
    for (int u = 0; u < MQ_K1B_AHEAD; ++u) {
        a[u] = u < N ? __ldcs(A + (long long)u * R) : 0.0;
        s[u] = u < N ? __ldcs(S + (long long)u * R) : 0.0;
    }
    for (long long n0 = 0; n0 < N; n0 += MQ_K1B_AHEAD) {
#pragma unroll
        for (int u = 0; u < MQ_K1B_AHEAD; ++u) {
            const long long n = n0 + u;
            const double gap = a[u], srv = s[u];
            // refill this slot with row n + AHEAD (uniform branch: N is the same for every lane)
            const long long nn = n + MQ_K1B_AHEAD;
            if (nn < N) {
                a[u] = __ldcs(A + nn * R);
                s[u] = __ldcs(S + nn * R);
            }
            if (n < N) {
                t = __dadd_rn(t, gap);
                const double start = dmax_pos<INT>(t, L[0]);
                const double w = __dsub_rn(start, t);
                const double d = __dadd_rn(start, srv);
                const double v = __dsub_rn(d, t);
                kw_replace_min<CT, INT>(L, d);
                d_max = dmax_pos<INT>(d_max, d);
                waited += w > 0.0 ? 1ull : 0ull;
                if (SAMPLES) {
                    if (P.W) __stcs(P.W + n * R + rep, w);
                    if (P.V) __stcs(P.V + n * R + rep, v);
                }
                // power sums in arrival order: x, x*x, fma(x*x, x, .), fma(x*x, x*x, .) -- every step correctly rounded,
                // restated with fma() by oracle/mq_oracle.c (mqo_fifo_jobs)
                const double w2 = __dmul_rn(w, w), v2 = __dmul_rn(v, v);
                sw[0] = __dadd_rn(sw[0], w);
                sw[1] = __dadd_rn(sw[1], w2);
                sw[2] = __fma_rn(w2, w, sw[2]);
                sw[3] = __fma_rn(w2, w2, sw[3]);
                sv[0] = __dadd_rn(sv[0], v);
                sv[1] = __dadd_rn(sv[1], v2);
                sv[2] = __fma_rn(v2, v, sv[2]);
                sv[3] = __fma_rn(v2, v2, sv[3]);
            }
        }
    }
    double *rec = P.rec + rep;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        rec[(long long)(MQ_JB_W + k) * R] = sw[k];
        rec[(long long)(MQ_JB_V + k) * R] = sv[k];
    }
    rec[(long long)MQ_JB_JOBS * R] = (double)N;
    rec[(long long)MQ_JB_TLAST * R] = t;
    rec[(long long)MQ_JB_DMAX * R] = d_max;
    rec[(long long)MQ_JB_WAITED * R] = (double)waited;
}

template <int CT>
static cudaError_t k1b_go(const K1bParams &P, cudaStream_t st)
{
    const int grid = (int)((P.reps + MQ_K1B_BLOCK - 1) / MQ_K1B_BLOCK);
    if (P.W || P.V)
        k1b_fifo_jobs<CT, true, false><<<grid, MQ_K1B_BLOCK, 0, st>>>(P);
    else if (std::getenv("MQ_K1B_INT"))
        k1b_fifo_jobs<CT, false, true><<<grid, MQ_K1B_BLOCK, 0, st>>>(P);
    else
        k1b_fifo_jobs<CT, false, false><<<grid, MQ_K1B_BLOCK, 0, st>>>(P);
    return cudaGetLastError();
}

static cudaError_t k1b_dispatch(const K1bParams &P, cudaStream_t st)
{
    switch (P.c) {
    case 1: return k1b_go<1>(P, st);
    case 2: return k1b_go<2>(P, st);
    case 3: return k1b_go<3>(P, st);
    case 4: return k1b_go<4>(P, st);
    case 5: return k1b_go<5>(P, st);
    case 6: return k1b_go<6>(P, st);
    case 7: return k1b_go<7>(P, st);
    case 8: return k1b_go<8>(P, st);
    default: return P.c <= 16 ? k1b_go<16>(P, st) : k1b_go<32>(P, st);
    }
}

}  // namespace mq

using namespace mq;

extern "C" int mq_replay_fifo_jobs(mq_ctx *ctx, const mq_jobs_cfg *cfg, const double *A, const double *S, double *agg,
                                   double *rep_out, double *W, double *V)
{
    if (!ctx || !cfg || !A || !S) return fail(MQ_E_INVALID, "mq_replay_fifo_jobs: NULL argument");
    if (cfg->c < 1 || cfg->c > MQ_MAX_CHANNELS)
        return fail(cfg->c < 1 ? MQ_E_INVALID : MQ_E_UNSUPPORTED, "mq_replay_fifo_jobs: num_of_channels must be in 1..%d",
                    MQ_MAX_CHANNELS);
    if (cfg->jobs < 1 || cfg->replications < 1) return fail(MQ_E_INVALID, "mq_replay_fifo_jobs: bad sizes");
    CU(use_device(ctx));
    const long long R = cfg->replications, N = cfg->jobs;
    const size_t arr = sizeof(double) * (size_t)N * (size_t)R;
    const size_t rec_bytes = sizeof(double) * (size_t)MQ_JB_ROWS * (size_t)R;
    if (int rc = ensure(ctx->rec, rec_bytes)) return rc;
    if (int rc = ensure_host_agg(ctx, sizeof(double) * MQ_JB_AGG_LEN)) return rc;
    K1bParams P;
    std::memset(&P, 0, sizeof(P));
    P.c = cfg->c;
    P.jobs = N;
    P.reps = R;
    P.rec = (double *)ctx->rec.p;
    ctx->launches = 0;
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    double *dW = nullptr, *dV = nullptr;
    if (cfg->device_ptrs) {
        P.A = A;
        P.S = S;
        dW = W;
        dV = V;
    } else {
        if (int rc = ensure(ctx->in_a, arr)) return rc;
        if (int rc = ensure(ctx->in_s, arr)) return rc;
        CU(cudaMemcpyAsync(ctx->in_a.p, A, arr, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->in_s.p, S, arr, cudaMemcpyHostToDevice, ctx->stream));
        P.A = (const double *)ctx->in_a.p;
        P.S = (const double *)ctx->in_s.p;
        if (W || V) {
            if (int rc = ensure(ctx->ring, arr * ((W ? 1 : 0) + (V ? 1 : 0)))) return rc;
            dW = W ? (double *)ctx->ring.p : nullptr;
            dV = V ? (double *)ctx->ring.p + (W ? (size_t)N * (size_t)R : 0) : nullptr;
        }
    }
    P.W = dW;
    P.V = dV;
    // the timed "sim" span starts after the input copies
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    CU(k1b_dispatch(P, ctx->stream));
    ctx->launches++;
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    if (agg) {
        // aggregate: sum over replications of sum(x^k) / jobs and of its square, k = 1..4 for w then v, then the counters
        static const RatioDesc desc[MQ_JB_AGG_LEN / 2] = {
            {MQ_JB_W, MQ_JB_JOBS}, {MQ_JB_W + 1, MQ_JB_JOBS}, {MQ_JB_W + 2, MQ_JB_JOBS}, {MQ_JB_W + 3, MQ_JB_JOBS},
            {MQ_JB_V, MQ_JB_JOBS}, {MQ_JB_V + 1, MQ_JB_JOBS}, {MQ_JB_V + 2, MQ_JB_JOBS}, {MQ_JB_V + 3, MQ_JB_JOBS},
            {MQ_JB_JOBS, -1}, {MQ_JB_TLAST, -1}, {MQ_JB_DMAX, -1}, {MQ_JB_WAITED, -1}};
        if (int rc = ensure(ctx->desc, sizeof(desc))) return rc;
        CU(cudaMemcpyAsync(ctx->desc.p, desc, sizeof(desc), cudaMemcpyHostToDevice, ctx->stream));
        CU(launch_reduce_ratio((const double *)ctx->rec.p, R, (const RatioDesc *)ctx->desc.p, MQ_JB_AGG_LEN / 2,
                               (double *)ctx->agg.p, ctx->stream));
        ctx->launches++;
    }
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    if (agg)
        CU(cudaMemcpyAsync(ctx->h_agg, ctx->agg.p, sizeof(double) * MQ_JB_AGG_LEN, cudaMemcpyDeviceToHost, ctx->stream));
    if (rep_out)
        CU(cudaMemcpyAsync(rep_out, ctx->rec.p, rec_bytes, cfg->device_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                           ctx->stream));
    if (!cfg->device_ptrs) {
        if (W) CU(cudaMemcpyAsync(W, dW, arr, cudaMemcpyDeviceToHost, ctx->stream));
        if (V) CU(cudaMemcpyAsync(V, dV, arr, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaEventRecord(ctx->ev[3], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (agg) std::memcpy(agg, ctx->h_agg, sizeof(double) * MQ_JB_AGG_LEN);
    return finish_timing(ctx);
}
