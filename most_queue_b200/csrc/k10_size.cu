// k10_size.cu -- K10: single-server queue with size-based scheduling (FCFS, SJF, PSJF, SRPT), one lane per
// replication.
//
// Walks SizeBasedQsSim.run (most_queue/sim/size_based.py) event for event:
//   arrival     size_based.py:216-247   the size is drawn at arrival (:155-164; FCFS draws the service at start);
//                                       a busy server is pre-empted when the new job's key is strictly smaller than
//                                       the key of the job in service (:182-214), whose remaining work is refreshed
//                                       first for SRPT (:166-180); the victim keeps its remaining work and its
//                                       accumulated wait and joins the line; every (re)start counts as `taked` and
//                                       samples the job's wait so far
//   waiting line sim/utils/queue_priority_size.py  min-heap on (rank, arrival time, creation order); the rank is
//                                       evaluated when the job joins the line: arrival time (FCFS), size (SJF, PSJF),
//                                       remaining work (SRPT)
//   serving     sim/base.py:249-310     base logic incl. busy periods; slowdown = sojourn / size (:110-119)
// With the default perfect predictor SPJF / PSPJF / SPRPT are the same three ranks.  The line is an unsorted array of
// 7-double records in HBM scratch, scanned for its minimum on a departure (lines are short in a stable M/G/1 queue;
// overflow beyond queue_cap is flagged).  fp64 with explicit round-to-nearest intrinsics in the reference's order:
// replay mode is bit-identical.  Generator streams: gaps = word 0, sizes = word 1 of stream 0 (as the FIFO kernel).
#include "k10_size.cuh"

namespace mq {

struct K10Job {
    double rank, arr, start_wait, wait, remaining, size;
    long long id;
    double pred;
};

__device__ __forceinline__ void k10_sample(double *run, double *sum, int n, double x, long long count)
{
    const double inv_count = __drcp_rn((double)count);
    const double one_minus = __dsub_rn(1.0, inv_count);
    double power = x;
    for (int i = 0; i < n; ++i) {
        run[i] = __dadd_rn(__dmul_rn(run[i], one_minus), __dmul_rn(power, inv_count));
        if (sum) sum[i] = __dadd_rn(sum[i], power);
        power = __dmul_rn(power, x);
    }
}

__device__ __forceinline__ bool k10_less(double ra, double aa, long long ia, double rb, double ab, long long ib)
{
    if (ra != rb) return ra < rb;
    if (aa != ab) return aa < ab;
    return ia < ib;
}

__global__ void __launch_bounds__(K10_BLOCK, 4) k10_size(const K10Params P)
{
    const long long gtid = (long long)blockIdx.x * K10_BLOCK + threadIdx.x;
    const long long R = P.reps;
    const DistF law_src = P.src, law_srv = P.srv, law_pred = P.pred_law;
    const int disc = P.discipline;
    const bool fcfs = disc == MQ_SZ_FCFS;
    const bool preemptive = disc == MQ_SZ_PSJF || disc == MQ_SZ_SRPT || disc == MQ_SZ_PSPJF || disc == MQ_SZ_SPRPT;
    const bool tracks_remaining = disc == MQ_SZ_SRPT || disc == MQ_SZ_SPRPT;

    for (long long rl = gtid; rl < R; rl += P.lanes) {
        const uint32_t rep = P.rep0 + (uint32_t)rl;
        double *rec = P.rec + rl;
        double *Q = P.queue + gtid;
        long long a_idx = 0, s_idx = 0, y_idx = 0;
        int status = 0;

        auto next_gap = [&]() -> double {
            const long long i = a_idx++;
            if (P.replay) {
                if (i >= P.n_a) {
                    status = MQ_E_INPUT_SHORT;
                    return 0.0;
                }
                return P.A[i * R + rl];
            }
            uint2 w = philox2x32_10((uint32_t)i, rep, P.key0);
            return (double)sample<-1>(law_src, w.x, (uint32_t)i, rep, P.key0, 0);
        };
        auto next_size = [&]() -> double {
            const long long i = s_idx++;
            if (P.replay) {
                if (i >= P.n_s) {
                    status = MQ_E_INPUT_SHORT;
                    return 0.0;
                }
                return P.S[i * R + rl];
            }
            uint2 w = philox2x32_10((uint32_t)i, rep, P.key0);
            return (double)sample<-1>(law_srv, w.y, (uint32_t)i, rep, P.key0, 1);
        };
        // predictor (size_based.py:163, sim/utils/predictor.py): perfect, x * Exp(1), x * exp(sigma Z), or replayed
        auto predict = [&](double x) -> double {
            if (P.pred_kind == MQ_PRED_PERFECT) return x;
            const long long i = y_idx++;
            if (P.replay) {
                if (i >= P.n_y) {
                    status = MQ_E_INPUT_SHORT;
                    return 0.0;
                }
                return P.Y[i * R + rl];
            }
            const uint32_t key = P.key0 + MQ_KEY_STRIDE * 64u;  // stream 1
            uint2 w = philox2x32_10((uint32_t)i, rep, key);
            const double z = (double)sample<-1>(law_pred, w.x, (uint32_t)i, rep, key, 0);
            return P.pred_kind == MQ_PRED_EXP ? __dmul_rn(x, z) : __dmul_rn(x, exp(__dmul_rn(P.pred_sigma, z)));
        };
        auto qf = [&](int slot, int f) -> double * { return Q + ((long long)slot * K10_JOBF + f) * P.lanes; };

        K10Job cur = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0, 0.0};
        bool busy = false;
        double end = 1e10, ttek = 0.0, start_busy = 0.0, p_over = 0.0;
        long long arrived = 0, taked = 0, served = 0, dropped = 0, in_sys = 0, busy_n = 0, next_id = 0, preempt = 0;
        int qsize = 0;
        double w_run[4] = {0, 0, 0, 0}, v_run[4] = {0, 0, 0, 0}, w_sum[4] = {0, 0, 0, 0}, v_sum[4] = {0, 0, 0, 0};
        double busy_run[3] = {0, 0, 0}, sd1 = 0.0, sd2 = 0.0;

        auto p_add = [&](double t_new) {
            const double dt = __dsub_rn(t_new, ttek);
            if (in_sys < P.p_bins)
                atomicAdd(&rec[(long long)(MQ_Z_P + in_sys) * R], dt);  // fire-and-forget fp64 add, same rounding
            else
                p_over = __dadd_rn(p_over, dt);
        };
        auto rank_of = [&](const K10Job &j) -> double {  // _rank_value size_based.py:128-153
            if (fcfs) return j.arr;
            if (disc == MQ_SZ_SJF || disc == MQ_SZ_PSJF) return j.size;
            if (disc == MQ_SZ_SRPT) return j.remaining;
            if (disc == MQ_SZ_SPJF || disc == MQ_SZ_PSPJF) return j.pred;
            const double pr = __dsub_rn(j.pred, __dsub_rn(j.size, j.remaining));  // SPRPT: predicted remaining work
            return pr > 0.0 ? pr : 0.0;
        };
        auto q_push = [&](const K10Job &j) {
            if (qsize >= P.qcap) {
                status = MQ_E_OVERFLOW;
                return;
            }
            *qf(qsize, 0) = rank_of(j);
            *qf(qsize, 1) = j.arr;
            *qf(qsize, 2) = j.start_wait;
            *qf(qsize, 3) = j.wait;
            *qf(qsize, 4) = j.remaining;
            *qf(qsize, 5) = j.size;
            *qf(qsize, 6) = (double)j.id;
            *qf(qsize, 7) = j.pred;
            qsize++;
        };
        auto start = [&](const K10Job &j) {  // Server.start_service servers.py:57-88
            cur = j;
            end = __dadd_rn(ttek, fcfs ? next_size() : j.remaining);
            busy = true;
        };

        double arrival_time = next_gap();

        while (served < P.jobs && status == 0) {
            // at most one job starts service per event -- the arriving job (idle server, or pre-emption) or the best job
            // of the line after a departure -- through ONE start site (wait sample, size draw for FCFS, completion time)
            bool do_start = false;
            K10Job sj = cur;
            if (arrival_time <= (busy ? end : 1e10)) {
                // arrival() size_based.py:216-247
                arrived++;
                p_add(arrival_time);
                ttek = arrival_time;
                arrival_time = __dadd_rn(ttek, next_gap());
                in_sys++;
                K10Job nj = {0.0, ttek, 0.0, 0.0, 0.0, 0.0, next_id++, 0.0};
                if (!fcfs) {
                    nj.size = next_size();
                    nj.remaining = nj.size;
                    nj.pred = predict(nj.size);
                }
                if (busy) {
                    nj.start_wait = ttek;
                    bool handled = false;
                    if (preemptive) {
                        const double rem0 = __dsub_rn(end, ttek);
                        const double rem = rem0 > 0.0 ? rem0 : 0.0;
                        if (tracks_remaining) cur.remaining = rem;  // _update_in_service_rank_fields :166-180
                        if (k10_less(rank_of(nj), nj.arr, nj.id, rank_of(cur), cur.arr, cur.id)) {
                            K10Job old = cur;
                            old.remaining = rem;  // preempt_service(preserve_remaining=True) servers.py:90-106
                            old.start_wait = ttek;
                            q_push(old);
                            do_start = true;
                            sj = nj;
                            preempt++;
                            handled = true;
                        }
                    }
                    if (!handled) {
                        if (P.buffer < 0 || qsize < P.buffer) {
                            q_push(nj);
                        } else {
                            dropped++;
                            in_sys--;
                        }
                    }
                } else {
                    do_start = true;
                    sj = nj;  // nj.wait == 0.0
                    if (in_sys == 1) start_busy = ttek;
                }
            } else {
                // serving() base.py:249-286
                p_add(end);
                ttek = end;
                end = 1e10;
                busy = false;
                served++;
                const double v = __dsub_rn(ttek, cur.arr);
                k10_sample(v_run, v_sum, 4, v, served);
                in_sys--;
                if (!fcfs && cur.size > 0.0) {
                    const double sd = __ddiv_rn(v, cur.size);
                    sd1 = __dadd_rn(sd1, sd);
                    sd2 = __dadd_rn(sd2, __dmul_rn(sd, sd));
                }
                if (qsize == 0 && in_sys == 0) {
                    busy_n++;
                    k10_sample(busy_run, nullptr, 3, __dsub_rn(ttek, start_busy), busy_n);
                }
                if (qsize != 0) {
                    // the best key in the line (queue_priority_size.py:70-84), then base.py:288-310
                    int best = 0;
                    double br = *qf(0, 0), ba = *qf(0, 1);
                    long long bi = (long long)*qf(0, 6);
                    for (int i = 1; i < qsize; ++i) {
                        const double r = *qf(i, 0), a = *qf(i, 1);
                        const long long id = (long long)*qf(i, 6);
                        if (k10_less(r, a, id, br, ba, bi)) {
                            best = i;
                            br = r;
                            ba = a;
                            bi = id;
                        }
                    }
                    K10Job qj = {br, ba, *qf(best, 2), *qf(best, 3), *qf(best, 4), *qf(best, 5), bi, *qf(best, 7)};
                    qsize--;
                    if (best != qsize)
                        for (int f = 0; f < K10_JOBF; ++f) *qf(best, f) = *qf(qsize, f);
                    start_busy = ttek;
                    qj.wait = __dadd_rn(qj.wait, __dsub_rn(ttek, qj.start_wait));
                    do_start = true;
                    sj = qj;
                }
            }
            if (do_start) {
                taked++;
                k10_sample(w_run, w_sum, 4, sj.wait, taked);
                start(sj);
            }
        }

        rec[(long long)MQ_Z_TTEK * R] = ttek;
        rec[(long long)MQ_Z_FLAGS * R] = (double)(status ? -status : 0);
        rec[(long long)MQ_Z_ARRIVED * R] = (double)arrived;
        rec[(long long)MQ_Z_TAKED * R] = (double)taked;
        rec[(long long)MQ_Z_SERVED * R] = (double)served;
        rec[(long long)MQ_Z_DROPPED * R] = (double)dropped;
        rec[(long long)MQ_Z_INSYS * R] = (double)in_sys;
        rec[(long long)MQ_Z_QUEUED * R] = (double)qsize;
        rec[(long long)MQ_Z_BUSYN * R] = (double)busy_n;
        rec[(long long)MQ_Z_POVER * R] = p_over;
        rec[(long long)MQ_Z_PREEMPTIONS * R] = (double)preempt;
        rec[(long long)MQ_Z_AUSED * R] = (double)a_idx;
        rec[(long long)MQ_Z_SUSED * R] = (double)s_idx;
        rec[(long long)MQ_Z_YUSED * R] = (double)y_idx;
        rec[(long long)MQ_Z_SLOWDOWN * R] = sd1;
        rec[(long long)(MQ_Z_SLOWDOWN + 1) * R] = sd2;
        for (int i = 0; i < 4; ++i) {
            rec[(long long)(MQ_Z_WSUM + i) * R] = w_sum[i];
            rec[(long long)(MQ_Z_VSUM + i) * R] = v_sum[i];
            rec[(long long)(MQ_Z_WRUN + i) * R] = w_run[i];
            rec[(long long)(MQ_Z_VRUN + i) * R] = v_run[i];
        }
        for (int i = 0; i < 3; ++i) rec[(long long)(MQ_Z_BUSYRUN + i) * R] = busy_run[i];
    }
}

cudaError_t k10_launch(const K10Params &P, int grid, cudaStream_t st)
{
    k10_size<<<grid, K10_BLOCK, 0, st>>>(P);
    return cudaGetLastError();
}

cudaError_t k10_occupancy(int *blocks_per_sm)
{
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k10_size, K10_BLOCK, 0);
}

}  // namespace mq
