// k9_negatives.cu -- K9: GI/G/c/r queue with negative arrivals, one lane per replication.
//
// Walks QsSimNegatives.run (most_queue/sim/negative.py) event for event:
//   event selection   negative.py:385-408   candidates in the order positive_arrival, negative_arrival, serving;
//                                           the first minimum wins
//   positive_arrival  negative.py:225-241   base arrival logic (queue when no channel is free)
//   negative_arrival  negative.py:243-383   no effect on an empty system;
//       DISASTER / CLEAR_SYSTEM (:277-304)  every job in service (index order), then every waiting job (FIFO
//                                           order) leaves as "broken"; waiting jobs also contribute their wait
//       RCS / REMOVE (:338-383)             one busy server, chosen uniformly, loses its job ("broken"); the head
//                                           of the waiting line takes the server
//       REQUEUE / RESUME / ..NO_RESAMPLING  (:253-276, :347-372) the interrupted jobs go back to the head of the line
//                                           and, servers being free, restart at once: one more `taked`, one more
//                                           sample of their unchanged wait, and a completion time from a new draw /
//                                           the remaining time / the original total; DISASTER hands the servers out
//                                           again in order of arrival (stable sort, lowest server first)
//   serving           negative.py:425-450   v over ALL departures counts `total`, v_served counts `served`
// The reference picks the RCS victim with Python's global random.choice; here it is the floor(u * busy)-th busy
// server in index order, u from the `choices` stream (replay array, or word 1 of generator stream 1).  fp64 with
// explicit round-to-nearest intrinsics in the reference's order: in replay mode every output is bit-identical.
// Generator streams: positive gaps = word 0 and services = word 1 of stream 0 (as the FIFO kernel), negative gaps =
// word 0 and choices = word 1 of stream 1.
#include "k9_negatives.cuh"

namespace mq {

// The four moment sets (wait, sojourn of all / served / broken jobs: 4 running means + 4 power sums each) live in shared
// memory [set][word][lane], so that one sample site can serve whichever set a lane's event needs.
enum { K9_W = 0, K9_V = 1, K9_VS = 2, K9_VB = 3 };

__device__ __forceinline__ void k9_sample(double *acc, int set, double x, long long count)
{
    // refresh_moments_stat (stats_update.py:6-22) + plain power sums
    double *run = acc + set * (8 * K9_BLOCK), *sum = run + 4 * K9_BLOCK;
    const double inv_count = __drcp_rn((double)count);
    const double one_minus = __dsub_rn(1.0, inv_count);
    double power = x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        run[i * K9_BLOCK] = __dadd_rn(__dmul_rn(run[i * K9_BLOCK], one_minus), __dmul_rn(power, inv_count));
        sum[i * K9_BLOCK] = __dadd_rn(sum[i * K9_BLOCK], power);
        power = __dmul_rn(power, x);
    }
}

// CT: server slots of the thread-private arrays (the smallest of 1 / 2 / 4 / 8 / 32 that holds c): a lane's stack is then
// a few hundred bytes instead of 1-2 KB and stays in L1
template <int CT>
__global__ void __launch_bounds__(K9_BLOCK) k9_negatives(const K9Params P)
{
    const long long gtid = (long long)blockIdx.x * K9_BLOCK + threadIdx.x;
    const long long R = P.reps;
    const int c = P.c;
    const DistF law_pos = P.positive, law_neg = P.negative, law_srv = P.service;
    __shared__ double k9_acc[4 * 8 * K9_BLOCK];
    double *acc = k9_acc + threadIdx.x;

    for (long long rl = gtid; rl < R; rl += P.lanes) {
        const uint32_t rep = P.rep0 + (uint32_t)rl;
        double *rec = P.rec + rl;
        double *Q = P.queue + gtid;
        double end[CT], arr[CT], wait[CT], tot[CT];
        uint32_t busy = 0;
        for (int j = 0; j < c; ++j) {
            end[j] = 1e10;
            arr[j] = wait[j] = tot[j] = 0.0;
        }
        long long used[MQ_NEG_STREAMS] = {0, 0, 0, 0};
        int status = 0;

        auto draw = [&](int s) -> double {
            const long long i = used[s]++;
            if (P.replay) {
                if (i >= P.n[s]) {
                    status = MQ_E_INPUT_SHORT;
                    return 0.0;
                }
                return P.X[P.off[s] + i * R + rl];
            }
            const uint32_t sid = (s == MQ_NS_NEGATIVES || s == MQ_NS_CHOICES) ? 1u : 0u;
            const uint32_t key = P.key0 + MQ_KEY_STRIDE * (64u * sid);
            uint2 w = philox2x32_10((uint32_t)i, rep, key);
            if (s == MQ_NS_POSITIVES) return (double)sample<-1>(law_pos, w.x, (uint32_t)i, rep, key, 0);
            if (s == MQ_NS_NEGATIVES) return (double)sample<-1>(law_neg, w.x, (uint32_t)i, rep, key, 0);
            if (s == MQ_NS_SERVICES) return (double)sample<-1>(law_srv, w.y, (uint32_t)i, rep, key, 1);
            return ((double)w.y + 0.5) * 0x1p-32;
        };

        double ttek = 0.0, p_over = 0.0;
        long long pos_arrived = 0, neg_arrived = 0, taked = 0, served = 0, broken = 0, total = 0, dropped = 0,
                  in_sys = 0;
        int free_channels = c, qhead = 0, qsize = 0;
        for (int i = 0; i < 32; ++i) acc[i * K9_BLOCK] = 0.0;

        auto p_add = [&](double t_new) {  // _update_state_probs base.py:327-340
            const double dt = __dsub_rn(t_new, ttek);
            if (in_sys < P.p_bins)
                atomicAdd(&rec[(long long)(MQ_N_P + in_sys) * R], dt);  // fire-and-forget fp64 add, same rounding
            else
                p_over = __dadd_rn(p_over, dt);
        };
        auto broken_sample = [&](double a) {  // negative.py:283-288
            broken++;
            total++;
            const double sj = __dsub_rn(ttek, a);
            k9_sample(acc, K9_V, sj, total);
            k9_sample(acc, K9_VB, sj, broken);
        };
        auto q_pop = [&]() -> double {
            const double a = Q[(long long)qhead * P.lanes];
            qsize--;
            qhead = (qsize == 0 || qhead + 1 >= P.qcap) ? 0 : qhead + 1;
            return a;
        };
        auto head_to_channel = [&](int ch) {  // send_head_of_queue_to_channel base.py:288-310
            const double a = q_pop();
            taked++;
            const double w = __dadd_rn(0.0, __dsub_rn(ttek, a));
            k9_sample(acc, K9_W, w, taked);
            arr[ch] = a;
            wait[ch] = w;
            tot[ch] = draw(MQ_NS_SERVICES);
            end[ch] = __dadd_rn(ttek, tot[ch]);
            busy |= 1u << ch;
            free_channels--;
        };
        // an interrupted job restarts on server ch (see the header): rem = its remaining time, t = its total
        auto restart = [&](int ch, double a, double w, double rem, double t) {
            taked++;
            const double wt = __dadd_rn(w, __dsub_rn(ttek, ttek));
            k9_sample(acc, K9_W, wt, taked);
            arr[ch] = a;
            wait[ch] = wt;
            if (P.scenario == MQ_NEG_REQUEUE) {
                tot[ch] = draw(MQ_NS_SERVICES);
                end[ch] = __dadd_rn(ttek, tot[ch]);
            } else {
                tot[ch] = t;
                end[ch] = __dadd_rn(ttek, P.scenario == MQ_NEG_RESUME ? rem : t);
            }
            busy |= 1u << ch;
            free_channels--;
        };

        double pos_time = draw(MQ_NS_POSITIVES);  // negative.py:135
        double neg_time = draw(MQ_NS_NEGATIVES);  // negative.py:153

        while (served < P.jobs && status == 0) {
            int midx = -1;
            double mt = 1e16;
            for (int j = 0; j < c; ++j) {
                if (end[j] < mt) {
                    mt = end[j];
                    midx = j;
                }
            }
            int ev = 0;  // 0 positive arrival, 1 negative arrival, 2 serving
            double et = pos_time;
            if (neg_time < et) {
                ev = 1;
                et = neg_time;
            }
            if (midx >= 0 && mt < et) ev = 2;

            if (ev == 1) {
                // negative_arrival negative.py:243-383
                neg_arrived++;
                p_add(neg_time);
                ttek = neg_time;
                neg_time = __dadd_rn(ttek, draw(MQ_NS_NEGATIVES));
                if (in_sys != 0) {
                    if (P.type == MQ_NEG_DISASTER && P.scenario != MQ_NEG_REMOVE) {
                        // the jobs in service, in order of arrival (stable), take servers 0, 1, ... again
                        int nb = 0, idx[CT];
                        double ja[CT], jw[CT], jr[CT], jt[CT];
                        for (int j = 0; j < c; ++j) {
                            if (!((busy >> j) & 1u)) continue;
                            const double rem = __dsub_rn(end[j], ttek);
                            ja[nb] = arr[j];
                            jw[nb] = wait[j];
                            jr[nb] = rem > 0.0 ? rem : 0.0;
                            jt[nb] = tot[j];
                            idx[nb] = nb;
                            nb++;
                            end[j] = 1e10;
                        }
                        for (int a = 1; a < nb; ++a) {
                            const int t = idx[a];
                            int b = a - 1;
                            while (b >= 0 && ja[idx[b]] > ja[t]) {
                                idx[b + 1] = idx[b];
                                --b;
                            }
                            idx[b + 1] = t;
                        }
                        busy = 0;
                        free_channels = c;
                        for (int q = 0; q < nb; ++q) restart(q, ja[idx[q]], jw[idx[q]], jr[idx[q]], jt[idx[q]]);
                    } else if (P.type == MQ_NEG_DISASTER) {
                        for (int j = 0; j < c; ++j) {
                            if (!((busy >> j) & 1u)) continue;
                            const double a = arr[j];
                            end[j] = 1e10;
                            broken_sample(a);
                        }
                        busy = 0;
                        in_sys = 0;
                        free_channels = c;
                        while (qsize > 0) {
                            const double a = q_pop();
                            taked++;
                            total++;
                            broken++;
                            k9_sample(acc, K9_W, __dadd_rn(0.0, __dsub_rn(ttek, a)), taked);
                            const double sj = __dsub_rn(ttek, a);
                            k9_sample(acc, K9_V, sj, total);
                            k9_sample(acc, K9_VB, sj, broken);
                        }
                    } else if (busy != 0) {
                        const int nb = __popc(busy);
                        int pick = (int)(draw(MQ_NS_CHOICES) * (double)nb);
                        if (pick >= nb) pick = nb - 1;
                        uint32_t left = busy;
                        for (int q = 0; q < pick; ++q) left &= left - 1u;  // drop the `pick` lowest busy servers
                        const int j = __ffs(left) - 1;
                        const double a = arr[j];
                        const double rem = __dsub_rn(end[j], ttek);
                        end[j] = 1e10;
                        busy &= ~(1u << j);
                        free_channels++;
                        if (P.scenario != MQ_NEG_REMOVE) {
                            restart(j, a, wait[j], rem > 0.0 ? rem : 0.0, tot[j]);
                        } else {
                            broken_sample(a);
                            in_sys--;
                            if (qsize != 0) head_to_channel(j);
                        }
                    }
                }
            } else {
                // positive_arrival (negative.py:225-241) and serving (negative.py:425-450) through ONE body: the clock and
                // time-in-state update, at most one "a job starts on a server" site (the arriving job on the lowest free
                // server, or the head of the line on the server that just finished) with its wait sample and its
                // service draw, and the sojourn samples of a departure
                const bool is_arr = ev == 0;
                const int jd = is_arr ? 0 : midx;
                const double te = is_arr ? pos_time : end[jd];
                const double a_dep = arr[jd];
                if (!is_arr) {
                    end[jd] = 1e10;
                    busy &= ~(1u << jd);
                }
                p_add(te);
                ttek = te;
                bool start;
                int slot;
                double a_job = ttek;
                if (is_arr) {
                    pos_arrived++;
                    in_sys++;
                    pos_time = __dadd_rn(ttek, draw(MQ_NS_POSITIVES));
                    start = free_channels != 0;
                    slot = __ffs(~busy) - 1;  // send_task_to_channel base.py:155-184, lowest free channel
                    if (!start) {
                        // send_task_to_queue base.py:186-212
                        if (P.buffer < 0 || qsize < P.buffer) {
                            if (qsize >= P.qcap) {
                                status = MQ_E_OVERFLOW;
                            } else {
                                int qs = qhead + qsize;
                                if (qs >= P.qcap) qs -= P.qcap;
                                Q[(long long)qs * P.lanes] = ttek;
                                qsize++;
                            }
                        } else {
                            dropped++;
                            in_sys--;
                        }
                    }
                } else {
                    free_channels++;
                    served++;
                    total++;
                    in_sys--;
                    start = qsize != 0;  // send_head_of_queue_to_channel base.py:288-310
                    slot = jd;
                    if (start) a_job = q_pop();
                    const double sj = __dsub_rn(ttek, a_dep);
                    k9_sample(acc, K9_V, sj, total);
                    k9_sample(acc, K9_VS, sj, served);
                }
                if (start) {
                    taked++;
                    const double w = is_arr ? 0.0 : __dadd_rn(0.0, __dsub_rn(ttek, a_job));
                    k9_sample(acc, K9_W, w, taked);
                    arr[slot] = a_job;
                    wait[slot] = w;
                    tot[slot] = draw(MQ_NS_SERVICES);
                    end[slot] = __dadd_rn(ttek, tot[slot]);
                    busy |= 1u << slot;
                    free_channels--;
                }
            }
        }

        rec[(long long)MQ_N_TTEK * R] = ttek;
        rec[(long long)MQ_N_FLAGS * R] = (double)(status ? -status : 0);
        rec[(long long)MQ_N_POS_ARRIVED * R] = (double)pos_arrived;
        rec[(long long)MQ_N_NEG_ARRIVED * R] = (double)neg_arrived;
        rec[(long long)MQ_N_TAKED * R] = (double)taked;
        rec[(long long)MQ_N_SERVED * R] = (double)served;
        rec[(long long)MQ_N_BROKEN * R] = (double)broken;
        rec[(long long)MQ_N_TOTAL * R] = (double)total;
        rec[(long long)MQ_N_DROPPED * R] = (double)dropped;
        rec[(long long)MQ_N_INSYS * R] = (double)in_sys;
        rec[(long long)MQ_N_QUEUED * R] = (double)qsize;
        rec[(long long)MQ_N_POVER * R] = p_over;
        for (int i = 0; i < 4; ++i) {
            rec[(long long)(MQ_N_WSUM + i) * R] = acc[(K9_W * 8 + 4 + i) * K9_BLOCK];
            rec[(long long)(MQ_N_VSUM + i) * R] = acc[(K9_V * 8 + 4 + i) * K9_BLOCK];
            rec[(long long)(MQ_N_VSSUM + i) * R] = acc[(K9_VS * 8 + 4 + i) * K9_BLOCK];
            rec[(long long)(MQ_N_VBSUM + i) * R] = acc[(K9_VB * 8 + 4 + i) * K9_BLOCK];
            rec[(long long)(MQ_N_WRUN + i) * R] = acc[(K9_W * 8 + i) * K9_BLOCK];
            rec[(long long)(MQ_N_VRUN + i) * R] = acc[(K9_V * 8 + i) * K9_BLOCK];
            rec[(long long)(MQ_N_VSRUN + i) * R] = acc[(K9_VS * 8 + i) * K9_BLOCK];
            rec[(long long)(MQ_N_VBRUN + i) * R] = acc[(K9_VB * 8 + i) * K9_BLOCK];
        }
        for (int s = 0; s < MQ_NEG_STREAMS; ++s) rec[(long long)(MQ_N_USED + s) * R] = (double)used[s];
    }
}

cudaError_t k9_launch(const K9Params &P, int grid, cudaStream_t st)
{
#define K9_GO(n) k9_negatives<n><<<grid, K9_BLOCK, 0, st>>>(P)
    if (P.c <= 1) K9_GO(1); else if (P.c <= 2) K9_GO(2); else if (P.c <= 4) K9_GO(4); else if (P.c <= 8) K9_GO(8); else K9_GO(32);
#undef K9_GO
    return cudaGetLastError();
}

cudaError_t k9_occupancy(int c, int *blocks_per_sm)
{
#define K9_OCC(n) cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k9_negatives<n>, K9_BLOCK, 0)
    return c <= 1 ? K9_OCC(1) : c <= 2 ? K9_OCC(2) : c <= 4 ? K9_OCC(4) : c <= 8 ? K9_OCC(8) : K9_OCC(32);
#undef K9_OCC
}

}  // namespace mq
