// k8_vacation.cu -- K8: GI/G/c/r queue with warm-up, cold (vacation) and cold-delay periods, and the N-policy
// queue; one lane per replication.
//
// Walks VacationQueueingSystemSimulator.run / NPolicyQueueSim.run (most_queue/sim/vacations.py) event for event:
//   event selection  sim/base.py:425-452     candidates in the order arrival, serving, warm_up_end, cold_end,
//                                            cold_delay_end; the first minimum wins
//   arrival          vacations.py:126-189    queue while cold / warming; a job that arrives during the cold delay
//                                            cancels it; an empty, switched-off system starts a warm-up (or serves
//                                            the first job with the warm law, is_service_on_warm_up)
//   serving          vacations.py:191-234    base serving + "the system emptied: start the cold delay / the cold"
//   period ends      vacations.py:236-296    warm end: up to c waiting jobs enter service on channels 0..c-1;
//                                            cold end: another vacation (multiple), a warm-up if somebody waits,
//                                            or straight to service; delay end: the cold period starts
//   N-policy         vacations.py:320-364    the cold law is the constant 1e100, ended by the N-th job present
//   QsPhase          sim/utils/phase.py      prob accumulates only when a period ENDS
// fp64 with explicit round-to-nearest intrinsics in the reference's order: in replay mode every output equals the
// reference's bit for bit.  Generator streams: arrivals = word 0 and services = word 1 of stream 0 (as the FIFO
// kernel), warm services = word 1 of stream 1, warm / cold / delay periods = word 0 of streams 2 / 3 / 4.
#include "k8_vacation.cuh"

namespace mq {

__device__ __forceinline__ void k8_refresh(double *m, int n, double x, long long count)
{
    // stats_update.py:6-22
    double power = x;
    const double inv_count = __drcp_rn((double)count);
    const double one_minus = __dsub_rn(1.0, inv_count);
    for (int i = 0; i < n; ++i) {
        m[i] = __dadd_rn(__dmul_rn(m[i], one_minus), __dmul_rn(power, inv_count));
        power = __dmul_rn(power, x);
    }
}

__device__ __forceinline__ void k8_add4(double *s, double x)
{
    double pw = x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s[i] = __dadd_rn(s[i], pw);
        pw = __dmul_rn(pw, x);
    }
}

struct K8Phase {
    int is_start;
    double end_time, prob, start_mom;
    long long starts;
};

// CT: server slots of the thread-private arrays (the smallest of 1 / 2 / 4 / 8 / 32 that holds c): a lane's stack is then
// a few hundred bytes instead of 1-2 KB and stays in L1
template <int CT>
__global__ void __launch_bounds__(K8_BLOCK) k8_vacation(const K8Params P)
{
    const long long gtid = (long long)blockIdx.x * K8_BLOCK + threadIdx.x;
    const long long R = P.reps;
    const int c = P.c;
    const bool warm_set = P.set[MQ_VS_WARM] != 0 || P.set[MQ_VS_WARM_SERVICES] != 0, cold_set = P.set[MQ_VS_COLD] != 0 || P.big_n > 0,
               delay_set = P.set[MQ_VS_DELAY] != 0;

    for (long long rl = gtid; rl < R; rl += P.lanes) {
        const uint32_t rep = P.rep0 + (uint32_t)rl;
        double *rec = P.rec + rl;
        double *Q = P.queue + gtid;
        double end[CT], arr[CT];
        uint32_t busy = 0;
        for (int j = 0; j < c; ++j) {
            end[j] = 1e10;
            arr[j] = 0.0;
        }
        long long used[MQ_VAC_STREAMS] = {0, 0, 0, 0, 0, 0};
        int status = 0;

        // stream s: replay array or counter-based generator (sid, word) as listed above
        auto draw = [&](int s) -> double {
            const long long i = used[s]++;
            if (P.replay) {
                if (i >= P.n[s]) {
                    status = MQ_E_INPUT_SHORT;
                    return 0.0;
                }
                return P.X[P.off[s] + i * R + rl];
            }
            const uint32_t sid = s <= MQ_VS_SERVICES ? 0u : (uint32_t)(s - 1);
            const int which = (s == MQ_VS_SERVICES || s == MQ_VS_WARM_SERVICES) ? 1 : 0;
            const uint32_t key = P.key0 + MQ_KEY_STRIDE * (64u * sid);
            uint2 w = philox2x32_10((uint32_t)i, rep, key);
            return (double)sample<-1>(P.dist[s], which ? w.y : w.x, (uint32_t)i, rep, key, which);
        };

        K8Phase warm = {0, 1e16, 0.0, 0.0, 0}, cold = {0, 1e16, 0.0, 0.0, 0}, delay = {0, 1e16, 0.0, 0.0, 0};
        auto phase_start = [&](K8Phase &ph, double ttek, double dur) {  // phase.py:38-47
            ph.is_start = 1;
            ph.start_mom = ttek;
            ph.starts++;
            ph.end_time = __dadd_rn(ttek, dur);
        };
        auto phase_end = [&](K8Phase &ph, double ttek) {  // phase.py:49-57
            ph.prob = __dadd_rn(ph.prob, __dsub_rn(ttek, ph.start_mom));
            ph.is_start = 0;
            ph.end_time = 1e16;
        };
        auto cold_duration = [&]() -> double { return P.big_n > 0 ? 1e100 : draw(MQ_VS_COLD); };

        double ttek = 0.0, start_busy = 0.0, p_over = 0.0;
        long long arrived = 0, taked = 0, served = 0, dropped = 0, in_sys = 0, busy_n = 0, warm_after_cold = 0;
        int free_channels = c, qhead = 0, qsize = 0;
        double w_run[4] = {0, 0, 0, 0}, v_run[4] = {0, 0, 0, 0}, w_sum[4] = {0, 0, 0, 0}, v_sum[4] = {0, 0, 0, 0};
        double busy_run[3] = {0, 0, 0};

        auto p_add = [&](double t_new) {  // _update_state_probs base.py:327-340
            const double dt = __dsub_rn(t_new, ttek);
            if (in_sys < P.p_bins)
                atomicAdd(&rec[(long long)(MQ_V_P + in_sys) * R], dt);  // fire-and-forget fp64 add, same rounding
            else
                p_over = __dadd_rn(p_over, dt);
        };
        auto to_queue = [&]() {  // send_task_to_queue base.py:186-212
            if (P.buffer < 0 || qsize < P.buffer) {
                if (qsize >= P.qcap) {
                    status = MQ_E_OVERFLOW;
                    return;
                }
                int slot = qhead + qsize;
                if (slot >= P.qcap) slot -= P.qcap;
                Q[(long long)slot * P.lanes] = ttek;
                qsize++;
            } else {
                dropped++;
                in_sys--;
            }
        };
        auto head_to_channel = [&](int ch) {  // send_head_of_queue_to_channel base.py:288-310
            const double a = Q[(long long)qhead * P.lanes];
            qsize--;
            qhead = (qsize == 0 || qhead + 1 >= P.qcap) ? 0 : qhead + 1;
            if (free_channels == 1) start_busy = ttek;
            taked++;
            const double w = __dadd_rn(0.0, __dsub_rn(ttek, a));
            k8_refresh(w_run, 4, w, taked);
            k8_add4(w_sum, w);
            arr[ch] = a;
            end[ch] = __dadd_rn(ttek, draw(MQ_VS_SERVICES));
            busy |= 1u << ch;
            free_channels--;
        };
        auto on_end_cold = [&]() {  // vacations.py:252-282
            p_add(cold.end_time);
            ttek = cold.end_time;
            phase_end(cold, ttek);
            if (P.multiple_vacations && qsize == 0) {
                phase_start(cold, ttek, cold_duration());
            } else if (warm_set) {
                if (qsize != 0) {
                    warm_after_cold++;
                    phase_start(warm, ttek, draw(MQ_VS_WARM));
                }
            } else {
                for (int i = 0; i < c; ++i)
                    if (qsize != 0) head_to_channel(i);
            }
        };

        double arrival_time = draw(MQ_VS_ARRIVALS);             // base.py:112
        if (P.big_n > 0) phase_start(cold, 0.0, 1e100);         // vacations.py:345-347

        while (served < P.jobs && status == 0) {
            // _get_min_server_time base.py:312-325 (strict '<', lowest index wins ties)
            int midx = -1;
            double mt = 1e16;
            for (int j = 0; j < c; ++j) {
                if (end[j] < mt) {
                    mt = end[j];
                    midx = j;
                }
            }
            int ev = 0;  // 0 arrival, 1 serving, 2 warm-up end, 3 cold end, 4 cold-delay end
            double et = arrival_time;
            if (midx >= 0 && mt < et) {
                ev = 1;
                et = mt;
            }
            if (warm_set && warm.is_start && warm.end_time < et) {
                ev = 2;
                et = warm.end_time;
            }
            if (cold_set && cold.is_start && cold.end_time < et) {
                ev = 3;
                et = cold.end_time;
            }
            if (delay_set && delay.is_start && delay.end_time < et) {
                ev = 4;
                et = delay.end_time;
            }

            if (ev <= 1) {
                // arrival() (vacations.py:126-189) and serving() (vacations.py:191-234) through ONE body: the clock and
                // time-in-state update, then at most one "a job starts on a server" site -- the arriving job on the
                // lowest free server (send_task_to_channel base.py:155-184) or the head of the line on the server that
                // just finished (send_head_of_queue_to_channel base.py:288-310) -- with its wait sample and service draw
                const bool is_arr = ev == 0;
                const int jd = is_arr ? 0 : midx;
                const double te = is_arr ? arrival_time : end[jd];
                const double a_dep = arr[jd];
                if (!is_arr) {
                    end[jd] = 1e10;
                    busy &= ~(1u << jd);
                }
                p_add(te);
                ttek = te;
                bool start = false, warm_srv = false;
                int slot = jd;
                double a_job = ttek;
                if (is_arr) {
                    arrived++;
                    in_sys++;
                    arrival_time = __dadd_rn(ttek, draw(MQ_VS_ARRIVALS));
                    bool queue = false;
                    if (free_channels == 0) {
                        queue = true;
                    } else if (cold_set && cold.is_start) {
                        queue = true;
                    } else if (delay_set && delay.is_start) {
                        delay.is_start = 0;
                        delay.end_time = 1e16;
                        delay.prob = __dadd_rn(delay.prob, __dsub_rn(ttek, delay.start_mom));
                        start = true;
                    } else if (warm_set) {
                        if (warm.is_start) {
                            queue = true;
                        } else if (free_channels == c) {
                            if (P.service_on_warm_up) {
                                start = true;
                                warm_srv = true;
                            } else {
                                phase_start(warm, ttek, draw(MQ_VS_WARM));
                                queue = true;
                            }
                        } else {
                            start = true;
                        }
                    } else {
                        start = true;
                    }
                    if (queue) to_queue();
                    slot = __ffs(~busy) - 1;
                } else {
                    served++;
                    free_channels++;
                    const double v = __dsub_rn(ttek, a_dep);
                    k8_refresh(v_run, 4, v, served);
                    k8_add4(v_sum, v);
                    in_sys--;
                    if (qsize == 0 && free_channels == 1 && in_sys == c - 1) {
                        busy_n++;
                        k8_refresh(busy_run, 3, __dsub_rn(ttek, start_busy), busy_n);
                    }
                    if (cold_set && qsize == 0 && free_channels == c) {
                        if (delay_set)
                            phase_start(delay, ttek, draw(MQ_VS_DELAY));
                        else
                            phase_start(cold, ttek, cold_duration());
                    }
                    start = qsize != 0;
                    if (start) {
                        a_job = Q[(long long)qhead * P.lanes];
                        qsize--;
                        qhead = (qsize == 0 || qhead + 1 >= P.qcap) ? 0 : qhead + 1;
                        if (free_channels == 1) start_busy = ttek;
                    }
                }
                if (start) {
                    taked++;
                    const double w = is_arr ? 0.0 : __dadd_rn(0.0, __dsub_rn(ttek, a_job));
                    k8_refresh(w_run, 4, w, taked);
                    k8_add4(w_sum, w);
                    arr[slot] = a_job;
                    end[slot] = __dadd_rn(ttek, draw(warm_srv ? MQ_VS_WARM_SERVICES : MQ_VS_SERVICES));
                    busy |= 1u << slot;
                    free_channels--;
                    if (is_arr && free_channels == 0 && in_sys == c) start_busy = ttek;
                }
                // NPolicyQueueSim.arrival vacations.py:354-364
                if (is_arr && P.big_n > 0 && cold.is_start && in_sys >= P.big_n) {
                    cold.end_time = ttek;
                    on_end_cold();
                }
            } else if (ev == 2) {
                // on_end_warming vacations.py:236-250
                p_add(warm.end_time);
                ttek = warm.end_time;
                phase_end(warm, ttek);
                for (int i = 0; i < c; ++i)
                    if (qsize != 0) head_to_channel(i);
            } else if (ev == 3) {
                on_end_cold();
            } else {
                // on_end_cold_delay vacations.py:284-296
                p_add(delay.end_time);
                ttek = delay.end_time;
                phase_end(delay, ttek);
                phase_start(cold, ttek, cold_duration());
            }
        }

        rec[(long long)MQ_V_TTEK * R] = ttek;
        rec[(long long)MQ_V_FLAGS * R] = (double)(status ? -status : 0);
        rec[(long long)MQ_V_ARRIVED * R] = (double)arrived;
        rec[(long long)MQ_V_TAKED * R] = (double)taked;
        rec[(long long)MQ_V_SERVED * R] = (double)served;
        rec[(long long)MQ_V_DROPPED * R] = (double)dropped;
        rec[(long long)MQ_V_INSYS * R] = (double)in_sys;
        rec[(long long)MQ_V_QUEUED * R] = (double)qsize;
        rec[(long long)MQ_V_BUSYN * R] = (double)busy_n;
        rec[(long long)MQ_V_POVER * R] = p_over;
        for (int i = 0; i < 4; ++i) {
            rec[(long long)(MQ_V_WSUM + i) * R] = w_sum[i];
            rec[(long long)(MQ_V_VSUM + i) * R] = v_sum[i];
            rec[(long long)(MQ_V_WRUN + i) * R] = w_run[i];
            rec[(long long)(MQ_V_VRUN + i) * R] = v_run[i];
        }
        for (int i = 0; i < 3; ++i) rec[(long long)(MQ_V_BUSYRUN + i) * R] = busy_run[i];
        rec[(long long)MQ_V_WARMPROB * R] = warm.prob;
        rec[(long long)MQ_V_COLDPROB * R] = cold.prob;
        rec[(long long)MQ_V_DELAYPROB * R] = delay.prob;
        rec[(long long)(MQ_V_STARTS + 0) * R] = (double)warm.starts;
        rec[(long long)(MQ_V_STARTS + 1) * R] = (double)cold.starts;
        rec[(long long)(MQ_V_STARTS + 2) * R] = (double)delay.starts;
        rec[(long long)MQ_V_WARM_AFTER_COLD * R] = (double)warm_after_cold;
        for (int s = 0; s < MQ_VAC_STREAMS; ++s) rec[(long long)(MQ_V_USED + s) * R] = (double)used[s];
    }
}

cudaError_t k8_launch(const K8Params &P, int grid, cudaStream_t st)
{
#define K8_GO(n) k8_vacation<n><<<grid, K8_BLOCK, 0, st>>>(P)
    if (P.c <= 1) K8_GO(1); else if (P.c <= 2) K8_GO(2); else if (P.c <= 4) K8_GO(4); else if (P.c <= 8) K8_GO(8); else K8_GO(32);
#undef K8_GO
    return cudaGetLastError();
}

cudaError_t k8_occupancy(int c, int *blocks_per_sm)
{
#define K8_OCC(n) cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k8_vacation<n>, K8_BLOCK, 0)
    return c <= 1 ? K8_OCC(1) : c <= 2 ? K8_OCC(2) : c <= 4 ? K8_OCC(4) : c <= 8 ? K8_OCC(8) : K8_OCC(32);
#undef K8_OCC
}

}  // namespace mq
