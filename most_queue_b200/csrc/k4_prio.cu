// k4_prio.cu -- K4: multi-class priority queue kernel (NP / PR / RS / RW), one lane per replication.
//
// Walks the event sequence of PriorityQueueSimulator.run (most_queue/sim/priority.py:622-657):
//   event selection  priority.py:564-608  class 0 arrival < class 1 arrival < ... < serving on ties
//   arrival(k)       priority.py:282-332  p[k][in_sys[k]] += t - t_old[k]
//   free channel     priority.py:334-369  lowest free server index (see tests/golden/make_golden.py)
//   pre-emption      priority.py:390-455  FIRST server in index order with a weaker class; victim goes
//                                         to the TAIL of its class queue; PR keeps the remaining time,
//                                         RS redraws from the ARRIVING class's stream, RW keeps the total
//   serving(c)       priority.py:457-539  v and w are both sampled here with count served[k];
//                                         NP scans all class queues, PR/RS/RW start at the finished class
//   start_service    sim/utils/servers.py:219-250
//
// fp64 throughout, every reference operation issued in the reference's order with explicit
// round-to-nearest intrinsics, so that in REPLAY mode (variates from caller arrays,
// A[k][i][rep], S[k][j][rep]) the running-mean moments, time-in-state sums and clock are
// bit-identical to the reference.  In generator mode the same loop consumes the virtual arrays
// of the counter-based generator (stream sid = class: word 0 of call i = i-th gap, word 1 of call
// j = j-th service draw).
//
// Per-lane state lives in thread-private arrays (local memory, L1-resident); the per-class waiting
// lines are rings of QCAP job records per class in HBM scratch ([class][slot][field][lane]).
#include "k4_prio.cuh"

namespace mq {

__device__ __forceinline__ void k4_refresh(double *m, double x, long long count)
{
    // stats_update.py:6-22
    double power = x;
    const double inv_count = __ddiv_rn(1.0, (double)count);
    const double one_minus = __dsub_rn(1.0, inv_count);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m[i] = __dadd_rn(__dmul_rn(m[i], one_minus), __dmul_rn(power, inv_count));
        power = __dmul_rn(power, x);
    }
}

__device__ __forceinline__ void k4_add4(double *s, double x)
{
    double pw = x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s[i] = __dadd_rn(s[i], pw);
        pw = __dmul_rn(pw, x);
    }
}

struct K4Job {
    double arr_time, start_wait, wait_time, tte;  // tte: remaining/forced service when is_pr
    int k, is_pr;
};

__global__ void __launch_bounds__(K4_BLOCK) k4_prio(const K4Params P)
{
    const long long gtid = (long long)blockIdx.x * K4_BLOCK + threadIdx.x;
    const long long R = P.reps;
    const int c = P.c, K = P.K;
    const int rows = K4_ROWS(P.p_bins);

    for (long long rl = gtid; rl < R; rl += P.lanes) {
        const uint32_t rep = P.rep0 + (uint32_t)rl;
        double *rec = P.rec + rl;
        double *Q = P.queue + gtid;

        double end[K4_MAXC], total_time[K4_MAXC];
        K4Job on[K4_MAXC];
        uint32_t busy = 0;
        for (int j = 0; j < K4_MAXC; ++j) {
            end[j] = 1e10;
            total_time[j] = 0.0;
        }
        double arrival_time[K4_MAXK], t_old[K4_MAXK], p_over[K4_MAXK];
        double w_run[K4_MAXK][4], v_run[K4_MAXK][4], w_sum[K4_MAXK][4], v_sum[K4_MAXK][4];
        long long arrived[K4_MAXK], taked[K4_MAXK], served[K4_MAXK], dropped[K4_MAXK], in_sys[K4_MAXK];
        long long a_idx[K4_MAXK], s_idx[K4_MAXK];
        int qhead[K4_MAXK], qsize[K4_MAXK];
        int status = 0;

        auto next_gap = [&](int k) -> double {
            const long long i = a_idx[k]++;
            if (P.replay) {
                if (i >= P.n_a[k]) {
                    status = MQ_E_INPUT_SHORT;
                    return 0.0;
                }
                return P.A[P.aoff[k] + i * R + rl];
            }
            const uint32_t key = P.key0 + MQ_KEY_STRIDE * (64u * (uint32_t)k);
            uint2 w = philox2x32_10((uint32_t)i, rep, key);
            return (double)sample<-1>(P.src[k], w.x, (uint32_t)i, rep, key, 0);
        };
        auto next_service = [&](int k) -> double {
            const long long j = s_idx[k]++;
            if (P.replay) {
                if (j >= P.n_s[k]) {
                    status = MQ_E_INPUT_SHORT;
                    return 0.0;
                }
                return P.S[P.soff[k] + j * R + rl];
            }
            const uint32_t key = P.key0 + MQ_KEY_STRIDE * (64u * (uint32_t)k);
            uint2 w = philox2x32_10((uint32_t)j, rep, key);
            return (double)sample<-1>(P.srv[k], w.y, (uint32_t)j, rep, key, 1);
        };
        auto qslot = [&](int k, int slot, int f) -> double * {
            return Q + (((long long)k * P.qcap + slot) * 4 + f) * P.lanes;
        };
        auto push = [&](const K4Job &jb) {
            const int k = jb.k;
            if (qsize[k] >= P.qcap) {
                status = MQ_E_OVERFLOW;
                return;
            }
            int slot = qhead[k] + qsize[k];
            if (slot >= P.qcap) slot -= P.qcap;
            *qslot(k, slot, 0) = jb.arr_time;
            *qslot(k, slot, 1) = jb.start_wait;
            *qslot(k, slot, 2) = jb.wait_time;
            *qslot(k, slot, 3) = jb.is_pr ? jb.tte : -1.0;
            qsize[k]++;
        };
        auto pop = [&](int k) -> K4Job {
            K4Job jb;
            const int slot = qhead[k];
            jb.k = k;
            jb.arr_time = *qslot(k, slot, 0);
            jb.start_wait = *qslot(k, slot, 1);
            jb.wait_time = *qslot(k, slot, 2);
            const double t = *qslot(k, slot, 3);
            jb.is_pr = t >= 0.0;
            jb.tte = t;
            qsize[k]--;
            // an emptied line restarts at slot 0: the ring only ever touches as many slots as the line was long,
            // so its footprint stays cache-resident instead of walking all of qcap
            qhead[k] = (qsize[k] == 0 || slot + 1 >= P.qcap) ? 0 : slot + 1;
            return jb;
        };
        auto p_add = [&](int k, double dt) {
            if (in_sys[k] < P.p_bins) {
                double *gp = &rec[(long long)(K4_G_ROWS + k * rows + K4_P + in_sys[k]) * R];
                *gp = __dadd_rn(*gp, dt);
            } else {
                p_over[k] = __dadd_rn(p_over[k], dt);
            }
        };

        for (int k = 0; k < K; ++k) {
            t_old[k] = 0.0;
            p_over[k] = 0.0;
            arrived[k] = taked[k] = served[k] = dropped[k] = in_sys[k] = 0;
            a_idx[k] = s_idx[k] = 0;
            qhead[k] = qsize[k] = 0;
            for (int i = 0; i < 4; ++i) w_run[k][i] = v_run[k][i] = w_sum[k][i] = v_sum[k][i] = 0.0;
        }
        for (int k = 0; k < K; ++k) arrival_time[k] = next_gap(k);  // priority.py:130

        double ttek = 0.0;
        int free_channels = c;
        long long served_total = 0;

        while (served_total < P.jobs && status == 0) {
            // _get_min_server_time priority.py:549-562 (starts at 1e10, strict '<')
            int midx = -1;
            double mt = 1e10;
            for (int j = 0; j < c; ++j) {
                if (end[j] < mt) {
                    mt = end[j];
                    midx = j;
                }
            }
            int ev = 0;
            double et = arrival_time[0];
            for (int k = 1; k < K; ++k) {
                if (arrival_time[k] < et) {
                    ev = k;
                    et = arrival_time[k];
                }
            }
            if (midx >= 0 && mt < et) ev = K;

            // The service draw of the job that starts in this event is made at ONE site after the event logic
            // (draw_cls / draw_srv), so lanes in different branches share the sampler code.  Streams are per
            // class, so the order of draws within a class is what the reference's is.
            int draw_cls = -1, draw_srv = 0;
            if (ev < K) {
                const int k = ev;
                ttek = arrival_time[k];
                p_add(k, __dsub_rn(arrival_time[k], t_old[k]));
                arrival_time[k] = __dadd_rn(ttek, next_gap(k));
                if (status) break;
                K4Job nj;
                nj.k = k;
                nj.arr_time = ttek;
                nj.start_wait = -1.0;
                nj.wait_time = 0.0;
                nj.tte = 0.0;
                nj.is_pr = 0;
                arrived[k]++;
                in_sys[k]++;
                t_old[k] = ttek;
                if (free_channels != 0) {
                    const int j = __ffs(~busy) - 1;  // lowest free index
                    taked[k]++;
                    draw_cls = k;
                    draw_srv = j;
                    on[j] = nj;
                    busy |= 1u << j;
                    free_channels--;
                } else if (P.prty == 0) {  // NP: priority.py:331-332
                    nj.start_wait = ttek;
                    push(nj);
                } else {
                    // arrive_with_absolute_prty priority.py:390-455
                    bool found = false;
                    for (int j = 0; j < c; ++j) {
                        if (on[j].k > k) {
                            const double time_to_end = end[j];
                            const double tot = total_time[j];
                            K4Job victim = on[j];
                            taked[k]++;
                            victim.start_wait = ttek;
                            victim.is_pr = 1;
                            if (P.prty == 1)
                                victim.tte = __dsub_rn(time_to_end, ttek);
                            else if (P.prty == 2)
                                victim.tte = next_service(k);  // priority.py:425, drawn before the new job's
                            else
                                victim.tte = tot;
                            push(victim);
                            draw_cls = k;
                            draw_srv = j;
                            on[j] = nj;
                            found = true;
                            break;
                        }
                    }
                    if (!found) {
                        long long qtot = 0;
                        for (int kk = 0; kk < K; ++kk) qtot += qsize[kk];
                        if (P.buffer <= 0 || qtot < P.buffer) {
                            nj.start_wait = ttek;
                            push(nj);
                        } else {
                            dropped[k]++;
                            in_sys[k]--;
                        }
                    }
                }
            } else {
                // serving(c) priority.py:457-539
                const int j = midx;
                const double time_to_end = end[j];
                const K4Job done = on[j];
                end[j] = 1e10;
                busy &= ~(1u << j);
                total_time[j] = 0.0;
                const int k = done.k;
                p_add(k, __dsub_rn(time_to_end, t_old[k]));
                ttek = time_to_end;
                t_old[k] = ttek;
                served[k]++;
                served_total++;
                free_channels++;
                const double v = __dsub_rn(ttek, done.arr_time);
                k4_refresh(v_run[k], v, served[k]);
                k4_refresh(w_run[k], done.wait_time, served[k]);
                k4_add4(v_sum[k], v);
                k4_add4(w_sum[k], done.wait_time);
                in_sys[k]--;
                const int start = P.prty == 0 ? 0 : k;
                for (int kk = start; kk < K; ++kk) {
                    if (qsize[kk] != 0) {
                        K4Job qj = pop(kk);
                        taked[kk]++;
                        qj.wait_time = __dadd_rn(qj.wait_time, __dsub_rn(ttek, qj.start_wait));
                        if (qj.is_pr) {
                            end[j] = __dadd_rn(ttek, qj.tte);  // servers.py:237-239
                        } else {
                            draw_cls = kk;
                            draw_srv = j;
                        }
                        on[j] = qj;
                        busy |= 1u << j;
                        free_channels--;
                        break;
                    }
                }
            }
            if (draw_cls >= 0) {  // ServerPriority.start_service, servers.py:241-243
                total_time[draw_srv] = next_service(draw_cls);
                end[draw_srv] = __dadd_rn(ttek, total_time[draw_srv]);
            }
        }

        rec[(long long)K4_G_TTEK * R] = ttek;
        rec[(long long)K4_G_FLAGS * R] = (double)(status ? -status : 0);
        for (int k = 0; k < K; ++k) {
            double *rk = rec + (long long)(K4_G_ROWS + k * rows) * R;
            for (int i = 0; i < 4; ++i) {
                rk[(long long)(K4_WSUM + i) * R] = w_sum[k][i];
                rk[(long long)(K4_VSUM + i) * R] = v_sum[k][i];
                rk[(long long)(K4_WRUN + i) * R] = w_run[k][i];
                rk[(long long)(K4_VRUN + i) * R] = v_run[k][i];
            }
            rk[(long long)K4_ARRIVED * R] = (double)arrived[k];
            rk[(long long)K4_TAKED * R] = (double)taked[k];
            rk[(long long)K4_SERVED * R] = (double)served[k];
            rk[(long long)K4_DROPPED * R] = (double)dropped[k];
            rk[(long long)K4_INSYS * R] = (double)in_sys[k];
            rk[(long long)K4_QUEUED * R] = (double)qsize[k];
            rk[(long long)K4_TOLD * R] = t_old[k];
            rk[(long long)K4_POVER * R] = p_over[k];
            rk[(long long)K4_AUSED * R] = (double)a_idx[k];
            rk[(long long)K4_SUSED * R] = (double)s_idx[k];
        }
    }
}

cudaError_t k4_launch(const K4Params &P, int grid, cudaStream_t st)
{
    k4_prio<<<grid, K4_BLOCK, 0, st>>>(P);
    return cudaGetLastError();
}

cudaError_t k4_occupancy(int *blocks_per_sm)
{
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k4_prio, K4_BLOCK, 0);
}

}  // namespace mq
