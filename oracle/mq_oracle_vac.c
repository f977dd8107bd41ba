/*
 * mq_oracle_vac.c -- CPU restatement of VacationQueueingSystemSimulator.run and NPolicyQueueSim.run
 * (most_queue/sim/vacations.py:12-364): QsSim (sim/base.py) with three extra event sources -- the end of a
 * warm-up, of a cold (vacation) period and of a cold-delay period.  TEST INFRASTRUCTURE ONLY.
 *
 * Event selection (base.py:425-452): the candidates are collected in the order arrival, serving, warm_up_end,
 * cold_end, cold_delay_end and the first minimum wins.  Free server = lowest index (tests/golden/make_golden.py
 * pins the reference's set iteration to that).
 */
#include "mq_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline void vac_refresh(double *m, int n, double x, int64_t count)
{
    /* stats_update.py:6-22 */
    double power = x;
    const double inv_count = 1.0 / (double)count;
    const double one_minus = 1.0 - inv_count;
    for (int i = 0; i < n; ++i) {
        m[i] = m[i] * one_minus + power * inv_count;
        power = power * x;
    }
}

typedef struct {
    int is_set, is_start;
    double end_time, prob, start_mom;
    int64_t starts;
} phase;

static void phase_start(phase *p, double ttek, double dur)
{ /* phase.py:38-47 */
    p->is_start = 1;
    p->start_mom = ttek;
    p->starts += 1;
    p->end_time = ttek + dur;
}

static void phase_end(phase *p, double ttek)
{ /* phase.py:49-57 */
    p->prob += ttek - p->start_mom;
    p->is_start = 0;
    p->end_time = 1e16;
}

int mqo_vac_run(const mqo_vac_cfg *cfg, mqo_src *arrivals, mqo_src *services, mqo_src *warm_services,
                mqo_src *warm_src, mqo_src *cold_src, mqo_src *delay_src, int64_t total_served, double *p,
                int64_t p_bins, mqo_vac_out *out)
{
    const int c = cfg->c;
    if (c < 1 || c > 64) return -2;
    memset(out, 0, sizeof(*out));
    for (int64_t i = 0; i < p_bins; ++i) p[i] = 0.0;
    double end[64], arr[64];
    int is_free[64];
    for (int j = 0; j < c; ++j) { end[j] = 1e10; arr[j] = 0.0; is_free[j] = 1; }
    int64_t qcap = 1024, qhead = 0, qsize = 0;
    double *queue = (double *)malloc(sizeof(double) * (size_t)qcap);
    phase warm = {cfg->warm_set, 0, 1e16, 0.0, 0.0, 0}, cold = {cfg->cold_set || cfg->big_n > 0, 0, 1e16, 0.0, 0.0, 0},
          delay = {cfg->delay_set, 0, 1e16, 0.0, 0.0, 0};
    int err = 0;
    double ttek = 0.0, start_busy = 0.0;
    int64_t arrived = 0, taked = 0, served = 0, dropped = 0, in_sys = 0, busy_n = 0, warm_after_cold = 0;
    int free_channels = c;
    double w_run[4] = {0}, v_run[4] = {0}, w_sum[4] = {0}, v_sum[4] = {0}, busy_run[3] = {0};
    double p_over = 0.0;
    double arrival_time = mqo_src_next(arrivals, &err); /* base.py:112 */
    if (cfg->big_n > 0) phase_start(&cold, 0.0, 1e100); /* vacations.py:345-347: set_cold(1e100, "D"); start(0) */

/* N-policy: the "cold" law is D(1e100) (vacations.py:345), a constant that consumes no draw */
#define COLD_DUR() (cfg->big_n > 0 ? 1e100 : mqo_src_next(cold_src, &err))
#define P_ADD(t_new)                                              \
    do {                                                          \
        const double dt_ = (t_new) - ttek;                        \
        if (in_sys < p_bins) p[in_sys] += dt_; else p_over += dt_; \
    } while (0)
#define Q_PUSH()                                                                   \
    do { /* send_task_to_queue base.py:186-212 */                                  \
        if (cfg->buffer < 0 || qsize < cfg->buffer) {                              \
            if (qsize == qcap) {                                                   \
                double *nq = (double *)malloc(sizeof(double) * (size_t)qcap * 2);  \
                for (int64_t i_ = 0; i_ < qsize; ++i_) nq[i_] = queue[(qhead + i_) % qcap]; \
                free(queue); queue = nq; qhead = 0; qcap *= 2;                     \
            }                                                                      \
            queue[(qhead + qsize) % qcap] = ttek;                                  \
            qsize++;                                                               \
        } else {                                                                   \
            dropped++;                                                             \
            in_sys--;                                                              \
        }                                                                          \
    } while (0)
#define TO_CHANNEL(is_warm_)                                                        \
    do { /* send_task_to_channel base.py:155-184 */                                 \
        int j_ = 0;                                                                 \
        while (!is_free[j_]) ++j_;                                                  \
        taked++;                                                                    \
        vac_refresh(w_run, 4, 0.0, taked);                                          \
        for (int i_ = 0; i_ < 4; ++i_) w_sum[i_] += 0.0;                            \
        arr[j_] = ttek;                                                             \
        end[j_] = ttek + mqo_src_next((is_warm_) ? warm_services : services, &err); \
        is_free[j_] = 0;                                                            \
        free_channels--;                                                            \
        if (free_channels == 0 && in_sys == c) start_busy = ttek;                   \
    } while (0)
#define HEAD_TO_CHANNEL(ch_)                                                    \
    do { /* send_head_of_queue_to_channel base.py:288-310 */                    \
        const double a_ = queue[qhead];                                         \
        qhead = (qhead + 1) % qcap;                                             \
        qsize--;                                                                \
        if (free_channels == 1) start_busy = ttek;                              \
        taked++;                                                                \
        const double w_ = 0.0 + (ttek - a_);                                    \
        vac_refresh(w_run, 4, w_, taked);                                       \
        { double pw_ = w_; for (int i_ = 0; i_ < 4; ++i_) { w_sum[i_] += pw_; pw_ *= w_; } } \
        arr[ch_] = a_;                                                          \
        end[ch_] = ttek + mqo_src_next(services, &err);                         \
        is_free[ch_] = 0;                                                       \
        free_channels--;                                                        \
    } while (0)
#define ON_END_COLD()                                                           \
    do { /* vacations.py:252-282 */                                             \
        P_ADD(cold.end_time);                                                   \
        ttek = cold.end_time;                                                   \
        phase_end(&cold, ttek);                                                 \
        if (cfg->multiple_vacations && qsize == 0) {                            \
            phase_start(&cold, ttek, COLD_DUR());             \
        } else if (warm.is_set) {                                               \
            if (qsize != 0) {                                                   \
                warm_after_cold++;                                              \
                phase_start(&warm, ttek, mqo_src_next(warm_src, &err));         \
            }                                                                   \
        } else {                                                                \
            for (int i_ = 0; i_ < c; ++i_)                                      \
                if (qsize != 0) HEAD_TO_CHANNEL(i_);                            \
        }                                                                       \
    } while (0)

    while (served < total_served && !err) {
        /* _get_min_server_time base.py:312-325 */
        int midx = -1;
        double mt = 1e16;
        for (int j = 0; j < c; ++j)
            if (end[j] < mt) { mt = end[j]; midx = j; }
        int ev = 0; /* 0 arrival, 1 serving, 2 warm end, 3 cold end, 4 cold delay end */
        double et = arrival_time;
        if (midx >= 0 && mt < et) { ev = 1; et = mt; }
        if (warm.is_set && warm.is_start && warm.end_time < et) { ev = 2; et = warm.end_time; }
        if (cold.is_set && cold.is_start && cold.end_time < et) { ev = 3; et = cold.end_time; }
        if (delay.is_set && delay.is_start && delay.end_time < et) { ev = 4; et = delay.end_time; }

        if (ev == 0) {
            /* arrival() vacations.py:126-189 */
            arrived++;
            P_ADD(arrival_time);
            in_sys++;
            ttek = arrival_time;
            arrival_time = ttek + mqo_src_next(arrivals, &err);
            if (free_channels == 0) {
                Q_PUSH();
            } else if (cold.is_set && cold.is_start) {
                Q_PUSH();
            } else if (delay.is_set && delay.is_start) {
                delay.is_start = 0;
                delay.end_time = 1e16;
                delay.prob += ttek - delay.start_mom;
                TO_CHANNEL(0);
            } else if (warm.is_set) {
                if (warm.is_start) {
                    Q_PUSH();
                } else if (free_channels == c) {
                    if (cfg->service_on_warm_up) {
                        TO_CHANNEL(1);
                    } else {
                        phase_start(&warm, ttek, mqo_src_next(warm_src, &err));
                        Q_PUSH();
                    }
                } else {
                    TO_CHANNEL(0);
                }
            } else {
                TO_CHANNEL(0);
            }
            /* NPolicyQueueSim.arrival vacations.py:354-364 */
            if (cfg->big_n > 0 && cold.is_start && in_sys >= cfg->big_n) {
                cold.end_time = ttek;
                ON_END_COLD();
            }
        } else if (ev == 1) {
            /* serving() vacations.py:191-234 */
            const double time_to_end = end[midx];
            const double a_dep = arr[midx];
            end[midx] = 1e10;
            is_free[midx] = 1;
            P_ADD(time_to_end);
            ttek = time_to_end;
            served++;
            free_channels++;
            const double v = ttek - a_dep;
            vac_refresh(v_run, 4, v, served);
            { double pw = v; for (int i = 0; i < 4; ++i) { v_sum[i] += pw; pw *= v; } }
            in_sys--;
            if (qsize == 0 && free_channels == 1 && in_sys == c - 1) {
                busy_n++;
                vac_refresh(busy_run, 3, ttek - start_busy, busy_n);
            }
            if (cold.is_set && qsize == 0 && free_channels == c) {
                if (delay.is_set)
                    phase_start(&delay, ttek, mqo_src_next(delay_src, &err));
                else
                    phase_start(&cold, ttek, COLD_DUR());
            }
            if (qsize != 0) HEAD_TO_CHANNEL(midx);
        } else if (ev == 2) {
            /* on_end_warming vacations.py:236-250 */
            P_ADD(warm.end_time);
            ttek = warm.end_time;
            phase_end(&warm, ttek);
            for (int i = 0; i < c; ++i)
                if (qsize != 0) HEAD_TO_CHANNEL(i);
        } else if (ev == 3) {
            ON_END_COLD();
        } else {
            /* on_end_cold_delay vacations.py:284-296 */
            P_ADD(delay.end_time);
            ttek = delay.end_time;
            phase_end(&delay, ttek);
            phase_start(&cold, ttek, COLD_DUR());
        }
    }
    free(queue);
    memcpy(out->w_run, w_run, sizeof(w_run));
    memcpy(out->v_run, v_run, sizeof(v_run));
    memcpy(out->w_sum, w_sum, sizeof(w_sum));
    memcpy(out->v_sum, v_sum, sizeof(v_sum));
    memcpy(out->busy_run, busy_run, sizeof(busy_run));
    out->busy_n = busy_n;
    out->arrived = arrived;
    out->taked = taked;
    out->served = served;
    out->dropped = dropped;
    out->in_sys = in_sys;
    out->queued = qsize;
    out->ttek = ttek;
    out->p_over = p_over;
    out->warm_prob = warm.prob;
    out->cold_prob = cold.prob;
    out->delay_prob = delay.prob;
    out->warm_starts = warm.starts;
    out->cold_starts = cold.starts;
    out->delay_starts = delay.starts;
    out->warm_after_cold = warm_after_cold;
    return err ? -1 : 0;
}
