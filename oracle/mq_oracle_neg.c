/*
 * mq_oracle_neg.c -- CPU restatement of QsSimNegatives.run (most_queue/sim/negative.py:57-575) for the two
 * negative-arrival types whose reference code is well defined: DISASTER with the CLEAR_SYSTEM scenario
 * (:277-304) and RCS with the REMOVE scenario (:338-383).  TEST INFRASTRUCTURE ONLY.
 *
 * Event selection (negative.py:385-408, base.py:438-452): candidates in the order positive_arrival,
 * negative_arrival, serving; the first minimum wins.  Free server = lowest index.  RCS picks the
 * floor(u * busy)-th busy server in index order, u from the `choice` stream (the reference calls Python's
 * global random.choice, negative.py:345; the golden fixtures replace it by exactly this rule).
 */
#include "mq_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline void neg_refresh(double *m, double x, int64_t count)
{
    double power = x;
    const double inv_count = 1.0 / (double)count;
    const double one_minus = 1.0 - inv_count;
    for (int i = 0; i < 4; ++i) {
        m[i] = m[i] * one_minus + power * inv_count;
        power = power * x;
    }
}

static inline void neg_add4(double *s, double x)
{
    double pw = x;
    for (int i = 0; i < 4; ++i) { s[i] += pw; pw *= x; }
}

int mqo_neg_run(int c, int64_t buffer, int type, int scenario, mqo_src *positives, mqo_src *negatives, mqo_src *services,
                mqo_src *choices, int64_t total_served, double *p, int64_t p_bins, mqo_neg_out *out)
{
    if (c < 1 || c > 64 || (type != MQO_NEG_DISASTER && type != MQO_NEG_RCS) || scenario < 1 || scenario > 4) return -2;
    memset(out, 0, sizeof(*out));
    for (int64_t i = 0; i < p_bins; ++i) p[i] = 0.0;
    double end[64], arr[64], wait[64], tot[64];
    int is_free[64];
    for (int j = 0; j < c; ++j) { end[j] = 1e10; arr[j] = 0.0; wait[j] = 0.0; tot[j] = 0.0; is_free[j] = 1; }
    int64_t qcap = 1024, qhead = 0, qsize = 0;
    double *queue = (double *)malloc(sizeof(double) * (size_t)qcap);
    int err = 0;
    double ttek = 0.0, p_over = 0.0;
    int64_t pos_arrived = 0, neg_arrived = 0, taked = 0, served = 0, broken = 0, total = 0, dropped = 0, in_sys = 0;
    int free_channels = c;
    double w_run[4] = {0}, v_run[4] = {0}, vs_run[4] = {0}, vb_run[4] = {0};
    double w_sum[4] = {0}, v_sum[4] = {0}, vs_sum[4] = {0}, vb_sum[4] = {0};
    double pos_time = mqo_src_next(positives, &err); /* negative.py:135 */
    double neg_time = mqo_src_next(negatives, &err); /* negative.py:153 */

#define P_ADD(t_new)                                              \
    do {                                                          \
        const double dt_ = (t_new) - ttek;                        \
        if (in_sys < p_bins) p[in_sys] += dt_; else p_over += dt_; \
    } while (0)
#define BROKEN_SAMPLE(a_)                                 \
    do { /* negative.py:283-288 */                        \
        broken++;                                         \
        total++;                                          \
        const double sj_ = ttek - (a_);                   \
        neg_refresh(v_run, sj_, total);                   \
        neg_add4(v_sum, sj_);                             \
        neg_refresh(vb_run, sj_, broken);                 \
        neg_add4(vb_sum, sj_);                            \
    } while (0)
#define HEAD_TO_CHANNEL(ch_)                                                    \
    do { /* send_head_of_queue_to_channel base.py:288-310 */                    \
        const double a_ = queue[qhead];                                         \
        qhead = (qhead + 1) % qcap;                                             \
        qsize--;                                                                \
        taked++;                                                                \
        const double w_ = 0.0 + (ttek - a_);                                    \
        neg_refresh(w_run, w_, taked);                                          \
        neg_add4(w_sum, w_);                                                    \
        arr[ch_] = a_;                                                          \
        wait[ch_] = w_;                                                         \
        tot[ch_] = mqo_src_next(services, &err);                                \
        end[ch_] = ttek + tot[ch_];                                             \
        is_free[ch_] = 0;                                                       \
        free_channels--;                                                        \
    } while (0)
    /* a job interrupted by a negative arrival goes back to the head of the line and, a server being free, restarts at
     * once (negative.py:262-275, 366-372 -> base.py:288-310 -> servers.py:57-88): one more `taked`, one more sample
     * of its unchanged wait, and a completion time that depends on the scenario */
#define RESTART(ch_, a_, w_, rem_, tot_)                                                \
    do {                                                                                \
        taked++;                                                                        \
        const double wt_ = (w_) + (ttek - ttek);                                        \
        neg_refresh(w_run, wt_, taked);                                                 \
        neg_add4(w_sum, wt_);                                                           \
        arr[ch_] = (a_);                                                                \
        wait[ch_] = wt_;                                                                \
        if (scenario == 2) { tot[ch_] = mqo_src_next(services, &err); end[ch_] = ttek + tot[ch_]; } \
        else if (scenario == 3) { tot[ch_] = (tot_); end[ch_] = ttek + (rem_); }        \
        else { tot[ch_] = (tot_); end[ch_] = ttek + (tot_); }                           \
        is_free[ch_] = 0;                                                               \
        free_channels--;                                                                \
    } while (0)

    while (served < total_served && !err) {
        int midx = -1;
        double mt = 1e16;
        for (int j = 0; j < c; ++j)
            if (end[j] < mt) { mt = end[j]; midx = j; }
        int ev = 0; /* 0 positive arrival, 1 negative arrival, 2 serving */
        double et = pos_time;
        if (neg_time < et) { ev = 1; et = neg_time; }
        if (midx >= 0 && mt < et) { ev = 2; et = mt; }

        if (ev == 0) {
            /* positive_arrival negative.py:225-241 */
            pos_arrived++;
            P_ADD(pos_time);
            in_sys++;
            ttek = pos_time;
            pos_time = ttek + mqo_src_next(positives, &err);
            if (free_channels == 0) {
                /* send_task_to_queue base.py:186-212 */
                if (buffer < 0 || qsize < buffer) {
                    if (qsize == qcap) {
                        double *nq = (double *)malloc(sizeof(double) * (size_t)qcap * 2);
                        for (int64_t i = 0; i < qsize; ++i) nq[i] = queue[(qhead + i) % qcap];
                        free(queue); queue = nq; qhead = 0; qcap *= 2;
                    }
                    queue[(qhead + qsize) % qcap] = ttek;
                    qsize++;
                } else {
                    dropped++;
                    in_sys--;
                }
            } else {
                /* send_task_to_channel base.py:155-184 */
                int j = 0;
                while (!is_free[j]) ++j;
                taked++;
                neg_refresh(w_run, 0.0, taked);
                neg_add4(w_sum, 0.0);
                arr[j] = ttek;
                wait[j] = 0.0;
                tot[j] = mqo_src_next(services, &err);
                end[j] = ttek + tot[j];
                is_free[j] = 0;
                free_channels--;
            }
        } else if (ev == 1) {
            /* negative_arrival negative.py:243-383 */
            neg_arrived++;
            P_ADD(neg_time);
            ttek = neg_time;
            neg_time = ttek + mqo_src_next(negatives, &err);
            if (in_sys == 0) continue;
            if (type == MQO_NEG_DISASTER && scenario != 1) {
                /* REQUEUE_ALL / RESUME_ALL / REQUEUE_ALL_NO_RESAMPLING negative.py:253-276 */
                int nb = 0, idx[64];
                double ja[64], jw[64], jrem[64], jtot[64];
                for (int j = 0; j < c; ++j) {
                    if (is_free[j]) continue;
                    const double rem = end[j] - ttek;
                    ja[nb] = arr[j]; jw[nb] = wait[j]; jrem[nb] = rem > 0.0 ? rem : 0.0; jtot[nb] = tot[j];
                    idx[nb] = nb;
                    nb++;
                    end[j] = 1e10;
                    is_free[j] = 1;
                }
                /* stable sort by arrival time (negative.py:269) */
                for (int a = 1; a < nb; ++a) {
                    const int t = idx[a];
                    int b = a - 1;
                    while (b >= 0 && ja[idx[b]] > ja[t]) { idx[b + 1] = idx[b]; --b; }
                    idx[b + 1] = t;
                }
                free_channels = c;
                for (int q = 0; q < nb; ++q) {
                    const int t = idx[q];
                    RESTART(q, ja[t], jw[t], jrem[t], jtot[t]);
                }
            } else if (type == MQO_NEG_DISASTER) {
                for (int j = 0; j < c; ++j) {
                    if (is_free[j]) continue;
                    const double a = arr[j];
                    end[j] = 1e10;
                    is_free[j] = 1;
                    BROKEN_SAMPLE(a);
                }
                in_sys = 0;
                free_channels = c;
                while (qsize > 0) {
                    const double a = queue[qhead];
                    qhead = (qhead + 1) % qcap;
                    qsize--;
                    const double wt = 0.0 + (ttek - a);
                    taked++;
                    total++;
                    broken++;
                    neg_refresh(w_run, wt, taked);
                    neg_add4(w_sum, wt);
                    const double sj = ttek - a;
                    neg_refresh(v_run, sj, total);
                    neg_add4(v_sum, sj);
                    neg_refresh(vb_run, sj, broken);
                    neg_add4(vb_sum, sj);
                }
            } else {
                int nb = 0, busy_idx[64];
                for (int j = 0; j < c; ++j)
                    if (!is_free[j]) busy_idx[nb++] = j;
                if (nb == 0) continue;
                const double u = mqo_src_next(choices, &err);
                int pick = (int)(u * (double)nb);
                if (pick >= nb) pick = nb - 1;
                const int j = busy_idx[pick];
                const double a = arr[j];
                const double rem0 = end[j] - ttek;
                end[j] = 1e10;
                is_free[j] = 1;
                free_channels++;
                if (scenario != 1) {
                    /* REQUEUE / RESUME / REQUEUE_NO_RESAMPLING negative.py:347-372 */
                    RESTART(j, a, wait[j], rem0 > 0.0 ? rem0 : 0.0, tot[j]);
                } else {
                    BROKEN_SAMPLE(a);
                    in_sys--;
                    if (qsize != 0) HEAD_TO_CHANNEL(j);
                }
            }
        } else {
            /* serving negative.py:425-450 */
            const double time_to_end = end[midx];
            const double a = arr[midx];
            end[midx] = 1e10;
            is_free[midx] = 1;
            P_ADD(time_to_end);
            ttek = time_to_end;
            free_channels++;
            served++;
            total++;
            const double sj = ttek - a;
            neg_refresh(v_run, sj, total);
            neg_add4(v_sum, sj);
            neg_refresh(vs_run, sj, served);
            neg_add4(vs_sum, sj);
            in_sys--;
            if (qsize != 0) HEAD_TO_CHANNEL(midx);
        }
    }
    free(queue);
    memcpy(out->w_run, w_run, sizeof(w_run));
    memcpy(out->v_run, v_run, sizeof(v_run));
    memcpy(out->vs_run, vs_run, sizeof(vs_run));
    memcpy(out->vb_run, vb_run, sizeof(vb_run));
    memcpy(out->w_sum, w_sum, sizeof(w_sum));
    memcpy(out->v_sum, v_sum, sizeof(v_sum));
    memcpy(out->vs_sum, vs_sum, sizeof(vs_sum));
    memcpy(out->vb_sum, vb_sum, sizeof(vb_sum));
    out->positive_arrived = pos_arrived;
    out->negative_arrived = neg_arrived;
    out->taked = taked;
    out->served = served;
    out->broken = broken;
    out->total = total;
    out->dropped = dropped;
    out->in_sys = in_sys;
    out->queued = qsize;
    out->ttek = ttek;
    out->p_over = p_over;
    return err ? -1 : 0;
}
