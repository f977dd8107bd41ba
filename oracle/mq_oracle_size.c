/*
 * mq_oracle_size.c -- CPU restatement of SizeBasedQsSim.run (most_queue/sim/size_based.py:64-248): one server, job
 * sizes drawn at arrival, the waiting line ordered by a rank (min first, ties by arrival time, then by creation
 * order), optional pre-emption of the job in service by a strictly better arrival.  With the default perfect
 * predictor SPJF / PSPJF / SPRPT rank exactly like SJF / PSJF / SRPT.  TEST INFRASTRUCTURE ONLY.
 */
#include "mq_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double rank, arr, start_wait, wait, remaining, size, pred;
    int64_t id;
} sjob;

static inline void sz_refresh(double *m, int n, double x, int64_t count)
{
    double power = x;
    const double inv_count = 1.0 / (double)count;
    const double one_minus = 1.0 - inv_count;
    for (int i = 0; i < n; ++i) {
        m[i] = m[i] * one_minus + power * inv_count;
        power = power * x;
    }
}

static inline void sz_add4(double *s, double x)
{
    double pw = x;
    for (int i = 0; i < 4; ++i) { s[i] += pw; pw *= x; }
}

static inline int key_less(const sjob *a, const sjob *b)
{ /* queue_priority_size.py:52-60: (rank, arr_time, id) */
    if (a->rank != b->rank) return a->rank < b->rank;
    if (a->arr != b->arr) return a->arr < b->arr;
    return a->id < b->id;
}

static inline double sz_rank(int discipline, const sjob *j)
{ /* _rank_value size_based.py:128-153 */
    switch (discipline) {
    case MQO_SZ_FCFS: return j->arr;
    case MQO_SZ_SJF:
    case MQO_SZ_PSJF: return j->size;
    case MQO_SZ_SRPT: return j->remaining;
    case MQO_SZ_SPJF:
    case MQO_SZ_PSPJF: return j->pred;
    default: { /* SPRPT: predicted remaining work */
        const double served = j->size - j->remaining;
        const double pr = j->pred - served;
        return pr > 0.0 ? pr : 0.0;
    }
    }
}

int mqo_size_run(int discipline, int64_t buffer, mqo_src *arrivals, mqo_src *sizes, int pred_kind, double pred_sigma,
                 mqo_src *preds, int64_t total_served, double *p, int64_t p_bins, mqo_size_out *out)
{
    if (discipline < MQO_SZ_FCFS || discipline > MQO_SZ_SPRPT) return -2;
    if (pred_kind < 0 || pred_kind > 3 || (pred_kind != 0 && !preds)) return -2;
    memset(out, 0, sizeof(*out));
    for (int64_t i = 0; i < p_bins; ++i) p[i] = 0.0;
    const int preemptive = discipline == MQO_SZ_PSJF || discipline == MQO_SZ_SRPT || discipline == MQO_SZ_PSPJF ||
                           discipline == MQO_SZ_SPRPT;
    const int tracks_remaining = discipline == MQO_SZ_SRPT || discipline == MQO_SZ_SPRPT;
    const int fcfs = discipline == MQO_SZ_FCFS;
    int64_t qcap = 1024, qsize = 0;
    sjob *queue = (sjob *)malloc(sizeof(sjob) * (size_t)qcap);
    sjob cur;
    memset(&cur, 0, sizeof(cur));
    int busy = 0, err = 0;
    double end = 1e10, ttek = 0.0, start_busy = 0.0, p_over = 0.0;
    int64_t arrived = 0, taked = 0, served = 0, dropped = 0, in_sys = 0, busy_n = 0, next_id = 0, preemptions = 0;
    double w_run[4] = {0}, v_run[4] = {0}, w_sum[4] = {0}, v_sum[4] = {0}, busy_run[3] = {0}, sd_sum[2] = {0};
    double arrival_time = mqo_src_next(arrivals, &err);

#define P_ADD(t_new)                                              \
    do {                                                          \
        const double dt_ = (t_new) - ttek;                        \
        if (in_sys < p_bins) p[in_sys] += dt_; else p_over += dt_; \
    } while (0)
#define RANK_OF(j_) sz_rank(discipline, &(j_))
#define Q_PUSH(j_)                                                                \
    do {                                                                          \
        if (qsize == qcap) { qcap *= 2; queue = (sjob *)realloc(queue, sizeof(sjob) * (size_t)qcap); } \
        queue[qsize] = (j_);                                                      \
        queue[qsize].rank = RANK_OF(j_);                                          \
        qsize++;                                                                  \
    } while (0)
#define START(j_)                                                                              \
    do { /* Server.start_service servers.py:57-88: remaining work if known, else a draw */     \
        cur = (j_);                                                                            \
        end = ttek + (fcfs ? mqo_src_next(sizes, &err) : cur.remaining);                       \
        busy = 1;                                                                              \
    } while (0)

    while (served < total_served && !err) {
        if (arrival_time <= (busy ? end : 1e10)) {
            /* arrival() size_based.py:216-247 */
            arrived++;
            P_ADD(arrival_time);
            ttek = arrival_time;
            arrival_time = ttek + mqo_src_next(arrivals, &err);
            in_sys++;
            sjob nj;
            memset(&nj, 0, sizeof(nj));
            nj.arr = ttek;
            nj.id = next_id++;
            if (!fcfs) { /* _sample_size :155-164 */
                nj.size = mqo_src_next(sizes, &err);
                nj.remaining = nj.size;
                /* predictor (size_based.py:163, sim/utils/predictor.py): perfect, x * Exp(1), x * exp(sigma * Z), or a
                 * replayed value */
                if (pred_kind == 0) nj.pred = nj.size;
                else {
                    const double z = mqo_src_next(preds, &err);
                    nj.pred = pred_kind == 1 ? nj.size * z : (pred_kind == 2 ? nj.size * exp(pred_sigma * z) : z);
                }
            }
            if (busy) {
                nj.start_wait = ttek;
                int handled = 0;
                if (preemptive) {
                    /* _try_preempt_on_arrival :182-214 */
                    if (tracks_remaining) {
                        const double rem = end - ttek;
                        cur.remaining = rem > 0.0 ? rem : 0.0; /* _update_in_service_rank_fields :166-180 */
                    }
                    sjob a = nj, b = cur;
                    a.rank = RANK_OF(nj);
                    b.rank = RANK_OF(cur);
                    if (key_less(&a, &b)) {
                        sjob old = cur;
                        const double rem = end - ttek; /* preempt_service(preserve_remaining=True) servers.py:90-106 */
                        old.remaining = rem > 0.0 ? rem : 0.0;
                        old.start_wait = ttek;
                        Q_PUSH(old);
                        taked++;
                        sz_refresh(w_run, 4, nj.wait, taked);
                        sz_add4(w_sum, nj.wait);
                        START(nj);
                        preemptions++;
                        handled = 1;
                    }
                }
                if (!handled) {
                    /* send_task_to_queue base.py:186-212 */
                    if (buffer < 0 || qsize < buffer) {
                        Q_PUSH(nj);
                    } else {
                        dropped++;
                        in_sys--;
                    }
                }
            } else {
                /* send_task_to_channel base.py:155-184 */
                taked++;
                sz_refresh(w_run, 4, 0.0, taked);
                sz_add4(w_sum, 0.0);
                START(nj);
                if (in_sys == 1) start_busy = ttek; /* free_channels == 0 and in_sys == n */
            }
        } else {
            /* serving() base.py:249-286 */
            P_ADD(end);
            ttek = end;
            end = 1e10;
            busy = 0;
            served++;
            const double v = ttek - cur.arr;
            sz_refresh(v_run, 4, v, served);
            sz_add4(v_sum, v);
            in_sys--;
            if (!fcfs && cur.size > 0.0) { /* _after_serving :110-119 */
                const double sd = v / cur.size;
                sd_sum[0] += sd;
                sd_sum[1] += sd * sd;
            }
            if (qsize == 0 && in_sys == 0) { /* free_channels == 1 and in_sys == n - 1 */
                busy_n++;
                sz_refresh(busy_run, 3, ttek - start_busy, busy_n);
            }
            if (qsize != 0) {
                /* pop the best rank (queue_priority_size.py:70-84), send_head_of_queue_to_channel base.py:288-310 */
                int64_t best = 0;
                for (int64_t i = 1; i < qsize; ++i)
                    if (key_less(&queue[i], &queue[best])) best = i;
                sjob qj = queue[best];
                queue[best] = queue[qsize - 1];
                qsize--;
                start_busy = ttek; /* free_channels == 1 */
                taked++;
                qj.wait = qj.wait + (ttek - qj.start_wait);
                sz_refresh(w_run, 4, qj.wait, taked);
                sz_add4(w_sum, qj.wait);
                START(qj);
            }
        }
    }
    free(queue);
    memcpy(out->w_run, w_run, sizeof(w_run));
    memcpy(out->v_run, v_run, sizeof(v_run));
    memcpy(out->w_sum, w_sum, sizeof(w_sum));
    memcpy(out->v_sum, v_sum, sizeof(v_sum));
    memcpy(out->busy_run, busy_run, sizeof(busy_run));
    memcpy(out->slowdown_sum, sd_sum, sizeof(sd_sum));
    out->busy_n = busy_n;
    out->arrived = arrived;
    out->taked = taked;
    out->served = served;
    out->dropped = dropped;
    out->in_sys = in_sys;
    out->queued = qsize;
    out->preemptions = preemptions;
    out->ttek = ttek;
    out->p_over = p_over;
    return err ? -1 : 0;
}
