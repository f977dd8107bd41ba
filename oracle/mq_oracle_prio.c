/*
 * mq_oracle_prio.c -- CPU restatement of PriorityQueueSimulator (most_queue/sim/priority.py).
 * TEST INFRASTRUCTURE ONLY (see mq_oracle.h).
 *
 *   priority.py:622-657  run                 -> loop below, stop when sum(served) >= total
 *   priority.py:564-608  event selection     -> class 0 arrival < class 1 arrival < ... < serving on ties
 *   priority.py:282-332  arrival(k)
 *   priority.py:334-369  arrive_on_free_channels
 *   priority.py:390-455  arrive_with_absolute_prty (PR / RS / RW)
 *   priority.py:457-539  serving(c)
 *   sim/utils/servers.py:219-261  ServerPriority.start_service / end_service
 */
#include "mq_oracle.h"

#include <stdlib.h>
#include <string.h>

#define PRIO_FREE 1e10
#define PRIO_MAX_C 64
#define PRIO_MAX_K 16

double mqo_src_next(mqo_src *s, int *err)
{
    if (s->generate) {
        double x = mqo_variate_sid(&s->dist, s->seed, s->sid, s->rep, (uint32_t)s->pos, s->which);
        s->pos++;
        return x;
    }
    if (s->pos >= s->n) {
        *err = -1;
        return 0.0;
    }
    double x = s->arr[s->pos * (s->stride > 0 ? s->stride : 1)];
    s->pos++;
    return x;
}

static inline void refresh4(double *m, double x, int64_t count)
{
    double power = x;
    double inv_count = 1.0 / (double)count;
    double one_minus = 1.0 - inv_count;
    for (int i = 0; i < 4; ++i) {
        m[i] = m[i] * one_minus + power * inv_count;
        power *= x;
    }
}

static inline void add4(double *s, double x)
{
    double pw = x;
    for (int i = 0; i < 4; ++i) {
        s[i] += pw;
        pw *= x;
    }
}

typedef struct {
    int k;
    double arr_time, start_waiting_time, wait_time, time_to_end_service;
    int is_pr;
} pjob;

typedef struct {
    pjob *buf;
    int64_t cap, head, size;
} pq;

static int pq_push(pq *q, pjob j)
{
    if (q->size == q->cap) {
        int64_t ncap = q->cap ? q->cap * 2 : 64;
        pjob *nb = (pjob *)malloc((size_t)ncap * sizeof(pjob));
        if (!nb) return -1;
        for (int64_t i = 0; i < q->size; ++i) nb[i] = q->buf[(q->head + i) % q->cap];
        free(q->buf);
        q->buf = nb;
        q->cap = ncap;
        q->head = 0;
    }
    q->buf[(q->head + q->size) % q->cap] = j;
    q->size++;
    return 0;
}

static pjob pq_pop(pq *q)
{
    pjob j = q->buf[q->head];
    q->head = (q->head + 1) % q->cap;
    q->size--;
    return j;
}

int mqo_prio_run(int c, int K, int prty, int64_t buffer, mqo_src *arrivals, mqo_src *services,
                 int64_t total_served, double *p, int64_t p_bins, mqo_prio_class_out *out,
                 double *ttek_out)
{
    if (c < 1 || c > PRIO_MAX_C || K < 1 || K > PRIO_MAX_K) return -2;
    double end[PRIO_MAX_C], total_time[PRIO_MAX_C];
    pjob on[PRIO_MAX_C];
    int busy[PRIO_MAX_C];
    for (int j = 0; j < c; ++j) { end[j] = PRIO_FREE; total_time[j] = 0.0; busy[j] = 0; }
    pq queue[PRIO_MAX_K];
    memset(queue, 0, sizeof(queue));
    memset(out, 0, sizeof(*out) * (size_t)K);
    for (int64_t i = 0; i < (int64_t)K * p_bins; ++i) p[i] = 0.0;
    double arrival_time[PRIO_MAX_K];
    int err = 0, rc = 0;
    for (int k = 0; k < K; ++k) arrival_time[k] = mqo_src_next(&arrivals[k], &err); /* :130 */
    if (err) return -1;
    double ttek = 0.0;
    int free_channels = c;
    int64_t served_total = 0;

    while (served_total < total_served) {
        /* _get_min_server_time :549-562: starts at 1e10 with strict '<' */
        int midx = -1;
        double mt = 1e10;
        for (int j = 0; j < c; ++j)
            if (end[j] < mt) { mt = end[j]; midx = j; }
        /* events dict order: arrival_class_0.., serving; min keeps the first minimum */
        int ev = -1;
        double et = 0.0;
        for (int k = 0; k < K; ++k)
            if (ev < 0 || arrival_time[k] < et) { ev = k; et = arrival_time[k]; }
        if (midx >= 0 && mt < et) { ev = K; et = mt; }

        if (ev < K) {
            const int k = ev;
            mqo_prio_class_out *o = &out[k];
            ttek = arrival_time[k];
            double dt = arrival_time[k] - o->t_old;
            if (o->in_sys < p_bins) p[k * p_bins + o->in_sys] += dt; else o->p_over += dt;
            arrival_time[k] = ttek + mqo_src_next(&arrivals[k], &err);
            if (err) { rc = -1; break; }
            pjob nj = {k, ttek, -1.0, 0.0, 0.0, 0};
            o->arrived++;
            o->in_sys++;
            o->t_old = ttek;
            if (free_channels != 0) {
                /* arrive_on_free_channels :334-369 (no warm-up) */
                int j = 0;
                while (busy[j]) ++j;
                o->taked++;
                total_time[j] = mqo_src_next(&services[k], &err);
                if (err) { rc = -1; break; }
                end[j] = ttek + total_time[j];
                nj.time_to_end_service = end[j];
                on[j] = nj;
                busy[j] = 1;
                free_channels--;
                continue;
            }
            if (prty == MQO_NP) {
                nj.start_waiting_time = ttek; /* :331-332 */
                if (pq_push(&queue[k], nj)) { rc = -3; break; }
                continue;
            }
            /* arrive_with_absolute_prty :390-455 */
            int found = 0;
            for (int j = 0; j < c; ++j) {
                if (on[j].k > k) {
                    double time_to_end = end[j];
                    double tot = total_time[j];
                    pjob victim = on[j];
                    o->taked++;
                    victim.start_waiting_time = ttek;
                    victim.is_pr = 1;
                    if (prty == MQO_PR) {
                        victim.time_to_end_service = time_to_end - ttek;
                    } else if (prty == MQO_RS) {
                        victim.time_to_end_service = mqo_src_next(&services[k], &err); /* :425 */
                        if (err) { rc = -1; break; }
                    } else {
                        victim.time_to_end_service = tot;
                    }
                    if (pq_push(&queue[victim.k], victim)) { rc = -3; break; }
                    total_time[j] = mqo_src_next(&services[k], &err);
                    if (err) { rc = -1; break; }
                    end[j] = ttek + total_time[j];
                    nj.time_to_end_service = end[j];
                    on[j] = nj;
                    found = 1;
                    break;
                }
            }
            if (rc) break;
            if (!found) {
                int64_t qtot = 0;
                for (int kk = 0; kk < K; ++kk) qtot += queue[kk].size;
                if (buffer <= 0 || qtot < buffer) {
                    nj.start_waiting_time = ttek;
                    if (pq_push(&queue[k], nj)) { rc = -3; break; }
                } else {
                    o->dropped++;
                    o->in_sys--;
                }
            }
        } else {
            /* serving(c) :457-539 */
            const int j = midx;
            double time_to_end = end[j];
            pjob done = on[j];
            end[j] = PRIO_FREE;
            busy[j] = 0;
            total_time[j] = 0.0;
            const int k = done.k;
            mqo_prio_class_out *o = &out[k];
            double dt = time_to_end - o->t_old;
            if (o->in_sys < p_bins) p[k * p_bins + o->in_sys] += dt; else o->p_over += dt;
            ttek = time_to_end;
            o->t_old = ttek;
            o->served++;
            served_total++;
            free_channels++;
            double v = ttek - done.arr_time;
            refresh4(o->v_run, v, o->served);
            refresh4(o->w_run, done.wait_time, o->served);
            add4(o->v_sum, v);
            add4(o->w_sum, done.wait_time);
            o->in_sys--;
            int start = (prty == MQO_NP) ? 0 : k;
            for (int kk = start; kk < K; ++kk) {
                if (queue[kk].size != 0) {
                    pjob qj = pq_pop(&queue[kk]);
                    out[kk].taked++;
                    qj.wait_time += ttek - qj.start_waiting_time;
                    if (qj.is_pr) {
                        end[j] = ttek + qj.time_to_end_service; /* servers.py:237-239 */
                    } else {
                        total_time[j] = mqo_src_next(&services[kk], &err);
                        if (err) { rc = -1; break; }
                        end[j] = ttek + total_time[j];
                    }
                    qj.time_to_end_service = end[j];
                    on[j] = qj;
                    busy[j] = 1;
                    free_channels--;
                    break;
                }
            }
            if (rc) break;
        }
    }
    for (int k = 0; k < K; ++k) {
        out[k].queued = queue[k].size;
        free(queue[k].buf);
    }
    *ttek_out = ttek;
    return rc;
}
