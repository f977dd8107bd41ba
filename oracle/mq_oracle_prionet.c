/*
 * mq_oracle_prionet.c -- CPU restatement of PriorityNetwork (most_queue/sim/networks/priority_network.py)
 * with its PriorityQueueSimulator nodes in network mode (sim/priority.py arrival(k, moment, ts),
 * serving(c, is_network=True)).  TEST INFRASTRUCTURE ONLY (see mq_oracle.h).
 *
 *   priority_network.py:241-276  run            -> until sum(served) >= job_served
 *   base_network_sim.py:184-227  event scan     -> class 0 arrival < class 1 < ... < busy servers, node-major
 *   priority_network.py:196-215  on_arrival(k)  -> next gap drawn BEFORE the routing uniform
 *   priority_network.py:217-240  on_serving     -> node.serving() (queue pop + its service draw) BEFORE routing
 *   priority_network.py:137-150  choose_next_node(real_class, current_node)
 *   priority_network.py:152-173  refresh_v_stat / refresh_w_stat (3 moments, math.pow)
 *   sim/priority.py:282-332,390-455,457-539  node arrival / pre-emption / serving
 */
#include "mq_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PN_MAX_CH 32

typedef struct {
    int k;          /* real class (TaskPriority.k) */
    int node_class; /* in_node_class_num */
    double arr_network, wait_network, arr_time, start_waiting_time, wait_time, time_to_end_service;
    int is_pr;
} pnjob;

typedef struct {
    pnjob *buf;
    int64_t cap, head, size;
} pnq;

static int pnq_push(pnq *q, pnjob j)
{
    if (q->size == q->cap) {
        int64_t ncap = q->cap ? q->cap * 2 : 32;
        pnjob *nb = (pnjob *)malloc((size_t)ncap * sizeof(pnjob));
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

static pnjob pnq_pop(pnq *q)
{
    pnjob j = q->buf[q->head];
    q->head = (q->head + 1) % q->cap;
    q->size--;
    return j;
}

static inline void refresh4(double *m, double x, int64_t count)
{
    double power = x, inv_count = 1.0 / (double)count, one_minus = 1.0 - inv_count;
    for (int i = 0; i < 4; ++i) {
        m[i] = m[i] * one_minus + power * inv_count;
        power *= x;
    }
}

static inline void addk(double *s, double x, int n)
{
    double pw = x;
    for (int i = 0; i < n; ++i) {
        s[i] += pw;
        pw *= x;
    }
}

static inline void refresh_net(double *m, double x, int64_t served)
{
    double factor = 1.0 - (1.0 / (double)served);
    for (int i = 0; i < 3; ++i) m[i] = m[i] * factor + pow(x, (double)(i + 1)) / (double)served;
}

typedef struct {
    int c, prty, free_channels;
    double end[PN_MAX_CH], total_time[PN_MAX_CH];
    pnjob on[PN_MAX_CH];
    int busy[PN_MAX_CH];
    pnq queue[MQO_PN_MAX_CLASSES];
} pnode;

int mqo_prionet_run(int m, int K, const int *channels, const int *prty, const int *nodes_prty, const double *R,
                    mqo_src *arrivals, mqo_src *routing, mqo_src *services, int64_t job_served,
                    mqo_pn_class_out *oc, mqo_pn_node_class_out *on_, double *ttek_out)
{
    if (m < 1 || m > MQO_NET_MAX_NODES || K < 1 || K > MQO_PN_MAX_CLASSES) return -2;
    pnode *nd = (pnode *)calloc((size_t)m, sizeof(pnode));
    if (!nd) return -3;
    memset(oc, 0, sizeof(*oc) * (size_t)K);
    memset(on_, 0, sizeof(*on_) * (size_t)m * K);
    for (int i = 0; i < m; ++i) {
        if (channels[i] < 1 || channels[i] > PN_MAX_CH) { free(nd); return -2; }
        nd[i].c = channels[i];
        nd[i].prty = prty[i];
        nd[i].free_channels = channels[i];
        for (int c = 0; c < channels[i]; ++c) { nd[i].end[c] = 1e10; nd[i].busy[c] = 0; }
    }
    int err = 0, rc = 0;
    double arrival_time[MQO_PN_MAX_CLASSES];
    for (int k = 0; k < K; ++k) arrival_time[k] = mqo_src_next(&arrivals[k], &err); /* :83 */
    if (err) { free(nd); return -1; }
    double ttek = 0.0;
    int64_t served_total = 0;
    const int RS = (m + 1) * (m + 1);

#define CHOOSE(real, cur, dest)                                                          \
    do {                                                                                 \
        double p_ = mqo_src_next(routing, &err), sum_ = 0.0;                             \
        (dest) = 0;                                                                      \
        for (int i_ = 0; i_ < m + 1; ++i_) {                                             \
            sum_ += R[(real)*RS + ((cur) + 1) * (m + 1) + i_];                           \
            if (sum_ > p_) { (dest) = i_; break; }                                       \
        }                                                                                \
    } while (0)

    /* node arrival in network mode, sim/priority.py:282-332 with (k, moment, ts) */
#define NODE_ARRIVAL(node, kcls, job)                                                                   \
    do {                                                                                                \
        pnode *N_ = &nd[node];                                                                          \
        mqo_pn_node_class_out *o_ = &on_[(node)*K + (kcls)];                                            \
        (job).node_class = (kcls);                                                                      \
        (job).arr_time = ttek;                                                                          \
        (job).wait_time = 0.0;                                                                          \
        (job).is_pr = 0;                                                                                \
        (job).start_waiting_time = -1.0;                                                                \
        (job).time_to_end_service = 0.0;                                                                \
        o_->arrived++;                                                                                  \
        o_->in_sys++;                                                                                   \
        if (N_->free_channels != 0) {                                                                   \
            int j_ = 0;                                                                                 \
            while (N_->busy[j_]) ++j_;                                                                  \
            o_->taked++;                                                                                \
            N_->total_time[j_] = mqo_src_next(&services[(node)*K + (kcls)], &err);                      \
            N_->end[j_] = ttek + N_->total_time[j_];                                                    \
            N_->on[j_] = (job);                                                                         \
            N_->busy[j_] = 1;                                                                           \
            N_->free_channels--;                                                                        \
        } else if (N_->prty == MQO_NP) {                                                                \
            (job).start_waiting_time = ttek;                                                            \
            if (pnq_push(&N_->queue[kcls], (job))) rc = -3;                                             \
        } else {                                                                                        \
            int found_ = 0;                                                                             \
            for (int j_ = 0; j_ < N_->c; ++j_) {                                                        \
                if (N_->on[j_].node_class > (kcls)) {                                                   \
                    double tte_ = N_->end[j_], tot_ = N_->total_time[j_];                               \
                    pnjob vic_ = N_->on[j_];                                                            \
                    o_->taked++;                                                                        \
                    vic_.start_waiting_time = ttek;                                                     \
                    vic_.is_pr = 1;                                                                     \
                    if (N_->prty == MQO_PR) vic_.time_to_end_service = tte_ - ttek;                     \
                    else if (N_->prty == MQO_RS)                                                        \
                        vic_.time_to_end_service = mqo_src_next(&services[(node)*K + (kcls)], &err);    \
                    else vic_.time_to_end_service = tot_;                                               \
                    if (pnq_push(&N_->queue[vic_.node_class], vic_)) rc = -3;                           \
                    N_->total_time[j_] = mqo_src_next(&services[(node)*K + (kcls)], &err);              \
                    N_->end[j_] = ttek + N_->total_time[j_];                                            \
                    N_->on[j_] = (job);                                                                 \
                    found_ = 1;                                                                         \
                    break;                                                                              \
                }                                                                                       \
            }                                                                                           \
            if (!found_) {                                                                              \
                (job).start_waiting_time = ttek; /* infinite buffer, priority.py:437-440 */             \
                if (pnq_push(&N_->queue[kcls], (job))) rc = -3;                                         \
            }                                                                                           \
        }                                                                                               \
    } while (0)

    while (served_total < job_served) {
        /* base_network_sim.py:184-227 */
        int ev_k = 0;
        double et = arrival_time[0];
        for (int k = 1; k < K; ++k)
            if (arrival_time[k] < et) { et = arrival_time[k]; ev_k = k; }
        int ev_node = -1, ev_ch = -1;
        for (int i = 0; i < m; ++i)
            for (int c = 0; c < nd[i].c; ++c)
                if (nd[i].busy[c] && nd[i].end[c] < et) { et = nd[i].end[c]; ev_node = i; ev_ch = c; }

        if (ev_node < 0) {
            /* on_arrival(k) :196-215 */
            const int k = ev_k;
            ttek = arrival_time[k];
            oc[k].arrived++;
            oc[k].in_sys++;
            arrival_time[k] = ttek + mqo_src_next(&arrivals[k], &err);
            int next;
            CHOOSE(k, -1, next);
            if (err) { rc = -1; break; }
            if (next >= m) { rc = -4; break; }
            pnjob ts;
            memset(&ts, 0, sizeof(ts));
            ts.k = k;
            ts.arr_network = ttek;
            ts.wait_network = 0.0;
            const int nc = nodes_prty[next * K + k];
            NODE_ARRIVAL(next, nc, ts);
            if (err) { rc = -1; break; }
            if (rc) break;
        } else {
            /* on_serving :217-240 ; node.serving(c, True) sim/priority.py:457-539 */
            pnode *N = &nd[ev_node];
            const int j = ev_ch;
            ttek = N->end[j];
            pnjob done = N->on[j];
            N->end[j] = 1e10;
            N->busy[j] = 0;
            N->total_time[j] = 0.0;
            const int kc = done.node_class;
            mqo_pn_node_class_out *o = &on_[ev_node * K + kc];
            o->served++;
            N->free_channels++;
            refresh4(o->v_run, ttek - done.arr_time, o->served);
            refresh4(o->w_run, done.wait_time, o->served);
            addk(o->v_sum, ttek - done.arr_time, 4);
            addk(o->w_sum, done.wait_time, 4);
            o->in_sys--;
            const int start = (N->prty == MQO_NP) ? 0 : kc;
            for (int kk = start; kk < K; ++kk) {
                if (N->queue[kk].size != 0) {
                    pnjob qj = pnq_pop(&N->queue[kk]);
                    on_[ev_node * K + kk].taked++;
                    qj.wait_time += ttek - qj.start_waiting_time;
                    qj.wait_network += ttek - qj.start_waiting_time;
                    if (qj.is_pr) {
                        N->end[j] = ttek + qj.time_to_end_service;
                    } else {
                        N->total_time[j] = mqo_src_next(&services[ev_node * K + kk], &err);
                        N->end[j] = ttek + N->total_time[j];
                    }
                    N->on[j] = qj;
                    N->busy[j] = 1;
                    N->free_channels--;
                    break;
                }
            }
            if (err) { rc = -1; break; }
            const int real = done.k;
            int next;
            CHOOSE(real, ev_node, next);
            if (err) { rc = -1; break; }
            if (next == m) {
                oc[real].served++;
                served_total++;
                oc[real].in_sys--;
                refresh_net(oc[real].v_run, ttek - done.arr_network, oc[real].served);
                refresh_net(oc[real].w_run, done.wait_network, oc[real].served);
                addk(oc[real].v_sum, ttek - done.arr_network, 3);
                addk(oc[real].w_sum, done.wait_network, 3);
            } else {
                pnjob ts = done;
                const int nc = nodes_prty[next * K + real];
                NODE_ARRIVAL(next, nc, ts);
                if (err) { rc = -1; break; }
                if (rc) break;
            }
        }
    }
    for (int i = 0; i < m; ++i)
        for (int k = 0; k < K; ++k) {
            on_[i * K + k].queued = nd[i].queue[k].size;
            free(nd[i].queue[k].buf);
        }
    free(nd);
    *ttek_out = ttek;
    return rc;
}
