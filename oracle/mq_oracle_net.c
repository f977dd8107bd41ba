/*
 * mq_oracle_net.c -- CPU restatement of NetworkSimulator (most_queue/sim/networks/network.py,
 * base_network_sim.py) with its QsSim nodes in network mode (sim/base.py arrival(moment, ts),
 * serving(c, is_network=True)).  TEST INFRASTRUCTURE ONLY (see mq_oracle.h).
 *
 *   network.py:196-229            run            -> loop until served == job_served
 *   base_network_sim.py:158-227   event scan     -> external arrival first, then busy servers in
 *                                                   node-major, channel-minor order; first minimum wins
 *   network.py:158-172            on_arrival     -> next gap drawn BEFORE the routing uniform
 *   network.py:174-194            on_serving     -> node.serving() (which may pop the queue head and
 *                                                   draw its service) BEFORE the routing uniform
 *   network.py:102-114            choose_next_node
 *   network.py:116-135            refresh_v_stat / refresh_w_stat (3 moments, math.pow)
 *   sim/base.py:214-247,249-310   node arrival / serving / send_head_of_queue_to_channel
 */
#include "mq_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NET_MAX_CH 64

typedef struct {
    double arr_network, wait_network, arr_time, start_waiting_time, wait_time;
} njob;

typedef struct {
    njob *buf;
    int64_t cap, head, size;
} nq;

static int nq_push(nq *q, njob j)
{
    if (q->size == q->cap) {
        int64_t ncap = q->cap ? q->cap * 2 : 64;
        njob *nb = (njob *)malloc((size_t)ncap * sizeof(njob));
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

static njob nq_pop(nq *q)
{
    njob j = q->buf[q->head];
    q->head = (q->head + 1) % q->cap;
    q->size--;
    return j;
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

static inline void addk(double *s, double x, int n)
{
    double pw = x;
    for (int i = 0; i < n; ++i) {
        s[i] += pw;
        pw *= x;
    }
}

/* network.py:116-135 */
static inline void refresh_net(double *m, double x, int64_t served)
{
    double factor = 1.0 - (1.0 / (double)served);
    for (int i = 0; i < 3; ++i) m[i] = m[i] * factor + pow(x, (double)(i + 1)) / (double)served;
}

/* network.py:102-114 */
static int choose_next(const double *R, int m, int current, double p)
{
    double sum_p = 0.0;
    for (int i = 0; i < m + 1; ++i) {
        sum_p += R[(current + 1) * (m + 1) + i];
        if (sum_p > p) return i;
    }
    return 0;
}

int mqo_net_run(int m, const int *channels, const double *R, mqo_src *arrivals, mqo_src *routing,
                mqo_src *services, int64_t job_served, mqo_net_out *out, mqo_net_node_out *nodes,
                double *v_samples, double *w_samples, int64_t cap)
{
    if (m < 1 || m > MQO_NET_MAX_NODES) return -2;
    double end[MQO_NET_MAX_NODES][NET_MAX_CH];
    njob on[MQO_NET_MAX_NODES][NET_MAX_CH];
    int busy[MQO_NET_MAX_NODES][NET_MAX_CH];
    int free_ch[MQO_NET_MAX_NODES];
    nq queue[MQO_NET_MAX_NODES];
    memset(queue, 0, sizeof(queue));
    memset(out, 0, sizeof(*out));
    memset(nodes, 0, sizeof(*nodes) * (size_t)m);
    for (int i = 0; i < m; ++i) {
        if (channels[i] < 1 || channels[i] > NET_MAX_CH) return -2;
        free_ch[i] = channels[i];
        for (int c = 0; c < channels[i]; ++c) { end[i][c] = 1e10; busy[i][c] = 0; }
    }
    int err = 0, rc = 0;
    double arrival_time = mqo_src_next(arrivals, &err); /* network.py:76 */
    if (err) return -1;
    double ttek = 0.0;

    /* node arrival: sim/base.py:214-247 with moment/ts given */
#define NODE_ARRIVAL(node, job)                                                                   \
    do {                                                                                          \
        mqo_net_node_out *no_ = &nodes[node];                                                     \
        no_->arrived++;                                                                           \
        (job).arr_time = ttek;                                                                    \
        (job).wait_time = 0.0;                                                                    \
        (job).start_waiting_time = -1.0;                                                          \
        no_->in_sys++;                                                                            \
        if (free_ch[node] == 0) {                                                                 \
            (job).start_waiting_time = ttek; /* base.py:199-201 */                                \
            if (nq_push(&queue[node], (job))) { rc = -3; }                                        \
        } else {                                                                                  \
            int c_ = 0;                                                                           \
            while (busy[node][c_]) ++c_;                                                          \
            no_->taked++;                                                                         \
            refresh4(no_->w_run, (job).wait_time, no_->taked);                                    \
            addk(no_->w_sum, (job).wait_time, 4);                                                 \
            end[node][c_] = ttek + mqo_src_next(&services[node], &err); /* servers.py:88 */       \
            on[node][c_] = (job);                                                                 \
            busy[node][c_] = 1;                                                                   \
            free_ch[node]--;                                                                      \
        }                                                                                         \
    } while (0)

    while (out->served < job_served) {
        /* base_network_sim.py:184-227 */
        int ev_node = -1, ev_ch = -1;
        double et = arrival_time;
        for (int i = 0; i < m; ++i)
            for (int c = 0; c < channels[i]; ++c)
                if (busy[i][c] && end[i][c] < et) { et = end[i][c]; ev_node = i; ev_ch = c; }

        if (ev_node < 0) {
            /* on_arrival network.py:158-172 */
            ttek = arrival_time;
            out->arrived++;
            out->in_sys++;
            arrival_time = ttek + mqo_src_next(arrivals, &err);
            int next = choose_next(R, m, -1, mqo_src_next(routing, &err));
            if (err) { rc = -1; break; }
            njob ts = {ttek, 0.0, 0.0, 0.0, 0.0};
            if (next >= m) { rc = -4; break; } /* source routed straight out: the reference would index qs[m] */
            NODE_ARRIVAL(next, ts);
            if (err) { rc = -1; break; }
            if (rc) break;
        } else {
            /* on_serving network.py:174-194 */
            const int node = ev_node, c = ev_ch;
            mqo_net_node_out *no = &nodes[node];
            ttek = end[node][c];
            /* node.serving(c, True): sim/base.py:249-286 */
            njob ts = on[node][c];
            end[node][c] = 1e10;
            busy[node][c] = 0;
            no->served++;
            free_ch[node]++;
            double vn = ttek - ts.arr_time;
            refresh4(no->v_run, vn, no->served);
            addk(no->v_sum, vn, 4);
            no->in_sys--;
            if (queue[node].size != 0) {
                /* send_head_of_queue_to_channel base.py:288-310 */
                njob qj = nq_pop(&queue[node]);
                no->taked++;
                qj.wait_time += ttek - qj.start_waiting_time;
                qj.wait_network += ttek - qj.start_waiting_time;
                refresh4(no->w_run, qj.wait_time, no->taked);
                addk(no->w_sum, qj.wait_time, 4);
                end[node][c] = ttek + mqo_src_next(&services[node], &err);
                if (err) { rc = -1; break; }
                on[node][c] = qj;
                busy[node][c] = 1;
                free_ch[node]--;
            }
            int next = choose_next(R, m, node, mqo_src_next(routing, &err));
            if (err) { rc = -1; break; }
            if (next == m) {
                out->served++;
                out->in_sys--;
                double v = ttek - ts.arr_network;
                refresh_net(out->v_run, v, out->served);
                refresh_net(out->w_run, ts.wait_network, out->served);
                addk(out->v_sum, v, 3);
                addk(out->w_sum, ts.wait_network, 3);
                if (v_samples && out->served <= cap) v_samples[out->served - 1] = v;
                if (w_samples && out->served <= cap) w_samples[out->served - 1] = ts.wait_network;
            } else {
                NODE_ARRIVAL(next, ts);
                if (err) { rc = -1; break; }
                if (rc) break;
            }
        }
    }
    out->ttek = ttek;
    for (int i = 0; i < m; ++i) {
        nodes[i].queued = queue[i].size;
        free(queue[i].buf);
    }
    return rc;
}
