/*
 * mq_oracle.c -- CPU restatement of most-queue's GI/G/c/r FIFO simulation loop.
 *
 * TEST INFRASTRUCTURE ONLY (see mq_oracle.h).  Plain C, fp64, no FMA contraction
 * (build with -ffp-contract=off) so that every add/mul rounds exactly like the
 * CPython floats of the reference.
 *
 * Restates (reference paths relative to most_queue/):
 *   sim/base.py:488-528   QsSim.run            -> fifo_loop()
 *   sim/base.py:420-454   event selection      -> "arrival wins ties", lowest server index
 *   sim/base.py:214-247   arrival()
 *   sim/base.py:186-212   send_task_to_queue() (finite buffer, drop)
 *   sim/base.py:155-184   send_task_to_channel()
 *   sim/base.py:249-310   serving(), send_head_of_queue_to_channel()
 *   sim/base.py:312-325   _get_min_server_time() (strict '<')
 *   sim/base.py:327-340   _update_state_probs()
 *   sim/utils/servers.py:53-118  start_service / end_service (1e10 sentinel)
 *   sim/utils/stats_update.py:6-22  refresh_moments_stat()
 *   random/distributions.py:407-429,749-761,807-820,996-1003  variate formulas
 */
#include "mq_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define MQO_FREE_SENTINEL 1e10 /* servers.py:35,114 */

const char *mqo_version(void) { return "mq_oracle 1"; }

/* ------------------------------------------------------------------------- */
/* counter-based generator spec (DESIGN.md "Generator streams")              */
/* ------------------------------------------------------------------------- */

#define PHILOX_M2 0xD256D193u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u
#define PHILOX_M4A 0xD2511F53u
#define PHILOX_M4B 0xCD9E8D57u
#define MQO_KEY_STRIDE 0xB5297A4Du

void mqo_philox2x32_10(uint32_t c0, uint32_t c1, uint32_t key, uint32_t out[2])
{
    for (int r = 0; r < 10; ++r) {
        if (r) key += PHILOX_W0;
        uint64_t prod = (uint64_t)PHILOX_M2 * c0;
        uint32_t hi = (uint32_t)(prod >> 32), lo = (uint32_t)prod;
        c0 = hi ^ key ^ c1;
        c1 = lo;
    }
    out[0] = c0;
    out[1] = c1;
}

void mqo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        if (r) { k0 += PHILOX_W0; k1 += PHILOX_W1; }
        uint64_t p0 = (uint64_t)PHILOX_M4A * c0, p1 = (uint64_t)PHILOX_M4B * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* word stream of one variate: word 0 comes from the primary call shared by the source
 * (which=0) and the server (which=1) of stream `sid`; further words come from
 * extension calls n = 1, 2, ... that belong to this variate alone. */
typedef struct {
    uint32_t seed, sid, rep, idx;
    int which;
    uint32_t w0;
} word_src;

static inline uint32_t ws_key(const word_src *s, uint32_t n)
{
    uint32_t slot = s->sid * 64u + (n ? 2u * n + (uint32_t)s->which : 0u);
    return s->seed + MQO_KEY_STRIDE * slot;
}

static inline void ws_ext(const word_src *s, uint32_t n, uint32_t out[2])
{
    mqo_philox2x32_10(s->idx, s->rep, ws_key(s, n), out);
}

static inline double u01(uint32_t w) { return ((double)w + 0.5) * 0x1p-32; }

/* "r < p1" on the raw word: T = floor(p1 * 2^32) */
static inline uint32_t branch_threshold(double p1)
{
    double t = floor(p1 * 0x1p32);
    return t >= 4294967295.0 ? 0xFFFFFFFFu : (t <= 0.0 ? 0u : (uint32_t)t);
}

static double variate_from(const mqo_dist *d, const word_src *s)
{
    const double *p = d->p;
    switch (d->kind) {
    case MQO_M: /* distributions.py:906,807-820 with r = 1 */
        return -log(u01(s->w0)) / p[0];
    case MQO_H2: { /* distributions.py:407-429; the branch word is reused as the
                      conditional uniform of the chosen phase */
        uint32_t T = branch_threshold(p[0]);
        if (s->w0 < T)
            return -log(((double)s->w0 + 0.5) / (double)T) / p[1];
        return -log(((double)(s->w0 - T) + 0.5) / (4294967296.0 - (double)T)) / p[2];
    }
    case MQO_E: { /* distributions.py:807-820: -log(prod U)/mu == -(sum log U)/mu */
        int r = (int)p[0];
        double acc = log(u01(s->w0));
        uint32_t e[2];
        for (int k = 1; k < r; ++k) {
            if (k & 1) ws_ext(s, (uint32_t)(k + 1) / 2, e);
            acc += log(u01(e[(k - 1) & 1]));
        }
        return -acc / p[1];
    }
    case MQO_PA: /* distributions.py:749-761: K * U^(-1/alpha) */
        return p[1] * exp(-log(u01(s->w0)) / p[0]);
    case MQO_D: /* distributions.py:614-626 */
        return p[0];
    case MQO_GAMMA: { /* distributions.py:996-1003 -> Generator.gamma(alpha, 1/mu);
                         restated as Marsaglia-Tsang (numpy's own algorithm family) */
        double mu = p[0], alpha = p[1];
        double a = alpha >= 1.0 ? alpha : alpha + 1.0;
        double dd = a - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * dd);
        double x = dd;
        for (uint32_t t = 0; t < 15; ++t) {
            uint32_t e0[2], e1[2];
            ws_ext(s, 2 * t + 1, e0);
            double z = sqrt(-2.0 * log(u01(e0[0]))) * cos(6.283185307179586 * u01(e0[1]));
            double v = 1.0 + cc * z;
            if (v <= 0.0) continue;
            v = v * v * v;
            x = dd * v;
            /* acceptance uniform: the primary word on the first attempt when alpha >= 1 (it is otherwise unused),
             * else word 0 of extension call 2t + 2 */
            uint32_t wa = s->w0;
            if (!(t == 0 && alpha >= 1.0)) { ws_ext(s, 2 * t + 2, e1); wa = e1[0]; }
            if (log(u01(wa)) < 0.5 * z * z + dd - dd * v + dd * log(v)) break;
        }
        if (alpha < 1.0) x *= exp(log(u01(s->w0)) / alpha);
        return x / mu;
    }
    case MQO_C2: { /* distributions.py:542-560 */
        uint32_t e[2];
        ws_ext(s, 1, e);
        double res = -log(u01(s->w0)) / p[1];
        if (e[0] < branch_threshold(p[0])) res += -log(u01(e[1])) / p[2];
        return res;
    }
    case MQO_UNIFORM: /* distributions.py:298-309 */
        return p[0] - p[1] + 2.0 * p[1] * u01(s->w0);
    case MQO_NORM: { /* distributions.py:209-216 (Box-Muller instead of ziggurat) */
        uint32_t e[2];
        ws_ext(s, 1, e);
        return p[0] + p[1] * sqrt(-2.0 * log(u01(s->w0))) * cos(6.283185307179586 * u01(e[0]));
    }
    case MQO_WEIBULL: /* distributions.py:94-102: pow(-log(p) * W, -1/k) as written there */
        return pow(-log(u01(s->w0)) * p[1], -1.0 / p[0]);
    default:
        return NAN;
    }
}

double mqo_variate_sid(const mqo_dist *d, uint32_t seed, uint32_t sid, uint32_t rep, uint32_t i, int which)
{
    word_src s = {seed, sid, rep, i, which, 0u};
    uint32_t w[2];
    mqo_philox2x32_10(i, rep, ws_key(&s, 0), w);
    s.w0 = w[which];
    return variate_from(d, &s);
}

double mqo_variate(const mqo_dist *d, uint32_t seed, uint32_t rep, uint32_t i, int which)
{
    word_src s = {seed, 0u, rep, i, which, 0u};
    uint32_t w[2];
    mqo_philox2x32_10(i, rep, ws_key(&s, 0), w);
    s.w0 = w[which];
    return variate_from(d, &s);
}

/* ------------------------------------------------------------------------- */
/* FIFO event loop                                                           */
/* ------------------------------------------------------------------------- */

/* stats_update.py:6-22 */
static inline void refresh_moments(double *m, int n, double x, int64_t count)
{
    double power = x;
    double inv_count = 1.0 / (double)count;
    double one_minus = 1.0 - inv_count;
    for (int i = 0; i < n; ++i) {
        m[i] = m[i] * one_minus + power * inv_count;
        power *= x;
    }
}

static inline void add_powers(double *s, double x)
{
    double pw = x;
    for (int i = 0; i < 4; ++i) {
        s[i] += pw;
        pw *= x;
    }
}

typedef struct {
    double arr;   /* arrival moment == start_waiting_time (base.py:196-197) */
    uint32_t idx; /* arrival index, generator mode only */
    uint32_t w1;  /* its service word (primary call word 1), generator mode only */
} qjob;

typedef struct {
    qjob *buf;
    int64_t cap, head, size;
} fifo_q;

static int q_push(fifo_q *q, qjob j)
{
    if (q->size == q->cap) {
        int64_t ncap = q->cap ? q->cap * 2 : 64;
        qjob *nb = (qjob *)malloc((size_t)ncap * sizeof(qjob));
        if (!nb) return -1;
        for (int64_t k = 0; k < q->size; ++k) nb[k] = q->buf[(q->head + k) % q->cap];
        free(q->buf);
        q->buf = nb;
        q->cap = ncap;
        q->head = 0;
    }
    q->buf[(q->head + q->size) % q->cap] = j;
    q->size++;
    return 0;
}

static qjob q_pop(fifo_q *q)
{
    qjob j = q->buf[q->head];
    q->head = (q->head + 1) % q->cap;
    q->size--;
    return j;
}

typedef struct {
    /* replay */
    const double *A, *S;
    int64_t nA, nS, strideA, strideS;
    /* generate */
    const mqo_dist *src, *srv;
    uint32_t seed, rep;
    int generate;
    int64_t warm; /* additive: statistics restart at the warm-th start of service (0 = reference) */
} fifo_input;

#define MAX_C 64

static int fifo_loop(int c, int64_t buffer, const fifo_input *in, int64_t total_served,
                     double *p, int64_t p_bins, mqo_fifo_out *o,
                     double *w_samples, int64_t w_cap, double *v_samples, int64_t v_cap)
{
    if (c < 1 || c > MAX_C) return -2;
    double end[MAX_C], arr[MAX_C];
    int busy[MAX_C], post[MAX_C];
    const int64_t warm = in->warm;
    double t_warm = 0.0;
    int64_t v_n = 0;
    for (int j = 0; j < c; ++j) { end[j] = MQO_FREE_SENTINEL; arr[j] = 0.0; busy[j] = 0; post[j] = 0; }
#define WARM_RESET()                                                                  \
    do {                                                                              \
        if (warm > 0 && o->taked == warm) {                                           \
            for (int k_ = 0; k_ < 4; ++k_) o->w_sum[k_] = o->v_sum[k_] = o->w_run[k_] = o->v_run[k_] = 0.0; \
            for (int64_t k_ = 0; k_ < p_bins; ++k_) p[k_] = 0.0;                      \
            o->p_over = 0.0;                                                          \
            t_warm = ttek;                                                            \
        }                                                                             \
    } while (0)
    memset(o, 0, sizeof(*o));
    for (int64_t k = 0; k < p_bins; ++k) p[k] = 0.0;
    fifo_q q = {0, 0, 0, 0};

    double ttek = 0.0, start_busy = 0.0;
    int free_channels = c;
    int64_t ia = 0, is = 0; /* variates consumed */
    int rc = 0;

    /* set_sources draws the first arrival (base.py:112) */
    double arrival_time;
    uint32_t pw[2] = {0, 0}; /* primary words of the most recent arrival index */
    uint32_t next_idx = 0;   /* index of the arrival whose moment is arrival_time */
    if (in->generate) {
        word_src s = {in->seed, 0u, in->rep, 0u, 0, 0u};
        mqo_philox2x32_10(0u, in->rep, ws_key(&s, 0), pw);
        s.w0 = pw[0];
        arrival_time = variate_from(in->src, &s);
        ia = 1;
    } else {
        if (in->nA < 1) return -1;
        arrival_time = in->A[0];
        ia = 1;
    }

    while (o->served < total_served) {
        /* _get_min_server_time (base.py:312-325): strict '<' -> lowest index wins */
        int midx = -1;
        double mt = 1e16;
        for (int j = 0; j < c; ++j)
            if (end[j] < mt) { mt = end[j]; midx = j; }

        /* _select_next_event (base.py:442-454): "arrival" is inserted first, min() keeps
         * the first minimum -> arrival wins ties */
        if (arrival_time <= mt) {
            /* arrival() base.py:214-247 */
            o->arrived++;
            double dt = arrival_time - ttek;
            if (o->in_sys < p_bins) p[o->in_sys] += dt; else o->p_over += dt;
            ttek = arrival_time;
            uint32_t this_idx = next_idx;
            uint32_t this_w1 = pw[1];
            if (in->generate) {
                next_idx++;
                word_src s = {in->seed, 0u, in->rep, next_idx, 0, 0u};
                mqo_philox2x32_10(next_idx, in->rep, ws_key(&s, 0), pw);
                s.w0 = pw[0];
                arrival_time = ttek + variate_from(in->src, &s);
                ia++;
            } else {
                if (ia >= in->nA) { rc = -1; break; }
                arrival_time = ttek + in->A[ia * in->strideA];
                ia++;
            }
            o->in_sys++;
            if (free_channels == 0) {
                /* send_task_to_queue base.py:186-212 */
                if (buffer < 0 || q.size < buffer) {
                    qjob j = {ttek, this_idx, this_w1};
                    if (q_push(&q, j)) { rc = -3; break; }
                } else {
                    o->dropped++;
                    o->in_sys--;
                }
            } else {
                /* send_task_to_channel base.py:155-184; free server = lowest index
                 * (the reference takes next(iter(set)); servers are exchangeable) */
                int j = 0;
                while (busy[j]) ++j;
                o->taked++;
                if (o->taked > warm) refresh_moments(o->w_run, 4, 0.0, o->taked - warm);
                add_powers(o->w_sum, 0.0);
                if (w_samples && o->taked <= w_cap) w_samples[o->taked - 1] = 0.0;
                post[j] = o->taked > warm;
                double sv;
                if (in->generate) {
                    word_src s = {in->seed, 0u, in->rep, this_idx, 1, this_w1};
                    sv = variate_from(in->srv, &s);
                    is++;
                } else {
                    if (is >= in->nS) { rc = -1; break; }
                    sv = in->S[is * in->strideS];
                    is++;
                }
                end[j] = ttek + sv; /* servers.py:88 */
                arr[j] = ttek;
                busy[j] = 1;
                free_channels--;
                if (free_channels == 0 && o->in_sys == c) start_busy = ttek;
                WARM_RESET();
            }
        } else {
            /* serving(c) base.py:249-286 */
            double time_to_end = end[midx];
            end[midx] = MQO_FREE_SENTINEL;
            busy[midx] = 0;
            double dt = time_to_end - ttek;
            if (o->in_sys < p_bins) p[o->in_sys] += dt; else o->p_over += dt;
            ttek = time_to_end;
            o->served++;
            free_channels++;
            double v = ttek - arr[midx];
            if (warm == 0 || post[midx]) {
                v_n++;
                refresh_moments(o->v_run, 4, v, v_n);
                add_powers(o->v_sum, v);
            }
            if (v_samples && o->served <= v_cap) v_samples[o->served - 1] = v;
            o->in_sys--;
            if (q.size == 0 && free_channels == 1 && o->in_sys == c - 1) {
                o->busy_n++;
                refresh_moments(o->busy_run, 3, ttek - start_busy, o->busy_n);
            }
            if (q.size != 0) {
                /* send_head_of_queue_to_channel base.py:288-310 */
                qjob h = q_pop(&q);
                if (free_channels == 1) start_busy = ttek;
                o->taked++;
                double w = 0.0 + (ttek - h.arr);
                if (o->taked > warm) refresh_moments(o->w_run, 4, w, o->taked - warm);
                add_powers(o->w_sum, w);
                if (w_samples && o->taked <= w_cap) w_samples[o->taked - 1] = w;
                post[midx] = o->taked > warm;
                double sv;
                if (in->generate) {
                    word_src s = {in->seed, 0u, in->rep, h.idx, 1, h.w1};
                    sv = variate_from(in->srv, &s);
                    is++;
                } else {
                    if (is >= in->nS) { rc = -1; break; }
                    sv = in->S[is * in->strideS];
                    is++;
                }
                end[midx] = ttek + sv;
                arr[midx] = h.arr;
                busy[midx] = 1;
                free_channels--;
                WARM_RESET();
            }
        }
    }
    o->ttek = ttek;
    o->v_count = warm ? v_n : o->served;
    o->w_count = o->taked - warm;
    o->t_obs = ttek - t_warm;
    o->queued = q.size;
    o->a_used = ia;
    o->s_used = is;
    free(q.buf);
    return rc;
}

int mqo_fifo_replay(int c, int64_t buffer, const double *A, int64_t nA, int64_t strideA,
                    const double *S, int64_t nS, int64_t strideS, int64_t total_served,
                    double *p, int64_t p_bins, mqo_fifo_out *out,
                    double *w_samples, int64_t w_cap, double *v_samples, int64_t v_cap)
{
    fifo_input in;
    memset(&in, 0, sizeof(in));
    in.A = A; in.S = S; in.nA = nA; in.nS = nS;
    in.strideA = strideA > 0 ? strideA : 1;
    in.strideS = strideS > 0 ? strideS : 1;
    return fifo_loop(c, buffer, &in, total_served, p, p_bins, out,
                     w_samples, w_cap, v_samples, v_cap);
}

/* Job-indexed form of the same simulator for an infinite buffer (the checker of csrc/k1b_fifo_jobs.cu): jobs 0..n-1 in
 * arrival order.  With FIFO and an unbounded line a job takes the server that frees first, so
 *   t_n = t_{n-1} + A[n]            (sim/base.py:230-236)      start = max(t_n, min completion)
 *   w_n = start - t_n               (base.py:300-303)          d = start + S[n]   (sim/utils/servers.py:88-90)
 *   v_n = d - t_n                   (base.py:269-272)
 * are the operations QsSim performs for job n on the same operands: w[n] / v[n] equal its samples bit for bit
 * (tests/test_oracle_fifo_jobs.py checks that against the reference's own fixtures).  Sums in arrival order with the
 * kernel's association: s1 += x, s2 += x*x, s3 = fma(x*x, x, s3), s4 = fma(x*x, x*x, s4). */
int mqo_fifo_jobs(int c, const double *A, int64_t strideA, const double *S, int64_t strideS, int64_t n,
                  double *w, double *v, double *w_sum, double *v_sum, double *t_last, double *d_max, int64_t *waited)
{
    if (c < 1 || c > 64 || n < 0) return -1;
    double L[64];
    for (int j = 0; j < c; ++j) L[j] = 0.0;
    double t = 0.0, dm = 0.0;
    int64_t nw = 0;
    for (int k = 0; k < 4; ++k) { w_sum[k] = 0.0; v_sum[k] = 0.0; }
    for (int64_t i = 0; i < n; ++i) {
        t = t + A[i * strideA];
        int jm = 0;
        for (int j = 1; j < c; ++j) if (L[j] < L[jm]) jm = j;
        const double start = L[jm] > t ? L[jm] : t;
        const double wi = start - t;
        const double d = start + S[i * strideS];
        const double vi = d - t;
        L[jm] = d;
        if (d > dm) dm = d;
        if (wi > 0.0) nw++;
        if (w) w[i] = wi;
        if (v) v[i] = vi;
        const double w2 = wi * wi, v2 = vi * vi;
        w_sum[0] = w_sum[0] + wi;
        w_sum[1] = w_sum[1] + w2;
        w_sum[2] = fma(w2, wi, w_sum[2]);
        w_sum[3] = fma(w2, w2, w_sum[3]);
        v_sum[0] = v_sum[0] + vi;
        v_sum[1] = v_sum[1] + v2;
        v_sum[2] = fma(v2, vi, v_sum[2]);
        v_sum[3] = fma(v2, v2, v_sum[3]);
    }
    *t_last = t;
    *d_max = dm;
    *waited = nw;
    return 0;
}

typedef struct {
    int c;
    int64_t warm;
    int64_t buffer, total_served, rep0, reps, p_bins;
    const mqo_dist *src, *srv;
    uint32_t seed;
    double *p;
    mqo_fifo_out *out;
    int64_t next; /* work counter, claimed with an atomic add */
    int rc;
} gen_job;

static void *gen_worker(void *arg)
{
    gen_job *g = (gen_job *)arg;
    for (;;) {
        int64_t r = __atomic_fetch_add(&g->next, 1, __ATOMIC_RELAXED);
        if (r >= g->reps) break;
        fifo_input in;
        memset(&in, 0, sizeof(in));
        in.generate = 1;
        in.warm = g->warm;
        in.src = g->src; in.srv = g->srv; in.seed = g->seed;
        in.rep = (uint32_t)(g->rep0 + r);
        int rc = fifo_loop(g->c, g->buffer, &in, g->total_served, g->p + r * g->p_bins,
                           g->p_bins, g->out + r, NULL, 0, NULL, 0);
        if (rc) __atomic_store_n(&g->rc, rc, __ATOMIC_RELAXED);
    }
    return NULL;
}

int mqo_fifo_generate(int c, int64_t buffer, const mqo_dist *src, const mqo_dist *srv,
                      int64_t total_served, uint32_t seed, int64_t rep0, int64_t reps,
                      double *p, int64_t p_bins, mqo_fifo_out *out, int threads)
{
    return mqo_fifo_generate_warm(c, buffer, src, srv, total_served, 0, seed, rep0, reps, p, p_bins, out, threads);
}

int mqo_fifo_generate_warm(int c, int64_t buffer, const mqo_dist *src, const mqo_dist *srv,
                           int64_t total_served, int64_t warm, uint32_t seed, int64_t rep0, int64_t reps,
                           double *p, int64_t p_bins, mqo_fifo_out *out, int threads)
{
    gen_job g = {c, warm, buffer, total_served, rep0, reps, p_bins, src, srv, seed, p, out, 0, 0};
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if (threads > reps) threads = (int)(reps > 0 ? reps : 1);
    pthread_t th[256];
    int started = 0;
    for (int t = 1; t < threads; ++t)
        if (pthread_create(&th[started], NULL, gen_worker, &g) == 0) started++;
    gen_worker(&g);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    return g.rc;
}
