/*
 * mq_oracle_scan.c -- CPU restatement of the GI/G/1 Lindley recursion that QsSim.run() reduces to
 * for one channel and an infinite buffer (most_queue/sim/base.py:488-528 with n = 1), sequentially
 * and with the fixed association tree of the scan kernel.  TEST INFRASTRUCTURE ONLY.
 */
#include "mq_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double a, b;
} mp; /* the map w -> max(a, w + b) */

static const mp MP_ID = {-INFINITY, 0.0};

/* later o earlier */
static inline mp mp_then(mp earlier, mp later)
{
    mp r;
    r.a = fmax(later.a, earlier.a + later.b);
    r.b = earlier.b + later.b;
    return r;
}

static inline mp mp_step(mp f, double x)
{
    mp r;
    r.a = fmax(0.0, f.a + x);
    r.b = f.b + x;
    return r;
}

static inline double mp_apply(mp f, double w) { return fmax(f.a, w + f.b); }

static void finish(const double *A, const double *S, int64_t N, const double *W, double w_in, mqo_lindley_out *o)
{
    memset(o, 0, sizeof(*o));
    for (int64_t n = 0; n < N; ++n) {
        double w = W[n], v = w + S[n];
        double pw = w, pv = v;
        for (int k = 0; k < 4; ++k) {
            o->w_sum[k] += pw;
            o->v_sum[k] += pv;
            pw *= w;
            pv *= v;
        }
        o->sum_a += A[n];
        o->sum_s += S[n];
    }
    o->served = N;
    o->taked = N;
    if (N > 0) {
        o->w_last_raw = W[N - 1] + (S[N - 1] - A[N]);
        if (o->w_last_raw >= 0.0) {
            /* job N arrived by the N-th departure: it is popped there and sampled (base.py:300-305) */
            double w = o->w_last_raw, pw = w;
            o->taked = N + 1;
            for (int k = 0; k < 4; ++k) {
                o->w_sum[k] += pw;
                pw *= w;
            }
        }
    }
    (void)w_in;
}

int mqo_lindley_seq(const double *A, const double *S, int64_t N, double w_in, double *W, mqo_lindley_out *out)
{
    double *buf = W ? W : (double *)calloc((size_t)(N > 0 ? N : 1), sizeof(double));
    if (!buf) return -3;
    double w = w_in;
    mp f = MP_ID;
    for (int64_t n = 0; n < N; ++n) {
        buf[n] = w;
        double x = S[n] - A[n + 1];
        w = fmax(0.0, w + x);
        f = mp_step(f, x);
    }
    finish(A, S, N, buf, w_in, out);
    out->a = f.a;
    out->b = f.b;
    if (!W) free(buf);
    return 0;
}

#define TILE_THREADS 256
#define WARP 32
#define TOP_THREADS 1024

/* inclusive Kogge-Stone over 32 lanes, as the kernel's __shfl_up loop does it */
static void ks32(mp *v)
{
    for (int d = 1; d < WARP; d <<= 1) {
        mp prev[WARP];
        memcpy(prev, v, sizeof(prev));
        for (int i = d; i < WARP; ++i) v[i] = mp_then(prev[i - d], prev[i]);
    }
}

/* exclusive prefixes of n summaries laid out as groups of 32 (Kogge-Stone) folded sequentially */
static mp scan_groups(const mp *in, int n, mp *excl)
{
    mp carry = MP_ID;
    for (int g = 0; g * WARP < n; ++g) {
        mp v[WARP];
        for (int i = 0; i < WARP; ++i) v[i] = (g * WARP + i < n) ? in[g * WARP + i] : MP_ID;
        mp inc[WARP];
        memcpy(inc, v, sizeof(inc));
        ks32(inc);
        for (int i = 0; i < WARP && g * WARP + i < n; ++i) {
            mp within = i ? inc[i - 1] : MP_ID;
            excl[g * WARP + i] = mp_then(carry, within);
        }
        carry = mp_then(carry, inc[WARP - 1]);
    }
    return carry;
}

int mqo_lindley_tree(const double *A, const double *S, int64_t N, double w_in, int L, double *W,
                     mqo_lindley_out *out)
{
    if (L < 1 || N < 0) return -2;
    const int64_t tile = (int64_t)TILE_THREADS * L;
    const int64_t ntiles = (N + tile - 1) / tile;
    double *buf = W ? W : (double *)calloc((size_t)(N > 0 ? N : 1), sizeof(double));
    mp *tsum = (mp *)malloc((size_t)(ntiles > 0 ? ntiles : 1) * sizeof(mp));
    mp *tex = (mp *)malloc((size_t)(ntiles > 0 ? ntiles : 1) * sizeof(mp));
    if (!buf || !tsum || !tex) return -3;

    /* pass 1: tile summaries */
    for (int64_t t = 0; t < ntiles; ++t) {
        mp seg[TILE_THREADS], ex[TILE_THREADS];
        for (int s = 0; s < TILE_THREADS; ++s) {
            mp f = MP_ID;
            for (int j = 0; j < L; ++j) {
                int64_t n = t * tile + (int64_t)s * L + j;
                if (n < N) f = mp_step(f, S[n] - A[n + 1]);
            }
            seg[s] = f;
        }
        tsum[t] = scan_groups(seg, TILE_THREADS, ex);
    }
    /* tile level: TOP_THREADS chunks folded sequentially, then groups of 32 */
    {
        const int64_t chunk = (ntiles + TOP_THREADS - 1) / TOP_THREADS;
        mp csum[TOP_THREADS], cex[TOP_THREADS];
        for (int c = 0; c < TOP_THREADS; ++c) {
            mp f = MP_ID;
            for (int64_t t = c * chunk; t < (c + 1) * chunk && t < ntiles; ++t) f = mp_then(f, tsum[t]);
            csum[c] = f;
        }
        mp total = scan_groups(csum, TOP_THREADS, cex);
        for (int c = 0; c < TOP_THREADS; ++c) {
            mp f = cex[c];
            for (int64_t t = c * chunk; t < (c + 1) * chunk && t < ntiles; ++t) {
                tex[t] = f;
                f = mp_then(f, tsum[t]);
            }
        }
        out->a = total.a;
        out->b = total.b;
    }
    /* pass 2: walk every segment from its carry */
    for (int64_t t = 0; t < ntiles; ++t) {
        mp seg[TILE_THREADS], ex[TILE_THREADS];
        for (int s = 0; s < TILE_THREADS; ++s) {
            mp f = MP_ID;
            for (int j = 0; j < L; ++j) {
                int64_t n = t * tile + (int64_t)s * L + j;
                if (n < N) f = mp_step(f, S[n] - A[n + 1]);
            }
            seg[s] = f;
        }
        scan_groups(seg, TILE_THREADS, ex);
        const double w_tile = mp_apply(tex[t], w_in);
        for (int s = 0; s < TILE_THREADS; ++s) {
            double w = mp_apply(ex[s], w_tile);
            for (int j = 0; j < L; ++j) {
                int64_t n = t * tile + (int64_t)s * L + j;
                if (n >= N) break;
                buf[n] = w;
                w = fmax(0.0, w + (S[n] - A[n + 1]));
            }
        }
    }
    double sa = out->a, sb = out->b;
    finish(A, S, N, buf, w_in, out);
    out->a = sa;
    out->b = sb;
    if (!W) free(buf);
    free(tsum);
    free(tex);
    return 0;
}
