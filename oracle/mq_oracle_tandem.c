/*
 * mq_oracle_tandem.c -- CPU restatement of TandemBlockingSim.run
 * (most_queue/sim/networks/tandem_blocking.py:50-142).  TEST INFRASTRUCTURE ONLY.
 */
#include "mq_oracle.h"

#include <math.h>
#include <string.h>

int mqo_tandem_run(int m, const int64_t *capacity, mqo_src *arrivals, mqo_src *services,
                   int64_t total_served, double warmup_fraction, mqo_tandem_out *out)
{
    if (m < 1 || m > MQO_NET_MAX_NODES) return -2;
    memset(out, 0, sizeof(*out));
    double t = 0.0, completion[MQO_NET_MAX_NODES], area[MQO_NET_MAX_NODES];
    int64_t jobs[MQO_NET_MAX_NODES];
    int blocked[MQO_NET_MAX_NODES];
    for (int i = 0; i < m; ++i) { completion[i] = INFINITY; area[i] = 0.0; jobs[i] = 0; blocked[i] = 0; }
    int err = 0;
    double next_arrival = mqo_src_next(arrivals, &err); /* :68 */
    if (err) return -1;
    const int64_t warm = (int64_t)((double)total_served * warmup_fraction); /* int(...) :70 */
    int64_t served = 0, arrived = 0, lost = 0, departures = 0;
    int stats_on = 0;
    double t0 = 0.0;

#define START_SERVICE(i) completion[i] = t + mqo_src_next(&services[i], &err)
#define ACCEPT(i)          \
    do {                   \
        jobs[i] += 1;      \
        if (jobs[i] == 1) START_SERVICE(i); \
    } while (0)
#define TRY_PUSH(i)                       \
    do {                                  \
        jobs[i] -= 1;                     \
        blocked[i] = 0;                   \
        completion[i] = INFINITY;         \
        if (jobs[i] > 0) START_SERVICE(i);\
        if ((i) + 1 < m) ACCEPT((i) + 1); \
    } while (0)

    while (served < total_served + warm) {
        double cmin = completion[0];
        int imin = 0;
        for (int i = 1; i < m; ++i)
            if (completion[i] < cmin) { cmin = completion[i]; imin = i; }
        double t_next = next_arrival < cmin ? next_arrival : cmin;
        if (stats_on) {
            double dt = t_next - t;
            for (int i = 0; i < m; ++i) area[i] += (double)jobs[i] * dt;
        }
        t = t_next;
        if (next_arrival <= cmin) {
            if (stats_on) arrived++;
            if (capacity[0] >= 0 && jobs[0] >= capacity[0]) {
                if (stats_on) lost++;
            } else {
                ACCEPT(0);
            }
            next_arrival = t + mqo_src_next(arrivals, &err);
            if (err) return -1;
            continue;
        }
        const int i = imin;
        if (i + 1 < m && capacity[i + 1] >= 0 && jobs[i + 1] >= capacity[i + 1]) {
            blocked[i] = 1;
            completion[i] = INFINITY;
        } else {
            if (i + 1 == m) {
                served++;
                if (stats_on) departures++;
                if (!stats_on && served >= warm) {
                    stats_on = 1;
                    t0 = t;
                    arrived = lost = departures = 0;
                    for (int j = 0; j < m; ++j) area[j] = 0.0;
                }
            }
            TRY_PUSH(i);
            int j = i;
            while (j > 0 && blocked[j - 1] && (capacity[j] < 0 || jobs[j] < capacity[j])) {
                TRY_PUSH(j - 1);
                j--;
            }
            if (err) return -1;
        }
    }
    const double elapsed = t - t0;
    out->t_end = t;
    out->t0 = t0;
    out->elapsed = elapsed;
    out->served = served;
    out->arrived = arrived;
    out->lost = lost;
    out->departures = departures;
    out->throughput = (double)departures / elapsed;
    out->loss_prob = arrived ? (double)lost / (double)arrived : 0.0;
    /* v = sum(mean_jobs) / throughput (:136): CPython >= 3.12 sums floats with Neumaier compensation
     * (Python/bltinmodule.c), which is what the reference ran with when the fixtures were generated */
    double sum = 0.0, comp = 0.0;
    for (int i = 0; i < m; ++i) {
        out->area[i] = area[i];
        out->mean_jobs[i] = area[i] / elapsed;
        const double x = out->mean_jobs[i], tt = sum + x;
        if (fabs(sum) >= fabs(x)) comp += (sum - tt) + x; else comp += (x - tt) + sum;
        sum = tt;
    }
    if (comp != 0.0 && isfinite(comp)) sum += comp;
    out->v0 = out->throughput > 0 ? sum / out->throughput : 0.0;
    return 0;
}
