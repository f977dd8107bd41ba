/*
 * mq_oracle.h -- CPU restatement of most-queue's simulation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or as
 * the timed CPU baseline.  The product path (most_queue_b200/) never links,
 * imports or calls it.
 *
 * Parity pin: the replay entry points are checked bit-for-bit against the
 * reference's own QsSim / PriorityQueueSimulator / NetworkSimulator (run in the
 * build container with pre-generated arrays injected through .generate())
 * via the fixtures in tests/golden/ (generator: tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it restates (paths relative to
 * the reference checkout, most_queue/...).
 */
#ifndef MQ_ORACLE_H
#define MQ_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* distribution kinds (most_queue/random/utils/create.py:19-80) */
enum {
    MQO_M = 0,      /* exponential, params[0] = mu                         */
    MQO_H2 = 1,     /* hyper-exponential, params = {p1, mu1, mu2}          */
    MQO_E = 2,      /* Erlang, params = {r, mu}                            */
    MQO_GAMMA = 3,  /* Gamma, params = {mu, alpha}                         */
    MQO_PA = 4,     /* Pareto, params = {alpha, K}                         */
    MQO_D = 5,      /* deterministic, params[0] = b                        */
    MQO_C2 = 6,     /* Cox-2, params = {p1, mu1, mu2}                      */
    MQO_UNIFORM = 7,/* uniform, params = {mean, half_interval}             */
    MQO_NORM = 8,   /* normal, params = {mean, std_dev}                    */
    MQO_WEIBULL = 9 /* Weibull (reference formula), params = {k, W}        */
};

typedef struct {
    int kind;
    double p[4];
} mqo_dist;

/* results of one FIFO replication */
typedef struct {
    double w_sum[4];   /* sum of w^k, k=1..4, accumulated in the reference's sample order */
    double v_sum[4];   /* sum of v^k, same                                                 */
    double w_run[4];   /* the reference's running-mean moments (stats_update.py:6-22)     */
    double v_run[4];
    double busy_run[3];/* busy-period running moments (base.py:276-281)                   */
    int64_t busy_n;
    int64_t arrived, taked, served, dropped, in_sys, queued;
    double ttek;       /* clock at stop                                                   */
    double p_over;     /* time spent in states >= p_bins                                  */
    int64_t a_used, s_used; /* how many interarrival / service variates were consumed     */
    int64_t w_count, v_count; /* sample counts behind w_* and v_* (differ from taked / served */
    double t_obs;             /* and ttek only when a warm-up is requested)                   */
} mqo_fifo_out;

/* QsSim.run() with pre-generated variates (sim/base.py:488-528).
 * A[0] is the draw made by set_sources (base.py:112); S[j] is the j-th service draw
 * (servers.py:88).  buffer < 0 means infinite (base.py:204).  p has p_bins entries
 * (time-in-state sums, NOT normalised).  w_samples / v_samples may be NULL; when given
 * they receive up to cap samples in the reference's sample order.
 * Returns 0, or -1 when the arrays ran out before total_served departures. */
/* job-indexed replay, infinite buffer: per-job waits / sojourn times (may be NULL) and power sums in arrival order */
int mqo_fifo_jobs(int c, const double *A, int64_t strideA, const double *S, int64_t strideS, int64_t n,
                  double *w, double *v, double *w_sum, double *v_sum, double *t_last, double *d_max, int64_t *waited);

int mqo_fifo_replay(int c, int64_t buffer, const double *A, int64_t nA, int64_t strideA,
                    const double *S, int64_t nS, int64_t strideS, int64_t total_served,
                    double *p, int64_t p_bins, mqo_fifo_out *out,
                    double *w_samples, int64_t w_cap, double *v_samples, int64_t v_cap);

/* Same event loop driven by the counter-based generator spec of DESIGN.md
 * (Philox2x32-10 keyed by seed, counter = (arrival index, replication id)),
 * evaluated in fp64 with libm.  One call simulates replications [rep0, rep0+reps)
 * with `threads` OpenMP threads; outputs are arrays of `reps` records, p is
 * [reps][p_bins]. */
int mqo_fifo_generate(int c, int64_t buffer, const mqo_dist *src, const mqo_dist *srv,
                      int64_t total_served, uint32_t seed, int64_t rep0, int64_t reps,
                      double *p, int64_t p_bins, mqo_fifo_out *out, int threads);

/* additive warm-up (not in the reference): statistics restart at the instant of the warm-th start of
 * service; w / v then cover the jobs started after it, p and t_obs the time after it */
int mqo_fifo_generate_warm(int c, int64_t buffer, const mqo_dist *src, const mqo_dist *srv,
                           int64_t total_served, int64_t warm, uint32_t seed, int64_t rep0, int64_t reps,
                           double *p, int64_t p_bins, mqo_fifo_out *out, int threads);

/* generator building blocks, exported so that tests can pin them */
void mqo_philox2x32_10(uint32_t c0, uint32_t c1, uint32_t key, uint32_t out[2]);
void mqo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* variate of `d` for arrival index i of replication rep; which = 0 source, 1 server */
double mqo_variate(const mqo_dist *d, uint32_t seed, uint32_t rep, uint32_t i, int which);

/* same for stream `sid` (priority class / network node): key = seed + STRIDE * 64 * sid */
double mqo_variate_sid(const mqo_dist *d, uint32_t seed, uint32_t sid, uint32_t rep, uint32_t i, int which);

/* ---------------------------------------------------------------------------------------
 * A variate source: either a caller array (replay) or the counter-based generator, which is
 * by definition the replay of the virtual array  x[i] = variate(word `which` of the primary
 * call (i, rep) of stream sid).
 * ------------------------------------------------------------------------------------- */
typedef struct {
    int generate;          /* 0: replay arr, 1: generator                                 */
    const double *arr;     /* replay: values, element i at arr[i * stride]                */
    int64_t n, stride;
    mqo_dist dist;         /* generator                                                    */
    uint32_t seed, sid, rep;
    int which;
    int64_t pos;           /* next index to hand out (in/out: number consumed)            */
} mqo_src;

/* ---- multi-class priority queue: PriorityQueueSimulator.run (sim/priority.py:622-657) ---- */
enum { MQO_NP = 0, MQO_PR = 1, MQO_RS = 2, MQO_RW = 3 };

typedef struct {
    double w_run[4], v_run[4]; /* running-mean moments, priority.py:676-690 */
    double w_sum[4], v_sum[4];
    int64_t arrived, taked, served, dropped, in_sys, queued;
    double t_old;
    double p_over;
} mqo_prio_class_out;

/* arrivals[k], services[k]: per-class sources (sources consumed in the reference's draw order:
 * arrival_time[k] draws at set_sources and at each class-k arrival, priority.py:130,300;
 * service draws at ServerPriority.start_service for non-resumed jobs, servers.py:241, and for
 * RS at the pre-emption, priority.py:425, from the ARRIVING class's stream).
 * p: [K][p_bins] time-in-state sums of in_sys[k] (priority.py:304,472).  Free server = lowest
 * index; pre-emption victim = first server in index order serving a weaker class
 * (priority.py:401-407).  buffer <= 0 means infinite (priority.py:373,437). */
int mqo_prio_run(int c, int K, int prty_type, int64_t buffer, mqo_src *arrivals, mqo_src *services,
                 int64_t total_served, double *p, int64_t p_bins, mqo_prio_class_out *out,
                 double *ttek_out);

double mqo_src_next(mqo_src *s, int *err);

/* ---- open network of FIFO nodes: NetworkSimulator.run (sim/networks/network.py:196-229) ---- */
#define MQO_NET_MAX_NODES 16
typedef struct {
    double v_run[4], w_run[4]; /* node-level running moments (QsSim.v / .w of qs[m]) */
    double v_sum[4], w_sum[4];
    int64_t arrived, taked, served, in_sys, queued;
} mqo_net_node_out;

typedef struct {
    double v_run[3], w_run[3]; /* v_network / w_network, network.py:116-135 (math.pow powers) */
    double v_sum[3], w_sum[3]; /* plain power sums (repeated multiplication) */
    int64_t served, arrived, in_sys;
    double ttek;
} mqo_net_out;

/* R: (m+1) x (m+1) row-major routing matrix (network.py:57-66): row 0 = source, row i+1 = node i,
 * column j < m = to node j, column m = leave.  arrivals: external gaps; routing: uniforms in the
 * order choose_next_node consumes them (network.py:109); services[i]: node i's service draws.
 * v_samples / w_samples (may be NULL): network sojourn / wait of each served job, in order. */
int mqo_net_run(int m, const int *channels, const double *R, mqo_src *arrivals, mqo_src *routing,
                mqo_src *services, int64_t job_served, mqo_net_out *out, mqo_net_node_out *nodes,
                double *v_samples, double *w_samples, int64_t cap);

/* ---- single long GI/G/1 FIFO chain: max-plus Lindley scan (DESIGN.md "K3") -------------------
 * Jobs n = 0..N-1; A[n] = gap before arrival n (A[0] = first arrival, sim/base.py:112), S[n] =
 * service of job n (servers.py:88).  With c = 1, infinite buffer, QsSim.run(N) is the Lindley
 * recursion  W[0] = 0,  W[n+1] = max(0, W[n] + x[n]),  x[n] = S[n] - A[n+1]  (wait = start - arrival,
 * base.py:302; start of job n+1 = departure of job n), v[n] = W[n] + S[n].
 * mqo_lindley_seq walks it sequentially; mqo_lindley_tree evaluates it with the kernel's fixed
 * association tree (segments of L jobs folded sequentially, Kogge-Stone over the 32 segments of a
 * warp, sequential over the warps of a 256-thread tile, then the tile level: 1024 chunks folded
 * sequentially, Kogge-Stone over 32 chunks, sequential over 32 groups), so that its W[] equals the
 * kernel's bit for bit.  A needs N+1 entries (x[N-1] uses A[N]).  w_in is the carry into job 0. */
typedef struct {
    double w_sum[4], v_sum[4];
    int64_t taked, served;
    double sum_a, sum_s; /* sum A[0..N-1], sum S[0..N-1] */
    double w_last_raw;   /* W[N-1] + x[N-1] before clamping: >= 0 iff job N was already waiting */
    double a, b;         /* max-plus summary of the whole range: W_out = max(a, W_in + b) */
} mqo_lindley_out;

int mqo_lindley_seq(const double *A, const double *S, int64_t N, double w_in, double *W, mqo_lindley_out *out);
int mqo_lindley_tree(const double *A, const double *S, int64_t N, double w_in, int L, double *W,
                     mqo_lindley_out *out);

/* ---- tandem line, single-server nodes, finite buffers, blocking after service:
 * TandemBlockingSim.run (sim/networks/tandem_blocking.py:50-142) ---- */
typedef struct {
    double throughput, loss_prob, v0; /* tandem_blocking.py:131-137 */
    double mean_jobs[MQO_NET_MAX_NODES];
    double area[MQO_NET_MAX_NODES];
    double t_end, t0, elapsed;
    int64_t served, arrived, lost, departures;
} mqo_tandem_out;

/* capacity[i] < 0 means unlimited (None).  services[i]: node i's draws in start_service order. */
int mqo_tandem_run(int m, const int64_t *capacity, mqo_src *arrivals, mqo_src *services,
                   int64_t total_served, double warmup_fraction, mqo_tandem_out *out);

/* ---- queue with warm-up, cold (vacation) and cold-delay periods: VacationQueueingSystemSimulator.run and
 * NPolicyQueueSim.run (sim/vacations.py:12-364) ---- */
typedef struct {
    int c;
    int64_t buffer;           /* < 0: infinite */
    int warm_set, cold_set, delay_set;
    int service_on_warm_up;   /* vacations.py:23 */
    int multiple_vacations;   /* vacations.py:24 */
    int64_t big_n;            /* > 0: N-policy (vacations.py:320-364) */
} mqo_vac_cfg;

typedef struct {
    double w_sum[4], v_sum[4], w_run[4], v_run[4], busy_run[3];
    int64_t busy_n, arrived, taked, served, dropped, in_sys, queued;
    double ttek, p_over;
    double warm_prob, cold_prob, delay_prob; /* QsPhase.prob: time in COMPLETED phases (phase.py:53) */
    int64_t warm_starts, cold_starts, delay_starts, warm_after_cold;
} mqo_vac_out;

/* services: the servers' dist draws in start_service order (servers.py:88); warm_services: the draws of
 * Server.warm_phase.dist (servers.py:90, is_service_on_warm_up); warm_src / cold_src / delay_src: durations of
 * the three QsPhase periods in start() order (phase.py:47).  Unused sources may be NULL. */
int mqo_vac_run(const mqo_vac_cfg *cfg, mqo_src *arrivals, mqo_src *services, mqo_src *warm_services,
                mqo_src *warm_src, mqo_src *cold_src, mqo_src *delay_src, int64_t total_served, double *p,
                int64_t p_bins, mqo_vac_out *out);

/* ---- queue with negative arrivals: QsSimNegatives.run (sim/negative.py:57-575), DISASTER / CLEAR_SYSTEM and
 * RCS / REMOVE ---- */
enum { MQO_NEG_DISASTER = 1, MQO_NEG_RCS = 2 }; /* NegativeServiceType values, negative.py:21-24 */

typedef struct {
    double w_sum[4], v_sum[4], vs_sum[4], vb_sum[4]; /* waits; sojourns of all / served / broken jobs */
    double w_run[4], v_run[4], vs_run[4], vb_run[4]; /* the reference's running moments                */
    int64_t positive_arrived, negative_arrived, taked, served, broken, total, dropped, in_sys, queued;
    double ttek, p_over;
} mqo_neg_out;

/* positives / negatives: gaps of the two sources; services: draws in start_service order; choices: uniforms that
 * pick the floor(u * busy)-th busy server (RCS only, may be NULL otherwise). */
/* scenario: 1 = CLEAR_SYSTEM / REMOVE, 2 = REQUEUE(_ALL), 3 = RESUME(_ALL), 4 = REQUEUE(_ALL)_NO_RESAMPLING (negative.py:27-54) */
int mqo_neg_run(int c, int64_t buffer, int type, int scenario, mqo_src *positives, mqo_src *negatives, mqo_src *services,
                mqo_src *choices, int64_t total_served, double *p, int64_t p_bins, mqo_neg_out *out);

/* ---- size-based single-server disciplines: SizeBasedQsSim.run (sim/size_based.py:64-248) ---- */
enum { MQO_SZ_FCFS = 0, MQO_SZ_SJF = 1, MQO_SZ_PSJF = 2, MQO_SZ_SRPT = 3, MQO_SZ_SPJF = 4, MQO_SZ_PSPJF = 5, MQO_SZ_SPRPT = 6 };

typedef struct {
    double w_sum[4], v_sum[4], w_run[4], v_run[4], busy_run[3];
    double slowdown_sum[2]; /* sum and sum of squares of sojourn / size over completed jobs (non-FCFS) */
    int64_t busy_n, arrived, taked, served, dropped, in_sys, queued, preemptions;
    double ttek, p_over;
} mqo_size_out;

/* sizes: job sizes in arrival order (size_based.py:155-164); for FCFS the service draws in start order */
/* pred_kind: 0 perfect predictor (Y = X), 1 Y = X * e with e from preds (Exp(1)), 2 Y = X * exp(sigma * z) with z from
 * preds (N(0,1)), 3 Y taken from preds as it is (replay) -- sim/utils/predictor.py */
int mqo_size_run(int discipline, int64_t buffer, mqo_src *arrivals, mqo_src *sizes, int pred_kind, double pred_sigma,
                 mqo_src *preds, int64_t total_served, double *p, int64_t p_bins, mqo_size_out *out);

/* ---- open network of PRIORITY nodes with per-class routing: PriorityNetwork.run
 * (sim/networks/priority_network.py:241-276) ---- */
#define MQO_PN_MAX_CLASSES 8
typedef struct {
    double v_run[3], w_run[3]; /* v_network[k], w_network[k] (math.pow moments, :152-173) */
    double v_sum[3], w_sum[3];
    int64_t served, arrived, in_sys;
} mqo_pn_class_out;

typedef struct {
    double v_run[4], w_run[4]; /* qs[node].v[j] / .w[j] for node class j */
    double v_sum[4], w_sum[4];
    int64_t arrived, taked, served, in_sys, queued;
} mqo_pn_node_class_out;

/* m nodes, K classes.  R: K matrices (m+1)x(m+1) row-major, class-major.  prty[i]: MQO_NP / PR / RS / RW of
 * node i.  nodes_prty[i*K + k]: node class of real class k at node i (:126-133).  arrivals[k]: external gaps of
 * class k; routing: uniforms in consumption order; services[i*K + j]: draws of node i, NODE class j, in draw
 * order.  out_cls[K]; out_node[m*K].  Free server = lowest index; victim = first weaker server. */
int mqo_prionet_run(int m, int K, const int *channels, const int *prty, const int *nodes_prty, const double *R,
                    mqo_src *arrivals, mqo_src *routing, mqo_src *services, int64_t job_served,
                    mqo_pn_class_out *out_cls, mqo_pn_node_class_out *out_node, double *ttek_out);

const char *mqo_version(void);

#ifdef __cplusplus
}
#endif
#endif
