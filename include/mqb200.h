/*
 * mqb200.h -- C ABI of libmqb200.so, the B200-native batched discrete-event simulation
 * engine that stands in for most-queue's pure-Python simulation loops.
 *
 * The reference has no FFI: its hot path sits behind Python objects
 * (QsSim / PriorityQueueSimulator / NetworkSimulator).  The entry points below are what a
 * ctypes binding of those objects' run() methods binds (INTEGRATION.md shows the stub);
 * each one names the reference interface it replaces (paths relative to most_queue/).
 *
 * Conventions: plain C types only, return 0 on success and a negative MQ_E_* code on
 * failure (mq_last_error() holds the message, thread-local); the library never keeps a
 * caller pointer after a call returns; every output buffer is allocated by the caller.
 * One mq_ctx drives one GPU and is not re-entrant; use one context (one process) per GPU.
 */
#ifndef MQB200_H
#define MQB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MQ_ABI_VERSION 1

/* error codes */
#define MQ_OK 0
#define MQ_E_INVALID (-1)     /* bad argument (the Python layer raises ValueError)        */
#define MQ_E_CUDA (-2)        /* CUDA runtime error                                       */
#define MQ_E_NODEVICE (-3)    /* no usable CUDA device                                    */
#define MQ_E_INPUT_SHORT (-4) /* replay arrays ran out before total_served departures     */
#define MQ_E_UNSUPPORTED (-5) /* configuration outside what the kernels cover             */
#define MQ_E_OVERFLOW (-6)    /* a bounded per-lane structure overflowed                  */

/* distribution kinds == Kendall codes of random/utils/create.py:19-80 */
#define MQ_M 0       /* "M"       p = {mu}                  distributions.py:887-914 */
#define MQ_H2 1      /* "H"       p = {p1, mu1, mu2}        distributions.py:407-429 */
#define MQ_E 2       /* "E"       p = {r, mu}               distributions.py:807-820 */
#define MQ_GAMMA 3   /* "Gamma"   p = {mu, alpha}           distributions.py:996-1003 */
#define MQ_PA 4      /* "Pa"      p = {alpha, K}            distributions.py:749-761 */
#define MQ_D 5       /* "D"       p = {b}                   distributions.py:614-626 */
#define MQ_C2 6      /* "C"       p = {p1, mu1, mu2}        distributions.py:542-560 */
#define MQ_UNIFORM 7 /* "Uniform" p = {mean, half_interval} distributions.py:298-309 */
#define MQ_NORM 8    /* "Norm"    p = {mean, std_dev}       distributions.py:209-216 */
#define MQ_WEIBULL 9 /* "Weibull" p = {k, W}                distributions.py:94-102  */

typedef struct {
    int32_t kind;
    int32_t reserved;
    double p[4];
} mq_dist;

typedef struct mq_ctx mq_ctx;

/* ---- per-replication record: one fp64 row per field, replication-minor ([field][rep]) ---- */
#define MQ_F_W 0        /* 4 rows: sum of w^k, k=1..4 over jobs that started service       */
#define MQ_F_V 4        /* 4 rows: sum of v^k over served jobs                             */
#define MQ_F_ARRIVED 8  /* QsSim.arrived  (base.py:222)                                    */
#define MQ_F_TAKED 9    /* QsSim.taked    (base.py:172,300) = number of w samples          */
#define MQ_F_SERVED 10  /* QsSim.served   (base.py:269)     = number of v samples          */
#define MQ_F_DROPPED 11 /* QsSim.dropped  (base.py:211)                                    */
#define MQ_F_TTEK 12    /* QsSim.ttek at stop                                              */
#define MQ_F_FLAGS 13   /* non-zero: replication hit a library limit (see MQ_FLAG_*)       */
#define MQ_F_POVER 14   /* time spent in states >= p_bins                                  */
#define MQ_F_RESERVED 15
#define MQ_F_P 16       /* p_bins rows: time spent with exactly j jobs in the system       */
#define MQ_REC_ROWS(p_bins) (16 + (p_bins))

#define MQ_FLAG_QUEUE_OVERFLOW 1

/* ---- aggregate over replications (fp64 vector, what the multi-GPU allreduce sums) ---- */
#define MQ_A_REPS 0      /* number of replications summed                                  */
#define MQ_A_W 1         /* 4: sum over reps of the per-rep moment estimate  sum(w^k)/taked */
#define MQ_A_W2 5        /* 4: sum over reps of that estimate squared (for CIs)            */
#define MQ_A_V 9         /* 4 */
#define MQ_A_V2 13       /* 4 */
#define MQ_A_ARRIVED 17
#define MQ_A_TAKED 18
#define MQ_A_SERVED 19
#define MQ_A_DROPPED 20
#define MQ_A_TTEK 21
#define MQ_A_WPOOL 22    /* 4: sum over reps of sum(w^k) (pooled moments)                  */
#define MQ_A_VPOOL 26    /* 4 */
#define MQ_A_FLAGGED 30  /* replications with a non-zero flag                              */
#define MQ_A_POVER 31    /* sum over reps of p_over/ttek                                   */
#define MQ_A_P 32        /* p_bins: sum over reps of p_time[j]/ttek, then p_bins squares   */
#define MQ_AGG_LEN(p_bins) (32 + 2 * (p_bins))

/* GI/G/c/r FIFO, generator mode.  Replaces QsSim.set_sources/set_servers/run
 * (sim/base.py:95,114,488) for `replications` independent runs. */
typedef struct {
    int32_t c;            /* num_of_channels                               base.py:29   */
    int32_t p_bins;       /* state-probability bins reported (1..MQ_MAX_PBINS)          */
    int64_t buffer;       /* waiting places, -1 = infinite                 base.py:204  */
    mq_dist src;          /* set_sources(params, kendall_notation)         base.py:95   */
    mq_dist srv;          /* set_servers(params, kendall_notation)         base.py:114  */
    int64_t total_served; /* run(total_served): departures per replication base.py:505  */
    int64_t replications; /* replications simulated by THIS call (this rank's shard)    */
    int64_t rep0;         /* global id of the first one: Philox counter word            */
    uint64_t seed;        /* QsSim(seed=...) -> Philox key                 base_core.py:23 */
    int32_t flags;        /* MQ_FIFO_V_AT_START | MQ_FIFO_FP64 | MQ_FIFO_BUSY or 0      */
    int32_t warmup;       /* additive (0 = the reference's behaviour): statistics restart at the     */
                          /* instant of the warmup-th start of service; w / v then cover the jobs     */
                          /* started after it (rows TAKED / SERVED hold those sample counts) and      */
                          /* p / TTEK the time after it.  Removes QsSim's O(1/N) start-up bias.       */
} mq_fifo_cfg;

/* additive: sample v for every job that STARTED service before the stop (its completion time is known
 * then) instead of only the jobs that left, which removes the right-censoring bias of the reference's
 * stop rule (the <= c jobs in service at the stop are long ones); SERVED then equals TAKED */
#define MQ_FIFO_V_AT_START 1
/* additive: run the same loop in fp64 (clocks, variates via the CUDA double math library, power sums) on the same
 * Philox words -- the A/B tier that measures what the fp32 production arithmetic costs against the reference's
 * Python floats; several times slower */
#define MQ_FIFO_FP64 2
/* also gather QsSim.busy, the busy-period raw moments (base.py:179-184, 274-281, 297-298).  The per-replication
 * record then has MQ_FIFO_BUSY_ROWS more rows behind MQ_REC_ROWS(p_bins) and the aggregate MQ_AGG_BUSY_LEN more
 * entries behind MQ_AGG_LEN(p_bins): [0..2] sum over replications of sum(b^k)/n, [3] sum of n, [4..7] their squares */
#define MQ_FIFO_BUSY 4
#define MQ_FB_BUSY 0     /* 3 rows: sum of (busy period)^k, k = 1..3 */
#define MQ_FB_BUSYN 3    /* number of busy periods (QsSim.busy_moments) */
#define MQ_FIFO_BUSY_ROWS 4
#define MQ_AGG_BUSY_LEN 8

#define MQ_MAX_PBINS 1024
#define MQ_MAX_CHANNELS 32

/* The same call on SEVERAL GPUs from one host thread: the replications of cfg (global ids rep0 .. rep0 + replications)
 * are cut into one contiguous share per context, every context (one per device, mq_create) simulates its share on its
 * own host thread, and the aggregates -- sums over replications -- are added.  This is what a single-process caller of
 * qs.run(n, replications=R) uses to fill a box without torchrun; results do not depend on the number of contexts
 * (Philox counters carry the global replication id).  rep_out is not available in this form. */
int mq_sim_fifo_multi(mq_ctx *const *ctxs, int n_ctx, const mq_fifo_cfg *cfg, double *agg);

/* GI/G/c/r FIFO, replay mode: variates come from caller arrays instead of the generator.
 * A[i * replications + r] is the i-th interarrival draw of replication r (A[0][r] is the
 * draw set_sources makes, base.py:112); S[j * replications + r] the j-th service draw
 * (servers.py:88).  fp64 throughout, bit-exact with the reference for the same arrays. */
typedef struct {
    int32_t c;
    int32_t p_bins;
    int64_t buffer;
    int64_t total_served;
    int64_t replications;
    int64_t n_a;          /* rows of A */
    int64_t n_s;          /* rows of S */
    int32_t device_ptrs;  /* 1: A, S (and rep_out when given) are device pointers       */
    int32_t flags;
} mq_replay_cfg;

/* GI/G/c FIFO with an INFINITE buffer, replay mode, job-indexed (csrc/k1b_fifo_jobs.cu): the HBM-roofline tier of the
 * replay.  Jobs n = 0..jobs-1 of every replication in arrival order (Kiefer-Wolfowitz recursion): arrival time
 * t_n = t_{n-1} + A[n] (base.py:230-236), the job takes the server that frees first, wait w_n = max(t_n, m) - t_n
 * (base.py:300-303), sojourn v_n = (max(t_n, m) + S[n]) - t_n (servers.py:88-90, base.py:269-272).  Every w_n, v_n equals
 * the sample QsSim produces for that job bit for bit; the sums run over the first `jobs` ARRIVALS in arrival order
 * (QsSim.run stops at the jobs-th DEPARTURE, which leaves out the <= c-1 jobs then in service), and no time-in-state
 * histogram is kept.  A[n * replications + r], S[n * replications + r]: `jobs` rows each. */
typedef struct {
    int32_t c;            /* num_of_channels                                             base.py:29 */
    int32_t device_ptrs;  /* 1: A, S (and rep_out, W, V when given) are device pointers              */
    int64_t jobs;
    int64_t replications;
    int32_t flags;
    int32_t reserved;
} mq_jobs_cfg;

/* per-replication record rows of mq_replay_fifo_jobs */
#define MQ_JB_W 0      /* 4: sum of w^k in arrival order: s1 += w, s2 += w*w, s3 = fma(w*w, w, s3), s4 = fma(w*w, w*w, s4) */
#define MQ_JB_V 4      /* 4: the same for v                                                        */
#define MQ_JB_JOBS 8   /* number of samples (= jobs)                                               */
#define MQ_JB_TLAST 9  /* arrival time of the last job                                             */
#define MQ_JB_DMAX 10  /* largest completion time: the clock when the last of the jobs has left    */
#define MQ_JB_WAITED 11 /* jobs with w > 0                                                         */
#define MQ_JB_ROWS 12
/* aggregate: for each of the 12 rows the sum over replications (rows 0-7 divided by jobs first: the per-replication
 * moment estimates), then the 12 sums of squares */
#define MQ_JB_AGG_LEN 24

/* agg, rep_out, W, V may be NULL.  W / V: [jobs][replications] per-job waits / sojourn times. */
int mq_replay_fifo_jobs(mq_ctx *ctx, const mq_jobs_cfg *cfg, const double *A, const double *S, double *agg,
                        double *rep_out, double *W, double *V);

/* extra rows appended to the replay record after MQ_REC_ROWS(p_bins): the reference's own
 * running-mean moments (stats_update.py:6-22), bit-exact */
#define MQ_RF_WRUN 0    /* 4 */
#define MQ_RF_VRUN 4    /* 4 */
#define MQ_RF_BUSYRUN 8 /* 3: QsSim.busy (base.py:276-281)                              */
#define MQ_RF_BUSYN 11
#define MQ_RF_INSYS 12
#define MQ_RF_QUEUED 13
#define MQ_RF_AUSED 14
#define MQ_RF_SUSED 15
#define MQ_REPLAY_EXTRA 16
#define MQ_REPLAY_ROWS(p_bins) (MQ_REC_ROWS(p_bins) + MQ_REPLAY_EXTRA)

/* ---- multi-class priority queue.  Replaces PriorityQueueSimulator(num_of_channels,
 * num_of_classes, prty_type, buffer).set_sources/set_servers/run (sim/priority.py:28,111,134,622) */
#define MQ_MAX_CLASSES 8
#define MQ_PRIO_MAX_CHANNELS 16
#define MQ_PRTY_NP 0 /* "NP" non-preemptive                       */
#define MQ_PRTY_PR 1 /* "PR" preemptive resume                    */
#define MQ_PRTY_RS 2 /* "RS" preemptive repeat with resampling    */
#define MQ_PRTY_RW 3 /* "RW" preemptive repeat without resampling */

typedef struct {
    int32_t c;             /* num_of_channels                                   priority.py:28 */
    int32_t K;             /* num_of_classes                                                   */
    int32_t prty_type;     /* MQ_PRTY_*                                                        */
    int32_t p_bins;        /* per-class state bins                                             */
    int64_t buffer;        /* total waiting places, <= 0 infinite         priority.py:373,437  */
    int64_t total_served;  /* run(total_served): departures of all classes priority.py:636     */
    int64_t replications;
    int64_t rep0;
    uint64_t seed;
    int32_t queue_cap;     /* per-class waiting-line capacity of the engine (0: default 512);  */
                           /* a replication that exceeds it is flagged, never silently wrong   */
    int32_t flags;
    mq_dist src[MQ_MAX_CLASSES]; /* set_sources([{"type","params"}]*K)          priority.py:111 */
    mq_dist srv[MQ_MAX_CLASSES]; /* set_servers([{"type","params"}]*K)          priority.py:134 */
} mq_prio_cfg;

/* per-replication record: 2 global rows, then for each class MQ_PRIO_ROWS(p_bins) rows */
#define MQ_PG_TTEK 0
#define MQ_PG_FLAGS 1
#define MQ_PG_ROWS 2
#define MQ_PF_WSUM 0   /* 4: sum of w^k over served jobs of the class (priority.py:480)          */
#define MQ_PF_VSUM 4   /* 4 */
#define MQ_PF_WRUN 8   /* 4: the reference's running-mean moments, exact in replay mode          */
#define MQ_PF_VRUN 12  /* 4 */
#define MQ_PF_ARRIVED 16
#define MQ_PF_TAKED 17
#define MQ_PF_SERVED 18
#define MQ_PF_DROPPED 19
#define MQ_PF_INSYS 20
#define MQ_PF_QUEUED 21
#define MQ_PF_TOLD 22
#define MQ_PF_POVER 23
#define MQ_PF_AUSED 24
#define MQ_PF_SUSED 25
#define MQ_PF_P 26     /* p_bins: time with exactly j class-k jobs in the system (priority.py:304,472) */
#define MQ_PRIO_ROWS(p_bins) (26 + (p_bins))
#define MQ_PRIO_REC_ROWS(K, p_bins) (2 + (K) * MQ_PRIO_ROWS(p_bins))

/* aggregate: [0] replications, [1] sum ttek, [2] flagged replications, [3] reserved, then per class
 * MQ_PRIO_AGG(p_bins) entries: w[4] (sum over reps of sum(w^k)/served), w2[4] (squares), v[4], v2[4],
 * arrived, taked, served, dropped, 4 reserved, p[p_bins] (sum of p_time/ttek), p2[p_bins] */
#define MQ_PRIO_AGG(p_bins) (24 + 2 * (p_bins))
#define MQ_PRIO_AGG_LEN(K, p_bins) (4 + (K) * MQ_PRIO_AGG(p_bins))

int mq_sim_prio(mq_ctx *ctx, const mq_prio_cfg *cfg, double *agg, double *rep_out);

/* replay form: A[k] / S[k] are class k's gap / service rows, [n][replications] fp64 host arrays
 * (A[k][0] is the draw set_sources makes, priority.py:130; S[k][j] the j-th draw from class k's
 * service stream in the reference's draw order).  src/srv of cfg are ignored. */
int mq_replay_prio(mq_ctx *ctx, const mq_prio_cfg *cfg, const double *const *A, const int64_t *n_a,
                   const double *const *S, const int64_t *n_s, double *agg, double *rep_out);

/* ---- open network of FIFO multi-server nodes with Markov routing.  Replaces
 * NetworkSimulator().set_sources(arrival_rate, R, source_kendall, source_params) /
 * set_nodes(serv_params, n) / run(job_served)  (sim/networks/network.py:53,82,196) */
#define MQ_NET_MAX_NODES 16
#define MQ_NET_MAX_SERVERS 64 /* sum of channels over all nodes */

typedef struct {
    int32_t m;                               /* number of nodes                     network.py:93 */
    int32_t flags;
    int32_t channels[MQ_NET_MAX_NODES];      /* n[i]                                network.py:82 */
    double R[(MQ_NET_MAX_NODES + 1) * (MQ_NET_MAX_NODES + 1)]; /* (m+1)x(m+1), row-major, packed   */
                                             /* with stride m+1: row 0 = source, row i+1 = node i, */
                                             /* column m = leave                    network.py:57-66 */
    mq_dist src;                             /* external flow ("M" arrival_rate)    network.py:72-76 */
    mq_dist srv[MQ_NET_MAX_NODES];           /* serv_params[i]                      network.py:95-97 */
    int64_t job_served;                      /* run(job_served)                     network.py:211 */
    int64_t replications;
    int64_t rep0;
    uint64_t seed;
    int32_t queue_cap;                       /* per-node waiting-line capacity (0: default 512)    */
    int32_t reserved;
} mq_net_cfg;

/* per-replication record rows */
#define MQ_NG_TTEK 0
#define MQ_NG_FLAGS 1
#define MQ_NG_SERVED 2
#define MQ_NG_ARRIVED 3
#define MQ_NG_INSYS 4
#define MQ_NG_AUSED 5
#define MQ_NG_UUSED 6
#define MQ_NG_VSUM 8   /* 3: sum of (network sojourn)^k over served jobs  network.py:188 */
#define MQ_NG_WSUM 11  /* 3: sum of (network wait)^k                      network.py:189 */
#define MQ_NG_VRUN 14  /* 3: v_network running moments (network.py:116-124)             */
#define MQ_NG_WRUN 17  /* 3 */
#define MQ_NG_ROWS 20
#define MQ_NN_VSUM 0   /* 4: node-level sojourn power sums (qs[i].v)                    */
#define MQ_NN_WSUM 4   /* 4 */
#define MQ_NN_VRUN 8   /* 4: qs[i].v running moments, exact in replay mode              */
#define MQ_NN_WRUN 12  /* 4 */
#define MQ_NN_ARRIVED 16
#define MQ_NN_TAKED 17
#define MQ_NN_SERVED 18
#define MQ_NN_SUSED 19
#define MQ_NN_INSYS 20
#define MQ_NN_QUEUED 21
#define MQ_NN_ROWS 22
#define MQ_NET_REC_ROWS(m) (MQ_NG_ROWS + (m) * MQ_NN_ROWS)

/* aggregate: [0] replications, [1] sum ttek, [2] flagged, [3] sum served, [4] sum arrived,
 * [5..7] sum over reps of v_sum[k]/served, [8..10] squares, [11..13] same for w, [14..16] squares,
 * then per node 16 entries: v[4] (v_sum/served), w[4] (w_sum/taked), served, taked, arrived,
 * v1 squared, w1 squared, 3 reserved */
#define MQ_NET_AGG_LEN(m) (20 + 16 * (m))

int mq_sim_net(mq_ctx *ctx, const mq_net_cfg *cfg, double *agg, double *rep_out);

/* replay form: A = external gaps [n_a][replications] (A[0] is the draw set_sources makes,
 * network.py:76), U = routing uniforms [n_u][replications] in the order choose_next_node consumes
 * them (network.py:109), S[i] = node i's service draws [n_s[i]][replications] (servers.py:88). */
int mq_replay_net(mq_ctx *ctx, const mq_net_cfg *cfg, const double *A, int64_t n_a, const double *U,
                  int64_t n_u, const double *const *S, const int64_t *n_s, double *agg, double *rep_out);

/* ---- one very long GI/G/1 FIFO chain: max-plus Lindley scan.  Replaces QsSim(1).run(N)
 * (sim/base.py:488 with num_of_channels = 1, buffer = None) for N far beyond what a sequential
 * walk can do.  Jobs n = 0..N-1, A[n] = gap before arrival n, S[n] = service of job n. */
typedef struct {
    mq_dist src;          /* generator form: set_sources                        base.py:95   */
    mq_dist srv;          /*                 set_servers                        base.py:114  */
    int64_t jobs;         /* jobs of THIS range                                              */
    int64_t job0;         /* global index of its first job (Philox counter; 0 on one GPU)    */
    uint64_t seed;
    uint32_t chain;       /* chain id (generator stream)                                     */
    int32_t device_ptrs;  /* replay form: A, S (and W) are device pointers                   */
    int32_t last_range;   /* 1: this range ends the chain (apply the stop rule, base.py:505) */
    int32_t flags;        /* MQ_SCAN_WANT_P | MQ_SCAN_KEEP or 0                              */
    int64_t jobs_ahead;   /* jobs of the chain after this range (0 for the last range); in   */
                          /* the replay form A must then hold jobs + 1 + jobs_ahead gaps     */
} mq_scan_cfg;

/* also produce the time-in-state levels (base.py:327-340 for one channel): stats[MQ_S_LEVELS + m - 1] =
 * time with N(t) >= m for m = 1..32 and stats[MQ_S_LEVELS + 32] = sum over m > 32 of that time;
 * p[0] = 1 - T[>=1]/ttek, p[m] = (T[>=m] - T[>=m+1]) / ttek.  Only jobs 0..N-1 are counted (the
 * reference also counts the few arrivals between the last arrival and the stop). */
#define MQ_SCAN_WANT_P 1
/* mq_scan_gg1_summary only: keep the range's intermediate results on the device; if the NEXT call on the context is
 * mq_scan_gg1_apply with the same range description (jobs, job0, seed, chain, laws / array pointers, jobs_ahead,
 * MQ_SCAN_WANT_P), it walks from them instead of forming them again.  This is the multi-GPU sequence: summary ->
 * all-gather of the (a, b) pairs -> apply from the carry. */
#define MQ_SCAN_KEEP 2

#define MQ_S_W 0       /* 4: sum of W^k over the jobs of the range (+ job N when last_range and it  */
                       /*    was already waiting at the N-th departure, base.py:300-305)            */
#define MQ_S_V 4       /* 4: sum of (W + S)^k                                                       */
#define MQ_S_SUMA 8    /* sum of the range's gaps                                                   */
#define MQ_S_SUMS 9    /* sum of the range's service times (= busy time of the single server)       */
#define MQ_S_TAKED 10  /* number of w samples                                                        */
#define MQ_S_SERVED 11 /* number of v samples                                                        */
#define MQ_S_A 12      /* max-plus summary of the range: W_out = max(a, W_in + b)                    */
#define MQ_S_B 13
#define MQ_S_WIN 14
#define MQ_S_WOUT 15   /* carry into the next range                                                  */
#define MQ_S_TAILRAW 16 /* W[n-1] + x[n-1] before clamping                                           */
#define MQ_S_TAILV 17  /* W[n-1] + S[n-1]: clock at the last departure minus the last arrival time   */
#define MQ_S_LEVELS 20 /* 33 entries, see MQ_SCAN_WANT_P                                             */
#define MQ_SCAN_LEN 56

/* A, S == NULL: generator form.  Replay form: A has jobs + 1 entries (x[n] uses A[n+1]), S has jobs.
 * summary: ab[2] = (a, b) of the range.  apply: walks the range from carry w_in; stats has
 * MQ_SCAN_LEN doubles; W (optional, may be NULL) receives the per-job waits.
 * mq_scan_gg1 = summary + apply with w_in = 0 for a chain held by one GPU. */
int mq_scan_gg1_summary(mq_ctx *ctx, const mq_scan_cfg *cfg, const double *A, const double *S, double ab[2]);
int mq_scan_gg1_apply(mq_ctx *ctx, const mq_scan_cfg *cfg, const double *A, const double *S, double w_in,
                      double *stats, double *W);
int mq_scan_gg1(mq_ctx *ctx, const mq_scan_cfg *cfg, const double *A, const double *S, double *stats, double *W);

/* ---- tandem line of single-server nodes with finite buffers and blocking after service.
 * Replaces TandemBlockingSim(seed).set_sources(arrival_rate, kendall) / set_nodes(serv_params,
 * capacity) / run(total_served, warmup_fraction)  (sim/networks/tandem_blocking.py:32,38,50) */
typedef struct {
    int32_t m;                              /* number of nodes                                    */
    int32_t flags;
    int64_t capacity[MQ_NET_MAX_NODES];     /* K_i: max jobs at node i incl. the one in service,  */
                                            /* < 0 = unlimited (None)          tandem_blocking.py:45 */
    mq_dist src;                            /* external flow into node 1        tandem_blocking.py:35 */
    mq_dist srv[MQ_NET_MAX_NODES];          /* per-node service                 tandem_blocking.py:44 */
    int64_t total_served;                   /* departures from the last node AFTER warm-up   :52   */
    double warmup_fraction;                 /* default 0.05 in the reference                 :50   */
    int64_t replications;
    int64_t rep0;
    uint64_t seed;
} mq_tandem_cfg;

/* per-replication record rows */
#define MQ_TG_ELAPSED 0   /* t - t0 (tandem_blocking.py:130)                                        */
#define MQ_TG_FLAGS 1
#define MQ_TG_ARRIVED 2   /* counted after warm-up                                                  */
#define MQ_TG_LOST 3
#define MQ_TG_DEPARTURES 4
#define MQ_TG_SERVED 5    /* including warm-up                                                      */
#define MQ_TG_AREASUM 6   /* sum over nodes of area_l[i]  (v[0] = this / departures)                */
#define MQ_TG_TEND 7
#define MQ_TG_T0 8
#define MQ_TG_AUSED 9
#define MQ_TG_ROWS 10
#define MQ_TN_AREA 0      /* area_l[i] = integral of jobs[i] dt after warm-up                       */
#define MQ_TN_SUSED 1
#define MQ_TN_ROWS 2
#define MQ_TANDEM_REC_ROWS(m) (MQ_TG_ROWS + (m) * MQ_TN_ROWS)

/* aggregate: [0] replications, [1] flagged, [2],[3] sum and sum of squares over replications of the
 * throughput departures/elapsed, [4],[5] of the loss probability lost/arrived, [6],[7] of
 * v[0] = sum(mean_jobs)/throughput, then per node [8+2i],[9+2i] of mean_jobs[i] = area/elapsed */
#define MQ_TANDEM_AGG_LEN(m) (8 + 2 * (m))

int mq_sim_tandem(mq_ctx *ctx, const mq_tandem_cfg *cfg, double *agg, double *rep_out);
/* replay form: A = external gaps [n_a][replications], S[i] = node i's draws in start_service order */
int mq_replay_tandem(mq_ctx *ctx, const mq_tandem_cfg *cfg, const double *A, int64_t n_a, const double *const *S,
                     const int64_t *n_s, double *agg, double *rep_out);

/* ---- open network of priority nodes with per-class routing.  Replaces PriorityNetwork(k_num)
 * .set_sources(L, R) / set_nodes(serv_params, n, prty, nodes_prty) / run(job_served)
 * (sim/networks/priority_network.py:28,60,87,241) */
#define MQ_PN_MAX_NODES 8
#define MQ_PN_MAX_CLASSES 4
#define MQ_PN_MAX_SERVERS 32 /* sum of channels over all nodes */

typedef struct {
    int32_t m;                                   /* nodes                                              */
    int32_t K;                                   /* classes (k_num)                      :28           */
    int32_t channels[MQ_PN_MAX_NODES];           /* n[i]                                 :87           */
    int32_t prty[MQ_PN_MAX_NODES];               /* MQ_PRTY_* of node i                  :104-110      */
    int32_t nodes_prty[MQ_PN_MAX_NODES * MQ_PN_MAX_CLASSES]; /* [i*K + k]: node class of real class k  */
    double R[MQ_PN_MAX_CLASSES * (MQ_PN_MAX_NODES + 1) * (MQ_PN_MAX_NODES + 1)]; /* K matrices (m+1)x(m+1),  */
                                                 /* packed with stride m+1, class-major  :60-76        */
    mq_dist src[MQ_PN_MAX_CLASSES];              /* external flow of class k ("M", L[k]) :82           */
    mq_dist srv[MQ_PN_MAX_NODES * MQ_PN_MAX_CLASSES]; /* [i*K + j]: service of NODE class j at node i, */
                                                 /* i.e. serv_params[i][nodes_prty[i][j]] (:129-131)   */
    int64_t job_served;                          /* run(job_served): departures of all classes :257    */
    int64_t replications;
    int64_t rep0;
    uint64_t seed;
    int32_t queue_cap;                           /* per (node, class) waiting-line capacity (0: 128)   */
    int32_t flags;
} mq_prionet_cfg;

/* per-replication record: 2 global rows (ttek, flags), then per class 16 rows, then per (node, node class) 20 */
#define MQ_PNC_VSUM 0   /* 3: sum of (network sojourn)^k                                   :232      */
#define MQ_PNC_WSUM 3   /* 3 */
#define MQ_PNC_VRUN 6   /* 3: v_network[k] running moments                                 :152-161  */
#define MQ_PNC_WRUN 9   /* 3 */
#define MQ_PNC_SERVED 12
#define MQ_PNC_ARRIVED 13
#define MQ_PNC_AUSED 14
#define MQ_PNC_ROWS 16
#define MQ_PNN_VSUM 0   /* 4: node-level sojourn power sums of node class j (qs[i].v[j])              */
#define MQ_PNN_WSUM 4
#define MQ_PNN_VRUN 8
#define MQ_PNN_WRUN 12
#define MQ_PNN_SERVED 16
#define MQ_PNN_TAKED 17
#define MQ_PNN_SUSED 18
#define MQ_PNN_ROWS 20
#define MQ_PN_G_TTEK 0
#define MQ_PN_G_FLAGS 1
#define MQ_PN_G_UUSED 2
#define MQ_PN_G_ROWS 4
#define MQ_PN_REC_ROWS(m, K) (MQ_PN_G_ROWS + (K) * MQ_PNC_ROWS + (m) * (K) * MQ_PNN_ROWS)

/* aggregate: [0] replications, [1] sum ttek, [2] flagged, [3] reserved, then per class 16 entries:
 * v[3] (sum over reps of v_sum/served), v2[3] (squares), w[3], w2[3], served, arrived, 2 reserved */
#define MQ_PN_AGG_LEN(K) (4 + 16 * (K))

int mq_sim_prionet(mq_ctx *ctx, const mq_prionet_cfg *cfg, double *agg, double *rep_out);
/* replay form: A[k] external gaps of class k, U routing uniforms, S[i*K + j] draws of node i, node class j */
int mq_replay_prionet(mq_ctx *ctx, const mq_prionet_cfg *cfg, const double *const *A, const int64_t *n_a,
                      const double *U, int64_t n_u, const double *const *S, const int64_t *n_s, double *agg,
                      double *rep_out);

/* ---- queue with warm-up, cold (vacation) and cold-delay periods, and the N-policy queue.  Replaces
 * VacationQueueingSystemSimulator(n, buffer, is_service_on_warm_up, is_multiple_vacations)
 * .set_sources / set_servers / set_warm / set_cold / set_cold_delay / run(total_served) and
 * NPolicyQueueSim(n, big_n) (sim/vacations.py:12,52,79,89,303,320) */
#define MQ_VAC_STREAMS 6 /* arrivals, services, warm services, warm-up, cold, cold-delay periods */
#define MQ_VS_ARRIVALS 0
#define MQ_VS_SERVICES 1
#define MQ_VS_WARM_SERVICES 2
#define MQ_VS_WARM 3
#define MQ_VS_COLD 4
#define MQ_VS_DELAY 5

typedef struct {
    int32_t c;                        /* num_of_channels (1..32)                             :19 */
    int32_t p_bins;
    int64_t buffer;                   /* < 0: infinite                                       :20 */
    mq_dist dist[MQ_VAC_STREAMS];     /* law of each stream (generator form); see set[]            */
    int32_t set[MQ_VAC_STREAMS];      /* set[MQ_VS_WARM]: set_warm was called (:52); set[MQ_VS_WARM_SERVICES]: */
                                      /* the servers got the warm law (:66-76); set[MQ_VS_COLD] :79; DELAY :89  */
    int32_t service_on_warm_up;       /* :23 */
    int32_t multiple_vacations;       /* :24 */
    int64_t big_n;                    /* > 0: N-policy, the cold law is the constant 1e100  :345  */
    int64_t total_served;
    int64_t replications;
    int64_t rep0;
    uint64_t seed;
    int32_t queue_cap;                /* infinite buffer: waiting-line capacity of the engine (0: 1024) */
    int32_t flags;
} mq_vac_cfg;

/* per-replication record rows */
#define MQ_V_TTEK 0
#define MQ_V_FLAGS 1
#define MQ_V_ARRIVED 2
#define MQ_V_TAKED 3
#define MQ_V_SERVED 4
#define MQ_V_DROPPED 5
#define MQ_V_INSYS 6
#define MQ_V_QUEUED 7
#define MQ_V_BUSYN 8
#define MQ_V_POVER 9
#define MQ_V_WSUM 10    /* 4 */
#define MQ_V_VSUM 14    /* 4 */
#define MQ_V_WRUN 18    /* 4: the reference's running moments (stats_update.py:6-22) */
#define MQ_V_VRUN 22    /* 4 */
#define MQ_V_BUSYRUN 26 /* 3 */
#define MQ_V_WARMPROB 29  /* QsPhase.prob of the three periods: time in COMPLETED periods (phase.py:53) */
#define MQ_V_COLDPROB 30
#define MQ_V_DELAYPROB 31
#define MQ_V_STARTS 32  /* 3: starts_times of warm, cold, delay */
#define MQ_V_WARM_AFTER_COLD 35
#define MQ_V_USED 36    /* 6: draws consumed per stream */
#define MQ_V_P 44       /* p_bins time-in-state sums */
#define MQ_V_ROWS(pb) (44 + (pb))

/* aggregate: [0] replications, [1] sum ttek, [2] flagged, [3] sum served, [4] sum arrived, [5] sum dropped,
 * [6] sum taked, [7] reserved; [8..11] sum over replications of w_sum/taked, [12..15] squares; [16..19] v_sum/served,
 * [20..23] squares; [24..26] warm/cold/delay time fractions (prob/ttek), [27..29] squares; [32..32+pb) p = time in
 * state / ttek, then pb squares */
#define MQ_V_AGG_LEN(pb) (32 + 2 * (pb))

int mq_sim_vacation(mq_ctx *ctx, const mq_vac_cfg *cfg, double *agg, double *rep_out);
/* replay form: X[s] = draws of stream s ([n[s]][replications], NULL / 0 when unused) */
int mq_replay_vacation(mq_ctx *ctx, const mq_vac_cfg *cfg, const double *const *X, const int64_t *n, double *agg,
                       double *rep_out);

/* ---- queue with negative arrivals.  Replaces QsSimNegatives(n, type_of_negatives, buffer).set_positive_sources /
 * set_negative_sources / set_servers / run(total_served) (sim/negative.py:62,116,134,184) for the DISASTER and RCS
 * types with all four scenarios each (negative.py:27-54); the RCH / RCE types are rejected with MQ_E_UNSUPPORTED */
#define MQ_NEG_DISASTER 1 /* NegativeServiceType.DISASTER, negative.py:21 */
#define MQ_NEG_RCS 2      /* NegativeServiceType.RCS,      negative.py:22 */
#define MQ_NEG_REMOVE 1        /* DisasterScenario.CLEAR_SYSTEM / RcsScenario.REMOVE: the jobs leave as "broken"        */
#define MQ_NEG_REQUEUE 2       /* REQUEUE_ALL / REQUEUE: the interrupted jobs restart with a new service draw         */
#define MQ_NEG_RESUME 3        /* RESUME_ALL / RESUME: they continue with their remaining service time                */
#define MQ_NEG_REQUEUE_FIXED 4 /* REQUEUE_ALL_NO_RESAMPLING / REQUEUE_NO_RESAMPLING: they restart with the same total */
#define MQ_NEG_STREAMS 4
#define MQ_NS_POSITIVES 0
#define MQ_NS_NEGATIVES 1
#define MQ_NS_SERVICES 2
#define MQ_NS_CHOICES 3 /* uniforms: RCS removes the floor(u * busy)-th busy server in index order */

typedef struct {
    int32_t c;               /* num_of_channels (1..32) */
    int32_t p_bins;
    int64_t buffer;          /* < 0: infinite */
    int32_t type;            /* MQ_NEG_DISASTER or MQ_NEG_RCS */
    int32_t queue_cap;       /* infinite buffer: waiting-line capacity of the engine (0: 1024) */
    int32_t scenario;        /* MQ_NEG_REMOVE (0 means the same) .. MQ_NEG_REQUEUE_FIXED */
    int32_t reserved;
    mq_dist positive, negative, service;
    int64_t total_served;
    int64_t replications;
    int64_t rep0;
    uint64_t seed;
} mq_neg_cfg;

/* per-replication record rows */
#define MQ_N_TTEK 0
#define MQ_N_FLAGS 1
#define MQ_N_POS_ARRIVED 2
#define MQ_N_NEG_ARRIVED 3
#define MQ_N_TAKED 4
#define MQ_N_SERVED 5
#define MQ_N_BROKEN 6
#define MQ_N_TOTAL 7
#define MQ_N_DROPPED 8
#define MQ_N_INSYS 9
#define MQ_N_QUEUED 10
#define MQ_N_POVER 11
#define MQ_N_WSUM 12   /* 4: waits                                   */
#define MQ_N_VSUM 16   /* 4: sojourn of all jobs (served and broken) */
#define MQ_N_VSSUM 20  /* 4: served jobs only                        */
#define MQ_N_VBSUM 24  /* 4: broken jobs only                        */
#define MQ_N_WRUN 28   /* 4 x 4: the reference's running moments in the same order */
#define MQ_N_VRUN 32
#define MQ_N_VSRUN 36
#define MQ_N_VBRUN 40
#define MQ_N_USED 44   /* 4: draws consumed per stream */
#define MQ_N_P 48
#define MQ_N_ROWS(pb) (48 + (pb))

/* aggregate: [0] replications, [1] sum ttek, [2] flagged, [3] sum served, [4] sum broken, [5] sum total, [6] sum
 * positive arrivals, [7] sum dropped; [8..11] w (sum over replications of w_sum/taked), [12..15] squares; [16..19] v
 * (/total), [20..23] squares; [24..27] v_served (/served), [28..31] squares; [32..35] v_broken (/broken), [36..39]
 * squares; [40] q = served/total, [41] its square; [48..48+pb) p = time in state / ttek, then pb squares */
#define MQ_N_AGG_LEN(pb) (48 + 2 * (pb))

int mq_sim_negatives(mq_ctx *ctx, const mq_neg_cfg *cfg, double *agg, double *rep_out);
/* replay form: X[s] = draws of stream s ([n[s]][replications]; choices may be NULL for DISASTER) */
int mq_replay_negatives(mq_ctx *ctx, const mq_neg_cfg *cfg, const double *const *X, const int64_t *n, double *agg,
                        double *rep_out);

/* ---- size-based single-server disciplines.  Replaces SizeBasedQsSim(1, discipline, buffer, track_slowdown)
 * .set_sources / set_servers / set_predictor / run(total_served) (sim/size_based.py:64,105,121,216).  The predictors
 * of sim/utils/predictor.py are built in: perfect (the default), exponential noise, log-normal noise */
#define MQ_SZ_FCFS 0
#define MQ_SZ_SJF 1
#define MQ_SZ_PSJF 2
#define MQ_SZ_SRPT 3
#define MQ_SZ_SPJF 4   /* ranks by the predicted size                                  size_based.py:141 */
#define MQ_SZ_PSPJF 5  /* the same, pre-emptive                                                          */
#define MQ_SZ_SPRPT 6  /* ranks by the predicted remaining work max(0, Y - served)     size_based.py:143-151 */
#define MQ_PRED_PERFECT 0    /* Y = X                                   size_based.py:33-38            */
#define MQ_PRED_EXP 1        /* Y | X = x ~ Exp(mean x)                 sim/utils/predictor.py:14-24   */
#define MQ_PRED_LOGNORMAL 2  /* Y | X = x ~ LogNormal(log x, sigma)     sim/utils/predictor.py:27-43   */
#define MQ_PRED_GIVEN 3      /* replay form: Y supplied by the caller                                    */

typedef struct {
    int32_t discipline;
    int32_t p_bins;
    int64_t buffer;      /* < 0: infinite */
    mq_dist src, srv;    /* interarrival law; job-size law (= service law, size_based.py:121-124) */
    int64_t total_served;
    int64_t replications;
    int64_t rep0;
    uint64_t seed;
    int32_t queue_cap;   /* infinite buffer: waiting-line capacity of the engine (0: 256) */
    int32_t pred_kind;   /* MQ_PRED_* (generator form; the replay form uses Y when it is given) */
    double pred_sigma;   /* MQ_PRED_LOGNORMAL */
} mq_size_cfg;

/* per-replication record rows */
#define MQ_Z_TTEK 0
#define MQ_Z_FLAGS 1
#define MQ_Z_ARRIVED 2
#define MQ_Z_TAKED 3
#define MQ_Z_SERVED 4
#define MQ_Z_DROPPED 5
#define MQ_Z_INSYS 6
#define MQ_Z_QUEUED 7
#define MQ_Z_BUSYN 8
#define MQ_Z_POVER 9
#define MQ_Z_PREEMPTIONS 10
#define MQ_Z_AUSED 11
#define MQ_Z_SUSED 12
#define MQ_Z_YUSED 13
#define MQ_Z_SLOWDOWN 14 /* 2: sum and sum of squares of sojourn / size (size_based.py:110-119) */
#define MQ_Z_WSUM 16     /* 4 */
#define MQ_Z_VSUM 20     /* 4 */
#define MQ_Z_WRUN 24     /* 4 */
#define MQ_Z_VRUN 28     /* 4 */
#define MQ_Z_BUSYRUN 32  /* 3 */
#define MQ_Z_P 36
#define MQ_Z_ROWS(pb) (36 + (pb))

/* aggregate: [0] replications, [1] sum ttek, [2] flagged, [3] sum served, [4] sum arrived, [5] sum dropped, [6] sum
 * taked, [7] sum preemptions; [8..11] w (w_sum/taked), [12..15] squares; [16..19] v (v_sum/served), [20..23] squares;
 * [24] mean slowdown (slowdown_sum/served), [25] its square; [32..32+pb) p, then pb squares */
#define MQ_Z_AGG_LEN(pb) (32 + 2 * (pb))

int mq_sim_size(mq_ctx *ctx, const mq_size_cfg *cfg, double *agg, double *rep_out);
/* replay form: A = gaps [n_a][replications], S = job sizes in arrival order (FCFS: services in start order), Y =
 * predicted sizes in arrival order (NULL: perfect predictor) */
int mq_replay_size(mq_ctx *ctx, const mq_size_cfg *cfg, const double *A, int64_t n_a, const double *S, int64_t n_s,
                   const double *Y, int64_t n_y, double *agg, double *rep_out);

int mq_version(void);
int mq_device_count(void);
const char *mq_last_error(void);

int mq_create(mq_ctx **out, int device_id);
int mq_destroy(mq_ctx *ctx);

/* agg: MQ_AGG_LEN(p_bins) doubles (host).  rep_out: NULL, or MQ_REC_ROWS(p_bins) *
 * replications doubles (host), [field][rep]. */
int mq_sim_fifo(mq_ctx *ctx, const mq_fifo_cfg *cfg, double *agg, double *rep_out);

/* rep_out: MQ_REPLAY_ROWS(p_bins) * replications doubles, [field][rep] (host, or device
 * when cfg->device_ptrs).  agg may be NULL. */
int mq_replay_fifo(mq_ctx *ctx, const mq_replay_cfg *cfg, const double *A, const double *S,
                   double *agg, double *rep_out);

/* device time of the kernels of the last call on this context, in milliseconds
 * (CUDA events on the context's stream): simulation kernel(s), reduction kernel, and the
 * whole call including copies. */
int mq_last_timing(mq_ctx *ctx, double *sim_ms, double *reduce_ms, double *total_ms);
/* number of kernels the last call launched */
int mq_last_launches(mq_ctx *ctx);

/* generator building blocks, exported for the known-answer tests */
int mq_debug_philox2x32(mq_ctx *ctx, uint32_t c0, uint32_t c1, uint32_t key, uint32_t out[2]);
/* n variates of `d` for arrival indices 0..n-1 of replication `rep` (which: 0 source word,
 * 1 server word), evaluated by the device sampler */
int mq_debug_variates(mq_ctx *ctx, const mq_dist *d, uint64_t seed, uint32_t rep, int which,
                      int64_t n, float *out);

#ifdef __cplusplus
}
#endif
#endif
