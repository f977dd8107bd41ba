// k1_fifo_replay.cuh -- K1: replay GI/G/c/r FIFO kernel (fp64, bit-exact tier).
//
// One lane = one replication of QsSim.run() (most_queue/sim/base.py:488-528) fed from
// caller arrays instead of a generator: A[i][rep] is the i-th interarrival draw
// (base.py:112,236), S[j][rep] the j-th service draw (sim/utils/servers.py:88).  Layout is
// replication-minor, so the 32 lanes of a warp read 256 contiguous bytes per row.
//
// Every fp64 operation is issued in the reference's order with explicit round-to-nearest
// intrinsics (no FMA contraction), so the outputs -- per-replication running-mean moments
// (sim/utils/stats_update.py:6-22), time-in-state sums (base.py:327-340), clock, counters,
// busy-period moments (base.py:276-281) -- equal the reference's bit for bit.
//
// Server slots are NOT sorted here: the reference scans slots with a strict '<'
// (base.py:312-325), so the lowest index wins ties, and the order of v samples at tied
// departures depends on it.
//
// Infinite buffer: the waiting jobs' arrival times are not stored; a lagging cursor re-reads
// A (the rows are L2-resident a few MB behind the lead) and rebuilds the same running sum
// a_i = a_{i-1} + A[i], which is bit-identical by construction.  Finite buffer r: arrival
// times of waiting jobs in an HBM ring [slot][rep].
#pragma once

#include "mq_device.cuh"

namespace mq {

constexpr int MQ_K1_BLOCK = 128;
constexpr int MQ_K1_BF = 32;  // fp64 histogram states in shared memory

struct K1Params {
    int c;
    int p_bins;
    long long buffer;
    long long jobs;
    long long reps;
    long long n_a, n_s;
    const double *A;  // [n_a][reps]
    const double *S;  // [n_s][reps]
    double *rec;      // [MQ_REPLAY_ROWS(p_bins)][reps], zeroed by the caller
    double *ring;     // finite buffer: [buffer][reps]
    const double2 *inv;  // inv[k] = {1/k, 1 - 1/k} (round-to-nearest), k = 0..n_inv-1; shared by all lanes
    long long n_inv;
};

// refresh_moments_stat, stats_update.py:6-22, op for op
template <int NM>
__device__ __forceinline__ void refresh_moments(double (&m)[NM], double x, long long count, const K1Params &P)
{
    double power = x;
    double inv_count, one_minus;
    if (count < P.n_inv) {  // 1/count and 1 - 1/count from the shared table (same roundings)
        const double2 t = __ldg(P.inv + count);
        inv_count = t.x;
        one_minus = t.y;
    } else {
        inv_count = __ddiv_rn(1.0, (double)count);
        one_minus = __dsub_rn(1.0, inv_count);
    }
#pragma unroll
    for (int i = 0; i < NM; ++i) {
        m[i] = __dadd_rn(__dmul_rn(m[i], one_minus), __dmul_rn(power, inv_count));
        power = __dmul_rn(power, x);
    }
}

__device__ __forceinline__ void add_powers(double (&s)[4], double x)
{
    double pw = x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s[i] = __dadd_rn(s[i], pw);
        pw = __dmul_rn(pw, x);
    }
}

template <int CT, bool FINITE>
__global__ void __launch_bounds__(MQ_K1_BLOCK) k1_fifo_replay(const K1Params P)
{
    __shared__ double hist[MQ_K1_BF * MQ_K1_BLOCK];  // [state][lane]
    const int tid = threadIdx.x;
    const long long rep = (long long)blockIdx.x * MQ_K1_BLOCK + tid;
    const long long R = P.reps;
    if (rep >= R) return;
    const int c = P.c;
    const long long N = P.jobs;
    const long long cap = FINITE ? P.buffer : 0;
    const double *A = P.A + rep;
    const double *S = P.S + rep;
    double *rec = P.rec + rep;
    double *ring = FINITE ? P.ring + rep : nullptr;
    const long long extra = (long long)MQ_REC_ROWS(P.p_bins);

#pragma unroll
    for (int b = 0; b < MQ_K1_BF; ++b) hist[b * MQ_K1_BLOCK + tid] = 0.0;

    double end[CT], arr[CT];
#pragma unroll
    for (int j = 0; j < CT; ++j) {
        end[j] = j < c ? 1e10 : 1e300;  // servers.py:35 ; slots >= c never win the scan
        arr[j] = 0.0;
    }
    uint32_t busy = 0;

    double ttek = 0.0, start_busy = 0.0;
    double w_run[4] = {0, 0, 0, 0}, v_run[4] = {0, 0, 0, 0}, busy_run[3] = {0, 0, 0};
    double w_sum[4] = {0, 0, 0, 0}, v_sum[4] = {0, 0, 0, 0};
    long long arrived = 0, taked = 0, served = 0, dropped = 0, busy_n = 0;
    long long in_sys = 0, q = 0;
    int free_channels = c;
    long long ia = 1, is = 0;
    // infinite: lag cursor (index and arrival time of the queue head); finite: ring head
    long long h = 0;
    double a_h = 0.0;
    long long head = 0;
    double p_over = 0.0;
    int status = 0;

    double arrival_time = P.n_a > 0 ? A[0] : 0.0;  // base.py:112
    if (P.n_a < 1) status = MQ_E_INPUT_SHORT;
    // the next draw of each stream is loaded one event ahead of its use (hides the HBM latency)
    double a_next = P.n_a > 1 ? A[R] : 0.0;  // default caching: the lag cursor re-reads these rows from L2
    double s_next = P.n_s > 0 ? __ldcs(S) : 0.0;

    while (served < N && status == 0) {
        // _get_min_server_time base.py:312-325
        int midx = -1;
        double mt = 1e16;
#pragma unroll
        for (int j = 0; j < CT; ++j) {
            if (end[j] < mt) {
                mt = end[j];
                midx = j;
            }
        }

        if (arrival_time <= mt) {
            // arrival() base.py:214-247
            arrived++;
            const double dt = __dsub_rn(arrival_time, ttek);
            if (in_sys < MQ_K1_BF) {
                double *hp = &hist[(int)in_sys * MQ_K1_BLOCK + tid];
                *hp = __dadd_rn(*hp, dt);
            } else if (in_sys < P.p_bins) {
                double *gp = &rec[(MQ_F_P + in_sys) * R];
                *gp = __dadd_rn(*gp, dt);
            } else {
                p_over = __dadd_rn(p_over, dt);
            }
            ttek = arrival_time;
            const long long this_idx = ia - 1;
            if (ia >= P.n_a) {
                status = MQ_E_INPUT_SHORT;
                break;
            }
            arrival_time = __dadd_rn(ttek, a_next);
            ia++;
            a_next = ia < P.n_a ? A[ia * R] : 0.0;
            in_sys++;
            if (free_channels == 0) {
                // send_task_to_queue base.py:186-212
                if (!FINITE) {
                    if (q == 0) {
                        h = this_idx;
                        a_h = ttek;
                    }
                    q++;
                } else if (q < cap) {
                    long long slot = head + q;
                    if (slot >= cap) slot -= cap;
                    ring[slot * R] = ttek;
                    q++;
                } else {
                    dropped++;
                    in_sys--;
                }
            } else {
                // send_task_to_channel base.py:155-184 (lowest free slot)
                const int j = __ffs(~busy & ((c >= 32) ? 0xffffffffu : ((1u << c) - 1u))) - 1;
                taked++;
                refresh_moments<4>(w_run, 0.0, taked, P);
                add_powers(w_sum, 0.0);
                if (is >= P.n_s) {
                    status = MQ_E_INPUT_SHORT;
                    break;
                }
                const double e = __dadd_rn(ttek, s_next);  // servers.py:88
                is++;
                s_next = is < P.n_s ? __ldcs(S + is * R) : 0.0;
#pragma unroll
                for (int k = 0; k < CT; ++k) {
                    if (k == j) {
                        end[k] = e;
                        arr[k] = ttek;
                    }
                }
                busy |= 1u << j;
                free_channels--;
                if (free_channels == 0 && in_sys == c) start_busy = ttek;
            }
        } else {
            // serving() base.py:249-286
            const double time_to_end = mt;
            double a_dep = 0.0;
#pragma unroll
            for (int k = 0; k < CT; ++k) {
                if (k == midx) {
                    a_dep = arr[k];
                    end[k] = 1e10;
                }
            }
            busy &= ~(1u << midx);
            const double dt = __dsub_rn(time_to_end, ttek);
            if (in_sys < MQ_K1_BF) {
                double *hp = &hist[(int)in_sys * MQ_K1_BLOCK + tid];
                *hp = __dadd_rn(*hp, dt);
            } else if (in_sys < P.p_bins) {
                double *gp = &rec[(MQ_F_P + in_sys) * R];
                *gp = __dadd_rn(*gp, dt);
            } else {
                p_over = __dadd_rn(p_over, dt);
            }
            ttek = time_to_end;
            served++;
            free_channels++;
            const double v = __dsub_rn(ttek, a_dep);
            refresh_moments<4>(v_run, v, served, P);
            add_powers(v_sum, v);
            in_sys--;
            if (q == 0 && free_channels == 1 && in_sys == c - 1) {
                busy_n++;
                refresh_moments<3>(busy_run, __dsub_rn(ttek, start_busy), busy_n, P);
            }
            if (q != 0) {
                // send_head_of_queue_to_channel base.py:288-310
                double a;
                if (!FINITE) {
                    a = a_h;
                    if (q > 1) {
                        h++;
                        a_h = __dadd_rn(a_h, A[h * R]);
                    }
                } else {
                    a = ring[head * R];
                    head++;
                    if (head >= cap) head = 0;
                }
                q--;
                if (free_channels == 1) start_busy = ttek;
                taked++;
                const double w = __dadd_rn(0.0, __dsub_rn(ttek, a));
                refresh_moments<4>(w_run, w, taked, P);
                add_powers(w_sum, w);
                if (is >= P.n_s) {
                    status = MQ_E_INPUT_SHORT;
                    break;
                }
                const double e = __dadd_rn(ttek, s_next);
                is++;
                s_next = is < P.n_s ? __ldcs(S + is * R) : 0.0;
#pragma unroll
                for (int k = 0; k < CT; ++k) {
                    if (k == midx) {
                        end[k] = e;
                        arr[k] = a;
                    }
                }
                busy |= 1u << midx;
                free_channels--;
            }
        }
    }

#pragma unroll
    for (int k = 0; k < 4; ++k) {
        rec[(MQ_F_W + k) * R] = w_sum[k];
        rec[(MQ_F_V + k) * R] = v_sum[k];
        rec[(extra + MQ_RF_WRUN + k) * R] = w_run[k];
        rec[(extra + MQ_RF_VRUN + k) * R] = v_run[k];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) rec[(extra + MQ_RF_BUSYRUN + k) * R] = busy_run[k];
    rec[(extra + MQ_RF_BUSYN) * R] = (double)busy_n;
    rec[(extra + MQ_RF_INSYS) * R] = (double)in_sys;
    rec[(extra + MQ_RF_QUEUED) * R] = (double)q;
    rec[(extra + MQ_RF_AUSED) * R] = (double)ia;
    rec[(extra + MQ_RF_SUSED) * R] = (double)is;
    rec[MQ_F_ARRIVED * R] = (double)arrived;
    rec[MQ_F_TAKED * R] = (double)taked;
    rec[MQ_F_SERVED * R] = (double)served;
    rec[MQ_F_DROPPED * R] = (double)dropped;
    rec[MQ_F_TTEK * R] = ttek;
    rec[MQ_F_FLAGS * R] = (double)(status != 0 ? -status : 0);
    rec[MQ_F_POVER * R] = p_over;
    const int nb = P.p_bins < MQ_K1_BF ? P.p_bins : MQ_K1_BF;
    for (int b = 0; b < nb; ++b) rec[(MQ_F_P + b) * R] = hist[b * MQ_K1_BLOCK + tid];
    if (P.p_bins < MQ_K1_BF) {
        // states beyond p_bins that were binned on chip: fold into p_over in state order
        double over = p_over;
        for (int b = P.p_bins; b < MQ_K1_BF; ++b) over = __dadd_rn(over, hist[b * MQ_K1_BLOCK + tid]);
        rec[MQ_F_POVER * R] = over;
    }
}

}  // namespace mq
