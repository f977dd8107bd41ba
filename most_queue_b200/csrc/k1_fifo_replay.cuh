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
// v2 structure (v1 had one code arm per event kind and ran 12.8 of 32 lanes per instruction): ONE
// predicated body per loop trip.  A trip is an arrival, a departure, or the second half of a departure
// (start of the queue head / end of a busy period, which the reference does inside serving(),
// base.py:249-310).  Each trip has at most one running-moment sample, so there is a single
// refresh site and a single "write a server slot" site; the moments, power sums, the sojourn
// time of the job on each server and the time-in-state histogram live in shared memory
// ([slot][lane], conflict-free for any per-lane slot), the completion times in registers.
//
// Infinite buffer: the waiting jobs' arrival times are not stored; a lagging cursor re-reads
// A (the rows are L2-resident a few MB behind the lead) and rebuilds the same running sum
// a_i = a_{i-1} + A[i], which is bit-identical by construction.  Finite buffer r: arrival
// times of waiting jobs in an HBM ring [slot][rep].
#pragma once

#include "mq_device.cuh"

namespace mq {

constexpr int MQ_K1_BLOCK = 128;
constexpr int MQ_K1_BF = 16;   // fp64 time-in-state bins in shared memory (further states go to the record)
constexpr int MQ_K1_MOM = 20;  // W: run[4] sum[4], V: run[4] sum[4], busy: run[4]

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
};

constexpr size_t k1_smem_bytes(int ct) { return sizeof(double) * MQ_K1_BLOCK * (size_t)(MQ_K1_BF + MQ_K1_MOM + ct); }

template <int CT, bool FINITE>
__global__ void __launch_bounds__(MQ_K1_BLOCK) k1_fifo_replay(const K1Params P)
{
    extern __shared__ double k1_sm[];
    const int tid = threadIdx.x;
    double *hist = k1_sm + tid;                                        // [MQ_K1_BF][lane]
    double *mom = k1_sm + MQ_K1_BF * MQ_K1_BLOCK + tid;                // [MQ_K1_MOM][lane]
    double *vv = k1_sm + (MQ_K1_BF + MQ_K1_MOM) * MQ_K1_BLOCK + tid;   // [CT][lane]: sojourn time of the job on slot j
    const long long rep = (long long)blockIdx.x * MQ_K1_BLOCK + tid;
    const long long R = P.reps;
    if (rep >= R) return;
    const int c = P.c;
    // counters and cursors are 32 bit (the C ABI rejects jobs, n_a, n_s or buffer >= 2^31 for this kernel)
    const uint32_t N = (uint32_t)P.jobs;
    const uint32_t cap = FINITE ? (uint32_t)P.buffer : 0u;
    const uint32_t n_a = (uint32_t)P.n_a, n_s = (uint32_t)P.n_s;
    const uint32_t p_bins = (uint32_t)P.p_bins;
    const double *A = P.A + rep;
    const double *S = P.S + rep;
    double *rec = P.rec + rep;
    double *ring = FINITE ? P.ring + rep : nullptr;
    const long long extra = (long long)MQ_REC_ROWS(P.p_bins);

#pragma unroll
    for (int b = 0; b < MQ_K1_BF + MQ_K1_MOM; ++b) k1_sm[b * MQ_K1_BLOCK + tid] = 0.0;

    double end[CT];
#pragma unroll
    for (int j = 0; j < CT; ++j) end[j] = j < c ? 1e10 : 1e300;  // servers.py:35 ; slots >= c never win the scan
    uint32_t busy = 0;
    const uint32_t cmask = (c >= 32) ? 0xffffffffu : ((1u << c) - 1u);

    double ttek = 0.0, start_busy = 0.0;
    uint32_t arrived = 0, taked = 0, served = 0, dropped = 0, busy_n = 0;
    uint32_t in_sys = 0, q = 0;
    int free_channels = c;
    uint32_t ia = 1, is = 0;
    // infinite: lag cursor (index and arrival time of the queue head); finite: ring head
    uint32_t h = 0;
    double a_h = 0.0, a_lag = 0.0;  // a_lag = A[h + 1], loaded one trip before the cursor moves
    uint32_t head = 0;
    double p_over = 0.0;
    int status = 0;
    int pend = 0;   // second half of a departure: 2 = start the head of the queue, 3 = a busy period ended
    int pslot = 0;  // the slot the pending start goes to

    double arrival_time = n_a > 0 ? A[0] : 0.0;  // base.py:112
    if (n_a < 1) status = MQ_E_INPUT_SHORT;
    // the next draw of each stream is loaded one trip ahead of its use (hides the HBM latency)
    double a_next = n_a > 1 ? A[R] : 0.0;  // default caching: the lag cursor re-reads these rows from L2
    double s_next = n_s > 0 ? S[0] : 0.0;

    while (status == 0 && (pend != 0 || served < N)) {
        // _get_min_server_time base.py:312-325: strict '<', lowest index wins ties (pairwise tree, left operand
        // kept on ties)
        double mv[CT];
        int mi[CT];
#pragma unroll
        for (int j = 0; j < CT; ++j) {
            mv[j] = end[j];
            mi[j] = j;
        }
#pragma unroll
        for (int st = 1; st < CT; st *= 2) {
#pragma unroll
            for (int j = 0; j + st < CT; j += 2 * st) {
                const bool lt = mv[j + st] < mv[j];
                mv[j] = lt ? mv[j + st] : mv[j];
                mi[j] = lt ? mi[j + st] : mi[j];
            }
        }
        const double mt = mv[0];
        const int midx = mi[0];

        const bool pending = pend != 0;
        const bool is_arr = !pending && arrival_time <= mt;
        const bool is_dep = !pending && !is_arr;
        const double te = pending ? ttek : (is_arr ? arrival_time : mt);
        {
            // _update_state_probs base.py:327-340 (a pending trip adds +0.0: no change)
            const double dt = __dsub_rn(te, ttek);
            if (in_sys < (uint32_t)MQ_K1_BF) {
                double *hp = hist + in_sys * MQ_K1_BLOCK;
                *hp = __dadd_rn(*hp, dt);
            } else if (in_sys < p_bins) {
                // far states accumulate in the (zeroed) record with a fire-and-forget fp64 reduction: same
                // round-to-nearest add, same order (one lane, one address), and no load to wait for
                atomicAdd(&rec[(long long)(MQ_F_P + in_sys) * R], dt);
            } else {
                p_over = __dadd_rn(p_over, dt);
            }
        }
        ttek = te;

        // ---- straight-line, predicated bookkeeping of the four trip kinds (no divergent arms) ----
        const bool is_pstart = pend == 2;  // send_head_of_queue_to_channel base.py:288-310
        const bool is_pbusy = pend == 3;   // a busy period is over, base.py:276-281
        // arrival() base.py:214-247
        arrived += is_arr;
        const bool short_a = is_arr && ia >= n_a;
        arrival_time = is_arr ? __dadd_rn(ttek, a_next) : arrival_time;
        ia += is_arr;
        if (is_arr) a_next = ia < n_a ? A[(long long)ia * R] : 0.0;
        in_sys += is_arr;
        const bool full = free_channels == 0;
        const bool arr_start = is_arr && !full;  // send_task_to_channel base.py:155-184
        const bool arr_q = is_arr && full;       // send_task_to_queue base.py:186-212
        // serving() base.py:249-286, first half
        const double v_dep = vv[midx * MQ_K1_BLOCK];
        busy = is_dep ? (busy & ~(1u << midx)) : busy;
        served += is_dep;
        free_channels += is_dep;
        in_sys -= is_dep;
        const bool dep_q = is_dep && q != 0;
        const bool dep_nq = is_dep && q == 0;
        const bool busy_over = dep_nq && free_channels == 1 && in_sys == (uint32_t)(c - 1);
        pslot = dep_q ? midx : pslot;
        // queue
        double a_head;
        if (!FINITE) {
            const bool first = arr_q && q == 0;
            h = first ? ia - 2 : h;
            a_h = first ? ttek : a_h;
            a_head = a_h;
            const bool more = is_pstart && q > 1;
            h += more;
            a_h = more ? __dadd_rn(a_h, a_lag) : a_h;
            // h + 1 <= ia - 1 < n_a while somebody waits; on the input-short path (ia >= n_a, status set at the end of the
            // trip) the row does not exist: the load is skipped instead of reading one row past the caller's array
            if ((first || more) && h + 1u < n_a) a_lag = A[(long long)(h + 1u) * R];
            q += (uint32_t)arr_q - (uint32_t)is_pstart;
        } else {
            const bool accept = arr_q && q < cap;
            const bool drop = arr_q && !accept;
            uint32_t sl = head + q;
            sl = sl >= cap ? sl - cap : sl;
            if (accept) ring[(long long)sl * R] = ttek;
            a_head = 0.0;
            if (is_pstart) a_head = ring[(long long)head * R];
            head += is_pstart;
            head = head >= cap ? 0u : head;
            q += (uint32_t)accept - (uint32_t)is_pstart;
            dropped += drop;
            in_sys -= drop;
        }
        start_busy = (is_pstart && free_channels == 1) ? ttek : start_busy;
        busy_n += is_pbusy;
        // a job starts service: the arriving one or the head of the queue
        const bool start = arr_start || is_pstart;
        const int slot = is_arr ? __ffs(~busy & cmask) - 1 : pslot;
        const double a_job = is_arr ? ttek : a_head;
        taked += start;
        const bool short_s = start && is >= n_s;
        const double e_new = __dadd_rn(ttek, s_next);  // servers.py:88
        is += start;
        if (start) {
            s_next = is < n_s ? S[(long long)is * R] : 0.0;
            vv[slot * MQ_K1_BLOCK] = __dsub_rn(e_new, a_job);  // what serving() will compute as ttek - arrival
        }
        busy = start ? (busy | (1u << slot)) : busy;
        free_channels -= start;
        start_busy = (arr_start && free_channels == 0 && in_sys == (uint32_t)c) ? ttek : start_busy;
        // the one server slot rewritten in this trip, the one running-moment sample of this trip
        const int wslot = start ? slot : (dep_nq ? midx : -1);
        const double wval = start ? e_new : 1e10;
        const int which = start ? 0 : (is_dep ? 1 : (is_pbusy ? 2 : -1));  // 0 = w, 1 = v, 2 = busy period
        const double x = start ? (is_arr ? 0.0 : __dadd_rn(0.0, __dsub_rn(ttek, a_job)))
                               : (is_dep ? v_dep : __dsub_rn(ttek, start_busy));
        const uint32_t count = start ? taked : (is_dep ? served : busy_n);
        pend = dep_q ? 2 : (busy_over ? 3 : 0);
        if (short_a || short_s) status = MQ_E_INPUT_SHORT;
#pragma unroll
        for (int k = 0; k < CT; ++k)
            if (k == wslot) end[k] = wval;

        if (which >= 0) {
            // refresh_moments_stat, stats_update.py:6-22, op for op (the busy period keeps 3 moments in the
            // reference; the 4th computed here is never read), plus the power sums the aggregate uses
            const double inv_count = __drcp_rn((double)count);  // == 1.0 / count, correctly rounded
            const double one_minus = __dsub_rn(1.0, inv_count);
            double *mp = mom + which * (8 * MQ_K1_BLOCK);
            const bool sums = which < 2;
            double power = x;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double *m = mp + i * MQ_K1_BLOCK;
                *m = __dadd_rn(__dmul_rn(*m, one_minus), __dmul_rn(power, inv_count));
                if (sums) {
                    double *sp = mp + (4 + i) * MQ_K1_BLOCK;
                    *sp = __dadd_rn(*sp, power);
                }
                power = __dmul_rn(power, x);
            }
        }
    }

#pragma unroll
    for (int k = 0; k < 4; ++k) {
        rec[(MQ_F_W + k) * R] = mom[(4 + k) * MQ_K1_BLOCK];
        rec[(MQ_F_V + k) * R] = mom[(12 + k) * MQ_K1_BLOCK];
        rec[(extra + MQ_RF_WRUN + k) * R] = mom[k * MQ_K1_BLOCK];
        rec[(extra + MQ_RF_VRUN + k) * R] = mom[(8 + k) * MQ_K1_BLOCK];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) rec[(extra + MQ_RF_BUSYRUN + k) * R] = mom[(16 + k) * MQ_K1_BLOCK];
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
    // states binned on chip; the caller zeroed the record, so states >= MQ_K1_BF accumulated there directly
    const int nb = P.p_bins < MQ_K1_BF ? P.p_bins : MQ_K1_BF;
    for (int b = 0; b < nb; ++b) rec[(MQ_F_P + b) * R] = hist[b * MQ_K1_BLOCK];
    if (P.p_bins < MQ_K1_BF) {
        // states beyond p_bins that were binned on chip: fold into p_over in state order
        double over = p_over;
        for (int b = P.p_bins; b < MQ_K1_BF; ++b) over = __dadd_rn(over, hist[b * MQ_K1_BLOCK]);
        rec[MQ_F_POVER * R] = over;
    }
}

}  // namespace mq
