// k2_fifo_gen.cuh -- K2: generate-and-simulate GI/G/c/r FIFO kernel.
//
// One lane = one independent replication of QsSim.run() (most_queue/sim/base.py:488-528).
// The lane walks the same event sequence as the reference (arrival wins ties, base.py:442-454;
// stop right after the total_served-th departure, base.py:505) and produces the same sample
// sets: w for every job that starts service (base.py:172-173,301-305), v for every served job
// (base.py:269-272), p[j] = time with j jobs in the system (base.py:327-340), finite-buffer
// drops (base.py:204-212).
//
// Layout of one lane's state (all on chip):
//   * clock, next arrival, c completion times: fp32 registers, RELATIVE to an epoch that is
//     re-based every MQ_EV events so magnitudes stay small; the elapsed epochs are summed in fp64.
//   * the c completion times are kept sorted (min/max insertion network) in Kiefer-Wolfowitz style:
//     a key <= now means "this server is free since then".  The low TB bits of each key carry the
//     id of the server slot, so keys are all distinct and the slot survives sorting.
//   * LAZY DEPARTURES: while nobody waits (q = 0) a departure changes nothing but the number in the
//     system, so it is not an event: at the next arrival the time-in-state of the interval is settled
//     for all servers at once from the sorted keys (state c-j lasts from key j-1 to key j, clamped
//     to the interval; c+1 register accumulators with static indices).  Only while jobs wait (q > 0)
//     are departures events (they start the head of the queue).  That is ~1 + P(wait) iterations per
//     job instead of 2.  From the total_served-th START on, an exact one-event-at-a-time end game
//     applies the reference's stop rule (stop right after the total_served-th departure).
//   * v of the job in service on each slot: shared memory [slot][lane] (bank = lane, conflict free).
//     v is accumulated when service STARTS and the <= c jobs still in service at the stop are
//     subtracted at the end, which yields exactly the reference's sample set.
//   * time-in-state: states 0..c in registers (see above), states c+1..c+MQ_BF in shared memory
//     [state][lane] (fp32, flushed to the fp64 per-replication record every MQ_FLUSH epochs), higher
//     states go straight to the fp64 record in HBM (rare).
//   * infinite buffer: the queue is NOT stored.  A lagging Philox cursor regenerates the arrival
//     time and service time of the job at the head of the queue (same words, same summation), so the
//     queue is unbounded with O(1) state.
//   * finite buffer r: (gap to previous queued arrival, service time) per waiting job in a ring
//     of r entries per lane in HBM scratch [slot][lane].
#pragma once

#include "mq_device.cuh"

namespace mq {

constexpr int MQ_BLOCK = 128;    // lanes per CTA
constexpr int MQ_BF = 32;        // histogram states kept in shared memory
constexpr int MQ_EV = 32;        // events between epoch re-bases
constexpr int MQ_FLUSH = 64;     // re-bases between fp32 -> fp64 flushes
// Arithmetic type of the kernel.  float is the production tier (fp32 clocks relative to an epoch, MUFU samplers);
// double is the A/B tier (MQ_FIFO_FP64): the same loop, the same Philox words, fp64 clocks, fp64 libm-grade samplers and
// fp64 power sums -- the precision the reference's Python floats have.  It exists to MEASURE what fp32 costs.
template <typename Real>
struct RT;

template <>
struct RT<float> {
    using U = uint32_t;
    using Ring = float2;
    static constexpr U KEY_GONE = 0xfe800000u;  // -8.5e37: server free "since ever" (plus slot id)
    static constexpr U KEY_OFF = 0x7e800000u;   // +8.5e37: slot not used (c < CT): always busy, never departs
    static constexpr float LIM = 1.0e30f;       // |key| below this is a real time
    static constexpr float HUGE_T = 3.0e38f;
    static __device__ __forceinline__ U bits(float x) { return __float_as_uint(x); }
    static __device__ __forceinline__ float real(U b) { return __uint_as_float(b); }
    static __device__ __forceinline__ float rcp(float x) { return fast_rcp(x); }
    static __device__ __forceinline__ float sat(float x) { return __saturatef(x); }
    static __device__ __forceinline__ float mn(float a, float b) { return fminf(a, b); }
    static __device__ __forceinline__ float mx(float a, float b) { return fmaxf(a, b); }
    static __device__ __forceinline__ float fma_(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ float abs_(float a) { return fabsf(a); }
    static __device__ __forceinline__ Ring ring(float a, float b) { return make_float2(a, b); }
    static __device__ __forceinline__ float ld(uint32_t addr)
    {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
        return v;
    }
    static __device__ __forceinline__ void st(uint32_t addr, float v)
    {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
    }
};

template <>
struct RT<double> {
    using U = unsigned long long;
    using Ring = double2;
    static constexpr U KEY_GONE = 0xC7D0000000000000ull;  // -2^126
    static constexpr U KEY_OFF = 0x47D0000000000000ull;   // +2^126
    static constexpr double LIM = 1.0e30;
    static constexpr double HUGE_T = 1.0e300;
    static __device__ __forceinline__ U bits(double x) { return (U)__double_as_longlong(x); }
    static __device__ __forceinline__ double real(U b) { return __longlong_as_double((long long)b); }
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double sat(double x) { return fmin(fmax(x, 0.0), 1.0); }
    static __device__ __forceinline__ double mn(double a, double b) { return fmin(a, b); }
    static __device__ __forceinline__ double mx(double a, double b) { return fmax(a, b); }
    static __device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
    static __device__ __forceinline__ double abs_(double a) { return fabs(a); }
    static __device__ __forceinline__ Ring ring(double a, double b) { return make_double2(a, b); }
    static __device__ __forceinline__ double ld(uint32_t addr)
    {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
        return v;
    }
    static __device__ __forceinline__ void st(uint32_t addr, double v)
    {
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
    }
};

struct K2Params {
    int c;
    int p_bins;
    long long buffer;  // -1 infinite
    long long jobs;
    long long reps;
    uint32_t rep0;
    uint32_t key0;
    uint32_t rk[10];   // key0 + r * MQ_PHILOX_W: the Philox round keys of the primary call
    uint32_t warm;     // statistics restart at the warm-th start of service (0: never)
    int v_at_start;    // MQ_FIFO_V_AT_START: keep the v of the jobs still in service at the stop
    DistF src, srv;
    DistD srcd, srvd; // the same laws with fp64 constants (MQ_FIFO_FP64 tier only)
    double *rec;      // [MQ_REC_ROWS(p_bins)][reps], zeroed by the caller
    void *ring;       // finite buffer: [buffer][total lanes] of (gap, service) pairs in the kernel's real type
    long long lanes;  // gridDim.x * MQ_BLOCK
};

template <int CT>
struct Tag {
    static constexpr int TB = CT <= 1 ? 0 : CT <= 2 ? 1 : CT <= 4 ? 2 : CT <= 8 ? 3 : CT <= 16 ? 4 : 5;
    static constexpr uint32_t MASK = (1u << TB) - 1u;
    static constexpr uint32_t HALF = TB ? (1u << (TB - 1)) : 0u;
    template <typename Real>
    static __device__ __forceinline__ uint32_t of(Real x)  // slot id carried by a key
    {
        return (uint32_t)RT<Real>::bits(x) & MASK;
    }
};

// round the key to the tag grid (nearest) and stamp the slot id into its low bits
template <int CT, typename Real>
__device__ __forceinline__ Real stamp(Real x, uint32_t tag)
{
    if (Tag<CT>::TB == 0) return x;
    using U = typename RT<Real>::U;
    const U b = RT<Real>::bits(x);
    return RT<Real>::real(((b + (U)Tag<CT>::HALF) & ~(U)Tag<CT>::MASK) | (U)tag);
}

// L[0..CT-1] sorted ascending.  new L = sorted(L[0..CT-2] + {y}); the old maximum is dropped.
template <int CT, typename Real>
__device__ __forceinline__ void insert_drop_last(Real (&L)[CT], Real y)
{
    Real prev = L[0];
    L[0] = RT<Real>::mn(prev, y);
#pragma unroll
    for (int j = 1; j < CT; ++j) {
        Real cur = L[j];
        Real up = RT<Real>::mx(prev, y);
        L[j] = (j == CT - 1) ? up : RT<Real>::mn(cur, up);
        prev = cur;
    }
    if (CT == 1) L[0] = y;
}

// new L = sorted(L[1..CT-1] + {y}); the old minimum is dropped.
template <int CT, typename Real>
__device__ __forceinline__ void insert_drop_first(Real (&L)[CT], Real y)
{
    if (CT == 1) {
        L[0] = y;
        return;
    }
    Real prev = L[1];
    L[0] = RT<Real>::mn(prev, y);
#pragma unroll
    for (int j = 1; j < CT - 1; ++j) {
        Real cur = L[j + 1];
        L[j] = RT<Real>::mn(cur, RT<Real>::mx(prev, y));
        prev = cur;
    }
    L[CT - 1] = RT<Real>::mx(prev, y);
}

// One wiring for both event kinds: an arrival keeps L[0..CT-2] (the old maximum -- a free slot, or
// y itself when nothing starts -- is dropped), a departure drops L[0]; y is inserted in order.
template <int CT, typename Real>
__device__ __forceinline__ void unified_insert(Real (&L)[CT], Real y, bool is_arr)
{
    if (CT == 1) {
        L[0] = y;
        return;
    }
    Real M[CT > 1 ? CT - 1 : 1];
#pragma unroll
    for (int j = 0; j < CT - 1; ++j) M[j] = is_arr ? L[j] : L[j + 1];
    L[0] = RT<Real>::mn(M[0], y);
#pragma unroll
    for (int j = 1; j < CT - 1; ++j) L[j] = RT<Real>::mn(M[j], RT<Real>::mx(M[j - 1], y));
    L[CT - 1] = RT<Real>::mx(M[CT - 2], y);
}

// After a re-base two keys can land on the same grid point and then differ only by their slot
// ids, which may invert an adjacent pair.  Restore strict order (odd-even transposition until
// clean; almost always the check alone).
template <int CT, typename Real>
__device__ __forceinline__ void restore_order(Real (&L)[CT])
{
    if (CT == 1) return;
    bool bad = false;
#pragma unroll
    for (int j = 0; j + 1 < CT; ++j) bad |= L[j] > L[j + 1];
    while (bad) {
        bad = false;
#pragma unroll
        for (int par = 0; par < 2; ++par) {
#pragma unroll
            for (int j = par; j + 1 < CT; j += 2) {
                const Real lo = RT<Real>::mn(L[j], L[j + 1]), hi = RT<Real>::mx(L[j], L[j + 1]);
                bad |= lo != L[j];
                L[j] = lo;
                L[j + 1] = hi;
            }
        }
    }
}

#ifndef MQ_K2_MINB
#define MQ_K2_MINB (CT == 8 ? 7 : 8)  // measured: 72 registers x 7 CTAs beat 64 x 8 once the round keys sit in uniform registers
#endif
#ifndef MQ_K2_UNROLL
#define MQ_K2_UNROLL 1
#endif
#ifndef MQ_K2_RK
#define MQ_K2_RK 1
#endif
constexpr int K2_UNROLL = MQ_K2_UNROLL;

template <typename Real, int CT>
constexpr int k2_min_blocks()
{
    return sizeof(Real) == 4 ? (CT <= 8 ? MQ_K2_MINB : 2) : (CT <= 8 ? 3 : 1);
}

// BUSY: also gather QsSim.busy, the reference's busy-period moments (base.py:179-184, 274-281, 297-298): the time from
// the last start of service that left no channel free to the first departure that finds the waiting line empty.  Extra
// rows MQ_FB_* behind the record; off by default so that the production loop does not pay for it.
template <typename Real, int CT, int SRCK, int SRVK, bool FINITE, bool BUSY>
__global__ void __launch_bounds__(MQ_BLOCK, (k2_min_blocks<Real, CT>())) k2_fifo_gen(const K2Params P)
{
    using T = RT<Real>;
    using Ring = typename T::Ring;
    extern __shared__ double smem_raw[];
    Real *smem = reinterpret_cast<Real *>(smem_raw);
    const int tid = threadIdx.x;
    Real *hist = smem + tid;                           // [MQ_BF][MQ_BLOCK]: states c+1 .. c+MQ_BF
    Real *vslot = smem + MQ_BF * MQ_BLOCK + tid;       // [CT][MQ_BLOCK]
    // 32-bit shared-window addresses of this lane's columns (keeps the address math out of the loop)
    const uint32_t hist_sa = (uint32_t)__cvta_generic_to_shared(hist);
    const uint32_t vslot_sa = (uint32_t)__cvta_generic_to_shared(vslot);
    constexpr uint32_t COL = MQ_BLOCK * (uint32_t)sizeof(Real);  // bytes between two rows of a [row][lane] array
    const long long gtid = (long long)blockIdx.x * MQ_BLOCK + tid;
    const int c = P.c;
    const uint32_t key0 = P.key0;
    const long long R = P.reps;
    const uint32_t N = (uint32_t)P.jobs;
    const uint32_t cap = FINITE ? (uint32_t)P.buffer : 0u;
    const int off = CT - c;  // register hs[m] holds state m - off
    Ring *ring = reinterpret_cast<Ring *>(P.ring);
    const Real ZERO = (Real)0;

    for (long long rl = gtid; rl < R; rl += P.lanes) {
        const uint32_t rep = P.rep0 + (uint32_t)rl;
        double *rec = P.rec + rl;
        auto draw_a = [&](uint32_t w, uint32_t idx) -> Real {
            if constexpr (sizeof(Real) == 4)
                return sample<SRCK>(P.src, w, idx, rep, key0, 0);
            else
                return sample_d(P.srcd, w, idx, rep, key0, 0);
        };
        auto draw_s = [&](uint32_t w, uint32_t idx) -> Real {
            if constexpr (sizeof(Real) == 4)
                return sample<SRVK>(P.srv, w, idx, rep, key0, 1);
            else
                return sample_d(P.srvd, w, idx, rep, key0, 1);
        };

        // ---- state ----
        Real L[CT];
#pragma unroll
        for (int j = 0; j < CT; ++j) {
            // ascending: the c real slots, free since ever (more negative = larger slot id), then unused
            L[j] = T::real(j < c ? (T::KEY_GONE | (typename T::U)(c - 1 - j)) : (T::KEY_OFF | (typename T::U)j));
        }
        Real hs[CT + 1];
#pragma unroll
        for (int m = 0; m <= CT; ++m) hs[m] = ZERO;
#pragma unroll
        for (int b = 0; b < MQ_BF; ++b) hist[b * MQ_BLOCK] = ZERO;

        Real t = ZERO;
        double tbase = 0.0;
        Real sw[4] = {ZERO, ZERO, ZERO, ZERO}, sv[4] = {ZERO, ZERO, ZERO, ZERO};
        // nxt = index of the arrival scheduled at tA (= number of arrivals so far); q = waiting jobs.
        // Infinite buffer: the waiting jobs are the last q arrivals, so the head of the queue is nxt - q.
        uint32_t nxt = 0, taked = 0, dropped = 0, q = 0;
        Real a_h = ZERO, S_h = ZERO;  // infinite buffer: arrival time and service time of the queue head
        uint32_t head = 0;            // finite buffer: ring bookkeeping
        Real a_anchor = ZERO, a_lastq = ZERO;
        uint32_t flags = 0;
        double t_warm = 0.0;
        bool warm_done = false;
        uint32_t next_stop = (P.warm != 0u && P.warm < N) ? P.warm : N;
        // busy periods (BUSY): power sums of the samples, their number, start of the current one
        Real sb[3] = {ZERO, ZERO, ZERO};
        uint32_t nbusy = 0;
        Real start_busy = ZERO;
        auto busy_sample = [&](Real x) {
            const Real x2 = x * x;
            sb[0] += x;
            sb[1] += x2;
            sb[2] = T::fma_(x2, x, sb[2]);
            nbusy++;
        };

        // first arrival: the draw set_sources makes (base.py:112)
        Real tA, S_pend;
        {
            uint2 w = philox2x32_10(0u, rep, key0);
            tA = draw_a(w.x, 0u);
            S_pend = draw_s(w.y, 0u);
        }

        auto flush = [&]() {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                rec[(long long)(MQ_F_W + k) * R] += (double)sw[k];
                rec[(long long)(MQ_F_V + k) * R] += (double)sv[k];
                sw[k] = ZERO;
                sv[k] = ZERO;
            }
            if (BUSY) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    rec[(long long)(MQ_REC_ROWS(P.p_bins) + MQ_FB_BUSY + k) * R] += (double)sb[k];
                    sb[k] = ZERO;
                }
            }
#pragma unroll
            for (int m = 0; m <= CT; ++m) {
                const int k = m - off;
                if (k >= 0) {
                    rec[(long long)(k < P.p_bins ? MQ_F_P + k : MQ_F_POVER) * R] += (double)hs[m];
                    hs[m] = ZERO;
                }
            }
            for (int b = 0; b < MQ_BF; ++b) {
                const int k = c + 1 + b;
                rec[(long long)(k < P.p_bins ? MQ_F_P + k : MQ_F_POVER) * R] += (double)hist[b * MQ_BLOCK];
                hist[b * MQ_BLOCK] = ZERO;
            }
        };
        auto rebase = [&]() {
            const Real base = t;
            tbase += (double)base;
            t = ZERO;
            tA -= base;
            a_h -= base;
            a_anchor -= base;
            a_lastq -= base;
            if (BUSY) start_busy -= base;
#pragma unroll
            for (int j = 0; j < CT; ++j) {
                const Real moved = stamp<CT, Real>(L[j] - base, Tag<CT>::of(L[j]));
                L[j] = T::abs_(L[j]) < T::LIM ? moved : L[j];
            }
            restore_order<CT, Real>(L);
        };
        // time with c + qq jobs in the system, qq >= 1
        auto add_q_state = [&](uint32_t qq, Real dt) {
            if (qq <= (uint32_t)MQ_BF) {
                const uint32_t ha = hist_sa + (qq - 1u) * COL;
                T::st(ha, T::ld(ha) + dt);
            } else {
                const long long k = (long long)c + qq;
                rec[(k < P.p_bins ? MQ_F_P + k : MQ_F_POVER) * R] += (double)dt;
            }
        };

        // ================= main phase: until the N-th job has STARTED service =================
        int epoch = 0;
        while (taked < N) {
#pragma unroll K2_UNROLL
            for (int ev = 0; ev < MQ_EV; ++ev) {
                const Real d0 = L[0];
                const bool q0 = q == 0u;
                // nobody waits: the next event is the arrival (departures are settled lazily);
                // otherwise arrival or departure, arrival wins ties (base.py:442-454)
                const bool is_arr = q0 || tA <= d0;
                const Real te = is_arr ? tA : d0;
                // _update_state_probs (base.py:327-340)
                const Real dt = te - t;
                {
                    // nobody waits: between t and te the c servers finish in key order, state (c - j) lasts
                    // until key j.  With r_j = saturate((key_j - t) / dt) the time in state (c - j) is
                    // (r_j - r_{j-1}) dt: one FFMA.SAT per key on the FMA pipe instead of two FMNMX clamps.
                    // Lanes with waiting jobs run the same code with a zero weight (no divergent arm).
                    const Real inv = (q0 && dt > ZERO) ? T::rcp(dt) : ZERO;
                    const Real wgt = q0 ? dt : ZERO;
                    const Real nt = -t * inv;
                    Real prev = ZERO;
#pragma unroll
                    for (int j = 0; j < CT; ++j) {
                        const Real r = T::sat(T::fma_(L[j], inv, nt));
                        hs[CT - j] = T::fma_(r - prev, wgt, hs[CT - j]);
                        prev = r;
                    }
                    hs[0] = T::fma_((Real)1 - prev, wgt, hs[0]);
                    // somebody waits: all servers are busy, the state is c + q (q = 0 adds 0 to the first bin)
                    add_q_state(q0 ? 1u : q, q0 ? ZERO : dt);
                }
                // a busy period ends at the first departure that finds the line empty while every channel was
                // taken (base.py:274-281): with nobody waiting that is key 0, if it lies inside (t, te)
                if (BUSY && q0 && d0 > t && d0 < te) busy_sample(d0 - start_busy);
                t = te;

                // one generator site per event: the arrival after this one (lead cursor) or, on a
                // departure that leaves jobs waiting, the next head of the queue (lag cursor)
                const bool pop_more = !FINITE && !is_arr && q > 1u;
                Real gx = ZERO, sx = ZERO;
                if (is_arr || pop_more) {
                    const uint32_t ridx = (is_arr ? nxt : nxt - q) + 1u;
#if MQ_K2_RK
                    uint2 w = philox2x32_10_rk(ridx, rep, P.rk);
#else
                    uint2 w = philox2x32_10(ridx, rep, key0);
#endif
                    gx = draw_a(w.x, ridx);
                    sx = draw_s(w.y, ridx);
                }

                // a job starts service: the arriving one if a server is free (send_task_to_channel
                // base.py:155-184), or the head of the queue on a departure (base.py:288-310)
                const bool start = is_arr ? (q0 && d0 <= te) : true;
                Real a_q = a_h, S_q = S_h;
                if (FINITE) {
                    a_q = ZERO;
                    S_q = ZERO;
                    if (!is_arr) {
                        const Ring e = ring[(long long)head * P.lanes + gtid];
                        a_q = a_anchor + e.x;
                        S_q = e.y;
                        a_anchor = a_q;
                        head++;
                        if (head >= cap) head = 0u;
                    }
                }
                const Real S_st = is_arr ? S_pend : S_q;
                const Real w = is_arr ? ZERO : T::mx(t - a_q, ZERO);
                const Real vv = w + S_st;
                const uint32_t tag = Tag<CT>::of(d0);
                const Real y = start ? stamp<CT, Real>(t + S_st, tag) : d0;  // d0 back in = no change
                insert_drop_first<CT, Real>(L, y);
                if (start) T::st(vslot_sa + tag * COL, vv);
                if (BUSY) start_busy = start ? t : start_busy;  // the last start before all channels are taken counts
                acc4(sw, start ? w : ZERO);
                acc4(sv, start ? vv : ZERO);
                taked += start ? 1u : 0u;
                // ---- queue ----
                const bool accept = start || !FINITE || q < cap;  // base.py:204-212
                const bool enq = is_arr && !start && accept;
                if (FINITE) {
                    if (enq) {
                        Real gap = ZERO;
                        if (q0) a_anchor = t; else gap = t - a_lastq;
                        a_lastq = t;
                        uint32_t slot = head + q;
                        if (slot >= cap) slot -= cap;
                        ring[(long long)slot * P.lanes + gtid] = T::ring(gap, S_pend);
                    }
                } else {
                    const bool enq_first = enq && q0;  // the arriving job becomes the head
                    a_h = enq_first ? t : (pop_more ? a_h + gx : a_h);
                    S_h = enq_first ? S_pend : (pop_more ? sx : S_h);
                }
                tA = is_arr ? t + gx : tA;
                S_pend = is_arr ? sx : S_pend;
                nxt += is_arr ? 1u : 0u;
                dropped += (is_arr && !accept) ? 1u : 0u;
                q += (enq ? 1u : 0u) - (is_arr ? 0u : 1u);
                if (taked >= next_stop) break;  // the N-th start, or the end of the warm-up (handled between chunks)
            }
            rebase();
            if (!warm_done && P.warm != 0u && taked == P.warm) {
                // end of the warm-up (the clock was just re-based to this very instant): forget every statistic
                // gathered so far, the job that just started included; jobs still in service are marked so that
                // the final correction skips them
                warm_done = true;
                next_stop = N;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    sw[k] = ZERO;
                    sv[k] = ZERO;
                    rec[(long long)(MQ_F_W + k) * R] = 0.0;
                    rec[(long long)(MQ_F_V + k) * R] = 0.0;
                }
                if (BUSY) {
                    for (int k = 0; k < 3; ++k) {
                        sb[k] = ZERO;
                        rec[(long long)(MQ_REC_ROWS(P.p_bins) + MQ_FB_BUSY + k) * R] = 0.0;
                    }
                    nbusy = 0;
                }
#pragma unroll
                for (int m = 0; m <= CT; ++m) hs[m] = ZERO;
                for (int b = 0; b < MQ_BF; ++b) hist[b * MQ_BLOCK] = ZERO;
                for (int k = 0; k < P.p_bins; ++k) rec[(long long)(MQ_F_P + k) * R] = 0.0;
                rec[(long long)MQ_F_POVER * R] = 0.0;
                for (int j = 0; j < CT; ++j) vslot[j * MQ_BLOCK] = (Real)-1;
                t_warm = tbase;
            }
            if ((++epoch % MQ_FLUSH) == 0) flush();
        }

        // ================= end game: one event at a time until the N-th departure =================
        // (base.py:505: the loop stops right after served == total_served)
        if (N > 0u) {
            int busy = 0;
#pragma unroll
            for (int j = 0; j < CT; ++j) {
                const uint32_t tg = Tag<CT>::of(L[j]);
                const bool real = L[j] < T::LIM;         // not an unused slot
                const bool gone = L[j] <= t;             // departed (settled lazily) or never used
                L[j] = (real && gone) ? T::real(T::KEY_GONE | (typename T::U)tg) : L[j];
                busy += (real && !gone) ? 1 : 0;
            }
            uint32_t served = taked - (uint32_t)busy;
            while (served < N) {
                Real dmin = T::HUGE_T;
#pragma unroll
                for (int j = 0; j < CT; ++j) dmin = L[j] > -T::LIM ? T::mn(dmin, L[j]) : dmin;
                const bool is_arr = busy == 0 || tA <= dmin;
                const Real te = is_arr ? tA : dmin;
                {
                    const long long k = (long long)busy + q;
                    rec[(k < P.p_bins ? MQ_F_P + k : MQ_F_POVER) * R] += (double)(te - t);
                }
                t = te;
                if (is_arr) {
                    const uint32_t idx1 = nxt + 1u;
                    nxt++;
                    uint2 w = philox2x32_10(idx1, rep, key0);
                    const Real gx = draw_a(w.x, idx1);
                    const Real sx = draw_s(w.y, idx1);
                    const Real S_i = S_pend;
                    tA = t + gx;
                    S_pend = sx;
                    if (busy < c) {
                        bool placed = false;
#pragma unroll
                        for (int j = 0; j < CT; ++j) {
                            if (!placed && L[j] < -T::LIM) {
                                const uint32_t tg = Tag<CT>::of(L[j]);
                                L[j] = stamp<CT, Real>(t + S_i, tg);
                                T::st(vslot_sa + tg * COL, S_i);
                                placed = true;
                            }
                        }
                        busy++;
                        taked++;
                        acc4(sv, S_i);
                        if (BUSY) start_busy = t;
                    } else if (!FINITE || q < cap) {
                        if (!FINITE) {
                            if (q == 0u) {
                                a_h = t;
                                S_h = S_i;
                            }
                        } else {
                            Real gap = ZERO;
                            if (q == 0u) a_anchor = t; else gap = t - a_lastq;
                            a_lastq = t;
                            uint32_t slot = head + q;
                            if (slot >= cap) slot -= cap;
                            ring[(long long)slot * P.lanes + gtid] = T::ring(gap, S_i);
                        }
                        q++;
                    } else {
                        dropped++;
                    }
                } else {
                    served++;
                    Real a = ZERO, S = ZERO;
                    const bool pop = q > 0u;
                    if (pop) {
                        if (!FINITE) {
                            a = a_h;
                            S = S_h;
                            if (q > 1u) {
                                const uint32_t ridx = nxt - q + 1u;
                                uint2 w = philox2x32_10(ridx, rep, key0);
                                a_h += draw_a(w.x, ridx);
                                S_h = draw_s(w.y, ridx);
                            }
                        } else {
                            const Ring e = ring[(long long)head * P.lanes + gtid];
                            a = a_anchor + e.x;
                            a_anchor = a;
                            S = e.y;
                            head++;
                            if (head >= cap) head = 0u;
                        }
                        q--;
                        taked++;
                        if (BUSY) start_busy = t;
                    } else {
                        if (BUSY && busy == c) busy_sample(t - start_busy);
                        busy--;
                    }
                    const Real w = T::mx(t - a, ZERO);
                    const Real vv = w + S;
#pragma unroll
                    for (int j = 0; j < CT; ++j) {
                        if (L[j] == dmin) {
                            const uint32_t tg = Tag<CT>::of(L[j]);
                            L[j] = pop ? stamp<CT, Real>(t + S, tg) : T::real(T::KEY_GONE | (typename T::U)tg);
                            if (pop) T::st(vslot_sa + tg * COL, vv);
                        }
                    }
                    if (pop) {
                        acc4(sw, w);
                        acc4(sv, vv);
                    }
                }
            }
        }
        tbase += (double)t;
        flush();

        // ---- jobs still in service at the stop were counted at start: take them out ----
        double vfix[4] = {0.0, 0.0, 0.0, 0.0};
        uint32_t in_service = 0;
#pragma unroll
        for (int j = 0; j < CT; ++j) {
            const Real xs = T::abs_(L[j]) < T::LIM ? vslot[Tag<CT>::of(L[j]) * MQ_BLOCK] : (Real)-1;
            if (xs >= ZERO && !P.v_at_start) {  // in service and started after the warm-up
                in_service++;
                const double x = (double)xs;
                const double x2 = x * x;
                vfix[0] += x;
                vfix[1] += x2;
                vfix[2] += x2 * x;
                vfix[3] += x2 * x2;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) rec[(long long)(MQ_F_V + k) * R] -= vfix[k];
        rec[(long long)MQ_F_ARRIVED * R] = (double)nxt;
        // sample counts: w over the jobs started (after the warm-up), v over those of them that left
        rec[(long long)MQ_F_TAKED * R] = (double)(taked - P.warm);
        rec[(long long)MQ_F_SERVED * R] = (P.warm || P.v_at_start) ? (double)(taked - P.warm - in_service) : (double)N;
        rec[(long long)MQ_F_DROPPED * R] = (double)dropped;
        rec[(long long)MQ_F_TTEK * R] = tbase - t_warm;
        rec[(long long)MQ_F_FLAGS * R] = (double)flags;
        if (BUSY) rec[(long long)(MQ_REC_ROWS(P.p_bins) + MQ_FB_BUSYN) * R] = (double)nbusy;
    }
}

}  // namespace mq
