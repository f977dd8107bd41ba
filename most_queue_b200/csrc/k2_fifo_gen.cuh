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
constexpr uint32_t MQ_KEY_GONE = 0xfe800000u;  // -8.5e37: server free "since ever" (plus slot id)
constexpr uint32_t MQ_KEY_OFF = 0x7e800000u;   // +8.5e37: slot not used (c < CT): always busy, never departs
constexpr float MQ_KEY_LIM = 1.0e30f;          // |key| below this is a real time

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
    double *rec;      // [MQ_REC_ROWS(p_bins)][reps], zeroed by the caller
    float2 *ring;     // finite buffer: [buffer][total lanes]
    long long lanes;  // gridDim.x * MQ_BLOCK
};

__device__ __forceinline__ float ld_shared(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ void st_shared(uint32_t addr, float v)
{
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

template <int CT>
struct Tag {
    static constexpr int TB = CT <= 1 ? 0 : CT <= 2 ? 1 : CT <= 4 ? 2 : CT <= 8 ? 3 : CT <= 16 ? 4 : 5;
    static constexpr uint32_t MASK = (1u << TB) - 1u;
    static constexpr uint32_t HALF = TB ? (1u << (TB - 1)) : 0u;
};

// round the key to the tag grid (nearest) and stamp the slot id into its low bits
template <int CT>
__device__ __forceinline__ float stamp(float x, uint32_t tag)
{
    if (Tag<CT>::TB == 0) return x;
    uint32_t b = __float_as_uint(x);
    return __uint_as_float(((b + Tag<CT>::HALF) & ~Tag<CT>::MASK) | tag);
}

// L[0..CT-1] sorted ascending.  new L = sorted(L[0..CT-2] + {y}); the old maximum is dropped.
template <int CT>
__device__ __forceinline__ void insert_drop_last(float (&L)[CT], float y)
{
    float prev = L[0];
    L[0] = fminf(prev, y);
#pragma unroll
    for (int j = 1; j < CT; ++j) {
        float cur = L[j];
        float up = fmaxf(prev, y);
        L[j] = (j == CT - 1) ? up : fminf(cur, up);
        prev = cur;
    }
    if (CT == 1) L[0] = y;
}

// new L = sorted(L[1..CT-1] + {y}); the old minimum is dropped.
template <int CT>
__device__ __forceinline__ void insert_drop_first(float (&L)[CT], float y)
{
    if (CT == 1) {
        L[0] = y;
        return;
    }
    float prev = L[1];
    L[0] = fminf(prev, y);
#pragma unroll
    for (int j = 1; j < CT - 1; ++j) {
        float cur = L[j + 1];
        L[j] = fminf(cur, fmaxf(prev, y));
        prev = cur;
    }
    L[CT - 1] = fmaxf(prev, y);
}

// One wiring for both event kinds: an arrival keeps L[0..CT-2] (the old maximum -- a free slot, or
// y itself when nothing starts -- is dropped), a departure drops L[0]; y is inserted in order.
template <int CT>
__device__ __forceinline__ void unified_insert(float (&L)[CT], float y, bool is_arr)
{
    if (CT == 1) {
        L[0] = y;
        return;
    }
    float M[CT > 1 ? CT - 1 : 1];
#pragma unroll
    for (int j = 0; j < CT - 1; ++j) M[j] = is_arr ? L[j] : L[j + 1];
    L[0] = fminf(M[0], y);
#pragma unroll
    for (int j = 1; j < CT - 1; ++j) L[j] = fminf(M[j], fmaxf(M[j - 1], y));
    L[CT - 1] = fmaxf(M[CT - 2], y);
}

// After a re-base two keys can land on the same grid point and then differ only by their slot
// ids, which may invert an adjacent pair.  Restore strict order (odd-even transposition until
// clean; almost always the check alone).
template <int CT>
__device__ __forceinline__ void restore_order(float (&L)[CT])
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
                const float lo = fminf(L[j], L[j + 1]), hi = fmaxf(L[j], L[j + 1]);
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

template <int CT, int SRCK, int SRVK, bool FINITE>
__global__ void __launch_bounds__(MQ_BLOCK, (CT <= 8 ? MQ_K2_MINB : 2)) k2_fifo_gen(const K2Params P)
{
    extern __shared__ float smem[];
    const int tid = threadIdx.x;
    float *hist = smem + tid;                           // [MQ_BF][MQ_BLOCK]: states c+1 .. c+MQ_BF
    float *vslot = smem + MQ_BF * MQ_BLOCK + tid;       // [CT][MQ_BLOCK]
    // 32-bit shared-window addresses of this lane's columns (keeps the address math out of the loop)
    const uint32_t hist_sa = (uint32_t)__cvta_generic_to_shared(hist);
    const uint32_t vslot_sa = (uint32_t)__cvta_generic_to_shared(vslot);
    const long long gtid = (long long)blockIdx.x * MQ_BLOCK + tid;
    const int c = P.c;
    const uint32_t key0 = P.key0;
    const long long R = P.reps;
    const uint32_t N = (uint32_t)P.jobs;
    const uint32_t cap = FINITE ? (uint32_t)P.buffer : 0u;
    const int off = CT - c;  // register hs[m] holds state m - off

    for (long long rl = gtid; rl < R; rl += P.lanes) {
        const uint32_t rep = P.rep0 + (uint32_t)rl;
        double *rec = P.rec + rl;

        // ---- state ----
        float L[CT];
#pragma unroll
        for (int j = 0; j < CT; ++j) {
            // ascending: the c real slots, free since ever (more negative = larger slot id), then unused
            L[j] = __uint_as_float(j < c ? (MQ_KEY_GONE | (uint32_t)(c - 1 - j)) : (MQ_KEY_OFF | (uint32_t)j));
        }
        float hs[CT + 1];
#pragma unroll
        for (int m = 0; m <= CT; ++m) hs[m] = 0.0f;
#pragma unroll
        for (int b = 0; b < MQ_BF; ++b) hist[b * MQ_BLOCK] = 0.0f;

        float t = 0.0f;
        double tbase = 0.0;
        float sw[4] = {0.f, 0.f, 0.f, 0.f}, sv[4] = {0.f, 0.f, 0.f, 0.f};
        // nxt = index of the arrival scheduled at tA (= number of arrivals so far); q = waiting jobs.
        // Infinite buffer: the waiting jobs are the last q arrivals, so the head of the queue is nxt - q.
        uint32_t nxt = 0, taked = 0, dropped = 0, q = 0;
        float a_h = 0.0f, S_h = 0.0f;  // infinite buffer: arrival time and service time of the queue head
        uint32_t head = 0;             // finite buffer: ring bookkeeping
        float a_anchor = 0.0f, a_lastq = 0.0f;
        uint32_t flags = 0;
        double t_warm = 0.0;
        bool warm_done = false;
        uint32_t next_stop = (P.warm != 0u && P.warm < N) ? P.warm : N;

        // first arrival: the draw set_sources makes (base.py:112)
        float tA, S_pend;
        {
            uint2 w = philox2x32_10(0u, rep, key0);
            tA = sample<SRCK>(P.src, w.x, 0u, rep, key0, 0);
            S_pend = sample<SRVK>(P.srv, w.y, 0u, rep, key0, 1);
        }

        auto flush = [&]() {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                rec[(long long)(MQ_F_W + k) * R] += (double)sw[k];
                rec[(long long)(MQ_F_V + k) * R] += (double)sv[k];
                sw[k] = 0.0f;
                sv[k] = 0.0f;
            }
#pragma unroll
            for (int m = 0; m <= CT; ++m) {
                const int k = m - off;
                if (k >= 0) {
                    rec[(long long)(k < P.p_bins ? MQ_F_P + k : MQ_F_POVER) * R] += (double)hs[m];
                    hs[m] = 0.0f;
                }
            }
            for (int b = 0; b < MQ_BF; ++b) {
                const int k = c + 1 + b;
                rec[(long long)(k < P.p_bins ? MQ_F_P + k : MQ_F_POVER) * R] += (double)hist[b * MQ_BLOCK];
                hist[b * MQ_BLOCK] = 0.0f;
            }
        };
        auto rebase = [&]() {
            const float base = t;
            tbase += (double)base;
            t = 0.0f;
            tA -= base;
            a_h -= base;
            a_anchor -= base;
            a_lastq -= base;
#pragma unroll
            for (int j = 0; j < CT; ++j) {
                const float moved = stamp<CT>(L[j] - base, __float_as_uint(L[j]) & Tag<CT>::MASK);
                L[j] = fabsf(L[j]) < MQ_KEY_LIM ? moved : L[j];
            }
            restore_order<CT>(L);
        };
        // time with c + qq jobs in the system, qq >= 1
        auto add_q_state = [&](uint32_t qq, float dt) {
            if (qq <= (uint32_t)MQ_BF) {
                const uint32_t ha = hist_sa + (qq - 1u) * (MQ_BLOCK * 4u);
                st_shared(ha, ld_shared(ha) + dt);
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
                const float d0 = L[0];
                const bool q0 = q == 0u;
                // nobody waits: the next event is the arrival (departures are settled lazily);
                // otherwise arrival or departure, arrival wins ties (base.py:442-454)
                const bool is_arr = q0 || tA <= d0;
                const float te = is_arr ? tA : d0;
                // _update_state_probs (base.py:327-340)
                const float dt = te - t;
                {
                    // nobody waits: between t and te the c servers finish in key order, state (c - j) lasts
                    // until key j.  With r_j = saturate((key_j - t) / dt) the time in state (c - j) is
                    // (r_j - r_{j-1}) dt: one FFMA.SAT per key on the FMA pipe instead of two FMNMX clamps.
                    // Lanes with waiting jobs run the same code with a zero weight (no divergent arm).
                    const float inv = (q0 && dt > 0.0f) ? fast_rcp(dt) : 0.0f;
                    const float wgt = q0 ? dt : 0.0f;
                    const float nt = -t * inv;
                    float prev = 0.0f;
#pragma unroll
                    for (int j = 0; j < CT; ++j) {
                        const float r = __saturatef(fmaf(L[j], inv, nt));
                        hs[CT - j] = fmaf(r - prev, wgt, hs[CT - j]);
                        prev = r;
                    }
                    hs[0] = fmaf(1.0f - prev, wgt, hs[0]);
                    // somebody waits: all servers are busy, the state is c + q (q = 0 adds 0 to the first bin)
                    add_q_state(q0 ? 1u : q, q0 ? 0.0f : dt);
                }
                t = te;

                // one generator site per event: the arrival after this one (lead cursor) or, on a
                // departure that leaves jobs waiting, the next head of the queue (lag cursor)
                const bool pop_more = !FINITE && !is_arr && q > 1u;
                float gx = 0.0f, sx = 0.0f;
                if (is_arr || pop_more) {
                    const uint32_t ridx = (is_arr ? nxt : nxt - q) + 1u;
#if MQ_K2_RK
                    uint2 w = philox2x32_10_rk(ridx, rep, P.rk);
#else
                    uint2 w = philox2x32_10(ridx, rep, key0);
#endif
                    gx = sample<SRCK>(P.src, w.x, ridx, rep, key0, 0);
                    sx = sample<SRVK>(P.srv, w.y, ridx, rep, key0, 1);
                }

                // a job starts service: the arriving one if a server is free (send_task_to_channel
                // base.py:155-184), or the head of the queue on a departure (base.py:288-310)
                const bool start = is_arr ? (q0 && d0 <= te) : true;
                float a_q = a_h, S_q = S_h;
                if (FINITE) {
                    a_q = 0.0f;
                    S_q = 0.0f;
                    if (!is_arr) {
                        const float2 e = P.ring[(long long)head * P.lanes + gtid];
                        a_q = a_anchor + e.x;
                        S_q = e.y;
                        a_anchor = a_q;
                        head++;
                        if (head >= cap) head = 0u;
                    }
                }
                const float S_st = is_arr ? S_pend : S_q;
                const float w = is_arr ? 0.0f : fmaxf(t - a_q, 0.0f);
                const float vv = w + S_st;
                const uint32_t tag = __float_as_uint(d0) & Tag<CT>::MASK;
                const float y = start ? stamp<CT>(t + S_st, tag) : d0;  // d0 back in = no change
                insert_drop_first<CT>(L, y);
                if (start) st_shared(vslot_sa + tag * (MQ_BLOCK * 4u), vv);
                acc4(sw, start ? w : 0.0f);
                acc4(sv, start ? vv : 0.0f);
                taked += start ? 1u : 0u;
                // ---- queue ----
                const bool accept = start || !FINITE || q < cap;  // base.py:204-212
                const bool enq = is_arr && !start && accept;
                if (FINITE) {
                    if (enq) {
                        float gap = 0.0f;
                        if (q0) a_anchor = t; else gap = t - a_lastq;
                        a_lastq = t;
                        uint32_t slot = head + q;
                        if (slot >= cap) slot -= cap;
                        P.ring[(long long)slot * P.lanes + gtid] = make_float2(gap, S_pend);
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
                    sw[k] = 0.0f;
                    sv[k] = 0.0f;
                    rec[(long long)(MQ_F_W + k) * R] = 0.0;
                    rec[(long long)(MQ_F_V + k) * R] = 0.0;
                }
#pragma unroll
                for (int m = 0; m <= CT; ++m) hs[m] = 0.0f;
                for (int b = 0; b < MQ_BF; ++b) hist[b * MQ_BLOCK] = 0.0f;
                for (int k = 0; k < P.p_bins; ++k) rec[(long long)(MQ_F_P + k) * R] = 0.0;
                rec[(long long)MQ_F_POVER * R] = 0.0;
                for (int j = 0; j < CT; ++j) vslot[j * MQ_BLOCK] = -1.0f;
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
                const uint32_t tg = __float_as_uint(L[j]) & Tag<CT>::MASK;
                const bool real = L[j] < MQ_KEY_LIM;     // not an unused slot
                const bool gone = L[j] <= t;             // departed (settled lazily) or never used
                L[j] = (real && gone) ? __uint_as_float(MQ_KEY_GONE | tg) : L[j];
                busy += (real && !gone) ? 1 : 0;
            }
            uint32_t served = taked - (uint32_t)busy;
            while (served < N) {
                float dmin = 3.0e38f;
#pragma unroll
                for (int j = 0; j < CT; ++j) dmin = L[j] > -MQ_KEY_LIM ? fminf(dmin, L[j]) : dmin;
                const bool is_arr = busy == 0 || tA <= dmin;
                const float te = is_arr ? tA : dmin;
                {
                    const long long k = (long long)busy + q;
                    rec[(k < P.p_bins ? MQ_F_P + k : MQ_F_POVER) * R] += (double)(te - t);
                }
                t = te;
                if (is_arr) {
                    const uint32_t idx1 = nxt + 1u;
                    nxt++;
                    uint2 w = philox2x32_10(idx1, rep, key0);
                    const float gx = sample<SRCK>(P.src, w.x, idx1, rep, key0, 0);
                    const float sx = sample<SRVK>(P.srv, w.y, idx1, rep, key0, 1);
                    const float S_i = S_pend;
                    tA = t + gx;
                    S_pend = sx;
                    if (busy < c) {
                        bool placed = false;
#pragma unroll
                        for (int j = 0; j < CT; ++j) {
                            if (!placed && L[j] < -MQ_KEY_LIM) {
                                const uint32_t tg = __float_as_uint(L[j]) & Tag<CT>::MASK;
                                L[j] = stamp<CT>(t + S_i, tg);
                                st_shared(vslot_sa + tg * (MQ_BLOCK * 4u), S_i);
                                placed = true;
                            }
                        }
                        busy++;
                        taked++;
                        acc4(sv, S_i);
                    } else if (!FINITE || q < cap) {
                        if (!FINITE) {
                            if (q == 0u) {
                                a_h = t;
                                S_h = S_i;
                            }
                        } else {
                            float gap = 0.0f;
                            if (q == 0u) a_anchor = t; else gap = t - a_lastq;
                            a_lastq = t;
                            uint32_t slot = head + q;
                            if (slot >= cap) slot -= cap;
                            P.ring[(long long)slot * P.lanes + gtid] = make_float2(gap, S_i);
                        }
                        q++;
                    } else {
                        dropped++;
                    }
                } else {
                    served++;
                    float a = 0.0f, S = 0.0f;
                    const bool pop = q > 0u;
                    if (pop) {
                        if (!FINITE) {
                            a = a_h;
                            S = S_h;
                            if (q > 1u) {
                                const uint32_t ridx = nxt - q + 1u;
                                uint2 w = philox2x32_10(ridx, rep, key0);
                                a_h += sample<SRCK>(P.src, w.x, ridx, rep, key0, 0);
                                S_h = sample<SRVK>(P.srv, w.y, ridx, rep, key0, 1);
                            }
                        } else {
                            const float2 e = P.ring[(long long)head * P.lanes + gtid];
                            a = a_anchor + e.x;
                            a_anchor = a;
                            S = e.y;
                            head++;
                            if (head >= cap) head = 0u;
                        }
                        q--;
                        taked++;
                    } else {
                        busy--;
                    }
                    const float w = fmaxf(t - a, 0.0f);
                    const float vv = w + S;
#pragma unroll
                    for (int j = 0; j < CT; ++j) {
                        if (L[j] == dmin) {
                            const uint32_t tg = __float_as_uint(L[j]) & Tag<CT>::MASK;
                            L[j] = pop ? stamp<CT>(t + S, tg) : __uint_as_float(MQ_KEY_GONE | tg);
                            if (pop) st_shared(vslot_sa + tg * (MQ_BLOCK * 4u), vv);
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
            const float xs = fabsf(L[j]) < MQ_KEY_LIM ? vslot[(__float_as_uint(L[j]) & Tag<CT>::MASK) * MQ_BLOCK] : -1.0f;
            if (xs >= 0.0f && !P.v_at_start) {  // in service and started after the warm-up
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
    }
}

}  // namespace mq
