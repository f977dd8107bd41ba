// k4v2_prio.cu -- K4 v2: the multi-class priority queue kernel (NP / PR / RS / RW) as ONE predicated trip body
// with the per-lane state in registers and shared memory, for c <= 8 channels and K <= 4 classes.
//
// Same event sequence, same arithmetic and same record rows as k4_prio.cu (which stays the kernel for larger c or
// K; see its header for the reference map, most_queue/sim/priority.py).  What changed is the shape of the code:
//   * v1 keeps the state in thread-private arrays (local memory) and has one code arm per event kind: ncu shows
//     7.2 of 32 lanes active per instruction and 2e10 local-memory sectors per 5e8 jobs.
//   * here a loop trip is an arrival (of whichever class) or a departure; both run through the same straight-line
//     body, with at most one "a job starts on a slot" site, one "a record joins a waiting line" site, one
//     service-draw site and one moment-update site per trip.
//   * completion times, the class of the job on each slot (4 bits per slot) and the next arrival time of every
//     class are registers; the job data of each slot, the per-class moments, counters, cursors and the first
//     few time-in-state bins are shared memory laid out [word][lane] (the bank depends on the lane only, so
//     per-lane slot / class indices never conflict); further bins go to the zeroed record with fire-and-forget
//     RED.ADD.F64 (same rounding, same order per address).
// Waiting lines stay rings of 4-double job records in HBM scratch ([class][slot][field][lane]).
#include "k4_prio.cuh"

namespace mq {

constexpr int K4V_BLOCK = 64;
#ifndef K4V_BF_GEN
#define K4V_BF_GEN 2
#endif
constexpr int K4V_BF_REPLAY = 4;  // time-in-state bins per class kept in shared memory (replay tier)
#ifndef K4V_SB_N
#define K4V_SB_N 8
#endif
constexpr int K4V_SB = K4V_SB_N;  // generator tier: service draws kept ahead per class and lane (power of two)

template <int CT, int KT, bool REPLAY>
struct K4VLayout {
    // doubles, [word][lane]
    static constexpr int ARR = 0, WAIT = CT, TOT = 2 * CT, CLS0 = 3 * CT;
    // replay: power sums and the reference's running means (read-modify-write) live here; generator tier: the power
    // sums go straight to the zeroed record (fire-and-forget RED.ADD.F64), which frees 128 bytes per lane
    static constexpr int T_OLD = 0, P_OVER = 1, VSUM = 2, WSUM = 6, VRUN = 10, WRUN = 14;
    static constexpr int HIST = REPLAY ? 18 : 2;
    static constexpr int BF = REPLAY ? K4V_BF_REPLAY : K4V_BF_GEN;  // time-in-state bins kept on chip
    static constexpr int DSTR = HIST + BF;
    static constexpr int ND = CLS0 + KT * DSTR;
    // 32-bit words, [word][lane]
    static constexpr int ARRIVED = 0, TAKED = 1, SERVED = 2, DROPPED = 3, INSYS = 4, AIDX = 5, SIDX = 6, QHEAD = 7,
                         QSIZE = 8, SFILL = 9, ISTR = 10;
    static constexpr int NI = KT * ISTR;
    static constexpr int NF = REPLAY ? 0 : KT * K4V_SB;  // fp32 service draws kept ahead
    static constexpr size_t BYTES = (size_t)K4V_BLOCK * (8 * ND + 4 * NI + 4 * NF);
};

template <int CT, int KT, bool REPLAY>
__global__ void __launch_bounds__(K4V_BLOCK) k4v2_prio(const K4Params P)
{
    using LY = K4VLayout<CT, KT, REPLAY>;
    extern __shared__ double k4v_sm[];
    const int tid = threadIdx.x;
    double *smd = k4v_sm + tid;
    int *smi = reinterpret_cast<int *>(k4v_sm + LY::ND * K4V_BLOCK) + tid;
    float *smf = reinterpret_cast<float *>(smi - tid + LY::NI * K4V_BLOCK) + tid;
#define SD(w) smd[(w) * K4V_BLOCK]
#define SI(w) smi[(w) * K4V_BLOCK]
#define DC(k, f) SD(LY::CLS0 + (k) * LY::DSTR + (f))
#define IC(k, f) SI((k) * LY::ISTR + (f))

    // the laws of the classes, read with a per-lane class index: shared memory instead of a thread-private copy
    __shared__ DistF sh_src[KT], sh_srv[KT];
    if (!REPLAY && tid < KT) {
        sh_src[tid] = P.src[tid];
        sh_srv[tid] = P.srv[tid];
    }
    __syncthreads();

    const long long gtid = (long long)blockIdx.x * K4V_BLOCK + tid;
    const long long R = P.reps;
    const int c = P.c, K = P.K;
    const int rows = K4_ROWS(P.p_bins);
    const uint32_t cmask = (1u << c) - 1u;

    for (long long rl = gtid; rl < R; rl += P.lanes) {
        const uint32_t rep = P.rep0 + (uint32_t)rl;
        double *rec = P.rec + rl;
        double *Q = P.queue + gtid;

        for (int w = 0; w < LY::ND; ++w) SD(w) = 0.0;
        for (int w = 0; w < LY::NI; ++w) SI(w) = 0;
        double end[CT], at[KT];
#pragma unroll
        for (int j = 0; j < CT; ++j) end[j] = j < c ? 1e10 : 1e300;
        uint32_t busy = 0, cls_pack = 0;
        uint32_t need = 1u;  // generator tier: some ring of this lane is (nearly) empty
        int status = 0;

        auto next_gap = [&](int k) -> double {
            const int i = IC(k, LY::AIDX);
            IC(k, LY::AIDX) = i + 1;
            if (REPLAY) {
                if (i >= P.n_a[k]) {
                    status = MQ_E_INPUT_SHORT;
                    return 0.0;
                }
                return P.A[P.aoff[k] + (long long)i * R + rl];
            }
            const uint32_t key = P.key0 + MQ_KEY_STRIDE * (64u * (uint32_t)k);
            uint2 w = philox2x32_10((uint32_t)i, rep, key);
            const DistF d = sh_src[k];
            return (double)sample<-1>(d, w.x, (uint32_t)i, rep, key, 0);
        };
        // Generator tier: service draws are produced up to K4V_SB ahead per class by all lanes of the warp together
        // (at the top of a trip) and consumed from a per-lane ring.  Draw j of a stream is a pure function of j, so
        // drawing ahead changes nothing; it keeps the rejection-loop samplers out of the predicated event body,
        // where only the lanes that start a service in this trip would run them.
        auto refill_services = [&]() {
            // when any lane of the warp is short of draws for some class (RS draws twice in one trip), all lanes top
            // up all their rings together
            if (!__any_sync(__activemask(), need != 0u)) return;
            need = 0u;
            for (int k = 0; k < K; ++k) {
                const uint32_t key = P.key0 + MQ_KEY_STRIDE * (64u * (uint32_t)k);
                const DistF d = sh_srv[k];
                for (;;) {
                    const int fill = IC(k, LY::SFILL);
                    const bool more = fill < K4V_SB;
                    if (!__any_sync(__activemask(), more)) break;
                    if (more) {
                        const uint32_t j = (uint32_t)(IC(k, LY::SIDX) + fill);
                        uint2 w = philox2x32_10(j, rep, key);
                        smf[(k * K4V_SB + (int)(j & (K4V_SB - 1))) * K4V_BLOCK] = sample<-1>(d, w.y, j, rep, key, 1);
                        IC(k, LY::SFILL) = fill + 1;
                    }
                }
            }
        };
        auto next_service = [&](int k) -> double {
            const int j = IC(k, LY::SIDX);
            IC(k, LY::SIDX) = j + 1;
            if (REPLAY) {
                if (j >= P.n_s[k]) {
                    status = MQ_E_INPUT_SHORT;
                    return 0.0;
                }
                return P.S[P.soff[k] + (long long)j * R + rl];
            }
            const int fill = IC(k, LY::SFILL) - 1;
            IC(k, LY::SFILL) = fill;
            need |= fill < 2 ? 1u : 0u;
            return (double)smf[(k * K4V_SB + (j & (K4V_SB - 1))) * K4V_BLOCK];
        };
        const long long QL = P.lanes;  // field stride of a waiting-line record
        auto qrec = [&](int k, int slot) -> double * { return Q + (long long)((k * P.qcap + slot) * 4) * QL; };

#pragma unroll
        for (int k = 0; k < KT; ++k) at[k] = 1e300;
        for (int k = 0; k < K; ++k) {
            const double g = next_gap(k);  // priority.py:130
#pragma unroll
            for (int kk = 0; kk < KT; ++kk)
                if (kk == k) at[kk] = g;
        }

        double ttek = 0.0;
        int free_channels = c;
        long long served_total = 0;

        while (served_total < P.jobs && status == 0) {
            if (!REPLAY) refill_services();
            // ---- next event: _get_min_server_time priority.py:549-562 (busy servers only, strict '<', lowest index
            // on ties) against the earliest class arrival (lowest class on ties); an arrival wins a tie
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
            double et = at[0];
            int ev = 0;
#pragma unroll
            for (int k = 1; k < KT; ++k) {
                const bool lt = at[k] < et;
                et = lt ? at[k] : et;
                ev = lt ? k : ev;
            }
            const bool is_dep = mt < 1e10 && mt < et;
            const bool is_arr = !is_dep;
            const int dcls = (int)((cls_pack >> (4 * midx)) & 15u);
            const int kev = is_dep ? dcls : ev;
            const double te = is_dep ? mt : et;

            // ---- time in state of the class concerned: p[k][in_sys[k]] += t - t_old[k]  (priority.py:304,472)
            {
                const int ins = IC(kev, LY::INSYS);
                const double dt = __dsub_rn(te, DC(kev, LY::T_OLD));
                if (ins < LY::BF) {
                    DC(kev, LY::HIST + ins) = __dadd_rn(DC(kev, LY::HIST + ins), dt);
                } else if (ins < P.p_bins) {
                    atomicAdd(&rec[(long long)(K4_G_ROWS + kev * rows + K4_P + ins) * R], dt);
                } else {
                    DC(kev, LY::P_OVER) = __dadd_rn(DC(kev, LY::P_OVER), dt);
                }
                DC(kev, LY::T_OLD) = te;
            }
            ttek = te;

            // what this trip does to the slots and the waiting lines
            bool start = false, push = false, s_ispr = false, s_new_busy = false;
            int s_slot = 0, s_cls = 0, p_cls = 0;
            double s_arr = 0.0, s_wait = 0.0, s_tte = 0.0;
            double p_arr = 0.0, p_sw = 0.0, p_wait = 0.0, p_tte = -1.0;
            int wslot = -1;
            double wval = 1e10;

            if (is_arr) {
                // arrival(k) priority.py:282-332
                const double g = next_gap(ev);
                const double na = __dadd_rn(ttek, g);
#pragma unroll
                for (int k = 0; k < KT; ++k)
                    if (k == ev) at[k] = na;
                IC(ev, LY::ARRIVED) += 1;
                IC(ev, LY::INSYS) += 1;
                const bool full = free_channels == 0;
                // arrive_with_absolute_prty priority.py:390-455: the FIRST server (index order) with a weaker class
                int vj = -1;
#pragma unroll
                for (int j = CT - 1; j >= 0; --j) {
                    const bool weaker = j < c && (int)((cls_pack >> (4 * j)) & 15u) > ev;
                    vj = weaker ? j : vj;
                }
                const bool pre_found = full && P.prty != 0 && vj >= 0;
                const bool pre_nf = full && P.prty != 0 && vj < 0;
                bool accept = full && P.prty == 0;  // NP: priority.py:331-332
                if (pre_nf) {
                    long long qtot = 0;
#pragma unroll
                    for (int kk = 0; kk < KT; ++kk) qtot += kk < K ? IC(kk, LY::QSIZE) : 0;
                    accept = P.buffer <= 0 || qtot < P.buffer;
                    if (!accept) {
                        IC(ev, LY::DROPPED) += 1;
                        IC(ev, LY::INSYS) -= 1;
                    }
                }
                if (pre_found) {
                    // the victim goes to the TAIL of its class line with its remaining / redrawn / total time
                    push = true;
                    p_cls = (int)((cls_pack >> (4 * vj)) & 15u);
                    p_arr = SD(LY::ARR + vj);
                    p_wait = SD(LY::WAIT + vj);
                    p_sw = ttek;
                    double e_v = 0.0;
#pragma unroll
                    for (int j = 0; j < CT; ++j)
                        if (j == vj) e_v = end[j];
                    if (P.prty == 1)
                        p_tte = __dsub_rn(e_v, ttek);
                    else if (P.prty == 2)
                        p_tte = next_service(ev);  // priority.py:425, drawn before the new job's
                    else
                        p_tte = SD(LY::TOT + vj);
                } else if (accept) {
                    push = true;
                    p_cls = ev;
                    p_arr = ttek;
                    p_sw = ttek;
                    p_wait = 0.0;
                    p_tte = -1.0;
                }
                if (!full || pre_found) {
                    start = true;
                    s_new_busy = !full;
                    s_slot = full ? vj : __ffs(~busy & cmask) - 1;
                    s_cls = ev;
                    s_arr = ttek;
                    s_wait = 0.0;
                    IC(ev, LY::TAKED) += 1;
                }
            } else {
                // serving(c) priority.py:457-539
                const int j = midx, k = dcls;
                const double a = SD(LY::ARR + j), wt = SD(LY::WAIT + j);
                busy &= ~(1u << j);
                SD(LY::TOT + j) = 0.0;
                const int n_served = IC(k, LY::SERVED) + 1;
                IC(k, LY::SERVED) = n_served;
                served_total++;
                free_channels++;
                const double v = __dsub_rn(ttek, a);
                {
                    double pv = v, pw = wt;
                    double inv_count = 0.0, one_minus = 0.0;
                    if (REPLAY) {
                        inv_count = __drcp_rn((double)n_served);  // == 1.0 / n, correctly rounded
                        one_minus = __dsub_rn(1.0, inv_count);
                    }
                    double *gv = rec + (long long)(K4_G_ROWS + k * rows + K4_VSUM) * R;
                    double *gw = rec + (long long)(K4_G_ROWS + k * rows + K4_WSUM) * R;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (REPLAY) {  // refresh_moments_stat, stats_update.py:6-22
                            DC(k, LY::VRUN + i) =
                                __dadd_rn(__dmul_rn(DC(k, LY::VRUN + i), one_minus), __dmul_rn(pv, inv_count));
                            DC(k, LY::WRUN + i) =
                                __dadd_rn(__dmul_rn(DC(k, LY::WRUN + i), one_minus), __dmul_rn(pw, inv_count));
                            DC(k, LY::VSUM + i) = __dadd_rn(DC(k, LY::VSUM + i), pv);
                            DC(k, LY::WSUM + i) = __dadd_rn(DC(k, LY::WSUM + i), pw);
                        } else {
                            atomicAdd(gv, pv);
                            atomicAdd(gw, pw);
                            gv += R;
                            gw += R;
                        }
                        pv = __dmul_rn(pv, v);
                        pw = __dmul_rn(pw, wt);
                    }
                }
                IC(k, LY::INSYS) -= 1;
                // next job: NP scans all class lines, the pre-emptive disciplines start at the finished class
                int kk = -1;
                const int q_lo = P.prty == 0 ? 0 : k;
#pragma unroll
                for (int q = KT - 1; q >= 0; --q) kk = (q < K && q >= q_lo && IC(q, LY::QSIZE) != 0) ? q : kk;
                if (kk >= 0) {
                    const int slot = IC(kk, LY::QHEAD);
                    const int qs = IC(kk, LY::QSIZE) - 1;
                    IC(kk, LY::QSIZE) = qs;
                    IC(kk, LY::QHEAD) = (qs == 0 || slot + 1 >= P.qcap) ? 0 : slot + 1;
                    IC(kk, LY::TAKED) += 1;
                    const double *qr = qrec(kk, slot);
                    s_arr = qr[0];
                    const double sw = qr[QL];
                    s_wait = __dadd_rn(qr[2 * QL], __dsub_rn(ttek, sw));
                    s_tte = qr[3 * QL];
                    s_ispr = s_tte >= 0.0;
                    start = true;
                    s_new_busy = true;
                    s_slot = j;
                    s_cls = kk;
                } else {
                    wslot = j;  // the slot is free again
                }
            }

            if (start) {
                // ServerPriority.start_service, servers.py:219-250
                double e;
                if (s_ispr) {
                    e = __dadd_rn(ttek, s_tte);  // servers.py:237-239: resume with the stored time
                } else {
                    const double sv = next_service(s_cls);
                    SD(LY::TOT + s_slot) = sv;
                    e = __dadd_rn(ttek, sv);
                }
                wslot = s_slot;
                wval = e;
                SD(LY::ARR + s_slot) = s_arr;
                SD(LY::WAIT + s_slot) = s_wait;
                cls_pack = (cls_pack & ~(15u << (4 * s_slot))) | ((uint32_t)s_cls << (4 * s_slot));
                if (s_new_busy) {
                    busy |= 1u << s_slot;
                    free_channels--;
                }
            }
#pragma unroll
            for (int j = 0; j < CT; ++j)
                if (j == wslot) end[j] = wval;

            if (push) {
                const int qs = IC(p_cls, LY::QSIZE);
                if (qs >= P.qcap) {
                    status = MQ_E_OVERFLOW;
                } else {
                    int slot = IC(p_cls, LY::QHEAD) + qs;
                    if (slot >= P.qcap) slot -= P.qcap;
                    double *qr = qrec(p_cls, slot);
                    qr[0] = p_arr;
                    qr[QL] = p_sw;
                    qr[2 * QL] = p_wait;
                    qr[3 * QL] = p_tte;
                    IC(p_cls, LY::QSIZE) = qs + 1;
                }
            }
        }

        rec[(long long)K4_G_TTEK * R] = ttek;
        rec[(long long)K4_G_FLAGS * R] = (double)(status ? -status : 0);
        for (int k = 0; k < K; ++k) {
            double *rk = rec + (long long)(K4_G_ROWS + k * rows) * R;
            const double n_served = (double)IC(k, LY::SERVED);
            for (int i = 0; i < 4; ++i) {
                // generator tier: the sums were accumulated in the record itself (this lane's own atomics)
                const double ws = REPLAY ? DC(k, LY::WSUM + i) : __ldcg(&rk[(long long)(K4_WSUM + i) * R]);
                const double vs = REPLAY ? DC(k, LY::VSUM + i) : __ldcg(&rk[(long long)(K4_VSUM + i) * R]);
                if (REPLAY) {
                    rk[(long long)(K4_WSUM + i) * R] = ws;
                    rk[(long long)(K4_VSUM + i) * R] = vs;
                }
                // generator tier: the running means are not tracked per sample; report the equivalent ratio
                rk[(long long)(K4_WRUN + i) * R] = REPLAY ? DC(k, LY::VRUN + 4 + i) : (n_served > 0 ? ws / n_served : 0.0);
                rk[(long long)(K4_VRUN + i) * R] = REPLAY ? DC(k, LY::VRUN + i) : (n_served > 0 ? vs / n_served : 0.0);
            }
            rk[(long long)K4_ARRIVED * R] = (double)IC(k, LY::ARRIVED);
            rk[(long long)K4_TAKED * R] = (double)IC(k, LY::TAKED);
            rk[(long long)K4_SERVED * R] = n_served;
            rk[(long long)K4_DROPPED * R] = (double)IC(k, LY::DROPPED);
            rk[(long long)K4_INSYS * R] = (double)IC(k, LY::INSYS);
            rk[(long long)K4_QUEUED * R] = (double)IC(k, LY::QSIZE);
            rk[(long long)K4_TOLD * R] = DC(k, LY::T_OLD);
            rk[(long long)K4_AUSED * R] = (double)IC(k, LY::AIDX);
            rk[(long long)K4_SUSED * R] = (double)IC(k, LY::SIDX);
            // bins kept on chip; bins >= LY::BF accumulated in the (zeroed) record directly
            double over = DC(k, LY::P_OVER);
            for (int b = 0; b < LY::BF; ++b) {
                if (b < P.p_bins)
                    rk[(long long)(K4_P + b) * R] = DC(k, LY::HIST + b);
                else
                    over = __dadd_rn(over, DC(k, LY::HIST + b));
            }
            rk[(long long)K4_POVER * R] = over;
        }
    }
#undef SD
#undef SI
#undef DC
#undef IC
}

template <int CT, int KT>
static cudaError_t launch_one(const K4Params &P, int grid, cudaStream_t st)
{
    cudaError_t e;
    if (P.replay) {
        constexpr size_t smem = K4VLayout<CT, KT, true>::BYTES;
        e = cudaFuncSetAttribute(k4v2_prio<CT, KT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k4v2_prio<CT, KT, true><<<grid, K4V_BLOCK, smem, st>>>(P);
    } else {
        constexpr size_t smem = K4VLayout<CT, KT, false>::BYTES;
        e = cudaFuncSetAttribute(k4v2_prio<CT, KT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k4v2_prio<CT, KT, false><<<grid, K4V_BLOCK, smem, st>>>(P);
    }
    return cudaGetLastError();
}

template <int CT, int KT>
static cudaError_t occupancy_one(bool replay, int *occ)
{
    if (replay) {
        constexpr size_t smem = K4VLayout<CT, KT, true>::BYTES;
        cudaError_t e = cudaFuncSetAttribute(k4v2_prio<CT, KT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k4v2_prio<CT, KT, true>, K4V_BLOCK, smem);
    }
    constexpr size_t smem = K4VLayout<CT, KT, false>::BYTES;
    cudaError_t e = cudaFuncSetAttribute(k4v2_prio<CT, KT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k4v2_prio<CT, KT, false>, K4V_BLOCK, smem);
}

bool k4v2_supported(int c, int K) { return c >= 1 && c <= 8 && K >= 1 && K <= 4; }

int k4v2_block() { return K4V_BLOCK; }

#define K4V_DISPATCH(fn, ...)                          \
    do {                                               \
        if (c <= 4 && K <= 2) return fn<4, 2>(__VA_ARGS__); \
        if (c <= 4) return fn<4, 4>(__VA_ARGS__);      \
        if (K <= 2) return fn<8, 2>(__VA_ARGS__);      \
        return fn<8, 4>(__VA_ARGS__);                  \
    } while (0)

cudaError_t k4v2_launch(const K4Params &P, int grid, cudaStream_t st)
{
    const int c = P.c, K = P.K;
    K4V_DISPATCH(launch_one, P, grid, st);
}

cudaError_t k4v2_occupancy(int c, int K, bool replay, int *blocks_per_sm)
{
    K4V_DISPATCH(occupancy_one, replay, blocks_per_sm);
}

}  // namespace mq
