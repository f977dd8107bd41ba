// k7v2_prionet.cu -- K7 v2: the priority-network kernel as ONE predicated trip body with the per-lane state in
// registers and shared memory, for networks of at most 16 servers in total (8 nodes, 4 classes).
//
// Same event sequence, same arithmetic and same record rows as k7_prionet.cu (which stays the kernel for larger
// networks; see its header for the reference map: most_queue/sim/networks/priority_network.py, sim/priority.py).
// It is the K5-v2 driver (k5v2_net.cu) around the K4-v2 node logic (k4v2_prio.cu):
//   * a trip is an external arrival of some class or a service completion; one routing-uniform site, one "pop a line
//     of the node that finished and start that job" site, one "a job arrives at a node" site (shared by external and
//     routed jobs) with its pre-emption arm;
//   * completion times, the real class and the node class of the job on every server (2 bits each) and the next
//     arrival of every class are registers; the five time stamps of the job on each server and the per-(node, class)
//     cursors are shared memory [word][lane]; routing matrices, class maps and laws are shared memory;
//   * generator tier: node-level and network-level power sums and the reported counters go straight into the zeroed
//     record with fire-and-forget RED.ADD.F64; service draws are produced up to K7V_SB ahead per (node, node class)
//     by all lanes of the warp in lock-step;
//   * replay tier: all moments in shared memory (the reference's running means are read-modify-write).
#include "k7_prionet.cuh"

namespace mq {

constexpr int K7V_BLOCK = 64;
constexpr int K7V_SB = 8;    // service draws kept ahead per (node, node class) and lane (power of two)
constexpr int K7V_NEED = 3;  // a trip can draw up to 3 times from one stream (pop start, RS redraw, new job)

__device__ __forceinline__ double k7v_cube(double x)
{
    const double p = __dmul_rn(x, x);
    const double e = __fma_rn(x, x, -p);
    const double q = __dmul_rn(p, x);
    const double e2 = __fma_rn(p, x, -q);
    return __dadd_rn(q, __fma_rn(e, x, e2));
}

template <bool REPLAY>
struct K7VLayout {
    // doubles: 5 per server; replay adds 16 per (node, class) and 12 per class
    __host__ __device__ static int nd(int nserv, int m, int K) { return 5 * nserv + (REPLAY ? 16 * m * K + 12 * K : 0); }
    // ints per (node, class): QHEAD, QSIZE, SIDX, SFILL, NSERVED; per class: AIDX, SERVED
    static constexpr int ISTR = 5;
    __host__ __device__ static int ni(int m, int K) { return ISTR * m * K + 2 * K; }
    __host__ __device__ static int nf(int m, int K) { return REPLAY ? 0 : K7V_SB * m * K; }
    __host__ __device__ static size_t bytes(int nserv, int m, int K)
    {
        return (size_t)K7V_BLOCK * (8 * (size_t)nd(nserv, m, K) + 4 * (size_t)ni(m, K) + 4 * (size_t)nf(m, K));
    }
};

template <int ST, bool REPLAY>
__global__ void __launch_bounds__(K7V_BLOCK) k7v2_prionet(const K7Params P)
{
    using LY = K7VLayout<REPLAY>;
    extern __shared__ double k7v_sm[];
    __shared__ DistF sh_srv[K7_MAXN * K7_MAXK], sh_src[K7_MAXK];
    __shared__ double sh_R[K7_MAXK * (K7_MAXN + 1) * (K7_MAXN + 1)];
    __shared__ int sh_base[K7_MAXN + 1], sh_ch[K7_MAXN], sh_prty[K7_MAXN], sh_map[K7_MAXN * K7_MAXK], sh_nodeof[32];
    const int tid = threadIdx.x;
    const int m = P.m, K = P.K, nserv = P.nserv;
    const int RS = (m + 1) * (m + 1);
    // sh_R holds the RUNNING row sums of the routing matrices, added left to right exactly as choose_next_node does
    // (priority_network.py:137-150), so a trip only compares
    for (int r = tid; r < K * (m + 1); r += K7V_BLOCK) {
        double sum_p = 0.0;
        for (int i = 0; i <= m; ++i) {
            sum_p = __dadd_rn(sum_p, P.R[r * (m + 1) + i]);
            sh_R[r * (m + 1) + i] = sum_p;
        }
    }
    if (tid < 32) {
        int node = 0;
        for (int i = 1; i < m; ++i) node += tid >= P.base[i];
        sh_nodeof[tid] = node;
    }
    for (int i = tid; i < m * K; i += K7V_BLOCK) {
        sh_map[i] = P.nodes_prty[i];
        if (!REPLAY) sh_srv[i] = P.srv[i];
    }
    if (tid <= m) sh_base[tid] = P.base[tid];
    if (tid < m) {
        sh_ch[tid] = P.channels[tid];
        sh_prty[tid] = P.prty[tid];
    }
    if (!REPLAY && tid < K) sh_src[tid] = P.src[tid];
    __syncthreads();

    const int ND = LY::nd(nserv, m, K), NI = LY::ni(m, K);
    double *smd = k7v_sm + tid;
    int *smi = reinterpret_cast<int *>(k7v_sm + ND * K7V_BLOCK) + tid;
    float *smf = reinterpret_cast<float *>(smi - tid + NI * K7V_BLOCK) + tid;
#define SD(w) smd[(w) * K7V_BLOCK]
#define SI(w) smi[(w) * K7V_BLOCK]
#define IQ(nk, f) SI((nk) * LY::ISTR + (f))
    enum { QHEAD = 0, QSIZE = 1, SIDX = 2, SFILL = 3, NSERVED = 4 };
    const int I_CLS = LY::ISTR * m * K;  // per class: AIDX, SERVED
    const int D_ARRN = 0, D_WAITN = nserv, D_ARRT = 2 * nserv, D_WAITT = 3 * nserv, D_TOT = 4 * nserv;
    const int D_NODE = 5 * nserv, D_NET = 5 * nserv + 16 * m * K;
    // replay: per (node, class) VSUM 0..3, WSUM 4..7, VRUN 8..11, WRUN 12..15; per class VSUM 0..2, WSUM 3..5,
    // VRUN 6..8, WRUN 9..11

    const long long gtid = (long long)blockIdx.x * K7V_BLOCK + tid;
    const long long R = P.reps;

    for (long long rl = gtid; rl < R; rl += P.lanes) {
        const uint32_t rep = P.rep0 + (uint32_t)rl;
        double *rec = P.rec + rl;
        double *Q = P.queue + gtid;
        for (int w = 0; w < ND; ++w) SD(w) = 0.0;
        for (int w = 0; w < NI; ++w) SI(w) = 0;

        double end[ST], at[K7_MAXK];
#pragma unroll
        for (int s = 0; s < ST; ++s) end[s] = s < nserv ? 1e10 : 1e300;
        uint32_t busy = 0, k_pack = 0, nc_pack = 0;  // 2 bits per server: real class, node class
        int status = 0, u_idx = 0;
        uint32_t need = 1u;  // generator tier: some ring of this lane is (nearly) empty
        long long served_total = 0;

        auto next_gap = [&](int k) -> double {
            const int i = SI(I_CLS + 2 * k);
            SI(I_CLS + 2 * k) = i + 1;
            if (REPLAY) {
                if (i >= P.n_a[k]) {
                    status = MQ_E_INPUT_SHORT;
                    return 0.0;
                }
                return P.A[P.aoff[k] + (long long)i * R + rl];
            }
            const uint32_t key = P.key0 + MQ_KEY_STRIDE * (64u * (uint32_t)(m * K + k));
            uint2 w = philox2x32_10((uint32_t)i, rep, key);
            const DistF d = sh_src[k];
            return (double)sample<-1>(d, w.x, (uint32_t)i, rep, key, 0);
        };
        auto next_uniform = [&]() -> double {
            const int i = u_idx++;
            if (REPLAY) {
                if (i >= P.n_u) {
                    status = MQ_E_INPUT_SHORT;
                    return 0.0;
                }
                return P.U[(long long)i * R + rl];
            }
            const uint32_t key = P.key0 + MQ_KEY_STRIDE * (64u * (uint32_t)(m * K + K));
            uint2 w = philox2x32_10((uint32_t)i, rep, key);
            return ((double)w.y + 0.5) * 0x1p-32;
        };
        auto refill_services = [&]() {
            // when any lane of the warp is short of draws for some stream, all lanes top up all their rings together
            if (!__any_sync(__activemask(), need != 0u)) return;
            need = 0u;
            for (int nk = 0; nk < m * K; ++nk) {
                const uint32_t key = P.key0 + MQ_KEY_STRIDE * (64u * (uint32_t)nk);
                const DistF d = sh_srv[nk];
                for (;;) {
                    const int fill = IQ(nk, SFILL);
                    const bool more = fill < K7V_SB;
                    if (!__any_sync(__activemask(), more)) break;
                    if (more) {
                        const uint32_t j = (uint32_t)(IQ(nk, SIDX) + fill);
                        uint2 w = philox2x32_10(j, rep, key);
                        smf[(nk * K7V_SB + (int)(j & (K7V_SB - 1))) * K7V_BLOCK] = sample<-1>(d, w.y, j, rep, key, 1);
                        IQ(nk, SFILL) = fill + 1;
                    }
                }
            }
        };
        auto next_service = [&](int nk) -> double {
            const int j = IQ(nk, SIDX);
            IQ(nk, SIDX) = j + 1;
            if (REPLAY) {
                if (j >= P.n_s[nk]) {
                    status = MQ_E_INPUT_SHORT;
                    return 0.0;
                }
                return P.S[P.soff[nk] + (long long)j * R + rl];
            }
            const int fill = IQ(nk, SFILL) - 1;
            IQ(nk, SFILL) = fill;
            need |= fill < K7V_NEED ? 1u : 0u;
            return (double)smf[(nk * K7V_SB + (j & (K7V_SB - 1))) * K7V_BLOCK];
        };
        const long long QL = P.lanes;  // field stride of a waiting-line record
        auto qrec = [&](int nk, int slot) -> double * { return Q + (long long)((nk * P.qcap + slot) * K7_JOBF) * QL; };
        double *rcls0 = rec + (long long)MQ_PN_G_ROWS * R;
        double *rnode0 = rec + (long long)(MQ_PN_G_ROWS + K * MQ_PNC_ROWS) * R;
        auto node_sample = [&](int nk, double v, double w, int count) {  // qs[i].v[j] / .w[j], sim/priority.py:676-690
            if (REPLAY) {
                double *b = &SD(D_NODE + 16 * nk);
                const double inv_count = __drcp_rn((double)count);
                const double one_minus = __dsub_rn(1.0, inv_count);
                double pv = v, pw = w;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    b[(8 + i) * K7V_BLOCK] = __dadd_rn(__dmul_rn(b[(8 + i) * K7V_BLOCK], one_minus), __dmul_rn(pv, inv_count));
                    b[(12 + i) * K7V_BLOCK] = __dadd_rn(__dmul_rn(b[(12 + i) * K7V_BLOCK], one_minus), __dmul_rn(pw, inv_count));
                    b[i * K7V_BLOCK] = __dadd_rn(b[i * K7V_BLOCK], pv);
                    b[(4 + i) * K7V_BLOCK] = __dadd_rn(b[(4 + i) * K7V_BLOCK], pw);
                    pv = __dmul_rn(pv, v);
                    pw = __dmul_rn(pw, w);
                }
            } else {
                double *g = rnode0 + (long long)(nk * MQ_PNN_ROWS) * R;
                double pv = v, pw = w;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    atomicAdd(g + (long long)(MQ_PNN_VSUM + i) * R, pv);
                    atomicAdd(g + (long long)(MQ_PNN_WSUM + i) * R, pw);
                    pv = __dmul_rn(pv, v);
                    pw = __dmul_rn(pw, w);
                }
            }
        };
        auto node_count = [&](int nk, int row) { atomicAdd(rnode0 + (long long)(nk * MQ_PNN_ROWS + row) * R, 1.0); };

#pragma unroll
        for (int k = 0; k < K7_MAXK; ++k) at[k] = 1e300;
        for (int k = 0; k < K; ++k) {
            const double g = next_gap(k);  // priority_network.py:83
#pragma unroll
            for (int kk = 0; kk < K7_MAXK; ++kk)
                if (kk == k) at[kk] = g;
        }
        double ttek = 0.0;

        while (served_total < P.jobs && status == 0) {
            if (!REPLAY) refill_services();
            // ---- next event, base_network_sim.py:184-227: class 0 arrival < class 1 < ... < busy servers in
            // node-major, channel-minor order; the first minimum wins
            double mv[ST];
            int mi[ST];
#pragma unroll
            for (int s = 0; s < ST; ++s) {
                mv[s] = end[s];
                mi[s] = s;
            }
#pragma unroll
            for (int st = 1; st < ST; st *= 2) {
#pragma unroll
                for (int s = 0; s + st < ST; s += 2 * st) {
                    const bool lt = mv[s + st] < mv[s];
                    mv[s] = lt ? mv[s + st] : mv[s];
                    mi[s] = lt ? mi[s + st] : mi[s];
                }
            }
            const double mt = mv[0];
            const int ev = mi[0];
            double et = at[0];
            int ev_k = 0;
#pragma unroll
            for (int k = 1; k < K7_MAXK; ++k) {
                const bool lt = at[k] < et;
                et = lt ? at[k] : et;
                ev_k = lt ? k : ev_k;
            }
            const bool is_dep = mt < 1e10 && mt < et;
            ttek = is_dep ? mt : et;

            // the job that moves in this trip
            double j_arrn = ttek, j_waitn = 0.0;
            int real = ev_k, cur = -1;
            int wslot = -1;
            double wval = 1e10;

            if (!is_dep) {
                // on_arrival(k) priority_network.py:196-215
                atomicAdd(rcls0 + (long long)(ev_k * MQ_PNC_ROWS + MQ_PNC_ARRIVED) * R, 1.0);
                const double na = __dadd_rn(ttek, next_gap(ev_k));
#pragma unroll
                for (int k = 0; k < K7_MAXK; ++k)
                    if (k == ev_k) at[k] = na;
            } else {
                // on_serving :217-240 ; node.serving(c, True) sim/priority.py:457-539
                const int node = sh_nodeof[ev];
                cur = node;
                const int kc = (int)((nc_pack >> (2 * ev)) & 3u);
                real = (int)((k_pack >> (2 * ev)) & 3u);
                const int nk = node * K + kc;
                j_arrn = SD(D_ARRN + ev);
                j_waitn = SD(D_WAITN + ev);
                const double a_t = SD(D_ARRT + ev), w_t = SD(D_WAITT + ev);
                busy &= ~(1u << ev);
                SD(D_TOT + ev) = 0.0;
                wslot = ev;
                const int n_served = IQ(nk, NSERVED) + 1;
                IQ(nk, NSERVED) = n_served;
                node_sample(nk, __dsub_rn(ttek, a_t), w_t, n_served);
                // next job of this node: NP scans all class lines, the pre-emptive disciplines start at the
                // finished class
                int kk = -1;
                for (int q = K - 1; q >= (sh_prty[node] == MQ_PRTY_NP ? 0 : kc); --q)
                    kk = IQ(node * K + q, QSIZE) != 0 ? q : kk;
                if (kk >= 0) {
                    const int nkk = node * K + kk;
                    const int slot = IQ(nkk, QHEAD);
                    const int qs = IQ(nkk, QSIZE) - 1;
                    IQ(nkk, QSIZE) = qs;
                    IQ(nkk, QHEAD) = (qs == 0 || slot + 1 >= P.qcap) ? 0 : slot + 1;
                    node_count(nkk, MQ_PNN_TAKED);
                    const double *qr = qrec(nkk, slot);
                    const double q_arrn = qr[0], q_waitn = qr[QL];
                    const double q_arrt = qr[2 * QL], q_sw = qr[3 * QL];
                    const double q_wt = qr[4 * QL], q_tte = qr[5 * QL];
                    const int pk = (int)qr[6 * QL];
                    const double d = __dsub_rn(ttek, q_sw);
                    if (pk >> 16) {
                        wval = __dadd_rn(ttek, q_tte);  // servers.py:237-239: resume with the stored time
                    } else {
                        const double sv = next_service(nkk);
                        SD(D_TOT + ev) = sv;
                        wval = __dadd_rn(ttek, sv);
                    }
                    SD(D_ARRN + ev) = q_arrn;
                    SD(D_WAITN + ev) = __dadd_rn(q_waitn, d);
                    SD(D_ARRT + ev) = q_arrt;
                    SD(D_WAITT + ev) = __dadd_rn(q_wt, d);
                    k_pack = (k_pack & ~(3u << (2 * ev))) | ((uint32_t)(pk & 255) << (2 * ev));
                    nc_pack = (nc_pack & ~(3u << (2 * ev))) | ((uint32_t)((pk >> 8) & 255) << (2 * ev));
                    busy |= 1u << ev;
                }
            }
#pragma unroll
            for (int s = 0; s < ST; ++s)
                if (s == wslot) end[s] = wval;

            // choose_next_node priority_network.py:137-150: per-class routing matrix
            const double pu = next_uniform();
            int next = -1;
            {
                // all columns at once, first hit from the bit mask (k5v2_net.cu): no chain of dependent selects
                const double *row = sh_R + real * RS + (cur + 1) * (m + 1);
                uint32_t hit = 0;
#pragma unroll
                for (int i = 0; i <= 8; ++i)
                    if (i <= m) hit |= (row[i] > pu ? 1u : 0u) << i;
                next = hit ? __ffs(hit) - 1 : 0;
            }
            if (status) break;
            if (!is_dep && next >= m) {
                status = MQ_E_INVALID;
                break;
            }

            if (next == m) {
                // the job leaves the network: priority_network.py:226-234
                const int n_out = SI(I_CLS + 2 * real + 1) + 1;
                SI(I_CLS + 2 * real + 1) = n_out;
                served_total++;
                const double v = __dsub_rn(ttek, j_arrn);
                if (REPLAY) {
                    const double sv = (double)n_out;
                    const double factor = __dsub_rn(1.0, __ddiv_rn(1.0, sv));
                    const double pv[3] = {v, __dmul_rn(v, v), k7v_cube(v)};
                    const double pw[3] = {j_waitn, __dmul_rn(j_waitn, j_waitn), k7v_cube(j_waitn)};
                    double *b = &SD(D_NET + 12 * real);
                    double sv_pw = v, sw_pw = j_waitn;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {  // priority_network.py:152-173
                        b[(6 + i) * K7V_BLOCK] = __dadd_rn(__dmul_rn(b[(6 + i) * K7V_BLOCK], factor), __ddiv_rn(pv[i], sv));
                        b[(9 + i) * K7V_BLOCK] = __dadd_rn(__dmul_rn(b[(9 + i) * K7V_BLOCK], factor), __ddiv_rn(pw[i], sv));
                        b[i * K7V_BLOCK] = __dadd_rn(b[i * K7V_BLOCK], sv_pw);
                        b[(3 + i) * K7V_BLOCK] = __dadd_rn(b[(3 + i) * K7V_BLOCK], sw_pw);
                        sv_pw = __dmul_rn(sv_pw, v);
                        sw_pw = __dmul_rn(sw_pw, j_waitn);
                    }
                } else {
                    double *g = rcls0 + (long long)(real * MQ_PNC_ROWS) * R;
                    double pv = v, pw = j_waitn;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        atomicAdd(g + (long long)(MQ_PNC_VSUM + i) * R, pv);
                        atomicAdd(g + (long long)(MQ_PNC_WSUM + i) * R, pw);
                        pv = __dmul_rn(pv, v);
                        pw = __dmul_rn(pw, j_waitn);
                    }
                }
            } else {
                // the job arrives at node `next` as node class kc: sim/priority.py:282-332 with (k, moment, ts)
                const int node = next, kc = sh_map[next * K + real], nk = node * K + kc;
                const int b = sh_base[node], ch = sh_ch[node], prty = sh_prty[node];
                const uint32_t full_mask = (1u << ch) - 1u;
                const uint32_t mask = (busy >> b) & full_mask;
                const bool full = mask == full_mask;
                // arrive_with_absolute_prty sim/priority.py:390-455: the FIRST server of the node with a weaker class
                int vj = -1;
                for (int j = ch - 1; j >= 0; --j) vj = (int)((nc_pack >> (2 * (b + j))) & 3u) > kc ? b + j : vj;
                const bool pre_found = full && prty != MQ_PRTY_NP && vj >= 0;
                const bool to_line = full && !pre_found;
                // record that joins a waiting line in this trip: the victim, or the arriving job
                if (pre_found || to_line) {
                    int p_nk = nk, p_pack = real | (kc << 8);
                    double p_arrn = j_arrn, p_waitn = j_waitn, p_arrt = ttek, p_wt = 0.0, p_tte = 0.0;
                    if (pre_found) {
                        const int vk = (int)((k_pack >> (2 * vj)) & 3u), vc = (int)((nc_pack >> (2 * vj)) & 3u);
                        p_nk = node * K + vc;
                        p_pack = vk | (vc << 8) | (1 << 16);
                        p_arrn = SD(D_ARRN + vj);
                        p_waitn = SD(D_WAITN + vj);
                        p_arrt = SD(D_ARRT + vj);
                        p_wt = SD(D_WAITT + vj);
                        double e_v = 0.0;
#pragma unroll
                        for (int s = 0; s < ST; ++s)
                            if (s == vj) e_v = end[s];
                        if (prty == MQ_PRTY_PR)
                            p_tte = __dsub_rn(e_v, ttek);
                        else if (prty == MQ_PRTY_RS)
                            p_tte = next_service(nk);  // drawn from the ARRIVING class's stream, before the new job's
                        else
                            p_tte = SD(D_TOT + vj);
                    }
                    const int qs = IQ(p_nk, QSIZE);
                    if (qs >= P.qcap) {
                        status = MQ_E_OVERFLOW;
                    } else {
                        int slot = IQ(p_nk, QHEAD) + qs;
                        if (slot >= P.qcap) slot -= P.qcap;
                        double *qr = qrec(p_nk, slot);
                        qr[0] = p_arrn;
                        qr[QL] = p_waitn;
                        qr[2 * QL] = p_arrt;
                        qr[3 * QL] = ttek;  // start_waiting_time
                        qr[4 * QL] = p_wt;
                        qr[5 * QL] = p_tte;
                        qr[6 * QL] = (double)p_pack;
                        IQ(p_nk, QSIZE) = qs + 1;
                    }
                }
                if (!full || pre_found) {
                    const int s = full ? vj : b + __ffs(~mask) - 1;  // lowest free channel, or the victim's
                    node_count(nk, MQ_PNN_TAKED);
                    const double sv = next_service(nk);
                    const double e = __dadd_rn(ttek, sv);
#pragma unroll
                    for (int q = 0; q < ST; ++q)
                        if (q == s) end[q] = e;
                    SD(D_TOT + s) = sv;
                    SD(D_ARRN + s) = j_arrn;
                    SD(D_WAITN + s) = j_waitn;
                    SD(D_ARRT + s) = ttek;
                    SD(D_WAITT + s) = 0.0;
                    k_pack = (k_pack & ~(3u << (2 * s))) | ((uint32_t)real << (2 * s));
                    nc_pack = (nc_pack & ~(3u << (2 * s))) | ((uint32_t)kc << (2 * s));
                    busy |= 1u << s;
                }
            }
        }

        rec[(long long)MQ_PN_G_TTEK * R] = ttek;
        rec[(long long)MQ_PN_G_FLAGS * R] = (double)(status ? -status : 0);
        rec[(long long)MQ_PN_G_UUSED * R] = (double)u_idx;
        for (int k = 0; k < K; ++k) {
            double *rc = rcls0 + (long long)(k * MQ_PNC_ROWS) * R;
            const double n_out = (double)SI(I_CLS + 2 * k + 1);
            if (REPLAY) {
                for (int q = 0; q < 3; ++q) {
                    rc[(long long)(MQ_PNC_VSUM + q) * R] = SD(D_NET + 12 * k + q);
                    rc[(long long)(MQ_PNC_WSUM + q) * R] = SD(D_NET + 12 * k + 3 + q);
                    rc[(long long)(MQ_PNC_VRUN + q) * R] = SD(D_NET + 12 * k + 6 + q);
                    rc[(long long)(MQ_PNC_WRUN + q) * R] = SD(D_NET + 12 * k + 9 + q);
                }
            }
            rc[(long long)MQ_PNC_SERVED * R] = n_out;
            rc[(long long)MQ_PNC_AUSED * R] = (double)SI(I_CLS + 2 * k);
        }
        for (int i = 0; i < m * K; ++i) {
            double *rn = rnode0 + (long long)(i * MQ_PNN_ROWS) * R;
            if (REPLAY) {
                for (int q = 0; q < 4; ++q) {
                    rn[(long long)(MQ_PNN_VSUM + q) * R] = SD(D_NODE + 16 * i + q);
                    rn[(long long)(MQ_PNN_WSUM + q) * R] = SD(D_NODE + 16 * i + 4 + q);
                    rn[(long long)(MQ_PNN_VRUN + q) * R] = SD(D_NODE + 16 * i + 8 + q);
                    rn[(long long)(MQ_PNN_WRUN + q) * R] = SD(D_NODE + 16 * i + 12 + q);
                }
            }
            rn[(long long)MQ_PNN_SERVED * R] = (double)IQ(i, NSERVED);
            rn[(long long)MQ_PNN_SUSED * R] = (double)IQ(i, SIDX);
        }
    }
#undef SD
#undef SI
#undef IQ
}

template <int ST>
static cudaError_t launch_one(const K7Params &P, int grid, cudaStream_t st)
{
    cudaError_t e;
    if (P.replay) {
        const size_t smem = K7VLayout<true>::bytes(P.nserv, P.m, P.K);
        e = cudaFuncSetAttribute(k7v2_prionet<ST, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k7v2_prionet<ST, true><<<grid, K7V_BLOCK, smem, st>>>(P);
    } else {
        const size_t smem = K7VLayout<false>::bytes(P.nserv, P.m, P.K);
        e = cudaFuncSetAttribute(k7v2_prionet<ST, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k7v2_prionet<ST, false><<<grid, K7V_BLOCK, smem, st>>>(P);
    }
    return cudaGetLastError();
}

template <int ST>
static cudaError_t occupancy_one(const K7Params &P, int *occ)
{
    if (P.replay) {
        const size_t smem = K7VLayout<true>::bytes(P.nserv, P.m, P.K);
        cudaError_t e = cudaFuncSetAttribute(k7v2_prionet<ST, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k7v2_prionet<ST, true>, K7V_BLOCK, smem);
    }
    const size_t smem = K7VLayout<false>::bytes(P.nserv, P.m, P.K);
    cudaError_t e = cudaFuncSetAttribute(k7v2_prionet<ST, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k7v2_prionet<ST, false>, K7V_BLOCK, smem);
}

bool k7v2_supported(int m, int K, int nserv)
{
    if (!(m >= 1 && m <= K7_MAXN && K >= 1 && K <= K7_MAXK && nserv >= 1 && nserv <= 16)) return false;
    return K7VLayout<true>::bytes(nserv, m, K) <= 200 * 1024;  // the replay form's moments must fit on chip
}

cudaError_t k7v2_launch(const K7Params &P, int grid, cudaStream_t st)
{
    return P.nserv <= 8 ? launch_one<8>(P, grid, st) : launch_one<16>(P, grid, st);
}

cudaError_t k7v2_occupancy(const K7Params &P, int *blocks_per_sm)
{
    return P.nserv <= 8 ? occupancy_one<8>(P, blocks_per_sm) : occupancy_one<16>(P, blocks_per_sm);
}

}  // namespace mq
