// k5v2_net.cu -- K5 v2: the open-network kernel as ONE predicated trip body with the per-lane state in registers
// and shared memory, for networks of at most 16 servers in total and at most 8 nodes.
//
// Same event sequence, same arithmetic and same record rows as k5_net.cu (which stays the kernel for larger
// networks; see its header for the reference map, most_queue/sim/networks/network.py).  Shape of the code:
//   * a trip is an external arrival or a service completion; both run through one straight-line body with one
//     uniform-draw site (routing), one "pop the node's waiting line and start it" site and one "a job arrives at a
//     node" site (the latter shared by external arrivals and routed jobs);
//   * completion times of all servers are registers (pairwise-tree argmin, node-major order, first minimum wins);
//     the three time stamps of the job on each server and the per-node cursors are shared memory [word][lane];
//   * generator tier: node-level and network-level power sums and the counters are added straight into the zeroed
//     per-replication record with fire-and-forget RED.ADD.F64 (one lane, one address: same order, same rounding),
//     and node service draws are produced up to K5V_SB ahead per node by all lanes of the warp in lock-step;
//   * replay tier: the reference's running moments need read-modify-write, so all moments live in shared memory.
#include "k5_net.cuh"

namespace mq {

constexpr int K5V_BLOCK = 64;
constexpr int K5V_MAXN = 8;
constexpr int K5V_SB = 4;  // service draws kept ahead per node and lane (generator tier; power of two)

__device__ __forceinline__ double k5v_cube(double x)
{
    // x^3 rounded once: (p + e) * x with p + e = x * x exactly
    const double p = __dmul_rn(x, x);
    const double e = __fma_rn(x, x, -p);
    const double q = __dmul_rn(p, x);
    const double e2 = __fma_rn(p, x, -q);
    return __dadd_rn(q, __fma_rn(e, x, e2));
}

// shared-memory words per lane
template <bool REPLAY>
struct K5VLayout {
    // doubles: 3 per server, then (replay) 16 per node + 12 for the network
    __host__ __device__ static int nd(int nserv, int m) { return 3 * nserv + (REPLAY ? 16 * m + 12 : 0); }
    // ints per node: QHEAD, QSIZE, SIDX, SFILL, NSERVED (+ NTAKED in replay)
    static constexpr int ISTR = REPLAY ? 6 : 5;
    __host__ __device__ static int ni(int m) { return ISTR * m; }
    __host__ __device__ static int nf(int m) { return REPLAY ? 0 : K5V_SB * m; }
    __host__ __device__ static size_t bytes(int nserv, int m)
    {
        return (size_t)K5V_BLOCK * (8 * (size_t)nd(nserv, m) + 4 * (size_t)ni(m) + 4 * (size_t)nf(m));
    }
};

// SRCK / SRVK: the law kinds of the external flow and of ALL nodes when they are known to be one kind (-1: switch on the
// kind at run time)
template <int ST, bool REPLAY, int SRCK = -1, int SRVK = -1>
__global__ void __launch_bounds__(K5V_BLOCK) k5v2_net(const K5Params P)
{
    using LY = K5VLayout<REPLAY>;
    extern __shared__ double k5v_sm[];
    __shared__ DistF sh_srv[K5V_MAXN];
    __shared__ double sh_R[(K5V_MAXN + 1) * (K5V_MAXN + 1)];
    __shared__ int sh_base[K5V_MAXN + 1], sh_ch[K5V_MAXN], sh_nodeof[16];
    const int tid = threadIdx.x;
    const int m = P.m, nserv = P.nserv;
    // sh_R holds the RUNNING row sums of the routing matrix, added left to right exactly as choose_next_node does
    // (network.py:102-114), so a trip only compares
    if (tid <= m) {
        double sum_p = 0.0;
        for (int i = 0; i <= m; ++i) {
            sum_p = __dadd_rn(sum_p, P.R[tid * (m + 1) + i]);
            sh_R[tid * (m + 1) + i] = sum_p;
        }
    }
    if (tid < 16) {
        int node = 0;
        for (int i = 1; i < m; ++i) node += tid >= P.base[i];
        sh_nodeof[tid] = node;
    }
    if (tid <= m) sh_base[tid] = P.base[tid];
    if (tid < m) {
        sh_ch[tid] = P.channels[tid];
        if (!REPLAY) sh_srv[tid] = P.srv[tid];
    }
    __syncthreads();

    const int ND = LY::nd(nserv, m), NI = LY::ni(m);
    double *smd = k5v_sm + tid;
    int *smi = reinterpret_cast<int *>(k5v_sm + ND * K5V_BLOCK) + tid;
    float *smf = reinterpret_cast<float *>(smi - tid + NI * K5V_BLOCK) + tid;
#define SD(w) smd[(w) * K5V_BLOCK]
#define SI(w) smi[(w) * K5V_BLOCK]
#define IN(node, f) SI((node) * LY::ISTR + (f))
    enum { QHEAD = 0, QSIZE = 1, SIDX = 2, SFILL = 3, NSERVED = 4, NTAKED = 5 };
    const int D_ARRN = 0, D_WAITN = nserv, D_ARRT = 2 * nserv, D_NODE = 3 * nserv, D_NET = 3 * nserv + 16 * m;
    // replay: per node VSUM 0..3, WSUM 4..7, VRUN 8..11, WRUN 12..15; network VSUM 0..2, WSUM 3..5, VRUN 6..8, WRUN 9..11

    const long long gtid = (long long)blockIdx.x * K5V_BLOCK + tid;
    const long long R = P.reps;
    const DistF src_law = P.src;  // by value: taking the address of a kernel parameter would spill all of P

    for (long long rl = gtid; rl < R; rl += P.lanes) {
        const uint32_t rep = P.rep0 + (uint32_t)rl;
        double *rec = P.rec + rl;
        double *Q = P.queue + gtid;
        for (int w = 0; w < ND; ++w) SD(w) = 0.0;
        for (int w = 0; w < NI; ++w) SI(w) = 0;

        double end[ST];
#pragma unroll
        for (int s = 0; s < ST; ++s) end[s] = s < nserv ? 1e10 : 1e300;
        uint32_t busy = 0;
        int status = 0, a_idx = 0, u_idx = 0;
        uint32_t need = 1u;  // generator tier: some ring of this lane is (nearly) empty
        long long served = 0, arrived = 0;
        const uint32_t key_m = P.key0 + MQ_KEY_STRIDE * (64u * (uint32_t)m);

        auto next_gap = [&]() -> double {
            const int i = a_idx++;
            if (REPLAY) {
                if (i >= P.n_a) {
                    status = MQ_E_INPUT_SHORT;
                    return 0.0;
                }
                return P.A[(long long)i * R + rl];
            }
            uint2 w = philox2x32_10((uint32_t)i, rep, key_m);
            return (double)sample<SRCK>(src_law, w.x, (uint32_t)i, rep, key_m, 0);
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
            uint2 w = philox2x32_10((uint32_t)i, rep, key_m);
            return ((double)w.y + 0.5) * 0x1p-32;
        };
        // generator tier: when any lane of the warp is short of draws for some node (fewer than the two starts a
        // trip can make), all lanes top up all their rings to K5V_SB together
        auto refill_services = [&]() {
            if (!__any_sync(__activemask(), need != 0u)) return;
            need = 0u;
            for (int node = 0; node < m; ++node) {
                const uint32_t key = P.key0 + MQ_KEY_STRIDE * (64u * (uint32_t)node);
                const DistF d = sh_srv[node];
                for (;;) {
                    const int fill = IN(node, SFILL);
                    const bool more = fill < K5V_SB;
                    if (!__any_sync(__activemask(), more)) break;
                    if (more) {
                        const uint32_t j = (uint32_t)(IN(node, SIDX) + fill);
                        uint2 w = philox2x32_10(j, rep, key);
                        smf[(node * K5V_SB + (int)(j & (K5V_SB - 1))) * K5V_BLOCK] = sample<SRVK>(d, w.y, j, rep, key, 1);
                        IN(node, SFILL) = fill + 1;
                    }
                }
            }
        };
        auto next_service = [&](int node) -> double {
            const int j = IN(node, SIDX);
            IN(node, SIDX) = j + 1;
            if (REPLAY) {
                if (j >= P.n_s[node]) {
                    status = MQ_E_INPUT_SHORT;
                    return 0.0;
                }
                return P.S[P.soff[node] + (long long)j * R + rl];
            }
            const int fill = IN(node, SFILL) - 1;
            IN(node, SFILL) = fill;
            need |= fill < 2 ? 1u : 0u;
            return (double)smf[(node * K5V_SB + (j & (K5V_SB - 1))) * K5V_BLOCK];
        };
        const long long QL = P.lanes;  // field stride of a waiting-line record
        auto qrec = [&](int node, int slot) -> double * { return Q + (long long)((node * P.qcap + slot) * 3) * QL; };
        double *rnode0 = rec + (long long)MQ_NG_ROWS * R;
        // a sample of a node-level moment set: row0 = MQ_NN_VSUM or MQ_NN_WSUM
        auto node_sample = [&](int node, int row0, double x, int count) {
            if (REPLAY) {
                double *sum = &SD(D_NODE + 16 * node + row0), *run = sum + 8 * K5V_BLOCK;
                const double inv_count = __drcp_rn((double)count);  // == 1.0 / count, correctly rounded
                const double one_minus = __dsub_rn(1.0, inv_count);
                double pw = x;
#pragma unroll
                for (int i = 0; i < 4; ++i) {  // refresh_moments_stat, stats_update.py:6-22
                    run[i * K5V_BLOCK] = __dadd_rn(__dmul_rn(run[i * K5V_BLOCK], one_minus), __dmul_rn(pw, inv_count));
                    sum[i * K5V_BLOCK] = __dadd_rn(sum[i * K5V_BLOCK], pw);
                    pw = __dmul_rn(pw, x);
                }
            } else if (x != 0.0) {
                // (a zero sample -- the wait of a job that finds a free server -- leaves the non-negative sums as they
                // are, bit for bit: no atomics for it)
                double *g = rnode0 + (long long)(node * MQ_NN_ROWS + row0) * R;
                double pw = x;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    atomicAdd(g, pw);
                    g += R;
                    pw = __dmul_rn(pw, x);
                }
            }
        };
        // counters that are only reported.  Replay tier: one atomic each.  Generator tier: only the departures of a
        // node are counted (an int in shared memory); taken = served + jobs in service and arrived = taken + jobs
        // waiting are derived when the replication ends
        auto node_count = [&](int node, int row, double d) {
            if (REPLAY) atomicAdd(rnode0 + (long long)(node * MQ_NN_ROWS + row) * R, d);
        };

        double arrival_time = next_gap();  // network.py:76
        double ttek = 0.0;

        while (served < P.jobs && status == 0) {
            // ---- next event, base_network_sim.py:184-227: external arrival first, then the busy servers in
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
            const bool is_dep = mt < 1e10 && mt < arrival_time;
            ttek = is_dep ? mt : arrival_time;
            // the head of the departing node's waiting line is needed half a trip further down: load it NOW, before the
            // draw-ahead refill and the Philox chain of the routing uniform, so that its L2 round trip hides behind them
            const int dep_node = is_dep ? sh_nodeof[ev] : 0;
            const int dep_qs = is_dep ? IN(dep_node, QSIZE) : 0;
            const int dep_slot = is_dep ? IN(dep_node, QHEAD) : 0;
            double q_arrn = 0.0, q_waitn = 0.0, q_arrt = 0.0;
            if (dep_qs != 0) {
                const double *qr = qrec(dep_node, dep_slot);
                q_arrn = qr[0];
                q_waitn = qr[QL];
                q_arrt = qr[2 * QL];
            }
            if (!REPLAY) refill_services();
            // every trip consumes exactly one routing uniform and it depends on nothing else
            const double pu = next_uniform();

            // the job that moves in this trip, and where it is now (-1: the source)
            double j_arrn = ttek, j_waitn = 0.0;
            int cur = -1;
            int wslot = -1;   // the one server slot rewritten by the first half of the trip
            double wval = 1e10;

            if (!is_dep) {
                // on_arrival network.py:158-172
                arrived++;
                arrival_time = __dadd_rn(ttek, next_gap());
            } else {
                // on_serving network.py:174-194: node.serving() first (base.py:249-310)
                const int node = dep_node;
                cur = node;
                const int qs = dep_qs, slot = dep_slot;
                j_arrn = SD(D_ARRN + ev);
                j_waitn = SD(D_WAITN + ev);
                const double a_t = SD(D_ARRT + ev);
                busy &= ~(1u << ev);
                wslot = ev;
                const int n_served = IN(node, NSERVED) + 1;
                IN(node, NSERVED) = n_served;
                node_count(node, MQ_NN_SERVED, 1.0);
                node_sample(node, MQ_NN_VSUM, __dsub_rn(ttek, a_t), n_served);
                if (qs != 0) {
                    // send_head_of_queue_to_channel base.py:288-310
                    IN(node, QSIZE) = qs - 1;
                    IN(node, QHEAD) = (qs == 1 || slot + 1 >= P.qcap) ? 0 : slot + 1;
                    int n_taked = 0;
                    if (REPLAY) {
                        n_taked = IN(node, NTAKED) + 1;
                        IN(node, NTAKED) = n_taked;
                    }
                    node_count(node, MQ_NN_TAKED, 1.0);
                    const double d = __dsub_rn(ttek, q_arrt);
                    node_sample(node, MQ_NN_WSUM, __dadd_rn(0.0, d), n_taked);
                    wval = __dadd_rn(ttek, next_service(node));
                    SD(D_ARRN + ev) = q_arrn;
                    SD(D_WAITN + ev) = __dadd_rn(q_waitn, d);
                    SD(D_ARRT + ev) = q_arrt;
                    busy |= 1u << ev;
                }
            }
#pragma unroll
            for (int s = 0; s < ST; ++s)
                if (s == wslot) end[s] = wval;

            // choose_next_node network.py:102-114: first column whose running row sum exceeds the uniform
            int next = -1;
            {
                // all columns are compared at once (independent loads and compares) and the first hit is picked from the
                // bit mask: the same decision as the sequential scan without its chain of dependent selects
                const double *row = sh_R + (cur + 1) * (m + 1);
                uint32_t hit = 0;
#pragma unroll
                for (int i = 0; i <= K5V_MAXN; ++i)
                    if (i <= m) hit |= (row[i] > pu ? 1u : 0u) << i;
                next = hit ? __ffs(hit) - 1 : 0;
            }
            if (status) break;
            if (!is_dep && next >= m) {
                status = MQ_E_INVALID;  // the reference would index qs[m]
                break;
            }

            if (next == m) {
                // the job leaves the network: network.py:184-190
                served++;
                const double v = __dsub_rn(ttek, j_arrn);
                if (REPLAY) {
                    const double sv = (double)served;
                    const double factor = __dsub_rn(1.0, __ddiv_rn(1.0, sv));
                    const double pv[3] = {v, __dmul_rn(v, v), k5v_cube(v)};
                    const double pw[3] = {j_waitn, __dmul_rn(j_waitn, j_waitn), k5v_cube(j_waitn)};
                    double sv_pw = v, sw_pw = j_waitn;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {  // network.py:116-135
                        SD(D_NET + 6 + i) = __dadd_rn(__dmul_rn(SD(D_NET + 6 + i), factor), __ddiv_rn(pv[i], sv));
                        SD(D_NET + 9 + i) = __dadd_rn(__dmul_rn(SD(D_NET + 9 + i), factor), __ddiv_rn(pw[i], sv));
                        SD(D_NET + i) = __dadd_rn(SD(D_NET + i), sv_pw);
                        SD(D_NET + 3 + i) = __dadd_rn(SD(D_NET + 3 + i), sw_pw);
                        sv_pw = __dmul_rn(sv_pw, v);
                        sw_pw = __dmul_rn(sw_pw, j_waitn);
                    }
                } else {
                    double pv = v, pw = j_waitn;
                    const bool waited = j_waitn != 0.0;  // a zero sample leaves the sums as they are
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        atomicAdd(&rec[(long long)(MQ_NG_VSUM + i) * R], pv);
                        if (waited) atomicAdd(&rec[(long long)(MQ_NG_WSUM + i) * R], pw);
                        pv = __dmul_rn(pv, v);
                        pw = __dmul_rn(pw, j_waitn);
                    }
                }
            } else {
                // the job arrives at node `next`: sim/base.py:214-247 with (moment, ts)
                const int node = next;
                node_count(node, MQ_NN_ARRIVED, 1.0);
                const int b = sh_base[node], ch = sh_ch[node];
                const uint32_t mask = (busy >> b) & ((1u << ch) - 1u);
                if (mask == (1u << ch) - 1u) {
                    const int qs = IN(node, QSIZE);
                    if (qs >= P.qcap) {
                        status = MQ_E_OVERFLOW;
                    } else {
                        int slot = IN(node, QHEAD) + qs;
                        if (slot >= P.qcap) slot -= P.qcap;
                        double *qr = qrec(node, slot);
                        qr[0] = j_arrn;
                        qr[QL] = j_waitn;
                        qr[2 * QL] = ttek;  // == start_waiting_time (base.py:199-201)
                        IN(node, QSIZE) = qs + 1;
                    }
                } else {
                    const int s = b + __ffs(~mask) - 1;  // lowest free channel
                    int n_taked = 0;
                    if (REPLAY) {
                        n_taked = IN(node, NTAKED) + 1;
                        IN(node, NTAKED) = n_taked;
                    }
                    node_count(node, MQ_NN_TAKED, 1.0);
                    node_sample(node, MQ_NN_WSUM, 0.0, n_taked);
                    const double e = __dadd_rn(ttek, next_service(node));
#pragma unroll
                    for (int q = 0; q < ST; ++q)
                        if (q == s) end[q] = e;
                    SD(D_ARRN + s) = j_arrn;
                    SD(D_WAITN + s) = j_waitn;
                    SD(D_ARRT + s) = ttek;
                    busy |= 1u << s;
                }
            }
        }

        rec[(long long)MQ_NG_TTEK * R] = ttek;
        rec[(long long)MQ_NG_FLAGS * R] = (double)(status ? -status : 0);
        rec[(long long)MQ_NG_SERVED * R] = (double)served;
        rec[(long long)MQ_NG_ARRIVED * R] = (double)arrived;
        rec[(long long)MQ_NG_INSYS * R] = (double)(arrived - served);
        rec[(long long)MQ_NG_AUSED * R] = (double)a_idx;
        rec[(long long)MQ_NG_UUSED * R] = (double)u_idx;
        if (REPLAY) {
            for (int k = 0; k < 3; ++k) {
                rec[(long long)(MQ_NG_VSUM + k) * R] = SD(D_NET + k);
                rec[(long long)(MQ_NG_WSUM + k) * R] = SD(D_NET + 3 + k);
                rec[(long long)(MQ_NG_VRUN + k) * R] = SD(D_NET + 6 + k);
                rec[(long long)(MQ_NG_WRUN + k) * R] = SD(D_NET + 9 + k);
            }
        }
        for (int i = 0; i < m; ++i) {
            double *rn = rnode0 + (long long)(i * MQ_NN_ROWS) * R;
            if (REPLAY) {
                for (int k = 0; k < 4; ++k) {
                    rn[(long long)(MQ_NN_VSUM + k) * R] = SD(D_NODE + 16 * i + k);
                    rn[(long long)(MQ_NN_WSUM + k) * R] = SD(D_NODE + 16 * i + 4 + k);
                    rn[(long long)(MQ_NN_VRUN + k) * R] = SD(D_NODE + 16 * i + 8 + k);
                    rn[(long long)(MQ_NN_WRUN + k) * R] = SD(D_NODE + 16 * i + 12 + k);
                }
            }
            if (REPLAY) {
                // jobs inside the node = arrived - served (this lane's own atomics on those two rows; exact small integers)
                rn[(long long)MQ_NN_INSYS * R] = __dsub_rn(__ldcg(&rn[(long long)MQ_NN_ARRIVED * R]), __ldcg(&rn[(long long)MQ_NN_SERVED * R]));
            } else {
                const int in_service = __popc((busy >> sh_base[i]) & ((1u << sh_ch[i]) - 1u));
                const int n_served = IN(i, NSERVED), n_taked = n_served + in_service, n_arrived = n_taked + IN(i, QSIZE);
                rn[(long long)MQ_NN_SERVED * R] = (double)n_served;
                rn[(long long)MQ_NN_TAKED * R] = (double)n_taked;
                rn[(long long)MQ_NN_ARRIVED * R] = (double)n_arrived;
                rn[(long long)MQ_NN_INSYS * R] = (double)(n_arrived - n_served);
            }
            rn[(long long)MQ_NN_SUSED * R] = (double)IN(i, SIDX);
            rn[(long long)MQ_NN_QUEUED * R] = (double)IN(i, QSIZE);
        }
    }
#undef SD
#undef SI
#undef IN
}

using K5Kernel = void (*)(const K5Params);
// generator tier: own instantiations for networks whose nodes all serve with one law kind (M or H2) under Poisson arrivals
template <int ST>
static K5Kernel gen_kernel(const K5Params &P)
{
    int kind = P.srv[0].kind;
    for (int i = 1; i < P.m; ++i)
        if (P.srv[i].kind != kind) kind = -1;
    if (P.src.kind == MQ_M && kind == MQ_H2) return k5v2_net<ST, false, MQ_M, MQ_H2>;
    if (P.src.kind == MQ_M && kind == MQ_M) return k5v2_net<ST, false, MQ_M, MQ_M>;
    return k5v2_net<ST, false, -1, -1>;
}

template <int ST>
static cudaError_t launch_one(const K5Params &P, int grid, cudaStream_t st)
{
    const K5Kernel kern = P.replay ? (K5Kernel)k5v2_net<ST, true> : gen_kernel<ST>(P);
    const size_t smem = P.replay ? K5VLayout<true>::bytes(P.nserv, P.m) : K5VLayout<false>::bytes(P.nserv, P.m);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, K5V_BLOCK, smem, st>>>(P);
    return cudaGetLastError();
}

template <int ST>
static cudaError_t occupancy_one(const K5Params &P, int *occ)
{
    const K5Kernel kern = P.replay ? (K5Kernel)k5v2_net<ST, true> : gen_kernel<ST>(P);
    const size_t smem = P.replay ? K5VLayout<true>::bytes(P.nserv, P.m) : K5VLayout<false>::bytes(P.nserv, P.m);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, K5V_BLOCK, smem);
}

bool k5v2_supported(int m, int nserv) { return m >= 1 && m <= K5V_MAXN && nserv >= 1 && nserv <= 16; }

cudaError_t k5v2_launch(const K5Params &P, int grid, cudaStream_t st)
{
    return P.nserv <= 8 ? launch_one<8>(P, grid, st) : launch_one<16>(P, grid, st);
}

cudaError_t k5v2_occupancy(const K5Params &P, int *blocks_per_sm)
{
    return P.nserv <= 8 ? occupancy_one<8>(P, blocks_per_sm) : occupancy_one<16>(P, blocks_per_sm);
}

}  // namespace mq
