// k5_net.cu -- K5: open network of FIFO multi-server nodes with Markov routing, one lane per
// replication.
//
// Walks the event sequence of NetworkSimulator.run (most_queue/sim/networks/network.py:196-229):
//   event scan      base_network_sim.py:158-227  external arrival first, then BUSY servers in node-major,
//                                                channel-minor order; the first minimum wins
//   on_arrival      network.py:158-172           next external gap is drawn BEFORE the routing uniform
//   on_serving      network.py:174-194           node.serving() -- which pops the queue head and draws
//                                                its service -- runs BEFORE the routing uniform
//   choose_next_node network.py:102-114          first column whose running row sum exceeds the uniform
//   node arrival / serving  sim/base.py:214-247, 249-310 (infinite buffer, lowest free server index)
//   network moments network.py:116-135           3 running moments with math.pow powers
//
// fp64 throughout with explicit round-to-nearest intrinsics in the reference's operation order, so
// in REPLAY mode per-job network sojourn / wait values, their power sums, node-level running
// moments, counters and the clock are bit-identical to the reference.  (The network-level running
// moments use pow(x, 2), pow(x, 3) in the reference; here x*x is exact and the cube is evaluated in
// double-double and rounded once, which matches a correctly rounded pow.)
//
// Generator mode consumes the virtual arrays of the counter-based generator: node i's service
// draws = word 1 of stream sid=i; external gaps = word 0 and routing uniforms = word 1 of stream
// sid=m.
#include "k5_net.cuh"

namespace mq {

struct K5Job {
    double arr_network, wait_network, arr_time;
};

__device__ __forceinline__ void k5_refresh4(double *m, double x, long long count)
{
    double power = x;
    const double inv_count = __ddiv_rn(1.0, (double)count);
    const double one_minus = __dsub_rn(1.0, inv_count);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m[i] = __dadd_rn(__dmul_rn(m[i], one_minus), __dmul_rn(power, inv_count));
        power = __dmul_rn(power, x);
    }
}

__device__ __forceinline__ double k5_cube(double x)
{
    // x^3 rounded once: (p + e) * x with p + e = x * x exactly
    const double p = __dmul_rn(x, x);
    const double e = __fma_rn(x, x, -p);
    const double q = __dmul_rn(p, x);
    const double e2 = __fma_rn(p, x, -q);
    return __dadd_rn(q, __fma_rn(e, x, e2));
}

// network.py:116-135
__device__ __forceinline__ void k5_refresh_net(double *m, double x, long long served)
{
    const double sv = (double)served;
    const double factor = __dsub_rn(1.0, __ddiv_rn(1.0, sv));
    const double pw[3] = {x, __dmul_rn(x, x), k5_cube(x)};
#pragma unroll
    for (int i = 0; i < 3; ++i) m[i] = __dadd_rn(__dmul_rn(m[i], factor), __ddiv_rn(pw[i], sv));
}

__device__ __forceinline__ void k5_addk(double *s, double x, int n)
{
    double pw = x;
    for (int i = 0; i < n; ++i) {
        s[i] = __dadd_rn(s[i], pw);
        pw = __dmul_rn(pw, x);
    }
}

__global__ void __launch_bounds__(K5_BLOCK) k5_net(const K5Params P)
{
    const long long gtid = (long long)blockIdx.x * K5_BLOCK + threadIdx.x;
    const long long R = P.reps;
    const int m = P.m;

    for (long long rl = gtid; rl < R; rl += P.lanes) {
        const uint32_t rep = P.rep0 + (uint32_t)rl;
        double *rec = P.rec + rl;
        double *Q = P.queue + gtid;

        double end[K5_MAXS];
        K5Job on[K5_MAXS];
        unsigned long long busy = 0ull;
        int free_ch[K5_MAXN], qhead[K5_MAXN], qsize[K5_MAXN];
        long long n_arrived[K5_MAXN], n_taked[K5_MAXN], n_served[K5_MAXN], n_insys[K5_MAXN], s_idx[K5_MAXN];
        double nv_run[K5_MAXN][4], nw_run[K5_MAXN][4], nv_sum[K5_MAXN][4], nw_sum[K5_MAXN][4];
        for (int i = 0; i < m; ++i) {
            free_ch[i] = P.channels[i];
            qhead[i] = qsize[i] = 0;
            n_arrived[i] = n_taked[i] = n_served[i] = n_insys[i] = s_idx[i] = 0;
            for (int k = 0; k < 4; ++k) nv_run[i][k] = nw_run[i][k] = nv_sum[i][k] = nw_sum[i][k] = 0.0;
        }
        for (int s = 0; s < P.nserv; ++s) end[s] = 1e10;
        double v_run[3] = {0, 0, 0}, w_run[3] = {0, 0, 0}, v_sum[3] = {0, 0, 0}, w_sum[3] = {0, 0, 0};
        long long served = 0, arrived = 0, in_sys = 0, a_idx = 0, u_idx = 0;
        int status = 0;
        const uint32_t key_m = P.key0 + MQ_KEY_STRIDE * (64u * (uint32_t)m);

        auto next_gap = [&]() -> double {
            const long long i = a_idx++;
            if (P.replay) {
                if (i >= P.n_a) {
                    status = MQ_E_INPUT_SHORT;
                    return 0.0;
                }
                return P.A[i * R + rl];
            }
            uint2 w = philox2x32_10((uint32_t)i, rep, key_m);
            return (double)sample<-1>(P.src, w.x, (uint32_t)i, rep, key_m, 0);
        };
        auto next_uniform = [&]() -> double {
            const long long i = u_idx++;
            if (P.replay) {
                if (i >= P.n_u) {
                    status = MQ_E_INPUT_SHORT;
                    return 0.0;
                }
                return P.U[i * R + rl];
            }
            uint2 w = philox2x32_10((uint32_t)i, rep, key_m);
            return ((double)w.y + 0.5) * 0x1p-32;
        };
        auto next_service = [&](int node) -> double {
            const long long j = s_idx[node]++;
            if (P.replay) {
                if (j >= P.n_s[node]) {
                    status = MQ_E_INPUT_SHORT;
                    return 0.0;
                }
                return P.S[P.soff[node] + j * R + rl];
            }
            const uint32_t key = P.key0 + MQ_KEY_STRIDE * (64u * (uint32_t)node);
            uint2 w = philox2x32_10((uint32_t)j, rep, key);
            return (double)sample<-1>(P.srv[node], w.y, (uint32_t)j, rep, key, 1);
        };
        auto qslot = [&](int node, int slot, int f) -> double * {
            return Q + (((long long)node * P.qcap + slot) * 3 + f) * P.lanes;
        };
        // choose_next_node network.py:102-114
        auto choose_next = [&](int current, double p) -> int {
            double sum_p = 0.0;
            for (int i = 0; i < m + 1; ++i) {
                sum_p = __dadd_rn(sum_p, P.R[(current + 1) * (m + 1) + i]);
                if (sum_p > p) return i;
            }
            return 0;
        };
        double ttek = 0.0;
        // node arrival, sim/base.py:214-247 with (moment, ts)
        auto node_arrival = [&](int node, K5Job job) {
            n_arrived[node]++;
            job.arr_time = ttek;
            n_insys[node]++;
            if (free_ch[node] == 0) {
                if (qsize[node] >= P.qcap) {
                    status = MQ_E_OVERFLOW;
                    return;
                }
                int slot = qhead[node] + qsize[node];
                if (slot >= P.qcap) slot -= P.qcap;
                *qslot(node, slot, 0) = job.arr_network;
                *qslot(node, slot, 1) = job.wait_network;
                *qslot(node, slot, 2) = job.arr_time;  // == start_waiting_time (base.py:199-201)
                qsize[node]++;
            } else {
                const int b = P.base[node];
                const unsigned long long mask = (busy >> b) & ((1ull << P.channels[node]) - 1ull);
                const int c = __ffsll((long long)~mask) - 1;  // lowest free channel
                n_taked[node]++;
                k5_refresh4(nw_run[node], 0.0, n_taked[node]);
                k5_addk(nw_sum[node], 0.0, 4);
                end[b + c] = __dadd_rn(ttek, next_service(node));
                on[b + c] = job;
                busy |= 1ull << (b + c);
                free_ch[node]--;
            }
        };

        double arrival_time = next_gap();  // network.py:76

        while (served < P.jobs && status == 0) {
            // base_network_sim.py:184-227
            int ev = -1;
            double et = arrival_time;
            for (int s = 0; s < P.nserv; ++s) {
                if (((busy >> s) & 1ull) && end[s] < et) {
                    et = end[s];
                    ev = s;
                }
            }
            if (ev < 0) {
                // on_arrival network.py:158-172
                ttek = arrival_time;
                arrived++;
                in_sys++;
                arrival_time = __dadd_rn(ttek, next_gap());
                const int next = choose_next(-1, next_uniform());
                if (status) break;
                if (next >= m) {
                    status = MQ_E_INVALID;  // the reference would index qs[m]
                    break;
                }
                K5Job ts;
                ts.arr_network = ttek;
                ts.wait_network = 0.0;
                ts.arr_time = ttek;
                node_arrival(next, ts);
            } else {
                // on_serving network.py:174-194
                int node = 0;
                while (ev >= P.base[node + 1]) ++node;
                ttek = end[ev];
                const K5Job ts = on[ev];
                end[ev] = 1e10;
                busy &= ~(1ull << ev);
                n_served[node]++;
                free_ch[node]++;
                const double vn = __dsub_rn(ttek, ts.arr_time);
                k5_refresh4(nv_run[node], vn, n_served[node]);
                k5_addk(nv_sum[node], vn, 4);
                n_insys[node]--;
                if (qsize[node] != 0) {
                    // send_head_of_queue_to_channel base.py:288-310
                    const int slot = qhead[node];
                    K5Job qj;
                    qj.arr_network = *qslot(node, slot, 0);
                    qj.wait_network = *qslot(node, slot, 1);
                    qj.arr_time = *qslot(node, slot, 2);
                    qsize[node]--;
                    // an emptied line restarts at slot 0 (keeps the ring's footprint cache-resident)
                    qhead[node] = (qsize[node] == 0 || slot + 1 >= P.qcap) ? 0 : slot + 1;
                    n_taked[node]++;
                    const double d = __dsub_rn(ttek, qj.arr_time);
                    const double wt = __dadd_rn(0.0, d);
                    qj.wait_network = __dadd_rn(qj.wait_network, d);
                    k5_refresh4(nw_run[node], wt, n_taked[node]);
                    k5_addk(nw_sum[node], wt, 4);
                    end[ev] = __dadd_rn(ttek, next_service(node));
                    on[ev] = qj;
                    busy |= 1ull << ev;
                    free_ch[node]--;
                }
                const int next = choose_next(node, next_uniform());
                if (status) break;
                if (next == m) {
                    served++;
                    in_sys--;
                    const double v = __dsub_rn(ttek, ts.arr_network);
                    k5_refresh_net(v_run, v, served);
                    k5_refresh_net(w_run, ts.wait_network, served);
                    k5_addk(v_sum, v, 3);
                    k5_addk(w_sum, ts.wait_network, 3);
                } else {
                    node_arrival(next, ts);
                }
            }
        }

        rec[(long long)MQ_NG_TTEK * R] = ttek;
        rec[(long long)MQ_NG_FLAGS * R] = (double)(status ? -status : 0);
        rec[(long long)MQ_NG_SERVED * R] = (double)served;
        rec[(long long)MQ_NG_ARRIVED * R] = (double)arrived;
        rec[(long long)MQ_NG_INSYS * R] = (double)in_sys;
        rec[(long long)MQ_NG_AUSED * R] = (double)a_idx;
        rec[(long long)MQ_NG_UUSED * R] = (double)u_idx;
        for (int k = 0; k < 3; ++k) {
            rec[(long long)(MQ_NG_VSUM + k) * R] = v_sum[k];
            rec[(long long)(MQ_NG_WSUM + k) * R] = w_sum[k];
            rec[(long long)(MQ_NG_VRUN + k) * R] = v_run[k];
            rec[(long long)(MQ_NG_WRUN + k) * R] = w_run[k];
        }
        for (int i = 0; i < m; ++i) {
            double *rn = rec + (long long)(MQ_NG_ROWS + i * MQ_NN_ROWS) * R;
            for (int k = 0; k < 4; ++k) {
                rn[(long long)(MQ_NN_VSUM + k) * R] = nv_sum[i][k];
                rn[(long long)(MQ_NN_WSUM + k) * R] = nw_sum[i][k];
                rn[(long long)(MQ_NN_VRUN + k) * R] = nv_run[i][k];
                rn[(long long)(MQ_NN_WRUN + k) * R] = nw_run[i][k];
            }
            rn[(long long)MQ_NN_ARRIVED * R] = (double)n_arrived[i];
            rn[(long long)MQ_NN_TAKED * R] = (double)n_taked[i];
            rn[(long long)MQ_NN_SERVED * R] = (double)n_served[i];
            rn[(long long)MQ_NN_SUSED * R] = (double)s_idx[i];
            rn[(long long)MQ_NN_INSYS * R] = (double)n_insys[i];
            rn[(long long)MQ_NN_QUEUED * R] = (double)qsize[i];
        }
    }
}

cudaError_t k5_launch(const K5Params &P, int grid, cudaStream_t st)
{
    k5_net<<<grid, K5_BLOCK, 0, st>>>(P);
    return cudaGetLastError();
}

cudaError_t k5_occupancy(int *blocks_per_sm)
{
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k5_net, K5_BLOCK, 0);
}

}  // namespace mq
