// k7_prionet.cu -- K7: open network of multi-class PRIORITY nodes with per-class Markov routing, one lane per
// replication.
//
// Walks PriorityNetwork.run (most_queue/sim/networks/priority_network.py:241-276) event for event:
//   event scan       base_network_sim.py:184-227   class 0 arrival < class 1 < ... < busy servers in node-major,
//                                                  channel-minor order; the first minimum wins
//   on_arrival(k)    priority_network.py:196-215   next gap drawn BEFORE the routing uniform; the job enters
//                                                  node `next` as node class nodes_prty[next][k]
//   on_serving       priority_network.py:217-240   node.serving() (queue pop + its service draw) BEFORE routing
//   choose_next_node priority_network.py:137-150   per-class routing matrix
//   node logic       sim/priority.py:282-332 (arrival in network mode), :390-455 (pre-emption: first server in
//                    index order with a weaker NODE class; victim to the tail of its class queue; PR keeps the
//                    remaining time, RS redraws from the arriving class, RW keeps the total), :457-539 (serving)
//   network moments  priority_network.py:152-173   3 running moments per real class with math.pow powers
// fp64 with explicit round-to-nearest intrinsics in the reference's order (replay tier bit-exact, the cube of the
// third network moment rounded once in double-double as in K5).  Streams in generator mode: stream sid =
// i*K + j for node i / node class j (word 1 = service draw), sid = m*K + k for class k's external gaps (word 0),
// sid = m*K + K for the routing uniforms (word 1).
#include "k7_prionet.cuh"

namespace mq {

struct K7Job {
    double arr_network, wait_network, arr_time, start_wait, wait_time, tte;
    int k, node_class, is_pr;
};

__device__ __forceinline__ void k7_refresh4(double *m, double x, long long count)
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

__device__ __forceinline__ double k7_cube(double x)
{
    const double p = __dmul_rn(x, x);
    const double e = __fma_rn(x, x, -p);
    const double q = __dmul_rn(p, x);
    const double e2 = __fma_rn(p, x, -q);
    return __dadd_rn(q, __fma_rn(e, x, e2));
}

__device__ __forceinline__ void k7_refresh_net(double *m, double x, long long served)
{
    const double sv = (double)served;
    const double factor = __dsub_rn(1.0, __ddiv_rn(1.0, sv));
    const double pw[3] = {x, __dmul_rn(x, x), k7_cube(x)};
#pragma unroll
    for (int i = 0; i < 3; ++i) m[i] = __dadd_rn(__dmul_rn(m[i], factor), __ddiv_rn(pw[i], sv));
}

__device__ __forceinline__ void k7_addk(double *s, double x, int n)
{
    double pw = x;
    for (int i = 0; i < n; ++i) {
        s[i] = __dadd_rn(s[i], pw);
        pw = __dmul_rn(pw, x);
    }
}

__global__ void __launch_bounds__(K7_BLOCK) k7_prionet(const K7Params P)
{
    const long long gtid = (long long)blockIdx.x * K7_BLOCK + threadIdx.x;
    const long long R = P.reps;
    const int m = P.m, K = P.K;
    const int RS = (m + 1) * (m + 1);

    for (long long rl = gtid; rl < R; rl += P.lanes) {
        const uint32_t rep = P.rep0 + (uint32_t)rl;
        double *rec = P.rec + rl;
        double *Q = P.queue + gtid;

        double end[K7_MAXS], total_time[K7_MAXS];
        K7Job on[K7_MAXS];
        unsigned int busy = 0u;
        int free_ch[K7_MAXN], qhead[K7_MAXN * K7_MAXK], qsize[K7_MAXN * K7_MAXK];
        long long n_taked[K7_MAXN * K7_MAXK], n_served[K7_MAXN * K7_MAXK], s_idx[K7_MAXN * K7_MAXK];
        double nv_run[K7_MAXN * K7_MAXK][4], nw_run[K7_MAXN * K7_MAXK][4], nv_sum[K7_MAXN * K7_MAXK][4],
            nw_sum[K7_MAXN * K7_MAXK][4];
        double arrival_time[K7_MAXK], v_run[K7_MAXK][3], w_run[K7_MAXK][3], v_sum[K7_MAXK][3], w_sum[K7_MAXK][3];
        long long served[K7_MAXK], arrived[K7_MAXK], a_idx[K7_MAXK];
        for (int i = 0; i < m; ++i) free_ch[i] = P.channels[i];
        for (int i = 0; i < m * K; ++i) {
            qhead[i] = qsize[i] = 0;
            n_taked[i] = n_served[i] = s_idx[i] = 0;
            for (int q = 0; q < 4; ++q) nv_run[i][q] = nw_run[i][q] = nv_sum[i][q] = nw_sum[i][q] = 0.0;
        }
        for (int k = 0; k < K; ++k) {
            served[k] = arrived[k] = a_idx[k] = 0;
            for (int q = 0; q < 3; ++q) v_run[k][q] = w_run[k][q] = v_sum[k][q] = w_sum[k][q] = 0.0;
        }
        for (int s = 0; s < P.nserv; ++s) {
            end[s] = 1e10;
            total_time[s] = 0.0;
        }
        long long u_idx = 0, served_total = 0;
        int status = 0;

        auto next_gap = [&](int k) -> double {
            const long long i = a_idx[k]++;
            if (P.replay) {
                if (i >= P.n_a[k]) {
                    status = MQ_E_INPUT_SHORT;
                    return 0.0;
                }
                return P.A[P.aoff[k] + i * R + rl];
            }
            const uint32_t key = P.key0 + MQ_KEY_STRIDE * (64u * (uint32_t)(m * K + k));
            uint2 w = philox2x32_10((uint32_t)i, rep, key);
            return (double)sample<-1>(P.src[k], w.x, (uint32_t)i, rep, key, 0);
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
            const uint32_t key = P.key0 + MQ_KEY_STRIDE * (64u * (uint32_t)(m * K + K));
            uint2 w = philox2x32_10((uint32_t)i, rep, key);
            return ((double)w.y + 0.5) * 0x1p-32;
        };
        auto next_service = [&](int nk) -> double {  // nk = node * K + node class
            const long long j = s_idx[nk]++;
            if (P.replay) {
                if (j >= P.n_s[nk]) {
                    status = MQ_E_INPUT_SHORT;
                    return 0.0;
                }
                return P.S[P.soff[nk] + j * R + rl];
            }
            const uint32_t key = P.key0 + MQ_KEY_STRIDE * (64u * (uint32_t)nk);
            uint2 w = philox2x32_10((uint32_t)j, rep, key);
            return (double)sample<-1>(P.srv[nk], w.y, (uint32_t)j, rep, key, 1);
        };
        auto qslot = [&](int nk, int slot, int f) -> double * {
            return Q + (((long long)nk * P.qcap + slot) * K7_JOBF + f) * P.lanes;
        };
        auto push = [&](int node, const K7Job &jb) {
            const int nk = node * K + jb.node_class;
            if (qsize[nk] >= P.qcap) {
                status = MQ_E_OVERFLOW;
                return;
            }
            int slot = qhead[nk] + qsize[nk];
            if (slot >= P.qcap) slot -= P.qcap;
            *qslot(nk, slot, 0) = jb.arr_network;
            *qslot(nk, slot, 1) = jb.wait_network;
            *qslot(nk, slot, 2) = jb.arr_time;
            *qslot(nk, slot, 3) = jb.start_wait;
            *qslot(nk, slot, 4) = jb.wait_time;
            *qslot(nk, slot, 5) = jb.tte;
            *qslot(nk, slot, 6) = (double)(jb.k | (jb.node_class << 8) | (jb.is_pr << 16));
            qsize[nk]++;
        };
        auto pop = [&](int nk) -> K7Job {
            K7Job jb;
            const int slot = qhead[nk];
            jb.arr_network = *qslot(nk, slot, 0);
            jb.wait_network = *qslot(nk, slot, 1);
            jb.arr_time = *qslot(nk, slot, 2);
            jb.start_wait = *qslot(nk, slot, 3);
            jb.wait_time = *qslot(nk, slot, 4);
            jb.tte = *qslot(nk, slot, 5);
            const int pk = (int)*qslot(nk, slot, 6);
            jb.k = pk & 255;
            jb.node_class = (pk >> 8) & 255;
            jb.is_pr = pk >> 16;
            qsize[nk]--;
            // an emptied line restarts at slot 0 (keeps the ring's footprint cache-resident)
            qhead[nk] = (qsize[nk] == 0 || slot + 1 >= P.qcap) ? 0 : slot + 1;
            return jb;
        };
        auto choose_next = [&](int real, int cur) -> int {
            const double p = next_uniform();
            double sum_p = 0.0;
            for (int i = 0; i < m + 1; ++i) {
                sum_p = __dadd_rn(sum_p, P.R[real * RS + (cur + 1) * (m + 1) + i]);
                if (sum_p > p) return i;
            }
            return 0;
        };
        double ttek = 0.0;
        // node arrival in network mode, sim/priority.py:282-332 with (k, moment, ts)
        auto node_arrival = [&](int node, int kc, K7Job job) {
            const int nk = node * K + kc, b = P.base[node], c = P.channels[node];
            job.node_class = kc;
            job.arr_time = ttek;
            job.wait_time = 0.0;
            job.is_pr = 0;
            job.start_wait = -1.0;
            job.tte = 0.0;
            if (free_ch[node] != 0) {
                const unsigned int mask = (busy >> b) & ((c >= 32) ? 0xffffffffu : ((1u << c) - 1u));
                const int j = b + __ffs(~mask) - 1;  // lowest free channel
                n_taked[nk]++;
                total_time[j] = next_service(nk);
                end[j] = __dadd_rn(ttek, total_time[j]);
                on[j] = job;
                busy |= 1u << j;
                free_ch[node]--;
            } else if (P.prty[node] == MQ_PRTY_NP) {
                job.start_wait = ttek;
                push(node, job);
            } else {
                bool found = false;
                for (int j = b; j < b + c; ++j) {
                    if (on[j].node_class > kc) {
                        const double time_to_end = end[j], tot = total_time[j];
                        K7Job victim = on[j];
                        n_taked[nk]++;
                        victim.start_wait = ttek;
                        victim.is_pr = 1;
                        if (P.prty[node] == MQ_PRTY_PR)
                            victim.tte = __dsub_rn(time_to_end, ttek);
                        else if (P.prty[node] == MQ_PRTY_RS)
                            victim.tte = next_service(nk);
                        else
                            victim.tte = tot;
                        push(node, victim);
                        total_time[j] = next_service(nk);
                        end[j] = __dadd_rn(ttek, total_time[j]);
                        on[j] = job;
                        found = true;
                        break;
                    }
                }
                if (!found) {
                    job.start_wait = ttek;
                    push(node, job);
                }
            }
        };

        for (int k = 0; k < K; ++k) arrival_time[k] = next_gap(k);  // priority_network.py:83

        while (served_total < P.jobs && status == 0) {
            int ev_k = 0;
            double et = arrival_time[0];
            for (int k = 1; k < K; ++k) {
                if (arrival_time[k] < et) {
                    et = arrival_time[k];
                    ev_k = k;
                }
            }
            int ev = -1;
            for (int s = 0; s < P.nserv; ++s) {
                if (((busy >> s) & 1u) && end[s] < et) {
                    et = end[s];
                    ev = s;
                }
            }
            if (ev < 0) {
                // on_arrival(k) :196-215
                const int k = ev_k;
                ttek = arrival_time[k];
                arrived[k]++;
                arrival_time[k] = __dadd_rn(ttek, next_gap(k));
                const int next = choose_next(k, -1);
                if (status) break;
                if (next >= m) {
                    status = MQ_E_INVALID;
                    break;
                }
                K7Job ts;
                ts.k = k;
                ts.arr_network = ttek;
                ts.wait_network = 0.0;
                ts.arr_time = ttek;
                ts.start_wait = -1.0;
                ts.wait_time = 0.0;
                ts.tte = 0.0;
                ts.node_class = 0;
                ts.is_pr = 0;
                node_arrival(next, P.nodes_prty[next * K + k], ts);
            } else {
                // on_serving :217-240 ; node.serving(c, True) sim/priority.py:457-539
                int node = 0;
                while (ev >= P.base[node + 1]) ++node;
                const int j = ev;
                ttek = end[j];
                const K7Job done = on[j];
                end[j] = 1e10;
                busy &= ~(1u << j);
                total_time[j] = 0.0;
                const int kc = done.node_class, nk = node * K + kc;
                n_served[nk]++;
                free_ch[node]++;
                const double vn = __dsub_rn(ttek, done.arr_time);
                k7_refresh4(nv_run[nk], vn, n_served[nk]);
                k7_refresh4(nw_run[nk], done.wait_time, n_served[nk]);
                k7_addk(nv_sum[nk], vn, 4);
                k7_addk(nw_sum[nk], done.wait_time, 4);
                const int start = P.prty[node] == MQ_PRTY_NP ? 0 : kc;
                for (int kk = start; kk < K; ++kk) {
                    const int nkk = node * K + kk;
                    if (qsize[nkk] != 0) {
                        K7Job qj = pop(nkk);
                        n_taked[nkk]++;
                        const double d = __dsub_rn(ttek, qj.start_wait);
                        qj.wait_time = __dadd_rn(qj.wait_time, d);
                        qj.wait_network = __dadd_rn(qj.wait_network, d);
                        if (qj.is_pr) {
                            end[j] = __dadd_rn(ttek, qj.tte);
                        } else {
                            total_time[j] = next_service(nkk);
                            end[j] = __dadd_rn(ttek, total_time[j]);
                        }
                        on[j] = qj;
                        busy |= 1u << j;
                        free_ch[node]--;
                        break;
                    }
                }
                const int real = done.k;
                const int next = choose_next(real, node);
                if (status) break;
                if (next == m) {
                    served[real]++;
                    served_total++;
                    const double v = __dsub_rn(ttek, done.arr_network);
                    k7_refresh_net(v_run[real], v, served[real]);
                    k7_refresh_net(w_run[real], done.wait_network, served[real]);
                    k7_addk(v_sum[real], v, 3);
                    k7_addk(w_sum[real], done.wait_network, 3);
                } else {
                    node_arrival(next, P.nodes_prty[next * K + real], done);
                }
            }
        }

        rec[(long long)MQ_PN_G_TTEK * R] = ttek;
        rec[(long long)MQ_PN_G_FLAGS * R] = (double)(status ? -status : 0);
        rec[(long long)MQ_PN_G_UUSED * R] = (double)u_idx;
        for (int k = 0; k < K; ++k) {
            double *rc = rec + (long long)(MQ_PN_G_ROWS + k * MQ_PNC_ROWS) * R;
            for (int q = 0; q < 3; ++q) {
                rc[(long long)(MQ_PNC_VSUM + q) * R] = v_sum[k][q];
                rc[(long long)(MQ_PNC_WSUM + q) * R] = w_sum[k][q];
                rc[(long long)(MQ_PNC_VRUN + q) * R] = v_run[k][q];
                rc[(long long)(MQ_PNC_WRUN + q) * R] = w_run[k][q];
            }
            rc[(long long)MQ_PNC_SERVED * R] = (double)served[k];
            rc[(long long)MQ_PNC_ARRIVED * R] = (double)arrived[k];
            rc[(long long)MQ_PNC_AUSED * R] = (double)a_idx[k];
        }
        for (int i = 0; i < m * K; ++i) {
            double *rn = rec + (long long)(MQ_PN_G_ROWS + K * MQ_PNC_ROWS + i * MQ_PNN_ROWS) * R;
            for (int q = 0; q < 4; ++q) {
                rn[(long long)(MQ_PNN_VSUM + q) * R] = nv_sum[i][q];
                rn[(long long)(MQ_PNN_WSUM + q) * R] = nw_sum[i][q];
                rn[(long long)(MQ_PNN_VRUN + q) * R] = nv_run[i][q];
                rn[(long long)(MQ_PNN_WRUN + q) * R] = nw_run[i][q];
            }
            rn[(long long)MQ_PNN_SERVED * R] = (double)n_served[i];
            rn[(long long)MQ_PNN_TAKED * R] = (double)n_taked[i];
            rn[(long long)MQ_PNN_SUSED * R] = (double)s_idx[i];
        }
    }
}

cudaError_t k7_launch(const K7Params &P, int grid, cudaStream_t st)
{
    k7_prionet<<<grid, K7_BLOCK, 0, st>>>(P);
    return cudaGetLastError();
}

cudaError_t k7_occupancy(int *blocks_per_sm)
{
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k7_prionet, K7_BLOCK, 0);
}

}  // namespace mq
