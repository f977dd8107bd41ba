// k6_tandem.cu -- K6: tandem line, single-server nodes, finite buffers, blocking after service;
// one lane per replication.
//
// Walks TandemBlockingSim.run (most_queue/sim/networks/tandem_blocking.py:50-142) event for event:
//   t_next = min(next_arrival, completion[*]); arrival first on ties (:103); first minimal node (:113)
//   blocked completion -> the job stays on the server until the next node has room (:114-116)
//   try_push(i): node i restarts its server BEFORE node i+1 accepts the job (:81-89, draw order)
//   cascade of unblocking towards the head of the line (:127-129)
//   statistics start at the departure that completes the warm-up (:121-125)
// fp64 with explicit round-to-nearest intrinsics in the reference's order: in replay mode throughput,
// loss probability and mean jobs per node are bit-identical to the reference.
#include "k6_tandem.cuh"

namespace mq {

__global__ void __launch_bounds__(K6_BLOCK) k6_tandem(const K6Params P)
{
    const long long gtid = (long long)blockIdx.x * K6_BLOCK + threadIdx.x;
    const long long R = P.reps;
    const int m = P.m;
    const double INF = __longlong_as_double(0x7ff0000000000000ll);

    for (long long rl = gtid; rl < R; rl += P.lanes) {
        const uint32_t rep = P.rep0 + (uint32_t)rl;
        double *rec = P.rec + rl;
        double completion[K6_MAXM], area[K6_MAXM];
        long long jobs[K6_MAXM], s_idx[K6_MAXM];
        uint32_t blocked = 0;
        for (int i = 0; i < m; ++i) {
            completion[i] = INF;
            area[i] = 0.0;
            jobs[i] = 0;
            s_idx[i] = 0;
        }
        long long a_idx = 0;
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
        double t = 0.0;
        auto start_service = [&](int i) {
            const long long j = s_idx[i]++;
            double sv;
            if (P.replay) {
                if (j >= P.n_s[i]) {
                    status = MQ_E_INPUT_SHORT;
                    sv = 0.0;
                } else {
                    sv = P.S[P.soff[i] + j * R + rl];
                }
            } else {
                const uint32_t key = P.key0 + MQ_KEY_STRIDE * (64u * (uint32_t)i);
                uint2 w = philox2x32_10((uint32_t)j, rep, key);
                sv = (double)sample<-1>(P.srv[i], w.y, (uint32_t)j, rep, key, 1);
            }
            completion[i] = __dadd_rn(t, sv);
        };
        auto accept = [&](int i) {
            jobs[i] += 1;
            if (jobs[i] == 1) start_service(i);
        };
        auto try_push = [&](int i) {
            jobs[i] -= 1;
            blocked &= ~(1u << i);
            completion[i] = INF;
            if (jobs[i] > 0) start_service(i);
            if (i + 1 < m) accept(i + 1);
        };

        double next_arrival = next_gap();
        long long served = 0, arrived = 0, lost = 0, departures = 0;
        bool stats_on = false;
        double t0 = 0.0;
        const long long target = P.jobs + P.warm;

        while (served < target && status == 0) {
            double cmin = completion[0];
            int imin = 0;
            for (int i = 1; i < m; ++i) {
                if (completion[i] < cmin) {
                    cmin = completion[i];
                    imin = i;
                }
            }
            const double t_next = next_arrival < cmin ? next_arrival : cmin;
            if (stats_on) {
                const double dt = __dsub_rn(t_next, t);
                for (int i = 0; i < m; ++i) area[i] = __dadd_rn(area[i], __dmul_rn((double)jobs[i], dt));
            }
            t = t_next;
            if (next_arrival <= cmin) {
                if (stats_on) arrived++;
                if (P.capacity[0] >= 0 && jobs[0] >= P.capacity[0]) {
                    if (stats_on) lost++;
                } else {
                    accept(0);
                }
                next_arrival = __dadd_rn(t, next_gap());
                continue;
            }
            const int i = imin;
            if (i + 1 < m && P.capacity[i + 1] >= 0 && jobs[i + 1] >= P.capacity[i + 1]) {
                blocked |= 1u << i;
                completion[i] = INF;
            } else {
                if (i + 1 == m) {
                    served++;
                    if (stats_on) departures++;
                    if (!stats_on && served >= P.warm) {
                        stats_on = true;
                        t0 = t;
                        arrived = lost = departures = 0;
                        for (int j = 0; j < m; ++j) area[j] = 0.0;
                    }
                }
                try_push(i);
                int j = i;
                while (j > 0 && ((blocked >> (j - 1)) & 1u) && (P.capacity[j] < 0 || jobs[j] < P.capacity[j])) {
                    try_push(j - 1);
                    j--;
                }
            }
        }

        double area_sum = 0.0;
        for (int i = 0; i < m; ++i) {
            // mean_jobs[i] = area[i] / elapsed and v[0] = sum(mean_jobs) / throughput are formed on the
            // host from these rows with the reference's operations
            rec[(long long)(MQ_TG_ROWS + i * MQ_TN_ROWS + MQ_TN_AREA) * R] = area[i];
            rec[(long long)(MQ_TG_ROWS + i * MQ_TN_ROWS + MQ_TN_SUSED) * R] = (double)s_idx[i];
            area_sum += area[i];
        }
        rec[(long long)MQ_TG_ELAPSED * R] = __dsub_rn(t, t0);
        rec[(long long)MQ_TG_FLAGS * R] = (double)(status ? -status : 0);
        rec[(long long)MQ_TG_ARRIVED * R] = (double)arrived;
        rec[(long long)MQ_TG_LOST * R] = (double)lost;
        rec[(long long)MQ_TG_DEPARTURES * R] = (double)departures;
        rec[(long long)MQ_TG_SERVED * R] = (double)served;
        rec[(long long)MQ_TG_AREASUM * R] = area_sum;
        rec[(long long)MQ_TG_TEND * R] = t;
        rec[(long long)MQ_TG_T0 * R] = t0;
        rec[(long long)MQ_TG_AUSED * R] = (double)a_idx;
    }
}

cudaError_t k6_launch(const K6Params &P, int grid, cudaStream_t st)
{
    k6_tandem<<<grid, K6_BLOCK, 0, st>>>(P);
    return cudaGetLastError();
}

cudaError_t k6_occupancy(int *blocks_per_sm)
{
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k6_tandem, K6_BLOCK, 0);
}

}  // namespace mq
