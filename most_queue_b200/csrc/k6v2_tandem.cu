// k6v2_tandem.cu -- K6 v2: the tandem-with-blocking kernel as ONE predicated trip body, for lines of at most 16 nodes.
//
// Same event sequence, same arithmetic and same record rows as k6_tandem.cu (see its header for the reference map,
// most_queue/sim/networks/tandem_blocking.py:50-142).  Shape of the code:
//   * a trip is an arrival, a completion (blocked, or pushed downstream) or ONE step of the unblocking cascade towards
//     the head of the line (:127-129) -- the cascade is a pending state, not an inner loop, so every lane of a warp
//     runs the same straight-line body whatever its event is;
//   * at most two services start in a trip -- the node that pushed restarts (:84-85) and the next node, or node 0 on
//     an arrival, accepts (:86-89) -- through one "restart" site and one "accept" site; the completion times are
//     registers, rewritten once at the end of the trip;
//   * jobs per node and the per-node stream cursors are shared memory [word][lane]; node laws and capacities are
//     shared memory; in the generator tier the service draws are produced up to K6V_SB ahead per node by all lanes of
//     the warp together (a draw is a pure function of its index).
#include <climits>

#include "k6_tandem.cuh"

namespace mq {

constexpr int K6V_BLOCK = 128;
constexpr int K6V_SB = 4;  // service draws kept ahead per node and lane (generator tier; power of two)

template <int MT, bool REPLAY>
struct K6VLayout {
    // 32-bit words per lane: JOBS[MT], SIDX[MT], then (generator) SFILL[MT] and K6V_SB * MT fp32 draws
    static constexpr int JOBS = 0, SIDX = MT, SFILL = 2 * MT, DRAWS = 3 * MT;
    static constexpr int NW = REPLAY ? 2 * MT : 3 * MT + K6V_SB * MT;
    static constexpr size_t BYTES = (size_t)K6V_BLOCK * 4 * NW;
};

template <int MT, bool REPLAY>
__global__ void __launch_bounds__(K6V_BLOCK) k6v2_tandem(const K6Params P)
{
    using LY = K6VLayout<MT, REPLAY>;
    extern __shared__ int k6v_sm[];
    __shared__ DistF sh_srv[MT];
    __shared__ int sh_cap[MT + 1];
    const int tid = threadIdx.x;
    const int m = P.m;
    if (tid < MT) {
        if (!REPLAY && tid < m) sh_srv[tid] = P.srv[tid];
        // "room at node i" is jobs[i] < cap[i]; no capacity = never full
        const long long cap = tid < m ? P.capacity[tid] : -1;
        sh_cap[tid] = cap < 0 ? INT_MAX : (cap > INT_MAX - 1 ? INT_MAX - 1 : (int)cap);
    }
    if (tid == 0) sh_cap[MT] = INT_MAX;
    __syncthreads();

    int *smi = k6v_sm + tid;
    float *smf = reinterpret_cast<float *>(k6v_sm) + tid;
#define SI(w) smi[(w) * K6V_BLOCK]
#define SF(w) smf[(w) * K6V_BLOCK]
    const long long gtid = (long long)blockIdx.x * K6V_BLOCK + tid;
    const long long R = P.reps;
    const double INF = __longlong_as_double(0x7ff0000000000000ll);
    const DistF src_law = P.src;
    const uint32_t key_m = P.key0 + MQ_KEY_STRIDE * (64u * (uint32_t)m);
    const long long target = P.jobs + P.warm;

    for (long long rl = gtid; rl < R; rl += P.lanes) {
        const uint32_t rep = P.rep0 + (uint32_t)rl;
        double *rec = P.rec + rl;
        for (int w = 0; w < LY::NW; ++w) SI(w) = 0;
        double completion[MT], area[MT];
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            completion[i] = INF;
            area[i] = 0.0;
        }
        uint32_t blocked = 0, need = 1u;
        long long a_idx = 0;
        int status = 0, casc = 0;
        long long served = 0;
        int arrived = 0, lost = 0, departures = 0;
        bool stats_on = false;
        double t = 0.0, t0 = 0.0;

        auto next_gap = [&]() -> double {
            const long long i = a_idx++;
            if (REPLAY) {
                if (i >= P.n_a) {
                    status = MQ_E_INPUT_SHORT;
                    return 0.0;
                }
                return P.A[i * R + rl];
            }
            uint2 w = philox2x32_10((uint32_t)i, rep, key_m);
            return (double)sample<-1>(src_law, w.x, (uint32_t)i, rep, key_m, 0);
        };
        auto refill_services = [&]() {
            if (!__any_sync(__activemask(), need != 0u)) return;
            need = 0u;
            for (int node = 0; node < m; ++node) {
                const uint32_t key = P.key0 + MQ_KEY_STRIDE * (64u * (uint32_t)node);
                const DistF d = sh_srv[node];
                for (;;) {
                    const int fill = SI(LY::SFILL + node);
                    const bool more = fill < K6V_SB;
                    if (!__any_sync(__activemask(), more)) break;
                    if (more) {
                        const uint32_t j = (uint32_t)(SI(LY::SIDX + node) + fill);
                        uint2 w = philox2x32_10(j, rep, key);
                        SF(LY::DRAWS + node * K6V_SB + (int)(j & (K6V_SB - 1))) = sample<-1>(d, w.y, j, rep, key, 1);
                        SI(LY::SFILL + node) = fill + 1;
                    }
                }
            }
        };
        auto next_service = [&](int node) -> double {
            const int j = SI(LY::SIDX + node);
            SI(LY::SIDX + node) = j + 1;
            if (REPLAY) {
                if (j >= P.n_s[node]) {
                    status = MQ_E_INPUT_SHORT;
                    return 0.0;
                }
                return P.S[P.soff[node] + (long long)j * R + rl];
            }
            const int fill = SI(LY::SFILL + node) - 1;
            SI(LY::SFILL + node) = fill;
            need |= fill < 2 ? 1u : 0u;
            return (double)SF(LY::DRAWS + node * K6V_SB + (j & (K6V_SB - 1)));
        };

        double next_arrival = next_gap();

        while ((served < target || casc > 0) && status == 0) {
            if (!REPLAY) refill_services();
            const bool casc_trip = casc > 0;
            // ---- next event: min(next_arrival, completion[*]); arrival first on ties (:103), first minimal node (:113)
            double cmin = completion[0];
            int imin = 0;
#pragma unroll
            for (int i = 1; i < MT; ++i) {
                const bool lt = completion[i] < cmin;
                cmin = lt ? completion[i] : cmin;
                imin = lt ? i : imin;
            }
            const bool is_arr = !casc_trip && next_arrival <= cmin;
            const bool comp = !casc_trip && !is_arr;
            const double t_next = casc_trip ? t : (next_arrival < cmin ? next_arrival : cmin);
            // ---- area_l[i] += jobs[i] * dt (:105-107).  Before the warm-up ends, and on a cascade step, the term is +0.0,
            // which leaves the (non-negative) sums bit for bit as they are
            {
                const double dt = stats_on ? __dsub_rn(t_next, t) : 0.0;
#pragma unroll
                for (int i = 0; i < MT; ++i) area[i] = __dadd_rn(area[i], __dmul_rn((double)SI(LY::JOBS + i), dt));
            }
            t = t_next;

            // ---- arrival (:108-112)
            const bool arr_ok = is_arr && SI(LY::JOBS + 0) < sh_cap[0];
            arrived += (is_arr && stats_on) ? 1 : 0;
            lost += (is_arr && !arr_ok && stats_on) ? 1 : 0;
            if (is_arr) next_arrival = __dadd_rn(t, next_gap());

            // ---- completion at node imin: blocked (:114-116), or the job moves on
            const int i = imin;
            const bool blk = comp && i + 1 < m && SI(LY::JOBS + (i + 1 < MT ? i + 1 : 0)) >= sh_cap[i + 1];
            const bool push = casc_trip || (comp && !blk);
            const int x = casc_trip ? casc - 1 : i;  // the node whose job moves downstream
            if (comp && !blk && i + 1 == m) {
                // a job leaves the line (:118-125)
                served++;
                if (stats_on) {
                    departures++;
                } else if (served >= P.warm) {
                    stats_on = true;
                    t0 = t;
                    arrived = lost = departures = 0;
#pragma unroll
                    for (int j = 0; j < MT; ++j) area[j] = 0.0;
                }
            }
            blocked = blk ? (blocked | (1u << i)) : blocked;

            // ---- try_push(x) (:81-89): node x restarts its server, then the next node accepts; an arrival is accept(0)
            int w1 = blk ? i : -1, w2 = -1;
            double v1 = INF, v2 = INF;
            int jx = 0;
            if (push) {
                jx = SI(LY::JOBS + x) - 1;
                SI(LY::JOBS + x) = jx;
                blocked &= ~(1u << x);
                w1 = x;
                if (jx > 0) v1 = __dadd_rn(t, next_service(x));
            }
            const int y = is_arr ? 0 : x + 1;
            if (arr_ok || (push && y < m)) {
                const int jy = SI(LY::JOBS + y) + 1;
                SI(LY::JOBS + y) = jy;
                if (jy == 1) {
                    w2 = y;
                    v2 = __dadd_rn(t, next_service(y));
                }
            }
#pragma unroll
            for (int k = 0; k < MT; ++k) completion[k] = (k == w2) ? v2 : ((k == w1) ? v1 : completion[k]);

            // ---- the cascade towards the head of the line (:127-129): node x - 1 is blocked and node x has room again
            casc = (push && x > 0 && ((blocked >> (x - 1)) & 1u) && jx < sh_cap[x]) ? x : 0;
        }

        double area_sum = 0.0;
        for (int i = 0; i < m; ++i) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < MT; ++k) a = (k == i) ? area[k] : a;
            rec[(long long)(MQ_TG_ROWS + i * MQ_TN_ROWS + MQ_TN_AREA) * R] = a;
            rec[(long long)(MQ_TG_ROWS + i * MQ_TN_ROWS + MQ_TN_SUSED) * R] = (double)SI(LY::SIDX + i);
            area_sum += a;
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
#undef SI
#undef SF
}

template <int MT>
static cudaError_t launch_one(const K6Params &P, int grid, cudaStream_t st)
{
    if (P.replay) {
        constexpr size_t smem = K6VLayout<MT, true>::BYTES;
        cudaError_t e = cudaFuncSetAttribute(k6v2_tandem<MT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k6v2_tandem<MT, true><<<grid, K6V_BLOCK, smem, st>>>(P);
    } else {
        constexpr size_t smem = K6VLayout<MT, false>::BYTES;
        cudaError_t e = cudaFuncSetAttribute(k6v2_tandem<MT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k6v2_tandem<MT, false><<<grid, K6V_BLOCK, smem, st>>>(P);
    }
    return cudaGetLastError();
}

template <int MT>
static cudaError_t occupancy_one(bool replay, int *occ)
{
    if (replay) {
        constexpr size_t smem = K6VLayout<MT, true>::BYTES;
        cudaError_t e = cudaFuncSetAttribute(k6v2_tandem<MT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k6v2_tandem<MT, true>, K6V_BLOCK, smem);
    }
    constexpr size_t smem = K6VLayout<MT, false>::BYTES;
    cudaError_t e = cudaFuncSetAttribute(k6v2_tandem<MT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k6v2_tandem<MT, false>, K6V_BLOCK, smem);
}

bool k6v2_supported(int m) { return m >= 1 && m <= 16; }

int k6v2_block() { return K6V_BLOCK; }

cudaError_t k6v2_launch(const K6Params &P, int grid, cudaStream_t st)
{
    if (P.m <= 4) return launch_one<4>(P, grid, st);
    if (P.m <= 8) return launch_one<8>(P, grid, st);
    return launch_one<16>(P, grid, st);
}

cudaError_t k6v2_occupancy(int m, bool replay, int *blocks_per_sm)
{
    if (m <= 4) return occupancy_one<4>(replay, blocks_per_sm);
    if (m <= 8) return occupancy_one<8>(replay, blocks_per_sm);
    return occupancy_one<16>(replay, blocks_per_sm);
}

}  // namespace mq
