// reduce.cuh -- deterministic reduction of the per-replication records to the aggregate
// vector (MQ_A_* layout of include/mqb200.h).  One CTA per aggregate entry, fixed-order tree,
// so the result does not depend on scheduling.  The aggregate is what a multi-GPU run sums
// with one allreduce.
#pragma once

#include "mq_device.cuh"

namespace mq {

constexpr int MQ_RED_BLOCK = 256;

__device__ __forceinline__ double agg_term(const double *rec, long long R, long long r, int p_bins, int qi)
{
    auto row = [&](int f) { return rec[(long long)f * R + r]; };
    if (qi == MQ_A_REPS) return 1.0;
    if (qi >= MQ_A_W && qi < MQ_A_W + 8) {
        const int k = (qi - MQ_A_W) & 3;
        const double cnt = row(MQ_F_TAKED);
        const double m = cnt > 0.0 ? row(MQ_F_W + k) / cnt : 0.0;
        return qi < MQ_A_W2 ? m : m * m;
    }
    if (qi >= MQ_A_V && qi < MQ_A_V + 8) {
        const int k = (qi - MQ_A_V) & 3;
        const double cnt = row(MQ_F_SERVED);
        const double m = cnt > 0.0 ? row(MQ_F_V + k) / cnt : 0.0;
        return qi < MQ_A_V2 ? m : m * m;
    }
    if (qi == MQ_A_ARRIVED) return row(MQ_F_ARRIVED);
    if (qi == MQ_A_TAKED) return row(MQ_F_TAKED);
    if (qi == MQ_A_SERVED) return row(MQ_F_SERVED);
    if (qi == MQ_A_DROPPED) return row(MQ_F_DROPPED);
    if (qi == MQ_A_TTEK) return row(MQ_F_TTEK);
    if (qi >= MQ_A_WPOOL && qi < MQ_A_WPOOL + 4) return row(MQ_F_W + qi - MQ_A_WPOOL);
    if (qi >= MQ_A_VPOOL && qi < MQ_A_VPOOL + 4) return row(MQ_F_V + qi - MQ_A_VPOOL);
    if (qi == MQ_A_FLAGGED) return row(MQ_F_FLAGS) != 0.0 ? 1.0 : 0.0;
    const double T = row(MQ_F_TTEK);
    const double invT = T > 0.0 ? 1.0 / T : 0.0;
    if (qi == MQ_A_POVER) return row(MQ_F_POVER) * invT;
    if (qi >= MQ_A_P && qi < MQ_A_P + p_bins) return row(MQ_F_P + qi - MQ_A_P) * invT;
    if (qi >= MQ_A_P + p_bins && qi < MQ_A_P + 2 * p_bins) {
        const double p = row(MQ_F_P + qi - MQ_A_P - p_bins) * invT;
        return p * p;
    }
    return 0.0;
}

__global__ void __launch_bounds__(MQ_RED_BLOCK) reduce_records(const double *rec, long long R, int p_bins,
                                                               double *agg)
{
    __shared__ double part[MQ_RED_BLOCK];
    const int qi = blockIdx.x;
    double acc = 0.0;
    for (long long r = threadIdx.x; r < R; r += MQ_RED_BLOCK) acc += agg_term(rec, R, r, p_bins, qi);
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int s = MQ_RED_BLOCK / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) agg[qi] = part[0];
}

}  // namespace mq
