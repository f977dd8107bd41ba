// reduce.cu -- deterministic reduction of the per-replication records to the aggregate
// vector (MQ_A_* layout of include/mqb200.h).  One CTA per aggregate entry, fixed-order tree,
// so the result does not depend on scheduling.  The aggregate is what a multi-GPU run sums
// with one allreduce.
#include "mq_host.cuh"

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

// generic form used by the priority and network kernels: entry q is the sum over replications of
// x = rec[num][r] / rec[den][r] (den < 0: x = rec[num][r]) and, in out[nq + q], of x^2.
__global__ void __launch_bounds__(MQ_RED_BLOCK) reduce_ratio(const double *rec, long long R, const RatioDesc *desc,
                                                             int nq, double *out)
{
    __shared__ double part[MQ_RED_BLOCK], part2[MQ_RED_BLOCK];
    const int qi = blockIdx.x;
    const RatioDesc d = desc[qi];
    double acc = 0.0, acc2 = 0.0;
    for (long long r = threadIdx.x; r < R; r += MQ_RED_BLOCK) {
        double x = rec[(long long)d.num * R + r];
        if (d.den >= 0) {
            const double dn = rec[(long long)d.den * R + r];
            x = dn != 0.0 ? x / dn : 0.0;
        }
        acc += x;
        acc2 += x * x;
    }
    part[threadIdx.x] = acc;
    part2[threadIdx.x] = acc2;
    __syncthreads();
    for (int s = MQ_RED_BLOCK / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            part[threadIdx.x] += part[threadIdx.x + s];
            part2[threadIdx.x] += part2[threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[qi] = part[0];
        out[nq + qi] = part2[0];
    }
}

cudaError_t launch_reduce_records(const double *rec, long long R, int p_bins, double *agg, cudaStream_t st)
{
    reduce_records<<<MQ_AGG_LEN(p_bins), MQ_RED_BLOCK, 0, st>>>(rec, R, p_bins, agg);
    return cudaGetLastError();
}

cudaError_t launch_reduce_ratio(const double *rec, long long R, const RatioDesc *desc, int nq, double *out,
                                cudaStream_t st)
{
    reduce_ratio<<<nq, MQ_RED_BLOCK, 0, st>>>(rec, R, desc, nq, out);
    return cudaGetLastError();
}

}  // namespace mq
