// k2_inst.cu -- instantiates K2 for one slot count (compile with -DMQ_CT=<n>) and the three
// variate pair classes: 0 = runtime switch on both kinds, 1 = M/M, 2 = M/H2, 3 = fp64 tier (runtime switch).
#include "k2_fifo_gen.cuh"

#ifndef MQ_CT
#error "compile with -DMQ_CT=<slot count>"
#endif

namespace mq {

#define MQ_CAT2(a, b) a##b
#define MQ_CAT(a, b) MQ_CAT2(a, b)

template <typename Real, int SRCK, int SRVK, bool BUSY>
static cudaError_t go(const K2Params &P, bool finite, int grid, size_t smem, cudaStream_t st, int *occ)
{
    auto run = [&](auto kern) -> cudaError_t {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        if (occ) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, MQ_BLOCK, smem);
        kern<<<grid, MQ_BLOCK, smem, st>>>(P);
        return cudaGetLastError();
    };
    return finite ? run(k2_fifo_gen<Real, MQ_CT, SRCK, SRVK, true, BUSY>)
                  : run(k2_fifo_gen<Real, MQ_CT, SRCK, SRVK, false, BUSY>);
}

// occ != nullptr: only query resident CTAs per SM for this variant.  pair_class 3 = the fp64 A/B tier (runtime switch
// on both kinds); `busy` selects the instantiations that also gather the busy-period moments (generic classes only)
cudaError_t MQ_CAT(k2_dispatch_ct, MQ_CT)(const K2Params &P, int pair_class, bool finite, bool busy, int grid,
                                          size_t smem, cudaStream_t st, int *occ)
{
    switch (pair_class) {
    case 1:
        return go<float, MQ_M, MQ_M, false>(P, finite, grid, smem, st, occ);
    case 2:
        return go<float, MQ_M, MQ_H2, false>(P, finite, grid, smem, st, occ);
    case 3:
        return busy ? go<double, -1, -1, true>(P, finite, grid, smem, st, occ)
                    : go<double, -1, -1, false>(P, finite, grid, smem, st, occ);
    default:
        return busy ? go<float, -1, -1, true>(P, finite, grid, smem, st, occ)
                    : go<float, -1, -1, false>(P, finite, grid, smem, st, occ);
    }
}

}  // namespace mq
