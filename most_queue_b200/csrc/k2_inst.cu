// k2_inst.cu -- instantiates K2 for one slot count (compile with -DMQ_CT=<n>) and the three
// variate pair classes: 0 = runtime switch on both kinds, 1 = M/M, 2 = M/H2.
#include "k2_fifo_gen.cuh"

#ifndef MQ_CT
#error "compile with -DMQ_CT=<slot count>"
#endif

namespace mq {

#define MQ_CAT2(a, b) a##b
#define MQ_CAT(a, b) MQ_CAT2(a, b)

template <int SRCK, int SRVK>
static cudaError_t go(const K2Params &P, bool finite, int grid, size_t smem, cudaStream_t st, int *occ)
{
    auto run = [&](auto kern) -> cudaError_t {
        if (occ) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, MQ_BLOCK, smem);
        kern<<<grid, MQ_BLOCK, smem, st>>>(P);
        return cudaGetLastError();
    };
    return finite ? run(k2_fifo_gen<MQ_CT, SRCK, SRVK, true>) : run(k2_fifo_gen<MQ_CT, SRCK, SRVK, false>);
}

// occ != nullptr: only query resident CTAs per SM for this variant
cudaError_t MQ_CAT(k2_dispatch_ct, MQ_CT)(const K2Params &P, int pair_class, bool finite, int grid,
                                          size_t smem, cudaStream_t st, int *occ)
{
    switch (pair_class) {
    case 1:
        return go<MQ_M, MQ_M>(P, finite, grid, smem, st, occ);
    case 2:
        return go<MQ_M, MQ_H2>(P, finite, grid, smem, st, occ);
    default:
        return go<-1, -1>(P, finite, grid, smem, st, occ);
    }
}

}  // namespace mq
