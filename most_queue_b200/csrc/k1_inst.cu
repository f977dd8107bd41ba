// k1_inst.cu -- instantiates K1 (replay) for one slot count (compile with -DMQ_CT=<n>).
#include "k1_fifo_replay.cuh"

#ifndef MQ_CT
#error "compile with -DMQ_CT=<slot count>"
#endif

namespace mq {

#define MQ_CAT2(a, b) a##b
#define MQ_CAT(a, b) MQ_CAT2(a, b)

cudaError_t MQ_CAT(k1_dispatch_ct, MQ_CT)(const K1Params &P, bool finite, cudaStream_t st)
{
    const int grid = (int)((P.reps + MQ_K1_BLOCK - 1) / MQ_K1_BLOCK);
    constexpr size_t smem = k1_smem_bytes(MQ_CT);
    static bool attr_set = false;
    if (!attr_set) {  // more than 48 KB of dynamic shared memory is opt-in
        cudaError_t e = cudaFuncSetAttribute(k1_fifo_replay<MQ_CT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(k1_fifo_replay<MQ_CT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    if (finite)
        k1_fifo_replay<MQ_CT, true><<<grid, MQ_K1_BLOCK, smem, st>>>(P);
    else
        k1_fifo_replay<MQ_CT, false><<<grid, MQ_K1_BLOCK, smem, st>>>(P);
    return cudaGetLastError();
}

}  // namespace mq
