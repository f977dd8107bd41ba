// k9_negatives.cuh -- parameters of the negative-arrivals queue kernel (k9_negatives.cu).
#pragma once

#include "mq_device.cuh"

namespace mq {

constexpr int K9_BLOCK = 128;
constexpr int K9_MAXC = 32;

struct K9Params {
    int c, p_bins, replay, qcap, type, scenario;
    long long buffer;
    long long jobs, reps;
    uint32_t rep0, key0;
    DistF positive, negative, service;
    const double *X;  // replay: stream s at X + off[s], [n[s]][reps]
    long long off[MQ_NEG_STREAMS], n[MQ_NEG_STREAMS];
    double *rec;    // [MQ_N_ROWS(p_bins)][reps], zeroed by the caller
    double *queue;  // [qcap][lanes] arrival times of the waiting jobs
    long long lanes;
};

cudaError_t k9_launch(const K9Params &P, int grid, cudaStream_t st);
cudaError_t k9_occupancy(int c, int *blocks_per_sm);

}  // namespace mq
