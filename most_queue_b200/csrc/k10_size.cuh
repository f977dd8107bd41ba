// k10_size.cuh -- parameters of the size-based single-server kernel (k10_size.cu).
#pragma once

#include "mq_device.cuh"

namespace mq {

constexpr int K10_BLOCK = 128;
constexpr int K10_JOBF = 7;  // doubles per waiting job: rank, arrival, start of wait, wait, remaining, size, id

struct K10Params {
    int discipline, p_bins, replay, qcap;
    long long buffer;
    long long jobs, reps;
    uint32_t rep0, key0;
    DistF src, srv;
    const double *A, *S;  // replay: [n_a][reps], [n_s][reps]
    long long n_a, n_s;
    double *rec;    // [MQ_Z_ROWS(p_bins)][reps], zeroed by the caller
    double *queue;  // [qcap][K10_JOBF][lanes]
    long long lanes;
};

cudaError_t k10_launch(const K10Params &P, int grid, cudaStream_t st);
cudaError_t k10_occupancy(int *blocks_per_sm);

}  // namespace mq
