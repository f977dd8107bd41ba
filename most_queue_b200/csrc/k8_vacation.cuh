// k8_vacation.cuh -- parameters of the vacation / N-policy queue kernel (k8_vacation.cu).
#pragma once

#include "mq_device.cuh"

namespace mq {

constexpr int K8_BLOCK = 128;
constexpr int K8_MAXC = 32;

struct K8Params {
    int c, p_bins, replay, qcap;
    int set[MQ_VAC_STREAMS];
    int service_on_warm_up, multiple_vacations;
    long long buffer, big_n;
    long long jobs, reps;
    uint32_t rep0, key0;
    DistF dist[MQ_VAC_STREAMS];
    const double *X;                                   // replay: all streams, stream s at X + off[s], [n[s]][reps]
    long long off[MQ_VAC_STREAMS], n[MQ_VAC_STREAMS];
    double *rec;    // [MQ_V_ROWS(p_bins)][reps], zeroed by the caller
    double *queue;  // [qcap][lanes] arrival times of the waiting jobs
    long long lanes;
};

cudaError_t k8_launch(const K8Params &P, int grid, cudaStream_t st);
cudaError_t k8_occupancy(int c, int *blocks_per_sm);

}  // namespace mq
