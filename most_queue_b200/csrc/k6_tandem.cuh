// k6_tandem.cuh -- parameters of the tandem-with-blocking kernel (k6_tandem.cu).
#pragma once

#include "mq_device.cuh"

namespace mq {

constexpr int K6_BLOCK = 128;
constexpr int K6_MAXM = MQ_NET_MAX_NODES;

struct K6Params {
    int m, replay;
    long long capacity[K6_MAXM];
    long long jobs, warm, reps;
    uint32_t rep0, key0;
    DistF src, srv[K6_MAXM];
    const double *A, *S;
    long long n_a, soff[K6_MAXM], n_s[K6_MAXM];
    double *rec;  // [MQ_TANDEM_REC_ROWS(m)][reps]
    long long lanes;
};

cudaError_t k6_launch(const K6Params &P, int grid, cudaStream_t st);
cudaError_t k6_occupancy(int *blocks_per_sm);

// k6v2_tandem.cu: the predicated, register / shared-memory resident form for lines of at most 16 nodes
bool k6v2_supported(int m);
cudaError_t k6v2_launch(const K6Params &P, int grid, cudaStream_t st);
cudaError_t k6v2_occupancy(int m, bool replay, int *blocks_per_sm);

}  // namespace mq
