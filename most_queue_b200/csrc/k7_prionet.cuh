// k7_prionet.cuh -- parameters of the priority-network kernel (k7_prionet.cu).
#pragma once

#include "mq_device.cuh"

namespace mq {

constexpr int K7_BLOCK = 64;
constexpr int K7_MAXN = MQ_PN_MAX_NODES;
constexpr int K7_MAXK = MQ_PN_MAX_CLASSES;
constexpr int K7_MAXS = MQ_PN_MAX_SERVERS;
constexpr int K7_JOBF = 7;  // doubles per queued job record

struct K7Params {
    int m, K, replay, qcap, nserv;
    int channels[K7_MAXN], base[K7_MAXN + 1], prty[K7_MAXN], nodes_prty[K7_MAXN * K7_MAXK];
    double R[K7_MAXK * (K7_MAXN + 1) * (K7_MAXN + 1)];
    long long jobs, reps;
    uint32_t rep0, key0;
    DistF src[K7_MAXK], srv[K7_MAXN * K7_MAXK];
    const double *A, *U, *S;
    long long aoff[K7_MAXK], n_a[K7_MAXK], n_u, soff[K7_MAXN * K7_MAXK], n_s[K7_MAXN * K7_MAXK];
    double *rec;
    double *queue;  // [m][K][qcap][K7_JOBF][lanes]
    long long lanes;
};

cudaError_t k7_launch(const K7Params &P, int grid, cudaStream_t st);
cudaError_t k7_occupancy(int *blocks_per_sm);

// k7v2_prionet.cu: the register / shared-memory resident predicated form for <= 16 servers in total
bool k7v2_supported(int m, int K, int nserv);
cudaError_t k7v2_launch(const K7Params &P, int grid, cudaStream_t st);
cudaError_t k7v2_occupancy(const K7Params &P, int *blocks_per_sm);

}  // namespace mq
