// k5_net.cuh -- parameters of the open-network kernel (k5_net.cu).
#pragma once

#include "mq_device.cuh"

namespace mq {

constexpr int K5_BLOCK = 64;
constexpr int K5_MAXN = MQ_NET_MAX_NODES;
constexpr int K5_MAXS = MQ_NET_MAX_SERVERS;

struct K5Params {
    int m, replay, qcap, nserv;
    int channels[K5_MAXN], base[K5_MAXN + 1];
    double R[(K5_MAXN + 1) * (K5_MAXN + 1)];  // stride m + 1
    long long jobs, reps;
    uint32_t rep0, key0;
    DistF src, srv[K5_MAXN];
    const double *A, *U, *S;  // replay: A [n_a][reps], U [n_u][reps], S + soff[i] = node i rows
    long long n_a, n_u, soff[K5_MAXN], n_s[K5_MAXN];
    double *rec;    // [MQ_NET_REC_ROWS(m)][reps]
    double *queue;  // [m][qcap][3][lanes]
    long long lanes;
};

cudaError_t k5_launch(const K5Params &P, int grid, cudaStream_t st);
cudaError_t k5_occupancy(int *blocks_per_sm);

// k5v2_net.cu: the register / shared-memory resident predicated form for <= 8 nodes and <= 16 servers
bool k5v2_supported(int m, int nserv);
cudaError_t k5v2_launch(const K5Params &P, int grid, cudaStream_t st);
cudaError_t k5v2_occupancy(const K5Params &P, int *blocks_per_sm);

}  // namespace mq
