// k10_size.cuh -- parameters of the size-based single-server kernel (k10_size.cu).
#pragma once

#include "mq_device.cuh"

namespace mq {

constexpr int K10_BLOCK = 128;
constexpr int K10_JOBF = 8;  // doubles per waiting job: rank, arrival, start of wait, wait, remaining, size, id, prediction

struct K10Params {
    int discipline, p_bins, replay, qcap, pred_kind;
    double pred_sigma;
    long long buffer;
    long long jobs, reps;
    uint32_t rep0, key0;
    DistF src, srv, pred_law;  // pred_law: Exp(1) or N(0,1) noise of the predictor (generator form)
    const double *A, *S, *Y;  // replay: [n_a][reps], [n_s][reps], [n_y][reps] (Y may be null: perfect predictor)
    long long n_a, n_s, n_y;
    double *rec;    // [MQ_Z_ROWS(p_bins)][reps], zeroed by the caller
    double *queue;  // [qcap][K10_JOBF][lanes]
    long long lanes;
};

cudaError_t k10_launch(const K10Params &P, int grid, cudaStream_t st);
cudaError_t k10_occupancy(int *blocks_per_sm);

}  // namespace mq
