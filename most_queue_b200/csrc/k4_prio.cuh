// k4_prio.cuh -- parameters and record layout of the priority kernel (k4_prio.cu).
#pragma once

#include "mq_device.cuh"

namespace mq {

constexpr int K4_BLOCK = 64;
constexpr int K4_MAXC = 16;
constexpr int K4_MAXK = 8;

struct K4Params {
    int c, K, prty, p_bins, replay, qcap;
    long long buffer;  // <= 0: infinite (priority.py:373,437)
    long long jobs, reps;
    uint32_t rep0, key0;
    DistF src[K4_MAXK], srv[K4_MAXK];
    // replay inputs: A + aoff[k] is class k's gap rows ([n_a[k]][reps]); same for S
    const double *A, *S;
    long long aoff[K4_MAXK], soff[K4_MAXK], n_a[K4_MAXK], n_s[K4_MAXK];
    double *rec;    // [2 + K * rows_per_class][reps]
    double *queue;  // [K][qcap][4][lanes]
    long long lanes;
};

// per-class rows
#define K4_WSUM 0
#define K4_VSUM 4
#define K4_WRUN 8
#define K4_VRUN 12
#define K4_ARRIVED 16
#define K4_TAKED 17
#define K4_SERVED 18
#define K4_DROPPED 19
#define K4_INSYS 20
#define K4_QUEUED 21
#define K4_TOLD 22
#define K4_POVER 23
#define K4_AUSED 24
#define K4_SUSED 25
#define K4_P 26
#define K4_ROWS(pb) (26 + (pb))
// global rows
#define K4_G_TTEK 0
#define K4_G_FLAGS 1
#define K4_G_ROWS 2

cudaError_t k4_launch(const K4Params &P, int grid, cudaStream_t st);
cudaError_t k4_occupancy(int *blocks_per_sm);

// k4v2_prio.cu: the register / shared-memory resident predicated form for c <= 8, K <= 4 (same block size)
bool k4v2_supported(int c, int K);
cudaError_t k4v2_launch(const K4Params &P, int grid, cudaStream_t st);
cudaError_t k4v2_occupancy(int c, int K, bool replay, int *blocks_per_sm);

}  // namespace mq
