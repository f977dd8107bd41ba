// k3_lindley.cuh -- parameters of the max-plus Lindley scan kernels (k3_lindley.cu).
#pragma once

#include "mq_device.cuh"

namespace mq {

constexpr int K3_THREADS = 256;   // segments per tile
constexpr int K3_L = 16;          // jobs per segment (one 128-byte line of fp64 per array per thread)
constexpr int K3_TOP = 1024;      // threads of the tile-level scan
constexpr int K3_NSTAT = 16;      // per-block partial sums
constexpr int K3_NLEVEL = 33;     // T[N >= 1..32] and the time-weighted excess above 32

struct MP {
    double a, b;  // the map w -> max(a, w + b)
};

struct K3Params {
    long long n;        // jobs in this range
    long long job0;     // global index of the first one (Philox counter)
    int replay;
    uint32_t key;       // key0 + STRIDE * 64 * chain
    uint32_t rk[10];    // key + r * MQ_PHILOX_W: the round keys of the primary Philox call
    DistF src, srv;
    const double *A;    // replay: A[0] = gap before job job0; n + 1 entries
    const double *S;    // replay: n entries
    MP *tile_sum;       // [ntiles]
    MP *tile_ex;        // [ntiles]
    MP *total;          // [1]
    MP *chunk_f;        // [K3_TOP] folds of the 1024 chunks of tile summaries
    MP *chunk_ex;       // [K3_TOP] their exclusive prefixes
    MP *thread_ex;      // optional [ntiles][K3_THREADS]: every segment's exclusive prefix inside its tile, kept by the
                        // summary pass so that the walk neither folds nor scans again (4 KB per 64 KB tile)
    double w_in;
    double *partial;    // [grid][K3_NSTAT]
    double *tail;       // [2]: W[n-1] + x[n-1] (before clamping), W[n-1] + S[n-1]
    double *W;          // optional per-job waits
    long long ntiles;
    long long n_ahead;       // jobs of the chain after this range (their gaps follow A[n])
    double *level_partial;   // optional [grid][K3_NLEVEL]: time-in-state on the scan path
    // generator form, optional: the summary pass stashes its fp32 draws (gap of job i in stG[i], i = 0..n; service in
    // stV[i]) and the walk pass reads them back instead of drawing everything a second time
    float *stG, *stV;
    // generator form, "speculative" mode (spec != 0): the summary pass walks every segment at once from its in-tile
    // prefix and leaves its power sums in spec_partial[grid_sum][K3_NSTAT]; k3_repair then re-draws and re-walks only the
    // segments whose start wait depends on the carry into their tile (see k3_onepass.cu for the argument).  No stash.
    int spec;
    double *spec_partial;
};

// number of valid entries among [base, base + span) when the valid ones are [0, limit)
__host__ __device__ __forceinline__ int k3_valid_count(long long limit, long long base, int span)
{
    const long long r = limit - base;
    return r <= 0 ? 0 : (r >= span ? span : (int)r);
}

// k3_onepass.cu: the replay form in one pass over A and S (TMA-staged tiles, decoupled look-back); same tree, same values
bool k3_onepass_usable(const K3Params &P);
size_t k3_onepass_scratch_bytes(long long ntiles);
cudaError_t k3_onepass_occupancy(int *blocks);
cudaError_t k3_launch_onepass(const K3Params &P, void *scratch, int grid, cudaStream_t st);
int k3_onepass_error();

cudaError_t k3_launch_summary(const K3Params &P, int grid, cudaStream_t st);
cudaError_t k3_launch_top(const K3Params &P, cudaStream_t st);
cudaError_t k3_launch_walk(const K3Params &P, int grid, cudaStream_t st);
cudaError_t k3_launch_repair(const K3Params &P, int grid, cudaStream_t st);
int k3_repair_grid(long long ntiles, int sms);
cudaError_t k3_launch_final(const double *partial, const double *level_partial, int grid, double *stats,
                            cudaStream_t st);
cudaError_t k3_occupancy(bool replay, int stage_mode, bool want_p, int src_kind, int srv_kind, int *summary_blocks, int *walk_blocks);
size_t k3_walk_smem(bool replay, bool want_p);

}  // namespace mq
