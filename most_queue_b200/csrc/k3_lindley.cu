// k3_lindley.cu -- K3: max-plus Lindley parallel scan for ONE very long GI/G/1 FIFO chain.
//
// With one channel and an infinite buffer QsSim.run() (most_queue/sim/base.py:488-528) is the
// Lindley recursion  W[n+1] = max(0, W[n] + x[n]),  x[n] = S[n] - A[n+1]  (wait = start - arrival,
// base.py:302; the start of job n+1 is the departure of job n), v[n] = W[n] + S[n] (base.py:272).
// The reference can only walk it sequentially (base.py:505).  Each step is the map
// w -> max(a, w + b) with (a, b) = (0, x[n]); maps compose associatively in max-plus algebra:
//   later o earlier = (max(a2, a1 + b2), b1 + b2),
// so the chain is cut into segments of K3_L jobs (one thread each), segment summaries are scanned
// with a FIXED association tree (Kogge-Stone over the 32 lanes of a warp, sequential over the 8 warps
// of a tile, then the tile level: 1024 chunks folded sequentially, Kogge-Stone over 32, sequential over
// 32), and every segment is then walked sequentially from its carry.  fp64 '+' is not associative,
// so the tree is part of the specification: oracle/mq_oracle_scan.c evaluates the same tree and the
// per-job waits agree bit for bit.
//
// Three launches: (1) tile summaries, (2) tile-level scan (one CTA), (3) walk + moment sums.  The
// replay form reads A and S twice (32 B/job); the generator form regenerates the variates with the
// counter-based generator (counter = global job index, so any range can be evaluated anywhere).
#include "k3_lindley.cuh"

namespace mq {

__device__ __forceinline__ MP mp_id() { return MP{-INFINITY, 0.0}; }

// max(0, x) and max(a, s) without fmax's NaN handling (no operand is ever a NaN): a shift and two ANDs, or one compare
// and two selects, instead of a seven-instruction sequence on every step of the recursion; same bits (k3_onepass.cu)
__device__ __forceinline__ double max0(double x)
{
    const int hi = __double2hiint(x);
    const int keep = ~(hi >> 31);
    return __hiloint2double(hi & keep, __double2loint(x) & keep);
}
__device__ __forceinline__ double max2(double a, double s) { return s > a ? s : a; }

__device__ __forceinline__ MP mp_then(MP earlier, MP later)
{
    return MP{max2(later.a, __dadd_rn(earlier.a, later.b)), __dadd_rn(earlier.b, later.b)};
}

__device__ __forceinline__ MP mp_step(MP f, double x)
{
    return MP{max0(__dadd_rn(f.a, x)), __dadd_rn(f.b, x)};
}

__device__ __forceinline__ double mp_apply(MP f, double w) { return max2(f.a, __dadd_rn(w, f.b)); }

__device__ __forceinline__ MP mp_shfl_up(MP v, int d)
{
    return MP{__shfl_up_sync(0xffffffffu, v.a, d), __shfl_up_sync(0xffffffffu, v.b, d)};
}

// exclusive prefix of `mine` over the CTA (groups of 32 lanes: Kogge-Stone; groups: sequential).
// Returns the exclusive prefix; *total receives the fold of all values (valid in every thread).
template <int NT>
__device__ __forceinline__ MP cta_scan(MP mine, MP *smem, MP *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    MP inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        MP up = mp_shfl_up(inc, d);
        if (lane >= d) inc = mp_then(up, inc);
    }
    MP within = mp_shfl_up(inc, 1);
    if (lane == 0) within = mp_id();
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    MP carry = mp_id(), tot = mp_id();
    for (int g = 0; g < NT / 32; ++g) {
        if (g == warp) carry = tot;
        tot = mp_then(tot, smem[g]);
    }
    __syncthreads();
    *total = tot;
    return mp_then(carry, within);
}

// Replay form: the CTA stages the tile's S[0..4096) and A[0..4096] in shared memory with coalesced
// streaming loads (a warp reads 256 contiguous bytes per instruction) and each thread then reads its own
// 16-job segment.  One pad word per 16 entries (index e + e/16) makes the strided per-thread reads
// conflict free (two wavefronts per 64-bit warp access, the minimum).
constexpr int K3_TILE = K3_THREADS * K3_L;
constexpr int K3_HALO = 256;  // gaps staged beyond the tile for the level walk (WANT_P)
constexpr int K3_PADDED = K3_TILE + K3_TILE / 16 + 2;
constexpr int K3_PADDED_A = (K3_TILE + K3_HALO) + (K3_TILE + K3_HALO) / 16 + 2;
constexpr int K3_PBL = 32;    // levels N >= 1..32 accumulated per lane in shared memory
constexpr size_t K3_STAGE_BYTES = sizeof(double) * (K3_PADDED + K3_PADDED_A);
#ifndef K3_SUM_MINB
#define K3_SUM_MINB 4
#endif
constexpr size_t K3_XSTAGE_BYTES = sizeof(double) * K3_PADDED;  // summary pass, replay form: x = S - A' only
constexpr size_t K3_LEVEL_BYTES = sizeof(float) * K3_PBL * K3_THREADS + sizeof(double) * (K3_THREADS / 32) * (K3_PBL + 1);

__device__ __forceinline__ int k3_pad(int e) { return e + (e >> 4); }

// number of valid entries among [base, base + span) when the valid ones are [0, limit): a 32-bit count, so that the
// per-job bounds checks are one compare against it instead of 64-bit index arithmetic
__device__ __forceinline__ int k3_valid(long long limit, long long base, int span)
{
    const long long r = limit - base;
    return r <= 0 ? 0 : (r >= span ? span : (int)r);
}

__device__ __forceinline__ double gen_gap(const K3Params &P, long long n);

// the same staging from the fp32 stash written by the summary pass of the generator form (gaps beyond the range --
// the halo of the level walk in the last tile -- are drawn)
// HALO: also stage the 256 gaps beyond the tile that the level walk (WANT_P) reads; otherwise only the one gap that the
// last job's x needs
template <bool HALO>
__device__ __forceinline__ void stage_tile_stash(const K3Params &P, long long tile_base, double *smS, double *smA)
{
    if (tile_base + K3_TILE + K3_HALO <= P.n) {
        // interior tile (uniform in the CTA): every index is inside the stash, so all loads are issued back to back
        // before the first conversion waits for one of them
        const float *G = P.stG + tile_base, *V = P.stV + tile_base;
        float vS[K3_L], vA[K3_L];
#pragma unroll
        for (int k = 0; k < K3_L; ++k) {
            vS[k] = __ldcs(V + k * K3_THREADS + threadIdx.x);
            vA[k] = __ldcs(G + k * K3_THREADS + threadIdx.x);
        }
        float h = 0.0f;
        if (HALO || threadIdx.x == 0) h = __ldcs(G + K3_TILE + threadIdx.x);
#pragma unroll
        for (int k = 0; k < K3_L; ++k) {
            const int e = k * K3_THREADS + threadIdx.x;
            smS[k3_pad(e)] = (double)vS[k];
            smA[k3_pad(e)] = (double)vA[k];
        }
        if (HALO || threadIdx.x == 0) smA[k3_pad(K3_TILE + threadIdx.x)] = (double)h;
        if (HALO && threadIdx.x == 0) smA[k3_pad(K3_TILE + K3_HALO)] = (double)__ldcs(G + K3_TILE + K3_HALO);
        __syncthreads();
        return;
    }
    auto gap_at = [&](long long n) -> double {
        if (n <= P.n) return (double)__ldcs(P.stG + n);
        return n <= P.n + P.n_ahead ? gen_gap(P, n) : 0.0;
    };
#pragma unroll
    for (int k = 0; k < K3_L; ++k) {
        const int e = k * K3_THREADS + threadIdx.x;
        const long long n = tile_base + e;
        smS[k3_pad(e)] = n < P.n ? (double)__ldcs(P.stV + n) : 0.0;
        smA[k3_pad(e)] = gap_at(n);
    }
    if (HALO || threadIdx.x == 0) {
        const int e = K3_TILE + threadIdx.x;
        smA[k3_pad(e)] = gap_at(tile_base + e);
        if (HALO && threadIdx.x == 0) smA[k3_pad(K3_TILE + K3_HALO)] = gap_at(tile_base + K3_TILE + K3_HALO);
    }
    __syncthreads();
}

template <bool STASH = false, bool HALO = true>
__device__ __forceinline__ void stage_tile(const K3Params &P, long long tile_base, double *smS, double *smA)
{
    if (STASH) {
        stage_tile_stash<HALO>(P, tile_base, smS, smA);
        return;
    }
    const int vs = k3_valid(P.n, tile_base, K3_TILE);                                    // services in range
    const int va = k3_valid(P.n + P.n_ahead + 1, tile_base, K3_TILE + K3_HALO + 1);      // gaps in range
    const double *S = P.S + tile_base, *A = P.A + tile_base;
#pragma unroll
    for (int k = 0; k < K3_L; ++k) {
        const int e = k * K3_THREADS + threadIdx.x;
        smS[k3_pad(e)] = e < vs ? __ldcs(S + e) : 0.0;
        smA[k3_pad(e)] = e < va ? __ldcs(A + e) : 0.0;
    }
    if (HALO || threadIdx.x == 0) {   // one more gap for x of the last job, and the halo used by the level walk
        const int e = K3_TILE + threadIdx.x;
        smA[k3_pad(e)] = e < va ? __ldcs(A + e) : 0.0;
        if (HALO && threadIdx.x == 0) smA[k3_pad(K3_TILE + K3_HALO)] = K3_TILE + K3_HALO < va ? __ldcs(A + K3_TILE + K3_HALO) : 0.0;
    }
    __syncthreads();
}

// x[j] = S[n] - A[n+1] and S[j] for the K3_L jobs of this thread's segment; returns the sum of the
// segment's gaps A[n]
__device__ __forceinline__ double gen_gap(const K3Params &P, long long n)
{
    const unsigned long long g = (unsigned long long)(P.job0 + n);
    uint2 w = philox2x32_10_rk((uint32_t)g, (uint32_t)(g >> 32), P.rk);
    return (double)sample<-1>(P.src, w.x, (uint32_t)g, (uint32_t)(g >> 32), P.key, 0);
}

// FULL: every job of the tile is inside the range (all tiles but the last one): no per-job bounds logic
template <bool REPLAY, bool STORE_GAPS = false, bool STASH_W = false, bool FULL = false>
__device__ __forceinline__ double load_segment(const K3Params &P, long long n0, double (&x)[K3_L], double (&s)[K3_L],
                                               const double *smS, double *smA, float *stg = nullptr,
                                               float *stv = nullptr)
{
    double sum_a = 0.0;
    if (REPLAY) {
        const int e0 = threadIdx.x * K3_L;
        double an = smA[k3_pad(e0)];
        const int valid = FULL ? K3_L : k3_valid(P.n, n0, K3_L);
#pragma unroll
        for (int j = 0; j < K3_L; ++j) {
            const bool ok = FULL || j < valid;
            s[j] = smS[k3_pad(e0 + j)];
            sum_a += ok ? an : 0.0;
            an = smA[k3_pad(e0 + j + 1)];
            x[j] = ok ? __dsub_rn(s[j], an) : 0.0;
        }
    } else {
        // word 0 of call n = gap before job n, word 1 of call n = service of job n
        unsigned long long g = (unsigned long long)(P.job0 + n0);
        uint2 w = philox2x32_10_rk((uint32_t)g, (uint32_t)(g >> 32), P.rk);
        double gap = (double)sample<-1>(P.src, w.x, (uint32_t)g, (uint32_t)(g >> 32), P.key, 0);
        const int valid = FULL ? K3_L : k3_valid(P.n, n0, K3_L);
#pragma unroll
        for (int j = 0; j < K3_L; ++j) {
            const long long n = n0 + j;
            const bool ok = FULL || j < valid;
            const double sv = (double)sample<-1>(P.srv, w.y, (uint32_t)g, (uint32_t)(g >> 32), P.key, 1);
            ++g;
            w = philox2x32_10_rk((uint32_t)g, (uint32_t)(g >> 32), P.rk);
            const double gnext = (double)sample<-1>(P.src, w.x, (uint32_t)g, (uint32_t)(g >> 32), P.key, 0);
            s[j] = ok ? sv : 0.0;
            sum_a += ok ? gap : 0.0;
            x[j] = ok ? __dsub_rn(sv, gnext) : 0.0;
            if (STORE_GAPS) smA[k3_pad(threadIdx.x * K3_L + j)] = gap;
            if (STASH_W) {  // both values came from fp32 samplers: the casts are exact.  Staged in shared memory
                // (padded: conflict free) and written out by the CTA with coalesced stores
                stg[k3_pad(threadIdx.x * K3_L + j)] = (float)gap;
                stv[k3_pad(threadIdx.x * K3_L + j)] = (float)sv;
                if (!FULL && n == P.n - 1) P.stG[P.n] = (float)gnext;
            }
            gap = gnext;
        }
        if (STORE_GAPS) smA[k3_pad(threadIdx.x * K3_L + K3_L)] = gap;
    }
    return sum_a;
}

template <bool FULL = false>
__device__ __forceinline__ MP fold_segment(const K3Params &P, long long n0, const double (&x)[K3_L])
{
    MP f = mp_id();
    const int valid = FULL ? K3_L : k3_valid(P.n, n0, K3_L);
#pragma unroll
    for (int j = 0; j < K3_L; ++j)
        if (FULL || j < valid) f = mp_step(f, x[j]);
    return f;
}

// Generator form of the summary pass: draw, fold and (optionally) stash one job at a time in a ROLLED loop.  The
// unrolled load_segment + fold_segment pair inlines the sampler switch 32 times (the kernel then stalls on
// instruction fetch); the fold is sequential in j either way, so the association order is unchanged.
#ifndef K3_GEN_UNROLL
#define K3_GEN_UNROLL 1
#endif
constexpr int K3_GEN_UNROLL_N = K3_GEN_UNROLL;
template <int STAGE, bool FULL, int SRCK, int SRVK>
__device__ __forceinline__ MP gen_fold_segment(const K3Params &P, long long n0, float *stg, float *stv)
{
    unsigned long long g = (unsigned long long)(P.job0 + n0);
    uint2 w = philox2x32_10_rk((uint32_t)g, (uint32_t)(g >> 32), P.rk);
    float gap = sample<SRCK>(P.src, w.x, (uint32_t)g, (uint32_t)(g >> 32), P.key, 0);
    const int valid = FULL ? K3_L : k3_valid(P.n, n0, K3_L);
    MP f = mp_id();
    if (SRCK == MQ_GAMMA) {
        // Gamma gaps: the rejection loop is the caller's.  A trip of this loop is ONE attempt at the gap that follows
        // job j of the lane; a lane whose attempt is rejected (a few per cent) tries again in the next trip while its
        // neighbours go on with their next job, so a warp makes 16 + (the most rejections of any of its lanes) trips
        // instead of 16 x (attempts of its unluckiest lane per job).  Same draws, same fold order per lane.
        const float dd = P.src.a, cc = P.src.b;
        const int small_alpha = P.src.r;
        int j = 0;
        uint32_t t = 0;
        float x = dd;
        float sv = sample<SRVK>(P.srv, w.y, (uint32_t)g, (uint32_t)(g >> 32), P.key, 1);
        ++g;
        w = philox2x32_10_rk((uint32_t)g, (uint32_t)(g >> 32), P.rk);
        while (__any_sync(0xffffffffu, j < K3_L)) {
            if (j < K3_L) {
                const bool ok = gamma_attempt(dd, cc, small_alpha, t, w.x, (uint32_t)g, (uint32_t)(g >> 32), P.key, 0, x) || t == 14;
                if (ok) {
                    const float gnext = gamma_finish(x, P.src.c, P.src.d, small_alpha, w.x);
                    if (FULL || j < valid) f = mp_step(f, __dsub_rn((double)sv, (double)gnext));
                    if (STAGE) {
                        stg[k3_pad(threadIdx.x * K3_L + j)] = gap;
                        stv[k3_pad(threadIdx.x * K3_L + j)] = sv;
                        if (STAGE == 1 && !FULL && n0 + j == P.n - 1) P.stG[P.n] = gnext;
                    }
                    gap = gnext;
                    ++j;
                    t = 0;
                    x = dd;
                    if (j < K3_L) {
                        sv = sample<SRVK>(P.srv, w.y, (uint32_t)g, (uint32_t)(g >> 32), P.key, 1);
                        ++g;
                        w = philox2x32_10_rk((uint32_t)g, (uint32_t)(g >> 32), P.rk);
                    }
                } else {
                    ++t;
                }
            }
        }
        if (STAGE == 2 && threadIdx.x == K3_THREADS - 1) stg[k3_pad(K3_TILE)] = gap;  // gap of the job after the tile
        return f;
    }
#pragma unroll K3_GEN_UNROLL_N
    for (int j = 0; j < K3_L; ++j) {
        const float sv = sample<SRVK>(P.srv, w.y, (uint32_t)g, (uint32_t)(g >> 32), P.key, 1);
        ++g;
        w = philox2x32_10_rk((uint32_t)g, (uint32_t)(g >> 32), P.rk);
        const float gnext = sample<SRCK>(P.src, w.x, (uint32_t)g, (uint32_t)(g >> 32), P.key, 0);
        if (FULL || j < valid) f = mp_step(f, __dsub_rn((double)sv, (double)gnext));
        if (STAGE) {
            stg[k3_pad(threadIdx.x * K3_L + j)] = gap;
            stv[k3_pad(threadIdx.x * K3_L + j)] = sv;
            if (STAGE == 1 && !FULL && n0 + j == P.n - 1) P.stG[P.n] = gnext;
        }
        gap = gnext;
    }
    if (STAGE == 2 && threadIdx.x == K3_THREADS - 1) stg[k3_pad(K3_TILE)] = gap;
    return f;
}

// SRCK / SRVK: the laws of the gaps / services known at compile time (-1: switch on the kind at run time; the switch is
// 9 % of the generator form's instructions, so the BASELINE shapes get their own instantiation)
#ifndef K3_SPEC_MINB
#define K3_SPEC_MINB 5
#endif
// STAGE: 0 nothing kept of the draws; 1 stash them in HBM for the walk pass; 2 "speculative": keep them in shared memory,
// walk every segment from its in-tile prefix right here and leave the repair of the carry-dependent segments to k3_repair
template <bool REPLAY, int STAGE = 0, int SRCK = -1, int SRVK = -1>
__global__ void __launch_bounds__(K3_THREADS, (REPLAY ? K3_SUM_MINB : (STAGE == 2 ? K3_SPEC_MINB : 6))) k3_summary(const K3Params P)
{
    constexpr bool STASH_W = STAGE == 1;
    __shared__ MP smem[K3_THREADS / 32];
    extern __shared__ double stage[];
    double *smS = stage;
    __shared__ float stash_stage[STAGE ? 2 * K3_PADDED : 2];
    float *stg = stash_stage, *stv = stash_stage + (STAGE ? K3_PADDED : 1);
    __shared__ double spec_red[STAGE == 2 ? K3_THREADS / 32 : 1][10];
    if (STAGE == 2 && threadIdx.x < (K3_THREADS / 32) * 10) (&spec_red[0][0])[threadIdx.x] = 0.0;
    for (long long t = blockIdx.x; t < P.ntiles; t += gridDim.x) {
        const long long n0 = (t * K3_THREADS + threadIdx.x) * K3_L;
        const bool full_tile = (t + 1) * (long long)K3_TILE < P.n;  // uniform in the CTA
        MP mine;
        if (!REPLAY) {
            mine = full_tile ? gen_fold_segment<STAGE, true, SRCK, SRVK>(P, n0, stg, stv)
                             : gen_fold_segment<STAGE, false, SRCK, SRVK>(P, n0, stg, stv);
        } else {
            // replay form: the fold only needs x[n] = S[n] - A[n+1], formed while the coalesced loads arrive and staged
            // as ONE array (35 KB instead of 72 KB per CTA: 6 resident CTAs instead of 3); jobs beyond the range are 0
            const long long tb = t * K3_TILE;
            const int vs = k3_valid(P.n, tb, K3_TILE);
            const double *S = P.S + tb, *A = P.A + tb + 1;
#pragma unroll
            for (int h = 0; h < 2; ++h) {  // two batches of 16 loads: enough in flight, half the registers
                double vS[K3_L / 2], vA[K3_L / 2];
#pragma unroll
                for (int k = 0; k < K3_L / 2; ++k) {
                    const int e = (h * (K3_L / 2) + k) * K3_THREADS + threadIdx.x;
                    vS[k] = e < vs ? __ldcs(S + e) : 0.0;
                    vA[k] = e < vs ? __ldcs(A + e) : 0.0;
                }
#pragma unroll
                for (int k = 0; k < K3_L / 2; ++k)
                    smS[k3_pad((h * (K3_L / 2) + k) * K3_THREADS + threadIdx.x)] = __dsub_rn(vS[k], vA[k]);
            }
            __syncthreads();
            const int e0 = threadIdx.x * K3_L;
            const int valid = full_tile ? K3_L : k3_valid(P.n, n0, K3_L);
            mine = mp_id();
#pragma unroll
            for (int j = 0; j < K3_L; ++j)
                if (j < valid) mine = mp_step(mine, smS[k3_pad(e0 + j)]);
        }
        MP total;
        const MP ex_seg = cta_scan<K3_THREADS>(mine, smem, &total);  // its barriers also fence the stage reuse
        if (P.thread_ex)
            __stcs(reinterpret_cast<double2 *>(P.thread_ex) + t * K3_THREADS + threadIdx.x, make_double2(ex_seg.a, ex_seg.b));
        if (STASH_W) {
            // (the scan's barriers have made the staged draws visible) one 128-byte line per warp and instruction
            const long long tb = t * K3_TILE;
#pragma unroll
            for (int k = 0; k < K3_L; ++k) {
                const int e = k * K3_THREADS + threadIdx.x;
                if (tb + e < P.n) {
                    __stcs(P.stG + tb + e, stg[k3_pad(e)]);
                    __stcs(P.stV + tb + e, stv[k3_pad(e)]);
                }
            }
            __syncthreads();  // before the next tile overwrites the staging arrays
        }
        if (STAGE == 2) {
            // (the scan's barriers have made the staged draws visible) speculative walk: the wait at the start of the
            // segment is max(a_s, w_tile + b_s), and after the first idle period of the tile that is a_s whatever the
            // carry w_tile is.  Segment 0 has no a_s and the last tile has bounds: both are walked by k3_repair alone.
            if (full_tile) {
                const int e0 = threadIdx.x * K3_L;
                double acc[10];
#pragma unroll
                for (int i = 0; i < 10; ++i) acc[i] = 0.0;
                if (threadIdx.x != 0) {
                    double w = ex_seg.a;
                    float g = stg[k3_pad(e0)];
#pragma unroll 4
                    for (int j = 0; j < K3_L; ++j) {
                        const float sv = stv[k3_pad(e0 + j)];
                        const float gn = stg[k3_pad(e0 + j + 1)];
                        const double sd = (double)sv;
                        const double v = __dadd_rn(w, sd);
                        const double w2 = w * w, v2 = v * v;
                        acc[0] += w;
                        acc[1] += w2;
                        acc[2] += w2 * w;
                        acc[3] += w2 * w2;
                        acc[4] += v;
                        acc[5] += v2;
                        acc[6] += v2 * v;
                        acc[7] += v2 * v2;
                        acc[8] += (double)g;
                        acc[9] += sd;
                        w = max0(__dadd_rn(w, __dsub_rn(sd, (double)gn)));
                        g = gn;
                    }
                }
                // the sums of the tile go to per-warp accumulators at once (fixed order): the kernel runs at 40 registers
                const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
                for (int i = 0; i < 10; ++i) {
                    double v = acc[i];
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
                    if (lane == 0) spec_red[warp][i] += v;
                }
            }
            __syncthreads();  // before the next tile overwrites the staging arrays
        }
        if (threadIdx.x == 0) P.tile_sum[t] = total;
    }
    if (STAGE == 2) {
        __syncthreads();
        if (threadIdx.x < K3_NSTAT) {
            double v = 0.0;
            if (threadIdx.x < 10)
                for (int wv = 0; wv < K3_THREADS / 32; ++wv) v += spec_red[wv][threadIdx.x];
            P.spec_partial[(long long)blockIdx.x * K3_NSTAT + threadIdx.x] = v;
        }
    }
}

// Speculative mode, second kernel: one warp per tile tests every segment's start wait max(a_s, w_tile + b_s) against the
// a_s the summary pass walked it from and re-draws and re-walks the segments where they differ (1.2 of 256 at rho = 0.7),
// adding (true - provisional) to the sums; segment 0 of every tile and the whole last tile are walked here only.
__global__ void __launch_bounds__(K3_THREADS) k3_repair(const K3Params P)
{
    __shared__ double red[K3_THREADS / 32][K3_NSTAT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] = 0.0;
    const long long warps = (long long)gridDim.x * (K3_THREADS / 32);
    for (long long t = (long long)blockIdx.x * (K3_THREADS / 32) + warp; t < P.ntiles; t += warps) {
        const bool full_tile = (t + 1) * (long long)K3_TILE < P.n;
        const double w_tile = mp_apply(P.tile_ex[t], P.w_in);
        const double2 *exs = reinterpret_cast<const double2 *>(P.thread_ex) + t * K3_THREADS;
#pragma unroll 1
        for (int q = 0; q < K3_THREADS / 32; ++q) {
            const int seg = q * 32 + lane;
            const double2 e = __ldcs(exs + seg);
            const double w0 = mp_apply(MP{e.x, e.y}, w_tile);
            const bool own = !full_tile || seg == 0;        // nothing provisional was accumulated for this segment
            if (!own && w0 == e.x) continue;
            const long long n0 = (t * K3_THREADS + seg) * K3_L;
            const int valid = full_tile ? K3_L : k3_valid(P.n, n0, K3_L);
            const int last_j = (!full_tile && P.n - 1 >= n0 && P.n - 1 < n0 + K3_L) ? (int)(P.n - 1 - n0) : -1;
            unsigned long long g = (unsigned long long)(P.job0 + n0);
            uint2 wd = philox2x32_10_rk((uint32_t)g, (uint32_t)(g >> 32), P.rk);
            float gap = sample<-1>(P.src, wd.x, (uint32_t)g, (uint32_t)(g >> 32), P.key, 0);
            double w = w0, wp = e.x;
#pragma unroll 1
            for (int j = 0; j < valid; ++j) {
                const float sv = sample<-1>(P.srv, wd.y, (uint32_t)g, (uint32_t)(g >> 32), P.key, 1);
                ++g;
                wd = philox2x32_10_rk((uint32_t)g, (uint32_t)(g >> 32), P.rk);
                const float gn = sample<-1>(P.src, wd.x, (uint32_t)g, (uint32_t)(g >> 32), P.key, 0);
                const double sd = (double)sv, x = __dsub_rn(sd, (double)gn);
                const double v = __dadd_rn(w, sd);
                const double w2 = w * w, v2 = v * v;
                double tt[8] = {w, w2, w2 * w, w2 * w2, v, v2, v2 * v, v2 * v2};
                if (own) {
                    acc[8] += (double)gap;
                    acc[9] += sd;
                } else {
                    const double vp = __dadd_rn(wp, sd);
                    const double p2 = wp * wp, q2 = vp * vp;
                    tt[0] -= wp;
                    tt[1] -= p2;
                    tt[2] -= p2 * wp;
                    tt[3] -= p2 * p2;
                    tt[4] -= vp;
                    tt[5] -= q2;
                    tt[6] -= q2 * vp;
                    tt[7] -= q2 * q2;
                    wp = max0(__dadd_rn(wp, x));
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] += tt[i];
                const double raw = __dadd_rn(w, x);
                if (j == last_j) {
                    P.tail[0] = raw;
                    P.tail[1] = v;
                }
                w = max0(raw);
                gap = gn;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        double v = acc[i];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < K3_NSTAT) {
        double v = 0.0;
        if (threadIdx.x < 10)
            for (int wv = 0; wv < K3_THREADS / 32; ++wv) v += red[wv][threadIdx.x];
        P.partial[(long long)blockIdx.x * K3_NSTAT + threadIdx.x] = v;
    }
}

// Tile-level scan, the specification's "1024 chunks folded sequentially, Kogge-Stone over 32, sequential over 32", in
// three small launches so that the strided gathers of the chunk folds are spread over 32 SMs instead of queueing behind
// one SM's memory pipe (one 1024-thread CTA took 220 us for 65 536 tiles; the operations and their order are unchanged):
//   k3_top_fold  <<<32, 32>>>   thread c folds the summaries of chunk c sequentially          -> chunk_f[c]
//   k3_top_scan  <<<1, 1024>>>  the CTA scan over the 1024 chunk folds                        -> chunk_ex[c], total
//   k3_top_apply <<<32, 32>>>   thread c walks its chunk again and writes every tile's prefix -> tile_ex[t]
constexpr int K3_TOP_CTA = 32;
constexpr int K3_TOP_U = 16;  // summaries loaded ahead of the (sequential, order-fixed) fold: the loops are load-latency bound

__global__ void __launch_bounds__(K3_TOP_CTA) k3_top_fold(const K3Params P)
{
    const int c = blockIdx.x * K3_TOP_CTA + threadIdx.x;
    const long long chunk = (P.ntiles + K3_TOP - 1) / K3_TOP;
    const long long lo = (long long)c * chunk;
    const long long hi = lo + chunk < P.ntiles ? lo + chunk : P.ntiles;
    constexpr int U = K3_TOP_U;
    MP f = mp_id();
    for (long long t = lo; t < hi; t += U) {
        MP v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (t + u < hi) v[u] = P.tile_sum[t + u];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (t + u < hi) f = mp_then(f, v[u]);
    }
    P.chunk_f[c] = f;
}

__global__ void __launch_bounds__(K3_TOP) k3_top_scan(const K3Params P)
{
    __shared__ MP smem[K3_TOP / 32];
    MP total;
    const MP ex = cta_scan<K3_TOP>(P.chunk_f[threadIdx.x], smem, &total);
    P.chunk_ex[threadIdx.x] = ex;
    if (threadIdx.x == 0) *P.total = total;
}

__global__ void __launch_bounds__(K3_TOP_CTA) k3_top_apply(const K3Params P)
{
    const int c = blockIdx.x * K3_TOP_CTA + threadIdx.x;
    const long long chunk = (P.ntiles + K3_TOP - 1) / K3_TOP;
    const long long lo = (long long)c * chunk;
    const long long hi = lo + chunk < P.ntiles ? lo + chunk : P.ntiles;
    constexpr int U = K3_TOP_U;
    MP ex = P.chunk_ex[c];
    for (long long t = lo; t < hi; t += U) {
        MP v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (t + u < hi) v[u] = P.tile_sum[t + u];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (t + u < hi) {
                P.tile_ex[t + u] = ex;
                ex = mp_then(ex, v[u]);
            }
        }
    }
}

// Time-in-state on the scan path.  With one server and FIFO, N(t) >= m while t lies in the (m-1)-th
// interarrival interval after job k's arrival iff job k is still in the system, so
//   T[N >= m] = sum over jobs k of | [t_k, t_k + v_k) intersected with [t_{k+m-1}, t_{k+m}) |.
// Each job walks forward over the gaps that follow it until its sojourn is used up (1 + arrivals during the
// sojourn steps); per-lane fp32 columns in shared memory, flushed to fp64 per warp.
// MODE: 0 = generator (everything is drawn), 1 = replay (caller arrays), 2 = the generator form walking its stash
template <int MODE>
__device__ __noinline__ void level_walk(const K3Params &P, double v, int e, long long n, const double *smA,
                                        float *lev, double &over)
{
    const long long n_all = P.n + P.n_ahead;  // jobs of the whole chain from this range's origin
    double cum = 0.0;
    for (int m = 1;; ++m) {
        const long long nn = n + m;
        double g;
        if (nn >= n_all)
            g = __longlong_as_double(0x7ff0000000000000ll);  // no counted arrival after the last job
        else if (e + m <= K3_TILE + K3_HALO)
            g = smA[k3_pad(e + m)];
        else
            g = MODE == 1 ? __ldg(P.A + nn) : ((MODE == 2 && nn <= P.n) ? (double)__ldg(P.stG + nn) : gen_gap(P, nn));
        const double hi = cum + g;
        const double ov = fmin(v, hi) - cum;
        if (m <= K3_PBL)
            lev[(m - 1) * K3_THREADS] += (float)ov;
        else
            over += ov;
        if (!(v > hi)) break;
        cum = hi;
    }
}

struct K3True {
    static constexpr bool value = true;
};
struct K3False {
    static constexpr bool value = false;
};

template <int MODE, bool WANT_P>
__global__ void __launch_bounds__(K3_THREADS, (MODE == 0 ? 1 : ((MODE == 2 && !WANT_P) ? 3 : 2))) k3_walk(const K3Params P)
{
    constexpr bool REPLAY = MODE != 0;  // staged inputs: caller arrays or the stash
    __shared__ MP smem[K3_THREADS / 32];
    __shared__ double red[K3_THREADS / 32][K3_NSTAT];
    double acc[K3_NSTAT];
#pragma unroll
    for (int i = 0; i < K3_NSTAT; ++i) acc[i] = 0.0;

    extern __shared__ double stage[];
    double *smS = stage, *smA = stage + K3_PADDED;
    float *lev = reinterpret_cast<float *>(stage + K3_PADDED + K3_PADDED_A) + threadIdx.x;  // [K3_PBL][lane]
    double *lev_w = reinterpret_cast<double *>(reinterpret_cast<float *>(stage + K3_PADDED + K3_PADDED_A) +
                                               K3_PBL * K3_THREADS);                       // [warp][K3_PBL + 1]
    double over = 0.0;
    int since_flush = 0;
    if (WANT_P) {
        for (int m = 0; m < K3_PBL; ++m) lev[m * K3_THREADS] = 0.0f;
        if (threadIdx.x < (K3_THREADS / 32) * (K3_PBL + 1)) lev_w[threadIdx.x] = 0.0;
        for (int i = threadIdx.x + K3_THREADS; i < (K3_THREADS / 32) * (K3_PBL + 1); i += K3_THREADS) lev_w[i] = 0.0;
    }
    auto flush_levels = [&]() {
        const int lane_ = threadIdx.x & 31, warp_ = threadIdx.x >> 5;
        for (int m = 0; m < K3_PBL; ++m) {
            double v = (double)lev[m * K3_THREADS];
            lev[m * K3_THREADS] = 0.0f;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
            if (lane_ == 0) lev_w[warp_ * (K3_PBL + 1) + m] += v;
        }
        double v = over;
        over = 0.0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
        if (lane_ == 0) lev_w[warp_ * (K3_PBL + 1) + K3_PBL] += v;
    };
    for (long long t = blockIdx.x; t < P.ntiles; t += gridDim.x) {
        const long long n0 = (t * K3_THREADS + threadIdx.x) * K3_L;
        double x[K3_L], s[K3_L];
        if (REPLAY) stage_tile<MODE == 2, WANT_P>(P, t * K3_TILE, smS, smA);
        const bool full_tile = (t + 1) * (long long)K3_TILE < P.n;  // uniform in the CTA
        // staged forms (replay, stash): the tile sits in shared memory, so nothing is carried in registers across the
        // CTA scan -- the fold and the walk both read their operands from there (halves the register count of v1)
        auto tile_body_staged = [&](auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            const int e0 = threadIdx.x * K3_L;
            const int valid = FULL ? K3_L : k3_valid(P.n, n0, K3_L);
            MP ex;
            if (P.thread_ex) {  // uniform: the summary pass kept every segment's prefix
                const double2 v = __ldcs(reinterpret_cast<const double2 *>(P.thread_ex) + t * K3_THREADS + threadIdx.x);
                ex = MP{v.x, v.y};
            } else {
                MP f = mp_id();
#pragma unroll
                for (int j = 0; j < K3_L; ++j) {
                    const double xj = __dsub_rn(smS[k3_pad(e0 + j)], smA[k3_pad(e0 + j + 1)]);
                    if (FULL || j < valid) f = mp_step(f, xj);
                }
                MP total;
                ex = cta_scan<K3_THREADS>(f, smem, &total);
            }
            const double w_tile = mp_apply(P.tile_ex[t], P.w_in);
            double w = mp_apply(ex, w_tile);
            const int last_j = (!FULL && P.n - 1 >= n0 && P.n - 1 < n0 + K3_L) ? (int)(P.n - 1 - n0) : -1;
            double *Wp = P.W ? P.W + n0 : nullptr;
            double an = smA[k3_pad(e0)];
#pragma unroll
            for (int j = 0; j < K3_L; ++j) {
                const long long n = n0 + j;
                const double sj = smS[k3_pad(e0 + j)];
                const double a0 = an;
                an = smA[k3_pad(e0 + j + 1)];
                if (FULL || j < valid) {
                    if (Wp) Wp[j] = w;
                    const double v = __dadd_rn(w, sj);
                    const double w2 = w * w, v2 = v * v;
                    acc[0] += w;
                    acc[1] += w2;
                    acc[2] += w2 * w;
                    acc[3] += w2 * w2;
                    acc[4] += v;
                    acc[5] += v2;
                    acc[6] += v2 * v;
                    acc[7] += v2 * v2;
                    acc[8] += a0;
                    acc[9] += sj;
                    if (WANT_P) level_walk<MODE>(P, v, e0 + j, n, smA, lev, over);
                    const double raw = __dadd_rn(w, __dsub_rn(sj, an));
                    if (!FULL && j == last_j) {
                        P.tail[0] = raw;
                        P.tail[1] = v;
                    }
                    w = max0(raw);
                }
            }
        };
        auto tile_body = [&](auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            acc[8] += load_segment<REPLAY, WANT_P && !REPLAY, false, FULL>(P, n0, x, s, smS, smA);
            if (WANT_P && !REPLAY) {
                // halo of generated gaps beyond the tile
                const int e = K3_TILE + 1 + threadIdx.x;
                if (e <= K3_TILE + K3_HALO) smA[k3_pad(e)] = gen_gap(P, t * K3_TILE + e);
            }
            MP ex;
            if (P.thread_ex) {
                const double2 v = __ldcs(reinterpret_cast<const double2 *>(P.thread_ex) + t * K3_THREADS + threadIdx.x);
                ex = MP{v.x, v.y};
                if (REPLAY || WANT_P) __syncthreads();  // (what the scan's barriers did) the staged tile has been read
            } else {
                MP total;
                ex = cta_scan<K3_THREADS>(fold_segment<FULL>(P, n0, x), smem, &total);
            }
            const double w_tile = mp_apply(P.tile_ex[t], P.w_in);
            double w = mp_apply(ex, w_tile);
            const int valid = FULL ? K3_L : k3_valid(P.n, n0, K3_L);
            const int last_j = (!FULL && P.n - 1 >= n0 && P.n - 1 < n0 + K3_L) ? (int)(P.n - 1 - n0) : -1;
            double *Wp = P.W ? P.W + n0 : nullptr;
#pragma unroll
            for (int j = 0; j < K3_L; ++j) {
                const long long n = n0 + j;
                if (FULL || j < valid) {
                    if (Wp) Wp[j] = w;
                    const double v = __dadd_rn(w, s[j]);
                    const double w2 = w * w, v2 = v * v;
                    acc[0] += w;
                    acc[1] += w2;
                    acc[2] += w2 * w;
                    acc[3] += w2 * w2;
                    acc[4] += v;
                    acc[5] += v2;
                    acc[6] += v2 * v;
                    acc[7] += v2 * v2;
                    acc[9] += s[j];
                    if (WANT_P) level_walk<MODE>(P, v, threadIdx.x * K3_L + j, n, smA, lev, over);
                    const double raw = __dadd_rn(w, x[j]);
                    if (!FULL && j == last_j) {
                        P.tail[0] = raw;
                        P.tail[1] = v;
                    }
                    w = max0(raw);
                }
            }
        };
        if (MODE == 2) {  // measured: the stash walk gains 7 % (3 CTAs/SM), the replay walk loses 2 % with it
            if (full_tile)
                tile_body_staged(K3True{});
            else
                tile_body_staged(K3False{});
        } else {
            if (full_tile)
                tile_body(K3True{});
            else
                tile_body(K3False{});
        }
        if (WANT_P) {
            if (++since_flush == 16) {
                since_flush = 0;
                flush_levels();
            }
            __syncthreads();  // every level walk has read the staged gaps before the next tile overwrites them
        } else if (MODE == 2) {
            // the staged body reads its operands from shared memory DURING the walk (the register form has read
            // them before the scan's barriers): every warp must be done before the next tile is staged
            __syncthreads();
        }
    }
    if (WANT_P) {
        flush_levels();
        __syncthreads();
        if (threadIdx.x <= K3_PBL) {
            double v = 0.0;
            for (int wv = 0; wv < K3_THREADS / 32; ++wv) v += lev_w[wv * (K3_PBL + 1) + threadIdx.x];
            P.level_partial[(long long)blockIdx.x * (K3_PBL + 1) + threadIdx.x] = v;
        }
    }
    // fixed-order CTA reduction of the partial sums
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < K3_NSTAT; ++i) {
        double v = acc[i];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < K3_NSTAT) {
        double v = 0.0;
        for (int wv = 0; wv < K3_THREADS / 32; ++wv) v += red[wv][threadIdx.x];
        P.partial[(long long)blockIdx.x * K3_NSTAT + threadIdx.x] = v;
    }
}

__global__ void __launch_bounds__(32) k3_final(const double *partial, const double *level_partial, int grid, double *stats)
{
    // one warp per statistic, fixed association: lane l adds the partials of CTAs l, l + 32, ... in order, then the
    // lanes are folded by the shuffle tree (all loads of a lane are independent: one memory round trip)
    const int i = blockIdx.x, lane = threadIdx.x;
    const double *src = nullptr;
    int stride = 0;
    if (i < K3_NSTAT) {
        src = partial + i;
        stride = K3_NSTAT;
    } else if (level_partial) {
        src = level_partial + (i - K3_NSTAT);
        stride = K3_PBL + 1;
    }
    double v = 0.0;
    if (src)
        for (int b = lane; b < grid; b += 32) v += src[(long long)b * stride];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if (lane == 0) stats[i] = v;
}

// the generator-form summary kernel for a pair of laws: own instantiations for the BASELINE chain (Gamma gaps, Pareto
// services), Gamma gaps with any service law, and M/M; everything else switches on the kind at run time
using SummaryKernel = void (*)(const K3Params);
template <int STAGE>
static SummaryKernel gen_summary_kernel_t(int src_kind, int srv_kind)
{
    if (src_kind == MQ_GAMMA) return srv_kind == MQ_PA ? k3_summary<false, STAGE, MQ_GAMMA, MQ_PA> : k3_summary<false, STAGE, MQ_GAMMA, -1>;
    if (src_kind == MQ_M && srv_kind == MQ_M) return k3_summary<false, STAGE, MQ_M, MQ_M>;
    return k3_summary<false, STAGE, -1, -1>;
}
static SummaryKernel gen_summary_kernel(int stage_mode, int src_kind, int srv_kind)
{
    if (stage_mode == 2) return gen_summary_kernel_t<2>(src_kind, srv_kind);
    return stage_mode == 1 ? gen_summary_kernel_t<1>(src_kind, srv_kind) : gen_summary_kernel_t<0>(src_kind, srv_kind);
}

cudaError_t k3_launch_summary(const K3Params &P, int grid, cudaStream_t st)
{
    if (P.replay) {
        cudaError_t e = cudaFuncSetAttribute(k3_summary<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)K3_XSTAGE_BYTES);
        if (e != cudaSuccess) return e;
        k3_summary<true><<<grid, K3_THREADS, K3_XSTAGE_BYTES, st>>>(P);
    } else {
        gen_summary_kernel(P.spec ? 2 : (P.stG != nullptr ? 1 : 0), P.src.kind, P.srv.kind)<<<grid, K3_THREADS, 0, st>>>(P);
    }
    return cudaGetLastError();
}

int k3_repair_grid(long long ntiles, int sms)
{
    const long long want = (ntiles + K3_THREADS / 32 - 1) / (K3_THREADS / 32);
    return (int)(want < 4ll * sms ? (want < 1 ? 1 : want) : 4ll * sms);
}

cudaError_t k3_launch_repair(const K3Params &P, int grid, cudaStream_t st)
{
    k3_repair<<<grid, K3_THREADS, 0, st>>>(P);
    return cudaGetLastError();
}

cudaError_t k3_launch_top(const K3Params &P, cudaStream_t st)
{
    k3_top_fold<<<K3_TOP / K3_TOP_CTA, K3_TOP_CTA, 0, st>>>(P);
    k3_top_scan<<<1, K3_TOP, 0, st>>>(P);
    k3_top_apply<<<K3_TOP / K3_TOP_CTA, K3_TOP_CTA, 0, st>>>(P);
    return cudaGetLastError();
}

template <int MODE, bool WANT_P>
static cudaError_t launch_walk_one(const K3Params &P, int grid, size_t smem, cudaStream_t st)
{
    if (smem > 0) {
        cudaError_t e = cudaFuncSetAttribute(k3_walk<MODE, WANT_P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k3_walk<MODE, WANT_P><<<grid, K3_THREADS, smem, st>>>(P);
    return cudaGetLastError();
}

cudaError_t k3_launch_walk(const K3Params &P, int grid, cudaStream_t st)
{
    const bool want_p = P.level_partial != nullptr;
    const int mode = P.replay ? 1 : (P.stG ? 2 : 0);
    const size_t smem = k3_walk_smem(mode != 0, want_p);
    if (mode == 1) return want_p ? launch_walk_one<1, true>(P, grid, smem, st) : launch_walk_one<1, false>(P, grid, smem, st);
    if (mode == 2) return want_p ? launch_walk_one<2, true>(P, grid, smem, st) : launch_walk_one<2, false>(P, grid, smem, st);
    return want_p ? launch_walk_one<0, true>(P, grid, smem, st) : launch_walk_one<0, false>(P, grid, smem, st);
}

size_t k3_walk_smem(bool replay, bool want_p)
{
    if (!replay && !want_p) return 0;
    return K3_STAGE_BYTES + (want_p ? K3_LEVEL_BYTES : 0);
}

cudaError_t k3_launch_final(const double *partial, const double *level_partial, int grid, double *stats,
                            cudaStream_t st)
{
    k3_final<<<K3_NSTAT + K3_PBL + 1, 32, 0, st>>>(partial, level_partial, grid, stats);
    return cudaGetLastError();
}

template <int MODE, bool WANT_P>
static cudaError_t occ_walk_one(size_t smem, int *b)
{
    if (smem > 0) {
        cudaError_t e = cudaFuncSetAttribute(k3_walk<MODE, WANT_P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(b, k3_walk<MODE, WANT_P>, K3_THREADS, smem);
}

cudaError_t k3_occupancy(bool replay, int stage_mode, bool want_p, int src_kind, int srv_kind, int *summary_blocks, int *walk_blocks)
{
    int a = 0, b = 0;
    cudaError_t e;
    const bool stash = stage_mode == 1;
    const int mode = replay ? 1 : (stash ? 2 : 0);
    const size_t wsm = k3_walk_smem(mode != 0, want_p);
    if (replay) {
        e = cudaFuncSetAttribute(k3_summary<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K3_XSTAGE_BYTES);
        if (e == cudaSuccess)
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k3_summary<true>, K3_THREADS, K3_XSTAGE_BYTES);
    } else {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, gen_summary_kernel(stage_mode, src_kind, srv_kind), K3_THREADS, 0);
    }
    if (e == cudaSuccess) {
        if (mode == 1)
            e = want_p ? occ_walk_one<1, true>(wsm, &b) : occ_walk_one<1, false>(wsm, &b);
        else if (mode == 2)
            e = want_p ? occ_walk_one<2, true>(wsm, &b) : occ_walk_one<2, false>(wsm, &b);
        else
            e = want_p ? occ_walk_one<0, true>(wsm, &b) : occ_walk_one<0, false>(wsm, &b);
    }
    *summary_blocks = a;
    *walk_blocks = b;
    return e;
}

}  // namespace mq
