// k3_onepass.cu -- K3 replay form with ONE read of A and S from HBM: a persistent, warp-specialised scan that WALKS EVERY
// TILE SPECULATIVELY while it is in shared memory and repairs the few jobs that depended on the carry afterwards.
//
// Same specification as k3_lindley.cu (and oracle/mq_oracle_scan.c): the max-plus Lindley recursion of QsSim.run() with
// one channel (most_queue/sim/base.py:488-528, wait sample base.py:302, sojourn base.py:272) evaluated with the FIXED
// association tree "16-job segments folded sequentially -> Kogge-Stone over the 32 lanes of a warp -> sequential over the
// 8 warps of a 4096-job tile -> tile level: chunks of ceil(ntiles/1024) tiles folded sequentially, Kogge-Stone over 32
// chunks, sequential over the 32 groups".  Every per-job wait, range summary and tail this kernel produces is the value
// the three-launch path produces, bit for bit; the power sums agree to rounding (another order of addition).
//
// The idea that removes the second read.  The wait at the start of segment s of a tile is max(a_s, w_tile + b_s), where
// (a_s, b_s) is the segment's exclusive prefix INSIDE the tile and w_tile the wait carried into the tile.  a_s is known as
// soon as the tile itself has been scanned; w_tile needs the look-back over all earlier tiles (microseconds).  But a
// queue forgets: after the first idle period inside the tile the carry no longer matters, and then max(.) == a_s exactly.
// At rho = 0.7 that holds for all but 1.2 of the 256 segments of a tile on average (6 at rho = 0.9, 24 at rho = 0.95).
// So the compute warps walk every segment from a_s at once, while the tile is still in shared memory, and a look-back
// warp -- when w_tile finally arrives -- tests max(a_s, w_tile + b_s) != a_s for all 256 segments and re-walks exactly
// those (the first segment from a copy in shared memory, the others from global memory through L1), adding (true -
// provisional) to the sums and
// overwriting their stored waits.  The look-back latency is off the critical path altogether: nothing waits for it but a
// 4 KB ring slot of segment prefixes.
//
// One CTA per SM, co-resident (cooperative launch), static tile order t = blockIdx.x + k * gridDim.x, 24 warps:
//   producer    warp 23 (one lane; it also prefetches the NEXT tile into L2 while it waits for a stage) brings tile k from HBM into one of three 64 KB stages with two TMA tensor copies (S and
//               A as [rows][16] fp64 boxes of 256 x 16 with the 128-byte swizzle: thread r reads ITS row -- its 16-job
//               segment -- with conflict-free LDS.128) + one 16-byte bulk copy for the gap that follows the tile.
//   compute     two groups of 8 warps (tile k goes to group k % 2): fold the rows (x = S - A'), scan the 256 segment
//               summaries, park every segment's exclusive prefix in an L2-resident ring, publish the tile summary, then
//               walk the segment from a_s out of the same stage and hand the stage back.
//   look-back A warps 16-18 (tile k on warp 16 + k % 3) publish what OTHER tiles wait for: the fold of the tile's chunk from
//               the identity up to the tile (needs published summaries only), the exclusive prefix of a chunk (first tile
//               of the chunk: Kogge-Stone over the earlier chunk folds of the group + the group carry) and the carry into
//               the next group.
//   look-back B warps 19-22 (tile k on warp 19 + k % 4) resolve the tile's own exclusive prefix (from the nearest predecessor that has published its
//               inclusive prefix, else from the chunk's exclusive prefix; the sequential association makes every such
//               route give the same bits), publish the inclusive one, and do the repair described above.  The last tile
//               of the range (bounds checks, the tail sample of base.py:300-305) is walked here only.
// Published words carry their own "ready" flag (see publish()): no fences anywhere on the look-back path.
// Every mbarrier that is waited on by parity has the SAME waiter in every phase (see OpShared::full): try_wait.parity can
// only tell the current phase from the preceding one.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda.h>
#include <cudaTypedefs.h>

#include "k3_lindley.cuh"

namespace mq {

namespace {

constexpr int OP_STAGES = 3;                     // 64 KB stages: two being computed on, one in flight
constexpr int OP_GROUPS = 2;                     // compute groups of 8 warps (tile k -> group k % OP_GROUPS)
constexpr int OP_RING = 24;                      // tiles whose repair may be pending (ring of hand-off slots)
constexpr int OP_TILE = K3_THREADS * K3_L;       // 4096 jobs
constexpr int OP_ROWS = K3_THREADS;              // 256 rows of 16 doubles
constexpr int OP_ARR_BYTES = OP_TILE * 8;        // 32 KB per array per stage
constexpr int OP_NLA = 3;                        // look-back A warps (tile k goes to warp k % OP_NLA)
constexpr int OP_NLB = 4;                        // look-back B / repair warps (tile k goes to warp k % OP_NLB)
constexpr int OP_CWARPS = OP_GROUPS * (K3_THREADS / 32);
constexpr int OP_WARP_A = OP_CWARPS;             // 16..18
constexpr int OP_WARP_B = OP_WARP_A + OP_NLA;    // 19..22
constexpr int OP_WARP_P = OP_WARP_B + OP_NLB;    // 23
constexpr int OP_THREADS = (OP_WARP_P + 1) * 32; // 768: 6 warps of 80 registers per scheduler
static_assert(OP_RING % (OP_GROUPS * OP_NLA * OP_NLB) == 0, "every hand-off slot must keep the same waiters from phase to phase");
constexpr int OP_LB = 128;                       // look-back window kept in shared memory (summaries), per look-back warp
constexpr uint32_t OP_TX = 2u * OP_ARR_BYTES + 16u;
#ifndef OP_V_CHECK
#define OP_V_CHECK 0
#endif
#ifndef OP_V_PROGRESS
#define OP_V_PROGRESS 0
#endif
#ifndef OP_V_SPIN_LIMIT
#define OP_V_SPIN_LIMIT 4000000u
#endif
constexpr uint32_t OP_SPIN_LIMIT = OP_V_SPIN_LIMIT;     // wait iterations before a wait gives up (op_give_up): a protocol bug traps with a code instead of hanging

struct OpTrue {
    static constexpr bool value = true;
};
struct OpFalse {
    static constexpr bool value = false;
};

struct OpStage {
    // 1024-byte aligned (the swizzle pattern is a function of the shared-memory address bits 4-9)
    unsigned char S[OP_ARR_BYTES];
    unsigned char A[OP_ARR_BYTES];
};

struct OpShared {
    OpStage st[OP_STAGES];
    double anext[OP_STAGES][2];                   // A[tile_base + 4096], A[tile_base + 4097] of the staged tile
    MP tile_total[OP_RING];                       // the tile's summary (compute group -> look-back warps)
    MP lbA[OP_NLA][OP_LB], lbB[OP_NLB][OP_LB];    // summaries gathered by the look-back warps, newest first
    MP warp_tot[OP_GROUPS][2][K3_THREADS / 32];   // double-buffered by tile: one barrier per tile
    MP ks_left[OP_NLA][5];                        // left operands of the late Kogge-Stone lane (look-back A)
    double seg0[OP_RING][2 * K3_L + 2];           // S[0..15], A[0..16] of the tile's first segment (always repaired)
    double red[OP_WARP_P][K3_NSTAT];
    // full[s][g]: tile k lands in stage k % 3 for group k % 2 -- one "full" barrier per (stage, group), so that a barrier has
    // ONE waiting group, which has seen every earlier phase itself.  With one barrier per stage the groups alternate on
    // it, and a group that arrives while the OTHER group's tile is still in flight asks for the parity of a phase two
    // ahead of the current one: try_wait.parity takes that for the preceding (completed) phase and lets it through onto
    // stale data (measured: 1 launch in 100 with wrong sums, 1 in 2000 hung).
    unsigned long long full[OP_STAGES][OP_GROUPS], empty[OP_STAGES];
    unsigned long long folded[OP_RING], walked[OP_RING], slot_free[OP_RING];
};

// The two maxima of the recursion without fmax's NaN handling (no operand is ever a NaN: sums of finite draws and
// -inf): every wait of this kernel sits on a chain of dependent "add, max" steps, and fmax(double) is a seven-instruction
// sequence (DSETP.MAX, FSEL, LOP3, SEL, SEL, two moves) where these are a shift and two ANDs, or DSETP.GT and two selects.
// Same values bit for bit (a -0.0 result would need a -0.0 operand, which max(0, .) never produces).
__device__ __forceinline__ double max0(double x)
{
    const int hi = __double2hiint(x);
    const int keep = ~(hi >> 31);
    return __hiloint2double(hi & keep, __double2loint(x) & keep);
}
__device__ __forceinline__ double max2(double a, double s) { return s > a ? s : a; }
__device__ __forceinline__ MP mp_id() { return MP{-INFINITY, 0.0}; }
__device__ __forceinline__ MP mp_then(MP earlier, MP later)
{
    return MP{max2(later.a, __dadd_rn(earlier.a, later.b)), __dadd_rn(earlier.b, later.b)};
}
__device__ __forceinline__ MP mp_step(MP f, double x)
{
    return MP{max0(__dadd_rn(f.a, x)), __dadd_rn(f.b, x)};
}
__device__ __forceinline__ double mp_apply(MP f, double w) { return max2(f.a, __dadd_rn(w, f.b)); }
__device__ __forceinline__ MP mp_shfl(MP v, int src)
{
    return MP{__shfl_sync(0xffffffffu, v.a, src), __shfl_sync(0xffffffffu, v.b, src)};
}
__device__ __forceinline__ MP mp_shfl_up(MP v, int d)
{
    return MP{__shfl_up_sync(0xffffffffu, v.a, d), __shfl_up_sync(0xffffffffu, v.b, d)};
}
// Kogge-Stone inclusive scan over the 32 lanes (the association of cta_scan in k3_lindley.cu)
__device__ __forceinline__ MP warp_inclusive(MP v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        MP up = mp_shfl_up(v, d);
        if (lane >= d) v = mp_then(up, v);
    }
    return v;
}

// ---- mbarrier / TMA primitives (PTX) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// arrive that may only be issued once `token` exists.  An mbarrier arrive does NOT wait for the data of earlier
// shared-memory LOADS of the thread to have come back (measured: a stage handed back right after its LDS were issued was
// overwritten by the next TMA copy before the loads had read it).  The caller folds one word of every loaded register
// into `token`; the instructions that do so cannot issue before the loads have returned.
// `zero_mask` is a kernel parameter that is 0 at run time: neither the compiler nor ptxas can fold token & zero_mask, so
// the address of the arrive really depends on every loaded register that went into the token.
__device__ __forceinline__ void mbar_arrive_after(unsigned long long *b, uint32_t token, uint32_t zero_mask)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b) + (token & zero_mask)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes or ~the hint has passed, so a
// waiting warp costs a handful of issue slots per microsecond instead of spinning (a spin loop with a clock read per
// iteration took more than half of all issue slots of the SM in the first version of this kernel)
__device__ __forceinline__ bool mbar_try(unsigned long long *b, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(b)), "r"(parity), "r"(20000u)
        : "memory");
    return ok != 0;
}
// A wait that ran out of patience (seconds: a protocol bug) leaves a code in MAPPED HOST memory and traps: the launch
// fails instead of hanging the GPU, and the host can still read the code after the context is gone.  The reporter does
// not return -- with a bail-out path instead of the trap the same kernel ran 6 % slower (the compiler keeps the state
// of the exit alive across every wait).
__device__ __noinline__ void op_give_up(int *err, int code)
{
    *(volatile int *)err = code;
    __threadfence_system();
    __trap();
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, uint32_t parity, int *err)
{
    for (uint32_t it = 0; !mbar_try(b, parity); ++it) {
        if (it > OP_SPIN_LIMIT) op_give_up(err, 0x100 | (int)(smem_u32(b) & 0xffff) << 12 | (int)parity);
    }
}
// the waits of the look-back warps (same loop, a shorter patience: they trap first and name the barrier that starved them)
__device__ __forceinline__ void mbar_wait_idle(unsigned long long *b, uint32_t parity, int *err)
{
    for (uint32_t it = 0; !mbar_try(b, parity); ++it) {
        if (it > OP_SPIN_LIMIT / 16) op_give_up(err, 0x200 | (int)(smem_u32(b) & 0xffff) << 12 | (int)parity);
    }
}
__device__ __forceinline__ void tma_tile_2d(void *dst, const CUtensorMap *map, int c0, int c1, unsigned long long *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
// the same copy with an L2 eviction-priority hint (createpolicy): a tile is read once, so it is fetched evict_first
__device__ __forceinline__ void tma_tile_2d_hint(void *dst, const CUtensorMap *map, int c0, int c1, unsigned long long *bar,
                                                 unsigned long long policy)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
// L2 prefetch of a tile box: the HBM latency of a tile is paid while its stage is still busy with an earlier tile, and
// the TMA copy that follows finds the lines in L2
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *map, int c0, int c1)
{
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ unsigned long long policy_evict_first()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_copy_16(void *dst, const void *src, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ long long gtime()
{
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void named_bar(int id, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---- publication through L2 without fences: a published (a, b) pair is ONE 16-byte word; "not yet" is the all-ones
// pattern (a NaN in b, which a sum of finite draws never is) that the launch memsets the arrays to.  128-bit relaxed
// loads / stores at GPU scope are single-copy atomic (PTX b128), so the word itself is the flag: no release fence on the
// writer, no acquire (and no L1 invalidation) on the reader, one L2 round trip per hop. ----
__device__ __forceinline__ void publish(MP *p, MP v)
{
    asm volatile(
        "{\n\t.reg .b128 q;\n\t"
        "mov.b128 q, {%1, %2};\n\t"
        "st.relaxed.gpu.global.b128 [%0], q;\n\t}"
        ::"l"(p), "l"(__double_as_longlong(v.a)), "l"(__double_as_longlong(v.b))
        : "memory");
}
__device__ __forceinline__ bool try_mp(const MP *p, MP &v)
{
    long long lo, hi;
    asm volatile(
        "{\n\t.reg .b128 q;\n\t"
        "ld.relaxed.gpu.global.b128 q, [%2];\n\t"
        "mov.b128 {%0, %1}, q;\n\t}"
        : "=l"(lo), "=l"(hi)
        : "l"(p)
        : "memory");
    v = MP{__longlong_as_double(lo), __longlong_as_double(hi)};
    return hi != -1ll;
}
__device__ __forceinline__ MP wait_mp(const MP *p, int *err)
{
    MP v;
    if (try_mp(p, v)) return v;
    for (uint32_t it = 0;; ++it) {
        __nanosleep(it < 8 ? 32 : 128);
        if (try_mp(p, v)) return v;
        if (it > OP_SPIN_LIMIT) op_give_up(err, 0x300);
    }
}

// element (row r, column col) of a swizzled [256][16] fp64 box: 16-byte chunk q = col / 2 sits at chunk q ^ (r & 7)
__device__ __forceinline__ const double2 *sw_chunk(const unsigned char *box, int r, int q)
{
    return reinterpret_cast<const double2 *>(box + r * 128 + ((q ^ (r & 7)) << 4));
}
__device__ __forceinline__ double *sw_elem(unsigned char *box, int e)
{
    const int r = e >> 4, col = e & 15;
    return reinterpret_cast<double *>(box + r * 128 + (((col >> 1) ^ (r & 7)) << 4) + ((col & 1) << 3));
}

// half h (8 jobs) of row r of a staged tile: S[16 r + 8 h ..] and the 9 gaps A[16 r + 8 h .. 16 r + 8 h + 8]
constexpr int OP_H = K3_L / 2;
__device__ __forceinline__ void load_half(const OpStage &st, const double *a_next, int r, int h, double (&sv)[OP_H],
                                          double (&av)[OP_H + 1])
{
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double2 x = *sw_chunk(st.S, r, 4 * h + q);
        const double2 y = *sw_chunk(st.A, r, 4 * h + q);
        sv[2 * q] = x.x;
        sv[2 * q + 1] = x.y;
        av[2 * q] = y.x;
        av[2 * q + 1] = y.y;
    }
    if (h == 0)
        av[OP_H] = sw_chunk(st.A, r, 4)->x;
    else
        av[OP_H] = r + 1 < OP_ROWS ? sw_chunk(st.A, r + 1, 0)->x : a_next[0];
}

// the last tile of the range is staged with plain loads: services beyond the range and gaps beyond the chain are 0
__device__ __forceinline__ void stage_ragged(const K3Params &P, long long tb, OpStage &st, double *a_next, int r)
{
    const int vs = k3_valid_count(P.n, tb, OP_TILE);
    const int va = k3_valid_count(P.n + P.n_ahead + 1, tb, OP_TILE + 2);
    for (int e = r; e < OP_TILE; e += K3_THREADS) {
        *sw_elem(st.S, e) = e < vs ? __ldcs(P.S + tb + e) : 0.0;
        *sw_elem(st.A, e) = e < va ? __ldcs(P.A + tb + e) : 0.0;
    }
    if (r < 2) a_next[r] = OP_TILE + r < va ? __ldcs(P.A + tb + OP_TILE + r) : 0.0;
}

}  // namespace

struct K3OnePassParams {
    K3Params P;            // n, A, S, w_in, partial, tail, total, W
    CUtensorMap mapS, mapA;
    // published words, all-ones until written (see publish()):
    MP *t_sum;             // [ntiles] tile summaries
    MP *t_incf;            // [ntiles] fold of the tile's chunk from the identity up to and including the tile
    MP *t_inc;             // [ntiles] inclusive prefix of the tile
    MP *chunk_ex;          // [K3_TOP] exclusive prefix of a chunk (published when chunks have more than one tile)
    MP *carry;             // [40]     carry into group g
    MP *seg_ring;          // [grid][OP_RING][256] exclusive prefix of every segment inside its tile (compute -> repair)
    long long chunk;       // tiles per chunk
    uint32_t zero_mask;    // 0 (see mbar_arrive_after)
    int hints;             // 1: the tile fetch carries the L2 evict_first hint (a tile is read once; MQ_SCAN_HINTS=0 turns it off)
    int prefetch;          // tiles the producer prefetches into L2 ahead of the stage it is waiting for (0: none)
    int ring;              // hand-off slots in use (OP_NLB..OP_RING): tiles whose look-back / repair may be pending
    int *err;
    long long *trace;      // optional [grid][OP_TRACE_TILES][24] SM clocks of the pipeline events (MQ_SCAN_TRACE)
};

#if OP_V_PROGRESS
#define OP_PROG(role, k_, code) (((volatile int *)Q.err)[16 + blockIdx.x * 16 + (role)] = (k_) * 16 + (code))
#else
#define OP_PROG(role, k_, code) ((void)0)
#endif
constexpr int OP_TRACE_TILES = 64;
#define OP_TRACE(ev, val)                                                                             \
    do {                                                                                              \
        if (Q.trace && k < OP_TRACE_TILES) Q.trace[((long long)blockIdx.x * OP_TRACE_TILES + k) * 24 + (ev)] = (val); \
    } while (0)

// f o lb[n-1] o ... o lb[0] (lb holds the newest summary first)
__device__ __forceinline__ MP fold_newest_first(MP f, const MP *lb, int n)
{
#pragma unroll 1
    for (int i = n - 1; i >= 0; --i) f = mp_then(f, lb[i]);
    return f;
}

// Backward scan over the predecessors of tile t inside its chunk, 32 at a time (lane i looks at tile t - 1 - cnt - i).
// `ready` is the array of published partial folds (t_incf or t_inc): the nearest predecessor that has one becomes the base
// (idx = number of summaries to fold after it, newest first in lb[]); every nearer predecessor contributes its summary,
// which the lane waits for.  idx stays -1 when the chunk start is reached first (all n_pred summaries are in lb[]).
__device__ __forceinline__ void gather(const K3OnePassParams &Q, MP *lb, const MP *ready, long long t, long long n_pred,
                                       int lane, long long &idx, MP &base)
{
    long long cnt = 0;
    while (idx < 0 && cnt < n_pred) {
        const long long i = cnt + lane;
        const bool ok = i < n_pred;
        const long long p = t - 1 - i;
        MP v = mp_id(), sum = mp_id();
        bool have = false, have_sum = false;
        if (ok) {
            have = try_mp(ready + p, v);          // both loads are in flight together: one L2 round trip
            have_sum = try_mp(Q.t_sum + p, sum);
            if (!have && i == OP_LB - 1) {  // the window is OP_LB summaries long: its last slot waits for a partial fold
                v = wait_mp(ready + p, Q.err);
                have = true;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, have);
        const int j = m ? __ffs(m) - 1 : 32;
        if (ok && lane < j) lb[i] = have_sum ? sum : wait_mp(Q.t_sum + p, Q.err);
        if (m) {
            base = mp_shfl(v, j);
            idx = cnt + j;
        }
        cnt += 32;
    }
    __syncwarp();
}


// One segment walked by one lane of a look-back warp: the repair of a segment whose start wait depends on the carry into
// the tile (FULL: adds true - provisional to acc), or the only walk of a segment of the last tile of the range (bounds,
// tail sample; nothing provisional was accumulated).  Sp[j] = S[n0 + j], Ap[j] = A[n0 + 1 + j]: the shared-memory copy
// of the tile's first segment, or global memory (prefetched into L1 by the caller: the loop then runs at L1 latency).
template <bool FULL>
__device__ __forceinline__ void repair_segment(const K3Params &P, long long n0, const double *Sp, const double *Ap,
                                               double w_true, double w_prov, bool prov, double (&acc)[8])
{
    double *Wp = P.W ? P.W + n0 : nullptr;
    const int valid = FULL ? K3_L : k3_valid_count(P.n, n0, K3_L);
    const int last_j = (!FULL && P.n - 1 >= n0 && P.n - 1 < n0 + K3_L) ? (int)(P.n - 1 - n0) : -1;
    double w = w_true, wp = w_prov;
#pragma unroll 4
    for (int j = 0; j < K3_L; ++j) {
        if (FULL || j < valid) {
            const double s = Sp[j], a1 = Ap[j];
            if (Wp) Wp[j] = w;
            const double x = __dsub_rn(s, a1);
            const double v = __dadd_rn(w, s);
            const double w2 = w * w, v2 = v * v;
            double t[8] = {w, w2, w2 * w, w2 * w2, v, v2, v2 * v, v2 * v2};
            if (FULL && prov) {
                const double vp = __dadd_rn(wp, s);
                const double p2 = wp * wp, q2 = vp * vp;
                t[0] -= wp;
                t[1] -= p2;
                t[2] -= p2 * wp;
                t[3] -= p2 * p2;
                t[4] -= vp;
                t[5] -= q2;
                t[6] -= q2 * vp;
                t[7] -= q2 * q2;
                wp = max0(__dadd_rn(wp, x));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += t[i];
            const double raw = __dadd_rn(w, x);
            if (!FULL && j == last_j) {
                P.tail[0] = raw;
                P.tail[1] = v;
            }
            w = max0(raw);
        }
    }
}
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__global__ void __maxnreg__(80) k3_onepass(const __grid_constant__ K3OnePassParams Q)
{
    extern __shared__ __align__(1024) unsigned char op_raw[];
    OpShared &sm = *reinterpret_cast<OpShared *>(op_raw);
    if (threadIdx.x == 0 && (smem_u32(op_raw) & 1023u)) op_give_up(Q.err, 3);  // the swizzle pattern assumes 1024-byte aligned boxes
    const K3Params &P = Q.P;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long ntiles = P.ntiles;
    const int RING = Q.ring;
    const int my_tiles = blockIdx.x < ntiles ? (int)((ntiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;
    MP *ring = Q.seg_ring + (long long)blockIdx.x * OP_RING * K3_THREADS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < OP_STAGES; ++s) {
            for (int g = 0; g < OP_GROUPS; ++g) mbar_init(&sm.full[s][g], 1);
            mbar_init(&sm.empty[s], K3_THREADS);
        }
        for (int d = 0; d < OP_RING; ++d) {
            mbar_init(&sm.folded[d], K3_THREADS);
            mbar_init(&sm.walked[d], K3_THREADS);
            mbar_init(&sm.slot_free[d], 2);  // both look-back warps of the tile are done with the slot
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == OP_WARP_P) {
        // ------------------------------------------------------------------ producer (HBM -> shared memory)
        if (lane == 0) {
            const unsigned long long once = policy_evict_first();
            const int PF = Q.prefetch;
            int next_pf = OP_STAGES;  // tiles 0..2 go straight into the stages
            for (int k = 0; k < my_tiles; ++k) {
                const int s = k % OP_STAGES;
                const long long t = blockIdx.x + (long long)k * gridDim.x;
                for (; PF > 0 && next_pf <= k + PF && next_pf < my_tiles; ++next_pf) {
                    const long long tp = blockIdx.x + (long long)next_pf * gridDim.x;
                    if (next_pf > k && (tp + 1) * (long long)OP_TILE < P.n) {
                        tma_prefetch_2d(&Q.mapS, 0, (int)(tp * OP_ROWS));
                        tma_prefetch_2d(&Q.mapA, 0, (int)(tp * OP_ROWS));
                    }
                }
                OP_PROG(15, k, 1);
                mbar_wait(&sm.empty[s], (uint32_t)((k / OP_STAGES) & 1) ^ 1u, Q.err);
                OP_PROG(15, k, 2);
                OP_TRACE(0, clock64());
                OP_TRACE(21, gtime());
                unsigned long long *fb = &sm.full[s][k % OP_GROUPS];
                if ((t + 1) * (long long)OP_TILE < P.n) {
                    mbar_expect_tx(fb, OP_TX);
                    if (Q.hints) {
                        tma_tile_2d_hint(sm.st[s].S, &Q.mapS, 0, (int)(t * OP_ROWS), fb, once);
                        tma_tile_2d_hint(sm.st[s].A, &Q.mapA, 0, (int)(t * OP_ROWS), fb, once);
                    } else {
                        tma_tile_2d(sm.st[s].S, &Q.mapS, 0, (int)(t * OP_ROWS), fb);
                        tma_tile_2d(sm.st[s].A, &Q.mapA, 0, (int)(t * OP_ROWS), fb);
                    }
                    bulk_copy_16(sm.anext[s], P.A + (t + 1) * OP_TILE, fb);
                } else {
                    mbar_arrive(fb);  // the last tile is staged by the compute group itself (bounds checks)
                }
            }
        }
    } else if (warp < OP_CWARPS) {
        // ------------------------------------------------------------------ compute groups: fold, scan, provisional walk
        const int grp = warp / (K3_THREADS / 32), gw = warp % (K3_THREADS / 32);
        const int r = threadIdx.x % K3_THREADS;
        const int bar_id = 1 + grp;
        double acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.0;
        double sum_a = 0.0, sum_s = 0.0;
        for (int k = grp; k < my_tiles; k += OP_GROUPS) {
            const int s = k % OP_STAGES, d = k % RING;
            const long long t = blockIdx.x + (long long)k * gridDim.x;
            const long long tb = t * OP_TILE;
            OpStage &stg = sm.st[s];
            const double *a_next = sm.anext[s];
            if (lane == 0 && gw < 4) OP_PROG(grp * 4 + gw, k, 1);
            mbar_wait(&sm.slot_free[d], (uint32_t)((k / RING) & 1) ^ 1u, Q.err);  // at most RING tiles await their repair
            if (lane == 0 && gw < 4) OP_PROG(grp * 4 + gw, k, 2);
            mbar_wait(&sm.full[s][grp], (uint32_t)((k / (OP_STAGES * OP_GROUPS)) & 1), Q.err);
            if (lane == 0 && gw < 4) OP_PROG(grp * 4 + gw, k, 3);
            if (r == 0) OP_TRACE(9, clock64());
            const bool full_tile = (t + 1) * (long long)OP_TILE < P.n;
#if OP_V_CHECK
            if (full_tile && (r & 31) == 5) {  // is the tile in the stage the tile we waited for?
                const double got = *sw_elem(stg.S, r * K3_L + 3), want = P.S[tb + r * K3_L + 3];
                if (got != want) {
                    atomicAdd(Q.err + 4, 1);
                    Q.err[6] = k * 1000 + blockIdx.x;
                }
            }
#endif
            int valid = K3_L;
            if (!full_tile) {
                stage_ragged(P, tb, stg, sm.anext[s], r);
                named_bar(bar_id, K3_THREADS);
                valid = k3_valid_count(P.n, tb + (long long)r * K3_L, K3_L);
            }
            // ---- fold the row (tile-uniform fast path: every tile but the last one has no per-job bounds logic)
            MP mine = mp_id();
            auto fold_row = [&](auto full_tag) {
                constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    double sv[OP_H], av[OP_H + 1];
                    load_half(stg, a_next, r, h, sv, av);
#pragma unroll
                    for (int j = 0; j < OP_H; ++j) {
                        if (FULL || h * OP_H + j < valid) {
                            mine = mp_step(mine, __dsub_rn(sv[j], av[j + 1]));
                            sum_a += av[j];
                            sum_s += sv[j];
                        }
                    }
                }
            };
            if (full_tile)
                fold_row(OpTrue{});
            else
                fold_row(OpFalse{});
            // ---- scan of the 256 segment summaries (Kogge-Stone per warp, sequential over the 8 warps)
            const MP inc = warp_inclusive(mine, lane);
            MP within = mp_shfl_up(inc, 1);
            if (lane == 0) within = mp_id();
            MP *wt = sm.warp_tot[grp][(k / OP_GROUPS) & 1];
            if (lane == 31) wt[gw] = inc;
            if (lane == 0 && gw < 4) OP_PROG(grp * 4 + gw, k, 4);
            named_bar(bar_id, K3_THREADS);
            if (lane == 0 && gw < 4) OP_PROG(grp * 4 + gw, k, 5);
            MP carry = mp_id(), tot = mp_id();
#pragma unroll
            for (int g = 0; g < K3_THREADS / 32; ++g) {
                if (g == gw) carry = tot;
                tot = mp_then(tot, wt[g]);
            }
            const MP ex = mp_then(carry, within);
            __stcg(reinterpret_cast<double2 *>(ring + d * K3_THREADS + r), make_double2(ex.a, ex.b));
            if (r == 0) {
                OP_TRACE(2, clock64());
                sm.tile_total[d] = tot;
                publish(Q.t_sum + t, tot);
                OP_TRACE(3, clock64());
                OP_TRACE(16, gtime());
            }
            mbar_arrive(&sm.folded[d]);  // (release, CTA scope) the ring slot and tile_total are visible to the waiters
            // ---- provisional walk from a_s (the first segment has no a_s: it is always repaired; the last tile of the
            // range is walked by the repair only)
            uint32_t token = (uint32_t)__double2hiint(mine.b);
            if (full_tile && r != 0) {
                double w = ex.a;
                double *Wp = P.W ? P.W + tb + (long long)r * K3_L : nullptr;
                auto walk_row = [&](auto w_tag) {
                    constexpr bool STORE_W = decltype(w_tag)::value;
#pragma unroll 1
                    for (int h = 0; h < 2; ++h) {
                        double sv[OP_H], av[OP_H + 1];
                        load_half(stg, a_next, r, h, sv, av);
#pragma unroll
                        for (int jj = 0; jj < OP_H; ++jj) {
                            if (STORE_W) Wp[h * OP_H + jj] = w;
                            const double v = __dadd_rn(w, sv[jj]);
                            const double w2 = w * w, v2 = v * v;
                            acc[0] += w;
                            acc[1] += w2;
                            acc[2] += w2 * w;
                            acc[3] += w2 * w2;
                            acc[4] += v;
                            acc[5] += v2;
                            acc[6] += v2 * v;
                            acc[7] += v2 * v2;
                            w = max0(__dadd_rn(w, __dsub_rn(sv[jj], av[jj + 1])));
                        }
                    }
                };
                if (Wp)
                    walk_row(OpTrue{});
                else
                    walk_row(OpFalse{});
                token |= (uint32_t)__double2hiint(w);  // w depends on every word the walk loaded
            } else if (full_tile) {
                // r == 0: the first segment is walked by the repair alone; leave it a copy of the row
                double *c0 = sm.seg0[d];
#pragma unroll
                for (int j = 0; j < K3_L; ++j) c0[j] = *sw_elem(stg.S, j);
#pragma unroll
                for (int j = 0; j <= K3_L; ++j) c0[K3_L + j] = *sw_elem(stg.A, j);
            }
#if OP_V_CHECK
            if (full_tile && (r & 31) == 7) {  // is it still there at the end of the walk?
                const double got = *sw_elem(stg.A, r * K3_L + 9), want = P.A[tb + r * K3_L + 9];
                if (got != want) {
                    atomicAdd(Q.err + 5, 1);
                    Q.err[7] = k * 1000 + blockIdx.x;
                }
                token |= (uint32_t)__double2hiint(got);
            }
#endif
            // the rows must have been READ before the stage goes back to the TMA producer (see mbar_arrive_after)
            mbar_arrive_after(&sm.empty[s], token, Q.zero_mask);
            mbar_arrive(&sm.walked[d]);
            if (lane == 0 && gw < 4) OP_PROG(grp * 4 + gw, k, 6);
            if (r == 0) {
                OP_TRACE(10, clock64());
                OP_TRACE(22, gtime());
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double v = acc[i];
#pragma unroll
            for (int dd = 16; dd > 0; dd >>= 1) v += __shfl_down_sync(0xffffffffu, v, dd);
            if (lane == 0) sm.red[warp][i] = v;
        }
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) {
            sum_a += __shfl_down_sync(0xffffffffu, sum_a, dd);
            sum_s += __shfl_down_sync(0xffffffffu, sum_s, dd);
        }
        if (lane == 0) {
            sm.red[warp][8] = sum_a;
            sm.red[warp][9] = sum_s;
        }
    } else if (warp >= OP_WARP_A && warp < OP_WARP_A + OP_NLA) {
        // ------------------------------------------------------------------ look-back A: everything other tiles wait for
        // (1) the chunk fold from the identity up to this tile -- depends on published SUMMARIES only; (2) for the first
        // tile of a chunk, the chunk's exclusive prefix: Kogge-Stone over the folds of the earlier chunks of the group
        // (products of step 1 of other tiles) and the group carry; (3) at the end of a group, the carry into the next one.
        // Step 1 of an iteration never waits for a step 2 / 3 of the same or a later iteration of any CTA, so waits do not
        // chain from CTA to CTA.
        const long long chunk = Q.chunk;
        MP *lb = sm.lbA[warp - OP_WARP_A];
        for (int k = warp - OP_WARP_A; k < my_tiles; k += OP_NLA) {
            const int d = k % RING;
            const long long t = blockIdx.x + (long long)k * gridDim.x;
            if (lane == 0) OP_PROG(warp - OP_WARP_A + 8, k, 1);
            mbar_wait_idle(&sm.folded[d], (uint32_t)((k / RING) & 1), Q.err);
            if (lane == 0) OP_PROG(warp - OP_WARP_A + 8, k, 2);
            if (lane == 0) OP_TRACE(4, clock64());
            const MP own = sm.tile_total[d];
            __syncwarp();
            if (lane == 0) mbar_arrive_after(&sm.slot_free[d], (uint32_t)__double2hiint(own.b), Q.zero_mask);  // tile_total[d] has been read
            const long long c = t / chunk, lo = c * chunk, n_pred = t - lo;
            const int g = (int)(c >> 5), l = (int)(c & 31);
            long long idx = -1;
            MP base = mp_id();
            gather(Q, lb, Q.t_incf, t, n_pred, lane, idx, base);
            if (idx < 0) idx = n_pred;  // fold from the identity at the chunk start
            if (lane == 0) OP_PROG(warp - OP_WARP_A + 8, k, 3);
            MP f = base;
            if (lane == 0) {
                OP_TRACE(5, clock64());
                OP_TRACE(11, idx);
                f = fold_newest_first(f, lb, (int)idx);
                f = mp_then(f, own);
                publish(Q.t_incf + t, f);
                OP_TRACE(6, clock64());
                OP_TRACE(17, gtime());
            }
            const long long group_hi = ((long long)(g + 1) * 32 * chunk < ntiles ? (long long)(g + 1) * 32 * chunk : ntiles);
            const bool group_end = t == group_hi - 1;
            if (n_pred == 0 || group_end) {
                f = mp_shfl(f, 0);
                MP cg = mp_id();
                if (g > 0) {
                    if (lane == 0) cg = wait_mp(Q.carry + g, Q.err);  // published a whole group ago (or about to be)
                    cg = mp_shfl(cg, 0);
                }
                if (n_pred == 0) {
                    // The chunk's exclusive prefix = group carry, then the Kogge-Stone inclusive value at lane l - 1 over
                    // the folds of the earlier chunks of the group.  All of those but the LAST (the chunk that ends at
                    // tile t - 1) were published long ago.  Kogge-Stone only looks to the left, so the five left operands
                    // lane l - 1 meets do not depend on its own value: the scan runs first without that lane and records
                    // them; the late fold then needs five dependent steps on one lane instead of five shuffle rounds.
                    const int last = l - 1;
                    MP v = mp_id();
                    if (lane < last) v = wait_mp(Q.t_incf + ((long long)g * 32 + lane + 1) * chunk - 1, Q.err);
                    MP *left = sm.ks_left[warp - OP_WARP_A];
#pragma unroll
                    for (int sft = 0; sft < 5; ++sft) {
                        const MP up = mp_shfl_up(v, 1 << sft);
                        if (lane == last) left[sft] = up;
                        if (lane >= (1 << sft) && lane != last) v = mp_then(up, v);
                    }
                    if (l == 0) {
                        if (lane == 0) publish(Q.chunk_ex + c, mp_then(cg, mp_id()));
                    } else if (lane == last) {
                        MP x = wait_mp(Q.t_incf + t - 1, Q.err);
#pragma unroll
                        for (int sft = 0; sft < 5; ++sft)
                            if (last >= (1 << sft)) x = mp_then(left[sft], x);
                        publish(Q.chunk_ex + c, mp_then(cg, x));
                    }
                    if (lane == 0) OP_TRACE(18, gtime());
                }
                if (group_end) {
                    // the group total is the Kogge-Stone value at lane 31 (a different association from lane l's when the
                    // range ends inside the group, so the full scan it is; one tile in 32 chunks comes here)
                    MP v = mp_id();
                    if (lane < l)
                        v = wait_mp(Q.t_incf + ((long long)g * 32 + lane + 1) * chunk - 1, Q.err);
                    else if (lane == l)
                        v = f;
                    const MP total_g = mp_shfl(warp_inclusive(v, lane), 31);
                    if (lane == 0) {
                        const MP next = mp_then(cg, total_g);
                        if (t == ntiles - 1)
                            *P.total = next;  // the groups after the last tile hold the identity
                        else
                            publish(Q.carry + g + 1, next);
                    }
                }
            }
            __syncwarp();
        }
    } else if (warp >= OP_WARP_B && warp < OP_WARP_B + OP_NLB) {
        // ------------------------------------------------------------------ look-back B: the tile's exclusive prefix, then
        // the repair of the segments whose start wait depends on it
        const long long chunk = Q.chunk;
        MP *lb = sm.lbB[warp - OP_WARP_B];
        double acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.0;
        for (int k = warp - OP_WARP_B; k < my_tiles; k += OP_NLB) {
            const int d = k % RING;
            const long long t = blockIdx.x + (long long)k * gridDim.x;
            const long long tb = t * OP_TILE;
            if (lane == 0) OP_PROG(warp - OP_WARP_B + 11, k, 1);
            mbar_wait_idle(&sm.folded[d], (uint32_t)((k / RING) & 1), Q.err);
            if (lane == 0) OP_PROG(warp - OP_WARP_B + 11, k, 2);
            const MP own = sm.tile_total[d];
            const long long c = t / chunk, n_pred = t - c * chunk;
            long long idx = -1;
            MP base = mp_id();
            gather(Q, lb, Q.t_inc, t, n_pred, lane, idx, base);
            if (lane == 0) {
                OP_TRACE(12, idx);
                OP_TRACE(20, gtime());
            }
            if (idx < 0) {
                // reached the chunk start: continue from the chunk's exclusive prefix (look-back A of its first tile)
                idx = n_pred;
                if (lane == 0) base = wait_mp(Q.chunk_ex + c, Q.err);
                base = mp_shfl(base, 0);
            }
            MP ex = mp_id();
            if (lane == 0) {
                ex = fold_newest_first(base, lb, (int)idx);
                OP_TRACE(7, clock64());
                OP_TRACE(19, gtime());
                publish(Q.t_inc + t, mp_then(ex, own));
                OP_TRACE(8, clock64());
            }
            ex = mp_shfl(ex, 0);
            const double w_tile = mp_apply(ex, P.w_in);
            // ---- repair.  The provisional walk (and its wait stores, which the repair overwrites) must be over.
            if (lane == 0) OP_PROG(warp - OP_WARP_B + 11, k, 4);
            mbar_wait_idle(&sm.walked[d], (uint32_t)((k / RING) & 1), Q.err);
            if (lane == 0) OP_PROG(warp - OP_WARP_B + 11, k, 5);
            const bool full_tile = (t + 1) * (long long)OP_TILE < P.n;
            double2 sx[K3_THREADS / 32];
#pragma unroll
            for (int q = 0; q < K3_THREADS / 32; ++q)
                sx[q] = __ldcg(reinterpret_cast<const double2 *>(ring + d * K3_THREADS + q * 32 + lane));
            int repaired = 0;
#pragma unroll 1
            for (int q = 0; q < K3_THREADS / 32; ++q) {
                const int sg = q * 32 + lane;
                const double w0 = mp_apply(MP{sx[q].x, sx[q].y}, w_tile);
                const long long n0 = tb + (long long)sg * K3_L;
                const double *Sp = P.S + n0, *Ap = P.A + n0 + 1;
                if (full_tile) {
                    if (sg == 0) {
                        repair_segment<true>(P, n0, sm.seg0[d], sm.seg0[d] + K3_L + 1, w0, 0.0, false, acc);
                        ++repaired;
                    } else if (w0 != sx[q].x) {
                        prefetch_l1(Sp);
                        prefetch_l1(Sp + K3_L - 1);
                        prefetch_l1(Ap);
                        prefetch_l1(Ap + K3_L - 1);
                        repair_segment<true>(P, n0, Sp, Ap, w0, sx[q].x, true, acc);
                        ++repaired;
                    }
                } else {
                    repair_segment<false>(P, n0, Sp, Ap, w0, 0.0, false, acc);
                }
            }
            if (Q.trace) {
                repaired = __reduce_add_sync(0xffffffffu, repaired);
                if (lane == 0) OP_TRACE(13, repaired);
            }
            __syncwarp();
            if (lane == 0) {
                OP_TRACE(14, clock64());
                mbar_arrive_after(&sm.slot_free[d], (uint32_t)__double2hiint(sx[0].y), Q.zero_mask);
                OP_PROG(warp - OP_WARP_B + 11, k, 6);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double v = acc[i];
#pragma unroll
            for (int dd = 16; dd > 0; dd >>= 1) v += __shfl_down_sync(0xffffffffu, v, dd);
            if (lane == 0) sm.red[warp][i] = v;
        }
    }
    __syncthreads();
    // fixed-order CTA reduction: statistics 0-7 from the compute warps and the repairs, 8-9 from the compute warps
    if (threadIdx.x < K3_NSTAT) {
        const int i = threadIdx.x;
        double v = 0.0;
        if (i < 8) {
            for (int wv = 0; wv < OP_CWARPS; ++wv) v += sm.red[wv][i];
            for (int wv = OP_WARP_B; wv < OP_WARP_B + OP_NLB; ++wv) v += sm.red[wv][i];
        } else if (i < 10) {
            for (int wv = 0; wv < OP_CWARPS; ++wv) v += sm.red[wv][i];
        }
        P.partial[(long long)blockIdx.x * K3_NSTAT + i] = v;
    }
}

size_t k3_onepass_smem() { return sizeof(OpShared); }

size_t k3_onepass_scratch_bytes(long long ntiles)
{
    // t_sum, t_incf, t_inc [ntiles], chunk_ex [K3_TOP], carry [40] (all memset to ones before every launch), err, and the
    // hand-off ring of segment prefixes (at most 256 CTAs)
    return sizeof(MP) * (3 * (size_t)ntiles + K3_TOP + 40) + 256 + sizeof(MP) * (size_t)256 * OP_RING * K3_THREADS;
}

static PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder()
{
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

// rows of 16 doubles, box of 256 rows, 128-byte swizzle
static bool make_map(CUtensorMap *map, const double *base, long long rows)
{
    auto enc = tensor_map_encoder();
    if (!enc || rows < 1) return false;
    cuuint64_t dims[2] = {16, (cuuint64_t)rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {16, (cuuint32_t)OP_ROWS};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool k3_onepass_usable(const K3Params &P)
{
    if (!P.replay || P.level_partial) return false;
    if ((reinterpret_cast<uintptr_t>(P.A) | reinterpret_cast<uintptr_t>(P.S)) & 15) return false;  // TMA: 16-byte aligned
    if (P.ntiles * (long long)OP_ROWS >= (1ll << 31)) return false;                                 // 32-bit box coordinate
    return tensor_map_encoder() != nullptr;
}

// one int of mapped pinned host memory per process: written by op_give_up(), readable by the host even after a trap
static int *error_word(bool device_side)
{
    struct Holder {  // (a function-local static: initialised once, thread-safe)
        int *h = nullptr, *d = nullptr;
        Holder()
        {
            if (cudaHostAlloc((void **)&h, 64 + 16 * 4 * 257, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
                h = nullptr;
                cudaGetLastError();
                return;
            }
            *h = 0;
            if (cudaHostGetDevicePointer((void **)&d, h, 0) != cudaSuccess) d = h;  // unified addressing: same pointer
        }
    };
    static Holder holder;
    return device_side ? holder.d : holder.h;
}

// `scratch` = k3_onepass_scratch_bytes(ntiles) bytes of device memory; grid <= resident CTAs (one per SM)
cudaError_t k3_launch_onepass(const K3Params &P, void *scratch, int grid, cudaStream_t st)
{
    const size_t smem = k3_onepass_smem();
    {   // (per device, so on every launch: a process may hold contexts on several GPUs)
        cudaError_t e = cudaFuncSetAttribute(k3_onepass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    K3OnePassParams Q;
    Q.P = P;
    // full tiles only go through TMA: rows that are completely inside the arrays
    if (!make_map(&Q.mapS, P.S, P.n / 16 > 0 ? P.n / 16 : 1) || !make_map(&Q.mapA, P.A, (P.n + 1) / 16 > 0 ? (P.n + 1) / 16 : 1))
        return cudaErrorInvalidValue;
    MP *mp = (MP *)scratch;
    const size_t words = 3 * (size_t)P.ntiles + K3_TOP + 40;
    Q.t_sum = mp;
    Q.t_incf = Q.t_sum + P.ntiles;
    Q.t_inc = Q.t_incf + P.ntiles;
    Q.chunk_ex = Q.t_inc + P.ntiles;
    Q.carry = Q.chunk_ex + K3_TOP;
    Q.err = error_word(true);
    if (!Q.err) return cudaErrorMemoryAllocation;
    Q.seg_ring = (MP *)((char *)(mp + words) + 256);
    if (grid > 256) return cudaErrorInvalidValue;
    Q.chunk = (P.ntiles + K3_TOP - 1) / K3_TOP;
    // hand-off slots: a tile's slot (4 KB of segment prefixes in L2) is held until its look-back and repair are done, so
    // the ring depth x the tile period is the look-back latency the pipeline tolerates; MQ_SCAN_LEAD overrides (3..24)
    Q.ring = 24;
    if (const char *ev = std::getenv("MQ_SCAN_LEAD")) Q.ring = std::atoi(ev);
    // a slot's barriers must have the SAME waiters in every phase (the compute group, the look-back A warp, the look-back
    // B warp of the tile): the ring depth is a multiple of the two groups, the three A warps and the four B warps
    Q.ring = Q.ring >= 24 ? 24 : 12;
    Q.hints = 1;
    Q.zero_mask = 0;
    Q.prefetch = 1;
    if (const char *ev = std::getenv("MQ_SCAN_PREFETCH")) Q.prefetch = std::atoi(ev) < 0 ? 0 : (std::atoi(ev) > 8 ? 8 : std::atoi(ev));
    if (const char *ev = std::getenv("MQ_SCAN_HINTS")) Q.hints = std::atoi(ev) != 0;
    cudaError_t e = cudaMemsetAsync(scratch, 0xFF, sizeof(MP) * words, st);
    if (e != cudaSuccess) return e;
    *error_word(false) = 0;
#if OP_V_PROGRESS
    for (int i = 16; i < 16 + 16 * 257; ++i) error_word(false)[i] = -1;
#endif
    // MQ_SCAN_TRACE=<file>: SM clocks of every pipeline event of the first 64 tiles of each CTA (a development aid)
    Q.trace = nullptr;
    const char *trace_path = std::getenv("MQ_SCAN_TRACE");
    const size_t trace_bytes = sizeof(long long) * (size_t)grid * OP_TRACE_TILES * 24;
    if (trace_path) {
        if (cudaMalloc(&Q.trace, trace_bytes) != cudaSuccess) return cudaErrorMemoryAllocation;
        cudaMemsetAsync(Q.trace, 0, trace_bytes, st);
    }
    void *args[] = {(void *)&Q};
    e = cudaLaunchCooperativeKernel((const void *)k3_onepass, dim3(grid), dim3(OP_THREADS), args, smem, st);
    if (trace_path) {
        std::vector<long long> h((size_t)grid * OP_TRACE_TILES * 24);
        cudaStreamSynchronize(st);
        cudaMemcpy(h.data(), Q.trace, trace_bytes, cudaMemcpyDeviceToHost);
        cudaFree(Q.trace);
        if (FILE *f = std::fopen(trace_path, "wb")) {
            std::fwrite(h.data(), 1, trace_bytes, f);
            std::fclose(f);
        }
    }
    return e;
}

// the code a wait that gave up left behind (0 = none); readable after a failed launch
int k3_onepass_error()
{
    int *h = error_word(false);
#if OP_V_CHECK
    if (h && (h[4] || h[5])) {
        std::fprintf(stderr, "k3_onepass stage check: %d wrong at fold start (last k*1000+cta %d), %d wrong at walk end (last %d)\n", h[4], h[6], h[5], h[7]);
        h[4] = h[5] = 0;
    }
#endif
#if OP_V_PROGRESS
    if (h && *h) {
        std::fprintf(stderr, "k3_onepass progress (tile * 16 + state) per CTA: compute warps 0-7 (group 0: 0-3? see code), A 8-10, B 11-14, producer 15\n");
        for (int b = 0; b < 148; b += 21) {
            std::fprintf(stderr, "  CTA %3d:", b);
            for (int i = 0; i < 16; ++i) std::fprintf(stderr, " %d.%d", h[16 + b * 16 + i] / 16, h[16 + b * 16 + i] % 16);
            std::fprintf(stderr, "\n");
        }
    }
#endif
    return h ? *h : 0;
}

cudaError_t k3_onepass_occupancy(int *blocks)
{
    const size_t smem = k3_onepass_smem();
    cudaError_t e = cudaFuncSetAttribute(k3_onepass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, k3_onepass, OP_THREADS, smem);
}

}  // namespace mq
