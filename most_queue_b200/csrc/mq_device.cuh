// mq_device.cuh -- device-side building blocks shared by the simulation kernels:
// Philox2x32-10, word -> variate samplers (fp32, MUFU log/exp), small helpers.
//
// Generator spec (DESIGN.md "Generator streams"): arrival index i of replication r makes
// one primary call philox2x32_10(c0 = i, c1 = r, key0); word 0 drives the interarrival
// gap before arrival i, word 1 the service time of arrival i.  Distributions that need
// more than one word take them from extension calls with key0 + STRIDE * (2n + which),
// n = 1, 2, ...  The samplers follow most_queue/random/distributions.py (formulas cited at
// each case) but evaluate in fp32 with MUFU.LG2 / MUFU.EX2.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/mqb200.h"

#define MQ_PHILOX_M 0xD256D193u
#define MQ_PHILOX_W 0x9E3779B9u
#define MQ_KEY_STRIDE 0xB5297A4Du

namespace mq {

// fp32 device form of an mq_dist, prepared on the host (see make_distf in mq_api.cu)
struct DistF {
    int kind;
    int r;       // Erlang order; Gamma: 1 when alpha < 1
    uint32_t T;  // H2 / Cox branch threshold on the raw 32-bit word
    float a, b, c, d;
};

__device__ __forceinline__ uint2 philox2x32_10(uint32_t c0, uint32_t c1, uint32_t key)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi = __umulhi(MQ_PHILOX_M, c0);
        uint32_t lo = MQ_PHILOX_M * c0;
        c0 = hi ^ (key + (uint32_t)r * MQ_PHILOX_W) ^ c1;
        c1 = lo;
    }
    return make_uint2(c0, c1);
}

// the same function with the ten round keys key + r * W precomputed by the host (rk lives in the kernel
// parameter block, so each round key is a constant-bank operand of the xor instead of an add per round)
__device__ __forceinline__ uint2 philox2x32_10_rk(uint32_t c0, uint32_t c1, const uint32_t (&rk)[10])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi = __umulhi(MQ_PHILOX_M, c0);
        uint32_t lo = MQ_PHILOX_M * c0;
        c0 = hi ^ rk[r] ^ c1;
        c1 = lo;
    }
    return make_uint2(c0, c1);
}

__device__ __forceinline__ float fast_lg2(float x)
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float fast_ex2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float fast_rcp(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// cos(2 pi u) for u in (0, 1] as -cos(2 pi u - pi): the argument lies in (-pi, pi], where MUFU.COS is good to
// 2^-21 absolute (cospif costs 25 instructions; the oracle evaluates the same formula in fp64)
__device__ __forceinline__ float fast_cos2pi(float u)
{
    return -__cosf(fmaf(u, 6.283185307179586f, -3.141592653589793f));
}

// (w + 0.5) * 2^-32, in (0, 1]
__device__ __forceinline__ float u01(uint32_t w)
{
    return fmaf(__uint2float_rn(w), 0x1p-32f, 0x1p-33f);
}

// extension words of one variate: call n >= 1 of stream `which`
__device__ __forceinline__ uint2 ext_words(uint32_t idx, uint32_t rep, uint32_t key0, int which, uint32_t n)
{
    return philox2x32_10(idx, rep, key0 + MQ_KEY_STRIDE * (2u * n + (uint32_t)which));
}

// -------------------------------------------------------------------------------------
// samplers.  KIND >= 0 compiles exactly one of them; KIND < 0 switches on d.kind at run
// time (uniform across the grid, so the branch does not diverge).
// -------------------------------------------------------------------------------------
__device__ __forceinline__ float sample_m(const DistF &d, uint32_t w0)
{
    // distributions.py:906 -> :807-820 with r = 1: -log(U)/mu ; d.a = -ln2/mu
    return fast_lg2(u01(w0)) * d.a;
}

__device__ __forceinline__ float sample_h2(const DistF &d, uint32_t w0)
{
    // distributions.py:407-429.  One word: the branch test "r < p1" is done on the raw word
    // against T = floor(p1 * 2^32); given the branch, the word is uniform on its side of
    // T and is reused as the exponential's uniform.  d.a = 1/T, d.b = 1/(2^32-T),
    // d.c = -ln2/mu1, d.d = -ln2/mu2
    bool first = w0 < d.T;
    uint32_t off = first ? w0 : w0 - d.T;
    float inv = first ? d.a : d.b;
    float sc = first ? d.c : d.d;
    float u = fmaf(__uint2float_rn(off), inv, 0.5f * inv);
    return fast_lg2(u) * sc;
}

__device__ __forceinline__ float sample_pa(const DistF &d, uint32_t w0)
{
    // distributions.py:749-761: K * U^(-1/alpha); d.a = -1/alpha, d.b = K
    return d.b * fast_ex2(fast_lg2(u01(w0)) * d.a);
}

// The two loop samplers are real functions (one copy per kernel); they take the law's fields BY VALUE: a
// `const DistF &` parameter would force every caller to keep its DistF in local memory.
static __device__ __noinline__ float sample_erlang(float scale, int order, uint32_t w0, uint32_t idx, uint32_t rep,
                                                   uint32_t key0, int which)
{
    // distributions.py:807-820: -log(prod U)/mu, as a sum of logs; scale = -ln2/mu
    float acc = fast_lg2(u01(w0));
    uint2 e = make_uint2(0u, 0u);
    for (int k = 1; k < order; ++k) {
        if (k & 1) e = ext_words(idx, rep, key0, which, (uint32_t)(k + 1) >> 1);
        acc += fast_lg2(u01((k & 1) ? e.x : e.y));
    }
    return acc * scale;
}

static __device__ __noinline__ float sample_gamma(float dd, float cc, float inv_mu, float inv_alpha, int small_alpha,
                                                  uint32_t w0, uint32_t idx, uint32_t rep, uint32_t key0, int which)
{
    // distributions.py:996-1003 (Generator.gamma(alpha, 1/mu)) restated as Marsaglia-Tsang
    // with Box-Muller normals.  dd = d.a, cc = d.b, inv_mu = d.c, inv_alpha = d.d, small_alpha = d.r (alpha < 1)
    float x = dd;
    for (uint32_t t = 0; t < 15; ++t) {
        uint2 e0 = ext_words(idx, rep, key0, which, 2 * t + 1);
        float rad = sqrtf(-1.3862943611198906f * fast_lg2(u01(e0.x)));  // sqrt(-2 ln u)
        float z = rad * fast_cos2pi(u01(e0.y));
        float v = fmaf(cc, z, 1.0f);
        if (v <= 0.0f) continue;
        v = v * v * v;
        x = dd * v;
        // the acceptance uniform of the first attempt is the primary word itself when it is not needed for the
        // alpha < 1 boost: one extension call per draw in the common case instead of two
        const uint32_t wa = (t == 0 && !small_alpha) ? w0 : ext_words(idx, rep, key0, which, 2 * t + 2).x;
        const float u = u01(wa), z2 = z * z;
        // Marsaglia-Tsang's squeeze: u < 1 - 0.0331 z^4 implies the logarithmic test below, so nine draws in ten are
        // accepted without the two logarithms (the decision, hence the variate, is the same)
        if (u < fmaf(-0.0331f * z2, z2, 1.0f)) break;
        float lhs = 0.6931471805599453f * fast_lg2(u);
        float rhs = 0.5f * z2 + dd - dd * v + dd * 0.6931471805599453f * fast_lg2(v);
        if (lhs < rhs) break;
    }
    if (small_alpha) x *= fast_ex2(fast_lg2(u01(w0)) * inv_alpha);
    return x * inv_mu;
}

// ONE attempt (number t) of sample_gamma's loop, for callers that keep the loop themselves so that a lane which rejects
// retries in its next trip instead of holding its warp inside the sampler.  Same operations in the same order as the
// loop body above, so `x` (start it at dd) ends up with the same bits; true = accepted (the caller also stops at t = 14).
__device__ __forceinline__ bool gamma_attempt(float dd, float cc, int small_alpha, uint32_t t, uint32_t w0, uint32_t idx,
                                              uint32_t rep, uint32_t key0, int which, float &x)
{
    uint2 e0 = ext_words(idx, rep, key0, which, 2 * t + 1);
    float rad = sqrtf(-1.3862943611198906f * fast_lg2(u01(e0.x)));
    float z = rad * fast_cos2pi(u01(e0.y));
    float v = fmaf(cc, z, 1.0f);
    if (v <= 0.0f) return false;
    v = v * v * v;
    x = dd * v;
    const uint32_t wa = (t == 0 && !small_alpha) ? w0 : ext_words(idx, rep, key0, which, 2 * t + 2).x;
    const float u = u01(wa), z2 = z * z;
    if (u < fmaf(-0.0331f * z2, z2, 1.0f)) return true;  // the squeeze, as in sample_gamma
    float lhs = 0.6931471805599453f * fast_lg2(u);
    float rhs = 0.5f * z2 + dd - dd * v + dd * 0.6931471805599453f * fast_lg2(v);
    return lhs < rhs;
}
// what sample_gamma does after its loop
__device__ __forceinline__ float gamma_finish(float x, float inv_mu, float inv_alpha, int small_alpha, uint32_t w0)
{
    if (small_alpha) x *= fast_ex2(fast_lg2(u01(w0)) * inv_alpha);
    return x * inv_mu;
}

template <int KIND>
__device__ __forceinline__ float sample(const DistF &d, uint32_t w0, uint32_t idx, uint32_t rep,
                                        uint32_t key0, int which)
{
    const int kind = KIND >= 0 ? KIND : d.kind;
    switch (kind) {
    case MQ_M:
        return sample_m(d, w0);
    case MQ_H2:
        return sample_h2(d, w0);
    case MQ_E:
        return sample_erlang(d.a, d.r, w0, idx, rep, key0, which);
    case MQ_GAMMA:
        return sample_gamma(d.a, d.b, d.c, d.d, d.r, w0, idx, rep, key0, which);
    case MQ_PA:
        return sample_pa(d, w0);
    case MQ_D:  // distributions.py:614-626
        return d.a;
    case MQ_C2: {  // distributions.py:542-560; d.a = -ln2/mu1, d.b = -ln2/mu2
        uint2 e = ext_words(idx, rep, key0, which, 1);
        float res = fast_lg2(u01(w0)) * d.a;
        if (e.x < d.T) res += fast_lg2(u01(e.y)) * d.b;
        return res;
    }
    case MQ_UNIFORM:  // distributions.py:298-309; d.a = mean - half, d.b = 2 half
        return fmaf(d.b, u01(w0), d.a);
    case MQ_NORM: {  // distributions.py:209-216 (Box-Muller); d.a = mean, d.b = std
        uint2 e = ext_words(idx, rep, key0, which, 1);
        float rad = sqrtf(-1.3862943611198906f * fast_lg2(u01(w0)));
        return fmaf(d.b * rad, fast_cos2pi(u01(e.x)), d.a);
    }
    case MQ_WEIBULL:  // distributions.py:94-102: pow(-log(p) * W, -1/k); d.a = W, d.b = -1/k
        return fast_ex2(fast_lg2(-0.6931471805599453f * fast_lg2(u01(w0)) * d.a) * d.b);
    default:
        return 0.0f;
    }
}

// -------------------------------------------------------------------------------------
// fp64 tier of the same samplers (MQ_FIFO_FP64: the A/B kernel that measures what the fp32 production tier costs in
// accuracy).  Same words, same formulas, evaluated with the CUDA math library's double log / exp / cos / sqrt -- the
// generator specification of DESIGN.md section 3 as oracle/mq_oracle.c states it in fp64.
// -------------------------------------------------------------------------------------
struct DistD {
    int kind;
    int r;
    uint32_t T;
    uint32_t pad;
    double a, b, c, d;
};

__device__ __forceinline__ double u01d(uint32_t w) { return ((double)w + 0.5) * 0x1p-32; }

static __device__ __noinline__ double sample_d(int kind, int r, uint32_t T, double a, double b, double c, double d,
                                               uint32_t w0, uint32_t idx, uint32_t rep, uint32_t key0, int which)
{
    switch (kind) {
    case MQ_M:  // a = 1/mu
        return -log(u01d(w0)) * a;
    case MQ_H2:  // a = 1/mu1, b = 1/mu2 (distributions.py:407-429; one word, see sample_h2)
        if (w0 < T) return -log(((double)w0 + 0.5) / (double)T) * a;
        return -log(((double)(w0 - T) + 0.5) / (4294967296.0 - (double)T)) * b;
    case MQ_E: {  // a = 1/mu
        double acc = log(u01d(w0));
        uint2 e = make_uint2(0u, 0u);
        for (int k = 1; k < r; ++k) {
            if (k & 1) e = ext_words(idx, rep, key0, which, (uint32_t)(k + 1) >> 1);
            acc += log(u01d((k & 1) ? e.x : e.y));
        }
        return -acc * a;
    }
    case MQ_GAMMA: {  // a = dd, b = cc, c = 1/mu, d = 1/alpha, r = (alpha < 1)
        double x = a;
        for (uint32_t t = 0; t < 15; ++t) {
            uint2 e0 = ext_words(idx, rep, key0, which, 2 * t + 1);
            const double z = sqrt(-2.0 * log(u01d(e0.x))) * cos(6.283185307179586 * u01d(e0.y));
            double v = 1.0 + b * z;
            if (v <= 0.0) continue;
            v = v * v * v;
            x = a * v;
            const uint32_t wa = (t == 0 && !r) ? w0 : ext_words(idx, rep, key0, which, 2 * t + 2).x;
            if (log(u01d(wa)) < 0.5 * z * z + a - a * v + a * log(v)) break;
        }
        if (r) x *= exp(log(u01d(w0)) * d);
        return x * c;
    }
    case MQ_PA:  // a = -1/alpha, b = K
        return b * exp(log(u01d(w0)) * a);
    case MQ_D:
        return a;
    case MQ_C2: {  // a = 1/mu1, b = 1/mu2
        uint2 e = ext_words(idx, rep, key0, which, 1);
        double res = -log(u01d(w0)) * a;
        if (e.x < T) res -= log(u01d(e.y)) * b;
        return res;
    }
    case MQ_UNIFORM:  // a = mean - half, b = 2 half
        return fma(b, u01d(w0), a);
    case MQ_NORM: {  // a = mean, b = std
        uint2 e = ext_words(idx, rep, key0, which, 1);
        return fma(b * sqrt(-2.0 * log(u01d(w0))), cos(6.283185307179586 * u01d(e.x)), a);
    }
    case MQ_WEIBULL:  // a = W, b = -1/k
        return exp(log(-log(u01d(w0)) * a) * b);
    default:
        return 0.0;
    }
}

__device__ __forceinline__ double sample_d(const DistD &q, uint32_t w0, uint32_t idx, uint32_t rep, uint32_t key0,
                                           int which)
{
    return sample_d(q.kind, q.r, q.T, q.a, q.b, q.c, q.d, w0, idx, rep, key0, which);
}

// sum of x^k, k = 1..4 (5 issue slots)
__device__ __forceinline__ void acc4(float (&s)[4], float x)
{
    float x2 = x * x;
    s[0] += x;
    s[1] += x2;
    s[2] = fmaf(x2, x, s[2]);
    s[3] = fmaf(x2, x2, s[3]);
}

__device__ __forceinline__ void acc4(double (&s)[4], double x)
{
    double x2 = x * x;
    s[0] += x;
    s[1] += x2;
    s[2] = fma(x2, x, s[2]);
    s[3] = fma(x2, x2, s[3]);
}

}  // namespace mq
