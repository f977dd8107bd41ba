// microbench_issue.cu -- issue rate (lane-operations per clock per SM) of the instruction kinds K2's inner
// loop is made of, on the GPU it runs on.  SURVEY 8(d) asks for the measured rate of IMAD.WIDE / LOP3 because
// the issue roofline's denominator depends on it.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -o tools/microbench_issue tools/microbench_issue.cu ; run on the GPU box; prints one JSON object.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096
#define CHAINS 8

template <int OP>
__global__ void __launch_bounds__(256) bench(uint32_t *out, uint32_t seed)
{
    uint32_t a[CHAINS], b[CHAINS];
    float f[CHAINS], g[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
        a[i] = seed + threadIdx.x * 7 + i;
        b[i] = seed * 3 + i + blockIdx.x;
        f[i] = 1.0f + 1e-3f * (float)(threadIdx.x + i);
        g[i] = 0.5f + 1e-4f * (float)i;
    }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (OP == 0) {  // IMAD.WIDE.U32 (Philox round multiply): 64-bit product, both halves consumed
                uint32_t lo, hi;
                asm volatile("{ .reg .u64 t; mul.wide.u32 t, %2, 0xD256D193; mov.b64 {%0, %1}, t; }"
                             : "=r"(lo), "=r"(hi) : "r"(a[i]));
                a[i] = lo + hi;  // one IADD keeps both halves live (IADD issues at full rate)
            } else if (OP == 1) {  // LOP3 (3-input xor)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
            } else if (OP == 2) {  // FFMA
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(g[i]), "f"(g[(i + 1) % CHAINS]));
            } else if (OP == 3) {  // FMNMX
                asm volatile("min.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g[i]));
                asm volatile("max.f32 %0, %0, %1;" : "+f"(g[i]) : "f"(f[(i + 1) % CHAINS]));
            } else if (OP == 4) {  // IADD3
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            } else if (OP == 5) {  // MUFU.LG2
                asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
            } else if (OP == 6) {  // IMAD (32-bit low product + add)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
            } else if (OP == 7) {  // FSEL / SEL through setp + selp
                asm volatile("{ .reg .pred p; setp.lt.f32 p, %0, %1; selp.f32 %0, %2, %0, p; }"
                             : "+f"(f[i]) : "f"(g[i]), "f"(g[(i + 3) % CHAINS]));
            } else if (OP == 9) {  // one Philox2x32 round: IMAD.WIDE.U32 + 3-input LOP3
                uint32_t lo, hi;
                asm volatile("{ .reg .u64 t; mul.wide.u32 t, %2, 0xD256D193; mov.b64 {%0, %1}, t; }"
                             : "=r"(lo), "=r"(hi) : "r"(a[i]));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a[i]) : "r"(hi), "r"(b[i]), "r"(seed + it));
                b[i] = lo;
            } else if (OP == 8) {  // DFMA
                double d;
                asm volatile("{ .reg .f64 x, y; cvt.f64.f32 x, %1; cvt.f64.f32 y, %2; fma.rn.f64 %0, x, y, x; }"
                             : "=d"(d) : "f"(f[i]), "f"(g[i]));
                f[i] = (float)d;
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc ^= a[i] ^ b[i] ^ __float_as_uint(f[i]) ^ __float_as_uint(g[i]);
    if (acc == 0x12345678u) out[0] = acc;
}

template <int OP>
static double run(int sms, double ops_per_iter, uint32_t *out, int clock_khz)
{
    const int grid = sms * 8;
    bench<OP><<<grid, 256>>>(out, 12345u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        bench<OP><<<grid, 256>>>(out, 12345u + r);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double lane_ops = (double)grid * 256.0 * ITER * CHAINS * ops_per_iter;
    const double clocks = best * 1e-3 * clock_khz * 1e3;
    return lane_ops / clocks / sms;  // lane-ops per clock per SM at the nominal max clock
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    uint32_t *out;
    cudaMalloc(&out, 4);
    const int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"unit\": \"lane-ops per clock per SM at clock_khz\"", p.name, sms, khz);
    printf(", \"imad_wide_u32_plus_iadd\": %.1f", run<0>(sms, 1, out, khz));
    printf(", \"lop3\": %.1f", run<1>(sms, 1, out, khz));
    printf(", \"ffma\": %.1f", run<2>(sms, 1, out, khz));
    printf(", \"fmnmx\": %.1f", run<3>(sms, 2, out, khz));
    printf(", \"iadd\": %.1f", run<4>(sms, 1, out, khz));
    printf(", \"mufu_lg2\": %.1f", run<5>(sms, 1, out, khz));
    printf(", \"imad_lo\": %.1f", run<6>(sms, 1, out, khz));
    printf(", \"setp_selp\": %.1f", run<7>(sms, 1, out, khz));
    printf(", \"cvt_dfma_cvt\": %.1f", run<8>(sms, 1, out, khz));
    printf(", \"philox_round_imadwide_plus_lop3\": %.1f", run<9>(sms, 1, out, khz));
    printf("}\n");
    return 0;
}
