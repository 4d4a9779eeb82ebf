// Issue-rate probes for the pipes a 255-bit multiplier could use on sm_100a (B200):
//   0: IMAD.WIDE.U32 (carry-chained, the form fe_mul uses)      1: DFMA (fma.rz.f64)
//   2: both interleaved 1:1 in one instruction stream           3: DFMA : IMAD.WIDE = 2:1
//   4: IADD3 (add.cc chains) interleaved 1:1 with IMAD.WIDE
// Prints ops/s per class over all SMs and lanes/clk/SM at the clock read from the device.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define MAD4(c0,c1,c2,c3,c4,c5,c6,c7,c8,a0,a1,a2,a3,b) \
    asm volatile("mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\tmadc.lo.cc.u32 %2, %10, %13, %2;\n\t" \
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\tmadc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\t" \
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\taddc.u32 %8, %8, 0;" \
        : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3), "+r"(c4), "+r"(c5), "+r"(c6), "+r"(c7), "+r"(c8) \
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b))
#define DF(x, a, b) asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(x) : "d"(a), "d"(b))
#define ADD8(c0,c1,c2,c3,c4,c5,c6,c7,d) \
    asm volatile("add.cc.u32 %0,%0,%8; addc.cc.u32 %1,%1,%8; addc.cc.u32 %2,%2,%8; addc.cc.u32 %3,%3,%8;" \
        "addc.cc.u32 %4,%4,%8; addc.cc.u32 %5,%5,%8; addc.cc.u32 %6,%6,%8; addc.u32 %7,%7,%8;" \
        : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3), "+r"(c4), "+r"(c5), "+r"(c6), "+r"(c7) : "r"(d))

__global__ void __launch_bounds__(256) k(int mode, int iters, uint32_t seed, double *out) {
    uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3;
    uint32_t c0 = 1 + blockIdx.x, c1 = 2, c2 = 3, c3 = 4, c4 = 5, c5 = 6, c6 = 7, c7 = 8, c8 = 0;
    uint32_t d0 = 9, d1 = 10, d2 = 11, d3 = 12, d4 = 13, d5 = 14, d6 = 15, d7 = 16, d8 = 0;
    uint32_t e0 = 1, e1 = 2, e2 = 3, e3 = 4, e4 = 5, e5 = 6, e6 = 7, e7 = 8;
    uint32_t g0 = 9, g1 = 8, g2 = 7, g3 = 6, g4 = 5, g5 = 4, g6 = 3, g7 = 2;
    double x0 = 1.0 + threadIdx.x, x1 = 2.0, x2 = 3.0, x3 = 4.0, x4 = 5.0, x5 = 6.0, x6 = 7.0, x7 = 8.0;
    double fa = 1.0000001 + 1e-9 * seed, fb = 0.9999999;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (mode == 0 || mode == 2 || mode == 3 || mode == 4) {
                MAD4(c0, c1, c2, c3, c4, c5, c6, c7, c8, a0, a1, a2, a3, d0 ^ d7);
                MAD4(d0, d1, d2, d3, d4, d5, d6, d7, d8, a1, a2, a3, a0, c0 ^ c7);   // 16 IMAD.WIDE
            }
            if (mode == 1 || mode == 2 || mode == 3) {
                DF(x0, fa, fb); DF(x1, fa, fb); DF(x2, fa, fb); DF(x3, fa, fb);
                DF(x4, fa, fb); DF(x5, fa, fb); DF(x6, fa, fb); DF(x7, fa, fb);
                DF(x0, fb, fa); DF(x1, fb, fa); DF(x2, fb, fa); DF(x3, fb, fa);
                DF(x4, fb, fa); DF(x5, fb, fa); DF(x6, fb, fa); DF(x7, fb, fa);   // 16 DFMA
            }
            if (mode == 3) {
                DF(x0, fa, fa); DF(x1, fa, fa); DF(x2, fa, fa); DF(x3, fa, fa);
                DF(x4, fa, fa); DF(x5, fa, fa); DF(x6, fa, fa); DF(x7, fa, fa);
                DF(x0, fb, fb); DF(x1, fb, fb); DF(x2, fb, fb); DF(x3, fb, fb);
                DF(x4, fb, fb); DF(x5, fb, fb); DF(x6, fb, fb); DF(x7, fb, fb);   // 16 more
            }
            if (mode == 4) {
                ADD8(e0, e1, e2, e3, e4, e5, e6, e7, g7);
                ADD8(g0, g1, g2, g3, g4, g5, g6, g7, e7);                        // 16 IADD3
            }
        }
    }
    uint32_t r = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7 ^ c8 ^ d0 ^ d1 ^ d2 ^ d3 ^ d4 ^ d5 ^ d6 ^ d7 ^ d8 ^ e0 ^ e7 ^ g0 ^ g7 ^ e3 ^ g3;
    double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (r == 0x12345678u || s == 1.2345) out[0] = s + r;
}

int main() {
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double *d;
    cudaMalloc(&d, 64);
    const int blocks = pr.multiProcessorCount * 8, iters = 4096;
    const char *names[5] = {"imad_wide", "dfma", "imad_wide+dfma 1:1", "imad_wide+dfma 1:2", "imad_wide+iadd3 1:1"};
    printf("{\"sms\": %d, \"clock_khz\": %d, \"probes\": {", pr.multiProcessorCount, khz);
    for (int mode = 0; mode < 5; mode++) {
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        k<<<blocks, 256>>>(mode, iters / 8, 1, d);
        cudaEventRecord(a);
        k<<<blocks, 256>>>(mode, iters, 7, d);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        double thr = (double)blocks * 256 * iters * 4 * 16;   // ops of each class present
        double imad = (mode != 1) ? thr : 0, df = mode == 1 || mode == 2 ? thr : mode == 3 ? 2 * thr : 0, ia = mode == 4 ? thr : 0;
        double sec = ms * 1e-3, clk = khz * 1e3, sm = pr.multiProcessorCount;
        printf("%s\"%s\": {\"ms\": %.3f, \"imad_wide_T\": %.3f, \"dfma_T\": %.3f, \"iadd3_T\": %.3f, \"imad_lanes_clk_sm\": %.1f, \"dfma_lanes_clk_sm\": %.1f, \"iadd3_lanes_clk_sm\": %.1f}",
               mode ? ", " : "", names[mode], ms, imad / sec / 1e12, df / sec / 1e12, ia / sec / 1e12, imad / sec / clk / sm, df / sec / clk / sm, ia / sec / clk / sm);
    }
    printf("}}\n");
    return 0;
}
