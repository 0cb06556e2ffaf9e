// Which issue pipe do the 16-bit packed minimum (HMNMX2) and the fp32 -> 16-bit pack (F2FP) use on sm_100a?
// One warp per SMSP-sized block set; each variant runs a long dependent-free stream of the instruction(s) and reports
// warp-instructions per cycle per SMSP.  If HMNMX2 / F2FP shared the ALU pipe with FMNMX3, the mixed streams would take the SUM
// of the separate times; on a different pipe they overlap.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#define ITERS 4096
template <int MODE>
__global__ void k(float* out, unsigned long long* cyc, float seed) {
    float a[8]; unsigned int h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; h[i] = __float_as_uint(seed * (i + 1)) + threadIdx.x; }
    float x = seed * 3.f, y = seed * 5.f;
    unsigned int hx = __float_as_uint(seed) ^ 0x12345u, hy = hx * 3u;
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0 || MODE == 3 || MODE == 4 || MODE == 6)      // FMNMX3
                asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(x), "f"(y));
            if (MODE == 1 || MODE == 3)                   // HMNMX2 (bf16x2)
                asm volatile("min.bf16x2 %0, %0, %1;" : "+r"(h[i]) : "r"(hx));
            if (MODE == 2 || MODE == 4)                   // F2FP pack: two fp32 -> bf16x2
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(y));
            if (MODE == 5 || MODE == 6)                   // f16x2 min
                asm volatile("min.f16x2 %0, %0, %1;" : "+r"(h[i]) : "r"(hy));
            if (MODE == 7)                                // FFMA reference (fma pipe)
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(x), "f"(y));
            if (MODE == 8) {                              // FMNMX3 + FFMA (known different pipes)
                asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(x), "f"(y));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(y), "f"(y));
            }
        }
    }
    unsigned long long t1 = clock64();
    float s = x; unsigned int hs = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += a[i]; hs ^= h[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(hs);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE> void run(const char* name, int per_iter) {
    float* out; unsigned long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    for (int warps = 1; warps <= 8; warps *= 2) {       // warps per SMSP (block = 4 SMSPs x warps x 32)
        k<MODE><<<148, 128 * warps>>>(out, cyc, 1.5f); cudaDeviceSynchronize();
        k<MODE><<<148, 128 * warps>>>(out, cyc, 1.5f); cudaDeviceSynchronize();
        unsigned long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-28s warps/SMSP %d: %.3f warp-instr / cycle / SMSP\n", name, warps, (double)ITERS * 8 * per_iter * warps / (double)c);
    }
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("FMNMX3", 1); run<1>("HMNMX2.BF16", 1); run<5>("HMNMX2.F16", 1); run<2>("F2FP.BF16.PACK", 1); run<7>("FFMA", 1);
    run<3>("FMNMX3 + HMNMX2.BF16", 2); run<6>("FMNMX3 + HMNMX2.F16", 2); run<4>("FMNMX3 + F2FP", 2); run<8>("FMNMX3 + FFMA", 2);
    return 0;
}
