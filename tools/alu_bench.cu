// Microbenchmark: per-SMSP issue throughput of the ALU ops the epilogue can be built from.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/alu_bench.bin tools/alu_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

enum Op { FMNMX2 = 0, FMNMX3, IMNMX2, IMNMX3, UMNMX3, PRMT, LOP3, FADD, FFMA, HMNMX2, FSETSEL, IADD, NOPS };
const char* NAMES[] = {"fmin(a,b)", "fmin3(a,b,c)", "imin(a,b)", "imin3", "umin3", "prmt", "lop3", "fadd", "ffma", "hmin2", "fsetp+fsel", "iadd"};

template <int OP>
__global__ void k(int iters, unsigned long long* out, float* sink, const float* in) {
    float a[8];
    uint32_t u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = in[threadIdx.x + 32 * i]; u[i] = __float_as_uint(a[i]); }
    const float c0 = in[1000], c1 = in[1001];
    const uint32_t k0 = __float_as_uint(c0), k1 = __float_as_uint(c1);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == FMNMX2) a[i] = fminf(a[i], c0 + (float)r);
                if (OP == FMNMX3) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c0), "f"(c1));
                if (OP == IMNMX2) asm volatile("min.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(k0));
                if (OP == IMNMX3) u[i] = (uint32_t)min(min((int)u[i], (int)k0), (int)k1 + r);
                if (OP == UMNMX3) u[i] = min(min(u[i], k0), k1 + (uint32_t)r);
                if (OP == PRMT) u[i] = __byte_perm(u[i], k0, 0x3214 + r);
                if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(u[i]) : "r"(k0), "r"(k1));
                if (OP == FADD) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c0));
                if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c0), "f"(c1));
                if (OP == HMNMX2) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(k0));
                if (OP == FSETSEL) a[i] = (a[i] < c0 + (float)r) ? a[i] : c1;
                if (OP == IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(k0));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int OP>
void run(int warps_per_sm, const float* d_in) {
    unsigned long long* d_out; float* d_sink;
    const int grid = 148, iters = 512;
    cudaMalloc(&d_out, grid * 8); cudaMalloc(&d_sink, grid * warps_per_sm * 32 * 4);
    k<OP><<<grid, warps_per_sm * 32>>>(iters, d_out, d_sink, d_in);
    k<OP><<<grid, warps_per_sm * 32>>>(iters, d_out, d_sink, d_in);
    cudaDeviceSynchronize();
    unsigned long long h[148];
    cudaMemcpy(h, d_out, grid * 8, cudaMemcpyDeviceToHost);
    double cyc = 0; for (int i = 0; i < grid; ++i) cyc += (double)h[i]; cyc /= grid;
    double ops_per_warp = (double)iters * 64;
    printf("%-14s warps/SM=%-2d : %8.0f cyc, %6.2f cyc per warp-instr per SMSP-warp, %6.2f warp-instr/cyc/SMSP\n", NAMES[OP], warps_per_sm, cyc,
           cyc / ops_per_warp, ops_per_warp * warps_per_sm / 4.0 / cyc);
    cudaFree(d_out); cudaFree(d_sink);
}

int main() {
    float* d_in; cudaMalloc(&d_in, 4096 * 4);
    float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = 1.0f + i * 0.001f;
    cudaMemcpy(d_in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int w : {4, 8, 16}) {
        run<FMNMX2>(w, d_in); run<FMNMX3>(w, d_in); run<IMNMX2>(w, d_in); run<IMNMX3>(w, d_in); run<UMNMX3>(w, d_in);
        run<PRMT>(w, d_in); run<LOP3>(w, d_in); run<FADD>(w, d_in); run<FFMA>(w, d_in); run<HMNMX2>(w, d_in);
        run<FSETSEL>(w, d_in); run<IADD>(w, d_in);
    }
    return 0;
}
