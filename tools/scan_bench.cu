// Microbenchmark: the epilogue's scan_unit in isolation (8 or 4 warps per SM), cycles per 256-column unit.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I vq_vae_2_pytorch_b200/csrc -o tools/scan_bench.bin tools/scan_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <cuda_runtime.h>
#include "tc_kernel.cuh"
using namespace vqb200::tc;

__global__ void k(int reps, unsigned long long* out, float* sink) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(smem_u32(&tptr), 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t lane_addr = tptr + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) & 1) * 256;
    float acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        float m1, m2;
        int j = scan_unit<false>(lane_addr, m1, m2, nullptr);
        acc += m1 + m2 + j;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tptr, 512);
}

int main() {
    for (int warps : {4, 8}) {
        unsigned long long* d_out; float* d_sink;
        const int grid = 148, reps = 200;
        cudaMalloc(&d_out, grid * 8); cudaMalloc(&d_sink, grid * warps * 32 * 4);
        k<<<grid, warps * 32>>>(reps, d_out, d_sink);
        k<<<grid, warps * 32>>>(reps, d_out, d_sink);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long h[148]; cudaMemcpy(h, d_out, grid * 8, cudaMemcpyDeviceToHost);
        double cyc = 0; for (int i = 0; i < grid; ++i) cyc += (double)h[i]; cyc /= grid;
        printf("scan_unit warps/SM=%d : %.0f cycles per unit (256 cols x 32 lanes per warp) (%s)\n", warps, cyc / reps, cudaGetErrorString(e));
    }
    return 0;
}
