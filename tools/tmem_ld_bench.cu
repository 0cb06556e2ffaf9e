// Microbenchmark: tcgen05.ld throughput per SM for several shapes / repeat counts / warp counts.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_ld_bench tools/tmem_ld_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X> __device__ __forceinline__ uint32_t ld_cols(uint32_t taddr);
#define LD_IMPL(X, REGLIST, ...)                                                                                   \
    template <> __device__ __forceinline__ uint32_t ld_cols<X>(uint32_t taddr) {                                     \
        uint32_t v[X];                                                                                               \
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x" #X ".b32 {" REGLIST "}, [%" #X "];" : __VA_ARGS__ : "r"(taddr) : "memory"); \
        uint32_t s = 0;                                                                                              \
        _Pragma("unroll") for (int i = 0; i < X; ++i) s ^= v[i];                                                     \
        return s;                                                                                                    \
    }
#define O8(b) "=r"(v[b+0]), "=r"(v[b+1]), "=r"(v[b+2]), "=r"(v[b+3]), "=r"(v[b+4]), "=r"(v[b+5]), "=r"(v[b+6]), "=r"(v[b+7])
LD_IMPL(8, "%0,%1,%2,%3,%4,%5,%6,%7", O8(0))
LD_IMPL(16, "%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15", O8(0), O8(8))
LD_IMPL(32, "%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31",
        O8(0), O8(8), O8(16), O8(24))
LD_IMPL(64, "%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
            "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63",
        O8(0), O8(8), O8(16), O8(24), O8(32), O8(40), O8(48), O8(56))

// each warp reads `cols` columns of its lane quarter, `reps` times; waits once per pass (loads pipelined)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// the LAST warp of the CTA issues `n_mma` back-to-back M128 N256 K16 bf16 MMAs into TMEM columns 256..511 (mma_on)
template <int X>
__global__ void k_bench(int cols, int reps, int mma_on, unsigned long long* out, uint32_t* sink) {
    __shared__ uint32_t tptr;
    __shared__ __align__(8) unsigned long long mbar;
    extern __shared__ __align__(1024) unsigned char dsm[];
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const int wait_each = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(dsm)[i] = 0x3f803f80u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    if (mma_on && warp == nwarps - 1) {
        if ((threadIdx.x & 31) == 0) {
            const uint32_t a0 = (smem_u32(dsm) + 1023u) & ~1023u, b0 = a0 + 16384u;
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (32u << 17) | (8u << 24);
            const long long t0 = clock64();
            for (int i = 0; i < mma_on; ++i) {
                const uint64_t ad = desc_sw128(a0) + 2u * (i & 3), bd = desc_sw128(b0) + 2u * (i & 3);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tptr + 256u), "l"(ad), "l"(bd), "r"(idesc), "r"(i > 0 ? 1u : 0u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
            uint32_t ok = 0;
            while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(smem_u32(&mbar)) : "memory");
            out[148 + blockIdx.x] = (unsigned long long)(clock64() - t0);
        }
        __syncwarp();
    } else {
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        for (int c = 0; c < cols; c += X) {
            acc ^= ld_cols<X>(base + c);
            if (wait_each) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
    }
    __syncthreads();
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tptr) : "memory");
}

template <int X>
void run(int warps, int n_mma) {
    unsigned long long* d_out; uint32_t* d_sink;
    const int grid = 148, cols = 256, reps = 128;
    const int threads = (warps + (n_mma ? 1 : 0)) * 32;
    cudaMalloc(&d_out, 2 * grid * 8); cudaMalloc(&d_sink, grid * threads * 4);
    cudaMemset(d_out, 0, 2 * grid * 8);
    cudaFuncSetAttribute(k_bench<X>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k_bench<X><<<grid, threads, 65536>>>(cols, reps, n_mma, d_out, d_sink);
    k_bench<X><<<grid, threads, 65536>>>(cols, reps, n_mma, d_out, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[296];
    cudaMemcpy(h, d_out, 2 * grid * 8, cudaMemcpyDeviceToHost);
    double cyc = 0, mcyc = 0; for (int i = 0; i < grid; ++i) { cyc += (double)h[i]; mcyc += (double)h[148 + i]; } cyc /= grid; mcyc /= grid;
    double bytes = (double)warps * 32 * cols * 4 * reps;
    printf("x%-3d read_warps=%-2d mma=%-5d : ld %9.0f cyc -> %7.1f B/cyc/SM, %6.1f cyc/LDTM | mma %9.0f cyc = %6.1f cyc/MMA (%s)\n", X, warps, n_mma,
           cyc, bytes / cyc, cyc / ((double)cols / X * reps), mcyc, n_mma ? mcyc / n_mma : 0.0, cudaGetErrorString(e));
    cudaFree(d_out); cudaFree(d_sink);
}

int main() {
    for (int w : {4, 8}) {
        for (int m : {0, 2000}) {
            run<16>(w, m); run<32>(w, m); run<64>(w, m);
        }
    }
    return 0;
}
