// tcgen05 / TMEM / bulk-copy (TMA engine) assignment kernel for sm_100a.
//
// One persistent, warp-specialised CTA per SM computes, for tiles of 128 input rows, a LOWER BOUND of
// the squared distance to every code on the 5th-gen tensor cores, certifies the arg-min with a rigorous
// error bound, and fuses gather / straight-through output / commitment loss / code statistics behind it
// (reference vqvae.py:43-56,72-73 in one pass over x).  Rows the bound cannot certify (near-ties) are
// appended to a list and re-scored by the exact fp32 kernels of simt_kernels.cuh, so the result is the
// exact-arithmetic arg-min regardless of tensor-core precision.
//
// GEMM formulation (all operands bf16, fp32 accumulation in TMEM; K-major, M=128, N=256, K=16 per MMA):
//   A row i  = [ xh | xl | xh | 1 1 1  o1 o2 o3  nx 1  0.. ]      xh+xl ~ x_i (split bf16, 16 bits)
//   B row k  = [ eh | eh | el | b1 b2 b3 1 1 1  -cA*ne  -cB*ne^2  0.. ]   eh+el ~ -2 e_k
//   acc_ik   = -2 x_i.e_k + ||e_k||^2 + off_i - E'_ik  =  lower bound of  ||x_i-e_k||^2 + (off_i-||x_i||^2)
// with b1+b2+b3 = ||e_k||^2 and o1+o2+o3 = off_i = ||x_i||^2 (1+2^-9) exactly (3-term bf16 splits), and
// E'_ik = cA ||x_i|| ||e_k|| + cB ||e_k||^2 an upper bound of the filter's own error, folded into the
// contraction.  NSPLIT=1 drops the xl/el blocks (plain bf16 filter, wider bound).
//
// Warp roles (640 threads = 5 warpgroups, registers re-balanced with setmaxnreg: 128/128/72/72/56/24):
//   WG0, WG1: two epilogue warpgroups, one per TMEM buffer (thread = row; arg-min and the runner-up from a
//   two-class min scan, ONE ALU op per score) | WG2: output (gather, straight-through value, loss) |
//   WG3: fp32->bf16 converters | WG4: w16 bulk-copy producer, w17 MMA issuer + TMEM owner (w18-19 idle).
#pragma once
#include <atomic>
#include <cuda_bf16.h>
#include <cudaTypedefs.h>   // CUtensorMap, PFN_cuTensorMapEncodeTiled (resolved at run time: no link-time libcuda dependency)

#include "common.cuh"

namespace vqb200 {
namespace tc {

constexpr int TILE_M = 128;        // rows per tile = UMMA M
constexpr int UNIT_N = 128;        // codes per MMA = UMMA N; one TMEM accumulator buffer = 128 columns
constexpr int NBUF = 4;            // TMEM buffers (4 x 128 columns = all of TMEM): the MMAs run up to 3 units ahead of the scans
constexpr int TC_D = 64;           // supported dim: one 128-byte swizzle row of bf16
constexpr int THREADS = 768;        // 6 warpgroups; register budgets re-balanced with setmaxnreg
// Warp -> role map.  The SM's issue arbiter favours HIGHER warp ids (measured: the warpgroup with the higher
// ids ran its identical scan 30% faster), so the latency-critical roles sit at the top: the MMA issuer and the
// producer, then the converters (their latency serialises with the MMAs while A is single-buffered), then
// the output warps; the two epilogue groups have slack and take what is left.
constexpr int W_EPI0 = 0, W_EPI1 = 4, W_OUT = 8, W_CONV = 16, W_PROD = 20, W_MMA = 21;
constexpr int NORM_RING = 4;        // converters run at most 3 tiles ahead of the certificate (A stages + TMEM buffers bound the lead)
constexpr int RES_RING = 2;
#ifndef VQB200_RELAX_NS
#define VQB200_RELAX_NS 0    /* measured: 200 ns polling costs 3 us per launch (late wake-ups outweigh the freed issue slots) */
#endif
constexpr int RELAX_NS = VQB200_RELAX_NS;   // poll interval of the relaxed waits

// error-bound constants (see DESIGN.md "certificate"): dot-product error <= c1 ||x|| ||e||
//   split-bf16 (3 products): 3.1 * 2^-18 rounding + fp32 accumulation  -> c1 = 2^-16, cA = 2 c1
//   plain bf16             : (1+2^-9)^2 - 1                            -> c1 = 2^-8 * 1.01
//   tf32 (nsplit == 0)     : x is read as tf32 by the tensor core straight from its fp32 TMA stage (the low 13 mantissa bits
//                            are dropped: relative error < 2^-10 whether the hardware truncates or rounds), -2e is rounded to
//                            nearest tf32 when the image is built (2^-11)  -> c1 = (2^-10 + 2^-11 + 2^-21) * 1.01, cA = 2 c1
__host__ __device__ constexpr float bound_cA(int nsplit) { return nsplit == 3 ? 3.0517578125e-5f : (nsplit == 0 ? 2.96e-3f : 7.9e-3f); }
constexpr float BOUND_CB = 4.0e-6f;        // fp32 accumulation of <= 14 MMAs, relative to ||x||^2 + ||e||^2
constexpr float BOUND_UP = 1.03125f;       // covers the upward bf16 roundings of ||x||, ||e||, ||e||^2

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// for roles with slack (producer, converters, output): poll rarely -- a waiting warp's try_wait / branch / nanosleep loop
// was 23 % of all issued instructions (ncu), taken from the schedulers the scan warps need
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
template <int NS>
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
    if constexpr (NS == 0) mbar_wait(bar, parity);
    else while (!mbar_test_wait(bar, parity)) __nanosleep(NS);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// same with an L2 cache policy (createpolicy): x tiles are read again by the output warps and by the statistics kernel
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
// 3-D tiled TMA load (tensor map in kernel parameter space): box -> shared memory, completion on an mbarrier
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, int c0, int c1, int c2, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
                 ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// ---- 2-CTA (cta_group::2) variants: one CTA pair = one cluster; the leader (rank 0) issues the MMAs ----------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {   // address from mapa_rank()
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // local barrier, remote arrivals
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    }
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {     // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
// the loaded registers are in/out operands so no use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :: "memory");
}

__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red_add_f32(float* p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// shared-memory matrix descriptors (cute::UMMA::SmemDescriptor bit layout), K-major operands
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {   // rows of 128 B, 8-row atoms of 1024 B
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t desc_sw32(uint32_t saddr) {    // rows of 32 B, 8-row atoms of 256 B
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, K-major both, N=256, M=128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(UNIT_N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
// same with M = 256 for cta_group::2 (each CTA of the pair owns 128 of the 256 rows and 128 of the 256 B rows)
// kind::tf32: A = B = tf32 (format 2), K-major both (NCHW-physical x: A is MN-major, bit 15)
constexpr uint32_t IDESC_TF = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(UNIT_N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
constexpr uint32_t IDESC_TF_AMN = IDESC_TF | (1u << 15);
constexpr uint32_t IDESC2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(UNIT_N >> 3) << 17) | ((uint32_t)((2 * TILE_M) >> 4) << 24);

// byte offset of bf16 element (row, col) inside a 128B-swizzled K-major block (Swizzle<3,4,3>)
__host__ __device__ __forceinline__ uint32_t sw128_off(uint32_t row, uint32_t col) {
    uint32_t chunk = col >> 3;
    return row * 128u + (((chunk ^ (row & 7u)) << 4) | ((col & 7u) << 1));
}
// byte offset of the 16-byte chunk c (0/1) of `row` inside a 32B-swizzled K-major block (Swizzle<1,4,3>)
__host__ __device__ __forceinline__ uint32_t sw32_chunk_off(uint32_t row, uint32_t c) {
    return (row >> 3) * 256u + (row & 7u) * 32u + ((c ^ ((row >> 2) & 1u)) << 4);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {   // element 0 in the low half
    __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float bf16_round(float f) { return __bfloat162float(__float2bfloat16_rn(f)); }
// exact 3-term bf16 split of an fp32 value
__device__ __forceinline__ void split3(float f, float& a, float& b, float& c) {
    a = bf16_round(f);
    float r = f - a;
    b = bf16_round(r);
    c = bf16_round(r - b);
}

// tf32 helpers: round to nearest tf32 (10 explicit mantissa bits; the result is an fp32 with the low 13 bits clear, which the
// kind::tf32 MMA reads exactly) and the exact 3-term tf32 split of an fp32 value
__device__ __forceinline__ float tf32_rn(float f) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(f));
    return __uint_as_float(u);
}
__device__ __forceinline__ void split3_tf32(float f, float& a, float& b, float& c) {
    a = tf32_rn(f);
    float r = f - a;
    b = tf32_rn(r);
    c = tf32_rn(r - b);
}

// ---------------------------------------------------------------------------------------------------
// tensor-core operand image of the codebook (global memory, byte-identical to its smem layout)
//   [ eh : K*128 | el : K*128 | misc : K*32 ]   + enorm[K] (fp32 ||e_k||) kept next to it
//   [ e32 : 2 x K*128 | misc32 : K*32 ]  tf32 operand of -2e: two 32-dim k-blocks of 128-byte swizzled rows, and its misc rows
// ---------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t image_off_lo(int K) { return (size_t)K * 128; }
__host__ __device__ inline size_t image_off_misc(int K) { return (size_t)K * 256; }
__host__ __device__ inline size_t image_off_enorm(int K) { return (size_t)K * 288; }
__host__ __device__ inline size_t image_off_misc1(int K) { return (size_t)K * 292; }     // misc block carrying the plain-bf16 bound
__host__ __device__ inline size_t image_off_tf(int K) { return (size_t)K * 324; }         // tf32 operand [2][K][128 B]
__host__ __device__ inline size_t image_off_tfmisc(int K) { return (size_t)K * 580; }     // tf32 misc rows [K][32 B]
__host__ __device__ inline size_t image_bytes(int K) { return (size_t)K * 612; }

// operand-image rows of code k, 8-dim chunk c (0..7): e = the code's 64 fp32 components, e2 = ||e_k||^2; requires D == 64
__device__ __forceinline__ void tc_image_rows(const float* e_row, float e2, unsigned char* __restrict__ img, int K, int k, int c,
                                              float cA, float cA1, float cB) {
    const float* e = e_row + c * 8;
    {   // tf32 operand: dims 8c .. 8c+7 = two 16-byte chunks of the 128-byte row of k-block c / 4
        unsigned char* t = img + image_off_tf(K) + (size_t)(c >> 2) * K * 128 + (size_t)k * 128;
        const uint32_t ch = 2u * (uint32_t)(c & 3), sw = (uint32_t)k & 7u;
        *reinterpret_cast<float4*>(t + (((ch) ^ sw) << 4)) =
            make_float4(tf32_rn(-2.f * e[0]), tf32_rn(-2.f * e[1]), tf32_rn(-2.f * e[2]), tf32_rn(-2.f * e[3]));
        *reinterpret_cast<float4*>(t + (((ch + 1u) ^ sw) << 4)) =
            make_float4(tf32_rn(-2.f * e[4]), tf32_rn(-2.f * e[5]), tf32_rn(-2.f * e[6]), tf32_rn(-2.f * e[7]));
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float a = -2.f * e[2 * j], b = -2.f * e[2 * j + 1];
        float ah = bf16_round(a), bh = bf16_round(b);
        hi[j] = pack_bf16(ah, bh);
        lo[j] = pack_bf16(a - ah, b - bh);
    }
    uint32_t off = sw128_off((uint32_t)k, (uint32_t)c * 8);
    *reinterpret_cast<uint4*>(img + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(img + image_off_lo(K) + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    if (c == 0) {
        float b1, b2, b3;
        split3(e2, b1, b2, b3);
        float ne = sqrtf(e2);
        float t6 = -bf16_round(cA * ne * 1.0078125f);           // rounded up in magnitude
        float t7 = -bf16_round(cB * e2 * 1.0078125f);
        unsigned char* m = img + image_off_misc(K);
        *reinterpret_cast<uint4*>(m + sw32_chunk_off((uint32_t)k, 0)) =
            make_uint4(pack_bf16(b1, b2), pack_bf16(b3, 1.f), pack_bf16(1.f, 1.f), pack_bf16(t6, t7));
        *reinterpret_cast<uint4*>(m + sw32_chunk_off((uint32_t)k, 1)) = make_uint4(0, 0, 0, 0);
        reinterpret_cast<float*>(img + image_off_enorm(K))[k] = ne;
        // same row with the wider bound of the plain-bf16 filter (one image serves both filter precisions)
        const float t6b = -bf16_round(cA1 * ne * 1.0078125f);
        unsigned char* m1 = img + image_off_misc1(K);
        *reinterpret_cast<uint4*>(m1 + sw32_chunk_off((uint32_t)k, 0)) =
            make_uint4(pack_bf16(b1, b2), pack_bf16(b3, 1.f), pack_bf16(1.f, 1.f), pack_bf16(t6b, t7));
        *reinterpret_cast<uint4*>(m1 + sw32_chunk_off((uint32_t)k, 1)) = make_uint4(0, 0, 0, 0);
        // tf32 misc row [b1 b2 b3 1 | 1 1 -cA||e|| -cB||e||^2] (8 tf32 = one 32-byte swizzle row = one K = 8 MMA)
        float c1, c2, c3;
        split3_tf32(e2, c1, c2, c3);
        const float u6 = -tf32_rn(bound_cA(0) * ne * 1.001953125f);      // rounded up in magnitude
        const float u7 = -tf32_rn(cB * e2 * 1.001953125f);
        unsigned char* mt = img + image_off_tfmisc(K);
        *reinterpret_cast<float4*>(mt + sw32_chunk_off((uint32_t)k, 0)) = make_float4(c1, c2, c3, 1.f);
        *reinterpret_cast<float4*>(mt + sw32_chunk_off((uint32_t)k, 1)) = make_float4(1.f, 1.f, u6, u7);
    }
}


// ---------------------------------------------------------------------------------------------------
// epilogue scan (thread = row).  A warpgroup sees up to 256 "virtual" columns per tile (two 128-column accumulator
// units).  Every score feeds two families of running minima: class A = column mod 16 (16 registers) and class
// B = column div 16 (16 registers), each updated with one 3-input minimum per two scores -> ONE ALU op per score, no
// index bookkeeping in the hot loop.  A and B together identify a column (j = 16 b + a), so afterwards
//   m1  = the global minimum, found in one A-class a* and one B-class b*  (arg-min = 16 b* + a*),
//   m2  = min( second-smallest A-class minimum, second-smallest B-class minimum ) = the exact runner-up value,
// because any other column differs from the winner in its A-class or its B-class.  "Second smallest" counts
// duplicates, so an exact tie gives m2 == m1 and the row goes to the exact re-score.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fmin3(float a, float b, float c) { return fminf(a, fminf(b, c)); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :: "memory");
}

template <int REGS> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }

__device__ __forceinline__ float min16(const float* e) {
    return fminf(fmin3(fmin3(e[0], e[1], e[2]), fmin3(e[3], e[4], e[5]), fmin3(e[6], e[7], e[8])),
                 fmin3(fmin3(e[9], e[10], e[11]), fmin3(e[12], e[13], e[14]), e[15]));
}

// 32 scores of columns 32C .. 32C+31: class A = j mod 16 (rA[16]), class B = j div 16 (rB[16])
template <int C, bool DBG>
__device__ __forceinline__ void scan_chunk(const uint32_t (&v)[32], float (&rA)[16], float (&rB)[16], float* dbg) {
    float key[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) key[i] = __uint_as_float(v[i]);
    if (DBG && dbg) {
#pragma unroll
        for (int j = 0; j < 32; ++j) dbg[C * 32 + j] = key[j];
    }
#pragma unroll
    for (int a = 0; a < 16; ++a) rA[a] = fmin3(rA[a], key[a], key[a + 16]);
    rB[2 * C] = min16(key);
    rB[2 * C + 1] = min16(key + 16);
}

// software-pipelined walk over the 4 chunks (32 columns each) of one 128-column TMEM buffer: the load of chunk c+1 is in
// flight while chunk c is reduced.  CB = index of the buffer's first chunk within the warpgroup's 256 virtual columns
// (0 or 4), so the class-B minima land in fixed registers.
template <int CB, bool DBG>
__device__ __forceinline__ void scan_buffer(uint32_t lane_addr, float (&rA)[16], float (&rB)[16], float* dbg) {
    uint32_t va[32], vb[32];
    tmem_ld32(lane_addr, va);
    tmem_ld_wait(va);
    tmem_ld32(lane_addr + 32, vb);
    scan_chunk<CB + 0, DBG>(va, rA, rB, dbg);
    tmem_ld_wait(vb);
    tmem_ld32(lane_addr + 64, va);
    scan_chunk<CB + 1, DBG>(vb, rA, rB, dbg);
    tmem_ld_wait(va);
    tmem_ld32(lane_addr + 96, vb);
    scan_chunk<CB + 2, DBG>(va, rA, rB, dbg);
    tmem_ld_wait(vb);
    scan_chunk<CB + 3, DBG>(vb, rA, rB, dbg);
}

// smallest and second-smallest (duplicates count) of 16 values: tournament on (lo, hi) pairs
__device__ __forceinline__ void two_smallest16(const float (&r)[16], float& lo, float& hi) {
    float l[8], h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { l[i] = fminf(r[2 * i], r[2 * i + 1]); h[i] = fmaxf(r[2 * i], r[2 * i + 1]); }
#pragma unroll
    for (int n = 4; n >= 1; n >>= 1) {
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const float nl = fminf(l[2 * i], l[2 * i + 1]);
            const float nh = fmin3(fmaxf(l[2 * i], l[2 * i + 1]), h[2 * i], h[2 * i + 1]);
            l[i] = nl; h[i] = nh;
        }
    }
    lo = l[0]; hi = h[0];
}
// index (0..15) of the smallest of 16 values: the index rides in the 4 low mantissa bits through a min tree.  Values that
// differ only in those bits may resolve to either index -- such rows have a runner-up gap of <= 16 ulp and are never
// certified (the certificate needs more than cB (||x||^2 + ||e||^2) >> 16 ulp of any score).  Unused classes hold +inf,
// which the tag turns into NaN: fminf ignores them.
__device__ __forceinline__ int argmin16(const float (&r)[16]) {
    float t[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = __uint_as_float((__float_as_uint(r[i]) & 0xFFFFFFF0u) | (uint32_t)i);
    return (int)(__float_as_uint(min16(t)) & 15u);
}

// after the warpgroup's (up to) 256 virtual columns went through scan_buffer: winning virtual column, its score m1
// and the runner-up score m2 (m2 == m1 on exact ties; NaN / inf rows are never certified)
__device__ __forceinline__ int scan_finish(const float (&rA)[16], const float (&rB)[16], float& m1, float& m2) {
    float loA, hiA, loB, hiB;
    two_smallest16(rA, loA, hiA);
    two_smallest16(rB, loB, hiB);
    m1 = fminf(loA, loB);
    m2 = fminf(hiA, hiB);
    return 16 * argmin16(rB) + argmin16(rA);
}

// ---------------------------------------------------------------------------------------------------
// MMA issue.  The issuer's own instruction stream is on the critical path (one thread, dependent uniform-datapath
// instructions of ~10 cycles each), so all MMAs of one accumulator unit and the commit go out in ONE asm block,
// executed by the converged warp with the tcgen05 instructions predicated on elect.sync (inside a divergent
// `if (lane == 0)` the compiler wraps every UTCHMMA in an elect-and-retry loop), and the shared-memory descriptors are
// assembled from 32-bit halves: lo = (address >> 4) | LBO, hi = constant per swizzle mode.
//   unit = misc.misc (overwrite)  +  4 x xh.eh  [+ 4 x xl.eh + 4 x xh.el when NSPLIT == 3]   (K = 16 per MMA)
// ---------------------------------------------------------------------------------------------------
constexpr uint32_t DESC_HI_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t DESC_HI_SW32 = (256u >> 4) | (1u << 14) | (6u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }

#define VQ_MMA_OP(CG) "@pe tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %7, "
#define VQ_MMA_KS(CG, A, B, KS)                                                              \
    "add.u32 ta, " A ", " KS ";\n\tadd.u32 tb, " B ", " KS ";\n\t"                            \
    "mov.b64 da, {ta, %6};\n\tmov.b64 db, {tb, %6};\n\t" VQ_MMA_OP(CG) "pt;\n\t"
#define VQ_MMA_BLOCK4(CG, A, B) VQ_MMA_KS(CG, A, B, "0") VQ_MMA_KS(CG, A, B, "2") VQ_MMA_KS(CG, A, B, "4") VQ_MMA_KS(CG, A, B, "6")
#define VQ_MMA_HEAD(CG)                                                                      \
    "{\n\t.reg .pred pf, pt, pe;\n\t.reg .b64 da, db;\n\t.reg .b32 ta, tb;\n\t"               \
    "elect.sync _|pe, 0xffffffff;\n\t"                                                       \
    "setp.ne.b32 pf, %7, %7;\n\tsetp.eq.b32 pt, %7, %7;\n\t"                                  \
    "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t" VQ_MMA_OP(CG) "pf;\n\t"
#define VQ_COMMIT1 "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t}"
#define VQ_COMMIT2 "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%8], %11;\n\t}"

template <bool CTA2>
__device__ __forceinline__ void commit_elected(uint32_t bar) {
    if constexpr (CTA2)
        asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
                     "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
                     :: "r"(bar), "h"((uint16_t)3) : "memory");
    else
        asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
                     "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
                     :: "r"(bar) : "memory");
}

template <int NSPLIT, bool CTA2>
__device__ __forceinline__ void issue_unit(uint32_t d_tmem, uint32_t am_lo, uint32_t bm_lo, uint32_t a_lo, uint32_t b_lo,
                                           uint32_t al_lo, uint32_t bl_lo, uint32_t bar_tf) {
    const uint32_t idesc = CTA2 ? IDESC2 : IDESC;
    if constexpr (NSPLIT == 3 && CTA2) {
        asm volatile(VQ_MMA_HEAD("2") VQ_MMA_BLOCK4("2", "%3", "%4") VQ_MMA_BLOCK4("2", "%9", "%4") VQ_MMA_BLOCK4("2", "%3", "%10") VQ_COMMIT2
                     :: "r"(d_tmem), "r"(am_lo), "r"(bm_lo), "r"(a_lo), "r"(b_lo), "r"(DESC_HI_SW32), "r"(DESC_HI_SW128), "r"(idesc),
                        "r"(bar_tf), "r"(al_lo), "r"(bl_lo), "h"((uint16_t)3) : "memory");
    } else if constexpr (NSPLIT == 3) {
        asm volatile(VQ_MMA_HEAD("1") VQ_MMA_BLOCK4("1", "%3", "%4") VQ_MMA_BLOCK4("1", "%9", "%4") VQ_MMA_BLOCK4("1", "%3", "%10") VQ_COMMIT1
                     :: "r"(d_tmem), "r"(am_lo), "r"(bm_lo), "r"(a_lo), "r"(b_lo), "r"(DESC_HI_SW32), "r"(DESC_HI_SW128), "r"(idesc),
                        "r"(bar_tf), "r"(al_lo), "r"(bl_lo) : "memory");
    } else if constexpr (CTA2) {
        asm volatile(VQ_MMA_HEAD("2") VQ_MMA_BLOCK4("2", "%3", "%4") VQ_COMMIT2
                     :: "r"(d_tmem), "r"(am_lo), "r"(bm_lo), "r"(a_lo), "r"(b_lo), "r"(DESC_HI_SW32), "r"(DESC_HI_SW128), "r"(idesc),
                        "r"(bar_tf), "r"(al_lo), "r"(bl_lo), "h"((uint16_t)3) : "memory");
    } else {
        asm volatile(VQ_MMA_HEAD("1") VQ_MMA_BLOCK4("1", "%3", "%4") VQ_COMMIT1
                     :: "r"(d_tmem), "r"(am_lo), "r"(bm_lo), "r"(a_lo), "r"(b_lo), "r"(DESC_HI_SW32), "r"(DESC_HI_SW128), "r"(idesc),
                        "r"(bar_tf) : "memory");
    }
}

// ---- kind::tf32 issue (NSPLIT == 0): x is multiplied straight from its fp32 TMA stage ---------------------------------
// One accumulator unit = 8 MMAs (M128 x N128 x K8, 32 bytes of K per step) over the two 32-dim k-blocks of x and of the
// -2e image, the first overwriting the accumulator, and LAST the misc MMA (bias, row offset, error bound: it needs the row
// norms, which the norm warps compute from the same stage while the products are already running) + the commit.
//   a_lo / b_lo: descriptor low words of k-block 0; a_kb / b_kb: descriptor-unit (16-byte) offset of k-block 1;
//   a_ks: descriptor-unit step per K = 8 (2 for a K-major stage, 64 for the MN-major x^T stage: 8 dims x 128 bytes)
__device__ __forceinline__ void issue_unit_tf_products(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t a_kb, uint32_t b_kb,
                                                       uint32_t a_ks, uint32_t a_hi, uint32_t idesc) {
#define VQ_TF_OP "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %7, "
#define VQ_TF_NEXT "add.u32 ta, ta, %6;\n\tadd.u32 tb, tb, 2;\n\tmov.b64 da, {ta, %8};\n\tmov.b64 db, {tb, %3};\n\t" VQ_TF_OP "pt;\n\t"
    asm volatile(
        "{\n\t.reg .pred pf, pt, pe;\n\t.reg .b64 da, db;\n\t.reg .b32 ta, tb;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "setp.ne.b32 pf, %7, %7;\n\tsetp.eq.b32 pt, %7, %7;\n\t"
        "mov.u32 ta, %1;\n\tmov.u32 tb, %2;\n\t"
        "mov.b64 da, {ta, %8};\n\tmov.b64 db, {tb, %3};\n\t" VQ_TF_OP "pf;\n\t"
        VQ_TF_NEXT VQ_TF_NEXT VQ_TF_NEXT
        "add.u32 ta, %1, %4;\n\tadd.u32 tb, %2, %5;\n\t"
        "mov.b64 da, {ta, %8};\n\tmov.b64 db, {tb, %3};\n\t" VQ_TF_OP "pt;\n\t"
        VQ_TF_NEXT VQ_TF_NEXT VQ_TF_NEXT
        "}"
        :: "r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(DESC_HI_SW128), "r"(a_kb), "r"(b_kb), "r"(a_ks), "r"(idesc), "r"(a_hi) : "memory");
#undef VQ_TF_NEXT
#undef VQ_TF_OP
}
__device__ __forceinline__ void issue_unit_tf_misc(uint32_t d_tmem, uint32_t am_lo, uint32_t bm_lo, uint32_t bar_tf) {
    asm volatile(
        "{\n\t.reg .pred pt, pe;\n\t.reg .b64 da, db;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "setp.eq.b32 pt, %4, %4;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, pt;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t}"
        :: "r"(d_tmem), "r"(am_lo), "r"(bm_lo), "r"(DESC_HI_SW32), "r"(IDESC_TF), "r"(bar_tf) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// shared-memory plan
// ---------------------------------------------------------------------------------------------------
template <int NSPLIT, int AS, int XS, bool CTA2 = false>
struct Plan {
    // NSPLIT == 0: tf32 filter -- no converted A operand (the MMAs read the fp32 x stage), an A stage is the misc rows only
    static constexpr bool TF = NSPLIT == 0;
    static_assert(!(TF && CTA2), "the tf32 filter has no CTA-pair variant");
    static constexpr uint32_t A_STAGE = TF ? 4096u : (NSPLIT == 3 ? 2u : 1u) * 16384u + 4096u;
    static constexpr uint32_t X_STAGE = TILE_M * TC_D * 4;
    static constexpr uint32_t BDIV = CTA2 ? 2u : 1u;          // a CTA of a pair holds half of the codes of every unit
    __host__ __device__ static uint32_t b_bytes(int K) { return ((uint32_t)K * ((NSPLIT == 3 || TF) ? 256u : 128u) + (uint32_t)K * 32u) / BDIV; }
    __host__ __device__ static uint32_t off_b_lo(int K) { return (uint32_t)K * 128u / BDIV; }      // split: el block; tf32: k-block 1
    __host__ __device__ static uint32_t off_b_misc(int K) { return (uint32_t)K * ((NSPLIT == 3 || TF) ? 256u : 128u) / BDIV; }
    __host__ __device__ static uint32_t off_a(int K) { return b_bytes(K); }
    __host__ __device__ static uint32_t off_x(int K) { return off_a(K) + AS * A_STAGE; }
    __host__ __device__ static uint32_t off_small(int K) { return off_x(K) + XS * X_STAGE; }
    // small region: enorm[K] | rownorm[NORM_RING][128] | codes[RES_RING][128] | partials[2][128] | barriers | tmem ptr
    __host__ __device__ static uint32_t off_rownorm(int K) { return off_small(K) + (uint32_t)K * 4u; }
    __host__ __device__ static uint32_t off_codes(int K) { return off_rownorm(K) + NORM_RING * TILE_M * 4u; }
    __host__ __device__ static uint32_t off_parts(int K) { return off_codes(K) + RES_RING * TILE_M * 4u; }
    __host__ __device__ static uint32_t off_bars(int K) { return off_parts(K) + 2u * TILE_M * 16u; }
    __host__ __device__ static uint32_t total(int K) { return off_bars(K) + 512u + 1024u /* base alignment slack */; }
};

struct Params {
    alignas(64) CUtensorMap tmap;    // NCHW only: x as a 3-D tensor [image][dim][row-in-image], box = 128 rows x 64 dims
    const float* x;
    int64_t n_rows;
    int K;
    const unsigned char* image;      // tensor-core operand image
    const float* cbT;                // [K][64] fp32
    float* quantize;                 // may be null
    int64_t* embed_ind;
    double* diff_acc;                // may be null
    float* stat_sums;                // may be null
    float* stat_counts;
    int* flagged_count;
    int* flagged_rows;
    float* dbg_scores;               // optional [n_rows][K] dump of the tensor-core scores
    unsigned long long* prof;        // optional [grid][PROF_SLOTS] cycle counters (pipeline bubble analysis)
    int dbg_skip;                    // DBG builds only: bit0 skip output work, bit1 skip conversion, bit2 skip scan, bit3 skip MMAs
    float cA, cB;
    // codebooks larger than the resident operand image (K > 512): one launch per 512-code slice.  Scores of a row are
    // comparable across slices (same row offset, per-code error terms), so the running (m1, m2, winner, ||e_winner||) is
    // carried in `partial` (one float4 per row); only the last slice certifies and writes outputs.
    // NCHW-physical input (kernel template flag NCHW): element (n, d) of x / quantize at
    //   (n / rpi) * img_stride + d * col_stride + (n % rpi), rpi % 128 == 0 (a tile never straddles two images)
    int64_t rpi, img_stride, col_stride;
    float* x_dense;                  // NCHW only, may be null: dense [N][64] copy of x for the statistics kernel
    int code_base;                   // global index of this launch's first code
    float4* partial;                 // may be null (single launch)
    int pass_first, pass_last;
};
// profile slots (cycles, summed over the CTA's tiles; one recording lane per role)
enum ProfSlot { PF_PROD_WAIT_XE = 0, PF_MMA_WAIT_AF, PF_MMA_WAIT_TE, PF_MMA_TOTAL, PF_CONV_WAIT_XF, PF_CONV_WAIT_AE,
                PF_CONV_TOTAL, PF_EPI0_WAIT_TF, PF_EPI0_SCAN, PF_EPI0_TOTAL, PF_EPI1_WAIT_TF, PF_EPI1_SCAN, PF_EPI1_WAIT_PF,
                PF_EPI1_WAIT_RE, PF_EPI1_TOTAL, PF_OUT_WAIT_RF, PF_OUT_TOTAL, PF_KERNEL, PF_CONV_LOOP, PF_CONV_TAIL, PF_CONV_FENCE,
                PROF_SLOTS = 24 };

enum BarId { BAR_B = 0, BAR_XF = 1, BAR_XE = 5, BAR_AF = 9, BAR_AE = 13, BAR_TF = 17, BAR_TE = 21, BAR_RF = 25, BAR_RE = 27,
             BAR_PF = 29, BAR_PE = 31, BAR_PB = 33, BAR_COUNT = 34 };

// CTA2 = true: launched as clusters of two CTAs (one SM pair).  Each CTA converts, scans and writes its own 128-row
// tile, but the pair shares ONE tcgen05.mma.cta_group::2 stream (M = 256) issued by the leader CTA, and each CTA
// holds only half of the codebook operand image (the B rows of its half of every 256-code unit).  That halves the
// image's shared-memory footprint (144 KB -> 72 KB at K = 512), which is what pays for double-buffered A and x
// stages, and halves the operand bandwidth each SM's tensor core draws from shared memory.
template <int NSPLIT, int AS, int XS, bool DBG, bool CTA2, bool NCHW = false>
__global__ void __launch_bounds__(THREADS, 1) k_vq_tc(const __grid_constant__ Params p) {
    using P = Plan<NSPLIT, AS, XS, CTA2>;
    constexpr bool TF = NSPLIT == 0;              // tf32 filter: MMAs read the fp32 x stage, "converters" only compute row norms
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    const int K = p.K;
    const int U = K / UNIT_N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const uint32_t sB = base, sA = base + P::off_a(K), sX = base + P::off_x(K);
    float* enorm_s = reinterpret_cast<float*>(sm + P::off_small(K));
    float* rownorm_s = reinterpret_cast<float*>(sm + P::off_rownorm(K));
    int* codes_s = reinterpret_cast<int*>(sm + P::off_codes(K));
    float4* part_s = reinterpret_cast<float4*>(sm + P::off_parts(K));
    const uint32_t bars = base + P::off_bars(K);
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(sm + P::off_bars(K) + BAR_COUNT * 8);
    auto bar = [&](int id) { return bars + 8u * (uint32_t)id; };

    pdl_wait();                                   // codebook image / previous step's outputs complete (no-op without PDL)
    pdl_trigger();
    const int64_t n_tiles = (p.n_rows + TILE_M - 1) / TILE_M;
    const uint32_t crank = CTA2 ? cluster_ctarank() : 0u;
    // both CTAs of a pair run the same number of trips (the pair's second tile may lie past the end on the last one)
    const int64_t first_tile = CTA2 ? (int64_t)(blockIdx.x & ~1u) : (int64_t)blockIdx.x;
    const uint32_t n_iter = first_tile < n_tiles ? (uint32_t)((n_tiles - first_tile + gridDim.x - 1) / gridDim.x) : 0u;
    constexpr uint32_t PAIR_THREADS = CTA2 ? 8u : 4u;         // arriving WARPS on the barriers the MMA issuer waits on
    unsigned long long* prof = (DBG && p.prof) ? p.prof + (size_t)blockIdx.x * PROF_SLOTS : nullptr;
    const long long t_kernel0 = clock64();
    // timed barrier wait: accumulates the stall into a REGISTER counter when profiling is on (flushed to
    // global memory once per role, so the instrumentation does not add memory round trips to the pipeline)
    long long pacc[6] = {0, 0, 0, 0, 0, 0};
    auto wait_t = [&](int id, uint32_t parity, int acc, bool rec) {
        if (DBG && prof && rec) {
            const long long t0 = clock64();
            mbar_wait(bar(id), parity);
            pacc[acc] += clock64() - t0;
        } else {
            mbar_wait(bar(id), parity);
        }
    };
    auto wait_r = [&](int id, uint32_t parity, int acc, bool rec) {     // relaxed polling (roles with slack)
        if (DBG && prof && rec) {
            const long long t0 = clock64();
            mbar_wait_relaxed<RELAX_NS>(bar(id), parity);
            pacc[acc] += clock64() - t0;
        } else {
            mbar_wait_relaxed<RELAX_NS>(bar(id), parity);
        }
    };
    auto flush = [&](int slot, int acc) { prof[slot] = (unsigned long long)pacc[acc]; };

    // ---- one-time setup -----------------------------------------------------------------------------
    if (threadIdx.x == 0) {
        mbar_init(bar(BAR_B), 1);
        // x stage released by: the converters (+ the output warps, NCHW: they read x from the stage); tf32: the commit behind the tile's last MMA
        for (int s = 0; s < XS; ++s) { mbar_init(bar(BAR_XF + s), 1); mbar_init(bar(BAR_XE + s), TF ? 1 : (NCHW ? 12 : 4)); }
        for (int s = 0; s < AS; ++s) { mbar_init(bar(BAR_AF + s), PAIR_THREADS); mbar_init(bar(BAR_AE + s), 1); }
        for (int s = 0; s < NBUF; ++s) { mbar_init(bar(BAR_TF + s), 1); mbar_init(bar(BAR_TE + s), PAIR_THREADS); }
        mbar_init(bar(BAR_PB), 1);
        for (int s = 0; s < RES_RING; ++s) { mbar_init(bar(BAR_RF + s), 4); mbar_init(bar(BAR_RE + s), 8); }
        for (int s = 0; s < 2; ++s) { mbar_init(bar(BAR_PF + s), 4); mbar_init(bar(BAR_PE + s), 4); }
        fence_barrier_init();
    }
    if (warp == W_MMA) { if (CTA2) tmem_alloc2(smem_u32(tmem_ptr_s), 512); else tmem_alloc(smem_u32(tmem_ptr_s), 512); }
    // constant part of the A "misc" blocks: [1 1 1 | o1 o2 o3 | nx 1 | 0 x8]; zero chunk 1 once
    for (int i = threadIdx.x; i < AS * TILE_M; i += THREADS) {
        int s = i / TILE_M, r = i % TILE_M;
        unsigned char* m = sm + P::off_a(K) + s * P::A_STAGE + (P::A_STAGE - 4096u);
        *reinterpret_cast<uint4*>(m + sw32_chunk_off((uint32_t)r, 1)) = make_uint4(0, 0, 0, 0);
    }
    fence_async_smem();
    int bad = 0;
    for (int i = threadIdx.x; i < K; i += THREADS) {
        const float ne = reinterpret_cast<const float*>(p.image + image_off_enorm(K))[i];
        enorm_s[i] = ne;
        bad |= !(ne < 1.0e18f);                   // NaN / inf / absurd norms: certify nothing, exact path decides
    }
    tc_fence_before();
    const bool cb_bad = __syncthreads_or(bad) != 0;
    if (CTA2) cluster_sync_all();                 // the peer's barriers exist before anybody arrives on them remotely
    tc_fence_after();
    // barriers the (leader's) MMA issuer waits on live in the leader CTA; both CTAs arrive there
    auto arrive_mma_side = [&](int id) {
        if (CTA2) mbar_arrive_cluster(mapa_rank(bar(id), 0)); else mbar_arrive(bar(id));
    };
    const uint32_t tmem_base = *tmem_ptr_s;

    if (warp == W_PROD) {
        // ================= bulk-copy producer =========================================================
        reg_dec<24>();
        if (lane == 0) {
            const uint32_t bbytes = P::b_bytes(K);
            mbar_expect_tx(bar(BAR_B), bbytes);
            // per 256-code unit: the rows this CTA feeds to the MMA (all 256, or its half of them in a pair)
            constexpr uint32_t UROWS = UNIT_N / P::BDIV;
            for (int u = 0; u < U; ++u) {
                const size_t krow = (size_t)u * UNIT_N + (size_t)crank * UROWS;
                if (TF) {                        // two 32-dim k-blocks of tf32 rows + the tf32 misc rows
                    bulk_g2s(sB + (uint32_t)u * UROWS * 128u, p.image + image_off_tf(K) + krow * 128, UROWS * 128u, bar(BAR_B));
                    bulk_g2s(sB + P::off_b_lo(K) + (uint32_t)u * UROWS * 128u, p.image + image_off_tf(K) + (size_t)K * 128 + krow * 128, UROWS * 128u, bar(BAR_B));
                    bulk_g2s(sB + P::off_b_misc(K) + (uint32_t)u * UROWS * 32u, p.image + image_off_tfmisc(K) + krow * 32, UROWS * 32u, bar(BAR_B));
                    continue;
                }
                bulk_g2s(sB + (uint32_t)u * UROWS * 128u, p.image + krow * 128, UROWS * 128u, bar(BAR_B));
                if (NSPLIT == 3)
                    bulk_g2s(sB + P::off_b_lo(K) + (uint32_t)u * UROWS * 128u, p.image + image_off_lo(K) + krow * 128, UROWS * 128u, bar(BAR_B));
                bulk_g2s(sB + P::off_b_misc(K) + (uint32_t)u * UROWS * 32u, p.image + (NSPLIT == 3 ? image_off_misc(K) : image_off_misc1(K)) + krow * 32, UROWS * 32u, bar(BAR_B));
            }
            const uint64_t keep = l2_policy_evict_last();
            for (uint32_t it = 0; it < n_iter; ++it) {
                const int64_t t = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
                const uint32_t s = it % XS, ph = (it / XS) & 1u;
                wait_r(BAR_XE + s, ph ^ 1u, 0, true);
                const int64_t r0 = t * TILE_M;
                const uint32_t rows = (uint32_t)max((int64_t)0, min((int64_t)TILE_M, p.n_rows - r0));
                const uint32_t bytes = TF ? (rows ? P::X_STAGE : 0u) : rows * TC_D * 4u;     // tensor-map boxes: rows past the end are zero-filled and counted
                mbar_expect_tx(bar(BAR_XF + s), bytes);
                if (TF && !NCHW) {               // two boxes of [128 rows][32 dims], 128-byte swizzle: the K-major tf32 A operand as it is
                    if (bytes) {
                        tma_load_2d_hint(sX + s * P::X_STAGE, &p.tmap, 0, (int)r0, bar(BAR_XF + s), keep);
                        tma_load_2d_hint(sX + s * P::X_STAGE + 16384u, &p.tmap, 32, (int)r0, bar(BAR_XF + s), keep);
                    }
                } else
                if (NCHW) {                      // one tiled TMA load: box [64 dims][128 rows] -> the stage holds x^T [d][row]
                    // (64 separate 512-byte bulk copies cost ~140 cycles each in the TMA unit: 126 us per launch)
                    if (bytes) tma_load_3d(sX + s * P::X_STAGE, &p.tmap, (int)(r0 % p.rpi), 0, (int)(r0 / p.rpi), bar(BAR_XF + s), keep);
                } else if (bytes) bulk_g2s_hint(sX + s * P::X_STAGE, p.x + r0 * TC_D, bytes, bar(BAR_XF + s), keep);
            }
            if (DBG && prof) flush(PF_PROD_WAIT_XE, 0);
        }
    } else if (warp == W_MMA) {
        // ================= MMA issuer (converged warp, tcgen05 instructions predicated on one elected lane) ===========
        reg_dec<24>();
        mbar_wait(bar(BAR_B), 0);
        if (CTA2) {
            if (crank != 0) { if (lane == 0) mbar_arrive_cluster(mapa_rank(bar(BAR_PB), 0)); }   // my half of B has landed
            else mbar_wait_cluster(bar(BAR_PB), 0);
        }
        const long long t_role0 = clock64();
        const bool mrec = lane == 0;
        constexpr uint32_t UROWS = UNIT_N / P::BDIV;
        constexpr uint32_t B_STEP = (UROWS * 128u) >> 4, BM_STEP = (UROWS * 32u) >> 4;   // per unit, in descriptor units
        const uint32_t b_lo0 = desc_lo(sB), bl_lo0 = desc_lo(sB + P::off_b_lo(K)), bm_lo0 = desc_lo(sB + P::off_b_misc(K));
        for (uint32_t it = 0; it < (crank == 0 ? n_iter : 0u); ++it) {     // the leader issues for the pair
            const uint32_t sa = it % AS, pha = (it / AS) & 1u;
            if constexpr (TF) {
                // products straight from the x stage as soon as it lands; the misc MMA of the first unit waits for the row norms
                const uint32_t sx = it % XS, phx = (it / XS) & 1u;
                wait_t(BAR_XF + sx, phx, 0, mrec);
                tc_fence_after();
                const uint32_t x_lo = desc_lo(sX + sx * P::X_STAGE), am_lo = desc_lo(sA + sa * P::A_STAGE);
                for (int u = 0; u < U; ++u) {
                    const uint32_t uc = it * (uint32_t)U + (uint32_t)u;
                    const uint32_t buf = uc % NBUF, pht = (uc / NBUF) & 1u;
                    wait_t(BAR_TE + buf, pht ^ 1u, 1, mrec);
                    tc_fence_after();
                    issue_unit_tf_products(tmem_base + buf * UNIT_N, x_lo, b_lo0 + (uint32_t)u * B_STEP, 16384u >> 4,
                                           P::off_b_lo(K) >> 4, 2u, DESC_HI_SW128, IDESC_TF);
                    if (u == 0) { wait_t(BAR_AF + sa, pha, 0, mrec); tc_fence_after(); }
                    issue_unit_tf_misc(tmem_base + buf * UNIT_N, am_lo, bm_lo0 + (uint32_t)u * BM_STEP, bar(BAR_TF + buf));
                }
                commit_elected<false>(bar(BAR_XE + sx));      // all MMAs that read this x stage / these misc rows are done
                commit_elected<false>(bar(BAR_AE + sa));
                continue;
            }
            if (CTA2) mbar_wait_cluster(bar(BAR_AF + sa), pha); else wait_t(BAR_AF + sa, pha, 0, mrec);
            tc_fence_after();
            const uint32_t a0 = sA + sa * P::A_STAGE;
            const uint32_t a_lo = desc_lo(a0), al_lo = desc_lo(a0 + 16384u), am_lo = desc_lo(a0 + (P::A_STAGE - 4096u));
            for (int u = 0; u < U; ++u) {
                const uint32_t uc = it * (uint32_t)U + (uint32_t)u;
                const uint32_t buf = uc % NBUF, pht = (uc / NBUF) & 1u;
                if (CTA2) mbar_wait_cluster(bar(BAR_TE + buf), pht ^ 1u); else wait_t(BAR_TE + buf, pht ^ 1u, 1, mrec);
                tc_fence_after();
                if (DBG && (p.dbg_skip & 8)) {
                    if (lane == 0) { if (CTA2) umma_commit_2cta(bar(BAR_TF + buf)); else umma_commit(bar(BAR_TF + buf)); }
                    __syncwarp();
                } else {
                    // misc block first (bias, offset, error bound: overwrites the accumulator), then the products
                    if constexpr (!TF)
                    issue_unit<NSPLIT, CTA2>(tmem_base + buf * UNIT_N, am_lo, bm_lo0 + (uint32_t)u * BM_STEP, a_lo,
                                             b_lo0 + (uint32_t)u * B_STEP, al_lo, bl_lo0 + (uint32_t)u * B_STEP, bar(BAR_TF + buf));
                }
            }
            // the elected lane of issue_unit and lane 0 may differ: tcgen05.commit tracks the MMAs of the executing
            // thread, so the stage release is committed by an elected lane as well
            commit_elected<CTA2>(bar(BAR_AE + sa));
        }
        if (DBG && prof && mrec) { flush(PF_MMA_WAIT_AF, 0); flush(PF_MMA_WAIT_TE, 1); prof[PF_MMA_TOTAL] = (unsigned long long)(clock64() - t_role0); }
    } else if (warp > W_MMA) {
        reg_dec<24>();                           // spare warps of the producer/MMA warpgroup
    } else if (warp >= W_CONV) {
        // ================= converters: fp32 rows -> split-bf16 K-major operand ==========================
        reg_dec<56>();
        const int cw = warp - W_CONV;            // rows cw*32 .. cw*32+31
        const int half = lane >> 4, q4 = lane & 15;
        const bool rec = (warp == W_CONV && lane == 0);
        const long long t_role0 = clock64();
        for (uint32_t it = 0; it < n_iter; ++it) {
            const int64_t t = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
            const uint32_t sx = it % XS, phx = (it / XS) & 1u, sa = it % AS, pha = (it / AS) & 1u;
            wait_r(BAR_XF + sx, phx, 0, rec);
            wait_r(BAR_AE + sa, pha ^ 1u, 1, rec);
            const unsigned char* xs = sm + P::off_x(K) + sx * P::X_STAGE;
            unsigned char* ah = sm + P::off_a(K) + sa * P::A_STAGE;
            unsigned char* al = ah + 16384u;
            unsigned char* am = ah + (P::A_STAGE - 4096u);
            const long long tc0 = (DBG && prof && rec) ? clock64() : 0;
            float my_sq = 1.f;
            if constexpr (TF && !NCHW) {
                // tf32: nothing to convert -- thread = row reads its 256 bytes from the two swizzled k-blocks (a quarter warp
                // covers 8 rows x 16 bytes at 8 distinct chunk positions: conflict-free) for the row norm only
                const int r = cw * 32 + lane;
                const unsigned char* xr = xs + r * 128;
                const uint32_t sw = (uint32_t)r & 7u;
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 v = *reinterpret_cast<const float4*>(xr + kb * 16384 + (((uint32_t)c ^ sw) << 4));
                        s0 = fmaf(v.x, v.x, s0); s1 = fmaf(v.y, v.y, s1); s2 = fmaf(v.z, v.z, s2); s3 = fmaf(v.w, v.w, s3);
                    }
                my_sq = (s0 + s1) + (s2 + s3);
            } else
            if (NCHW) {
                // the stage holds x^T [d][row]: thread = row (conflict-free 128-byte reads per dim), all 64 components pass
                // through this thread, so the row norm needs no shuffles
                if (!(DBG && (p.dbg_skip & 2))) {
                    const int r = cw * 32 + lane;
                    const float* xr = reinterpret_cast<const float*>(xs) + r;
                    float4* xd = p.x_dense ? reinterpret_cast<float4*>(p.x_dense + (t * TILE_M + r) * TC_D) : nullptr;
                    float sq = 0.f;
#pragma unroll 2
                    for (int c = 0; c < 8; ++c) {         // 8 dims -> one 16-byte chunk of the K-major operand row
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = xr[(c * 8 + j) * TILE_M];
                        const uint32_t p01 = pack_bf16(v[0], v[1]), p23 = pack_bf16(v[2], v[3]);
                        const uint32_t p45 = pack_bf16(v[4], v[5]), p67 = pack_bf16(v[6], v[7]);
                        const uint32_t off = sw128_off((uint32_t)r, (uint32_t)c * 8);
                        *reinterpret_cast<uint4*>(ah + off) = make_uint4(p01, p23, p45, p67);
                        if (NSPLIT == 3) {
                            const uint32_t pk[4] = {p01, p23, p45, p67};
                            uint32_t lo[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                lo[j] = pack_bf16(v[2 * j] - __uint_as_float(pk[j] << 16), v[2 * j + 1] - __uint_as_float(pk[j] & 0xFFFF0000u));
                            *reinterpret_cast<uint4*>(al + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) sq = fmaf(v[j], v[j], sq);
                        if (xd) stg_v8(reinterpret_cast<float*>(xd + 2 * c), v);   // dense copy for the statistics kernel: one full sector per lane
                    }
                    my_sq = sq;
                }
            } else
            if (!(DBG && (p.dbg_skip & 2))) {
#pragma unroll 1
                for (int g4 = 0; g4 < 4; ++g4) {          // 4 row pairs per trip
                    float4 v[4];
                    float sq[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        v[u] = *reinterpret_cast<const float4*>(xs + (cw * 32 + 2 * (4 * g4 + u) + half) * 256 + q4 * 16);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = cw * 32 + 2 * (4 * g4 + u) + half;
                        const uint32_t p01 = pack_bf16(v[u].x, v[u].y), p23 = pack_bf16(v[u].z, v[u].w);
                        const uint32_t off = sw128_off((uint32_t)r, (uint32_t)q4 * 4);
                        *reinterpret_cast<uint2*>(ah + off) = make_uint2(p01, p23);
                        if (NSPLIT == 3) {
                            const float h0 = __uint_as_float(p01 << 16), h1 = __uint_as_float(p01 & 0xFFFF0000u);
                            const float h2 = __uint_as_float(p23 << 16), h3 = __uint_as_float(p23 & 0xFFFF0000u);
                            *reinterpret_cast<uint2*>(al + off) =
                                make_uint2(pack_bf16(v[u].x - h0, v[u].y - h1), pack_bf16(v[u].z - h2, v[u].w - h3));
                        }
                        sq[u] = fmaf(v[u].x, v[u].x, fmaf(v[u].y, v[u].y, fmaf(v[u].z, v[u].z, v[u].w * v[u].w)));
                    }
                    {
                        const bool up = (q4 & 2) != 0;
                        const float s0 = up ? sq[0] : sq[2], k0 = up ? sq[2] : sq[0];
                        const float s1 = up ? sq[1] : sq[3], k1 = up ? sq[3] : sq[1];
                        sq[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, 2);
                        sq[1] = k1 + __shfl_xor_sync(0xffffffffu, s1, 2);
                        const bool up1 = (q4 & 1) != 0;
                        const float s2 = up1 ? sq[0] : sq[1], k2 = up1 ? sq[1] : sq[0];
                        float tot = k2 + __shfl_xor_sync(0xffffffffu, s2, 1);
                        tot += __shfl_xor_sync(0xffffffffu, tot, 4);
                        tot += __shfl_xor_sync(0xffffffffu, tot, 8);
                        if ((q4 >> 2) == g4) my_sq = tot;
                    }
                }
            }
            const long long tc1 = (DBG && prof && rec) ? clock64() : 0;
            if constexpr (TF) {
                const int r = cw * 32 + lane;
                const float nx = sqrtf(my_sq);
                float o1, o2, o3;
                split3_tf32(my_sq * 1.001953125f, o1, o2, o3);          // off_i = ||x||^2 (1 + 2^-9)
                const float nxu = tf32_rn(nx * 1.001953125f);           // ||x|| rounded up
                *reinterpret_cast<float4*>(am + sw32_chunk_off((uint32_t)r, 0)) = make_float4(1.f, 1.f, 1.f, o1);
                *reinterpret_cast<float4*>(am + sw32_chunk_off((uint32_t)r, 1)) = make_float4(o2, o3, nxu, 1.f);
                rownorm_s[(it % NORM_RING) * TILE_M + r] = nx;
            } else {
                const int r = NCHW ? cw * 32 + lane : cw * 32 + 2 * q4 + half;
                const float nx = sqrtf(my_sq);
                float o1, o2, o3;
                split3(my_sq * 1.001953125f, o1, o2, o3);            // off_i = ||x||^2 (1 + 2^-9)
                const float nxu = bf16_round(nx * 1.0078125f);        // ||x|| rounded up
                *reinterpret_cast<uint4*>(am + sw32_chunk_off((uint32_t)r, 0)) =
                    make_uint4(pack_bf16(1.f, 1.f), pack_bf16(1.f, o1), pack_bf16(o2, o3), pack_bf16(nxu, 1.f));
                rownorm_s[(it % NORM_RING) * TILE_M + r] = nx;
            }
            const long long tc2 = (DBG && prof && rec) ? clock64() : 0;
            fence_async_smem();
            __syncwarp();
            if (lane == 0) { arrive_mma_side(BAR_AF + sa); if (!TF) mbar_arrive(bar(BAR_XE + sx)); }
            if (DBG && prof && rec) {
                const long long tc3 = clock64();
                pacc[2] += tc1 - tc0; pacc[3] += tc2 - tc1; pacc[4] += tc3 - tc2;
            }
        }
        if (DBG && prof && rec) { flush(PF_CONV_WAIT_XF, 0); flush(PF_CONV_WAIT_AE, 1); flush(PF_CONV_LOOP, 2); flush(PF_CONV_TAIL, 3); flush(PF_CONV_FENCE, 4); prof[PF_CONV_TOTAL] = (unsigned long long)(clock64() - t_role0); }
    } else if (warp < W_OUT) {
        // ================= epilogue: TMEM -> two-class min scan -> certified arg-min =====================
        // The codes of a tile come as U = K/128 accumulator units of 128 columns; group g takes units g, g+2 (its
        // "256 virtual columns"), the two groups work on the same tile concurrently and group 1 merges.
        reg_inc<128>();
        const int g = warp >> 2;
        const int wq = warp & 3;                 // TMEM lane quarter this warp may access
        const int row_in_tile = wq * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16);
        const bool rec = ((warp & 3) == 0 && lane == 0);      // first warp of each group
        const int pbase = g ? PF_EPI1_WAIT_TF : PF_EPI0_WAIT_TF;
        const long long t_role0 = clock64();
        // units u = g, g+2, ... of every tile belong to this group (U = K/128 units per tile: 2 or 4); unit counter
        // uc = it*U + u selects TMEM buffer uc % 4 and its phase
        for (uint32_t it = 0; it < n_iter; ++it) {
            const int64_t t = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
            const int64_t grow = t * TILE_M + row_in_tile;
            float rA[16], rB[16];
#pragma unroll
            for (int a = 0; a < 16; ++a) { rA[a] = INFINITY; rB[a] = INFINITY; }
            float* dbg = (DBG && p.dbg_scores && grow < p.n_rows) ? p.dbg_scores + grow * K : nullptr;
            const long long t_scan0 = clock64();
            {   // first unit of the group
                const uint32_t uc = it * (uint32_t)U + (uint32_t)g, buf = uc % NBUF, pht = (uc / NBUF) & 1u;
                wait_t(BAR_TF + buf, pht, 0, rec);
                tc_fence_after();
                if (!(DBG && (p.dbg_skip & 4))) scan_buffer<0, DBG>(lane_base + buf * UNIT_N, rA, rB, dbg ? dbg + g * UNIT_N - 0 : nullptr);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_mma_side(BAR_TE + buf);
            }
            if (U == 4) {   // second unit (codes 128*(g+2) ..)
                const uint32_t uc = it * 4u + (uint32_t)g + 2u, buf = uc % NBUF, pht = (uc / NBUF) & 1u;
                wait_t(BAR_TF + buf, pht, 0, rec);
                tc_fence_after();
                if (!(DBG && (p.dbg_skip & 4))) scan_buffer<4, DBG>(lane_base + buf * UNIT_N, rA, rB, dbg ? dbg + (g + 2) * UNIT_N - 4 * 32 : nullptr);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_mma_side(BAR_TE + buf);
            }
            float m1, m2;
            int k1;
            {
                const int v = scan_finish(rA, rB, m1, m2);        // virtual column 0..255 of this group
                k1 = (v < UNIT_N ? g : g + 2) * UNIT_N + (v & (UNIT_N - 1));
                if (DBG && (p.dbg_skip & 4)) { m1 = 1.f; m2 = 1e9f; k1 = 0; }
            }
            if (DBG && prof && rec) pacc[1] += clock64() - t_scan0;
            {
                const uint32_t ps = it & 1u, php = (it >> 1) & 1u;
                if (g == 0) {                    // hand this group's result to group 1
                    mbar_wait(bar(BAR_PE + ps), php ^ 1u);
                    part_s[ps * TILE_M + row_in_tile] = make_float4(m1, m2, __int_as_float(k1), 0.f);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar(BAR_PF + ps));
                    continue;
                }
                wait_t(BAR_PF + ps, php, 2, rec);
                const float4 o = part_s[ps * TILE_M + row_in_tile];
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(BAR_PE + ps));
                const int ko = __float_as_int(o.z);
                m2 = fminf(fminf(o.y, m2), fmaxf(o.x, m1));
                const bool take_other = (o.x < m1) || (o.x == m1 && ko < k1);   // equal scores: lowest code (gap 0 -> flagged anyway)
                if (take_other) { m1 = o.x; k1 = ko; }
            }
            // certificate: every other code's lower bound must clear the winner's upper bound
            const float xn = rownorm_s[(it % NORM_RING) * TILE_M + row_in_tile];
            float en = enorm_s[k1 < K ? k1 : 0];
            const bool in_range = grow < p.n_rows;
            k1 += p.code_base;
            bool bad = cb_bad;
            if (p.partial) {                     // sliced codebook: fold in the earlier slices / hand on to the later ones
                if (!p.pass_first && in_range) {
                    const float4 o = p.partial[grow];
                    const int ko = __float_as_int(o.z);
                    bad = bad || ((__float_as_uint(o.w) >> 31) != 0u);   // an earlier slice saw a non-finite codebook
                    m2 = fminf(fminf(o.y, m2), fmaxf(o.x, m1));
                    if ((o.x < m1) || (o.x == m1 && ko < k1)) { m1 = o.x; k1 = ko; en = fabsf(o.w); }
                }
                if (!p.pass_last && in_range) p.partial[grow] = make_float4(m1, m2, __int_as_float(k1), bad ? __uint_as_float(__float_as_uint(en) | 0x80000000u) : en);
            }
            const float need = 2.f * (p.cA * BOUND_UP * xn * en + p.cB * (BOUND_UP * en * en + xn * xn));
            const bool certified = ((m2 - m1) > need) && (xn < 1.0e18f) && !bad;   // NaN -> false
            const uint32_t rs = it % RES_RING, phr = (it / RES_RING) & 1u;
            wait_t(BAR_RE + rs, phr ^ 1u, 3, rec);
            int code = -2;
            if (in_range && p.pass_last) {
                code = certified ? k1 : -1;
                if (certified) p.embed_ind[grow] = (int64_t)k1;
            }
            codes_s[rs * TILE_M + row_in_tile] = code;
            const unsigned fl = __ballot_sync(0xffffffffu, code == -1);
            if (fl) {
                int basei = 0;
                const int leader = __ffs(fl) - 1;
                if (lane == leader) basei = atomicAdd(p.flagged_count, __popc(fl));
                basei = __shfl_sync(0xffffffffu, basei, leader);
                if (code == -1) p.flagged_rows[basei + __popc(fl & ((1u << lane) - 1u))] = (int)grow;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(BAR_RF + rs));
        }
        if (DBG && prof && rec) { flush(pbase, 0); flush(pbase + 1, 1); if (g) { flush(PF_EPI1_WAIT_PF, 2); flush(PF_EPI1_WAIT_RE, 3); } prof[g ? PF_EPI1_TOTAL : PF_EPI0_TOTAL] = (unsigned long long)(clock64() - t_role0); }
    } else {
        // ================= output: gather, straight-through value, loss =================================
        // 8 warps x 16 rows; a half-warp handles one row per instruction (16 lanes x 16 bytes = one 256-byte row), 8 rows
        // per batch with all loads in flight together.  Fast path (every row of the batch certified, the common case):
        // straight-line code, row addresses are compile-time offsets from one per-tile pointer.
        reg_dec<72>();
        const int ow = warp - W_OUT;             // rows ow*16 .. ow*16+15 of the tile
        const int half = lane >> 4, q4 = lane & 15;
        float dacc = 0.f;
        const bool rec = (warp == W_OUT && lane == 0);
        constexpr int RQ = TC_D / 4;             // float4 per row
        const size_t my_off = (size_t)(ow * 16 + half) * RQ + q4;       // my first row of a tile, my 16-byte column
        const float4* cb4 = reinterpret_cast<const float4*>(p.cbT) + q4;
        const long long t_role0 = clock64();
        for (uint32_t it = 0; it < n_iter; ++it) {
            const int64_t t = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
            const uint32_t rs = it % RES_RING, phr = (it / RES_RING) & 1u;
            wait_r(BAR_RF + rs, phr, 0, rec);
            if (NCHW) {
                // NCHW-physical quantize: lanes = 32 consecutive rows at a fixed dim (128-byte coalesced stores); warp ow
                // takes row group ow & 3 and the dims [32 (ow >> 2), +32).  x comes from the shared-memory stage (x^T
                // [d][row], kept until the output warps release it), the code row is gathered per lane with all 8 loads
                // in flight: one L2 round trip per tile.
                const uint32_t sx = it % XS;
                if (!(DBG && (p.dbg_skip & 1))) {
                    const int r = (ow & 3) * 32 + lane, d0 = (ow >> 2) * 32;
                    const int k = codes_s[rs * TILE_M + r];
                    if (k >= 0) {
                        const int64_t n0 = t * TILE_M;
                        const int64_t base = (n0 / p.rpi) * p.img_stride + (n0 % p.rpi) + r + (int64_t)d0 * p.col_stride;
                        float* og = p.quantize ? p.quantize + base : nullptr;
                        const float* xsr = reinterpret_cast<const float*>(sm + P::off_x(K) + sx * P::X_STAGE) + d0 * TILE_M + r;
                        const float4* q4p = reinterpret_cast<const float4*>(p.cbT + (size_t)k * TC_D + d0);
                        float qq[4][8];
#pragma unroll
                        for (int c = 0; c < 4; ++c) ldg_nc_v8(reinterpret_cast<const float*>(q4p) + 8 * c, qq[c]);   // 32 B per lane and instruction
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float qv[4] = {qq[c >> 1][(c & 1) * 4], qq[c >> 1][(c & 1) * 4 + 1], qq[c >> 1][(c & 1) * 4 + 2], qq[c >> 1][(c & 1) * 4 + 3]};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float xv = xsr[(c * 4 + j) * TILE_M];
                                const float dl = qv[j] - xv;
                                dacc = fmaf(dl, dl, dacc);
                                if (og) __stcs(og + (int64_t)(c * 4 + j) * p.col_stride, xv + dl);
                                if (p.stat_sums) red_add_f32(p.stat_sums + (size_t)k * TC_D + d0 + c * 4 + j, xv);
                            }
                        }
                        if (p.stat_sums && d0 == 0) red_add_f32(p.stat_counts + k, 1.0f);
                    }
                }
                __syncwarp();
                if (lane == 0) { mbar_arrive(bar(BAR_RE + rs)); mbar_arrive(bar(BAR_XE + sx)); }
                continue;
            }
            const float4* xt = reinterpret_cast<const float4*>(p.x) + (size_t)t * TILE_M * RQ + my_off;
            float4* qt = p.quantize ? reinterpret_cast<float4*>(p.quantize) + (size_t)t * TILE_M * RQ + my_off : nullptr;
            const int* cs = codes_s + rs * TILE_M + ow * 16 + half;
            if (!(DBG && (p.dbg_skip & 1)))
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                int kk[4];
                float4 xv[4], qv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) kk[i] = cs[b * 8 + i * 2];
                const bool all_ok = __all_sync(0xffffffffu, (kk[0] | kk[1] | kk[2] | kk[3]) >= 0);
                if (all_ok && !p.stat_sums) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) xv[i] = __ldcg(xt + (b * 8 + i * 2) * RQ);
#pragma unroll
                    for (int i = 0; i < 4; ++i) qv[i] = __ldcg(cb4 + (uint32_t)kk[i] * RQ);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float4 d, o;
                        d.x = qv[i].x - xv[i].x; d.y = qv[i].y - xv[i].y; d.z = qv[i].z - xv[i].z; d.w = qv[i].w - xv[i].w;
                        o.x = xv[i].x + d.x; o.y = xv[i].y + d.y; o.z = xv[i].z + d.z; o.w = xv[i].w + d.w;
                        dacc = fmaf(d.x, d.x, fmaf(d.y, d.y, fmaf(d.z, d.z, fmaf(d.w, d.w, dacc))));
                        if (qt) __stcs(qt + (b * 8 + i * 2) * RQ, o);
                    }
                } else {                         // some row flagged / past the end (or the atomics fallback for statistics)
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (kk[i] >= 0) {
                            xv[i] = __ldcg(xt + (b * 8 + i * 2) * RQ);
                            qv[i] = __ldcg(cb4 + (uint32_t)kk[i] * RQ);
                        }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (kk[i] < 0) continue;
                        float4 d, o;
                        d.x = qv[i].x - xv[i].x; d.y = qv[i].y - xv[i].y; d.z = qv[i].z - xv[i].z; d.w = qv[i].w - xv[i].w;
                        o.x = xv[i].x + d.x; o.y = xv[i].y + d.y; o.z = xv[i].z + d.z; o.w = xv[i].w + d.w;
                        dacc = fmaf(d.x, d.x, fmaf(d.y, d.y, fmaf(d.z, d.z, fmaf(d.w, d.w, dacc))));
                        if (qt) __stcs(qt + (b * 8 + i * 2) * RQ, o);
                        if (p.stat_sums) {       // only when the caller has no room for the segmented-reduction kernel
                            red_add_v4(p.stat_sums + (size_t)kk[i] * TC_D + q4 * 4, xv[i]);
                            if (q4 == 0) red_add_f32(p.stat_counts + kk[i], 1.0f);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(BAR_RE + rs));
        }
        if (DBG && prof && rec) { flush(PF_OUT_WAIT_RF, 0); prof[PF_OUT_TOTAL] = (unsigned long long)(clock64() - t_role0); }
        if (p.diff_acc) {
            dacc = warp_sum(dacc);
            if (lane == 0) atomicAdd(p.diff_acc, (double)dacc);
        }
    }

    // ---- teardown -------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (CTA2) cluster_sync_all();                 // nobody leaves while the pair still reads its smem / arrives on its barriers
    if (warp == W_MMA) { if (CTA2) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
    if (DBG && prof && threadIdx.x == 0) prof[PF_KERNEL] = (unsigned long long)(clock64() - t_kernel0);
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
inline int tc_nsplit() {
    static int v = [] {
        const char* e = getenv("VQB200_TC_SPLIT");
        int n = e ? atoi(e) : 3;
        return n == 1 ? 1 : 3;
    }();
    return v;
}
inline int tc_num_sms() {
    static int v = [] {
        int dev = 0, n = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        return n > 0 ? n : 148;
    }();
    return v;
}

inline bool tc_shape_ok(int dim, int n_embed) { return dim == tc::TC_D && (n_embed == 256 || n_embed == 512); }
// larger codebooks run the same kernel once per 512-code slice (operand image = K / 512 sub-images)
constexpr int TC_SLICE = 512;
inline bool tc_sliced_ok(int dim, int n_embed) {
    return dim == tc::TC_D && n_embed > TC_SLICE && n_embed % TC_SLICE == 0 && n_embed <= 16384;
}
inline bool tc_any_ok(int dim, int n_embed) { return tc_shape_ok(dim, n_embed) || tc_sliced_ok(dim, n_embed); }

// NCHW-physical rows (the permute(0,2,3,1) view of vqvae.py:227,235) the kernel consumes in place: unit row stride, whole
// 128-row tiles inside one image, 16-byte aligned 512-byte pieces
inline bool tc_layout_nchw(const RowLayout& L, const float* x, int dim) {
    if (L.row_stride != 1 || L.col_stride < L.rows_per_image || dim != tc::TC_D) return false;
    if (L.rows_per_image % tc::TILE_M != 0 || L.n_rows % L.rows_per_image != 0) return false;
    if ((L.col_stride & 3) != 0 || (L.image_stride & 3) != 0) return false;
    if (L.n_rows > L.rows_per_image && L.image_stride < (int64_t)dim * L.col_stride) return false;
    return (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
}
inline bool tc_layout_dense(const RowLayout& L, const float* x, int dim) {
    if (L.col_stride != 1 || L.row_stride != dim) return false;
    if (L.n_rows > L.rows_per_image && L.image_stride != L.rows_per_image * dim) return false;
    return (reinterpret_cast<uintptr_t>(x) & 15u) == 0;                          // bulk copies need 16-byte alignment
}
inline bool tc_supported(const RowLayout& L, const float* x, int dim, int n_embed) {
    if (!tc_any_ok(dim, n_embed) || L.n_rows < 1) return false;
    static const bool disabled = getenv("VQB200_DISABLE_TC") != nullptr;     // environment switches are read once per process
    if (disabled) return false;
    return tc_layout_dense(L, x, dim) || tc_layout_nchw(L, x, dim);
}

// x (NCHW-physical) as a 3-D tensor map: dims (fastest first) row-in-image, dim, image; box = one 128-row x 64-dim tile
inline int tc_encode_tmap(CUtensorMap* tm, const float* x, const RowLayout& L, int dim) {
    static PFN_cuTensorMapEncodeTiled encode = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) f = nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(f);
    }();
    if (!encode) return 1;
    const cuuint64_t n_img = (cuuint64_t)(L.n_rows / L.rows_per_image);
    const cuuint64_t img_stride = n_img > 1 ? (cuuint64_t)L.image_stride : (cuuint64_t)dim * (cuuint64_t)L.col_stride;
    cuuint64_t gdim[3] = {(cuuint64_t)L.rows_per_image, (cuuint64_t)dim, n_img};
    cuuint64_t gstr[2] = {(cuuint64_t)L.col_stride * 4u, img_stride * 4u};
    cuuint32_t box[3] = {(cuuint32_t)tc::TILE_M, (cuuint32_t)dim, 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}

// dense x [N][64] for the tf32 filter: 2-D tensor map, box = 128 rows x 32 dims (one 128-byte swizzle row per x row and k-block);
// two boxes per tile land as the K-major SWIZZLE_128B operand the tf32 MMAs read directly
inline int tc_encode_tmap_dense_tf(CUtensorMap* tm, const float* x, int64_t n_rows) {
    static PFN_cuTensorMapEncodeTiled encode = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) f = nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(f);
    }();
    if (!encode) return 1;
    cuuint64_t gdim[2] = {(cuuint64_t)tc::TC_D, (cuuint64_t)n_rows};
    cuuint64_t gstr[1] = {(cuuint64_t)tc::TC_D * 4u};
    cuuint32_t box[2] = {32u, (cuuint32_t)tc::TILE_M};
    cuuint32_t estr[2] = {1u, 1u};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}

// CTAs the tensor-core kernel runs for n_rows rows (= number of private statistics tables it fills)
inline int tc_grid(int64_t n_rows, bool pair) {
    int64_t n_tiles = (n_rows + tc::TILE_M - 1) / tc::TILE_M;
    int grid = (int)std::min<int64_t>(n_tiles, tc_num_sms());
    if (!pair) return grid;
    grid = std::max(2, (grid + 1) / 2 * 2);      // whole CTA pairs; 148 SMs = 74 pairs
    if (grid > tc_num_sms()) grid = tc_num_sms() / 2 * 2;
    return grid;
}

template <int NSPLIT, int AS, int XS, bool DBG, bool CTA2, bool NCHW = false>
inline int tc_launch(const tc::Params& prm, cudaStream_t st) {
    using P = tc::Plan<NSPLIT, AS, XS, CTA2>;
    auto kern = tc::k_vq_tc<NSPLIT, AS, XS, DBG, CTA2, NCHW>;
    const int smem = (int)P::total(prm.K);
    // opt-in shared-memory size is a per-device function attribute: cache it per device (several GPUs in one process)
    static std::atomic<int> configured_dev[64];          // (atomic: several host threads / GPUs per process)
    int dev_id = 0;
    if (cudaGetDevice(&dev_id) != cudaSuccess || dev_id < 0 || dev_id >= 64) dev_id = 0, configured_dev[0] = 0;
    std::atomic<int>& configured = configured_dev[dev_id];
    if (configured.load(std::memory_order_relaxed) < smem) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1;
        configured = smem;
    }
    const int grid = tc_grid(prm.n_rows, CTA2);
    if (!CTA2) return launch_pdl(kern, dim3((unsigned)grid), dim3(tc::THREADS), (size_t)smem, st, prm) != cudaSuccess;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(tc::THREADS);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    return cudaLaunchKernelEx(&cfg, kern, prm) != cudaSuccess;
}

// main kernel only; the caller runs the exact fix-up over the flagged rows afterwards
inline int tc_forward(const float* x, const RowLayout& L, int dim, int n_embed, const CodebookImage& cb,
                      float* quantize, int64_t* embed_ind, const ForwardScratch& sc, double* diff_acc,
                      float* sums, float* counts, float* dbg_scores, cudaStream_t st,
                      unsigned long long* prof = nullptr, int nsplit = -1, int* grid_out = nullptr,
                      float* x_dense = nullptr) {
    // nsplit: 3 = split-bf16 filter, 1 = plain bf16, 0 = tf32 (MMAs read the fp32 x stage), -1 = the build's default
    if (nsplit != 0 && nsplit != 1 && nsplit != 3) nsplit = tc_nsplit();
    const bool nchw = !tc_layout_dense(L, x, dim);
    if (nsplit == 0 && nchw) nsplit = 1;           // (the tf32 filter covers dense rows; NCHW-physical rows keep the bf16 kernels)
    (void)dim;
    static const bool pair = [] { const char* e = getenv("VQB200_TC_CTA2"); return e ? atoi(e) != 0 : true; }();
    static const bool pair1 = [] { const char* e = getenv("VQB200_TC_CTA2_BF16"); return e ? atoi(e) != 0 : false; }();
    const bool sliced = tc_sliced_ok(dim, n_embed);
    const int K_launch = sliced ? TC_SLICE : n_embed;
    const int n_slices = sliced ? n_embed / TC_SLICE : 1;
    const bool use_pair = K_launch == 512 && (nsplit == 3 ? pair : pair1);
    if (grid_out) *grid_out = tc_grid(L.n_rows, use_pair);
    const bool dbg = dbg_scores || prof;           // diagnostics live in a separate instantiation
    if (sliced && (dbg || !sc.partial)) return 1;
    if (nchw && dbg) return 1;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (nchw || nsplit == 0) {
        // tensor maps depend only on (pointer, layout, kind): training loops call with the same few buffers over and over,
        // so the last maps are kept (cuTensorMapEncodeTiled costs ~1-2 us on the host)
        struct Entry { const float* x; RowLayout L; int kind; CUtensorMap map; };
        static thread_local Entry cache[8];
        static thread_local int next = 0;
        const int kind = nchw ? 1 : 2;
        const Entry* hit = nullptr;
        for (const Entry& e : cache)
            if (e.x == x && e.kind == kind && e.L.n_rows == L.n_rows && e.L.rows_per_image == L.rows_per_image &&
                e.L.image_stride == L.image_stride && e.L.row_stride == L.row_stride && e.L.col_stride == L.col_stride) { hit = &e; break; }
        if (hit) tmap = hit->map;
        else {
            if (nchw ? tc_encode_tmap(&tmap, x, L, dim) : tc_encode_tmap_dense_tf(&tmap, x, L.n_rows)) return 1;
            cache[next] = Entry{x, L, kind, tmap};
            next = (next + 1) % 8;
        }
    }
    for (int sl = 0; sl < n_slices; ++sl) {
        tc::Params prm;
        prm.tmap = tmap;
        prm.x = x; prm.n_rows = L.n_rows; prm.K = K_launch;
        prm.image = cb.tc + (size_t)sl * tc::image_bytes(TC_SLICE); prm.cbT = cb.cbT;
        prm.quantize = quantize; prm.embed_ind = embed_ind; prm.diff_acc = diff_acc;
        prm.stat_sums = sums; prm.stat_counts = counts;
        prm.flagged_count = sc.flagged_count; prm.flagged_rows = sc.flagged_rows; prm.dbg_scores = dbg_scores;
        prm.prof = prof;
        static const int dbg_skip_env = [] { const char* e = getenv("VQB200_DBG_SKIP"); return e ? atoi(e) : 0; }();
        prm.dbg_skip = dbg_skip_env;
        prm.cA = tc::bound_cA(nsplit); prm.cB = tc::BOUND_CB;
        prm.rpi = L.rows_per_image; prm.img_stride = L.image_stride; prm.col_stride = L.col_stride;
        prm.x_dense = (nchw && sl == 0) ? x_dense : nullptr;
        prm.code_base = sl * TC_SLICE; prm.partial = sliced ? sc.partial : nullptr;
        prm.pass_first = sl == 0; prm.pass_last = sl == n_slices - 1;
        int rc;
        if (nchw) {                                // x stages x3: the output warps read x from the stage too
            if (nsplit == 3) rc = (pair && K_launch == 512) ? tc_launch<3, 1, 3, false, true, true>(prm, st) : tc_launch<3, 1, 1, false, false, true>(prm, st);
            else rc = dbg_skip_env ? tc_launch<1, 2, 3, true, false, true>(prm, st)      // (diagnostics: role skipping)
                                                : tc_launch<1, 2, 3, false, false, true>(prm, st);
        } else if (nsplit == 0) {
            rc = dbg ? tc_launch<0, 2, 2, true, false>(prm, st) : tc_launch<0, 2, 2, false, false>(prm, st);
        } else if (nsplit == 3) {
            if (pair && K_launch == 512)           // CTA pairs: half the operand image per CTA -> double-buffered A and x
                rc = dbg ? tc_launch<3, 2, 2, true, true>(prm, st) : tc_launch<3, 2, 2, false, true>(prm, st);
            else
                rc = dbg ? tc_launch<3, 1, 1, true, false>(prm, st) : tc_launch<3, 1, 1, false, false>(prm, st);
        } else if (pair1 && K_launch == 512) {
            rc = dbg ? tc_launch<1, 2, 3, true, true>(prm, st) : tc_launch<1, 2, 3, false, true>(prm, st);
        } else {
            rc = dbg ? tc_launch<1, 2, 2, true, false>(prm, st) : tc_launch<1, 2, 2, false, false>(prm, st);
        }
        if (rc) return rc;
    }
    return 0;
}

}  // namespace vqb200
