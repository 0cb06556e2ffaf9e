// tcgen05 assignment kernel for WIDE codes: dim = 64 * DB, DB = 2 or 4 (the D axis of BASELINE.json's cfg-5 sweep and the
// D = 256 quantizers of vqvae_deep.py:252,257).  Same contract, certificate and epilogue as tc::k_vq_tc (tc_kernel.cuh);
// what changes is the contraction: it runs over DB blocks of 64 dims, and neither a whole fp32 x tile (128 rows x D) nor
// all bf16 A blocks of a tile fit in shared memory next to the codebook image.  So the loop order is BLOCK-outer,
// UNIT-inner: all accumulator units of a tile are live in TMEM at once, every 64-dim block of x is staged (2-D TMA tensor
// map, box = 128 rows x 64 dims), converted to one 16-KB bf16 A stage and multiplied against the resident operand image
// of that block for every unit; the "misc" MMA (bias, row offset, error bound -- it needs the full row norm) goes LAST.
//
//   resident per CTA : eh blocks [DB][KL][64] bf16 (128 KB) + misc rows, KL = codes per launch (512 at D = 128, 256 at
//                      D = 256); larger codebooks run one launch per KL-code slice with the running (best, runner-up,
//                      winner) carried per row, exactly like the sliced mode of tc::k_vq_tc
//   filter           : plain bf16 (cA = 7.9e-3); rows it cannot certify go to the exact fp32 fix-up (k_fixup)
//   accumulation term: cB scales with the number of accumulated MMAs (BOUND_CB * DB)
#pragma once
#include <atomic>
#include "tc_kernel.cuh"

namespace vqb200 {
namespace tcw {
using namespace tc;

constexpr int AS_CONV = 2;                   // bf16 A stages (one 64-dim block of a 128-row tile each) filled by the converters
// streamed-operand kernels (PRE): 4 stages.  At DB = 2 (four scan groups) three groups hand partial results to
// the merging group through ONE slot of 3 x 2 KB, and the code norms are read from global memory instead of a 2-KB copy.
__host__ __device__ constexpr int as_pre(int) { return 4; }
__host__ __device__ constexpr int n_partials(bool pre, int DB) { return (pre && DB == 2) ? 3 : 1; }
__host__ __device__ constexpr int n_part_slots(bool pre, int DB) { return (pre && DB == 2) ? 1 : 2; }
__host__ __device__ constexpr bool enorm_in_smem(bool pre, int DB) { return !(pre && DB == 2); }
constexpr uint32_t A_STAGE = 16384u, AM_STAGE = 4096u;
constexpr uint32_t X_ROWS = 64, X_STAGE = X_ROWS * 64 * 4;   // x stage = HALF a block (64 rows x 64 dims fp32): the upper half is
                                                             // reloaded while the converters still work on the lower half

__host__ __device__ inline size_t wimage_off_misc(int KL, int DB) { return (size_t)KL * 128 * DB; }
__host__ __device__ inline size_t wimage_off_enorm(int KL, int DB) { return (size_t)KL * (128 * DB + 32); }
__host__ __device__ inline size_t wimage_bytes(int KL, int DB) { return align_up((size_t)KL * (128 * DB + 36), 1024); }
__host__ __device__ constexpr float bound_cB(int DB) { return BOUND_CB * (float)DB; }

// operand image of the wide engine from cbT [K][D] / ee [K]: one thread per (code, 8-dim chunk)
__global__ void __launch_bounds__(256) k_prepare_wide(const float* __restrict__ cbT, const float* __restrict__ ee,
                                                       unsigned char* __restrict__ img, int K, int D, int KL, float cA1, float cB) {
    pdl_wait();
    pdl_trigger();
    const int nch = D / 8, DB = D / 64;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= K * nch) return;
    const int k = idx / nch, c = idx % nch, sl = k / KL, kl = k % KL;
    unsigned char* base = img + (size_t)sl * wimage_bytes(KL, DB);
    const float* e = cbT + (size_t)k * D + c * 8;
    uint32_t hi[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) hi[j] = pack_bf16(-2.f * e[2 * j], -2.f * e[2 * j + 1]);
    *reinterpret_cast<uint4*>(base + (size_t)(c >> 3) * KL * 128 + sw128_off((uint32_t)kl, (uint32_t)(c & 7) * 8)) =
        make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (c == 0) {
        const float e2 = ee[k];
        float b1, b2, b3;
        split3(e2, b1, b2, b3);
        const float ne = sqrtf(e2);
        const float t6 = -bf16_round(cA1 * ne * 1.0078125f);     // rounded up in magnitude
        const float t7 = -bf16_round(cB * e2 * 1.0078125f);
        unsigned char* m = base + wimage_off_misc(KL, DB);
        *reinterpret_cast<uint4*>(m + sw32_chunk_off((uint32_t)kl, 0)) =
            make_uint4(pack_bf16(b1, b2), pack_bf16(b3, 1.f), pack_bf16(1.f, 1.f), pack_bf16(t6, t7));
        *reinterpret_cast<uint4*>(m + sw32_chunk_off((uint32_t)kl, 1)) = make_uint4(0, 0, 0, 0);
        reinterpret_cast<float*>(base + wimage_off_enorm(KL, DB))[kl] = ne;
    }
}

// x [N][64 DB] fp32 -> per 128-row tile: DB bf16 operand blocks, the tile's misc rows and the row norms, all byte-identical
// to the shared-memory stages of k_vq_tcw<PRE> (one plain bulk copy each).  Sliced codebooks convert x ONCE per call
// instead of once per slice.  A half-warp owns a row: coalesced 256-byte reads, 128-byte writes.
template <int DB>
__global__ void __launch_bounds__(256) k_convert_wide(const float* __restrict__ x, int64_t n_rows, unsigned char* __restrict__ a_img,
                                                       unsigned char* __restrict__ m_img, float* __restrict__ norms) {
    pdl_wait();
    pdl_trigger();
    constexpr int D = 64 * DB;
    const int lane = threadIdx.x & 31, q4 = lane & 15;
    const int64_t padded = (n_rows + TILE_M - 1) / TILE_M * TILE_M;
    const int64_t n_hw = ((int64_t)gridDim.x * blockDim.x) >> 4;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4; row < padded; row += n_hw) {
        const int64_t t = row >> 7;
        const uint32_t r = (uint32_t)(row & 127);
        const bool in = row < n_rows;                              // rows past the end: zeros (finite scores, never used)
        float4 v[DB];
#pragma unroll
        for (int b = 0; b < DB; ++b)
            v[b] = in ? __ldcs(reinterpret_cast<const float4*>(x + row * D + b * 64) + q4) : make_float4(0.f, 0.f, 0.f, 0.f);
        float sq = 0.f;
#pragma unroll
        for (int b = 0; b < DB; ++b) {
            *reinterpret_cast<uint2*>(a_img + ((size_t)t * DB + b) * A_STAGE + sw128_off(r, (uint32_t)q4 * 4)) =
                make_uint2(pack_bf16(v[b].x, v[b].y), pack_bf16(v[b].z, v[b].w));
            sq = fmaf(v[b].x, v[b].x, fmaf(v[b].y, v[b].y, fmaf(v[b].z, v[b].z, fmaf(v[b].w, v[b].w, sq))));
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);      // stays inside the half-warp
        if (q4 == 0) {
            const float nx = sqrtf(sq);
            float o1, o2, o3;
            split3(sq * 1.001953125f, o1, o2, o3);                 // off_i = ||x||^2 (1 + 2^-9)
            const float nxu = bf16_round(nx * 1.0078125f);         // ||x|| rounded up
            unsigned char* m = m_img + (size_t)t * AM_STAGE;
            *reinterpret_cast<uint4*>(m + sw32_chunk_off(r, 0)) =
                make_uint4(pack_bf16(1.f, 1.f), pack_bf16(1.f, o1), pack_bf16(o2, o3), pack_bf16(nxu, 1.f));
            *reinterpret_cast<uint4*>(m + sw32_chunk_off(r, 1)) = make_uint4(0, 0, 0, 0);
            norms[row] = nx;
        }
    }
}

struct Plan {
    int KL, DB, XS, AS, NP, NSLOT, ENORM;
    int BDIV = 1;                                // 2: CTA pair, every CTA holds half of the B rows of each 128-code unit
    __host__ __device__ uint32_t off_bmisc() const { return (uint32_t)(KL / BDIV) * 128u * (uint32_t)DB; }
    __host__ __device__ uint32_t off_a() const { return off_bmisc() + (uint32_t)(KL / BDIV) * 32u; }
    __host__ __device__ uint32_t off_am() const { return off_a() + (uint32_t)AS * A_STAGE; }
    __host__ __device__ uint32_t off_x() const { return off_am() + 2u * AM_STAGE; }
    __host__ __device__ uint32_t off_small() const { return off_x() + (uint32_t)XS * X_STAGE; }
    __host__ __device__ uint32_t off_rownorm() const { return off_small() + (ENORM ? (uint32_t)KL * 4u : 0u); }
    __host__ __device__ uint32_t off_codes() const { return off_rownorm() + NORM_RING * TILE_M * 4u; }
    __host__ __device__ uint32_t off_parts() const { return off_codes() + RES_RING * TILE_M * 4u; }
    __host__ __device__ uint32_t off_bars() const { return off_parts() + (uint32_t)NSLOT * (uint32_t)NP * TILE_M * 16u; }
    __host__ __device__ uint32_t total() const { return off_bars() + 384u + 1024u /* base alignment slack */; }
};

struct WParams {
    alignas(64) CUtensorMap tmap;    // x as a 2-D tensor [row][dim], box = 64 rows x 64 dims
    const float* x;
    int64_t n_rows;
    int KL;                          // codes in this launch (256 or 512)
    int K_total;                     // row pitch of dbg_scores
    const unsigned char* image;      // this slice's operand image
    const float* cbT;                // [K][D] fp32
    float* quantize;                 // may be null
    int64_t* embed_ind;
    double* diff_acc;                // may be null
    float* stat_sums;                // may be null (global-atomics statistics)
    float* stat_counts;
    int* flagged_count;
    int* flagged_rows;
    float* dbg_scores;               // DBG builds: [n_rows][K_total] dump of the tensor-core scores
    float cA, cB;
    int code_base;
    float4* partial;                 // may be null (single launch)
    int pass_first, pass_last;
    // PRE = true (sliced codebooks): x was converted ONCE by k_convert_wide; the producer streams ready-made operand stages
    const unsigned char* a_img;      // [tile][DB][16384]: bf16 A blocks, byte-identical to their 128B-swizzled shared-memory layout
    const unsigned char* m_img;      // [tile][4096]: misc rows of the tile (32-byte swizzle)
    const float* norms;              // [tile][128]: ||x_row||
    int64_t row_base;                // global id of row 0 of this launch (row-chunked launches; flagged-row list entries are global)
};

enum WBar { WB_B = 0, WB_XF = 1, WB_XE = 5, WB_AF = 9, WB_AE = 13, WB_TF = 17, WB_TE = 21, WB_RF = 25, WB_RE = 27, WB_PF = 29,
            WB_PE = 31, WB_PB = 33, WB_AFP = 34, WB_COUNT = 38 };

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar), "l"(policy) : "memory");
}

// four K = 16 MMAs of one 64-dim block into one accumulator unit (converged warp, one elected lane issues); the first one
// overwrites the accumulator when acc0 == 0
template <bool CTA2 = false>
__device__ __forceinline__ void issue_block4(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t acc0) {
#define VQW_KS(CG, KS)                                                                       \
    "add.u32 ta, %1, " KS ";\n\tadd.u32 tb, %2, " KS ";\n\t"                                   \
    "mov.b64 da, {ta, %3};\n\tmov.b64 db, {tb, %3};\n\t"                                        \
    "@pe tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %4, pt;\n\t"
#define VQW_BLOCK4(CG)                                                                       \
    asm volatile("{\n\t.reg .pred pa, pt, pe;\n\t.reg .b64 da, db;\n\t.reg .b32 ta, tb;\n\t"  \
                 "elect.sync _|pe, 0xffffffff;\n\t"                                           \
                 "setp.ne.b32 pa, %5, 0;\n\tsetp.eq.b32 pt, %4, %4;\n\t"                      \
                 "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"                          \
                 "@pe tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %4, pa;\n\t"        \
                 VQW_KS(CG, "2") VQW_KS(CG, "4") VQW_KS(CG, "6") "}"                          \
                 :: "r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(DESC_HI_SW128), "r"(CTA2 ? IDESC2 : IDESC), "r"(acc0) : "memory")
    if constexpr (CTA2) VQW_BLOCK4("2"); else VQW_BLOCK4("1");
#undef VQW_BLOCK4
#undef VQW_KS
}
// the misc MMA (K = 16, 32-byte swizzle rows): accumulates bias, row offset and error bound onto the finished products
template <bool CTA2 = false>
__device__ __forceinline__ void issue_misc(uint32_t d_tmem, uint32_t am_lo, uint32_t bm_lo) {
#define VQW_MISC(CG)                                                                         \
    asm volatile("{\n\t.reg .pred pt, pe;\n\t.reg .b64 da, db;\n\t"                            \
                 "elect.sync _|pe, 0xffffffff;\n\t"                                           \
                 "setp.eq.b32 pt, %4, %4;\n\t"                                                \
                 "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"                          \
                 "@pe tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %4, pt;\n\t}"       \
                 :: "r"(d_tmem), "r"(am_lo), "r"(bm_lo), "r"(DESC_HI_SW32), "r"(CTA2 ? IDESC2 : IDESC) : "memory")
    if constexpr (CTA2) VQW_MISC("2"); else VQW_MISC("1");
#undef VQW_MISC
}

// CTA2 = true (conversion path only): clusters of two CTAs share ONE tcgen05.mma.cta_group::2 stream (M = 256) issued by the
// leader; every CTA converts, scans and writes its own 128-row tile but holds only its 64 rows of each 128-code unit of the
// operand image -- KL = 512 codes stay resident at D = 256 (128 KB per CTA), i.e. ONE pass over x instead of two.
template <int DB, int XS, bool DBG, bool PRE, bool CTA2 = false>
__global__ void __launch_bounds__(THREADS, 1) k_vq_tcw(const __grid_constant__ WParams p) {
    constexpr int D = 64 * DB;
    constexpr int AS = PRE ? as_pre(DB) : AS_CONV;
    constexpr int NP = n_partials(PRE, DB), NSLOT = n_part_slots(PRE, DB);
    constexpr bool ENORM_S = enorm_in_smem(PRE, DB);
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    const int KL = p.KL;
    const int U = KL / UNIT_N;                  // accumulator units per tile: 2 or 4
    // Streamed passes that write no outputs (every slice but the last) have idle output warps: they join as scan groups 2 and
    // 3, so that each of the four 128-column units of a tile is scanned by its own warpgroup (four scanning warps per
    // scheduler instead of two: the scan is bound by TMEM-load / dependent-minimum latency).
    const bool helper = PRE && DB == 2 && !DBG && U == 4 && !p.pass_last;
    const int NGRP = helper ? 4 : 2;
    const Plan P{KL, DB, PRE ? 0 : XS, AS, NP, NSLOT, ENORM_S ? 1 : 0, CTA2 ? 2 : 1};
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KLB = CTA2 ? KL / 2 : KL;          // operand-image rows resident in this CTA
    constexpr uint32_t UROWS = CTA2 ? UNIT_N / 2 : UNIT_N;    // ... of each 128-code unit
    const uint32_t crank = CTA2 ? cluster_ctarank() : 0u;

    const uint32_t sB = base, sBm = base + P.off_bmisc(), sA = base + P.off_a(), sAm = base + P.off_am(), sX = base + P.off_x();
    float* enorm_s = reinterpret_cast<float*>(sm + P.off_small());
    float* rownorm_s = reinterpret_cast<float*>(sm + P.off_rownorm());
    int* codes_s = reinterpret_cast<int*>(sm + P.off_codes());
    float4* part_s = reinterpret_cast<float4*>(sm + P.off_parts());
    const uint32_t bars = base + P.off_bars();
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(sm + P.off_bars() + WB_COUNT * 8);
    auto bar = [&](int id) { return bars + 8u * (uint32_t)id; };

    pdl_wait();
    pdl_trigger();
    const int64_t n_tiles = (p.n_rows + TILE_M - 1) / TILE_M;
    // both CTAs of a pair run the same number of trips (the pair's second tile may lie past the end on the last one)
    const int64_t first_tile = CTA2 ? (int64_t)(blockIdx.x & ~1u) : (int64_t)blockIdx.x;
    const uint32_t n_iter = first_tile < n_tiles ? (uint32_t)((n_tiles - first_tile + gridDim.x - 1) / gridDim.x) : 0u;
    constexpr uint32_t PAIR_WARPS = CTA2 ? 8u : 4u;           // arriving warps on the barriers the MMA issuer waits on

    if (threadIdx.x == 0) {
        mbar_init(bar(WB_B), 1);
        mbar_init(bar(WB_PB), 1);
        for (int s = 0; s < AS; ++s) mbar_init(bar(WB_AFP + s), 2);      // streamed pair: one relay arrival per CTA (below)
        for (int s = 0; s < XS; ++s) { mbar_init(bar(WB_XF + s), 1); mbar_init(bar(WB_XE + s), 4); }
        for (int s = 0; s < AS; ++s) { mbar_init(bar(WB_AF + s), PRE ? 1 : PAIR_WARPS); mbar_init(bar(WB_AE + s), 1); }
        for (int s = 0; s < NBUF; ++s) { mbar_init(bar(WB_TF + s), 1); mbar_init(bar(WB_TE + s), PAIR_WARPS); }
        for (int s = 0; s < RES_RING; ++s) { mbar_init(bar(WB_RF + s), 4); mbar_init(bar(WB_RE + s), 8); }
        for (int s = 0; s < 2; ++s) { mbar_init(bar(WB_PF + s), helper ? 12 : 4); mbar_init(bar(WB_PE + s), 4); }
        fence_barrier_init();
    }
    if (warp == W_MMA) { if (CTA2) tmem_alloc2(smem_u32(tmem_ptr_s), 512); else tmem_alloc(smem_u32(tmem_ptr_s), 512); }
    for (int i = threadIdx.x; i < 2 * TILE_M; i += THREADS)      // constant zero half of the A misc rows
        *reinterpret_cast<uint4*>(sm + P.off_am() + (i / TILE_M) * AM_STAGE + sw32_chunk_off((uint32_t)(i % TILE_M), 1)) = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    int bad = 0;
    for (int i = threadIdx.x; i < KL; i += THREADS) {
        const float ne = reinterpret_cast<const float*>(p.image + wimage_off_enorm(KL, DB))[i];
        if (ENORM_S) enorm_s[i] = ne;
        bad |= !(ne < 1.0e18f);                  // NaN / inf / absurd norms: certify nothing, the exact path decides
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
        // ================= producer: resident operand image once, then one 2-D TMA box per (tile, block) ==========
        reg_dec<24>();
        if (lane == 0) {
            const uint32_t img_bytes = (uint32_t)KLB * 128u * DB, misc_bytes = (uint32_t)KLB * 32u;
            mbar_expect_tx(bar(WB_B), img_bytes + misc_bytes);
            if constexpr (CTA2) {                // my 64 rows of every 128-code unit, block by block
                for (int b = 0; b < DB; ++b)
                    for (int u = 0; u < U; ++u)
                        bulk_g2s(sB + ((uint32_t)b * KLB + (uint32_t)u * UROWS) * 128u,
                                 p.image + ((size_t)b * KL + (size_t)u * UNIT_N + (size_t)crank * UROWS) * 128u, UROWS * 128u, bar(WB_B));
                for (int u = 0; u < U; ++u)
                    bulk_g2s(sBm + (uint32_t)u * UROWS * 32u, p.image + wimage_off_misc(KL, DB) + ((size_t)u * UNIT_N + (size_t)crank * UROWS) * 32u,
                             UROWS * 32u, bar(WB_B));
            } else {
                for (uint32_t o = 0; o < img_bytes; o += 16384u) bulk_g2s(sB + o, p.image + o, 16384u, bar(WB_B));
                bulk_g2s(sBm, p.image + wimage_off_misc(KL, DB), misc_bytes, bar(WB_B));
            }
            const uint64_t keep = l2_policy_evict_last();       // the output warps read the tile again through L2
            for (uint32_t it = 0; it < n_iter; ++it) {
                const int64_t t = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
                if constexpr (PRE) {             // ready-made operand stages: 16 KB per block, misc rows and norms with the last block
                    for (int b = 0; b < DB; ++b) {
                        const uint32_t g = it * DB + b, s = g % AS, ph = (g / AS) & 1u;
                        mbar_wait(bar(WB_AE + s), ph ^ 1u);
                        const bool last = b == DB - 1;
                        mbar_expect_tx(bar(WB_AF + s), A_STAGE + (last ? AM_STAGE + TILE_M * 4u : 0u));
                        bulk_g2s(sA + s * A_STAGE, p.a_img + ((size_t)t * DB + b) * A_STAGE, A_STAGE, bar(WB_AF + s));
                        if (last) {
                            bulk_g2s(sAm + (it & 1u) * AM_STAGE, p.m_img + (size_t)t * AM_STAGE, AM_STAGE, bar(WB_AF + s));
                            bulk_g2s(base + P.off_rownorm() + (it % NORM_RING) * TILE_M * 4u, p.norms + (size_t)t * TILE_M, TILE_M * 4u, bar(WB_AF + s));
                        }
                    }
                } else {
                    for (int bh = 0; bh < 2 * DB; ++bh) {            // (block, half) in conversion order
                        const uint32_t g = it * 2u * DB + bh, s = g % XS, ph = (g / XS) & 1u;
                        mbar_wait(bar(WB_XE + s), ph ^ 1u);
                        mbar_expect_tx(bar(WB_XF + s), X_STAGE);     // rows past the end are zero-filled and still counted
                        tma_load_2d(sX + s * X_STAGE, &p.tmap, (bh >> 1) * 64, (int)(t * TILE_M + (bh & 1) * X_ROWS), bar(WB_XF + s), keep);
                    }
                }
            }
        }
    } else if (warp == W_MMA) {
        // ================= MMA issuer (converged warp; tcgen05 instructions predicated on one elected lane) ======
        reg_dec<24>();
        mbar_wait(bar(WB_B), 0);
        if (CTA2) {
            if (crank != 0) { if (lane == 0) mbar_arrive_cluster(mapa_rank(bar(WB_PB), 0)); }   // my half of B has landed
            else mbar_wait_cluster(bar(WB_PB), 0);
        }
        const uint32_t b_lo0 = desc_lo(sB), bm_lo0 = desc_lo(sBm);
        for (uint32_t it = 0; it < (crank == 0 ? n_iter : 0u); ++it) {     // the leader issues for the pair
            const uint32_t am_lo = desc_lo(sAm + (it & 1u) * AM_STAGE);
            if constexpr (PRE && DB * 2 <= AS) {
                // Streamed operands and two whole tiles fit in the A ring: UNIT-outer order.  The units of a tile finish one
                // after the other, every TMEM buffer runs its own scan -> MMA -> scan cycle (with four scan groups the period
                // is one unit scan + one unit of MMAs instead of the scans of a whole tile).
                for (int b = 0; b < DB; ++b) {
                    const uint32_t g = it * DB + b;
                    if (CTA2) mbar_wait_cluster(bar(WB_AFP + g % AS), (g / AS) & 1u); else mbar_wait(bar(WB_AF + g % AS), (g / AS) & 1u);
                }
                tc_fence_after();
                for (int u = 0; u < U; ++u) {
                    const uint32_t uc = it * (uint32_t)U + (uint32_t)u, buf = uc % NBUF, pht = (uc / NBUF) & 1u;
                    if (CTA2) mbar_wait_cluster(bar(WB_TE + buf), pht ^ 1u); else mbar_wait(bar(WB_TE + buf), pht ^ 1u);
                    tc_fence_after();
                    for (int b = 0; b < DB; ++b)
                        issue_block4<CTA2>(tmem_base + buf * UNIT_N, desc_lo(sA + ((it * DB + b) % AS) * A_STAGE),
                                           b_lo0 + ((((uint32_t)b * KLB + (uint32_t)u * UROWS) * 128u) >> 4), b != 0 ? 1u : 0u);
                    issue_misc<CTA2>(tmem_base + buf * UNIT_N, am_lo, bm_lo0 + (((uint32_t)u * UROWS * 32u) >> 4));
                    commit_elected<CTA2>(bar(WB_TF + buf));
                }
                for (int b = 0; b < DB; ++b) commit_elected<CTA2>(bar(WB_AE + (it * DB + b) % AS));
            } else
            for (int b = 0; b < DB; ++b) {
                const uint32_t g = it * DB + b, sa = g % AS, pha = (g / AS) & 1u;
                if (CTA2) mbar_wait_cluster(bar((PRE ? WB_AFP : WB_AF) + sa), pha); else mbar_wait(bar(WB_AF + sa), pha);
                tc_fence_after();
                const uint32_t a_lo = desc_lo(sA + sa * A_STAGE);
                for (int u = 0; u < U; ++u) {
                    const uint32_t uc = it * (uint32_t)U + (uint32_t)u, buf = uc % NBUF, pht = (uc / NBUF) & 1u;
                    if (b == 0) {
                        if (CTA2) mbar_wait_cluster(bar(WB_TE + buf), pht ^ 1u); else mbar_wait(bar(WB_TE + buf), pht ^ 1u);
                        tc_fence_after();
                    }
                    issue_block4<CTA2>(tmem_base + buf * UNIT_N, a_lo, b_lo0 + ((((uint32_t)b * KLB + (uint32_t)u * UROWS) * 128u) >> 4),
                                       b != 0 ? 1u : 0u);
                    if (b == DB - 1) {           // the misc rows of this tile were written before the last block's AF arrival
                        issue_misc<CTA2>(tmem_base + buf * UNIT_N, am_lo, bm_lo0 + (((uint32_t)u * UROWS * 32u) >> 4));
                        commit_elected<CTA2>(bar(WB_TF + buf));
                    }
                }
                commit_elected<CTA2>(bar(WB_AE + sa));           // also covers every earlier MMA (the misc rows of older tiles)
            }
        }
    } else if (warp > W_MMA) {
        reg_dec<24>();
    } else if (warp >= W_CONV) {
        // ================= converters: one fp32 [128][64] block -> bf16 K-major A stage ==========================
        if constexpr (PRE) reg_dec<24>(); else reg_dec<56>();
        const int cw = warp - W_CONV;            // rows h*64 + cw*16 .. +15 of both halves h of a block
        const int half = lane >> 4, q4 = lane & 15;
        const uint32_t conv_iter = PRE ? 0u : n_iter;             // PRE: nothing to convert
        if constexpr (PRE && CTA2) {
            // streamed pair: the A stages are filled by each CTA's own bulk copies (local transaction barrier); the first idle
            // converter warp relays "my stage s has landed" to the leader's barrier the MMA issuer waits on
            if (cw == 0) {
                for (uint32_t it = 0; it < n_iter; ++it)
                    for (int b = 0; b < DB; ++b) {
                        const uint32_t g = it * DB + b, sa = g % AS, pha = (g / AS) & 1u;
                        mbar_wait(bar(WB_AF + sa), pha);
                        if (lane == 0) arrive_mma_side(WB_AFP + sa);
                        __syncwarp();
                    }
            }
        }
        // row (within the tile) handled in trip gi (0..3: half gi >> 1), slot u, by this half-warp
        auto row_of = [&](int gi, int u) { return (gi >> 1) * 64 + cw * 16 + 2 * (4 * (gi & 1) + u) + half; };
        for (uint32_t it = 0; it != conv_iter; ++it) {
            float row_sq = 0.f;                  // ||x||^2 of row row_of(q4 >> 2, q4 & 3), accumulated over the blocks
            for (int b = 0; b < DB; ++b) {
                const uint32_t g = it * DB + b, sa = g % AS, pha = (g / AS) & 1u;
                mbar_wait(bar(WB_AE + sa), pha ^ 1u);
                unsigned char* ah = sm + P.off_a() + sa * A_STAGE;
                float my_sq = 0.f;
#pragma unroll 1
                for (int g4 = 0; g4 < 4; ++g4) {          // 4 row pairs per trip; trips 0,1: upper half, 2,3: lower half
                    const uint32_t gx = 2u * g + (uint32_t)(g4 >> 1), sx = gx % XS, phx = (gx / XS) & 1u;
                    if ((g4 & 1) == 0) mbar_wait(bar(WB_XF + sx), phx);
                    const unsigned char* xs = sm + P.off_x() + sx * X_STAGE;
                    float4 v[4];
                    float sq[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        v[u] = *reinterpret_cast<const float4*>(xs + (row_of(g4, u) & 63) * 256 + q4 * 16);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        *reinterpret_cast<uint2*>(ah + sw128_off((uint32_t)row_of(g4, u), (uint32_t)q4 * 4)) =
                            make_uint2(pack_bf16(v[u].x, v[u].y), pack_bf16(v[u].z, v[u].w));
                        sq[u] = fmaf(v[u].x, v[u].x, fmaf(v[u].y, v[u].y, fmaf(v[u].z, v[u].z, v[u].w * v[u].w)));
                    }
                    {   // transposing butterfly: lane (half, q4) ends with the sum of the row of trip q4 >> 2, slot q4 & 3
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
                    if (g4 & 1) {                 // this half of the block is in registers / the A stage: free its x stage
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar(WB_XE + sx));
                    }
                }
                row_sq += my_sq;
                if (b == DB - 1) {
                    const int r = row_of(q4 >> 2, q4 & 3);
                    const float nx = sqrtf(row_sq);
                    float o1, o2, o3;
                    split3(row_sq * 1.001953125f, o1, o2, o3);            // off_i = ||x||^2 (1 + 2^-9)
                    const float nxu = bf16_round(nx * 1.0078125f);        // ||x|| rounded up
                    *reinterpret_cast<uint4*>(sm + P.off_am() + (it & 1u) * AM_STAGE + sw32_chunk_off((uint32_t)r, 0)) =
                        make_uint4(pack_bf16(1.f, 1.f), pack_bf16(1.f, o1), pack_bf16(o2, o3), pack_bf16(nxu, 1.f));
                    rownorm_s[(it % NORM_RING) * TILE_M + r] = nx;
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) arrive_mma_side(WB_AF + sa);
            }
        }
    } else if (warp < W_OUT || helper) {
        // ================= epilogue: TMEM -> two-class min scan -> certified arg-min (as tc::k_vq_tc) =============
        // setmaxnreg draws from the CTA's LAUNCH allocation (768 x 80 registers): four scanning warpgroups get 104 each
        // (4 x 128 x 104 + 2 x 128 x 24 = 59 392 <= 61 440); one immediate for every path that reaches the scan code
        if constexpr (PRE && DB == 2) reg_inc<104>(); else reg_inc<128>();
        const int g = warp >> 2;                 // scan group 0 .. NGRP-1; group 1 merges and certifies
        const int wq = warp & 3;
        const int row_in_tile = wq * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16);
        for (uint32_t it = 0; it < n_iter; ++it) {
            const int64_t t = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
            const int64_t grow = t * TILE_M + row_in_tile;
            float rA[16], rB[16];
#pragma unroll
            for (int a = 0; a < 16; ++a) { rA[a] = INFINITY; rB[a] = INFINITY; }
            float* dbg = (DBG && p.dbg_scores && grow < p.n_rows) ? p.dbg_scores + grow * p.K_total + p.code_base : nullptr;
            {   // first unit of the group
                const uint32_t uc = it * (uint32_t)U + (uint32_t)g, buf = uc % NBUF, pht = (uc / NBUF) & 1u;
                mbar_wait(bar(WB_TF + buf), pht);
                tc_fence_after();
                scan_buffer<0, DBG>(lane_base + buf * UNIT_N, rA, rB, dbg ? dbg + g * UNIT_N : nullptr);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_mma_side(WB_TE + buf);
            }
            if (U == 4 && !helper) {   // second unit (codes 128*(g+2) ..)
                const uint32_t uc = it * 4u + (uint32_t)g + 2u, buf = uc % NBUF, pht = (uc / NBUF) & 1u;
                mbar_wait(bar(WB_TF + buf), pht);
                tc_fence_after();
                scan_buffer<4, DBG>(lane_base + buf * UNIT_N, rA, rB, dbg ? dbg + (g + 2) * UNIT_N - 4 * 32 : nullptr);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_mma_side(WB_TE + buf);
            }
            float m1, m2;
            int k1;
            {
                const int v = scan_finish(rA, rB, m1, m2);        // virtual column 0..255 of this group
                k1 = (v < UNIT_N ? g : g + 2) * UNIT_N + (v & (UNIT_N - 1));
            }
            {
                const uint32_t ps = it % NSLOT, php = (it / NSLOT) & 1u;
                if (g != 1) {                    // hand this group's result to group 1 (slot 0: group 0, slots 1, 2: helpers)
                    mbar_wait(bar(WB_PE + ps), php ^ 1u);
                    part_s[(ps * NP + (g == 0 ? 0 : g - 1)) * TILE_M + row_in_tile] = make_float4(m1, m2, __int_as_float(k1), 0.f);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar(WB_PF + ps));
                    continue;
                }
                mbar_wait(bar(WB_PF + ps), php);
                float4 o[NP];
#pragma unroll
                for (int q = 0; q < NP; ++q) o[q] = part_s[(ps * NP + q) * TILE_M + row_in_tile];
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(WB_PE + ps));
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    if (q >= NGRP - 1) break;
                    const int ko = __float_as_int(o[q].z);
                    m2 = fminf(fminf(o[q].y, m2), fmaxf(o[q].x, m1));
                    const bool take_other = (o[q].x < m1) || (o[q].x == m1 && ko < k1);
                    if (take_other) { m1 = o[q].x; k1 = ko; }
                }
            }
            const float xn = rownorm_s[(it % NORM_RING) * TILE_M + row_in_tile];
            float en = ENORM_S ? enorm_s[k1 < KL ? k1 : 0] : __ldg(reinterpret_cast<const float*>(p.image + wimage_off_enorm(KL, DB)) + (k1 < KL ? k1 : 0));
            const bool in_range = grow < p.n_rows;
            k1 += p.code_base;
            bool badrow = cb_bad;
            if (p.partial) {                     // sliced codebook: fold in the earlier slices / hand on to the later ones
                if (!p.pass_first && in_range) {
                    const float4 o = p.partial[grow];
                    const int ko = __float_as_int(o.z);
                    badrow = badrow || ((__float_as_uint(o.w) >> 31) != 0u);
                    m2 = fminf(fminf(o.y, m2), fmaxf(o.x, m1));
                    if ((o.x < m1) || (o.x == m1 && ko < k1)) { m1 = o.x; k1 = ko; en = fabsf(o.w); }
                }
                if (!p.pass_last && in_range)
                    p.partial[grow] = make_float4(m1, m2, __int_as_float(k1), badrow ? __uint_as_float(__float_as_uint(en) | 0x80000000u) : en);
            }
            const float need = 2.f * (p.cA * BOUND_UP * xn * en + p.cB * (BOUND_UP * en * en + xn * xn));
            const bool certified = ((m2 - m1) > need) && (xn < 1.0e18f) && !badrow;   // NaN -> false
            if (helper) continue;                // no outputs in this pass: the output warps are scanning
            const uint32_t rs = it % RES_RING, phr = (it / RES_RING) & 1u;
            mbar_wait(bar(WB_RE + rs), phr ^ 1u);
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
                if (code == -1) p.flagged_rows[basei + __popc(fl & ((1u << lane) - 1u))] = (int)(grow + p.row_base);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(WB_RF + rs));
        }
    } else {
        // ================= output: gather, straight-through value, loss ==========================================
        // 8 warps x 16 rows; a half-warp covers one 64-dim block of one row per instruction (16 lanes x 16 bytes)
        reg_dec<72>();
        const int ow = warp - W_OUT;
        const int half = lane >> 4, q4 = lane & 15;
        float dacc = 0.f;
        constexpr int RQ = D / 4;                // float4 per row
        const float4* x4 = reinterpret_cast<const float4*>(p.x);
        float4* o4 = reinterpret_cast<float4*>(p.quantize);
        const float4* cb4 = reinterpret_cast<const float4*>(p.cbT);
        for (uint32_t it = 0; it < n_iter; ++it) {
            const int64_t t = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
            const uint32_t rs = it % RES_RING, phr = (it / RES_RING) & 1u;
            mbar_wait(bar(WB_RF + rs), phr);
            const int* cs = codes_s + rs * TILE_M + ow * 16 + half;
            const size_t row0 = (size_t)t * TILE_M + ow * 16 + half;
#pragma unroll
            for (int b2 = 0; b2 < 2; ++b2) {
                int kk[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) kk[i] = cs[b2 * 8 + i * 2];
#pragma unroll
                for (int blk = 0; blk < DB; ++blk) {
                    float4 xv[4], qv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (kk[i] >= 0) {
                            xv[i] = __ldcg(x4 + (row0 + b2 * 8 + i * 2) * RQ + blk * 16 + q4);
                            qv[i] = __ldcg(cb4 + (size_t)kk[i] * RQ + blk * 16 + q4);
                        }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (kk[i] < 0) continue;
                        float4 d, o;
                        d.x = qv[i].x - xv[i].x; d.y = qv[i].y - xv[i].y; d.z = qv[i].z - xv[i].z; d.w = qv[i].w - xv[i].w;
                        o.x = xv[i].x + d.x; o.y = xv[i].y + d.y; o.z = xv[i].z + d.z; o.w = xv[i].w + d.w;
                        dacc = fmaf(d.x, d.x, fmaf(d.y, d.y, fmaf(d.z, d.z, fmaf(d.w, d.w, dacc))));
                        if (o4) __stcs(o4 + (row0 + b2 * 8 + i * 2) * RQ + blk * 16 + q4, o);
                        if (p.stat_sums) {
                            red_add_v4(p.stat_sums + (size_t)kk[i] * D + blk * 64 + q4 * 4, xv[i]);
                            if (q4 == 0 && blk == 0) red_add_f32(p.stat_counts + kk[i], 1.0f);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(WB_RE + rs));
        }
        if (p.diff_acc) {
            dacc = warp_sum(dacc);
            if (lane == 0) atomicAdd(p.diff_acc, (double)dacc);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CTA2) cluster_sync_all();                 // nobody leaves while the pair still reads its smem / arrives on its barriers
    if (warp == W_MMA) { if (CTA2) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

}  // namespace tcw

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
// codes per launch: the bf16 operand image of a launch is 128 KB of shared memory
// CTA-pair variant (VQB200_TCW_CTA2, default on): D = 256, K = 512 -- the deep fork's quantizers (vqvae_deep.py:252,257) -- keeps
// all 512 codes resident over a pair of SMs (128 KB of image per CTA) and runs ONE pass over x instead of two
// (larger codebooks at D = 256: 512-code pair passes -- K = 1024 as two converting passes instead of four streamed 256-code
//  passes, K >= 2048 as streamed pair passes: half as many passes and half the operand-image reads per MMA and SM)
inline bool tcw_pair_available() {
    // one probe per process: can a cluster of two CTAs of the pair kernel (its full shared-memory footprint) be co-scheduled here?
    // (an SM partition without whole SM pairs, or a driver without cluster launch, keeps the single-CTA passes)
    static const bool ok = [] {
        auto kern = tcw::k_vq_tcw<4, 2, false, false, true>;
        const tcw::Plan P{512, 4, 2, tcw::AS_CONV, tcw::n_partials(false, 4), tcw::n_part_slots(false, 4), tcw::enorm_in_smem(false, 4) ? 1 : 0, 2};
        const int smem = (int)P.total();
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) { (void)cudaGetLastError(); return false; }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2); cfg.blockDim = dim3(tc::THREADS); cfg.dynamicSmemBytes = (size_t)smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { (void)cudaGetLastError(); return false; }
        return n > 0;
    }();
    return ok;
}
inline bool tcw_pair(int dim, int n_embed) {
    static const bool on = [] { const char* e = getenv("VQB200_TCW_CTA2"); return e ? atoi(e) != 0 : true; }() && tcw_pair_available();
    static const int kmax = [] { const char* e = getenv("VQB200_TCW_CTA2_KMAX"); return e ? atoi(e) : 16384; }();
    // D = 128 streamed passes as pairs: built, parity-green and measured -- no gain (K = 2048: 377 vs 375 us, K = 8192: 1150 vs
    // 1164 us; those passes are bound by the scan, not by operand reads; eight A stages in the freed shared memory instead of
    // four did not help either: 1193 vs 1167 us), so it stays opt-in
    static const bool d128 = [] { const char* e = getenv("VQB200_TCW_CTA2_D128"); return e ? atoi(e) != 0 : false; }();
    if (on && d128 && dim == 128 && n_embed % 512 == 0 && n_embed / 512 > 2) return true;     // streamed passes (K >= 2048)
    return on && dim == 256 && n_embed % 512 == 0 && n_embed <= kmax;
}
inline int tcw_slice(int dim, int n_embed) {                  // codes per launch: 2 or 4 units of 128 codes
    return ((dim == 128 && n_embed >= 512) || tcw_pair(dim, n_embed)) ? 512 : 256;
}
inline bool tcw_shape_ok(int dim, int n_embed) {
    if (dim != 128 && dim != 256) return false;
    if (n_embed < 256 || n_embed > 16384) return false;
    return n_embed % tcw_slice(dim, n_embed) == 0;
}
inline bool tcw_supported(const RowLayout& L, const float* x, int dim, int n_embed) {
    if (!tcw_shape_ok(dim, n_embed) || L.n_rows < 1) return false;
    static const bool disabled = getenv("VQB200_DISABLE_TC") != nullptr || getenv("VQB200_DISABLE_TCW") != nullptr;
    if (disabled) return false;
    return tc_layout_dense(L, x, dim);
}

inline int tcw_encode_tmap(CUtensorMap* tm, const float* x, int64_t n_rows, int dim) {
    static PFN_cuTensorMapEncodeTiled encode = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) f = nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(f);
    }();
    if (!encode) return 1;
    cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)n_rows};
    cuuint64_t gstr[1] = {(cuuint64_t)dim * 4u};
    cuuint32_t box[2] = {64u, (cuuint32_t)tcw::X_ROWS};
    cuuint32_t estr[2] = {1u, 1u};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}

template <int DB, int XS, bool DBG, bool PRE, bool CTA2 = false>
inline int tcw_launch(const tcw::WParams& prm, cudaStream_t st) {
    auto kern = tcw::k_vq_tcw<DB, XS, DBG, PRE, CTA2>;
    const tcw::Plan P{prm.KL, DB, PRE ? 0 : XS, PRE ? tcw::as_pre(DB) : tcw::AS_CONV, tcw::n_partials(PRE, DB), tcw::n_part_slots(PRE, DB),
                       tcw::enorm_in_smem(PRE, DB) ? 1 : 0, CTA2 ? 2 : 1};
    const int smem = (int)P.total();
    // opt-in shared-memory size is a per-device function attribute: cache it per device (several GPUs in one process)
    static std::atomic<int> configured_dev[64];          // (atomic: several host threads / GPUs per process)
    int dev_id = 0;
    if (cudaGetDevice(&dev_id) != cudaSuccess || dev_id < 0 || dev_id >= 64) dev_id = 0, configured_dev[0] = 0;
    std::atomic<int>& configured = configured_dev[dev_id];
    if (configured.load(std::memory_order_relaxed) < smem) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
            fprintf(stderr, "vqb200: wide tensor-core kernel needs %d bytes of shared memory (DB=%d XS=%d KL=%d)\n", smem, DB, XS, prm.KL);
            return 1;
        }
        configured = smem;
    }
    cudaError_t e;
    if (!CTA2) {
        e = launch_pdl(kern, dim3((unsigned)tc_grid(prm.n_rows, false)), dim3(tc::THREADS), (size_t)smem, st, prm);
    } else {                                     // clusters of two CTAs (one SM pair each), programmatic dependent launch as everywhere
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)tc_grid(prm.n_rows, true));
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
        e = cudaLaunchKernelEx(&cfg, kern, prm);
    }
    if (e != cudaSuccess) fprintf(stderr, "vqb200: wide tensor-core kernel launch failed: %s (smem %d)\n", cudaGetErrorString(e), smem);
    return e != cudaSuccess;
}

// operand image(s) of the wide engine from the fp32 code-major copy
inline cudaError_t tcw_prepare(const CodebookImage& cb, int dim, int n_embed, cudaStream_t st) {
    const int threads = n_embed * (dim / 8);
    return launch_pdl(tcw::k_prepare_wide, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, st, (const float*)cb.cbT, (const float*)cb.ee,
                      cb.tc, n_embed, dim, tcw_slice(dim, n_embed), tc::bound_cA(1), tcw::bound_cB(dim / 64));
}

// main kernel only (one launch per slice); the caller runs the exact fix-up over the flagged rows afterwards
inline int tcw_forward(const float* x, const RowLayout& L, int dim, int n_embed, const CodebookImage& cb, float* quantize,
                       int64_t* embed_ind, const ForwardScratch& sc, double* diff_acc, float* sums, float* counts,
                       float* dbg_scores, cudaStream_t st, unsigned long long* n_launches = nullptr) {
    const int DB = dim / 64, KL = tcw_slice(dim, n_embed), n_slices = n_embed / KL;
    // sliced codebook: convert x once (k_convert_wide), then stream ready-made operand stages.  Measured on B200 at
    // N = 524 288: with two slices the extra pass over x costs more than converting twice (D = 256, K = 512: 0.53 vs 0.49 ms),
    // from four slices on it wins (D = 256, K = 8192: 6.15 -> 2.68 ms)
    const bool pre = n_slices > 2;
    if (pre && (!sc.partial || !sc.wide_a)) return 1;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (!pre && tcw_encode_tmap(&tmap, x, L.n_rows, dim)) {
        fprintf(stderr, "vqb200: cuTensorMapEncodeTiled failed for x [%lld, %d]\n", (long long)L.n_rows, dim);
        return 1;
    }
    // Row chunks whose bf16 image fits in L2 (all slices of a chunk back to back, so that only the first slice reads it from
    // HBM) were measured and are OFF by default: 48 / 64 / 96 MB chunks were 5-20 % slower than whole-batch passes (more
    // wave tails and launches; the A stream is bound by the 64 KB of stages in flight per SM, not by HBM).
    static const int64_t chunk_mb = [] { const char* e = getenv("VQB200_TCW_CHUNK_MB"); return (int64_t)(e ? atoi(e) : 0); }();
    int64_t chunk_rows = L.n_rows;
    if (pre && chunk_mb > 0) {
        chunk_rows = std::max<int64_t>(tc::TILE_M, (chunk_mb << 20) / ((int64_t)dim * 2) / tc::TILE_M * tc::TILE_M);
        const int64_t per_wave = (int64_t)tc_num_sms() * tc::TILE_M;             // whole waves of tiles
        if (chunk_rows > per_wave) chunk_rows = chunk_rows / per_wave * per_wave;
    }
    for (int64_t r0 = 0; r0 < L.n_rows; r0 += chunk_rows) {
        const int64_t rows = std::min(chunk_rows, L.n_rows - r0);
        const int64_t t0 = r0 / tc::TILE_M;
        if (pre) {
            const int64_t tiles = (rows + tc::TILE_M - 1) / tc::TILE_M;
            const unsigned grid = (unsigned)std::min<int64_t>(tiles * 8, (int64_t)tc_num_sms() * 16);
            unsigned char* a = sc.wide_a + (size_t)t0 * DB * tcw::A_STAGE;
            unsigned char* m = sc.wide_m + (size_t)t0 * tcw::AM_STAGE;
            float* nr = sc.wide_norm + (size_t)t0 * tc::TILE_M;
            cudaError_t e = DB == 2 ? launch_pdl(tcw::k_convert_wide<2>, dim3(grid), dim3(256), 0, st, x + r0 * dim, rows, a, m, nr)
                                    : launch_pdl(tcw::k_convert_wide<4>, dim3(grid), dim3(256), 0, st, x + r0 * dim, rows, a, m, nr);
            if (e != cudaSuccess) return 1;
            if (n_launches) ++*n_launches;
        }
        for (int sl = 0; sl < n_slices; ++sl) {
            tcw::WParams prm;
            prm.tmap = tmap;
            prm.x = x + r0 * dim; prm.n_rows = rows; prm.KL = KL; prm.K_total = n_embed; prm.row_base = r0;
            prm.image = cb.tc + (size_t)sl * tcw::wimage_bytes(KL, DB); prm.cbT = cb.cbT;
            prm.quantize = quantize ? quantize + r0 * dim : nullptr; prm.embed_ind = embed_ind + r0; prm.diff_acc = diff_acc;
            prm.stat_sums = sums; prm.stat_counts = counts;
            prm.flagged_count = sc.flagged_count; prm.flagged_rows = sc.flagged_rows;
            prm.dbg_scores = dbg_scores ? dbg_scores + r0 * n_embed : nullptr;
            prm.cA = tc::bound_cA(1); prm.cB = tcw::bound_cB(DB);
            prm.code_base = sl * KL; prm.partial = n_slices > 1 ? sc.partial + r0 : nullptr;
            prm.pass_first = sl == 0; prm.pass_last = sl == n_slices - 1;
            prm.a_img = pre ? sc.wide_a + (size_t)t0 * DB * tcw::A_STAGE : nullptr;
            prm.m_img = pre ? sc.wide_m + (size_t)t0 * tcw::AM_STAGE : nullptr;
            prm.norms = pre ? sc.wide_norm + (size_t)t0 * tc::TILE_M : nullptr;
            int rc;
            if (pre) {
                if (DB == 2 && tcw_pair(dim, n_embed)) rc = dbg_scores ? tcw_launch<2, 1, true, true, true>(prm, st) : tcw_launch<2, 1, false, true, true>(prm, st);
                else if (DB == 2) rc = dbg_scores ? tcw_launch<2, 1, true, true>(prm, st) : tcw_launch<2, 1, false, true>(prm, st);
                else if (KL == 512) rc = dbg_scores ? tcw_launch<4, 1, true, true, true>(prm, st) : tcw_launch<4, 1, false, true, true>(prm, st);
                else rc = dbg_scores ? tcw_launch<4, 1, true, true>(prm, st) : tcw_launch<4, 1, false, true>(prm, st);
            } else if (DB == 2) {
                // (D = 128 converting passes as CTA pairs with four x stages instead of two were measured: assign 123 -> 120 us, eval
                //  forward 142 -> 157 us, training step 223 -> 230 us at K = 512 -- the converters, not the x loads, bound them: not kept)
                if (KL == 512) rc = dbg_scores ? tcw_launch<2, 2, true, false>(prm, st) : tcw_launch<2, 2, false, false>(prm, st);
                else rc = dbg_scores ? tcw_launch<2, 4, true, false>(prm, st) : tcw_launch<2, 4, false, false>(prm, st);
            } else if (KL == 512) {              // D = 256, all 512 codes resident over a CTA pair
                rc = dbg_scores ? tcw_launch<4, 2, true, false, true>(prm, st) : tcw_launch<4, 2, false, false, true>(prm, st);
            } else {
                rc = dbg_scores ? tcw_launch<4, 2, true, false>(prm, st) : tcw_launch<4, 2, false, false>(prm, st);
            }
            if (rc) return rc;
            if (n_launches) ++*n_launches;
        }
    }
    return 0;
}

}  // namespace vqb200
