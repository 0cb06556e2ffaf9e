// Code statistics of the D = 64 shapes (vqvae.py:50,55-56 without a one-hot, without floating-point atomics and without a
// sort): the second pass over x, fed by the TMA engine.
//
// One persistent CTA per SM owns a PRIVATE [K][64] fp32 table in shared memory (128 KB at K = 512) and streams 64-row
// tiles of x through a ring of shared-memory stages (bulk copies for dense rows, 3-D tensor-map boxes with the 128-byte
// swizzle for the NCHW-physical view of vqvae.py:227,235 -- no dense side copy of x is needed any more).
//   warp 0      ROUTER + producer: arms a free stage, starts the tile's copy, and deals the tile's 64 rows to their OWNER
//               warps (owner = 1 + code mod 31) as per-owner row lists in shared memory (ranks from match.any: no atomics,
//               fixed order).  It runs up to STAGES tiles ahead of the consumers.
//   warps 1-31  CONSUMERS: a warp walks its list of the tile, lanes = dims (one conflict-free 128-byte wavefront per half
//               row for dense rows; 4 wavefronts for the swizzled x^T stage), accumulates runs of the same code in
//               registers and adds to the table row when the code changes.  A code has exactly one owner, so every table
//               update is a plain read-modify-write, and the summation order is fixed for a fixed grid: the statistics
//               are bit-reproducible run to run (the reference's cuBLAS GEMM is not).
// Tiles are walked from the END of x: the assignment kernel touched those rows last (L2 evict_last on its x tiles).
#pragma once
#include <cudaTypedefs.h>

#include "common.cuh"
#include "tc_kernel.cuh"

namespace vqb200 {
namespace st {

constexpr int THREADS = 1024, WARPS = 32, CONSUMERS = 31;
constexpr int TILE = 64, STAGES = 5;
constexpr uint32_t STAGE_BYTES = TILE * 64 * 4;

struct Params {
    alignas(64) CUtensorMap tmap;    // NCHW only: x as a 3-D tensor [image][dim][row-in-image], box = 32 rows x 64 dims, SWIZZLE_128B
    const float* x;
    int64_t n_rows;
    int K;
    const int64_t* embed_ind;
    float* partials;                 // [gridDim.x][K * 65]: K*64 sums (code-major) then K counts, as k_stats_fold expects
    int64_t rpi;                     // NCHW: rows per image (multiple of 128)
};

// ring | table [K][64] | cnt [K] | lists [STAGES][32][64] u8 | codes [STAGES][64] u16 | counts [STAGES][32] | barriers
__host__ __device__ inline size_t smem_bytes(int K) {
    return (size_t)STAGES * STAGE_BYTES + (size_t)K * 260 + (size_t)STAGES * (32 * TILE + TILE * 2 + 32 * 4) + 3 * STAGES * 8 + 16 + 1024;
}

__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

template <bool NCHW>
__global__ void __launch_bounds__(THREADS, 1) k_code_stats_tma(const __grid_constant__ Params p) {
    extern __shared__ unsigned char st_smem_raw[];
    const uint32_t raw = tc::smem_u32(st_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = st_smem_raw + (base - raw);
    const int K = p.K;
    float* table = reinterpret_cast<float*>(sm + STAGES * STAGE_BYTES);                 // [K][64]
    int* cnt = reinterpret_cast<int*>(table + (size_t)K * 64);                           // [K]
    int* counts = cnt + K;                                                              // [STAGES][32] rows dealt to each owner
    unsigned short* codes_s = reinterpret_cast<unsigned short*>(counts + STAGES * 32);  // [STAGES][64]
    unsigned char* lists = reinterpret_cast<unsigned char*>(codes_s + STAGES * TILE);   // [STAGES][32][64] row numbers
    const uint32_t bars = (tc::smem_u32(lists + STAGES * 32 * TILE) + 7u) & ~7u;        // full | routed | free, STAGES each
    auto bar_full = [&](int s) { return bars + 8u * (uint32_t)s; };
    auto bar_routed = [&](int s) { return bars + 8u * (uint32_t)(STAGES + s); };
    auto bar_free = [&](int s) { return bars + 8u * (uint32_t)(2 * STAGES + s); };
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    pdl_trigger();
    for (int i = tid; i < K * 16; i += THREADS) reinterpret_cast<float4*>(table)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < K; i += THREADS) cnt[i] = 0;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { tc::mbar_init(bar_full(s), 1); tc::mbar_init(bar_routed(s), 1); tc::mbar_init(bar_free(s), CONSUMERS); }
        tc::fence_barrier_init();
    }
    __syncthreads();
    pdl_wait();                                  // the indices (and x) of the upstream kernels are complete

    const int64_t n_tiles = (p.n_rows + TILE - 1) / TILE;
    const int64_t first = n_tiles - 1 - (int64_t)blockIdx.x;        // walk from the end of x
    const int n_my = first >= 0 ? (int)(first / gridDim.x) + 1 : 0;

    if (warp == 0) {
        // ================= router + producer =============================================================================
        const uint64_t pol = l2_policy_evict_first();               // last reader of x in the step
        auto load_codes = [&](int j, int (&code)[2]) {              // lane l: rows l and l + 32 of my j-th tile (-1 past the end)
            const int64_t r0 = (first - (int64_t)j * gridDim.x) * TILE;
            const int rows = j < n_my ? (int)min((int64_t)TILE, p.n_rows - r0) : 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = 32 * h + lane;
                const long long k = r < rows ? __ldcs(reinterpret_cast<const long long*>(p.embed_ind) + r0 + r) : -1ll;
                code[h] = (k >= 0 && k < K) ? (int)k : -1;        // (out-of-range indices cannot occur: written by our own kernels)
            }
        };
        int code[2], code_next[2];
        load_codes(0, code);
        for (int j = 0; j < n_my; ++j) {
            const int s = j % STAGES;
            load_codes(j + 1, code_next);        // in flight while this tile is dealt
            if (j >= STAGES) tc::mbar_wait(bar_free(s), (uint32_t)(j / STAGES - 1) & 1u);     // every consumer has left the stage
            if (lane == 0) {
                const int64_t r0 = (first - (int64_t)j * gridDim.x) * TILE;
                const uint32_t rows = (uint32_t)min((int64_t)TILE, p.n_rows - r0);
                const uint32_t dst = base + (uint32_t)s * STAGE_BYTES;
                tc::fence_async_smem();           // generic-proxy readers of the stage are done: order the async-proxy write behind them
                if (NCHW) {
                    tc::mbar_expect_tx(bar_full(s), STAGE_BYTES);
                    tc::tma_load_3d(dst, &p.tmap, (int)(r0 % p.rpi), 0, (int)(r0 / p.rpi), bar_full(s), pol);
                    tc::tma_load_3d(dst + 8192u, &p.tmap, (int)(r0 % p.rpi) + 32, 0, (int)(r0 / p.rpi), bar_full(s), pol);
                } else {
                    tc::mbar_expect_tx(bar_full(s), rows * 256u);
                    tc::bulk_g2s_hint(dst, p.x + r0 * 64, rows * 256u, bar_full(s), pol);
                }
            }
            counts[s * 32 + lane] = 0;
            __syncwarp();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = code[h];
                const int owner = k >= 0 ? 1 + k % CONSUMERS : 0;          // owner 0 = nobody (rows past the end)
                const unsigned m = __match_any_sync(0xffffffffu, owner);
                const int leader = __ffs(m) - 1;
                int pos = 0;
                if (lane == leader) { pos = counts[s * 32 + owner]; counts[s * 32 + owner] = pos + __popc(m); }
                pos = __shfl_sync(0xffffffffu, pos, leader) + __popc(m & ((1u << lane) - 1u));
                if (k >= 0) {
                    lists[(s * 32 + owner) * TILE + pos] = (unsigned char)(32 * h + lane);
                    codes_s[s * TILE + 32 * h + lane] = (unsigned short)k;
                }
                __syncwarp();
            }
            if (lane == 0) tc::mbar_arrive(bar_routed(s));          // release: the lists of this tile are visible
            code[0] = code_next[0]; code[1] = code_next[1];
        }
    } else {
        // ================= consumers ======================================================================================
        int cur = -1, run = 0;
        float a0 = 0.f, a1 = 0.f;
        auto flush = [&]() {
            if (cur >= 0) {
                float* t = table + (size_t)cur * 64;
                t[lane] += a0;
                t[lane + 32] += a1;
                if (lane == 0) cnt[cur] += run;
            }
        };
        auto eat = [&](int k, float v0, float v1) {
            if (k != cur) {
                flush();
                cur = k; a0 = 0.f; a1 = 0.f; run = 0;
            }
            a0 += v0;
            a1 += v1;
            ++run;
        };
        auto row_loads = [&](const unsigned char* xs, int r, float& v0, float& v1) {
            if (NCHW) {                          // x^T boxes of 32 rows: [dim][32 rows], 16-byte chunks XOR-swizzled by dim mod 8
                const uint32_t rr = (uint32_t)r & 31u;
                const uint32_t off = ((uint32_t)r >> 5) * 8192u + (uint32_t)lane * 128u + ((((rr >> 2) ^ ((uint32_t)lane & 7u)) << 4) | ((rr & 3u) << 2));
                v0 = *reinterpret_cast<const float*>(xs + off);
                v1 = *reinterpret_cast<const float*>(xs + off + 4096u);
            } else {
                const float* xr = reinterpret_cast<const float*>(xs) + r * 64;
                v0 = xr[lane];
                v1 = xr[lane + 32];
            }
        };
        for (int j = 0; j < n_my; ++j) {
            const int s = j % STAGES;
            const uint32_t ph = (uint32_t)(j / STAGES) & 1u;
            tc::mbar_wait(bar_routed(s), ph);
            const int n = counts[s * 32 + warp];
            if (n > 0) {
                tc::mbar_wait(bar_full(s), ph);
                const unsigned char* xs = sm + (size_t)s * STAGE_BYTES;
                const unsigned char* my = lists + (s * 32 + warp) * TILE;
                const unsigned short* cs = codes_s + s * TILE;
                int i = 0;
                for (; i + 2 <= n; i += 2) {      // two rows per trip: their shared-memory reads overlap
                    const int r0 = my[i], r1 = my[i + 1];
                    const int k0 = cs[r0], k1 = cs[r1];
                    float v0, v1, w0, w1;
                    row_loads(xs, r0, v0, v1);
                    row_loads(xs, r1, w0, w1);
                    eat(k0, v0, v1);
                    eat(k1, w0, w1);
                }
                if (i < n) {
                    const int r0 = my[i];
                    float v0, v1;
                    row_loads(xs, r0, v0, v1);
                    eat(cs[r0], v0, v1);
                }
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(bar_free(s));
        }
        flush();
    }
    __syncthreads();
    float* out = p.partials + (size_t)blockIdx.x * K * 65;
    for (int i = tid; i < K * 16; i += THREADS) reinterpret_cast<float4*>(out)[i] = reinterpret_cast<const float4*>(table)[i];
    for (int i = tid; i < K; i += THREADS) out[(size_t)K * 64 + i] = (float)cnt[i];
}

}  // namespace st

// x (NCHW-physical) as a 3-D tensor map for the statistics kernel: box = 32 rows x 64 dims, 128-byte swizzle
inline int st_encode_tmap(CUtensorMap* tm, const float* x, const RowLayout& L) {
    static PFN_cuTensorMapEncodeTiled encode = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) f = nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(f);
    }();
    if (!encode) return 1;
    const cuuint64_t n_img = (cuuint64_t)(L.n_rows / L.rows_per_image);
    const cuuint64_t img_stride = n_img > 1 ? (cuuint64_t)L.image_stride : (cuuint64_t)64 * (cuuint64_t)L.col_stride;
    cuuint64_t gdim[3] = {(cuuint64_t)L.rows_per_image, 64u, n_img};
    cuuint64_t gstr[2] = {(cuuint64_t)L.col_stride * 4u, img_stride * 4u};
    cuuint32_t box[3] = {32u, 64u, 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}

// shapes / layouts the TMA-fed statistics kernel covers: dim 64, table + ring within shared memory, dense rows or the
// NCHW-physical view in whole 128-row tiles (the layouts tc_layout_dense / tc_layout_nchw accept)
inline bool st_supported(const RowLayout& L, const float* x, int dim, int n_embed) {
    if (dim != 64 || n_embed < 1 || n_embed > 512 || L.n_rows < 1) return false;
    if (getenv("VQB200_DISABLE_STATS_TMA")) return false;
    return tc_layout_dense(L, x, dim) || tc_layout_nchw(L, x, dim);
}

// launches the kernel; *parts_out = number of per-CTA tables written to `partials` ([parts][K * 65] floats)
inline int st_launch(const float* x, const RowLayout& L, int n_embed, const int64_t* embed_ind, float* partials, int max_parts,
                     cudaStream_t stream, int* parts_out) {
    const bool nchw = !tc_layout_dense(L, x, 64);
    st::Params prm;
    memset(&prm.tmap, 0, sizeof(prm.tmap));
    if (nchw && st_encode_tmap(&prm.tmap, x, L)) return 1;
    prm.x = x; prm.n_rows = L.n_rows; prm.K = n_embed; prm.embed_ind = embed_ind; prm.partials = partials;
    prm.rpi = L.rows_per_image;
    const int64_t n_tiles = (L.n_rows + st::TILE - 1) / st::TILE;
    const int grid = (int)std::min<int64_t>(n_tiles, std::min(tc_num_sms(), max_parts));
    const int smem = (int)st::smem_bytes(n_embed);
    static int configured_dev[64][2] = {};
    int dev_id = 0;
    if (cudaGetDevice(&dev_id) != cudaSuccess || dev_id < 0 || dev_id >= 64) dev_id = 0, configured_dev[0][0] = configured_dev[0][1] = 0;
    int& configured = configured_dev[dev_id][nchw ? 1 : 0];
    if (configured < smem) {
        cudaError_t e = nchw ? cudaFuncSetAttribute(st::k_code_stats_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                             : cudaFuncSetAttribute(st::k_code_stats_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return 1;
        configured = smem;
    }
    if (parts_out) *parts_out = grid;
    cudaError_t e = nchw ? launch_pdl(st::k_code_stats_tma<true>, dim3((unsigned)grid), dim3(st::THREADS), (size_t)smem, stream, prm)
                         : launch_pdl(st::k_code_stats_tma<false>, dim3((unsigned)grid), dim3(st::THREADS), (size_t)smem, stream, prm);
    return e != cudaSuccess;
}

}  // namespace vqb200
