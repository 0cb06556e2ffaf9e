// Shared device/host helpers for the vqb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vqb200 {

// Physical placement of the N logical rows of a [..., D] tensor (see include/vqb200.h "vq_layout").
struct RowLayout {
    int64_t n_rows;
    int64_t rows_per_image;
    int64_t image_stride;
    int64_t row_stride;
    int64_t col_stride;
};

__host__ __device__ __forceinline__ int64_t row_offset(const RowLayout& L, int64_t n) {
    int64_t img = n / L.rows_per_image;
    int64_t r = n - img * L.rows_per_image;
    return img * L.image_stride + r * L.row_stride;
}

// Prepared codebook image (what vqb200_codebook_prepare / vqb200_ema_update maintain).
//   cbT   [K][D]   fp32, code-major: row k is code e_k (coalesced gather, exact re-score operand)
//   ee    [K]      fp32 ||e_k||^2, summed in a fixed order
//   tc    tensor-core operand image + per-code norms for the error bound (see tc_kernel.cuh)
struct CodebookImage {
    float* cbT;
    float* ee;
    float* enorm_max;   // 1 float: max_k ||e_k|| over "near" codes (tcgen05 bound), [1]=far threshold
    unsigned char* tc;  // bf16 operand image, 1024-byte aligned
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// tensor-core operand image (tc_kernel.cuh): per code 2 x dim bf16 (hi, lo) + 32 B misc row + fp32 norm + 32 B misc row of
// the plain-bf16 bound, then the tf32 operand (dim fp32 words, low 13 bits clear) + its 32 B misc row
__host__ __device__ inline size_t tc_image_bytes(int dim, int n_embed) {
    size_t dpad = (size_t)(dim + 63) / 64 * 64;
    return (size_t)n_embed * (dpad * 4 + 32 + 4 + 32 + dpad * 4 + 32) + 1024;
}

__host__ __device__ inline size_t codebook_bytes(int dim, int n_embed) {
    size_t b = align_up((size_t)n_embed * dim * 4, 1024);
    b += align_up((size_t)n_embed * 4, 1024);
    b += 1024;                                   // scalars
    b += align_up(tc_image_bytes(dim, n_embed), 1024);
    return b;
}

__host__ __device__ inline CodebookImage codebook_view(void* base, int dim, int n_embed) {
    unsigned char* p = (unsigned char*)base;
    CodebookImage v;
    v.cbT = (float*)p;            p += align_up((size_t)n_embed * dim * 4, 1024);
    v.ee = (float*)p;             p += align_up((size_t)n_embed * 4, 1024);
    v.enorm_max = (float*)p;      p += 1024;
    v.tc = p;
    return v;
}

// per-call scratch of the forward
struct ForwardScratch {
    double* diff_acc;      // 1 double: sum (q-x)^2
    int* flagged_count;    // 1 int: rows sent to the exact re-score
    unsigned int* ticket;  // 1 word: last-block-done ticket of the fix-up / gather kernels (kept zero between launches)
    unsigned int* ema_ticket;  // 1 word: last-block-done ticket of the fold + EMA kernel
    unsigned int* n_parts;     // 1 word: number of per-CTA statistics tables the statistics kernel of this call wrote
    int* code_counts;      // [n_embed] rows per code of this call (integer atomics of the statistics kernel; cleared with the header)
    unsigned int* fix_tickets;         // [FIX_CAP / 64] per-chunk tickets of the code-block-parallel fix-up (cleared with the header)
    unsigned long long* fix_partial;   // [min(n_rows, FIX_CAP)][FIX_KB] partial arg-min keys of the flagged rows
    int* flagged_rows;     // [n_rows] row ids
    float4* partial;       // [n_rows] running (m1, m2, winner, ||e_winner||) of the sliced tensor-core engine (n_embed > 512), else null
    float* stat_partials;  // [STAT_PARTS][K*(D+1)] per-CTA statistics tables
    // sliced WIDE tensor-core engine (dim 128 / 256, tc_wide_kernel.cuh): x converted once per call, per 128-row tile
    unsigned char* wide_a; // [tiles][dim/64][16384] bf16 operand blocks (128-byte swizzle image)
    unsigned char* wide_m; // [tiles][4096] misc rows (32-byte swizzle image)
    float* wide_norm;      // [tiles][128] ||x_row||
};
constexpr int STAT_PARTS = 160;
// fix-up of few flagged rows: the exact re-score of a 64-row chunk is split over the 64-code blocks of the codebook (up to
// FIX_KB of them) so that a handful of chunks still fills the GPU; at most FIX_CAP rows take that route
constexpr int FIX_CAP = 65536, FIX_KB = 8;
__host__ __device__ inline size_t scratch_header_bytes(int n_embed) {     // header + code counters + fix-up tickets (zeroed per call)
    return 256 + align_up((size_t)n_embed * 4, 256) + (size_t)(FIX_CAP / 64) * 4;
}
__host__ __device__ inline bool scratch_has_partial(int dim, int n_embed) {
    return (dim == 64 && n_embed > 512) || (dim == 128 && n_embed > 512) || (dim == 256 && n_embed > 256);   // sliced tensor-core launches
}
__host__ __device__ inline bool scratch_has_wide(int dim, int n_embed) { return dim != 64 && scratch_has_partial(dim, n_embed); }
__host__ __device__ inline size_t scratch_wide_bytes(int64_t n_rows, int dim) {
    const size_t tiles = (size_t)((n_rows + 127) / 128);
    return tiles * ((size_t)(dim / 64) * 16384 + 4096 + 512);
}
// per-CTA statistics tables exist only for shapes whose [K][D] table fits the statistics kernel's shared memory (other shapes
// use global atomics): 21 MB at D = 64, K = 512 -- and nothing (instead of 1.35 GB) at D = 256, K = 8192
__host__ __device__ inline bool scratch_has_stat_tables(int dim, int n_embed) {
    return (size_t)n_embed * dim * 4 + (size_t)(4 * n_embed + 2) * 4 + 4096 * 12 + 16 <= 200 * 1024 && n_embed <= 65535;   // = code_stats_smem_bytes <= 200 KB
}
__host__ __device__ inline int64_t std_min64(int64_t a, int64_t b) { return a < b ? a : b; }
__host__ __device__ inline size_t forward_scratch_bytes(int64_t n_rows, int dim, int n_embed) {
    return scratch_header_bytes(n_embed) + (size_t)std_min64(n_rows, FIX_CAP) * FIX_KB * 8 + align_up((size_t)n_rows * 4, 256) +
           (scratch_has_partial(dim, n_embed) ? align_up((size_t)n_rows * 16, 256) : 0) +
           (scratch_has_wide(dim, n_embed) ? scratch_wide_bytes(n_rows, dim) : 0) +
           (scratch_has_stat_tables(dim, n_embed) ? (size_t)STAT_PARTS * n_embed * (dim + 1) * 4 : 0);
}
__host__ __device__ inline ForwardScratch scratch_view(void* base, int64_t n_rows, int dim, int n_embed) {
    unsigned char* p = (unsigned char*)base;
    ForwardScratch s;
    s.diff_acc = (double*)p;
    s.flagged_count = (int*)(p + 16);
    s.ticket = (unsigned int*)(p + 32);
    s.ema_ticket = (unsigned int*)(p + 48);
    s.n_parts = (unsigned int*)(p + 52);
    s.code_counts = (int*)(p + 256);
    s.fix_tickets = (unsigned int*)(p + 256 + align_up((size_t)n_embed * 4, 256));
    p += scratch_header_bytes(n_embed);
    s.fix_partial = (unsigned long long*)p;
    p += (size_t)std_min64(n_rows, FIX_CAP) * FIX_KB * 8;
    s.flagged_rows = (int*)p;
    p += align_up((size_t)n_rows * 4, 256);
    s.partial = nullptr;
    if (scratch_has_partial(dim, n_embed)) { s.partial = (float4*)p; p += align_up((size_t)n_rows * 16, 256); }
    s.wide_a = nullptr; s.wide_m = nullptr; s.wide_norm = nullptr;
    if (scratch_has_wide(dim, n_embed)) {
        const size_t tiles = (size_t)((n_rows + 127) / 128);
        s.wide_a = p;            p += tiles * (size_t)(dim / 64) * 16384;
        s.wide_m = p;            p += tiles * 4096;
        s.wide_norm = (float*)p; p += tiles * 512;
    }
    s.stat_partials = (float*)p;
    return s;
}

// Programmatic dependent launch (PDL): the kernels of one training step form a chain on one stream; each is launched with
// programmatic stream serialisation, announces at its top that its dependents may be scheduled (pdl_trigger) and waits
// for the completion + memory flush of its own prerequisite before touching any global memory (pdl_wait).  This only
// hides launch latency between the 6 launches of a step (~1-2 us each); both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): a lane that owns 32 contiguous bytes of a row touches ONE full sector per
// instruction instead of two half sectors -- the per-lane row accesses of the NCHW variant are bound by L1TEX sector throughput
__device__ __forceinline__ void ldg_nc_v8(const float* p, float (&v)[8]) {     // p 32-byte aligned
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ void stg_v8(float* p, const float (&v)[8]) {        // p 32-byte aligned
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace vqb200
