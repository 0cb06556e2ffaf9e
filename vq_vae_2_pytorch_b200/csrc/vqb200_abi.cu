// C ABI of the B200-native VQ-VAE-2 quantizer path (see include/vqb200.h for the contract).
// One translation unit: nvcc -gencode arch=compute_100a,code=sm_100a -shared -> libvqb200.so
#include "../../include/vqb200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>

#include "common.cuh"
#include "simt_kernels.cuh"
#include "tc_kernel.cuh"
#include "fused_kernels.cuh"
#include "tc_wide_kernel.cuh"

using namespace vqb200;

namespace {

thread_local int g_last_cuda_error = 0;
std::atomic<unsigned long long> g_launches{0};   // kernels this library launched (bench.py's gpu_launches)

inline int cuda_fail(cudaError_t e) {
    g_last_cuda_error = (int)e;
    return VQB200_ECUDA;
}
#define VQ_CUDA(call)                                     \
    do {                                                  \
        cudaError_t e__ = (call);                         \
        if (e__ != cudaSuccess) return cuda_fail(e__);    \
    } while (0)
#define VQ_LAUNCH_CHECK()              \
    do {                               \
        g_launches.fetch_add(1);       \
        VQ_CUDA(cudaGetLastError());   \
    } while (0)

// per (device, kernel slot): the largest opt-in dynamic shared-memory size already configured
bool smem_attr_needed(int slot, size_t bytes) {
    static std::atomic<size_t> configured[64][4];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    size_t cur = configured[dev][slot].load(std::memory_order_relaxed);
    if (cur >= bytes) return false;
    configured[dev][slot].store(bytes, std::memory_order_relaxed);
    return true;
}

bool layout_ok(int64_t n_rows, int32_t dim, int64_t rpi, int64_t img_stride, int64_t row_stride,
               int64_t col_stride) {
    if (n_rows < 0 || dim <= 0 || rpi <= 0 || row_stride <= 0 || col_stride <= 0) return false;
    if (n_rows > rpi && img_stride <= 0) return false;
    // the two layouts the streaming kernels coalesce for
    return col_stride == 1 || row_stride == 1;
}

int prepare_codebook(const float* d_embed, int dim, int n_embed, void* d_codebook, cudaStream_t st,
                     unsigned int* zero_header = nullptr, bool* header_zeroed = nullptr) {
    if (header_zeroed) *header_zeroed = false;
    CodebookImage cb = codebook_view(d_codebook, dim, n_embed);
    if (tc_any_ok(dim, n_embed)) {                // one launch: transpose + norms + tensor-core operand image(s)
        VQ_CUDA(launch_pdl(k_prepare64, dim3(n_embed / PREP_CODES), dim3(256), 0, st, d_embed, cb.cbT, cb.ee, cb.tc, n_embed,
                           tc_sliced_ok(dim, n_embed) ? TC_SLICE : 0, tc::bound_cA(3), tc::bound_cA(1), tc::BOUND_CB, zero_header));
        g_launches.fetch_add(1);
        if (header_zeroed) *header_zeroed = zero_header != nullptr;
        return VQB200_OK;
    }
    dim3 grid((n_embed + 31) / 32, (dim + 31) / 32), block(32, 8);
    k_codebook_transpose<<<grid, block, 0, st>>>(d_embed, cb.cbT, dim, n_embed);
    VQ_LAUNCH_CHECK();
    int warps_per_block = 8;
    k_codebook_norms<<<(n_embed + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>(
        cb.cbT, cb.ee, dim, n_embed);
    VQ_LAUNCH_CHECK();
    if (tcw_shape_ok(dim, n_embed)) {             // bf16 operand image(s) of the wide tensor-core engine (D = 128 / 256)
        VQ_CUDA(tcw_prepare(cb, dim, n_embed, st));
        g_launches.fetch_add(1);
    }
    return VQB200_OK;
}

// The forward, split so the host-buffer path can stream row chunks through it:
//   zero_first : clear statistics / diff accumulator before accumulating
//   finalize   : write diff = acc / (total_rows * dim)
int forward_impl(const float* d_x, const RowLayout& L, int dim, int n_embed, const void* d_codebook,
                 float* d_quantize, int64_t* d_ind, float* d_diff, float* d_stats, void* d_scratch,
                 int engine, bool zero_first, bool finalize, int64_t total_rows, cudaStream_t st,
                 float* dbg_scores = nullptr, int64_t scratch_rows = -1, unsigned long long* prof = nullptr,
                 float* d_x_dense = nullptr, const float* d_embed_prepare = nullptr, bool defer_fold = false) {
    if (scratch_rows < 0) scratch_rows = total_rows;
    CodebookImage cb = codebook_view(const_cast<void*>(d_codebook), dim, n_embed);
    ForwardScratch sc = scratch_view(d_scratch, scratch_rows, dim, n_embed);
    float* sums = d_stats;
    float* counts = d_stats ? d_stats + (size_t)n_embed * dim : nullptr;
    const double inv = total_rows > 0 ? 1.0 / ((double)total_rows * (double)dim)
                                      : std::numeric_limits<double>::quiet_NaN();   // mean over zero elements is NaN in the reference
    float* fin_diff = (finalize && d_diff) ? d_diff : nullptr;
    bool finalized = false;
    bool use_tc = false;
    const bool tc_engine = engine == VQB200_ENGINE_TCGEN05 || engine == VQB200_ENGINE_TCGEN05_BF16 || engine == VQB200_ENGINE_TCGEN05_TF32;
    if (L.n_rows > 0 && (tc_engine || engine == VQB200_ENGINE_AUTO))
        use_tc = tc_supported(L, d_x, dim, n_embed) || tcw_supported(L, d_x, dim, n_embed);
    const bool wide = use_tc && dim != tc::TC_D;  // tc_wide_kernel.cuh
    if (L.n_rows > 0 && tc_engine && !use_tc)
        return VQB200_EUNSUPPORTED;
    // statistics: private-table segmented reduction when [K][D] fp32 fits in shared memory (its fold kernel writes
    // d_stats, no memset needed), else the gather kernels fall back to global atomics on a cleared d_stats
    const size_t cs_smem = code_stats_smem_bytes(dim, n_embed);
    const bool stats_kernel = d_stats && cs_smem <= 200 * 1024 && n_embed <= 65535;
    // NCHW-physical rows consumed in place by the tensor-core kernel: the statistics kernel gathers rows by code, which
    // only coalesces on dense rows, so training needs the dense copy the kernel's converters can write on the side
    const bool nchw = use_tc && !tc_layout_dense(L, d_x, dim);
    if (nchw && d_x_dense && (reinterpret_cast<uintptr_t>(d_x_dense) & 31u)) return VQB200_EINVAL;   // written with 256-bit stores
    if (nchw && stats_kernel && !d_x_dense) {
        if (tc_engine) return VQB200_EUNSUPPORTED;
        use_tc = false;
    }
    bool header_zeroed = false;
    if (zero_first && d_stats && (!stats_kernel || L.n_rows == 0)) VQ_CUDA(cudaMemsetAsync(d_stats, 0, vqb200_stats_bytes(dim, n_embed), st));
    if (d_embed_prepare) {                        // ahead of the main kernel, adjacent launches (PDL); also clears the scratch header
        int rc = prepare_codebook(d_embed_prepare, dim, n_embed, const_cast<void*>(d_codebook), st,
                                  zero_first ? reinterpret_cast<unsigned int*>(sc.diff_acc) : nullptr, &header_zeroed);
        if (rc) return rc;
    }
    if (zero_first) {
        // loss accumulator, flagged-row counter, tickets + the rows-per-code counters behind them
        if (!header_zeroed) VQ_CUDA(cudaMemsetAsync(sc.diff_acc, 0, scratch_header_bytes(n_embed), st));
    } else if (use_tc) {
        VQ_CUDA(cudaMemsetAsync(sc.flagged_count, 0, sizeof(int), st));
    }
    if (L.n_rows > 0) {
        const int nsplit = engine == VQB200_ENGINE_TCGEN05_BF16 ? 1 : (engine == VQB200_ENGINE_TCGEN05 ? 3 : (engine == VQB200_ENGINE_TCGEN05_TF32 ? 0 : -1));
        if (stats_kernel) { sums = nullptr; counts = nullptr; }
        const bool want_gather = d_quantize || d_diff || sums;
        const size_t gsmem = (size_t)GS_BM * (dim + 1) * sizeof(float);
        if (gsmem > 200 * 1024) return VQB200_EUNSUPPORTED;
        // opt-in shared-memory sizes are per-device function attributes: set once per (device, size), not on every call
        if (gsmem > 48 * 1024 && smem_attr_needed(0, gsmem))
            VQ_CUDA(cudaFuncSetAttribute(k_gather_stats, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
        // k_fixup also carries ~18 KB of static shared memory (the exact re-score tiles): opt in as soon as the sum passes 48 KB
        if (gsmem > 24 * 1024 && smem_attr_needed(1, gsmem))
            VQ_CUDA(cudaFuncSetAttribute(k_fixup, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
        const int sms = tc_num_sms();
        if (use_tc) {
            // tensor-core filter + fused output for certified rows; flagged rows -> exact SIMT fix-up (one launch:
            // re-score, gather / output / loss of those rows, loss finalisation)
            unsigned long long wide_launches = 0;
            int rc = wide ? tcw_forward(d_x, L, dim, n_embed, cb, d_quantize, d_ind, sc, d_diff ? sc.diff_acc : nullptr, sums,
                                        counts, dbg_scores, st, &wide_launches)
                          : tc_forward(d_x, L, dim, n_embed, cb, d_quantize, d_ind, sc, d_diff ? sc.diff_acc : nullptr, sums,
                                       counts, dbg_scores, st, prof, nsplit, nullptr, (nchw && stats_kernel) ? d_x_dense : nullptr);
            g_launches.fetch_add(wide ? wide_launches : 1ull);
            if (rc) return cuda_fail(cudaGetLastError());
            VQ_CUDA(launch_pdl(k_fixup, dim3(2 * sms), dim3(AS_THREADS), gsmem, st, d_x, L, dim, n_embed, cb.cbT, cb.ee, d_ind,
                               d_quantize, d_diff ? sc.diff_acc : nullptr, sums, counts, sc.flagged_rows, sc.flagged_count,
                               want_gather ? 1 : 0, fin_diff, inv, sc.ticket, sc.fix_partial, sc.fix_tickets));
            g_launches.fetch_add(1);
            finalized = true;
        } else {
            int64_t blocks = std::min<int64_t>((L.n_rows + AS_BM - 1) / AS_BM, (int64_t)sms * 64);
            k_assign_exact<<<(unsigned)blocks, AS_THREADS, 0, st>>>(d_x, L, dim, n_embed, cb.cbT, cb.ee, d_ind,
                                                                      nullptr, nullptr);
            VQ_LAUNCH_CHECK();
            if (want_gather) {
                int64_t gblocks = std::min<int64_t>((L.n_rows + GS_BM - 1) / GS_BM, (int64_t)sms * 64);
                k_gather_stats<<<(unsigned)gblocks, GS_THREADS, gsmem, st>>>(
                    d_x, L, dim, n_embed, cb.cbT, d_ind, d_quantize, d_diff ? sc.diff_acc : nullptr, sums, counts,
                    nullptr, nullptr, fin_diff, inv, sc.ticket);
                VQ_LAUNCH_CHECK();
                finalized = true;
            }
        }
        if (stats_kernel) {
            if (smem_attr_needed(2, cs_smem))
                VQ_CUDA(cudaFuncSetAttribute(k_code_stats, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cs_smem));
            // rows per trip chosen so the trips divide evenly over the SMs (one CTA per SM, private table each)
            const int sms_cs = std::min(tc_num_sms(), STAT_PARTS);
            int64_t waves = (L.n_rows + (int64_t)sms_cs * CS_CHUNK - 1) / ((int64_t)sms_cs * CS_CHUNK);
            int64_t chunk = (L.n_rows + sms_cs * waves - 1) / (sms_cs * waves);
            chunk = std::min<int64_t>(CS_CHUNK, std::max<int64_t>(256, (chunk + 31) / 32 * 32));
            int64_t n_chunks = (L.n_rows + chunk - 1) / chunk;
            int parts = (int)std::min<int64_t>(n_chunks, sms_cs);
            const bool from_dense = use_tc && nchw;
            const RowLayout Ld{L.n_rows, L.n_rows, 0, dim, 1};
            VQ_CUDA(launch_pdl(k_code_stats, dim3(parts), dim3(CS_THREADS), cs_smem, st, from_dense ? d_x_dense : d_x,
                               from_dense ? Ld : L, dim, n_embed, d_ind, sc.stat_partials, (int)chunk, sc.code_counts, sc.n_parts));
            g_launches.fetch_add(1);
            if (defer_fold) return VQB200_OK;     // the caller's fold + EMA kernel consumes the per-CTA tables (finalize ran in k_fixup / k_gather_stats)
            // d_stats (+)= sum of the per-CTA tables: overwrite on the first call, accumulate on host-path continuation chunks
            const int nstat = n_embed * (dim + 1);
            VQ_CUDA(launch_pdl(k_stats_fold, dim3((nstat + 127) / 128), dim3(32, FOLD_Y), 0, st, sc.stat_partials, parts, nstat,
                               d_stats, zero_first ? 0 : 1));
            g_launches.fetch_add(1);
        }
    }
    if (fin_diff && !finalized) {
        k_finalize_diff<<<1, 32, 0, st>>>(sc.diff_acc, d_diff, inv);
        VQ_LAUNCH_CHECK();
    }
    return VQB200_OK;
}

int ema_impl(const float* d_stats, const PeerStats* peers, float* d_cluster_size, float* d_embed_avg, float* d_embed,
             int dim, int n_embed, float decay, float one_minus_decay, float eps, void* d_codebook, cudaStream_t st) {
    // d_stats = [sums K*D | counts K | 4 spare words: n = sum(cluster_size), EMA ticket] -- see vqb200_stats_bytes
    const float* sums = d_stats;
    const float* counts = d_stats + (size_t)n_embed * dim;
    float* spare = const_cast<float*>(d_stats) + (size_t)n_embed * (dim + 1);
    CodebookImage cb{nullptr, nullptr, nullptr, nullptr};
    if (d_codebook) cb = codebook_view(d_codebook, dim, n_embed);
    if (tc_shape_ok(dim, n_embed)) {              // one launch: [all-reduce over peer memory +] EMA + renormalise + next image
        PeerStats none{};
        unsigned int* ticket = reinterpret_cast<unsigned int*>(spare + 1);
        if (peers)
            VQ_CUDA(launch_pdl(k_ema64<true>, dim3(n_embed / PREP_CODES), dim3(256), 0, st, d_stats, *peers, d_cluster_size,
                               d_embed_avg, d_embed, cb.cbT, cb.ee, cb.tc, n_embed, decay, one_minus_decay, eps,
                               tc::bound_cA(3), tc::bound_cA(1), tc::BOUND_CB, ticket));
        else
            VQ_CUDA(launch_pdl(k_ema64<false>, dim3(n_embed / PREP_CODES), dim3(256), 0, st, d_stats, none, d_cluster_size,
                               d_embed_avg, d_embed, cb.cbT, cb.ee, cb.tc, n_embed, decay, one_minus_decay, eps,
                               tc::bound_cA(3), tc::bound_cA(1), tc::BOUND_CB, ticket));
        g_launches.fetch_add(1);
        return VQB200_OK;
    }
    if (peers) return VQB200_EUNSUPPORTED;
    k_ema_cluster<<<1, 1024, 0, st>>>(counts, d_cluster_size, n_embed, decay, one_minus_decay, spare);
    VQ_LAUNCH_CHECK();
    int wpb = 8;
    k_ema_embed<<<(n_embed + wpb - 1) / wpb, wpb * 32, 0, st>>>(sums, d_cluster_size, spare, d_embed_avg,
                                                                 d_embed, cb.cbT, cb.ee, dim, n_embed, decay,
                                                                 one_minus_decay, eps);
    VQ_LAUNCH_CHECK();
    if (d_codebook && tcw_shape_ok(dim, n_embed)) {
        VQ_CUDA(tcw_prepare(cb, dim, n_embed, st));
        g_launches.fetch_add(1);
    }
    return VQB200_OK;
}

// fold of the per-CTA statistics tables + [exchange over peer memory] + EMA in one launch (k_ema64f); dim 64, n_embed 256 / 512
int fold_ema_launch(const ForwardScratch& sc, const PeerFold* peers, float* d_cluster_size, float* d_embed_avg, float* d_embed,
                    int n_embed, float decay, float one_minus_decay, float eps, cudaStream_t st) {
    PeerFold none{};
    const dim3 grid(n_embed / EF_CODES), block(EF_THREADS);
    if (peers)
        VQ_CUDA(launch_pdl(k_ema64f<true>, grid, block, 0, st, sc.stat_partials, sc.n_parts, sc.code_counts, *peers, d_cluster_size,
                           d_embed_avg, d_embed, (float*)nullptr, (float*)nullptr, (unsigned char*)nullptr, n_embed, decay,
                           one_minus_decay, eps, tc::bound_cA(3), tc::bound_cA(1), tc::BOUND_CB, sc.ema_ticket));
    else
        VQ_CUDA(launch_pdl(k_ema64f<false>, grid, block, 0, st, sc.stat_partials, sc.n_parts, sc.code_counts, none, d_cluster_size,
                           d_embed_avg, d_embed, (float*)nullptr, (float*)nullptr, (unsigned char*)nullptr, n_embed, decay,
                           one_minus_decay, eps, tc::bound_cA(3), tc::bound_cA(1), tc::BOUND_CB, sc.ema_ticket));
    g_launches.fetch_add(1);
    return VQB200_OK;
}

}  // namespace

extern "C" {

int vqb200_abi_version(void) { return VQB200_ABI_VERSION; }

const char* vqb200_error_string(int code) {
    switch (code) {
        case VQB200_OK: return "ok";
        case VQB200_EINVAL: return "invalid argument";
        case VQB200_EUNSUPPORTED: return "unsupported shape or layout";
        case VQB200_ECUDA: return "CUDA error (see vqb200_last_cuda_error)";
        case VQB200_ENODEVICE: return "no sm_100 CUDA device";
        default: return "unknown error";
    }
}

int vqb200_last_cuda_error(void) { return g_last_cuda_error; }
uint64_t vqb200_launch_count(void) { return (uint64_t)g_launches.load(); }

size_t vqb200_codebook_bytes(int32_t dim, int32_t n_embed) {
    if (dim <= 0 || n_embed <= 0) return 0;
    return codebook_bytes(dim, n_embed);
}
size_t vqb200_forward_scratch_bytes(int64_t n_rows, int32_t dim, int32_t n_embed) {
    if (n_rows < 0) return 0;
    if (dim <= 0 || n_embed <= 0) return 0;
    return forward_scratch_bytes(n_rows, dim, n_embed);
}
size_t vqb200_stats_bytes(int32_t dim, int32_t n_embed) {
    if (dim <= 0 || n_embed <= 0) return 0;
    return ((size_t)n_embed * (dim + 1) + 4) * sizeof(float);   // + spare scalars (n = sum cluster_size)
}

int vqb200_codebook_prepare(const float* d_embed, int32_t dim, int32_t n_embed, void* d_codebook, void* stream) {
    if (!d_embed || !d_codebook || dim <= 0 || n_embed <= 0) return VQB200_EINVAL;
    return prepare_codebook(d_embed, dim, n_embed, d_codebook, (cudaStream_t)stream);
}

int vqb200_quantize_forward(const float* d_x, int64_t n_rows, int32_t dim, int32_t n_embed,
                            int64_t rows_per_image, int64_t image_stride, int64_t row_stride,
                            int64_t col_stride, const void* d_codebook, float* d_quantize,
                            int64_t* d_embed_ind, float* d_diff, float* d_stats, void* d_scratch,
                            int32_t engine, void* stream) {
    if (!d_codebook || !d_scratch || dim <= 0 || n_embed <= 0 || n_rows < 0) return VQB200_EINVAL;
    if (n_rows > 0 && (!d_x || !d_embed_ind)) return VQB200_EINVAL;
    if (n_rows > (int64_t)INT32_MAX) return VQB200_EUNSUPPORTED;
    if (engine < VQB200_ENGINE_AUTO || engine > VQB200_ENGINE_TCGEN05_TF32) return VQB200_EINVAL;
    if (n_rows > 0 && !layout_ok(n_rows, dim, rows_per_image, image_stride, row_stride, col_stride))
        return VQB200_EUNSUPPORTED;
    RowLayout L{n_rows, rows_per_image > 0 ? rows_per_image : 1, image_stride, row_stride, col_stride};
    return forward_impl(d_x, L, dim, n_embed, d_codebook, d_quantize, d_embed_ind, d_diff, d_stats, d_scratch,
                        engine, true, true, n_rows, (cudaStream_t)stream);
}

int vqb200_ema_update(const float* d_stats, float* d_cluster_size, float* d_embed_avg, float* d_embed,
                      int32_t dim, int32_t n_embed, float decay, float one_minus_decay, float eps,
                      void* d_codebook, void* stream) {
    if (!d_stats || !d_cluster_size || !d_embed_avg || !d_embed || dim <= 0 || n_embed <= 0)
        return VQB200_EINVAL;
    return ema_impl(d_stats, nullptr, d_cluster_size, d_embed_avg, d_embed, dim, n_embed, decay, one_minus_decay, eps,
                    d_codebook, (cudaStream_t)stream);
}

int vqb200_ema_update_p2p(const void* const* h_stats_ptrs, void* const* h_flag_ptrs, int32_t rank, int32_t world,
                          uint32_t step, float* d_cluster_size, float* d_embed_avg, float* d_embed, int32_t dim,
                          int32_t n_embed, float decay, float one_minus_decay, float eps, void* d_codebook, void* stream) {
    if (!h_stats_ptrs || !h_flag_ptrs || !d_cluster_size || !d_embed_avg || !d_embed || dim <= 0 || n_embed <= 0)
        return VQB200_EINVAL;
    if (world < 1 || world > P2P_MAX_RANKS || rank < 0 || rank >= world || step == 0) return VQB200_EINVAL;
    if (!tc_shape_ok(dim, n_embed)) return VQB200_EUNSUPPORTED;
    PeerStats ps{};
    for (int r = 0; r < world; ++r) {
        if (!h_stats_ptrs[r] || !h_flag_ptrs[r]) return VQB200_EINVAL;
        ps.stats[r] = static_cast<const float*>(h_stats_ptrs[r]);
        ps.flags[r] = static_cast<unsigned int*>(h_flag_ptrs[r]);
    }
    ps.rank = rank; ps.world = world; ps.step = step;
    return ema_impl(ps.stats[rank], &ps, d_cluster_size, d_embed_avg, d_embed, dim, n_embed, decay, one_minus_decay, eps,
                    d_codebook, (cudaStream_t)stream);
}

static int fill_peer_fold(PeerFold& pf, void* const* h_push_dst, const void* const* h_recv, void* d_err, void* d_step_counter,
                          int rank, int world) {
    if (!h_push_dst || !h_recv || !d_err || !d_step_counter) return VQB200_EINVAL;
    if (world < 1 || world > P2P_MAX_RANKS || rank < 0 || rank >= world) return VQB200_EINVAL;
    for (int par = 0; par < 2; ++par)
        for (int r = 0; r < world; ++r) {           // h_push_dst / h_recv: [2 parities][world] pointers, parity-major
            void* dst = h_push_dst[par * world + r];
            const void* rcv = h_recv[par * world + r];
            if (!dst || !rcv) return VQB200_EINVAL;
            if ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(rcv)) & 7u) return VQB200_EINVAL;
            pf.push_dst[par][r] = static_cast<uint2*>(dst);
            pf.recv[par][r] = static_cast<const uint2*>(rcv);
        }
    pf.err = static_cast<unsigned int*>(d_err);
    pf.step_counter = static_cast<unsigned int*>(d_step_counter);
    pf.rank = rank; pf.world = world;
    return VQB200_OK;
}

int vqb200_stats_exchange_peers(float* d_stats, int64_t n_words, void* const* h_push_dst, const void* const* h_recv,
                                void* d_err, void* d_step_counter, int32_t rank, int32_t world, void* stream) {
    if (!d_stats || n_words <= 0 || n_words > (int64_t)(1 << 24)) return VQB200_EINVAL;
    PeerFold pf{};
    int rc = fill_peer_fold(pf, h_push_dst, h_recv, d_err, d_step_counter, rank, world);
    if (rc) return rc;
    const int blocks = (int)std::min<int64_t>((n_words + 1023) / 1024, tc_num_sms());     // ONE wave (see the kernel)
    VQ_CUDA(launch_pdl(k_exchange_ll, dim3(blocks), dim3(1024), 0, (cudaStream_t)stream, d_stats, (int)n_words, pf));
    g_launches.fetch_add(1);
    return VQB200_OK;
}

int vqb200_quantize_step(const float* d_x, int64_t n_rows, int32_t dim, int32_t n_embed, int64_t rows_per_image,
                         int64_t image_stride, int64_t row_stride, int64_t col_stride, float* d_embed,
                         float* d_cluster_size, float* d_embed_avg, void* d_codebook, float* d_quantize,
                         int64_t* d_embed_ind, float* d_diff, float* d_stats, void* d_scratch, float* d_x_dense,
                         int32_t engine, int32_t ema, float decay, float one_minus_decay, float eps, void* stream) {
    if (!d_embed) return VQB200_EINVAL;
    if (ema && d_stats && (!d_cluster_size || !d_embed_avg)) return VQB200_EINVAL;
    int rc;
    if (!d_codebook || !d_scratch || dim <= 0 || n_embed <= 0 || n_rows < 0) return VQB200_EINVAL;
    if (n_rows > 0 && (!d_x || !d_embed_ind)) return VQB200_EINVAL;
    if (n_rows > (int64_t)INT32_MAX) return VQB200_EUNSUPPORTED;
    if (engine < VQB200_ENGINE_AUTO || engine > VQB200_ENGINE_TCGEN05_TF32) return VQB200_EINVAL;
    if (n_rows > 0 && !layout_ok(n_rows, dim, rows_per_image, image_stride, row_stride, col_stride))
        return VQB200_EUNSUPPORTED;
    {
        RowLayout L{n_rows, rows_per_image > 0 ? rows_per_image : 1, image_stride, row_stride, col_stride};
        // dim 64, n_embed 256 / 512 with the EMA requested: the per-CTA statistics tables are folded INSIDE the EMA kernel
        // (one launch instead of k_stats_fold + k_ema64; d_stats is not written)
        if (ema && d_stats && tc_shape_ok(dim, n_embed)) {
            ForwardScratch sc = scratch_view(d_scratch, n_rows, dim, n_embed);
            rc = forward_impl(d_x, L, dim, n_embed, d_codebook, d_quantize, d_embed_ind, d_diff, d_stats, d_scratch, engine, true,
                              true, n_rows, (cudaStream_t)stream, nullptr, -1, nullptr, d_x_dense, d_embed, true);
            if (rc) return rc;
            return fold_ema_launch(sc, nullptr, d_cluster_size, d_embed_avg, d_embed, n_embed, decay, one_minus_decay, eps,
                                   (cudaStream_t)stream);
        }
        rc = forward_impl(d_x, L, dim, n_embed, d_codebook, d_quantize, d_embed_ind, d_diff, d_stats, d_scratch, engine, true,
                          true, n_rows, (cudaStream_t)stream, nullptr, -1, nullptr, d_x_dense, d_embed);
    }
    if (rc || !ema || !d_stats) return rc;
    return vqb200_ema_update(d_stats, d_cluster_size, d_embed_avg, d_embed, dim, n_embed, decay, one_minus_decay, eps,
                             nullptr, stream);
}

int vqb200_quantize_step_peers(const float* d_x, int64_t n_rows, int32_t dim, int32_t n_embed, int64_t rows_per_image,
                               int64_t image_stride, int64_t row_stride, int64_t col_stride, float* d_embed,
                               float* d_cluster_size, float* d_embed_avg, void* d_codebook, float* d_quantize,
                               int64_t* d_embed_ind, float* d_diff, void* d_scratch, float* d_x_dense, int32_t engine,
                               float decay, float one_minus_decay, float eps, void* const* h_push_dst,
                               const void* const* h_recv, void* d_err, void* d_step_counter, int32_t rank, int32_t world,
                               void* stream) {
    if (!d_embed || !d_cluster_size || !d_embed_avg || !d_codebook || !d_scratch || dim <= 0 || n_embed <= 0) return VQB200_EINVAL;
    if (n_rows < 0 || (n_rows > 0 && (!d_x || !d_embed_ind))) return VQB200_EINVAL;
    if (n_rows > (int64_t)INT32_MAX) return VQB200_EUNSUPPORTED;
    if (engine < VQB200_ENGINE_AUTO || engine > VQB200_ENGINE_TCGEN05_TF32) return VQB200_EINVAL;
    if (n_rows > 0 && !layout_ok(n_rows, dim, rows_per_image, image_stride, row_stride, col_stride)) return VQB200_EUNSUPPORTED;
    if (!tc_shape_ok(dim, n_embed)) return VQB200_EUNSUPPORTED;      // the fold + EMA kernel: dim 64, n_embed 256 / 512
    PeerFold pf{};
    {
        int prc = fill_peer_fold(pf, h_push_dst, h_recv, d_err, d_step_counter, rank, world);
        if (prc) return prc;
    }
    RowLayout L{n_rows, rows_per_image > 0 ? rows_per_image : 1, image_stride, row_stride, col_stride};
    ForwardScratch sc = scratch_view(d_scratch, n_rows, dim, n_embed);
    int rc = forward_impl(d_x, L, dim, n_embed, d_codebook, d_quantize, d_embed_ind, d_diff, sc.stat_partials, d_scratch, engine,
                          true, true, n_rows, (cudaStream_t)stream, nullptr, -1, nullptr, d_x_dense, d_embed, true);
    if (rc) return rc;
    return fold_ema_launch(sc, &pf, d_cluster_size, d_embed_avg, d_embed, n_embed, decay, one_minus_decay, eps, (cudaStream_t)stream);
}

int vqb200_repack_rows(const float* d_src, float* d_dst, int64_t n_rows, int32_t dim, int64_t rows_per_image,
                       int64_t image_stride, int64_t row_stride, int64_t col_stride, int32_t to_dense, void* stream) {
    if (dim <= 0 || n_rows < 0) return VQB200_EINVAL;
    if (n_rows == 0) return VQB200_OK;
    if (!d_src || !d_dst) return VQB200_EINVAL;
    if (!layout_ok(n_rows, dim, rows_per_image, image_stride, row_stride, col_stride)) return VQB200_EUNSUPPORTED;
    RowLayout L{n_rows, rows_per_image, image_stride, row_stride, col_stride};
    const size_t smem = (size_t)RP_ROWS * (dim + 1) * sizeof(float);
    if (smem > 200 * 1024) return VQB200_EUNSUPPORTED;
    const int64_t tiles = (n_rows + RP_ROWS - 1) / RP_ROWS;
    const unsigned blocks = (unsigned)std::min<int64_t>(tiles, (int64_t)tc_num_sms() * 32);
    cudaStream_t st = (cudaStream_t)stream;
    if (to_dense) {
        if (smem > 48 * 1024) VQ_CUDA(cudaFuncSetAttribute(k_repack_rows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_repack_rows<true><<<blocks, RP_THREADS, smem, st>>>(d_src, d_dst, L, dim);
    } else {
        if (smem > 48 * 1024) VQ_CUDA(cudaFuncSetAttribute(k_repack_rows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_repack_rows<false><<<blocks, RP_THREADS, smem, st>>>(d_src, d_dst, L, dim);
    }
    VQ_LAUNCH_CHECK();
    return VQB200_OK;
}

int vqb200_quantize_backward(const float* d_x, const int64_t* d_embed_ind, const void* d_codebook,
                             const float* d_grad_quantize, const float* d_grad_diff, float* d_grad_x,
                             int64_t n_rows, int32_t dim, int32_t n_embed, int64_t rows_per_image,
                             int64_t image_stride, int64_t row_stride, int64_t col_stride, void* stream) {
    if (!d_codebook || dim <= 0 || n_embed <= 0 || n_rows < 0) return VQB200_EINVAL;
    if (n_rows == 0) return VQB200_OK;
    if (!d_x || !d_embed_ind || !d_grad_x) return VQB200_EINVAL;
    if (!layout_ok(n_rows, dim, rows_per_image, image_stride, row_stride, col_stride)) return VQB200_EUNSUPPORTED;
    RowLayout L{n_rows, rows_per_image, image_stride, row_stride, col_stride};
    CodebookImage cb = codebook_view(const_cast<void*>(d_codebook), dim, n_embed);
    int64_t total = n_rows * (int64_t)dim;
    const bool dense = col_stride == 1 && row_stride == dim && (n_rows <= rows_per_image || image_stride == rows_per_image * (int64_t)dim) &&
                       dim % 4 == 0 && ((reinterpret_cast<uintptr_t>(d_x) | reinterpret_cast<uintptr_t>(d_grad_x) |
                                         reinterpret_cast<uintptr_t>(d_grad_quantize)) & 15u) == 0;
    if (dense) {
        const int64_t n_vec = total / 4;
        int blocks = (int)std::min<int64_t>((n_vec + 255) / 256, (int64_t)tc_num_sms() * 32);
        k_backward_dense<<<blocks, 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4*>(d_x), n_vec, dim / 4, d_embed_ind, reinterpret_cast<const float4*>(cb.cbT),
            reinterpret_cast<const float4*>(d_grad_quantize), d_grad_diff, reinterpret_cast<float4*>(d_grad_x),
            2.0 / (double)total);
        VQ_LAUNCH_CHECK();
        return VQB200_OK;
    }
    const bool nchw = row_stride == 1 && col_stride >= rows_per_image && rows_per_image % 4 == 0 && col_stride % 4 == 0 &&
                      image_stride % 4 == 0 && n_rows % 4 == 0 &&
                      ((reinterpret_cast<uintptr_t>(d_x) | reinterpret_cast<uintptr_t>(d_grad_x) |
                        reinterpret_cast<uintptr_t>(d_grad_quantize)) & 15u) == 0;
    if (nchw) {
        const int dims_per_block = dim >= 64 ? 16 : dim;
        dim3 grid((unsigned)((n_rows / 4 + 255) / 256), (unsigned)((dim + dims_per_block - 1) / dims_per_block));
        k_backward_nchw<<<grid, 256, 0, (cudaStream_t)stream>>>(d_x, L, dim, dims_per_block, d_embed_ind, cb.cbT, d_grad_quantize,
                                                                 d_grad_diff, d_grad_x, 2.0 / (double)total);
        VQ_LAUNCH_CHECK();
        return VQB200_OK;
    }
    int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
    k_backward<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_x, L, dim, d_embed_ind, cb.cbT, d_grad_quantize,
                                                          d_grad_diff, d_grad_x, 2.0 / (double)total);
    VQ_LAUNCH_CHECK();
    return VQB200_OK;
}

int vqb200_embed_code(const int64_t* d_embed_id, int64_t n_rows, const void* d_codebook, int32_t dim,
                      int32_t n_embed, float* d_out, int32_t* d_status, void* stream) {
    if (!d_codebook || dim <= 0 || n_embed <= 0 || n_rows < 0) return VQB200_EINVAL;
    if (n_rows == 0) return VQB200_OK;
    if (!d_embed_id || !d_out) return VQB200_EINVAL;
    CodebookImage cb = codebook_view(const_cast<void*>(d_codebook), dim, n_embed);
    if (d_status) VQ_CUDA(cudaMemsetAsync(d_status, 0, sizeof(int32_t), (cudaStream_t)stream));
    int64_t blocks = (n_rows + 7) / 8;
    k_embed_code<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_embed_id, n_rows, cb.cbT, dim, n_embed, d_out,
                                                                      d_status);
    VQ_LAUNCH_CHECK();
    return VQB200_OK;
}

int vqb200_pack_indices(const int64_t* d_embed_ind, int64_t n, int32_t n_embed, int32_t out_bytes, void* d_out,
                        int32_t* d_status, void* stream) {
    if (n < 0 || n_embed <= 0 || (out_bytes != 2 && out_bytes != 4)) return VQB200_EINVAL;
    if (out_bytes == 2 && n_embed > 65536) return VQB200_EUNSUPPORTED;
    if (n == 0) return VQB200_OK;
    if (!d_embed_ind || !d_out) return VQB200_EINVAL;
    if ((reinterpret_cast<uintptr_t>(d_embed_ind) & 15u) || (reinterpret_cast<uintptr_t>(d_out) & 15u)) return VQB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (d_status) VQ_CUDA(cudaMemsetAsync(d_status, 0, sizeof(int32_t), st));
    const unsigned blocks = (unsigned)((n + 1023) / 1024);
    if (out_bytes == 2) k_pack_indices<unsigned short><<<blocks, 256, 0, st>>>(d_embed_ind, n, n_embed, (unsigned short*)d_out, d_status);
    else k_pack_indices<int><<<blocks, 256, 0, st>>>(d_embed_ind, n, n_embed, (int*)d_out, d_status);
    VQ_LAUNCH_CHECK();
    return VQB200_OK;
}

int vqb200_unpack_indices(const void* d_codes, int64_t n, int32_t in_bytes, int64_t* d_embed_id, void* stream) {
    if (n < 0 || (in_bytes != 2 && in_bytes != 4)) return VQB200_EINVAL;
    if (n == 0) return VQB200_OK;
    if (!d_codes || !d_embed_id) return VQB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (in_bytes == 2) k_unpack_indices<unsigned short><<<blocks, 256, 0, st>>>((const unsigned short*)d_codes, n, d_embed_id);
    else k_unpack_indices<int><<<blocks, 256, 0, st>>>((const int*)d_codes, n, d_embed_id);
    VQ_LAUNCH_CHECK();
    return VQB200_OK;
}

int vqb200_debug_tc_scores(const float* d_x, int64_t n_rows, int32_t dim, int32_t n_embed, const void* d_codebook,
                           int64_t* d_embed_ind, float* d_scores, int32_t* d_flagged_count, void* d_scratch,
                           void* stream) {
    return vqb200_debug_tc_scores_ex(d_x, n_rows, dim, n_embed, d_codebook, d_embed_ind, d_scores, d_flagged_count, d_scratch,
                                     VQB200_ENGINE_TCGEN05, stream);
}

int vqb200_debug_tc_scores_ex(const float* d_x, int64_t n_rows, int32_t dim, int32_t n_embed, const void* d_codebook,
                              int64_t* d_embed_ind, float* d_scores, int32_t* d_flagged_count, void* d_scratch,
                              int32_t engine, void* stream) {
    if (!d_x || !d_codebook || !d_embed_ind || !d_scores || !d_scratch || n_rows <= 0) return VQB200_EINVAL;
    if (engine != VQB200_ENGINE_TCGEN05 && engine != VQB200_ENGINE_TCGEN05_BF16 && engine != VQB200_ENGINE_TCGEN05_TF32) return VQB200_EINVAL;
    RowLayout L{n_rows, n_rows, 0, dim, 1};
    if (!tc_supported(L, d_x, dim, n_embed) && !tcw_supported(L, d_x, dim, n_embed)) return VQB200_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = forward_impl(d_x, L, dim, n_embed, d_codebook, nullptr, d_embed_ind, nullptr, nullptr, d_scratch,
                          engine, true, false, n_rows, st, d_scores);
    if (rc) return rc;
    if (d_flagged_count)
        VQ_CUDA(cudaMemcpyAsync(d_flagged_count, scratch_view(d_scratch, n_rows, dim, n_embed).flagged_count, sizeof(int),
                                cudaMemcpyDeviceToDevice, st));
    return VQB200_OK;
}

int vqb200_tc_split(void) { return tc_nsplit(); }

int vqb200_tc_supported(const float* d_x, int64_t n_rows, int32_t dim, int32_t n_embed, int64_t rows_per_image,
                        int64_t image_stride, int64_t row_stride, int64_t col_stride) {
    if (n_rows <= 0 || rows_per_image <= 0) return 0;
    RowLayout L{n_rows, rows_per_image, image_stride, row_stride, col_stride};
    return (tc_supported(L, d_x, dim, n_embed) || tcw_supported(L, d_x, dim, n_embed)) ? 1 : 0;
}

int vqb200_debug_tc_profile(const float* d_x, int64_t n_rows, int32_t dim, int32_t n_embed, const void* d_codebook,
                            float* d_quantize, int64_t* d_embed_ind, void* d_scratch, uint64_t* d_prof, int32_t engine,
                            void* stream) {
    if (!d_x || !d_codebook || !d_embed_ind || !d_scratch || !d_prof || n_rows <= 0) return VQB200_EINVAL;
    if (engine != VQB200_ENGINE_TCGEN05 && engine != VQB200_ENGINE_TCGEN05_BF16 && engine != VQB200_ENGINE_TCGEN05_TF32) return VQB200_EINVAL;
    RowLayout L{n_rows, n_rows, 0, dim, 1};
    if (!tc_supported(L, d_x, dim, n_embed)) return VQB200_EUNSUPPORTED;
    return forward_impl(d_x, L, dim, n_embed, d_codebook, d_quantize, d_embed_ind, nullptr, nullptr, d_scratch,
                        engine, true, false, n_rows, (cudaStream_t)stream, nullptr, -1,
                        reinterpret_cast<unsigned long long*>(d_prof));
}
int vqb200_debug_tc_kernel(const float* d_x, int64_t n_rows, int32_t dim, int32_t n_embed, const void* d_codebook,
                           float* d_quantize, int64_t* d_embed_ind, void* d_scratch, int32_t engine, void* stream) {
    if (!d_x || !d_codebook || !d_embed_ind || !d_scratch || n_rows <= 0) return VQB200_EINVAL;
    if (engine != VQB200_ENGINE_TCGEN05 && engine != VQB200_ENGINE_TCGEN05_BF16 && engine != VQB200_ENGINE_TCGEN05_TF32) return VQB200_EINVAL;
    RowLayout L{n_rows, n_rows, 0, dim, 1};
    if (!tc_supported(L, d_x, dim, n_embed)) return VQB200_EUNSUPPORTED;
    CodebookImage cb = codebook_view(const_cast<void*>(d_codebook), dim, n_embed);
    ForwardScratch sc = scratch_view(d_scratch, n_rows, dim, n_embed);
    int rc = tc_forward(d_x, L, dim, n_embed, cb, d_quantize, d_embed_ind, sc, sc.diff_acc, nullptr, nullptr, nullptr,
                        (cudaStream_t)stream, nullptr, engine == VQB200_ENGINE_TCGEN05_BF16 ? 1 : (engine == VQB200_ENGINE_TCGEN05_TF32 ? 0 : 3));
    g_launches.fetch_add(1);
    return rc ? cuda_fail(cudaGetLastError()) : VQB200_OK;
}
int vqb200_tc_profile_slots(void) { return (int)tc::PROF_SLOTS; }

int vqb200_debug_pingpong(void* d_my_flag, void* d_peer_flag, int32_t iters, int32_t initiator, int32_t with_fence,
                          uint64_t* d_ns, void* stream) {
    if (!d_my_flag || !d_peer_flag || !d_ns || iters <= 0) return VQB200_EINVAL;
    k_pingpong<<<1, 32, 0, (cudaStream_t)stream>>>(static_cast<unsigned int*>(d_my_flag), static_cast<unsigned int*>(d_peer_flag), iters,
                                                    initiator, with_fence, reinterpret_cast<unsigned long long*>(d_ns));
    VQ_LAUNCH_CHECK();
    return VQB200_OK;
}

// ---- host-buffer path ----------------------------------------------------------------------------
struct vqb200_host_ctx {
    int64_t max_rows;
    int dim, n_embed;
    int64_t chunk_rows;
    float* d_x;
    float* d_q;
    int64_t* d_ind;
    float* d_diff;
    float* d_stats;
    void* d_scratch;
    void* d_codebook;
    cudaStream_t s_in, s_run, s_out;
    cudaEvent_t ev_in[64], ev_run[64], ev_start;
    int device;                  // the device the context was created on (made current by every call)
    cudaStream_t caller;         // stream whose earlier work (writes to d_embed / the EMA buffers) the call must observe
};

int vqb200_host_ctx_create(int64_t max_rows, int32_t dim, int32_t n_embed, vqb200_host_ctx** out) {
    if (!out || max_rows <= 0 || dim <= 0 || n_embed <= 0) return VQB200_EINVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return VQB200_ENODEVICE;
    vqb200_host_ctx* c = new (std::nothrow) vqb200_host_ctx();
    if (!c) return VQB200_EINVAL;
    std::memset(c, 0, sizeof(*c));
    c->max_rows = max_rows; c->dim = dim; c->n_embed = n_embed;
    if (cudaGetDevice(&c->device) != cudaSuccess) c->device = 0;
    c->caller = nullptr;         // legacy default stream (= torch's default stream) until vqb200_host_ctx_set_stream
    // ~8 MiB of fp32 rows per chunk keeps PCIe busy in both directions while the kernels run
    int64_t chunk_mb = 16;      // measured on B200 / PCIe Gen5: 16 MiB chunks move 272 MB per call in 3.43 ms, 8 MiB in 3.75 ms
    if (const char* e = getenv("VQB200_HOST_CHUNK_MB")) { int v = atoi(e); if (v >= 1 && v <= 1024) chunk_mb = v; }
    int64_t cr = std::max<int64_t>(4096, (chunk_mb << 20) / ((int64_t)dim * 4));
    cr = (cr + 127) / 128 * 128;
    while ((max_rows + cr - 1) / cr > 64) cr *= 2;
    c->chunk_rows = cr;
#define VQ_CTX(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { vqb200_host_ctx_destroy(c); return cuda_fail(e__); } } while (0)
    VQ_CTX(cudaMalloc(&c->d_x, (size_t)max_rows * dim * 4));
    VQ_CTX(cudaMalloc(&c->d_q, (size_t)max_rows * dim * 4));
    VQ_CTX(cudaMalloc(&c->d_ind, (size_t)max_rows * 8));
    VQ_CTX(cudaMalloc(&c->d_diff, 256));
    VQ_CTX(cudaMalloc(&c->d_stats, vqb200_stats_bytes(dim, n_embed)));
    VQ_CTX(cudaMalloc(&c->d_scratch, forward_scratch_bytes(max_rows, dim, n_embed)));
    VQ_CTX(cudaMalloc(&c->d_codebook, codebook_bytes(dim, n_embed)));
    VQ_CTX(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
    VQ_CTX(cudaStreamCreateWithFlags(&c->s_run, cudaStreamNonBlocking));
    VQ_CTX(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < 64; ++i) {
        VQ_CTX(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
        VQ_CTX(cudaEventCreateWithFlags(&c->ev_run[i], cudaEventDisableTiming));
    }
    VQ_CTX(cudaEventCreateWithFlags(&c->ev_start, cudaEventDisableTiming));
#undef VQ_CTX
    *out = c;
    return VQB200_OK;
}

int vqb200_host_ctx_set_stream(vqb200_host_ctx* c, void* stream) {
    if (!c) return VQB200_EINVAL;
    c->caller = (cudaStream_t)stream;
    return VQB200_OK;
}

void vqb200_host_ctx_destroy(vqb200_host_ctx* c) {
    if (!c) return;
    cudaFree(c->d_x); cudaFree(c->d_q); cudaFree(c->d_ind); cudaFree(c->d_diff);
    cudaFree(c->d_stats); cudaFree(c->d_scratch); cudaFree(c->d_codebook);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_run) cudaStreamDestroy(c->s_run);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    for (int i = 0; i < 64; ++i) {
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
        if (c->ev_run[i]) cudaEventDestroy(c->ev_run[i]);
    }
    if (c->ev_start) cudaEventDestroy(c->ev_start);
    delete c;
}

// `stats_out` != nullptr: leave the batch statistics there and skip the EMA (the caller reduces them across ranks first)
static int host_run(vqb200_host_ctx* c, const float* h_x, int64_t n_rows, float* d_embed,
                    float* d_cluster_size, float* d_embed_avg, float decay, float one_minus_decay,
                    float eps, int32_t training, float* stats_out, float* h_quantize, int64_t* h_embed_ind, float* h_diff,
                    int32_t engine) {
    if (!c || !d_embed || n_rows < 0 || n_rows > c->max_rows) return VQB200_EINVAL;
    if (n_rows > 0 && (!h_x || !h_embed_ind)) return VQB200_EINVAL;
    if (training && !stats_out && (!d_cluster_size || !d_embed_avg)) return VQB200_EINVAL;
    const int D = c->dim, K = c->n_embed;
    VQ_CUDA(cudaSetDevice(c->device));
    // the private streams are non-blocking: order them behind whatever the caller's stream did to d_embed / cluster_size /
    // embed_avg before this call (the call itself returns only after its own streams have drained)
    VQ_CUDA(cudaEventRecord(c->ev_start, c->caller));
    VQ_CUDA(cudaStreamWaitEvent(c->s_run, c->ev_start, 0));
    int rc = prepare_codebook(d_embed, D, K, c->d_codebook, c->s_run);
    if (rc) return rc;
    float* stats = stats_out ? stats_out : training ? c->d_stats : nullptr;
    // row chunks: full-size chunks in the middle (large copies keep both PCIe directions efficient), a short ramp at
    // the start (the first device->host copy can begin early) and at the end (short tail after the last host->device copy)
    int64_t begin[65];
    int64_t nchunks = 0;
    {
        const int64_t full = c->chunk_rows;
        int64_t pos = 0, step = std::max<int64_t>(4096, full / 4);
        begin[0] = 0;
        while (pos < n_rows && nchunks < 63) {
            int64_t left = n_rows - pos;
            int64_t take = std::min(step, left);
            if (left - take > 0 && left - take < full / 4) take = left;          // no tiny last chunk
            else if (left > take && left <= full + full / 2 && take == full) take = left - full / 4;   // taper the end
            pos += take;
            begin[++nchunks] = pos;
            step = std::min(full, step * 2);
        }
        if (pos < n_rows) begin[nchunks] = n_rows;                                 // (ran out of slots: last chunk takes the rest)
        if (nchunks == 0) { nchunks = 1; begin[1] = 0; }
    }
    // all host->device copies are queued first: they depend on nothing but the host buffer, and a copy stream that is
    // fed chunk by chunk between kernel launches runs with ~40 us gaps per chunk (host-issue bound)
    for (int64_t i = 0; i < nchunks; ++i) {
        int64_t r0 = begin[i], rows = begin[i + 1] - begin[i];
        if (rows > 0) {
            VQ_CUDA(cudaMemcpyAsync(c->d_x + r0 * D, h_x + r0 * D, (size_t)rows * D * 4, cudaMemcpyHostToDevice, c->s_in));
            VQ_CUDA(cudaEventRecord(c->ev_in[i], c->s_in));
        }
    }
    for (int64_t i = 0; i < nchunks; ++i) {
        int64_t r0 = begin[i], rows = begin[i + 1] - begin[i];
        if (rows > 0) VQ_CUDA(cudaStreamWaitEvent(c->s_run, c->ev_in[i], 0));
        RowLayout L{rows, rows > 0 ? rows : 1, 0, D, 1};
        rc = forward_impl(c->d_x + r0 * D, L, D, K, c->d_codebook, h_quantize ? c->d_q + r0 * D : nullptr,
                          c->d_ind + r0, c->d_diff, stats, c->d_scratch, engine, i == 0, i == nchunks - 1, n_rows,
                          c->s_run, nullptr, c->max_rows);
        if (rc) return rc;
        if (rows > 0) {
            VQ_CUDA(cudaEventRecord(c->ev_run[i], c->s_run));
            VQ_CUDA(cudaStreamWaitEvent(c->s_out, c->ev_run[i], 0));
            if (h_quantize)
                VQ_CUDA(cudaMemcpyAsync(h_quantize + r0 * D, c->d_q + r0 * D, (size_t)rows * D * 4,
                                        cudaMemcpyDeviceToHost, c->s_out));
            VQ_CUDA(cudaMemcpyAsync(h_embed_ind + r0, c->d_ind + r0, (size_t)rows * 8, cudaMemcpyDeviceToHost, c->s_out));
        }
    }
    if (h_diff) VQ_CUDA(cudaMemcpyAsync(h_diff, c->d_diff, 4, cudaMemcpyDeviceToHost, c->s_run));
    if (training && !stats_out) {
        rc = ema_impl(c->d_stats, nullptr, d_cluster_size, d_embed_avg, d_embed, D, K, decay, one_minus_decay, eps,
                      c->d_codebook, c->s_run);
        if (rc) return rc;
    }
    VQ_CUDA(cudaStreamSynchronize(c->s_out));
    VQ_CUDA(cudaStreamSynchronize(c->s_run));
    return VQB200_OK;
}

int vqb200_host_quantize(vqb200_host_ctx* c, const float* h_x, int64_t n_rows, float* d_embed,
                         float* d_cluster_size, float* d_embed_avg, float decay, float one_minus_decay,
                         float eps, int32_t training, float* h_quantize, int64_t* h_embed_ind, float* h_diff,
                         int32_t engine) {
    return host_run(c, h_x, n_rows, d_embed, d_cluster_size, d_embed_avg, decay, one_minus_decay, eps, training, nullptr,
                    h_quantize, h_embed_ind, h_diff, engine);
}

int vqb200_host_quantize_stats(vqb200_host_ctx* c, const float* h_x, int64_t n_rows, float* d_embed, float* d_stats,
                               float* h_quantize, int64_t* h_embed_ind, float* h_diff, int32_t engine) {
    if (!d_stats) return VQB200_EINVAL;
    return host_run(c, h_x, n_rows, d_embed, nullptr, nullptr, 0.f, 0.f, 0.f, 1, d_stats, h_quantize, h_embed_ind, h_diff,
                    engine);
}

}  // extern "C"
