/*
 * vqb200.h -- C ABI of the B200-native VQ-VAE-2 quantizer hot path.
 *
 * This is the drop-in boundary for ONE path of alehdaghi/vq-vae-2-pytorch: the `Quantize`
 * module (reference vqvae.py:28-78).  The reference has no FFI of its own (it is pure PyTorch),
 * so every entry point below names the reference expression(s) it replaces; the Python class
 * `vq_vae_2_pytorch_b200.Quantize` binds them with ctypes and keeps the reference's module
 * surface (INTEGRATION.md shows the binding a reference maintainer would add).
 *
 * Conventions
 *   - plain C types only: raw pointers, sizes, strides, a `void* stream` (cudaStream_t);
 *   - every pointer named d_* is DEVICE memory on the current device, h_* is HOST memory;
 *   - no allocation inside the device entry points: the caller passes workspaces whose sizes
 *     come from the vqb200_*_bytes() queries; all work is enqueued on `stream`, nothing syncs;
 *   - return value: 0 on success, a negative VQB200_E* code otherwise (vqb200_error_string()).
 *   - arithmetic type: fp32 in / fp32 out, int64 indices (the reference's dtypes).
 *
 * Row layout ("vq_layout"): the module receives a [..., D] tensor whose leading dims flatten to
 * N rows (vqvae.py:43).  Two physical layouts are accepted without a copy:
 *     element (n, d) lives at   (n / rows_per_image) * image_stride
 *                             + (n % rows_per_image) * row_stride + d * col_stride
 *   contiguous [N, D]:            rows_per_image = N,   image_stride = 0, row_stride = D, col_stride = 1
 *   permute(0,2,3,1) of NCHW      rows_per_image = H*W, image_stride = D*H*W, row_stride = 1,
 *   (what VQVAE.encode passes,                          col_stride = H*W
 *    vqvae.py:227,235):
 * `quantize` is written with the same layout as the input (vqvae.py:73 keeps the input's strides).
 */
#ifndef VQB200_H_
#define VQB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQB200_ABI_VERSION 1

#define VQB200_OK             0
#define VQB200_EINVAL        -1   /* bad argument (null pointer, non-positive size, ...)            */
#define VQB200_EUNSUPPORTED  -2   /* shape / layout outside what the kernels cover                  */
#define VQB200_ECUDA         -3   /* a CUDA call failed; see vqb200_last_cuda_error()               */
#define VQB200_ENODEVICE     -4   /* no sm_100 device                                               */

/* assignment engines for vqb200_quantize_forward / vqb200_assign */
/* tcgen05 coverage: dim 64 with n_embed 256 / 512 / k*512 <= 16384 (dense or NCHW-physical rows), and the wide engine
 * for dim 128 (n_embed 256 or k*512) and dim 256 (n_embed k*256), n_embed <= 16384, dense rows (vqvae_deep.py:252,257).
 * At dim 256 with n_embed k*512 the wide engine runs as CTA pairs (tcgen05.mma.cta_group::2): 512 codes per pass.           */
#define VQB200_ENGINE_AUTO    0   /* tcgen05 path when the shape is covered, else exact SIMT        */
#define VQB200_ENGINE_SIMT    1   /* exact fp32 SIMT distance kernel                                */
#define VQB200_ENGINE_TCGEN05 2   /* TMA + tcgen05 split-bf16 filter with exact fp32 re-score       */
#define VQB200_ENGINE_TCGEN05_BF16 3 /* same kernel, plain-bf16 filter: 1/3 of the MMAs, wider bound (more exact re-scores) */
#define VQB200_ENGINE_TCGEN05_TF32 4 /* kind::tf32 MMAs straight from the fp32 x tile (no conversion pass), dense rows, dim 64 */

int         vqb200_abi_version(void);
const char* vqb200_error_string(int code);
int         vqb200_last_cuda_error(void);      /* cudaError_t of the last failing CUDA call, 0 if none */
uint64_t    vqb200_launch_count(void);         /* kernels launched by this library so far (process-wide) */

/* ---- workspace sizes (bytes) ------------------------------------------------------------------ */
/* prepared codebook image: code-major fp32 copy, ||e_k||^2, tensor-core operand images            */
size_t vqb200_codebook_bytes(int32_t dim, int32_t n_embed);
/* per-call scratch of the forward (diff accumulator, flagged-row list, counters, rows-per-code counters, per-CTA statistics
 * tables where the shape allows them).  Layout contract used by adaptive callers: int32 at byte offset 16 = number of
 * rows the last tcgen05 forward sent to the exact re-score; int32 at byte offset 56 != 0 = the code-statistics kernel
 * met an index outside [0, n_embed) (internal error: the indices are written by this library's own kernels).         */
size_t vqb200_forward_scratch_bytes(int64_t n_rows, int32_t dim, int32_t n_embed);
/* packed codebook statistics: n_embed*dim per-code sums (code-major), then n_embed counts, then 4
 * spare words the EMA kernel uses as scalars; only the first n_embed*(dim+1) floats are all-reduced.
 * The spare words must be zero when vqb200_ema_update starts: vqb200_quantize_forward leaves them zero
 * and so does vqb200_ema_update.                                                                    */
size_t vqb200_stats_bytes(int32_t dim, int32_t n_embed);

/* ---- codebook ---------------------------------------------------------------------------------- */
/* Re-derive the prepared image from `embed` [dim, n_embed] (the module buffer, vqvae.py:37-38).
 * Must be called after any external write to `embed` (load_state_dict, .to(), DDP broadcast).     */
int vqb200_codebook_prepare(const float* d_embed, int32_t dim, int32_t n_embed,
                            void* d_codebook, void* stream);

/* ---- forward ----------------------------------------------------------------------------------- */
/* vqvae.py:43-52,72-73 (+ the statistics of :50,55-56 when d_stats != NULL), one pass over x:
 *   embed_ind[n] = argmin_k ||x_n - e_k||^2 (lowest k on ties)              -> d_embed_ind (int64 [N])
 *   quantize     = x + (e[embed_ind] - x)                                   -> d_quantize (layout of x)
 *   diff         = mean((e[embed_ind] - x)^2)                               -> d_diff (1 float)
 *   stats        = [sum of rows per code | rows per code], zeroed here      -> d_stats (training only)
 * d_quantize may be NULL (index extraction only, extract_code.py:23), d_diff may be NULL.          */
int vqb200_quantize_forward(const float* d_x, int64_t n_rows, int32_t dim, int32_t n_embed,
                            int64_t rows_per_image, int64_t image_stride,
                            int64_t row_stride, int64_t col_stride,
                            const void* d_codebook,
                            float* d_quantize, int64_t* d_embed_ind, float* d_diff,
                            float* d_stats, void* d_scratch, int32_t engine, void* stream);

/* vqvae.py:61-70 after the (optional) all-reduce of vqvae.py:58-59: EMA of cluster_size / embed_avg,
 * Laplace-smoothed renormalisation, in-place write of `embed`, and refresh of the prepared codebook
 * image for the next forward.  `one_minus_decay` is passed separately because the reference evaluates
 * `1 - decay` in double precision (vqvae.py:62).  d_codebook may be NULL (no image refresh).         */
int vqb200_ema_update(const float* d_stats, float* d_cluster_size, float* d_embed_avg, float* d_embed,
                      int32_t dim, int32_t n_embed, float decay, float one_minus_decay, float eps,
                      void* d_codebook, void* stream);

/* Multi-GPU form of vqb200_ema_update with the all-reduce of vqvae.py:58-59 (distributed.py:64-72) FUSED into the EMA
 * kernel: no NCCL call and no extra launch.  Every rank keeps its packed statistics of the step in peer-mapped
 * (symmetric) memory; h_stats_ptrs[r] / h_flag_ptrs[r] are HOST arrays of `world` DEVICE pointers, valid in THIS
 * process, to rank r's statistics buffer of this step and to rank r's flag array (world x uint32, zero before the
 * first step).  The kernel publishes "rank `rank` finished step `step`" to every peer (system-scope release store),
 * waits until all ranks have published `step`, sums the per-rank statistics over NVLink in rank order (same bits on
 * every rank) and applies vqvae.py:61-70.  `step` must grow by one per call (never 0) and the caller alternates
 * between two statistics buffers (and two flag arrays) so that a rank one step ahead cannot overwrite what a peer
 * still reads.  world <= 8; shapes the fused EMA kernel covers (dim 64, n_embed 256/512), else VQB200_EUNSUPPORTED. */
int vqb200_ema_update_p2p(const void* const* h_stats_ptrs, void* const* h_flag_ptrs, int32_t rank, int32_t world,
                          uint32_t step, float* d_cluster_size, float* d_embed_avg, float* d_embed,
                          int32_t dim, int32_t n_embed, float decay, float one_minus_decay, float eps,
                          void* d_codebook, void* stream);
/* The same exchange for ANY packed statistics buffer (shapes outside the fused kernel below: D = 128 / 256, K >= 1024):
 * an in-place all-reduce (SUM) of d_stats[0 .. n_words) over peer memory -- what dist_fn.all_reduce does at vqvae.py:58-59,
 * without NCCL: every word is stored as a {value, step} pair into every rank's receive slot and the local slots are summed
 * in rank order (bit-identical replicas).  Pointer arrays, d_err and d_step_counter as in vqb200_quantize_step_peers; every
 * slot holds n_words pairs; d_step_counter points at TWO zero-initialised words (counter, launch ticket).  Follow with
 * vqb200_ema_update(d_stats, ...).                                                                                      */
int vqb200_stats_exchange_peers(float* d_stats, int64_t n_words, void* const* h_push_dst, const void* const* h_recv,
                                void* d_err, void* d_step_counter, int32_t rank, int32_t world, void* stream);
/* Multi-rank training forward in ONE call, with the exchange that replaces dist_fn.all_reduce (vqvae.py:58-59 ->
 * distributed/distributed.py:64-72) fused into the EMA kernel over peer memory; dim 64 / n_embed 256 or 512.  Forward as
 * vqb200_quantize_step, then ONE kernel folds the per-CTA statistics tables, stores every word of this rank's packed
 * statistics [K*64 sums | K counts] as an 8-byte {value, step} pair into rank r's receive slot for THIS rank (peer-mapped
 * pointers, 8-byte aligned, K*65 pairs each; r == rank: the local slot), polls the words it needs in its LOCAL receive slots
 * until their tag equals `step` (flag-in-data: no fence, no flag array, no NCCL call, no remote load), adds them in rank
 * order (bit-identical replicas) and applies vqvae.py:61-70.
 *   h_push_dst / h_recv : [2][world] pointers, parity-major -- the slots are double-buffered on the parity of the step;
 *   d_step_counter      : one local device word, zero-initialised once, owned by the kernel: step = counter + 1, advanced
 *                         by the last block -- the tag never passes through the host, so the call sequence can be replayed
 *                         from a CUDA graph and a host-side exception cannot put one rank out of step;
 *   d_err               : 16 words of local device memory, zero-initialised; a word that does not arrive within 2 s makes
 *                         the kernel record d_err[0] = step, d_err[1] = missing rank and carry on.                          */
int vqb200_quantize_step_peers(const float* d_x, int64_t n_rows, int32_t dim, int32_t n_embed, int64_t rows_per_image,
                               int64_t image_stride, int64_t row_stride, int64_t col_stride, float* d_embed,
                               float* d_cluster_size, float* d_embed_avg, void* d_codebook, float* d_quantize,
                               int64_t* d_embed_ind, float* d_diff, void* d_scratch, float* d_x_dense, int32_t engine,
                               float decay, float one_minus_decay, float eps, void* const* h_push_dst,
                               const void* const* h_recv, void* d_err, void* d_step_counter, int32_t rank, int32_t world,
                               void* stream);

/* The module's whole forward in one call (fewer host round trips per step):
 *   vqb200_codebook_prepare(d_embed) + vqb200_quantize_forward(...) and, when `ema` != 0 and d_stats != NULL,
 *   the EMA of vqvae.py:61-70 on the statistics of this call (single-process training; at dim 64 / n_embed 256 or 512 ONE
 *   kernel folds the per-CTA statistics tables and applies the EMA, and d_stats is then not written).  With several
 *   ranks the caller uses vqb200_quantize_step_peers, or passes ema = 0, all-reduces d_stats (vqvae.py:58-59) and calls
 *   vqb200_ema_update itself.
 * d_cluster_size / d_embed_avg are only touched when the EMA runs.
 * d_x_dense (may be NULL): n_rows*dim floats of scratch, 32-byte aligned (written with 256-bit stores).  With it, NCHW-physical rows (unit row stride, the
 * permute(0,2,3,1) view of vqvae.py:227,235; whole 128-row tiles per image) are consumed IN PLACE by the tensor-core
 * kernel in training mode too: its converters write the dense copy the code-statistics kernel gathers from.     */
int vqb200_quantize_step(const float* d_x, int64_t n_rows, int32_t dim, int32_t n_embed,
                         int64_t rows_per_image, int64_t image_stride, int64_t row_stride, int64_t col_stride,
                         float* d_embed, float* d_cluster_size, float* d_embed_avg, void* d_codebook,
                         float* d_quantize, int64_t* d_embed_ind, float* d_diff, float* d_stats,
                         void* d_scratch, float* d_x_dense, int32_t engine, int32_t ema, float decay,
                         float one_minus_decay, float eps, void* stream);

/* Re-pack between a strided row layout (same layout arguments as vqb200_quantize_forward; e.g. the NCHW-physical
 * permute(0,2,3,1) view VQVAE.encode passes, vqvae.py:227,235) and dense [n_rows, dim] rows, coalesced on both sides.
 * to_dense != 0: d_src strided -> d_dst dense;  to_dense == 0: d_src dense -> d_dst strided.  The module uses it to
 * feed strided inputs to the tcgen05 engine (which stages dense 256-byte rows) and to write `quantize` back with the
 * input's strides (vqvae.py:73).                                                                                    */
int vqb200_repack_rows(const float* d_src, float* d_dst, int64_t n_rows, int32_t dim, int64_t rows_per_image,
                       int64_t image_stride, int64_t row_stride, int64_t col_stride, int32_t to_dense, void* stream);

/* Gradient implied by vqvae.py:72-73:  grad_x = grad_quantize + grad_diff * 2 (x - e[ind]) / (N*D).
 * d_grad_quantize (same layout as x) and d_grad_diff (1 float) may each be NULL (= zero).           */
int vqb200_quantize_backward(const float* d_x, const int64_t* d_embed_ind, const void* d_codebook,
                             const float* d_grad_quantize, const float* d_grad_diff, float* d_grad_x,
                             int64_t n_rows, int32_t dim, int32_t n_embed,
                             int64_t rows_per_image, int64_t image_stride,
                             int64_t row_stride, int64_t col_stride, void* stream);

/* vqvae.py:77-78 `embed_code`: out[n, :] = e[:, embed_id[n]]  (contiguous [N, dim] output).
 * Returns VQB200_EINVAL semantics on the device side: out-of-range ids are reported through
 * d_status (1 int32, set non-zero) like F.embedding's index check; d_status may be NULL.           */
int vqb200_embed_code(const int64_t* d_embed_id, int64_t n_rows, const void* d_codebook,
                      int32_t dim, int32_t n_embed, float* d_out, int32_t* d_status, void* stream);

/* Index egress for code extraction (extract_code.py:23-33 copies the int64 `id_t` / `id_b` of every batch to the host and
 * pickles them per image; dataset.py:45-51 reads them back).  n_embed <= 65536 fits 16 bits: pack embed_ind [n] int64 into
 * out_bytes = 2 (uint16) or 4 (int32) per code on the device so that the D2H copy moves 4x / 2x fewer bytes.  d_status
 * (1 int32, may be NULL) is set non-zero when an index lies outside [0, n_embed).  Both pointers 16-byte aligned.       */
int vqb200_pack_indices(const int64_t* d_embed_ind, int64_t n, int32_t n_embed, int32_t out_bytes, void* d_out,
                        int32_t* d_status, void* stream);
/* the inverse (codes uploaded for `embed_code` / `decode_code`, vqvae.py:251-259, sample.py:92-97): widen to int64 */
int vqb200_unpack_indices(const void* d_codes, int64_t n, int32_t in_bytes, int64_t* d_embed_id, void* stream);

/* ---- diagnostics ---------------------------------------------------------------------------------- */
/* Runs the tcgen05 engine on contiguous rows and additionally dumps the tensor-core scores
 * (certified lower bounds of ||x_n - e_k||^2 + row offset) to d_scores [n_rows, n_embed]; the tests use
 * it to check the error-bound certificate against float64.  d_flagged_count (1 int32, may be NULL)
 * receives the number of rows that were sent to the exact re-score.                                  */
int vqb200_debug_tc_scores(const float* d_x, int64_t n_rows, int32_t dim, int32_t n_embed,
                           const void* d_codebook, int64_t* d_embed_ind, float* d_scores,
                           int32_t* d_flagged_count, void* d_scratch, void* stream);
/* same, with the filter chosen by `engine` (VQB200_ENGINE_TCGEN05 / _BF16 / _TF32) */
int vqb200_debug_tc_scores_ex(const float* d_x, int64_t n_rows, int32_t dim, int32_t n_embed,
                              const void* d_codebook, int64_t* d_embed_ind, float* d_scores,
                              int32_t* d_flagged_count, void* d_scratch, int32_t engine, void* stream);
/* Same run with per-role cycle counters: d_prof [n_CTAs(<=160)][vqb200_tc_profile_slots()] uint64, zeroed by
 * the caller; slot meaning = enum ProfSlot in csrc/tc_kernel.cuh (pipeline bubble analysis for profiles/). */
int vqb200_debug_tc_profile(const float* d_x, int64_t n_rows, int32_t dim, int32_t n_embed,
                            const void* d_codebook, float* d_quantize, int64_t* d_embed_ind,
                            void* d_scratch, uint64_t* d_prof, int32_t engine, void* stream);
/* The tensor-core kernel ALONE (tc::k_vq_tc: assignment + gather + straight-through value + loss partial sums), no
 * fix-up / finalisation launches: what bench.py times with CUDA events for the roofline of the dominant kernel.  The
 * caller prepares the codebook image and clears the first 256 bytes of d_scratch once; uncertified rows are appended
 * to the flagged list and otherwise left untouched.                                                           */
int vqb200_debug_tc_kernel(const float* d_x, int64_t n_rows, int32_t dim, int32_t n_embed, const void* d_codebook,
                           float* d_quantize, int64_t* d_embed_ind, void* d_scratch, int32_t engine, void* stream);
int vqb200_tc_profile_slots(void);
/* Flag ping-pong between two GPUs over peer-mapped memory (tools/p2p_latency.py): `iters` round trips of a system-scope
 * release store to d_peer_flag and an acquire spin on d_my_flag; *d_ns = elapsed globaltimer nanoseconds.  Both sides must
 * be running at the same time on DIFFERENT GPUs (never two ranks on one GPU).                                              */
int vqb200_debug_pingpong(void* d_my_flag, void* d_peer_flag, int32_t iters, int32_t initiator, int32_t with_fence,
                          uint64_t* d_ns, void* stream);
/* 1 when vqb200_quantize_forward would take the tcgen05 engine for this shape / layout / pointer alignment */
int vqb200_tc_supported(const float* d_x, int64_t n_rows, int32_t dim, int32_t n_embed,
                        int64_t rows_per_image, int64_t image_stride, int64_t row_stride, int64_t col_stride);
int vqb200_tc_split(void);   /* 3 = split-bf16 filter (default), 1 = plain bf16 (env VQB200_TC_SPLIT) */

/* ---- host-buffer convenience path (what bench.py's `e2e` times) -------------------------------- */
/* Same contract as vqb200_quantize_forward (+ optional EMA update when `training`), but x / quantize /
 * embed_ind / diff are HOST buffers (pinned for full speed); H2D and D2H copies are pipelined in row
 * chunks across internal streams.  The context owns device staging memory sized for max_rows.       */
typedef struct vqb200_host_ctx vqb200_host_ctx;
int  vqb200_host_ctx_create(int64_t max_rows, int32_t dim, int32_t n_embed, vqb200_host_ctx** out);
/* The context's private copy / compute streams are non-blocking.  Every vqb200_host_quantize call first orders them behind
 * the work already queued on the CALLER's stream (so writes to d_embed / d_cluster_size / d_embed_avg issued there just
 * before the call are observed), makes the context's device current, and returns only after its own streams have drained.
 * The caller's stream is the legacy default stream unless set here (e.g. torch.cuda.current_stream().cuda_stream).        */
int  vqb200_host_ctx_set_stream(vqb200_host_ctx* ctx, void* stream);
void vqb200_host_ctx_destroy(vqb200_host_ctx* ctx);
/* device-resident module buffers (embed / cluster_size / embed_avg) stay on the device */
int  vqb200_host_quantize(vqb200_host_ctx* ctx, const float* h_x, int64_t n_rows,
                          float* d_embed, float* d_cluster_size, float* d_embed_avg,
                          float decay, float one_minus_decay, float eps, int32_t training,
                          float* h_quantize, int64_t* h_embed_ind, float* h_diff, int32_t engine);
/* Data-parallel form: the same pipelined forward, with the batch statistics of vqvae.py:55-56 left in the caller's
 * DEVICE buffer d_stats (vqb200_stats_bytes) and NO EMA -- the caller reduces d_stats across ranks where the reference
 * calls dist_fn.all_reduce (vqvae.py:58-59; one all-reduce of the packed buffer) and then applies vqvae.py:61-70 with
 * vqb200_ema_update(d_stats, ...).                                                                                       */
int  vqb200_host_quantize_stats(vqb200_host_ctx* ctx, const float* h_x, int64_t n_rows, float* d_embed, float* d_stats,
                                float* h_quantize, int64_t* h_embed_ind, float* h_diff, int32_t engine);

#ifdef __cplusplus
}
#endif
#endif /* VQB200_H_ */
