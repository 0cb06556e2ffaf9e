// Small-kernel fusions for the shapes the tcgen05 engine covers (D = 64): at the headline shape the step is ~100 us, so
// every extra launch (~2.5 us each) shows.  One launch each for: the codebook image, the flagged-row fix-up with the
// loss finalisation, and the EMA update with the next image.
#pragma once
#include "common.cuh"
#include "simt_kernels.cuh"
#include "tc_kernel.cuh"

namespace vqb200 {

constexpr int PREP_CODES = 8;      // codes per block of the image / EMA kernels (one warp per code)

// ------------------------------------------------------------------------------------------------
// codebook image in one launch: embed [64, K] -> cbT [K, 64], ||e_k||^2 (same summation order as
// k_codebook_norms) and the tensor-core operand image.  grid = K / 8, block = 256.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_prepare64(const float* __restrict__ embed, float* __restrict__ cbT,
                                                    float* __restrict__ ee, unsigned char* __restrict__ img, int K,
                                                    int slice_codes /* 0: one image of K codes; else sub-images of that many */,
                                                    float cA, float cA1, float cB,
                                                    unsigned int* __restrict__ zero_header /* may be null: 256-byte forward scratch header */) {
    __shared__ float es[PREP_CODES][65];
    __shared__ float e2s[PREP_CODES];
    pdl_wait();
    pdl_trigger();
    const int k0 = blockIdx.x * PREP_CODES, tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    if (zero_header) {                        // (saves a memset node in the chain)
        if (blockIdx.x == 0 && tid < 64) zero_header[tid] = 0u;              // loss accumulator, flagged-row counter, tickets
        if (tid < PREP_CODES) zero_header[64 + k0 + tid] = 0u;               // rows-per-code counters of this call (ForwardScratch::code_counts)
        // per-chunk tickets of the fix-up (ForwardScratch::fix_tickets, behind the 256-byte aligned counters)
        const int n_tick = FIX_CAP / 64, per = (n_tick + (int)gridDim.x - 1) / (int)gridDim.x;
        const int tbase = 64 + (int)(align_up((size_t)K * 4, 256) / 4);
        if (tid < per && blockIdx.x * per + tid < n_tick) zero_header[tbase + blockIdx.x * per + tid] = 0u;
    }
    for (int i = tid; i < 64 * PREP_CODES; i += 256) {          // 32-byte segments of 8 consecutive codes per dim
        const int d = i >> 3, j = i & 7;
        es[j][d] = embed[(size_t)d * K + k0 + j];
    }
    __syncthreads();
    const float v0 = es[w][lane], v1 = es[w][lane + 32];
    cbT[(size_t)(k0 + w) * 64 + lane] = v0;
    cbT[(size_t)(k0 + w) * 64 + lane + 32] = v1;
    float s = fmaf(v0, v0, 0.f);
    s = fmaf(v1, v1, s);
    s = warp_sum(s);
    if (lane == 0) { ee[k0 + w] = s; e2s[w] = s; }
    __syncthreads();
    if (tid < 8 * PREP_CODES) {
        const int k = k0 + (tid >> 3);
        if (slice_codes == 0) tc::tc_image_rows(&es[tid >> 3][0], e2s[tid >> 3], img, K, k, tid & 7, cA, cA1, cB);
        else tc::tc_image_rows(&es[tid >> 3][0], e2s[tid >> 3], img + (size_t)(k / slice_codes) * tc::image_bytes(slice_codes),
                               slice_codes, k % slice_codes, tid & 7, cA, cA1, cB);
    }
}

// ------------------------------------------------------------------------------------------------
// EMA update + Laplace-smoothed renormalisation (vqvae.py:61-70) + next codebook image in one launch.
// grid = K / 8, block = 256.  Every block derives n = sum_k cluster_size_new[k] from the OLD cluster sizes,
// which nobody overwrites until the last block to finish (ticket) stores the new ones.
//
// P2P = true fuses the all-reduce of vqvae.py:58-59 into this kernel (no NCCL call, no extra launch): every rank's
// packed statistics live in peer-mapped (symmetric) memory, block 0 publishes "my statistics for this step are
// complete" to all peers with a system-scope release store, every block waits for all ranks' flags and then sums the
// per-rank statistics over NVLink in RANK ORDER (identical result on every rank, so the replicas stay bit-identical).
// ------------------------------------------------------------------------------------------------
constexpr int P2P_MAX_RANKS = 8;
constexpr unsigned long long P2P_TIMEOUT_NS = 2000000000ull;   // 2 s: far beyond any step, short enough to fail loudly
struct PeerStats {
    const float* stats[P2P_MAX_RANKS];     // rank r's packed statistics of this step (this process' mapping)
    unsigned int* flags[P2P_MAX_RANKS];    // rank r's flag array [world]: flags[r][w] = last step rank w has published
    int rank, world;
    unsigned int step;
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// sum over the ranks, in rank order, of element i of every rank's statistics buffer.  All loads are issued before the
// first add (one NVLink round trip instead of `world`); volatile: never from a stale local cache line, never hoisted
// above the flag wait.
__device__ __forceinline__ void peer_load(const PeerStats& peers, size_t i, bool on, float (&v)[P2P_MAX_RANKS]) {
#pragma unroll
    for (int r = 0; r < P2P_MAX_RANKS; ++r)
        v[r] = (on && r < peers.world) ? *reinterpret_cast<const volatile float*>(peers.stats[r] + i) : 0.f;
}
__device__ __forceinline__ float rank_ordered_sum(const float (&v)[P2P_MAX_RANKS]) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < P2P_MAX_RANKS; ++r) s += v[r];          // + 0.f for absent ranks: exact
    return s;
}

template <bool P2P>
__global__ void __launch_bounds__(256) k_ema64(const float* __restrict__ stats, PeerStats peers,
                                                float* __restrict__ cluster_size,
                                                float* __restrict__ embed_avg, float* __restrict__ embed,
                                                float* __restrict__ cbT, float* __restrict__ ee,
                                                unsigned char* __restrict__ img, int K, float decay,
                                                float one_minus_decay, float eps, float cA, float cA1, float cB,
                                                unsigned int* __restrict__ ticket) {
    __shared__ float es[PREP_CODES][65];
    __shared__ float ssum[PREP_CODES][65];
    __shared__ float cnt_s[512];                                // K <= 512 (tc_shape_ok)
    __shared__ float e2s[PREP_CODES];
    __shared__ float part[8];
    __shared__ float n_s;
    __shared__ unsigned int last_s;
    pdl_wait();
    pdl_trigger();
    const int k0 = blockIdx.x * PREP_CODES, tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    if (P2P) {
        if (blockIdx.x == 0 && tid < peers.world) {             // my statistics (written by the previous kernels) are complete
            __threadfence_system();
            st_release_sys(peers.flags[tid] + peers.rank, peers.step);
        }
#ifdef VQB200_P2P_TRACE
        unsigned long long t0 = 0, t1 = 0;
        if (blockIdx.x == 0 && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
#endif
        if (tid < peers.world) {
            // bounded wait (a crashed or desynchronised peer must not hang this GPU for ever): after P2P_TIMEOUT_NS the
            // block gives up, records (step, missing rank) in word 32 / 33 of its flag array -- the host reads it back
            // (Quantize raises) -- and carries on with whatever the slots hold
            const unsigned int* f = peers.flags[peers.rank] + tid;
            unsigned long long t0 = 0;
            unsigned int spins = 0;
            while ((int)(ld_acquire_sys(f) - peers.step) < 0) {
                __nanosleep(64);
                if ((++spins & 1023u) == 0) {
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    if (t0 == 0) t0 = now;
                    else if (now - t0 > P2P_TIMEOUT_NS) {
                        unsigned int* err = peers.flags[peers.rank] + 32;
                        err[1] = (unsigned int)tid;
                        atomicExch(err, peers.step);
                        break;
                    }
                }
            }
        }
        __syncthreads();
#ifdef VQB200_P2P_TRACE
        if (blockIdx.x == 0 && tid == 0) {       // ring of (kernel start, wait) in the unused tail of my flag array
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            unsigned int* ring = peers.flags[peers.rank] + 8 + ((peers.step >> 1) % 14) * 4;
            ring[0] = (unsigned int)(t0 & 0xffffffffu); ring[1] = (unsigned int)(t1 - t0); ring[2] = peers.step;
        }
#endif

    }
    // counts / sums of this step (summed over the ranks in rank order) -> shared memory, all peer loads in one phase
    float cv[2], sv[2];
    if constexpr (P2P) {
        // ALL peer loads of the thread are issued before the first add: with the adds of one sum between the load groups
        // (what the compiler emits for four separate sums of volatile loads) the warp paid four NVLink round trips in a row
        float vc[2][P2P_MAX_RANKS], vs[2][P2P_MAX_RANKS];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int k = tid + 256 * u, j = k >> 6, d = k & 63;
            peer_load(peers, (size_t)K * 64 + k, k < K, vc[u]);
            peer_load(peers, (size_t)(k0 + j) * 64 + d, true, vs[u]);
        }
        asm volatile("" ::: "memory");               // keep the adds behind the last load
#pragma unroll
        for (int u = 0; u < 2; ++u) { cv[u] = rank_ordered_sum(vc[u]); sv[u] = rank_ordered_sum(vs[u]); }
    } else {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int k = tid + 256 * u;
            cv[u] = k < K ? stats[(size_t)K * 64 + k] : 0.f;
            const int i = tid + 256 * u, j = i >> 6, d = i & 63;   // coalesced 256-byte rows of the 8 codes of this block
            sv[u] = stats[(size_t)(k0 + j) * 64 + d];
        }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int k = tid + 256 * u;
        if (k < K) cnt_s[k] = cv[u];
        const int i = tid + 256 * u;
        ssum[i >> 6][i & 63] = sv[u];
    }
    __syncthreads();
    float s = 0.f;
    for (int k = tid; k < K; k += 256) {
        const float c = __fmaf_rn(cnt_s[k], one_minus_decay, __fmul_rn(cluster_size[k], decay));   // vqvae.py:61-63
        s += c;
    }
    s = warp_sum(s);
    if (lane == 0) part[w] = s;
    __syncthreads();
    if (tid == 0) {
        float n = 0.f;
        for (int i = 0; i < 8; ++i) n += part[i];
        n_s = n;                                                // vqvae.py:65
    }
    __syncthreads();
    const float n = n_s;
    const float denom = n + (float)((double)K * (double)eps);
    for (int i = tid; i < 64 * PREP_CODES; i += 256) {          // 32-byte segments of 8 consecutive codes per dim
        const int d = i >> 3, j = i & 7, k = k0 + j;
        const float c = __fmaf_rn(cnt_s[k], one_minus_decay, __fmul_rn(cluster_size[k], decay));
        const float cs = (c + eps) / denom * n;                 // vqvae.py:66-68
        const size_t o = (size_t)d * K + k;
        const float a = __fmaf_rn(ssum[j][d], one_minus_decay, __fmul_rn(embed_avg[o], decay));       // vqvae.py:64
        embed_avg[o] = a;
        const float e = a / cs;                                 // vqvae.py:69-70
        embed[o] = e;
        es[j][d] = e;
    }
    __syncthreads();
    const float v0 = es[w][lane], v1 = es[w][lane + 32];
    if (cbT) {
        cbT[(size_t)(k0 + w) * 64 + lane] = v0;
        cbT[(size_t)(k0 + w) * 64 + lane + 32] = v1;
    }
    float s2 = fmaf(v0, v0, 0.f);
    s2 = fmaf(v1, v1, s2);
    s2 = warp_sum(s2);
    if (lane == 0) { if (ee) ee[k0 + w] = s2; e2s[w] = s2; }
    __syncthreads();
    if (img && tid < 8 * PREP_CODES)
        tc::tc_image_rows(&es[tid >> 3][0], e2s[tid >> 3], img, K, k0 + (tid >> 3), tid & 7, cA, cA1, cB);
    // ---- deferred cluster_size store: only after every block has read the old values
    __threadfence();
    if (tid == 0) last_s = (atomicAdd(ticket, 1u) == gridDim.x - 1u) ? 1u : 0u;
    __syncthreads();
    if (last_s) {
        for (int k = tid; k < K; k += 256) {
            cluster_size[k] = __fmaf_rn(cnt_s[k], one_minus_decay, __fmul_rn(cluster_size[k], decay));
        }
        if (tid == 0) *ticket = 0u;                             // clean for the next launch
    }
}

// ------------------------------------------------------------------------------------------------
// flagged-row fix-up of the tcgen05 engine in one launch: exact fp32 re-score, then gather / straight-through value /
// loss / statistics (gather_chunk) of the same rows, then -- by the last block to finish -- the loss finalisation
// diff = acc * inv_count.  Exits after the ticket when no row was flagged.
//   many rows : one work item = one 64-row chunk against the whole codebook (assign_chunk); 4 CTAs per SM keep the loads of
//               several chunks in flight (one CTA per SM ran at 5 TFLOP/s).
//   few rows  (fewer chunks than half the CTAs -- measured: beyond that the 8x re-read of the x tiles costs more than the
//               idle SMs --, codebooks of 2 .. FIX_KB 64-code blocks): one work item = one
//               chunk against ONE 64-code block; the partial arg-mins are 64-bit (distance, code) keys whose integer
//               minimum is the reference's arg-min; the last block of a chunk (per-chunk ticket) merges them and produces the
//               chunk's outputs.  492 flagged rows (split-bf16 filter on N(0,1) rows) used to occupy 8 CTAs for ~100 us.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AS_THREADS, 2)
k_fixup(const float* __restrict__ x, RowLayout L, int D, int K, const float* __restrict__ cbT,
        const float* __restrict__ ee, int64_t* __restrict__ embed_ind, float* __restrict__ quantize,
        double* __restrict__ diff_acc, float* __restrict__ stat_sums, float* __restrict__ stat_counts,
        const int* __restrict__ row_list, const int* __restrict__ row_count, int want_gather,
        float* __restrict__ diff, double inv_count, unsigned int* __restrict__ ticket,
        unsigned long long* __restrict__ fix_partial, unsigned int* __restrict__ fix_tickets) {
    extern __shared__ float gs_tile[];
    __shared__ float warp_part[AS_THREADS / 32];
    __shared__ unsigned int last_s;
    __shared__ unsigned int chunk_last_s;
    pdl_wait();
    pdl_trigger();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t total = (int64_t)(*row_count);
    if (total == 0) {
        // nothing was flagged (the common case): no block has anything to add, so block 0 finalises the loss at once instead of
        // every block taking a ticket (296 serialised atomics on one word were ~1.5 us of the step's critical path)
        if (blockIdx.x == 0 && tid == 0 && diff) diff[0] = (float)(*reinterpret_cast<volatile double*>(diff_acc) * inv_count);
        return;
    }
    const int64_t chunks = (total + AS_BM - 1) / AS_BM;
    const int KB = (K + AS_BN - 1) / AS_BN;
    const bool split = fix_partial && fix_tickets && KB > 1 && KB <= FIX_KB && total <= FIX_CAP && 2 * chunks < (int64_t)gridDim.x;
    const int64_t items = split ? chunks * KB : chunks;
    float acc = 0.f;
    for (int64_t w = blockIdx.x; w < items; w += gridDim.x) {
        const int64_t chunk = split ? w / KB : w;
        const int64_t n0 = chunk * AS_BM;
        if (split) {
            const int kb = (int)(w - chunk * KB);
            assign_part(x, L, D, K, cbT, ee, row_list, total, n0, kb * AS_BN, min(K, (kb + 1) * AS_BN), fix_partial + kb, FIX_KB);
            __threadfence();
            __syncthreads();
            if (tid == 0) chunk_last_s = (atomicAdd(fix_tickets + chunk, 1u) == (unsigned int)(KB - 1)) ? 1u : 0u;
            __syncthreads();
            if (!chunk_last_s) continue;                        // (block-uniform)
            __threadfence();
            if (tid < AS_BM && n0 + tid < total) {              // merge the code blocks' partial arg-mins of my row
                const volatile unsigned long long* pr = fix_partial + (size_t)(n0 + tid) * FIX_KB;
                unsigned long long best = pr[0];
                for (int b = 1; b < KB; ++b) { const unsigned long long v = pr[b]; best = v < best ? v : best; }
                embed_ind[row_list[n0 + tid]] = (int64_t)(unsigned int)(best & 0xffffffffull);
            }
            if (tid == 0) fix_tickets[chunk] = 0u;              // clean for the next call
        } else {
            assign_chunk(x, L, D, K, cbT, ee, embed_ind, row_list, total, n0);
        }
        if (want_gather) {
            __syncthreads();                                    // the chunk's indices are written (same block reads them)
            for (int64_t g0 = n0; g0 < min(total, n0 + AS_BM); g0 += GS_BM)
                gather_chunk(x, L, D, cbT, embed_ind, quantize, stat_sums, stat_counts, row_list, total, g0, gs_tile, acc);
        }
    }
    if (want_gather && diff_acc) {
        acc = warp_sum(acc);
        if (lane == 0) warp_part[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            float s = 0.f;
            for (int w = 0; w < AS_THREADS / 32; ++w) s += warp_part[w];
            if (s != 0.f) atomicAdd(diff_acc, (double)s);
        }
    }
    __threadfence();
    if (tid == 0) last_s = (atomicAdd(ticket, 1u) == gridDim.x - 1u) ? 1u : 0u;
    __syncthreads();
    if (last_s && tid == 0) {
        *ticket = 0u;
        if (diff) diff[0] = (float)(*reinterpret_cast<volatile double*>(diff_acc) * inv_count);
    }
}


// diagnostics: flag ping-pong between two GPUs over peer-mapped memory (tools/p2p_latency.py).  The initiator stores i to the
// peer's flag and waits for the echo in its own; the responder echoes.  *ns_out = globaltimer nanoseconds for `iters` round trips.
__global__ void k_pingpong(unsigned int* my_flag, unsigned int* peer_flag, int iters, int initiator, int with_fence,
                           unsigned long long* ns_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int i = 1; i <= iters; ++i) {
        if (initiator) {
            if (with_fence) __threadfence_system();
            st_release_sys(peer_flag, (unsigned int)i);
            while ((int)(ld_acquire_sys(my_flag) - (unsigned int)i) < 0) {}
        } else {
            while ((int)(ld_acquire_sys(my_flag) - (unsigned int)i) < 0) {}
            if (with_fence) __threadfence_system();
            st_release_sys(peer_flag, (unsigned int)i);
        }
    }
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    *ns_out = t1 - t0;
}

}  // namespace vqb200

namespace vqb200 {

// ------------------------------------------------------------------------------------------------
// Fold + [exchange] + EMA in ONE launch (vqvae.py:55-70): replaces k_stats_fold + [all-reduce] + k_ema64.
// grid = K / 4 blocks of 1024 threads; block b owns codes 4b .. 4b+3.
//   fold      the block sums ITS 256 statistics elements over the per-CTA tables of the statistics kernel (4 thread groups,
//             each over every 4th table with all its loads in flight, combined in a fixed order -> deterministic); the
//             rows-per-code counts come from the integer atomics of the statistics kernel (ForwardScratch::code_counts);
//   exchange  (P2P) flag-in-data protocol over peer memory (what NCCL calls LL): every statistics word travels as ONE 8-byte
//             store {value, step} into slot `rank` of every rank's receive buffer -- 8-byte stores are single transactions,
//             so a word whose tag equals `step` is valid and NO fence, flag array or acknowledgement round trip is needed
//             (a release flag behind the data cost a full NVLink round trip, 4.9 us measured, before the 2.5 us one-way
//             trip of the flag itself).  Every thread then polls exactly the words it needs in its LOCAL slots (its element
//             of every rank; every block needs all K counts for n = sum cluster_size) and adds them in RANK ORDER ->
//             identical bits on every rank, no remote load, no NCCL call;
//   EMA       as k_ema64: decay, Laplace-smoothed renormalisation, embed / embed_avg in place, next codebook image,
//             cluster_size stored by the last block (every block derives n from the OLD values).  The old embed_avg /
//             cluster_size values are loaded BEFORE the fold so that their latency hides behind it.
// A word that does not arrive within P2P_TIMEOUT_NS: (step, rank) recorded in peers.err, the host raises.
// ------------------------------------------------------------------------------------------------
constexpr int EF_CODES = 4, EF_THREADS = 1024;
struct PeerFold {
    uint2* push_dst[2][P2P_MAX_RANKS];         // [step parity][rank r]: r's receive slot for MY statistics: [K*64 sums | K counts] x {value, step}
    const uint2* recv[2][P2P_MAX_RANKS];       // [step parity][rank]: my local receive slots, one per rank
    unsigned int* err;                         // 2 words: (step, rank) of a time-out
    unsigned int* step_counter;                // local device word: number of exchanges completed so far (the step tag lives on the
                                               // device, so a captured graph can be replayed and the host cannot fall out of step)
    int rank, world;
};

__device__ __forceinline__ void st_ll(uint2* p, float v, unsigned int tag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ uint2 ld_ll(const uint2* p) {
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
// rank-ordered sum of word i of every rank's local slot, each word polled until its tag is `step`
__device__ __forceinline__ float ll_gather(const PeerFold& peers, unsigned int step, size_t i) {
    const int par = (int)(step & 1u);
    uint2 v[P2P_MAX_RANKS];
    unsigned int pending = 0;
#pragma unroll
    for (int r = 0; r < P2P_MAX_RANKS; ++r)
        if (r < peers.world) { v[r] = ld_ll(peers.recv[par][r] + i); pending |= (v[r].y != step) ? (1u << r) : 0u; }
    unsigned long long t0 = 0;
    unsigned int spins = 0;
    while (pending) {
#pragma unroll
        for (int r = 0; r < P2P_MAX_RANKS; ++r)
            if (pending & (1u << r)) { v[r] = ld_ll(peers.recv[par][r] + i); if (v[r].y == step) pending &= ~(1u << r); }
        if (pending && (++spins & 255u) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > P2P_TIMEOUT_NS) {
                peers.err[1] = (unsigned int)(__ffs(pending) - 1);
                atomicExch(peers.err, step);
                break;
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < P2P_MAX_RANKS; ++r) s += (r < peers.world) ? __uint_as_float(v[r].x) : 0.f;      // + 0.f for absent ranks: exact
    return s;
}

template <bool P2P>
__global__ void __launch_bounds__(EF_THREADS) k_ema64f(const float* __restrict__ partials, const unsigned int* __restrict__ n_parts_ptr,
                                                        const int* __restrict__ code_counts, PeerFold peers,
                                                        float* __restrict__ cluster_size, float* __restrict__ embed_avg,
                                                        float* __restrict__ embed, float* __restrict__ cbT, float* __restrict__ ee,
                                                        unsigned char* __restrict__ img, int K, float decay, float one_minus_decay,
                                                        float eps, float cA, float cA1, float cB, unsigned int* __restrict__ ticket) {
    __shared__ float red[4][EF_CODES * 64];
    __shared__ float es[EF_CODES][65];
    __shared__ float ssum[EF_CODES][65];
    __shared__ float cnt_s[512];                                // K <= 512 (tc_shape_ok)
    __shared__ float csn_s[512];                                // new cluster sizes
    __shared__ float e2s[EF_CODES];
    __shared__ float part[32];
    __shared__ float n_s;
    __shared__ unsigned int last_s;
    pdl_wait();
    pdl_trigger();
    const int k0 = blockIdx.x * EF_CODES, tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int n_parts = (int)*n_parts_ptr, nstat = K * 65;
    // exchange number of this launch (1, 2, ...): every block reads the counter before the last block (ticket) advances it
    unsigned int step = 0;
    if constexpr (P2P) step = *reinterpret_cast<volatile unsigned int*>(peers.step_counter) + 1u;
#ifdef VQB200_P2P_TRACE
    unsigned long long tr[5] = {0, 0, 0, 0, 0};
    auto stamp = [&](int i) { if (P2P && tid == 0 && blockIdx.x == gridDim.x - 1) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr[i])); };
    stamp(0);
#endif
    // ---- old values first: nothing below depends on them until the EMA, so their latency hides behind the fold / exchange
    float old_avg = 0.f, old_cs = 0.f;
    if (tid < 64 * EF_CODES) old_avg = embed_avg[(size_t)(tid >> 2) * K + k0 + (tid & 3)];     // 16-byte segments of 4 consecutive codes per dim
    if (tid < K) old_cs = cluster_size[tid];
    // ---- fold: element e of my 256 (code e / 64, dim e % 64), thread group g over the tables g, g + 4, ... (<= 40 each)
    {
        const int e = tid & 255, g = tid >> 8;
        const float* src = partials + (size_t)k0 * 64 + e + (size_t)g * nstat;
        float s = 0.f;
        for (int c0 = 0; c0 < n_parts; c0 += 80) {                // rounds of 20 tables per group, all 20 loads in flight
            float v[20];
#pragma unroll
            for (int u = 0; u < 20; ++u) {
                const int c = c0 + g + 4 * u;
                v[u] = c < n_parts ? __ldcs(src + (size_t)(c0 + 4 * u) * nstat) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 20; ++u) s += v[u];
        }
        red[g][e] = s;
    }
    __syncthreads();
    float tot = 0.f;
    if (tid < 256) tot = ((red[0][tid] + red[1][tid]) + red[2][tid]) + red[3][tid];
#ifdef VQB200_P2P_TRACE
    stamp(1);
#endif
    if constexpr (P2P) {
        // ---- every word to every rank (mine included) as {value, step}; then poll the words this thread needs
        if (tid < 256) {
#pragma unroll
            for (int r = 0; r < P2P_MAX_RANKS; ++r)
                if (r < peers.world) st_ll(peers.push_dst[step & 1u][r] + (size_t)k0 * 64 + tid, tot, step);
        } else if (tid < 256 + EF_CODES) {
            const float c = (float)code_counts[k0 + tid - 256];
#pragma unroll
            for (int r = 0; r < P2P_MAX_RANKS; ++r)
                if (r < peers.world) st_ll(peers.push_dst[step & 1u][r] + (size_t)K * 64 + k0 + tid - 256, c, step);
        }
#ifdef VQB200_P2P_TRACE
        stamp(2);
#endif
        if (tid < 256) tot = ll_gather(peers, step, (size_t)k0 * 64 + tid);
        if (tid >= 512 && tid - 512 < K) cnt_s[tid - 512] = ll_gather(peers, step, (size_t)K * 64 + tid - 512);
#ifdef VQB200_P2P_TRACE
        __syncthreads();
        stamp(3);
#endif
    } else {
        if (tid >= 512 && tid - 512 < K) cnt_s[tid - 512] = (float)code_counts[tid - 512];
    }
    if (tid < 256) ssum[tid >> 6][tid & 63] = tot;
    __syncthreads();
    // ---- n = sum_k cluster_size_new[k] from the OLD cluster sizes (vqvae.py:61-65)
    float cs_new = 0.f;
    if (tid < K) { cs_new = __fmaf_rn(cnt_s[tid], one_minus_decay, __fmul_rn(old_cs, decay)); csn_s[tid] = cs_new; }
    float s = warp_sum(cs_new);
    if (lane == 0) part[w] = s;
    __syncthreads();
    if (tid == 0) {
        float n = 0.f;
        for (int i = 0; i < (K + 31) / 32; ++i) n += part[i];
        n_s = n;
    }
    __syncthreads();
    const float n = n_s;
    const float denom = n + (float)((double)K * (double)eps);
    if (tid < 64 * EF_CODES) {                                   // 16-byte segments of 4 consecutive codes per dim
        const int d = tid >> 2, j = tid & 3, k = k0 + j;
        const float cs = (csn_s[k] + eps) / denom * n;          // vqvae.py:66-68
        const size_t o = (size_t)d * K + k;
        const float a = __fmaf_rn(ssum[j][d], one_minus_decay, __fmul_rn(old_avg, decay));       // vqvae.py:64
        embed_avg[o] = a;
        const float e = a / cs;                                 // vqvae.py:69-70
        embed[o] = e;
        es[j][d] = e;
    }
    __syncthreads();
    if (w < EF_CODES) {
        const float v0 = es[w][lane], v1 = es[w][lane + 32];
        if (cbT) {
            cbT[(size_t)(k0 + w) * 64 + lane] = v0;
            cbT[(size_t)(k0 + w) * 64 + lane + 32] = v1;
        }
        float s2 = fmaf(v0, v0, 0.f);
        s2 = fmaf(v1, v1, s2);
        s2 = warp_sum(s2);
        if (lane == 0) { if (ee) ee[k0 + w] = s2; e2s[w] = s2; }
    }
    __syncthreads();
    if (img && tid < 8 * EF_CODES)
        tc::tc_image_rows(&es[tid >> 3][0], e2s[tid >> 3], img, K, k0 + (tid >> 3), tid & 7, cA, cA1, cB);
    // ---- deferred cluster_size store: only after every block has read the old values
    __threadfence();
    if (tid == 0) last_s = (atomicAdd(ticket, 1u) == gridDim.x - 1u) ? 1u : 0u;
    __syncthreads();
    if (last_s) {
        if (tid < K) cluster_size[tid] = cs_new;
        if (tid == 0) {
            *ticket = 0u;                                       // clean for the next launch
            if constexpr (P2P) *peers.step_counter = step;
        }
    }
#ifdef VQB200_P2P_TRACE
    if constexpr (P2P) {
        stamp(4);
        if (tid == 0 && blockIdx.x == gridDim.x - 1) {          // (kernel start, fold, push, wait, rest) of the last block, ns
            unsigned int* t = peers.err + 4;
            t[0] = (unsigned int)(tr[0] & 0xffffffffu); t[1] = (unsigned int)(tr[1] - tr[0]); t[2] = (unsigned int)(tr[2] - tr[1]);
            t[3] = (unsigned int)(tr[3] - tr[2]); t[4] = (unsigned int)(tr[4] - tr[3]); t[5] = step;
        }
    }
#endif
}

// The same flag-in-data exchange for ANY packed statistics buffer (shapes outside the fused fold + EMA kernel: D = 128 / 256,
// K >= 1024): every thread stores its words of d_stats as {value, step} pairs into every rank's receive slot, then polls
// the LOCAL slots and overwrites d_stats with the rank-ordered sum -- an in-place all-reduce without NCCL, bit-identical on
// every rank.  One wave of CTAs (grid <= SM count): every CTA pushes ALL its words before it polls any, so no CTA waits
// for a word whose sender cannot run.  peers.step_counter[1] is the launch's ticket word.
__global__ void __launch_bounds__(1024, 1) k_exchange_ll(float* __restrict__ stats, int n, PeerFold peers) {
    __shared__ unsigned int last_s;
    pdl_wait();
    pdl_trigger();
    const unsigned int step = *reinterpret_cast<volatile unsigned int*>(peers.step_counter) + 1u;
    const int par = (int)(step & 1u);
    const int stride = (int)(gridDim.x * blockDim.x);
    const int first = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    for (int i = first; i < n; i += stride) {
        const float v = stats[i];
#pragma unroll
        for (int r = 0; r < P2P_MAX_RANKS; ++r)
            if (r < peers.world) st_ll(peers.push_dst[par][r] + i, v, step);
    }
    for (int i = first; i < n; i += stride) stats[i] = ll_gather(peers, step, (size_t)i);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last_s = (atomicAdd(peers.step_counter + 1, 1u) == gridDim.x - 1u) ? 1u : 0u;
        if (last_s) {
            peers.step_counter[1] = 0u;
            *peers.step_counter = step;
        }
    }
}

}  // namespace vqb200
