// Exact fp32 SIMT kernels of the quantizer path.
//
// These pin the semantics (Appendix A of SURVEY.md, reference vqvae.py:42-78) and stay in the product
// as (a) the exact re-score / fallback engine of the tcgen05 path and (b) the engine for shapes the
// tensor-core kernel does not cover.  Streaming kernels are laid out for coalesced 128-byte accesses
// in both accepted row layouts (contiguous rows, or NCHW-physical rows with unit row stride).
#pragma once
#include "common.cuh"

namespace vqb200 {

// ------------------------------------------------------------------------------------------------
// codebook image: transpose embed [D,K] -> cbT [K,D] and ||e_k||^2
// ------------------------------------------------------------------------------------------------
__global__ void k_codebook_transpose(const float* __restrict__ embed, float* __restrict__ cbT, int D, int K) {
    __shared__ float tile[32][33];
    int k0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int d = d0 + j, k = k0 + threadIdx.x;
        tile[j][threadIdx.x] = (d < D && k < K) ? embed[(size_t)d * K + k] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int k = k0 + j, d = d0 + threadIdx.x;
        if (k < K && d < D) cbT[(size_t)k * D + d] = tile[threadIdx.x][j];
    }
}

// one warp per code; fixed summation order (lane-strided partials, xor tree)
__global__ void k_codebook_norms(const float* __restrict__ cbT, float* __restrict__ ee, int D, int K) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= K) return;
    const float* e = cbT + (size_t)warp * D;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s = fmaf(e[d], e[d], s);
    s = warp_sum(s);
    if (lane == 0) ee[warp] = s;
}

// ------------------------------------------------------------------------------------------------
// exact distance + argmin (vqvae.py:44-49): register-tiled fp32 FMA, 64 rows x 64 codes per step
// ------------------------------------------------------------------------------------------------
constexpr int AS_BM = 64, AS_BN = 64, AS_DC = 32, AS_THREADS = 256;

// (distance, code) as ONE unsigned 64-bit key whose integer order is the reference's order: smaller distance first (fp32
// bits mapped to an order-preserving unsigned), lower code index on exact ties -- partial arg-mins of different code blocks
// of one row are merged with a plain integer minimum
__device__ __forceinline__ unsigned long long dist_key(float dist, int k) {
    unsigned int u = __float_as_uint(dist);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned int)k;
}

// exact distance + arg-min (vqvae.py:44-49) of one chunk of AS_BM rows starting at position n0 against the codes [c_lo, c_hi):
// register-tiled fp32 FMA, 64 rows x 64 codes per step.  Rows come either from [0, total) or, when row_list != nullptr,
// from row_list[0 .. total).  The result of row position n0 + r goes to embed_ind[row] (whole codebook) and / or, as a
// (distance, code) key, to part[(n0 + r) * part_stride] (partial arg-min of a code block).
__device__ __forceinline__ void assign_core(const float* __restrict__ x, const RowLayout& L, int D, int K,
                                            const float* __restrict__ cbT, const float* __restrict__ ee,
                                            const int* __restrict__ row_list, int64_t total, int64_t n0, int c_lo, int c_hi,
                                            int64_t* __restrict__ embed_ind, unsigned long long* __restrict__ part, int part_stride) {
    __shared__ __align__(16) float xs[AS_DC][AS_BM + 4];
    __shared__ __align__(16) float es[AS_DC][AS_BN + 4];
    __shared__ int64_t row_off[AS_BM];
    __shared__ int row_id[AS_BM];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    __syncthreads();
    if (tid < AS_BM) {
        int64_t i = n0 + tid;
        int64_t n = -1;
        if (i < total) n = row_list ? (int64_t)row_list[i] : i;
        row_id[tid] = (int)n;
        row_off[tid] = n >= 0 ? row_offset(L, n) : 0;
    }
    __syncthreads();
    float best[4];
    int best_k[4];
    float xx[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) { best[i] = INFINITY; best_k[i] = c_lo; }
    for (int c0 = c_lo; c0 < c_hi; c0 += AS_BN) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int d0 = 0; d0 < D; d0 += AS_DC) {
            for (int i = tid; i < AS_BM * AS_DC; i += AS_THREADS) {
                int r, dd;
                if (L.col_stride == 1) { r = i / AS_DC; dd = i % AS_DC; } else { dd = i / AS_BM; r = i % AS_BM; }
                int d = d0 + dd;
                float v = 0.f;
                if (row_id[r] >= 0 && d < D) v = x[row_off[r] + (int64_t)d * L.col_stride];
                xs[dd][r] = v;
            }
            for (int i = tid; i < AS_BN * AS_DC; i += AS_THREADS) {
                int c = i / AS_DC, dd = i % AS_DC;
                int k = c0 + c, d = d0 + dd;
                es[dd][c] = (k < K && d < D) ? cbT[(size_t)k * D + d] : 0.f;
            }
            __syncthreads();
#pragma unroll 8
            for (int dd = 0; dd < AS_DC; ++dd) {
                float4 xv = *reinterpret_cast<const float4*>(&xs[dd][ty * 4]);
                float4 ev = *reinterpret_cast<const float4*>(&es[dd][tx * 4]);
                float xr[4] = {xv.x, xv.y, xv.z, xv.w};
                float er[4] = {ev.x, ev.y, ev.z, ev.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (c0 == c_lo) xx[i] = fmaf(xr[i], xr[i], xx[i]);      // same summation order as assign_chunk
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xr[i], er[j], acc[i][j]);
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = c0 + tx * 4 + j;
            if (k < K && k < c_hi) {
                float e2 = ee[k];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float dist = (xx[i] - 2.f * acc[i][j]) + e2;   // vqvae.py:44-48, left to right
                    if (dist < best[i]) { best[i] = dist; best_k[i] = k; }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float b = best[i];
        int bk = best_k[i];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            float ob = __shfl_xor_sync(0xffffffffu, b, o);
            int obk = __shfl_xor_sync(0xffffffffu, bk, o);
            if (ob < b || (ob == b && obk < bk)) { b = ob; bk = obk; }
        }
        int r = ty * 4 + i;
        if (tx == 0 && row_id[r] >= 0) {
            if (embed_ind) embed_ind[row_id[r]] = (int64_t)bk;
            if (part) part[(size_t)(n0 + r) * part_stride] = dist_key(b, bk);
        }
    }
}
__device__ __forceinline__ void assign_chunk(const float* __restrict__ x, const RowLayout& L, int D, int K,
                                             const float* __restrict__ cbT, const float* __restrict__ ee,
                                             int64_t* __restrict__ embed_ind, const int* __restrict__ row_list,
                                             int64_t total, int64_t n0) {
    assign_core(x, L, D, K, cbT, ee, row_list, total, n0, 0, K, embed_ind, nullptr, 0);
}
__device__ __forceinline__ void assign_part(const float* __restrict__ x, const RowLayout& L, int D, int K,
                                            const float* __restrict__ cbT, const float* __restrict__ ee,
                                            const int* __restrict__ row_list, int64_t total, int64_t n0, int c_lo, int c_hi,
                                            unsigned long long* __restrict__ part, int part_stride) {
    assign_core(x, L, D, K, cbT, ee, row_list, total, n0, c_lo, c_hi, nullptr, part, part_stride);
}

__global__ void __launch_bounds__(AS_THREADS)
k_assign_exact(const float* __restrict__ x, RowLayout L, int D, int K,
               const float* __restrict__ cbT, const float* __restrict__ ee,
               int64_t* __restrict__ embed_ind,
               const int* __restrict__ row_list, const int* __restrict__ row_count) {
    const int64_t total = row_list ? (int64_t)(*row_count) : L.n_rows;
    for (int64_t n0 = (int64_t)blockIdx.x * AS_BM; n0 < total; n0 += (int64_t)gridDim.x * AS_BM)
        assign_chunk(x, L, D, K, cbT, ee, embed_ind, row_list, total, n0);
}

// ------------------------------------------------------------------------------------------------
// gather + straight-through output + commitment loss + code statistics, one pass over x
// (vqvae.py:52, 72-73 and the statistics of :50,55-56 without a one-hot)
// ------------------------------------------------------------------------------------------------
constexpr int GS_BM = 32, GS_THREADS = 256;

// one chunk of GS_BM rows starting at position n0 (tile = dynamic shared memory, GS_BM x (D+1) floats)
__device__ __forceinline__ void gather_chunk(const float* __restrict__ x, const RowLayout& L, int D,
                                             const float* __restrict__ cbT, const int64_t* __restrict__ embed_ind,
                                             float* __restrict__ quantize, float* __restrict__ stat_sums,
                                             float* __restrict__ stat_counts, const int* __restrict__ row_list,
                                             int64_t total, int64_t n0, float* tile, float& acc) {
    __shared__ int64_t row_off[GS_BM];
    __shared__ int code[GS_BM];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld = D + 1;
    const int rows = (int)min((int64_t)GS_BM, total - n0);
    __syncthreads();                              // previous chunk fully written out
    if (tid < GS_BM) {
        bool ok = tid < rows;
        int64_t n = ok ? (row_list ? (int64_t)row_list[n0 + tid] : n0 + tid) : 0;
        row_off[tid] = ok ? row_offset(L, n) : 0;
        int64_t k = ok ? embed_ind[n] : 0;
        code[tid] = (int)k;
    }
    __syncthreads();
    for (int i = tid; i < GS_BM * D; i += GS_THREADS) {
        int r, d;
        if (L.col_stride == 1) { r = i / D; d = i % D; } else { d = i / GS_BM; r = i % GS_BM; }
        if (r < rows) tile[r * ld + d] = x[row_off[r] + (int64_t)d * L.col_stride];
    }
    __syncthreads();
    for (int r = warp; r < rows; r += GS_THREADS / 32) {
        const int k = code[r];
        const float* e = cbT + (size_t)k * D;
        for (int d = lane; d < D; d += 32) {
            float xv = tile[r * ld + d];
            float dl = e[d] - xv;                 // (quantize - input), vqvae.py:72-73
            acc = fmaf(dl, dl, acc);
            tile[r * ld + d] = xv + dl;           // input + (quantize - input).detach()
            if (stat_sums) atomicAdd(&stat_sums[(size_t)k * D + d], xv);
        }
        if (stat_counts && lane == 0) atomicAdd(&stat_counts[k], 1.0f);
    }
    __syncthreads();
    if (quantize) {
        for (int i = tid; i < GS_BM * D; i += GS_THREADS) {
            int r, d;
            if (L.col_stride == 1) { r = i / D; d = i % D; } else { d = i / GS_BM; r = i % GS_BM; }
            if (r < rows) quantize[row_off[r] + (int64_t)d * L.col_stride] = tile[r * ld + d];
        }
    }
}

// diff (may be null) is finalised by the last block to finish: diff = diff_acc * inv_count  (vqvae.py:72 mean)
__global__ void __launch_bounds__(GS_THREADS)
k_gather_stats(const float* __restrict__ x, RowLayout L, int D, int K,
               const float* __restrict__ cbT, const int64_t* __restrict__ embed_ind,
               float* __restrict__ quantize, double* __restrict__ diff_acc,
               float* __restrict__ stat_sums, float* __restrict__ stat_counts,
               const int* __restrict__ row_list, const int* __restrict__ row_count,
               float* __restrict__ diff, double inv_count, unsigned int* __restrict__ ticket) {
    extern __shared__ float tile[];              // [GS_BM][D + 1]
    __shared__ float warp_part[GS_THREADS / 32];
    __shared__ unsigned int last_s;
    (void)K;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t total = row_list ? (int64_t)(*row_count) : L.n_rows;
    float acc = 0.f;
    for (int64_t n0 = (int64_t)blockIdx.x * GS_BM; n0 < total; n0 += (int64_t)gridDim.x * GS_BM)
        gather_chunk(x, L, D, cbT, embed_ind, quantize, stat_sums, stat_counts, row_list, total, n0, tile, acc);
    acc = warp_sum(acc);
    if (lane == 0) warp_part[warp] = acc;
    __syncthreads();
    if (tid == 0 && diff_acc) {
        float s = 0.f;
        for (int w = 0; w < GS_THREADS / 32; ++w) s += warp_part[w];
        atomicAdd(diff_acc, (double)s);
    }
    if (ticket) {
        __threadfence();
        if (tid == 0) last_s = (atomicAdd(ticket, 1u) == gridDim.x - 1u) ? 1u : 0u;
        __syncthreads();
        if (last_s && tid == 0) {
            *ticket = 0u;
            if (diff && diff_acc) diff[0] = (float)(*reinterpret_cast<volatile double*>(diff_acc) * inv_count);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// code statistics without a one-hot and without floating-point atomics (vqvae.py:50,55-56):
// a segmented reduction keyed by code index.  Each CTA owns a slice of rows and a PRIVATE [K][D] fp32
// table in shared memory; rows are bucketed by code with a counting sort (integer smem atomics only),
// every code bucket is summed by exactly one warp (exclusive ownership -> plain adds), and the per-CTA
// tables are written out and folded by k_stats_reduce in a fixed order.
// (Global fp32 atomics top out near 170 G adds/s on B200: N*D = 33.5 M adds cost ~200 us at cfg-2.)
// ------------------------------------------------------------------------------------------------
constexpr int CS_THREADS = 1024, CS_CHUNK = 4096, CS_MAX_CTAS = 160;

// float add on shared memory (compiles to a CAS loop): only for the <= 2 buckets per warp and chunk that straddle the
// boundary between two warps' ranges
__device__ __forceinline__ void smem_add(float* p, float v) { atomicAdd(p, v); }

__global__ void __launch_bounds__(CS_THREADS, 1)
k_code_stats(const float* __restrict__ x, RowLayout L, int D, int K, const int64_t* __restrict__ embed_ind,
             float* __restrict__ partials /* [gridDim.x][K*(D+1)] */, int chunk /* rows per trip, <= CS_CHUNK */,
             int* __restrict__ code_counts /* may be null: [K] += rows per code (integer atomics: exact, order-free) */,
             unsigned int* __restrict__ n_parts_out /* may be null: receives gridDim.x */) {
    // layout: the int64 array first (dynamic shared memory is 16-byte aligned; behind [K][D] floats it would only be
    // 4-byte aligned when K*D is odd, e.g. Quantize(3, 5): misaligned 8-byte shared stores), then the 4-byte arrays,
    // then the 2-byte arrays
    extern __shared__ __align__(16) unsigned char cs_smem_raw[];
    int64_t* order = reinterpret_cast<int64_t*>(cs_smem_raw);                     // [CS_CHUNK] element offsets of the rows, bucketed by code
    float* table = reinterpret_cast<float*>(order + CS_CHUNK);                   // [K][D]
    int* cnt_total = reinterpret_cast<int*>(table + (size_t)K * D);              // [K]
    int* hist = cnt_total + K;                                // [K]   bucket sizes of this chunk
    int* start = hist + K;                                    // [K+1] bucket offsets
    int* cursor = start + K + 1;                              // [K]
    unsigned short* code = reinterpret_cast<unsigned short*>(cursor + K);         // [CS_CHUNK] code of row i
    unsigned short* scode = code + CS_CHUNK;                  // [CS_CHUNK] code at sorted position p
    __shared__ int warp_tot[CS_THREADS / 32];
    static_assert(CS_THREADS / 32 == 32, "the scan folds one warp total per lane");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = CS_THREADS / 32;

    pdl_trigger();
    for (int i = tid; i < K * D; i += CS_THREADS) table[i] = 0.f;
    for (int i = tid; i < K; i += CS_THREADS) cnt_total[i] = 0;
    pdl_wait();                                               // (the shared-memory table is cleared while the upstream kernel drains)
    const int64_t n_chunks = (L.n_rows + chunk - 1) / chunk;
    const bool fast = (D == 64 && L.col_stride == 1 && (reinterpret_cast<uintptr_t>(x) & 7u) == 0 && (L.row_stride & 1) == 0 &&
                       (L.image_stride & 1) == 0);
    // walk the chunks from the END of x: those rows were touched last by the assignment kernel and are
    // the most likely to still sit in L2
    for (int64_t j = n_chunks - 1 - blockIdx.x; j >= 0; j -= gridDim.x) {
        const int64_t r0 = j * chunk;
        const int rows = (int)min((int64_t)chunk, L.n_rows - r0);
        // --- bucket the rows of this trip by code (counting sort).  The (at most CS_CHUNK / CS_THREADS = 4) codes of a thread
        // are loaded up front -- independent loads, all in flight together -- and stay in registers for both passes.
        int kk[CS_CHUNK / CS_THREADS];
#pragma unroll
        for (int u = 0; u < CS_CHUNK / CS_THREADS; ++u) {
            const int i = tid + u * CS_THREADS;
            kk[u] = -1;
            if (i < rows) {
                const long long kl = embed_ind[r0 + i];
                int k = (int)kl;
                if (kl < 0 || kl >= K) {         // cannot happen for indices written by our own kernels: clamp for memory safety AND report
                    if (n_parts_out) atomicExch(n_parts_out + 1, 1u);   // scratch header word 14 (byte 56): internal-error flag
                    k = kl < 0 ? 0 : K - 1;
                }
                kk[u] = k;
            }
        }
        __syncthreads();
        for (int i = tid; i < K; i += CS_THREADS) hist[i] = 0;
        __syncthreads();
#pragma unroll
        for (int u = 0; u < CS_CHUNK / CS_THREADS; ++u)
            if (kk[u] >= 0) atomicAdd(&hist[kk[u]], 1);
        __syncthreads();
        {                                                     // exclusive scan of hist -> start, all warps: warp w owns SB bins
            const int SB = ((K + nwarps - 1) / nwarps + 31) / 32 * 32;
            const int b_lo = min(K, warp * SB), b_hi = min(K, b_lo + SB);
            int carry = 0;
            for (int b = b_lo; b < b_hi; b += 32) {
                const int v = (b + lane < b_hi) ? hist[b + lane] : 0;
                int incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                if (b + lane < b_hi) start[b + lane] = carry + incl - v;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) warp_tot[warp] = carry;
            __syncthreads();
            const int t = warp_tot[lane];                     // nwarps == 32: one total per lane
            int incl = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            const int offset = __shfl_sync(0xffffffffu, incl - t, warp);
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            for (int b = b_lo + lane; b < b_hi; b += 32) {
                const int st = start[b] + offset;
                start[b] = st;
                cursor[b] = st;
            }
            if (tid == 0) start[K] = total;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < CS_CHUNK / CS_THREADS; ++u) {
            if (kk[u] >= 0) {
                const int dst = atomicAdd(&cursor[kk[u]], 1);
                order[dst] = row_offset(L, r0 + tid + u * CS_THREADS);
                scode[dst] = (unsigned short)kk[u];
            }
        }
        for (int k = tid; k < K; k += CS_THREADS) cnt_total[k] += hist[k];
        __syncthreads();
        if (fast) {
            // flattened segmented reduction: every warp streams an equal share of the sorted positions with 16 rows
            // (256 B each, lane owns dims 2l, 2l+1) in flight, accumulates while the code stays the same and adds to
            // the table when it changes.  A bucket that lies inside the warp's range is owned exclusively (plain
            // add); the at most two that straddle a range boundary use a shared-memory float add.
            const int per = (rows + nwarps - 1) / nwarps;
            const int my_begin = warp * per, my_end = min(rows, my_begin + per);
            if (my_begin < my_end) {
                int cur = scode[my_begin];
                float a0 = 0.f, a1 = 0.f;
                auto flush = [&](int k) {
                    float* t = table + (size_t)k * 64 + 2 * lane;
                    if (start[k] >= my_begin && start[k + 1] <= my_end) {
                        float2 tv = *reinterpret_cast<float2*>(t);
                        tv.x += a0; tv.y += a1;
                        *reinterpret_cast<float2*>(t) = tv;
                    } else {
                        smem_add(t, a0);
                        smem_add(t + 1, a1);
                    }
                };
                // software pipeline: two register batches of 8 rows; the loads of one batch are issued before the other is
                // consumed, so a warp always has 8 .. 16 rows (2 .. 4 KB) in flight instead of draining to zero per trip
                const float2* xl = reinterpret_cast<const float2*>(x) + lane;
                auto load8 = [&](float2 (&v)[8], int p) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int pp = p + u;
                        v[u] = pp < my_end ? __ldcs(reinterpret_cast<const float2*>(reinterpret_cast<const float*>(xl) + order[pp])) : make_float2(0.f, 0.f);
                    }
                };
                auto eat8 = [&](const float2 (&v)[8], int p) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int cd = p + u < my_end ? (int)scode[p + u] : cur;      // warp-uniform
                        if (cd != cur) {
                            flush(cur);
                            a0 = 0.f; a1 = 0.f;
                            cur = cd;
                        }
                        a0 += v[u].x;
                        a1 += v[u].y;
                    }
                };
                float2 va[8], vb[8];
                load8(va, my_begin);
                for (int p = my_begin; p < my_end; p += 16) {
                    load8(vb, p + 8);
                    eat8(va, p);
                    load8(va, p + 16);
                    eat8(vb, p + 8);
                }
                flush(cur);
            }
        } else {
            // one warp per code bucket: register accumulation, exclusive table update
            for (int k = warp; k < K; k += nwarps) {
                const int b0 = start[k], b1 = start[k + 1];
                if (b0 == b1) continue;
                for (int d0 = 0; d0 < D; d0 += 64) {
                    float a0 = 0.f, a1 = 0.f;
                    const int da = d0 + lane, db = d0 + 32 + lane;
                    int b = b0;
                    for (; b + 4 <= b1; b += 4) {
                        float v0[4], v1[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int64_t off = order[b + u];
                            v0[u] = da < D ? x[off + (int64_t)da * L.col_stride] : 0.f;
                            v1[u] = db < D ? x[off + (int64_t)db * L.col_stride] : 0.f;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) { a0 += v0[u]; a1 += v1[u]; }
                    }
                    for (; b < b1; ++b) {
                        const int64_t off = order[b];
                        if (da < D) a0 += x[off + (int64_t)da * L.col_stride];
                        if (db < D) a1 += x[off + (int64_t)db * L.col_stride];
                    }
                    if (da < D) table[(size_t)k * D + da] += a0;
                    if (db < D) table[(size_t)k * D + db] += a1;
                }
            }
        }
    }
    __syncthreads();
    float* out = partials + (size_t)blockIdx.x * K * (D + 1);
    for (int i = tid; i < K * D; i += CS_THREADS) out[i] = table[i];
    for (int i = tid; i < K; i += CS_THREADS) {
        out[(size_t)K * D + i] = (float)cnt_total[i];
        if (code_counts && cnt_total[i]) atomicAdd(code_counts + i, cnt_total[i]);
    }
    if (n_parts_out && blockIdx.x == 0 && tid == 0) *n_parts_out = gridDim.x;
}

// stats[i] (+)= sum over CTAs of partials[c][i].  block = (32, FOLD_Y): 128 consecutive elements per block as four
// 32-wide column groups; thread-row y sums the tables c = y, y + FOLD_Y, ... (all its loads in flight together), the
// FOLD_Y partial sums are folded in a fixed order -> deterministic for a fixed grid.  accumulate == 0 overwrites stats
// (no memset needed before); the spare words behind the n elements (EMA ticket) are cleared either way.
constexpr int FOLD_Y = 16;
__global__ void __launch_bounds__(32 * FOLD_Y)
k_stats_fold(const float* __restrict__ partials, int n_parts, int n, float* __restrict__ stats, int accumulate) {
    __shared__ float part[FOLD_Y][4][32];
    pdl_wait();
    pdl_trigger();
    const int col = threadIdx.x, y = threadIdx.y;
    const int i0 = blockIdx.x * 128;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = y; c < n_parts; c += 2 * FOLD_Y) {
        const int c2 = c + FOLD_Y;
        float v[8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 32 + col;
            v[u] = i < n ? __ldcs(partials + (size_t)c * n + i) : 0.f;
            v[4 + u] = (i < n && c2 < n_parts) ? __ldcs(partials + (size_t)c2 * n + i) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) s[u] += v[u] + v[4 + u];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) part[y][u][col] = s[u];
    __syncthreads();
    if (y < 4) {
        const int u = y, i = i0 + u * 32 + col;
        if (i < n) {
            float t = 0.f;
#pragma unroll
            for (int r = 0; r < FOLD_Y; ++r) t += part[r][u][col];
            stats[i] = accumulate ? stats[i] + t : t;
        }
    }
    if (blockIdx.x == 0 && y == 0 && col < 4) stats[n + col] = 0.f;
}

__host__ __device__ inline size_t code_stats_smem_bytes(int D, int K) {
    return (size_t)K * D * 4 + (size_t)(4 * K + 2) * 4 + (size_t)CS_CHUNK * 8 + (size_t)CS_CHUNK * 2 * 2 + 16;
}

__global__ void k_finalize_diff(const double* __restrict__ diff_acc, float* __restrict__ diff, double inv_count) {
    if (threadIdx.x == 0 && blockIdx.x == 0) diff[0] = (float)(diff_acc[0] * inv_count);
}

// ------------------------------------------------------------------------------------------------
// EMA update + Laplace-smoothed renormalisation (vqvae.py:61-70), also refreshes cbT / ee
// ------------------------------------------------------------------------------------------------
// single block: cluster_size <- cluster_size*decay + counts*(1-decay);  n = sum(cluster_size)
__global__ void k_ema_cluster(const float* __restrict__ counts, float* __restrict__ cluster_size, int K,
                              float decay, float one_minus_decay, float* __restrict__ n_out) {
    __shared__ float part[32];
    float s = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float c = cluster_size[k] * decay;
        c = c + counts[k] * one_minus_decay;
        cluster_size[k] = c;
        s += c;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) n_out[0] = v;
    }
}

// one warp per code: embed_avg / embed columns, code-major copy and norm
__global__ void k_ema_embed(const float* __restrict__ sums, const float* __restrict__ cluster_size,
                            const float* __restrict__ n_in, float* __restrict__ embed_avg,
                            float* __restrict__ embed, float* __restrict__ cbT, float* __restrict__ ee,
                            int D, int K, float decay, float one_minus_decay, float eps) {
    int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k >= K) return;
    const float n = n_in[0];
    const float denom = n + (float)((double)K * (double)eps);
    const float cs = (cluster_size[k] + eps) / denom * n;      // vqvae.py:66-68
    float s2 = 0.f;
    for (int d = lane; d < D; d += 32) {
        size_t o = (size_t)d * K + k;
        float a = embed_avg[o] * decay;
        a = a + sums[(size_t)k * D + d] * one_minus_decay;     // vqvae.py:64
        embed_avg[o] = a;
        float e = a / cs;                                      // vqvae.py:69-70
        embed[o] = e;
        if (cbT) cbT[(size_t)k * D + d] = e;
        s2 = fmaf(e, e, s2);
    }
    s2 = warp_sum(s2);
    if (lane == 0 && ee) ee[k] = s2;
}

// ------------------------------------------------------------------------------------------------
// re-pack between a strided row layout (e.g. the NCHW-physical permute(0,2,3,1) view VQVAE.encode passes,
// vqvae.py:227,235) and dense [N, D] rows: the tcgen05 engine stages dense 256-byte rows with bulk copies.
// Tiles of RP_ROWS rows go through shared memory so that both sides are coalesced.
// ------------------------------------------------------------------------------------------------
constexpr int RP_ROWS = 64, RP_THREADS = 256;

template <bool PACK>   // PACK: strided -> dense;  !PACK: dense -> strided
__global__ void __launch_bounds__(RP_THREADS)
k_repack_rows(const float* __restrict__ src, float* __restrict__ dst, RowLayout L, int D) {
    extern __shared__ float rp_tile[];           // [RP_ROWS][D + 1]
    __shared__ int64_t roff[RP_ROWS];
    const int tid = threadIdx.x, ld = D + 1;
    for (int64_t n0 = (int64_t)blockIdx.x * RP_ROWS; n0 < L.n_rows; n0 += (int64_t)gridDim.x * RP_ROWS) {
        const int rows = (int)min((int64_t)RP_ROWS, L.n_rows - n0);
        __syncthreads();
        if (tid < RP_ROWS) roff[tid] = tid < rows ? row_offset(L, n0 + tid) : 0;
        __syncthreads();
        if (PACK) {
            for (int i = tid; i < RP_ROWS * D; i += RP_THREADS) {
                int r, d;
                if (L.col_stride == 1) { r = i / D; d = i % D; } else { d = i / RP_ROWS; r = i % RP_ROWS; }
                if (r < rows) rp_tile[r * ld + d] = src[roff[r] + (int64_t)d * L.col_stride];
            }
            __syncthreads();
            for (int i = tid; i < rows * D; i += RP_THREADS) dst[n0 * D + i] = rp_tile[(i / D) * ld + (i % D)];
        } else {
            for (int i = tid; i < rows * D; i += RP_THREADS) rp_tile[(i / D) * ld + (i % D)] = src[n0 * D + i];
            __syncthreads();
            for (int i = tid; i < RP_ROWS * D; i += RP_THREADS) {
                int r, d;
                if (L.col_stride == 1) { r = i / D; d = i % D; } else { d = i / RP_ROWS; r = i % RP_ROWS; }
                if (r < rows) dst[roff[r] + (int64_t)d * L.col_stride] = rp_tile[r * ld + d];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward implied by vqvae.py:72-73
// ------------------------------------------------------------------------------------------------
__global__ void k_backward(const float* __restrict__ x, RowLayout L, int D,
                           const int64_t* __restrict__ embed_ind, const float* __restrict__ cbT,
                           const float* __restrict__ grad_q, const float* __restrict__ grad_diff,
                           float* __restrict__ grad_x, double two_over_count) {
    const int64_t total = L.n_rows * (int64_t)D;
    const float c = grad_diff ? (float)(two_over_count * (double)grad_diff[0]) : 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t n;
        int d;
        if (L.col_stride == 1) { n = i / D; d = (int)(i - n * D); }
        else {  // walk the physical NCHW order: (image, d, row-in-image)
            int64_t per_img = L.rows_per_image * D;
            int64_t img = i / per_img, rem = i - img * per_img;
            d = (int)(rem / L.rows_per_image);
            n = img * L.rows_per_image + (rem - (int64_t)d * L.rows_per_image);
        }
        int64_t off = row_offset(L, n) + (int64_t)d * L.col_stride;
        float g = grad_q ? grad_q[off] : 0.f;
        if (c != 0.f) {
            float q = cbT[(size_t)embed_ind[n] * D + d];
            g = fmaf(c, x[off] - q, g);
        }
        grad_x[off] = g;
    }
}

// dense rows, D % 4 == 0, 16-byte aligned pointers: 128-bit accesses, one 64-bit division per thread and trip
// (grad = grad_q + c (x - e[ind]); three streams of N*D*4 bytes -> HBM-bound)
__global__ void __launch_bounds__(256)
k_backward_dense(const float4* __restrict__ x, int64_t n_vec /* N*D/4 */, int vec_per_row, const int64_t* __restrict__ embed_ind,
                 const float4* __restrict__ cbT, const float4* __restrict__ grad_q, const float* __restrict__ grad_diff,
                 float4* __restrict__ grad_x, double two_over_count) {
    const float c = grad_diff ? (float)(two_over_count * (double)grad_diff[0]) : 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
        float4 g = grad_q ? __ldcs(grad_q + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (c != 0.f) {
            const int64_t n = i / vec_per_row;
            const int v = (int)(i - n * vec_per_row);
            const float4 xv = __ldcs(x + i);
            const float4 q = __ldg(cbT + (size_t)embed_ind[n] * vec_per_row + v);
            g.x = fmaf(c, xv.x - q.x, g.x); g.y = fmaf(c, xv.y - q.y, g.y);
            g.z = fmaf(c, xv.z - q.z, g.z); g.w = fmaf(c, xv.w - q.w, g.w);
        }
        __stcs(grad_x + i, g);
    }
}

// NCHW-physical rows (unit row stride; rows_per_image, col_stride, image_stride multiples of 4; 16-byte aligned): a
// thread owns 4 consecutive rows and walks a chunk of dims, so x / grad_q / grad_x move as row-coalesced float4 and
// the four code rows are read sequentially (L1-resident sectors).  grid = (ceil(N / 1024), dim chunks).
__global__ void __launch_bounds__(256)
k_backward_nchw(const float* __restrict__ x, RowLayout L, int D, int dims_per_block, const int64_t* __restrict__ embed_ind,
                const float* __restrict__ cbT, const float* __restrict__ grad_q, const float* __restrict__ grad_diff,
                float* __restrict__ grad_x, double two_over_count) {
    const float c = grad_diff ? (float)(two_over_count * (double)grad_diff[0]) : 0.f;
    const int64_t n = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (n >= L.n_rows) return;
    const int64_t img = n / L.rows_per_image, r = n - img * L.rows_per_image;
    const int64_t base = img * L.image_stride + r;
    const int d0 = blockIdx.y * dims_per_block, d1 = min(D, d0 + dims_per_block);
    int64_t k[4] = {0, 0, 0, 0};
    if (c != 0.f) {
#pragma unroll
        for (int j = 0; j < 4; ++j) k[j] = embed_ind[n + j] * D;
    }
    if (c != 0.f && (D & 7) == 0 && (d0 & 7) == 0 && ((d1 - d0) & 7) == 0 && (reinterpret_cast<uintptr_t>(cbT) & 31u) == 0) {
        // eight dims of each of the four code rows per 256-bit load: 4 L1TEX sector accesses per 8 dims instead of 32
        for (int d = d0; d < d1; d += 8) {
            float q[4][8];
#pragma unroll
            for (int j = 0; j < 4; ++j) ldg_nc_v8(cbT + k[j] + d, q[j]);
#pragma unroll
            for (int dd = 0; dd < 8; ++dd) {
                const int64_t off = base + (int64_t)(d + dd) * L.col_stride;
                float4 g = grad_q ? __ldcs(reinterpret_cast<const float4*>(grad_q + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 xv = __ldcs(reinterpret_cast<const float4*>(x + off));
                g.x = fmaf(c, xv.x - q[0][dd], g.x);
                g.y = fmaf(c, xv.y - q[1][dd], g.y);
                g.z = fmaf(c, xv.z - q[2][dd], g.z);
                g.w = fmaf(c, xv.w - q[3][dd], g.w);
                __stcs(reinterpret_cast<float4*>(grad_x + off), g);
            }
        }
        return;
    }
#pragma unroll 4
    for (int d = d0; d < d1; ++d) {
        const int64_t off = base + (int64_t)d * L.col_stride;
        float4 g = grad_q ? __ldcs(reinterpret_cast<const float4*>(grad_q + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (c != 0.f) {
            const float4 xv = __ldcs(reinterpret_cast<const float4*>(x + off));
            g.x = fmaf(c, xv.x - __ldg(cbT + k[0] + d), g.x);
            g.y = fmaf(c, xv.y - __ldg(cbT + k[1] + d), g.y);
            g.z = fmaf(c, xv.z - __ldg(cbT + k[2] + d), g.z);
            g.w = fmaf(c, xv.w - __ldg(cbT + k[3] + d), g.w);
        }
        __stcs(reinterpret_cast<float4*>(grad_x + off), g);
    }
}

// ------------------------------------------------------------------------------------------------
// embed_code (vqvae.py:77-78): contiguous [N, D] gather
// ------------------------------------------------------------------------------------------------
__global__ void k_embed_code(const int64_t* __restrict__ ids, int64_t n_rows, const float* __restrict__ cbT,
                             int D, int K, float* __restrict__ out, int* __restrict__ status) {
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    int64_t k = ids[row];
    if (k < 0 || k >= K) {
        if (lane == 0 && status) atomicExch(status, 1);
        return;
    }
    const float* e = cbT + (size_t)k * D;
    float* o = out + row * D;
    for (int d = lane; d < D; d += 32) o[d] = e[d];
}

// ------------------------------------------------------------------------------------------------
// index egress (extract_code.py:23-33 copies int64 indices D2H per batch): n_embed <= 65536 fits 16 bits, so the codes
// leave the device as uint16 (or int32): 4x (2x) fewer bytes over PCIe.  Thread = 4 indices: 32 B in, 8 / 16 B out.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_pack_indices(const int64_t* __restrict__ ids, int64_t n, int K, T* __restrict__ out,
                                                      int* __restrict__ status) {
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= n) return;
    int64_t v[4];
    if (i0 + 4 <= n) {
        const longlong2 a = __ldcs(reinterpret_cast<const longlong2*>(ids + i0));
        const longlong2 b = __ldcs(reinterpret_cast<const longlong2*>(ids + i0) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = i0 + j < n ? ids[i0 + j] : 0;
    }
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) bad |= (v[j] < 0 || v[j] >= K);
    if (bad && status) atomicExch(status, 1);
    if (i0 + 4 <= n) {
        if (sizeof(T) == 2) *reinterpret_cast<ushort4*>(out + i0) = make_ushort4((unsigned short)v[0], (unsigned short)v[1], (unsigned short)v[2], (unsigned short)v[3]);
        else *reinterpret_cast<int4*>(out + i0) = make_int4((int)v[0], (int)v[1], (int)v[2], (int)v[3]);
    } else {
        for (int j = 0; j < 4 && i0 + j < n; ++j) out[i0 + j] = (T)v[j];
    }
}
template <typename T>
__global__ void __launch_bounds__(256) k_unpack_indices(const T* __restrict__ codes, int64_t n, int64_t* __restrict__ ids) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ids[i] = (int64_t)codes[i];
}

}  // namespace vqb200
