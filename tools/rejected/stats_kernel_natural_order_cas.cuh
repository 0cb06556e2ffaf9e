// REJECTED EXPERIMENT (round 2), kept for the record -- not compiled into libvqb200.so.
// Measured on B200, cfg-2 (tools/run_stats_nat.sh at the time): statistics kernels 74.7 us against 37.4 us for the counting-sort
// kernel on dense rows, independent of code skew (512 / 40 / 3 live codes: 75.4 / 74.9 / 76.3 us) -> bound by the CAS loop
// itself (~20 cycles per warp-wide shared-memory float add per SM), not by contention.  NCHW-physical rows read in place
// (no dense side copy written by the main kernel, which saves 26 us there): step 169.4 us against 156.7 us.
// EXPERIMENT (VQB200_STATS_NAT=1): the same statistics in NATURAL row order -- no sort, perfectly streaming loads -- with
// shared-memory float adds (CAS loops, `ATOMS.CAST.SPIN`) into the private table.  D = 64.  Same output contract as
// k_code_stats ([K*64 sums | K counts] per CTA, rows-per-code counters, number of parts).
//   NCHW == false: rows dense; a warp owns a row, lane l adds dims l and l + 32 (conflict-free banks); consecutive rows of
//                  a batch with the same code are pre-added in registers
//   NCHW == true : rows are pixels of an NCHW-physical tensor (row_stride 1, col_stride = pixels per image): a warp owns 32
//                  consecutive pixels, lane = pixel (128-byte coalesced loads per dim), table rows padded to 65 floats so
//                  that lanes with different codes fall into different banks
template <bool NCHW>
__global__ void __launch_bounds__(CS_THREADS, 1)
k_code_stats_nat(const float* __restrict__ x, RowLayout L, int K, const int64_t* __restrict__ embed_ind,
                 float* __restrict__ partials, int* __restrict__ code_counts, unsigned int* __restrict__ n_parts_out) {
    extern __shared__ __align__(16) unsigned char cs_smem_raw[];
    constexpr int TS = NCHW ? 65 : 64;
    float* table = reinterpret_cast<float*>(cs_smem_raw);            // [K][TS]
    int* cnt = reinterpret_cast<int*>(table + (size_t)K * TS);       // [K]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = CS_THREADS / 32;
    pdl_trigger();
    for (int i = tid; i < K * TS; i += CS_THREADS) table[i] = 0.f;
    for (int i = tid; i < K; i += CS_THREADS) cnt[i] = 0;
    __syncthreads();
    pdl_wait();
    const int64_t groups = (L.n_rows + 31) / 32;                     // units of 32 rows
    const int64_t per = (groups + gridDim.x - 1) / gridDim.x;
    const int64_t g_begin = (int64_t)blockIdx.x * per, g_end = min(groups, g_begin + per);
    auto code_of = [&](int64_t row) -> int {
        const long long kl = embed_ind[row];
        int k = (int)kl;
        if (kl < 0 || kl >= K) {
            if (n_parts_out) atomicExch(n_parts_out + 1, 1u);
            k = kl < 0 ? 0 : K - 1;
        }
        return k;
    };
    if constexpr (!NCHW) {
        // warp w takes groups g_end-1-w, -nwarps, ... (from the end: those rows were written last by the assignment kernel)
        for (int64_t g = g_end - 1 - warp; g >= g_begin; g -= nwarps) {
            const int64_t r0 = g * 32;
            const int n = (int)min((int64_t)32, L.n_rows - r0);
            const int my_k = lane < n ? code_of(r0 + lane) : 0;
            if (lane < n) atomicAdd(&cnt[my_k], 1);
            for (int b = 0; b < n; b += 8) {
                float v0[8], v1[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const bool in = b + u < n;
                    const float* xr = x + row_offset(L, r0 + (in ? b + u : 0));
                    v0[u] = in ? __ldcs(xr + lane) : 0.f;
                    v1[u] = in ? __ldcs(xr + 32 + lane) : 0.f;
                }
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (b + u < n) {                                               // warp-uniform
                        const int k = __shfl_sync(0xffffffffu, my_k, b + u);
                        const int kn = (u < 7 && b + u + 1 < n) ? __shfl_sync(0xffffffffu, my_k, b + u + 1) : -1;
                        a0 += v0[u];
                        a1 += v1[u];
                        if (kn != k) {
                            atomicAdd(table + (size_t)k * 64 + lane, a0);
                            atomicAdd(table + (size_t)k * 64 + 32 + lane, a1);
                            a0 = 0.f; a1 = 0.f;
                        }
                    }
                }
            }
        }
    } else {
        for (int64_t g = g_end - 1 - warp; g >= g_begin; g -= nwarps) {
            const int64_t row = g * 32 + lane;
            const bool in = row < L.n_rows;
            const int k = in ? code_of(row) : 0;
            if (in) atomicAdd(&cnt[k], 1);
            const float* xr = x + (in ? row_offset(L, row) : 0);
            float* t = table + (size_t)k * 65;
#pragma unroll 1
            for (int d0 = 0; d0 < 64; d0 += 16) {
                float v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) v[u] = in ? __ldcs(xr + (int64_t)(d0 + u) * L.col_stride) : 0.f;
#pragma unroll
                for (int u = 0; u < 16; ++u)
                    if (in) atomicAdd(t + d0 + u, v[u]);
            }
        }
    }
    __syncthreads();
    float* out = partials + (size_t)blockIdx.x * K * 65;
    for (int i = tid; i < K * 64; i += CS_THREADS) out[i] = NCHW ? table[(size_t)(i >> 6) * 65 + (i & 63)] : table[i];
    for (int i = tid; i < K; i += CS_THREADS) {
        out[(size_t)K * 64 + i] = (float)cnt[i];
        if (code_counts && cnt[i]) atomicAdd(code_counts + i, cnt[i]);
    }
    if (n_parts_out && blockIdx.x == 0 && tid == 0) *n_parts_out = gridDim.x;
}

