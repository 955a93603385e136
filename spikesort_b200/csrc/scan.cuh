// Exclusive prefix sum of int32 block counts -> int64 offsets (single CTA,
// deterministic).  Shared by detect (K3) and remove_doubles (K4).
#pragma once
#include "common.cuh"

namespace ssb {

constexpr int SCAN_THREADS = 1024;

// offsets[i] = sum(counts[0..i)), offsets[n] = total.
// If row_offsets != nullptr: row_offsets[r] = offsets[r * per_row] for r <= n_rows.
static __global__ void __launch_bounds__(SCAN_THREADS)
scan_counts_kernel(const int32_t *__restrict__ counts, int64_t n, int64_t *__restrict__ offsets,
                   int64_t per_row, int64_t n_rows, int64_t *__restrict__ row_offsets)
{
    __shared__ int64_t warp_tot[32];
    __shared__ int64_t grand;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t per = (n + SCAN_THREADS - 1) / SCAN_THREADS;
    const int64_t lo = min((int64_t)tid * per, n), hi = min(lo + per, n);
    int64_t s = 0;
    for (int64_t i = lo; i < hi; ++i) s += counts[i];
    // inclusive warp scan
    int64_t inc = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int64_t o = __shfl_up_sync(SS_FULL, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int64_t w = warp_tot[lane];
        int64_t winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t o = __shfl_up_sync(SS_FULL, winc, d);
            if (lane >= d) winc += o;
        }
        warp_tot[lane] = winc - w;           // exclusive over warps
        if (lane == 31) grand = winc;
    }
    __syncthreads();
    int64_t run = warp_tot[warp] + inc - s;  // exclusive prefix of this thread's chunk
    for (int64_t i = lo; i < hi; ++i) {
        offsets[i] = run;
        if (row_offsets && per_row > 0 && i % per_row == 0) row_offsets[i / per_row] = run;
        run += counts[i];
    }
    if (tid == 0) {
        offsets[n] = grand;
        if (row_offsets) row_offsets[n_rows] = grand;
    }
}

}  // namespace ssb
