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
    // warp w owns the contiguous segment [w*seg, (w+1)*seg), seg a multiple of 32; lanes
    // read consecutive elements (coalesced), two passes: segment totals, then the scan
    __shared__ int64_t warp_base[32];
    __shared__ int64_t grand;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t seg = ((n + 32 * 32 - 1) / (32 * 32)) * 32;
    const int64_t lo = min((int64_t)warp * seg, n), hi = min(lo + seg, n);
    int64_t s = 0;
    for (int64_t i = lo + lane; i < hi; i += 32) s += counts[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(SS_FULL, s, d);
    if (lane == 0) warp_base[warp] = s;
    __syncthreads();
    if (warp == 0) {
        const int64_t w = warp_base[lane];
        int64_t winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t o = __shfl_up_sync(SS_FULL, winc, d);
            if (lane >= d) winc += o;
        }
        warp_base[lane] = winc - w;          // exclusive over warps
        if (lane == 31) grand = winc;
    }
    __syncthreads();
    int64_t run = warp_base[warp];
    for (int64_t i0 = lo; i0 < hi; i0 += 32) {
        const int64_t i = i0 + lane;
        const int64_t c = i < hi ? (int64_t)counts[i] : 0;
        int64_t inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t o = __shfl_up_sync(SS_FULL, inc, d);
            if (lane >= d) inc += o;
        }
        if (i < hi) {
            const int64_t ex = run + inc - c;
            offsets[i] = ex;
        }
        run += __shfl_sync(SS_FULL, inc, 31);
    }
    if (tid == 0) offsets[n] = grand;
    if (row_offsets) {
        __syncthreads();                     // offsets[] of the other warps (same CTA)
        for (int64_t r = tid; r <= n_rows; r += SCAN_THREADS)
            row_offsets[r] = r < n_rows && r * per_row < n ? offsets[r * per_row] : grand;
    }
}

}  // namespace ssb
