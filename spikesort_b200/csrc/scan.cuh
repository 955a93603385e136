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

// The same for counts laid out [n_rows][per_row] with MANY rows of blocks (detection: 32-384
// contacts x thousands of 8192-sample blocks): one CTA per row, two launches - row totals, then
// every row scans its own counts from the sum of the rows before it.  (The single CTA above
// took 36 us for the 70 k counts of configs[1].)
static __global__ void __launch_bounds__(SCAN_THREADS)
scan_row_totals_kernel(const int32_t *__restrict__ counts, int64_t per_row, int64_t *__restrict__ row_tot)
{
    __shared__ int64_t wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t *c = counts + (int64_t)blockIdx.x * per_row;
    int64_t s = 0;
    for (int64_t i = tid; i < per_row; i += SCAN_THREADS) s += c[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(SS_FULL, s, d);
    if (lane == 0) wsum[warp] = s;
    __syncthreads();
    if (warp == 0) {
        s = wsum[lane];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(SS_FULL, s, d);
        if (lane == 0) row_tot[blockIdx.x] = s;
    }
}

static __global__ void __launch_bounds__(SCAN_THREADS)
scan_rows_kernel(const int32_t *__restrict__ counts, int64_t per_row, int64_t n_rows,
                 const int64_t *__restrict__ row_tot, int64_t *__restrict__ offsets,
                 int64_t *__restrict__ row_offsets)
{
    __shared__ int64_t warp_base[32];
    __shared__ int64_t s_base, s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t r = blockIdx.x;
    if (warp == 0) {                                      // rows before this one, in row order
        int64_t b = 0;
        for (int64_t k = lane; k < r; k += 32) b += row_tot[k];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) b += __shfl_xor_sync(SS_FULL, b, d);
        if (lane == 0) { s_base = b; s_carry = b; }
    }
    __syncthreads();
    const int32_t *c = counts + r * per_row;
    int64_t *o = offsets + r * per_row;
    for (int64_t i0 = 0; i0 < per_row; i0 += SCAN_THREADS) {
        const int64_t i = i0 + tid;
        const int64_t v = i < per_row ? (int64_t)c[i] : 0;
        int64_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t up = __shfl_up_sync(SS_FULL, inc, d);
            if (lane >= d) inc += up;
        }
        if (lane == 31) warp_base[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const int64_t w = warp_base[lane];
            int64_t winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int64_t up = __shfl_up_sync(SS_FULL, winc, d);
                if (lane >= d) winc += up;
            }
            warp_base[lane] = winc - w;
        }
        __syncthreads();
        const int64_t carry = s_carry;
        if (i < per_row) o[i] = carry + warp_base[warp] + inc - v;
        __syncthreads();
        if (tid == SCAN_THREADS - 1) s_carry = carry + warp_base[warp] + inc;
        __syncthreads();
    }
    if (tid == 0) {
        row_offsets[r] = s_base;
        if (r == n_rows - 1) {
            row_offsets[n_rows] = s_carry;
            offsets[n_rows * per_row] = s_carry;
        }
    }
}

}  // namespace ssb
