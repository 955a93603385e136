// K2 (auto threshold) and K3 (threshold-crossing detection as an ordered
// stream compaction).
//
// Replaces, in reference src/spike_sort/core/extract.py:
//   :79     thresh = frac * sqrt(row[:10*FS].var(dtype=float64))
//   :86-93  i, = where(op1(row[:-1], thr) & op2(row[1:], thr)); spt = i*1000.0/FS
//
// Detection is three launches and ONE pass over the recording:
//   detect_mark_kernel : warp-ballot crossing bitmask (1 bit/sample) + a count
//                        per 8192-sample block;
//   scan_counts_kernel : exclusive prefix sum of the block counts (row-major, so
//                        output is ordered by contact then by time) + row offsets;
//   detect_emit_kernel : reads only the bitmask, ranks set bits with popc +
//                        warp/block prefix sums, writes t = i*1000.0/FS (IEEE
//                        double mul then div, as NumPy evaluates it).
#include "common.cuh"
#include "scan.cuh"
#include "internal.cuh"

namespace {

using namespace ssi;
constexpr int TH_THREADS = 256;

constexpr int DT_BATCH = 8;     // words (of 32 samples) whose loads are in flight together

// comparisons in float64: exact for int16/float32/float64 inputs (NumPy promotes the
// array to the float64 threshold; the float32-array-vs-Python-float case is handled
// by the host, which rounds the threshold to float32)
template <typename T>
__global__ void __launch_bounds__(32 * DT_WARPS)
detect_mark_kernel(const T *__restrict__ x, int64_t n_samples, int64_t ld,
                   const double *__restrict__ thr, int falling, int64_t blocks_per_row,
                   uint32_t *__restrict__ mask, int32_t *__restrict__ counts)
{
    __shared__ int warp_cnt[DT_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = blockIdx.y, b = blockIdx.x;
    const int64_t blk = r * blocks_per_row + b;
    const T *row = x + r * ld;
    const double th = __ldg(thr + r);
    const int64_t word0 = b * DT_BLOCK_WORDS + warp * DT_WORDS_PER_WARP;
    uint32_t mine = 0;
    int cnt = 0;
    // every sample is loaded ONCE: x[i+1] comes from the neighbouring lane / next word
#pragma unroll 1
    for (int w0 = 0; w0 < DT_WORDS_PER_WARP; w0 += DT_BATCH) {
        const int64_t i0 = (word0 + w0) * 32;
        double v[DT_BATCH + 1];
        if (i0 + 32 * DT_BATCH + 1 <= n_samples) {
#pragma unroll
            for (int k = 0; k < DT_BATCH; ++k) v[k] = (double)__ldg(row + i0 + 32 * k + lane);
            v[DT_BATCH] = (double)__ldg(row + i0 + 32 * DT_BATCH);        // broadcast load
        } else {
#pragma unroll
            for (int k = 0; k <= DT_BATCH; ++k) {
                const int64_t i = i0 + 32 * k + (k < DT_BATCH ? lane : 0);
                // past the end: NaN compares false on both sides -> no crossing
                v[k] = i < n_samples ? (double)__ldg(row + i) : __longlong_as_double(0x7ff8000000000000LL);
            }
        }
#pragma unroll
        for (int k = 0; k < DT_BATCH; ++k) {
            const double dn = ss_shfl_down(v[k], 1);
            const double first_next = ss_shfl(v[k + 1], 0);
            const double nxt = lane == 31 ? (k + 1 < DT_BATCH ? first_next : v[DT_BATCH]) : dn;
            const uint32_t m = __ballot_sync(SS_FULL, crossing(v[k], nxt, th, falling));
            if (lane == w0 + k) mine = m;
            cnt += __popc(m);
        }
    }
    mask[r * blocks_per_row * DT_BLOCK_WORDS + word0 + lane] = mine;   // coalesced
    if (lane == 0) warp_cnt[warp] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
#pragma unroll
        for (int w = 0; w < DT_WARPS; ++w) s += warp_cnt[w];
        counts[blk] = s;
    }
}

// one WARP per 8192-sample block (most blocks hold no crossing and leave at once)
__global__ void __launch_bounds__(32 * DT_WARPS)
detect_emit_kernel(const uint32_t *__restrict__ mask, const int64_t *__restrict__ offsets,
                   int64_t n_blocks, int64_t blocks_per_row, double FS, int64_t index_offset,
                   double *__restrict__ spt, int64_t cap)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t blk = (int64_t)blockIdx.x * DT_WARPS + warp;
    if (blk >= n_blocks) return;
    const int64_t first = offsets[blk];
    if (offsets[blk + 1] == first) return;
    const int64_t b = blk % blocks_per_row;
    int64_t pos = first;
    // all 8 words of a lane are loaded up front; groups of 32 words without a crossing (most
    // of them: a block holds a handful of spikes) cost one ballot
    uint32_t mw[DT_BLOCK_WORDS / 32];
#pragma unroll
    for (int j = 0; j < DT_BLOCK_WORDS / 32; ++j) mw[j] = mask[blk * DT_BLOCK_WORDS + 32 * j + lane];
#pragma unroll
    for (int j = 0; j < DT_BLOCK_WORDS / 32; ++j) {
        uint32_t m = mw[j];
        const unsigned any = __ballot_sync(SS_FULL, m != 0);
        if (any == 0) continue;
        const int c = __popc(m);
        int inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(SS_FULL, inc, d);
            if (lane >= d) inc += o;
        }
        int64_t p = pos + inc - c;
        const int64_t i0 = (b * DT_BLOCK_WORDS + 32 * j + lane) * 32 + index_offset;
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            if (p < cap) spt[p] = __ddiv_rn(__dmul_rn((double)(i0 + bit), 1000.0), FS);
            ++p;
        }
        pos += __shfl_sync(SS_FULL, inc, 31);
    }
}

// counts[blk] = number of set bits in the block's 256 mask words (fused path: the mask
// is filled by atomicOr from the filter kernel, counts are derived afterwards)
__global__ void __launch_bounds__(32 * DT_WARPS)
detect_count_kernel(const uint32_t *__restrict__ mask, int32_t *__restrict__ counts, int64_t n_blocks)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t blk = (int64_t)blockIdx.x * DT_WARPS + warp;       // one warp per block
    if (blk >= n_blocks) return;
    const uint4 *m4 = reinterpret_cast<const uint4 *>(mask + blk * DT_BLOCK_WORDS);
    int c = 0;
#pragma unroll
    for (int k = 0; k < DT_BLOCK_WORDS / 128; ++k) {
        const uint4 v = __ldg(m4 + 32 * k + lane);
        c += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(SS_FULL, c, d);
    if (lane == 0) counts[blk] = c;
}

// ---- auto threshold: shifted sums over the first n0 samples of each row ----
template <typename T>
__global__ void __launch_bounds__(TH_THREADS)
var_partial_kernel(const T *__restrict__ x, int64_t ld, int64_t n0, double *__restrict__ part)
{
    __shared__ double s1s[TH_THREADS / 32], s2s[TH_THREADS / 32];
    const int64_t r = blockIdx.y;
    const T *row = x + r * ld;
    const double shift = (double)__ldg(row);
    const int64_t per = ss_cdiv_dev(n0, TH_PARTS);
    const int64_t lo = min((int64_t)blockIdx.x * per, n0), hi = min(lo + per, n0);
    double s1 = 0.0, s2 = 0.0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += TH_THREADS) {
        const double d = (double)__ldg(row + i) - shift;
        s1 += d;
        s2 += d * d;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        s1 += __shfl_down_sync(SS_FULL, s1, d);
        s2 += __shfl_down_sync(SS_FULL, s2, d);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s1s[warp] = s1; s2s[warp] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int w = 0; w < TH_THREADS / 32; ++w) { a += s1s[w]; b += s2s[w]; }
        part[(r * TH_PARTS + blockIdx.x) * 2 + 0] = a;
        part[(r * TH_PARTS + blockIdx.x) * 2 + 1] = b;
    }
}

__global__ void var_final_kernel(const double *__restrict__ part, int64_t n_rows, int64_t n0,
                                 double factor, double *__restrict__ thr)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    double s1 = 0.0, s2 = 0.0;
    for (int p = 0; p < TH_PARTS; ++p) {
        s1 += part[(r * TH_PARTS + p) * 2 + 0];
        s2 += part[(r * TH_PARTS + p) * 2 + 1];
    }
    const double n = (double)n0;
    double var = (s2 - s1 * s1 / n) / n;
    if (var < 0.0) var = 0.0;
    thr[r] = factor * sqrt(var);
}

}  // namespace

extern "C" size_t ss_threshold_workspace_bytes(int64_t n_rows)
{
    return n_rows > 0 ? (size_t)n_rows * TH_PARTS * 2 * sizeof(double) : 0;
}

extern "C" int ss_var_threshold(const void *d_x, int dtype, int64_t n_rows, int64_t ld, int64_t n0,
                                double factor, double *d_thr, void *d_ws, size_t ws_bytes,
                                void *stream)
{
    SS_REQUIRE(d_x && d_thr && d_ws, SS_ERR_ARG, "var_threshold: null pointer");
    SS_REQUIRE(n_rows > 0 && n_rows <= 65535 && n0 > 0, SS_ERR_ARG,
               "var_threshold: bad shape (%lld rows, n0 %lld)", (long long)n_rows, (long long)n0);
    SS_REQUIRE(ws_bytes >= ss_threshold_workspace_bytes(n_rows), SS_ERR_WORKSPACE,
               "var_threshold: workspace too small");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    double *part = reinterpret_cast<double *>(d_ws);
    const dim3 grid(TH_PARTS, (unsigned)n_rows);
    switch (dtype) {
    case SS_I16: var_partial_kernel<int16_t><<<grid, TH_THREADS, 0, st>>>((const int16_t *)d_x, ld, n0, part); break;
    case SS_F32: var_partial_kernel<float><<<grid, TH_THREADS, 0, st>>>((const float *)d_x, ld, n0, part); break;
    case SS_F64: var_partial_kernel<double><<<grid, TH_THREADS, 0, st>>>((const double *)d_x, ld, n0, part); break;
    default: ss_set_error("var_threshold: bad dtype %d", dtype); return SS_ERR_ARG;
    }
    SS_LAUNCH_CHECK("var_partial_kernel");
    var_final_kernel<<<(unsigned)ss_cdiv(n_rows, 128), 128, 0, st>>>(part, n_rows, n0, factor, d_thr);
    SS_LAUNCH_CHECK("var_final_kernel");
    return SS_OK;
}

extern "C" size_t ss_detect_workspace_bytes(int64_t n_rows, int64_t n_samples)
{
    if (n_rows <= 0 || n_samples <= 0) return 0;
    return det_layout(n_rows, n_samples).total;
}

namespace {
int launch_mark(const void *d_x, int dtype, int64_t n_rows, int64_t n_valid, int64_t ld,
                const double *d_thr, int falling, const DetLayout &L, char *ws, cudaStream_t st)
{
    uint32_t *mask = reinterpret_cast<uint32_t *>(ws + L.off_mask);
    int32_t *counts = reinterpret_cast<int32_t *>(ws + L.off_counts);
    // only the blocks that hold valid samples; rows keep the full-recording layout
    const int64_t blocks_valid = ss_cdiv(n_valid, DT_BLOCK_SAMPLES);
    const dim3 grid((unsigned)blocks_valid, (unsigned)n_rows);
    switch (dtype) {
    case SS_I16: detect_mark_kernel<int16_t><<<grid, 32 * DT_WARPS, 0, st>>>((const int16_t *)d_x, n_valid, ld, d_thr, falling, L.blocks_per_row, mask, counts); break;
    case SS_F32: detect_mark_kernel<float><<<grid, 32 * DT_WARPS, 0, st>>>((const float *)d_x, n_valid, ld, d_thr, falling, L.blocks_per_row, mask, counts); break;
    case SS_F64: detect_mark_kernel<double><<<grid, 32 * DT_WARPS, 0, st>>>((const double *)d_x, n_valid, ld, d_thr, falling, L.blocks_per_row, mask, counts); break;
    default: ss_set_error("detect_mark: bad dtype %d", dtype); return SS_ERR_ARG;
    }
    SS_LAUNCH_CHECK("detect_mark_kernel");
    return SS_OK;
}
}  // namespace

int ssi::mark_head(const void *d_x, int dtype, int64_t n_rows, int64_t n_valid, int64_t n_samples,
                   int64_t ld, const double *d_thr, int falling, void *d_ws, cudaStream_t st)
{
    if (n_valid <= 0) return SS_OK;
    const DetLayout L = det_layout(n_rows, n_samples);
    return launch_mark(d_x, dtype, n_rows, n_valid, ld, d_thr, falling, L,
                       reinterpret_cast<char *>(d_ws), st);
}

namespace {
// exclusive scan of the block counts (row-major) + row offsets
int scan_blocks(const DetLayout &L, int64_t n_rows, char *ws, int64_t *d_offsets, cudaStream_t st)
{
    const int32_t *counts = reinterpret_cast<const int32_t *>(ws + L.off_counts);
    int64_t *offsets = reinterpret_cast<int64_t *>(ws + L.off_offsets);
    if (n_rows >= 2 && L.blocks_per_row >= 256) {
        // one CTA per row (the scratch area behind the offsets holds the row totals: the
        // threshold partial sums that also live there are dead by now)
        int64_t *row_tot = reinterpret_cast<int64_t *>(ws + L.off_scratch);
        ssb::scan_row_totals_kernel<<<(unsigned)n_rows, ssb::SCAN_THREADS, 0, st>>>(counts, L.blocks_per_row, row_tot);
        SS_LAUNCH_CHECK("scan_row_totals_kernel");
        ssb::scan_rows_kernel<<<(unsigned)n_rows, ssb::SCAN_THREADS, 0, st>>>(counts, L.blocks_per_row, n_rows,
                                                                          row_tot, offsets, d_offsets);
        SS_LAUNCH_CHECK("scan_rows_kernel");
        return SS_OK;
    }
    ssb::scan_counts_kernel<<<1, ssb::SCAN_THREADS, 0, st>>>(counts, L.n_blocks, offsets, L.blocks_per_row,
                                                         n_rows, d_offsets);
    SS_LAUNCH_CHECK("scan_counts_kernel");
    return SS_OK;
}
}  // namespace

int ssi::count_and_scan(int64_t n_rows, int64_t n_samples, void *d_ws, int64_t *d_offsets,
                        cudaStream_t st)
{
    const DetLayout L = det_layout(n_rows, n_samples);
    char *ws = reinterpret_cast<char *>(d_ws);
    int32_t *counts = reinterpret_cast<int32_t *>(ws + L.off_counts);
    detect_count_kernel<<<(unsigned)ss_cdiv(L.n_blocks, DT_WARPS), 32 * DT_WARPS, 0, st>>>(
        reinterpret_cast<const uint32_t *>(ws + L.off_mask), counts, L.n_blocks);
    SS_LAUNCH_CHECK("detect_count_kernel");
    return scan_blocks(L, n_rows, ws, d_offsets, st);
}

extern "C" int ss_detect_mark(const void *d_x, int dtype, int64_t n_rows, int64_t n_samples,
                              int64_t ld, const double *d_thr, int falling, void *d_ws,
                              size_t ws_bytes, int64_t *d_offsets, void *stream)
{
    SS_REQUIRE(d_x && d_thr && d_ws && d_offsets, SS_ERR_ARG, "detect_mark: null pointer");
    SS_REQUIRE(n_rows > 0 && n_rows <= 65535 && n_samples > 0, SS_ERR_ARG, "detect_mark: bad shape");
    const DetLayout L = det_layout(n_rows, n_samples);
    SS_REQUIRE(ws_bytes >= L.total, SS_ERR_WORKSPACE, "detect_mark: workspace %zu < %zu", ws_bytes, L.total);
    SS_REQUIRE(L.blocks_per_row < 0x7fffffffLL, SS_ERR_ARG, "detect_mark: too many blocks");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    char *ws = reinterpret_cast<char *>(d_ws);
    const int rc = launch_mark(d_x, dtype, n_rows, n_samples, ld, d_thr, falling, L, ws, st);
    if (rc) return rc;
    return scan_blocks(L, n_rows, ws, d_offsets, st);
}

extern "C" int ss_detect_emit(int64_t n_rows, int64_t n_samples, double FS, int64_t index_offset,
                              const void *d_ws, size_t ws_bytes, double *d_spt, int64_t cap,
                              void *stream)
{
    SS_REQUIRE(d_ws, SS_ERR_ARG, "detect_emit: null pointer");
    SS_REQUIRE(n_rows > 0 && n_samples > 0, SS_ERR_ARG, "detect_emit: bad shape");
    const DetLayout L = det_layout(n_rows, n_samples);
    SS_REQUIRE(ws_bytes >= L.total, SS_ERR_WORKSPACE, "detect_emit: workspace too small");
    if (cap <= 0) return SS_OK;
    SS_REQUIRE(d_spt, SS_ERR_ARG, "detect_emit: null output");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const char *ws = reinterpret_cast<const char *>(d_ws);
    detect_emit_kernel<<<(unsigned)ss_cdiv(L.n_blocks, DT_WARPS), 32 * DT_WARPS, 0, st>>>(
        reinterpret_cast<const uint32_t *>(ws + L.off_mask),
        reinterpret_cast<const int64_t *>(ws + L.off_offsets), L.n_blocks, L.blocks_per_row, FS,
        index_offset, d_spt, cap);
    SS_LAUNCH_CHECK("detect_emit_kernel");
    return SS_OK;
}
