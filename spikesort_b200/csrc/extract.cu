// K5 - waveform extraction: gather the sp_win neighbourhood of every spike on
// the selected contacts into the float32 (n_win, n_spk, n_c) block.
//
// Replaces extract_spikes, reference src/spike_sort/core/extract.py:146-177
// (resample == 1): indices = (spt/1000.0*FS).astype(int32) (truncation, :150),
// inner spikes copied whole (:160-163), outer spikes clipped to the recording and
// zero-filled (:165-170), is_valid = inner (:174-177).
//
// A CTA stages a tile of EX_PAIRS (spike, contact) windows in shared memory:
//   load  : one warp per window, lanes along time -> coalesced reads of the
//           n_win contiguous samples, rounded to float32;
//   store : threads along (spike, contact) for each time step -> contiguous
//           runs of the output's fastest axes.
// The smem tile is padded to EX_PAIRS+1 words per time step, which makes both
// phases bank-conflict free.
#include "common.cuh"

namespace {

constexpr int EX_PAIRS = 256;           // (spike, contact) windows per CTA = threads per CTA
constexpr int EX_MAX_WIN = 192;          // 192 * 257 * 4 B = 197 KB of shared memory

__device__ __forceinline__ int64_t ex_find_row(const int64_t *__restrict__ offsets, int64_t n_rows,
                                               int64_t s)
{
    int64_t lo = 0, hi = n_rows;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid) <= s) lo = mid; else hi = mid;
    }
    return lo;
}

// grid: (ceil(n_spk / SB), ceil(n_c / CB)), SB = EX_PAIRS / CB
__global__ void __launch_bounds__(EX_PAIRS)
extract_kernel(const void *__restrict__ x, int dtype, int64_t n_samples, int64_t ld, double FS,
               const int32_t *__restrict__ contacts, int n_c, int CB,
               const int32_t *__restrict__ rows, int64_t n_rows, const int64_t *__restrict__ offsets,
               int win0, int n_win, double t_min, double t_max,
               const double *__restrict__ spt, int64_t n_spk,
               float *__restrict__ waves, uint8_t *__restrict__ valid, int32_t *__restrict__ n_invalid,
               int64_t index_offset, int64_t halo_before, int64_t halo_after,
               int32_t *__restrict__ halo_miss)
{
    extern __shared__ float tile[];                 // [n_win][EX_PAIRS + 1]
    __shared__ int64_t s_first[EX_PAIRS];           // first sample of the window, per local spike
    __shared__ int64_t s_row0[EX_PAIRS];            // element offset of the row (row mode)
    const int SB = EX_PAIRS / CB;
    const int64_t s0 = (int64_t)blockIdx.x * SB;
    const int c0 = blockIdx.y * CB;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // the list may have spare capacity: entries at or beyond offsets[n_rows] are not spikes
    // (their waveforms are written as zeros, their valid flag as 0)
    const int64_t n_live = (!contacts && offsets) ? min(n_spk, __ldg(offsets + n_rows)) : n_spk;

    if (tid < SB) {
        const int64_t s = s0 + tid;
        int64_t first = 0, row0 = 0;
        if (s >= n_live && s < n_spk && blockIdx.y == 0) valid[s] = 0;
        if (s < n_live) {
            const double t = spt[s];
            const int centre = __double2int_rz(__dmul_rn(__ddiv_rn(t, 1000.0), FS));
            first = (int64_t)centre + win0 - index_offset;          // local index
            if (!contacts) {
                int64_t r = n_rows > 1 ? ex_find_row(offsets, n_rows, s) : 0;
                if (rows) r = rows[r];
                row0 = r * ld;
            }
            if (blockIdx.y == 0) {
                const bool ok = (t >= t_min) && (t <= t_max);
                valid[s] = ok ? 1 : 0;
                if (!ok) atomicAdd(n_invalid, 1);
            }
        }
        s_first[tid] = first;
        s_row0[tid] = row0;
    }
    __syncthreads();

    // load phase: warp w takes windows p = w, w + 8, ...
    for (int p = warp; p < EX_PAIRS; p += EX_PAIRS / 32) {
        const int sl = p / CB, cl = p - sl * CB;
        const int64_t s = s0 + sl;
        const int c = c0 + cl;
        const bool live = s < n_live && c < n_c;
        int64_t base = 0;
        if (live) base = contacts ? (int64_t)__ldg(contacts + c) * ld : s_row0[sl];
        const int64_t first = s_first[sl];
        for (int w = lane; w < n_win; w += 32) {
            const int64_t i = first + w;
            float v = 0.f;
            if (live && i >= -halo_before && i < n_samples + halo_after)
                v = ss_load_f32(x, dtype, base + i);
            else if (live && halo_miss && ss_beyond_halo(i, n_samples, halo_before, halo_after))
                atomicAdd(halo_miss, 1);
            tile[w * (EX_PAIRS + 1) + p] = v;
        }
    }
    __syncthreads();

    // store phase: thread p = (sl, cl) writes one element per time step
    const int sl = tid / CB, cl = tid - sl * CB;
    const int64_t s = s0 + sl;
    const int c = c0 + cl;
    if (s < n_spk && c < n_c) {
        float *dst = waves + s * n_c + c;
        const int64_t step = n_spk * (int64_t)n_c;
        for (int w = 0; w < n_win; ++w) dst[w * step] = tile[w * (EX_PAIRS + 1) + tid];
    }
}

// Row mode (one contact per spike list, n_c == 1): one THREAD per spike.  The n_win loads of
// a thread are independent and stay inside the 2-3 cache lines of its window; for every time
// step the threads of a warp store 32 consecutive floats of the output row - coalesced
// without a shared-memory transposition.
__global__ void __launch_bounds__(256)
extract_rows_kernel(const void *__restrict__ x, int dtype, int64_t n_samples, int64_t ld, double FS,
                    const int32_t *__restrict__ rows, int64_t n_rows, const int64_t *__restrict__ offsets,
                    int win0, int n_win, double t_min, double t_max,
                    const double *__restrict__ spt, int64_t n_spk,
                    float *__restrict__ waves, uint8_t *__restrict__ valid, int32_t *__restrict__ n_invalid,
                    int64_t index_offset, int64_t halo_before, int64_t halo_after,
                    int32_t *__restrict__ halo_miss)
{
    const int64_t s = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (s >= n_spk) return;
    const int64_t n_live = offsets ? min(n_spk, __ldg(offsets + n_rows)) : n_spk;
    float *dst = waves + s;
    if (s >= n_live) {                                  // spare capacity: zeros, not a spike
        valid[s] = 0;
        for (int w = 0; w < n_win; ++w) dst[(int64_t)w * n_spk] = 0.f;
        return;
    }
    const double t = spt[s];
    const int centre = __double2int_rz(__dmul_rn(__ddiv_rn(t, 1000.0), FS));
    const int64_t first = (int64_t)centre + win0 - index_offset;          // local index
    int64_t r = n_rows > 1 ? ex_find_row(offsets, n_rows, s) : 0;
    if (rows) r = rows[r];
    const int64_t row0 = r * ld;
    const bool ok = (t >= t_min) && (t <= t_max);
    valid[s] = ok ? 1 : 0;
    if (!ok) atomicAdd(n_invalid, 1);
    if (first >= -halo_before && first + n_win <= n_samples + halo_after) {
        const int64_t base = row0 + first;
        if (dtype == SS_F64) {
            const double *p = reinterpret_cast<const double *>(x) + base;
            for (int w = 0; w < n_win; ++w) dst[(int64_t)w * n_spk] = __double2float_rn(__ldg(p + w));
        } else {
            for (int w = 0; w < n_win; ++w) dst[(int64_t)w * n_spk] = ss_load_f32(x, dtype, base + w);
        }
    } else {
        bool missed = false;
        for (int w = 0; w < n_win; ++w) {
            const int64_t i = first + w;
            float v = 0.f;
            if (i >= -halo_before && i < n_samples + halo_after) v = ss_load_f32(x, dtype, row0 + i);
            else missed = missed || ss_beyond_halo(i, n_samples, halo_before, halo_after);
            dst[(int64_t)w * n_spk] = v;
        }
        if (missed && halo_miss) atomicAdd(halo_miss, 1);
    }
}

}  // namespace

extern "C" int ss_max_window(void) { return EX_MAX_WIN; }

extern "C" int ss_extract(const void *d_x, int dtype, int64_t n_samples, int64_t ld, double FS,
                          const int32_t *d_contacts, int n_c, const int32_t *d_rows, int64_t n_rows,
                          const int64_t *d_offsets, int win0, int win1, double t_min, double t_max,
                          const double *d_spt, int64_t n_spk, float *d_waves, uint8_t *d_valid,
                          int32_t *d_n_invalid, int64_t index_offset, int64_t halo_before,
                          int64_t halo_after, int32_t *d_halo_miss, void *stream)
{
    SS_REQUIRE(d_n_invalid, SS_ERR_ARG, "extract: null d_n_invalid");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    SS_CUDA_CHECK(cudaMemsetAsync(d_n_invalid, 0, sizeof(int32_t), st));
    if (n_spk == 0) return SS_OK;
    SS_REQUIRE(d_x && d_spt && d_waves && d_valid, SS_ERR_ARG, "extract: null pointer");
    SS_REQUIRE(dtype >= SS_I16 && dtype <= SS_F64, SS_ERR_ARG, "extract: bad dtype %d", dtype);
    SS_REQUIRE(n_c >= 1 && n_spk > 0, SS_ERR_ARG, "extract: bad shape");
    SS_REQUIRE(win1 > win0, SS_ERR_ARG, "extract: empty window [%d, %d)", win0, win1);
    const int n_win = win1 - win0;
    SS_REQUIRE(n_win <= EX_MAX_WIN, SS_ERR_UNSUPPORTED, "extract: window of %d samples > %d", n_win,
               EX_MAX_WIN);
    if (!d_contacts) {
        SS_REQUIRE(n_c == 1, SS_ERR_ARG, "extract: row mode needs n_c == 1");
        SS_REQUIRE(n_rows >= 1 && (n_rows == 1 || d_offsets), SS_ERR_ARG,
                   "extract: offsets required for %lld rows", (long long)n_rows);
    }
    if (!d_contacts) {
        extract_rows_kernel<<<(unsigned)ss_cdiv(n_spk, 256), 256, 0, st>>>(
            d_x, dtype, n_samples, ld, FS, d_rows, n_rows, d_offsets, win0, n_win, t_min, t_max,
            d_spt, n_spk, d_waves, d_valid, d_n_invalid, index_offset, halo_before, halo_after,
            d_halo_miss);
        SS_LAUNCH_CHECK("extract_rows_kernel");
        return SS_OK;
    }
    int CB = 1;
    while (CB < n_c && CB < 16) CB <<= 1;           // contacts per CTA: 1, 2, 4, 8 or 16
    const int SB = EX_PAIRS / CB;
    const size_t smem = (size_t)n_win * (EX_PAIRS + 1) * sizeof(float);
    SS_CUDA_CHECK(cudaFuncSetAttribute(extract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)((size_t)EX_MAX_WIN * (EX_PAIRS + 1) * sizeof(float))));
    const dim3 grid((unsigned)ss_cdiv(n_spk, SB), (unsigned)ss_cdiv(n_c, CB));
    extract_kernel<<<grid, EX_PAIRS, smem, st>>>(d_x, dtype, n_samples, ld, FS, d_contacts, n_c, CB,
                                                d_rows, n_rows, d_offsets, win0, n_win, t_min, t_max,
                                                d_spt, n_spk, d_waves, d_valid, d_n_invalid,
                                                index_offset, halo_before, halo_after, d_halo_miss);
    SS_LAUNCH_CHECK("extract_kernel");
    return SS_OK;
}
