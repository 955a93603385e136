// "Next" row, SURVEY.md 8f rank 1: spline upsampling of spike waveforms and alignment on
// the upsampled waveform.
//
// Replaces, in reference src/spike_sort/core/extract.py:
//   :187-205  resample_spikes: per spike and contact, splrep(time, wave, s=0) (interpolating
//             cubic B-spline, not-a-knot ends) evaluated by splev on
//             arange(time[0], time[-1], 1000/FS_new)
//   :245-258  align_spikes(resample != 1): arg-max/min of the UPSAMPLED float64 window,
//             shift = resamp_time[k]
//
// An interpolating spline on a fixed grid is a LINEAR map of the samples: out = W y with
// W (n_out x n_win) depending only on the two time grids.  The host builds W once with the
// same SciPy routines the reference calls (W[:, j] = splev(rt, splrep(time, e_j, s=0)));
// the kernels apply it in float64.  Agreement with FITPACK's per-waveform solve is at
// round-off (~1e-13 relative), not bit-exact; exact ties between upsampled points can
// therefore resolve differently (counted in the tests).
#include "common.cuh"

namespace {

constexpr int RS_MAX_WIN = 192;
constexpr int RS_WARPS = 8;
constexpr int RS_NONE = 0x7fffffff;

// thread per (spike, contact); W staged in shared memory; out is (n_out, n_spk, n_c) float64
__global__ void __launch_bounds__(256)
resample_kernel(const void *__restrict__ waves, int dtype, int n_win, int64_t n_pairs,
                const double *__restrict__ W, int n_out, double *__restrict__ out)
{
    extern __shared__ double Ws[];                         // [n_out][n_win]
    for (int i = threadIdx.x; i < n_out * n_win; i += blockDim.x) Ws[i] = W[i];
    __syncthreads();
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_pairs) return;
    for (int j = 0; j < n_out; ++j) {
        double acc = 0.0;
        for (int w = 0; w < n_win; ++w)
            acc += Ws[j * n_win + w] * ss_load(waves, dtype, (int64_t)w * n_pairs + idx);
        out[(int64_t)j * n_pairs + idx] = acc;
    }
}

__device__ __forceinline__ int64_t rs_find_row(const int64_t *__restrict__ offsets, int64_t n_rows, int64_t s)
{
    int64_t lo = 0, hi = n_rows;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid) <= s) lo = mid; else hi = mid;
    }
    return lo;
}

// one warp per spike, as align_kernel, but the extremum is searched on W y
__global__ void __launch_bounds__(32 * RS_WARPS)
align_resampled_kernel(const void *__restrict__ x, int dtype, int64_t n_samples, int64_t ld, double FS,
                       const int32_t *__restrict__ rows, int64_t n_rows,
                       const int64_t *__restrict__ offsets, int win0, int n_win, double t_min,
                       double t_max, double lo, double hi, int use_min,
                       const double *__restrict__ W, int n_out, const double *__restrict__ rtime,
                       double *__restrict__ spt, int64_t n_spk, int max_iter,
                       int64_t index_offset, int64_t halo_before, int64_t halo_after)
{
    extern __shared__ double sm[];                         // W [n_out][n_win], then y per warp
    double *Ws = sm;
    double *ys = sm + (size_t)n_out * n_win + (threadIdx.x >> 5) * n_win;
    for (int i = threadIdx.x; i < n_out * n_win; i += blockDim.x) Ws[i] = W[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * RS_WARPS + (threadIdx.x >> 5);
    if (s >= n_spk) return;
    int64_t r = n_rows > 1 ? rs_find_row(offsets, n_rows, s) : 0;
    if (rows) r = rows[r];
    const int64_t row0 = r * ld;
    double t = spt[s];
    bool moved = false;
    for (int it = 0; it < max_iter; ++it) {
        if (!(t >= t_min && t <= t_max)) break;
        const int centre = __double2int_rz(__dmul_rn(__ddiv_rn(t, 1000.0), FS));
        const int64_t first = (int64_t)centre + win0 - index_offset;   // local index
        for (int k = lane; k < n_win; k += 32) {
            const int64_t i = first + k;
            ys[k] = (i >= -halo_before && i < n_samples + halo_after)
                        ? (double)ss_load_f32(x, dtype, row0 + i) : 0.0;
        }
        __syncwarp();
        double best = 0.0;
        int best_k = RS_NONE;
        for (int j = lane; j < n_out; j += 32) {
            double acc = 0.0;
            for (int w = 0; w < n_win; ++w) acc += Ws[j * n_win + w] * ys[w];
            if (use_min) acc = -acc;
            if (best_k == RS_NONE || acc > best) { best = acc; best_k = j; }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const double ov = ss_shfl(best, lane ^ d);
            const int ok = __shfl_xor_sync(SS_FULL, best_k, d);
            if (ok != RS_NONE && (best_k == RS_NONE || ov > best || (ov == best && ok < best_k))) {
                best = ov;
                best_k = ok;
            }
        }
        __syncwarp();
        const double shift = __ldg(rtime + best_k);          // time[i] of the upsampled grid
        t = __dadd_rn(t, shift);
        moved = true;
        if (!(shift < lo || shift > hi)) break;
    }
    if (moved && lane == 0) spt[s] = t;
}

}  // namespace

extern "C" int ss_resample(const void *d_waves, int dtype, int n_win, int64_t n_spk, int n_c,
                           const double *d_W, int n_out, double *d_out, void *stream)
{
    if (n_spk == 0 || n_out == 0) return SS_OK;
    SS_REQUIRE(d_waves && d_W && d_out, SS_ERR_ARG, "resample: null pointer");
    SS_REQUIRE(dtype >= SS_I16 && dtype <= SS_F64, SS_ERR_ARG, "resample: bad dtype %d", dtype);
    SS_REQUIRE(n_win >= 1 && n_win <= RS_MAX_WIN && n_out >= 1, SS_ERR_UNSUPPORTED,
               "resample: n_win %d outside [1, %d]", n_win, RS_MAX_WIN);
    const size_t smem = (size_t)n_out * n_win * sizeof(double);
    SS_REQUIRE(smem <= 200 * 1024, SS_ERR_UNSUPPORTED, "resample: %d x %d table does not fit in shared memory", n_out, n_win);
    SS_CUDA_CHECK(cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t n_pairs = n_spk * n_c;
    resample_kernel<<<(unsigned)ss_cdiv(n_pairs, 256), 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        d_waves, dtype, n_win, n_pairs, d_W, n_out, d_out);
    SS_LAUNCH_CHECK("resample_kernel");
    return SS_OK;
}

extern "C" int ss_align_resampled(const void *d_x, int dtype, int64_t n_samples, int64_t ld, double FS,
                                  const int32_t *d_rows, int64_t n_rows, const int64_t *d_offsets,
                                  int win0, int win1, double t_min, double t_max, double lo, double hi,
                                  int use_min, const double *d_W, int n_out, const double *d_rtime,
                                  double *d_spt, int64_t n_spk, int max_iter, int64_t index_offset,
                                  int64_t halo_before, int64_t halo_after, void *stream)
{
    if (n_spk == 0) return SS_OK;
    SS_REQUIRE(d_x && d_spt && d_W && d_rtime, SS_ERR_ARG, "align_resampled: null pointer");
    SS_REQUIRE(dtype >= SS_I16 && dtype <= SS_F64, SS_ERR_ARG, "align_resampled: bad dtype %d", dtype);
    SS_REQUIRE(n_rows >= 1 && (n_rows == 1 || d_offsets), SS_ERR_ARG, "align_resampled: offsets required");
    const int n_win = win1 - win0;
    SS_REQUIRE(n_win >= 1 && n_win <= RS_MAX_WIN && n_out >= 1 && max_iter > 0, SS_ERR_UNSUPPORTED,
               "align_resampled: window of %d samples outside [1, %d]", n_win, RS_MAX_WIN);
    const size_t smem = ((size_t)n_out * n_win + (size_t)RS_WARPS * n_win) * sizeof(double);
    SS_REQUIRE(smem <= 200 * 1024, SS_ERR_UNSUPPORTED, "align_resampled: %d x %d table does not fit in shared memory", n_out, n_win);
    SS_CUDA_CHECK(cudaFuncSetAttribute(align_resampled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    align_resampled_kernel<<<(unsigned)ss_cdiv(n_spk, RS_WARPS), 32 * RS_WARPS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        d_x, dtype, n_samples, ld, FS, d_rows, n_rows, d_offsets, win0, n_win, t_min, t_max, lo, hi,
        use_min, d_W, n_out, d_rtime, d_spt, n_spk, max_iter, index_offset, halo_before, halo_after);
    SS_LAUNCH_CHECK("align_resampled_kernel");
    return SS_OK;
}
