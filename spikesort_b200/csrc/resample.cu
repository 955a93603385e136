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
// splrep(s=0) is FITPACK's fpcurf: Givens rotations turn the banded B-spline collocation
// matrix into a triangle, the same rotations are applied to the waveform, back substitution
// gives the coefficients, splev forms 4-term sums.  Angles, triangle and B-spline values
// depend only on the two time grids: the host computes them once in FITPACK's operation
// order (spikesort_b200/_spline.py).  The kernels replay the waveform-dependent operations
// with unfused float64 instructions in the same order, so every upsampled sample - and with
// it every arg-max - is bit-identical to SciPy's.
//
// Work per waveform: 4m rotations (6 flop), 7m for the back substitution, 8 n_out for the
// evaluation: ~3 kflop at m = 30, resample = 10, against 17 kflop for a dense (n_out x m)
// interpolation matrix.  One thread owns one waveform; its m-vector lives in shared memory
// in [k][thread] order (conflict-free), the plan is read as shared-memory broadcasts.
#include "common.cuh"

namespace {

constexpr int RS_MAX_WIN = 192;
constexpr int RS_BLOCK = 128;          // resample: threads (= waveforms) per CTA
constexpr int RA_BLOCK = 64;           // align: spikes per CTA

struct Plan {
    const double *rc, *rs, *tri, *eh;  // shared-memory copies
    const int *rj, *el;
    int m, n_out;
};

// copies the plan to shared memory at `base` (8-byte aligned); all threads of the CTA call it
__device__ __forceinline__ Plan plan_stage(char *base, const double *__restrict__ f64,
                                           const int32_t *__restrict__ i32, int m, int n_out)
{
    double *d = reinterpret_cast<double *>(base);
    const int nd = 12 * m + 4 * n_out, ni = 4 * m + n_out;
    int *iv = reinterpret_cast<int *>(d + nd);
    for (int i = threadIdx.x; i < nd; i += blockDim.x) d[i] = __ldg(f64 + i);
    for (int i = threadIdx.x; i < ni; i += blockDim.x) iv[i] = __ldg(i32 + i);
    Plan P;
    P.rc = d; P.rs = d + 4 * m; P.tri = d + 8 * m; P.eh = d + 12 * m;
    P.rj = iv; P.el = iv + 4 * m;
    P.m = m; P.n_out = n_out;
    return P;
}

// z[k*stride] holds y on entry (k < m), the B-spline coefficients on exit.
// fpcurf.f rows 'do 130' (right-hand side only) followed by fpback.f.
// y and z share storage: row `it` reads y[it] first and its rotations touch z[j], j <= it
// only (asserted where the plan is built).
__device__ __forceinline__ void spline_coefficients(const Plan &P, double *z, int stride)
{
    const int m = P.m;
    for (int it = 0; it < m; ++it) {
        double yi = z[it * stride];
        z[it * stride] = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = 4 * it + q;
            const int j = P.rj[r];
            if (j < 0) continue;
            const double cs = P.rc[r], sn = P.rs[r];
            const double s2 = z[j * stride];
            z[j * stride] = __dadd_rn(__dmul_rn(cs, s2), __dmul_rn(sn, yi));      // fprota: b
            yi = __dsub_rn(__dmul_rn(cs, yi), __dmul_rn(sn, s2));                 //         a
        }
    }
    for (int i = m - 1; i >= 0; --i) {
        double store = z[i * stride];
        const int nt = min(3, m - 1 - i);
        for (int q = 1; q <= nt; ++q)
            store = __dsub_rn(store, __dmul_rn(z[(i + q) * stride], P.tri[4 * i + q]));
        z[i * stride] = __ddiv_rn(store, P.tri[4 * i]);
    }
}

// splev.f 'do 60': value of the spline at new abscissa q
__device__ __forceinline__ double spline_value(const Plan &P, const double *z, int stride, int q)
{
    const int l = P.el[q];
    double sp = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        sp = __dadd_rn(sp, __dmul_rn(z[(l + j) * stride], P.eh[4 * q + j]));
    return sp;
}

// thread per (spike, contact); out is (n_out, n_spk, n_c) float64
__global__ void __launch_bounds__(RS_BLOCK)
resample_kernel(const void *__restrict__ waves, int dtype, int n_win, int64_t n_pairs,
                const double *__restrict__ plan_f64, const int32_t *__restrict__ plan_i32,
                int n_out, double *__restrict__ out)
{
    extern __shared__ __align__(16) char smem[];
    double *zs = reinterpret_cast<double *>(smem);                      // [n_win][RS_BLOCK]
    const Plan P = plan_stage(smem + (size_t)n_win * RS_BLOCK * sizeof(double), plan_f64, plan_i32,
                              n_win, n_out);
    __syncthreads();
    const int64_t idx = (int64_t)blockIdx.x * RS_BLOCK + threadIdx.x;
    if (idx >= n_pairs) return;
    double *z = zs + threadIdx.x;
    for (int k = 0; k < n_win; ++k) z[k * RS_BLOCK] = ss_load(waves, dtype, (int64_t)k * n_pairs + idx);
    spline_coefficients(P, z, RS_BLOCK);
    for (int q = 0; q < n_out; ++q)
        __stcs(out + (int64_t)q * n_pairs + idx, spline_value(P, z, RS_BLOCK, q));
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

// thread per spike; a warp loads the windows of its active spikes cooperatively (coalesced)
__global__ void __launch_bounds__(RA_BLOCK)
align_resampled_kernel(const void *__restrict__ x, int dtype, int64_t n_samples, int64_t ld, double FS,
                       const int32_t *__restrict__ rows, int64_t n_rows,
                       const int64_t *__restrict__ offsets, int win0, int n_win, double t_min,
                       double t_max, double lo, double hi, int use_min,
                       const double *__restrict__ plan_f64, const int32_t *__restrict__ plan_i32,
                       int n_out, const double *__restrict__ rtime,
                       double *__restrict__ spt, int64_t n_spk, int max_iter,
                       int64_t index_offset, int64_t halo_before, int64_t halo_after,
                       int32_t *__restrict__ halo_miss)
{
    extern __shared__ __align__(16) char smem[];
    double *zs = reinterpret_cast<double *>(smem);                      // [n_win][RA_BLOCK]
    const Plan P = plan_stage(smem + (size_t)n_win * RA_BLOCK * sizeof(double), plan_f64, plan_i32,
                              n_win, n_out);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp0 = threadIdx.x & ~31;
    const int64_t s = (int64_t)blockIdx.x * RA_BLOCK + threadIdx.x;
    double t = 0.0;
    int64_t row0 = 0;
    bool active = s < n_spk && !(offsets && s >= __ldg(offsets + n_rows)), moved = false;
    if (active) {
        int64_t r = n_rows > 1 ? rs_find_row(offsets, n_rows, s) : 0;
        if (rows) r = rows[r];
        row0 = r * ld;
        t = spt[s];
    }
    double *z = zs + threadIdx.x;
    for (int it = 0; it < max_iter; ++it) {
        active = active && (t >= t_min && t <= t_max);                  // filter_spt, :108-110
        const unsigned todo = __ballot_sync(SS_FULL, active);
        if (todo == 0) break;
        const int centre = active ? __double2int_rz(__dmul_rn(__ddiv_rn(t, 1000.0), FS)) : 0;   // :150
        const int64_t first = (int64_t)centre + win0 - index_offset;    // local index in the row
        // float32-rounded window (extract.py:146) of every active spike of the warp
        for (unsigned rest = todo; rest; rest &= rest - 1) {
            const int src = __ffs(rest) - 1;
            const int64_t f = __shfl_sync(SS_FULL, first, src);
            const int64_t base = __shfl_sync(SS_FULL, row0, src);
            for (int k = lane; k < n_win; k += 32) {
                const int64_t i = f + k;
                const bool stored = i >= -halo_before && i < n_samples + halo_after;
                zs[k * RA_BLOCK + warp0 + src] = stored ? (double)ss_load_f32(x, dtype, base + i) : 0.0;
                if (!stored && halo_miss && ss_beyond_halo(i, n_samples, halo_before, halo_after))
                    atomicAdd(halo_miss, 1);
            }
        }
        __syncwarp();
        if (active) {
            spline_coefficients(P, z, RA_BLOCK);
            double best = 0.0;
            int best_q = 0;
            for (int q = 0; q < n_out; ++q) {                            // first extremum, :251-254
                double v = spline_value(P, z, RA_BLOCK, q);
                if (use_min) v = -v;
                if (q == 0 || v > best) { best = v; best_q = q; }
            }
            const double shift = __ldg(rtime + best_q);                 // time[i], :257
            t = __dadd_rn(t, shift);
            moved = true;
            active = (shift < lo || shift > hi);                        // :263-264
        }
        __syncwarp();
    }
    if (moved) spt[s] = t;
}

size_t rs_smem(int n_win, int n_out, int block)
{
    return (size_t)n_win * block * sizeof(double) + (size_t)(12 * n_win + 4 * n_out) * sizeof(double) +
           (size_t)(4 * n_win + n_out) * sizeof(int);
}

}  // namespace

extern "C" int ss_resample(const void *d_waves, int dtype, int n_win, int64_t n_spk, int n_c,
                           const double *d_plan_f64, const int32_t *d_plan_i32, int n_out,
                           double *d_out, void *stream)
{
    if (n_spk == 0 || n_out == 0 || n_c == 0) return SS_OK;
    SS_REQUIRE(d_waves && d_plan_f64 && d_plan_i32 && d_out, SS_ERR_ARG, "resample: null pointer");
    SS_REQUIRE(dtype >= SS_I16 && dtype <= SS_F64, SS_ERR_ARG, "resample: bad dtype %d", dtype);
    SS_REQUIRE(n_win >= 4 && n_win <= RS_MAX_WIN && n_out >= 1, SS_ERR_UNSUPPORTED,
               "resample: n_win %d outside [4, %d]", n_win, RS_MAX_WIN);
    const size_t smem = rs_smem(n_win, n_out, RS_BLOCK);
    SS_REQUIRE(smem <= 220 * 1024, SS_ERR_UNSUPPORTED,
               "resample: window %d -> %d points does not fit in shared memory", n_win, n_out);
    SS_CUDA_CHECK(cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t n_pairs = n_spk * n_c;
    resample_kernel<<<(unsigned)ss_cdiv(n_pairs, RS_BLOCK), RS_BLOCK, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        d_waves, dtype, n_win, n_pairs, d_plan_f64, d_plan_i32, n_out, d_out);
    SS_LAUNCH_CHECK("resample_kernel");
    return SS_OK;
}

extern "C" int ss_align_resampled(const void *d_x, int dtype, int64_t n_samples, int64_t ld, double FS,
                                  const int32_t *d_rows, int64_t n_rows, const int64_t *d_offsets,
                                  int win0, int win1, double t_min, double t_max, double lo, double hi,
                                  int use_min, const double *d_plan_f64, const int32_t *d_plan_i32,
                                  int n_out, const double *d_rtime, double *d_spt, int64_t n_spk,
                                  int max_iter, int64_t index_offset, int64_t halo_before,
                                  int64_t halo_after, int32_t *d_halo_miss, void *stream)
{
    if (n_spk == 0) return SS_OK;
    SS_REQUIRE(d_x && d_spt && d_plan_f64 && d_plan_i32 && d_rtime, SS_ERR_ARG, "align_resampled: null pointer");
    SS_REQUIRE(dtype >= SS_I16 && dtype <= SS_F64, SS_ERR_ARG, "align_resampled: bad dtype %d", dtype);
    SS_REQUIRE(n_rows >= 1 && (n_rows == 1 || d_offsets), SS_ERR_ARG, "align_resampled: offsets required");
    const int n_win = win1 - win0;
    SS_REQUIRE(n_win >= 4 && n_win <= RS_MAX_WIN && n_out >= 1 && max_iter > 0, SS_ERR_UNSUPPORTED,
               "align_resampled: window of %d samples outside [4, %d]", n_win, RS_MAX_WIN);
    const size_t smem = rs_smem(n_win, n_out, RA_BLOCK);
    SS_REQUIRE(smem <= 220 * 1024, SS_ERR_UNSUPPORTED,
               "align_resampled: window %d -> %d points does not fit in shared memory", n_win, n_out);
    SS_CUDA_CHECK(cudaFuncSetAttribute(align_resampled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    align_resampled_kernel<<<(unsigned)ss_cdiv(n_spk, RA_BLOCK), RA_BLOCK, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        d_x, dtype, n_samples, ld, FS, d_rows, n_rows, d_offsets, win0, n_win, t_min, t_max, lo, hi,
        use_min, d_plan_f64, d_plan_i32, n_out, d_rtime, d_spt, n_spk, max_iter, index_offset,
        halo_before, halo_after, d_halo_miss);
    SS_LAUNCH_CHECK("align_resampled_kernel");
    return SS_OK;
}
