// "Next" rows of the hot-path scope: the cheap per-spike features
// every example combines with the PCA scores, and the column normalisation of combine().
//
// Replaces, in reference src/spike_sort/core/features.py:
//   :473-474  fetP2P  np.ptp(spikes, axis=0)            (float32: max - min in float32)
//   :483-484  fetMin  np.min(spikes, axis=0)
//   :191-193  normalize: data - data.min(0), then / data.max(0)   (float64)
// All three are bit-exact: min/max do not depend on the reduction order, and
// max_i fl(x_i - m) == fl(max_i x_i - m) because rounding is monotone.
#include "common.cuh"

namespace {

constexpr int NM_THREADS = 256;
constexpr int NM_PARTS = 592;           // row ranges (4 per SM)

__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }

// one thread per (spike, contact): waves is (n_win, n_spk, n_c), outputs (n_spk, n_c) in the
// waveforms' own dtype (np.ptp / np.min keep it: float32 from extract_spikes, float64 from
// resample_spikes)
template <typename T>
__global__ void __launch_bounds__(256)
wave_minmax_kernel(const T *__restrict__ waves, int n_win, int64_t n_pairs,
                   T *__restrict__ mn_out, T *__restrict__ p2p_out)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_pairs) return;
    T mn = __ldg(waves + idx), mx = mn;
    for (int w = 1; w < n_win; ++w) {
        const T v = __ldg(waves + (int64_t)w * n_pairs + idx);
        mn = v < mn ? v : mn;
        mx = v > mx ? v : mx;
    }
    if (mn_out) mn_out[idx] = mn;
    if (p2p_out) p2p_out[idx] = sub_rn(mx, mn);
}

// partial column min/max of a row-major (n_rows, n_cols) float64 matrix: block b takes a
// row range, threads run over its elements linearly (coalesced) and fold into per-column
// shared slots with an ordered-integer atomic min/max
__device__ __forceinline__ long long dbl_key(double v)
{
    const long long b = __double_as_longlong(v);
    return b >= 0 ? b : b ^ 0x7fffffffffffffffLL;          // monotone map double -> int64
}
__device__ __forceinline__ double key_dbl(long long k)
{
    return __longlong_as_double(k >= 0 ? k : k ^ 0x7fffffffffffffffLL);
}

__global__ void __launch_bounds__(NM_THREADS)
colminmax_partial_kernel(const double *__restrict__ x, int64_t n_rows, int n_cols,
                         long long *__restrict__ part)     // [parts][n_cols][2] keys
{
    extern __shared__ long long sh[];                       // [n_cols][2]
    for (int i = threadIdx.x; i < 2 * n_cols; i += NM_THREADS)
        sh[i] = (i & 1) ? dbl_key(-INFINITY) : dbl_key(INFINITY);
    __syncthreads();
    const int64_t per = (n_rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = min((int64_t)blockIdx.x * per, n_rows), r1 = min(r0 + per, n_rows);
    const int64_t e0 = r0 * n_cols, e1 = r1 * n_cols;
    for (int64_t e = e0 + threadIdx.x; e < e1; e += NM_THREADS) {
        const int c = (int)(e % n_cols);
        const long long k = dbl_key(x[e]);
        atomicMin(&sh[2 * c], k);
        atomicMax(&sh[2 * c + 1], k);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * n_cols; i += NM_THREADS)
        part[(int64_t)blockIdx.x * 2 * n_cols + i] = sh[i];
}

__global__ void colminmax_final_kernel(const long long *__restrict__ part, int parts, int n_cols,
                                       double *__restrict__ mn, double *__restrict__ range)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    long long kmin = part[2 * c], kmax = part[2 * c + 1];
    for (int p = 1; p < parts; ++p) {
        kmin = min(kmin, part[(int64_t)p * 2 * n_cols + 2 * c]);
        kmax = max(kmax, part[(int64_t)p * 2 * n_cols + 2 * c + 1]);
    }
    const double m = key_dbl(kmin);
    mn[c] = m;
    range[c] = __dsub_rn(key_dbl(kmax), m);                 // == (data - min).max(0)
}

__global__ void __launch_bounds__(256)
normalize_apply_kernel(double *__restrict__ x, int64_t n_elems, int n_cols,
                       const double *__restrict__ mn, const double *__restrict__ range)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elems) return;
    const int c = (int)(e % n_cols);
    x[e] = __ddiv_rn(__dsub_rn(x[e], mn[c]), range[c]);
}

}  // namespace

extern "C" int ss_wave_minmax(const void *d_waves, int dtype, int n_win, int64_t n_spk, int n_c,
                              void *d_min, void *d_p2p, void *stream)
{
    if (n_spk == 0) return SS_OK;
    SS_REQUIRE(d_waves && (d_min || d_p2p), SS_ERR_ARG, "wave_minmax: null pointer");
    SS_REQUIRE(n_win >= 1 && n_spk > 0 && n_c >= 1, SS_ERR_ARG, "wave_minmax: bad shape");
    SS_REQUIRE(dtype == SS_F32 || dtype == SS_F64, SS_ERR_UNSUPPORTED,
               "wave_minmax: waveforms must be float32 or float64 (dtype %d)", dtype);
    const int64_t n_pairs = n_spk * n_c;
    const unsigned blocks = (unsigned)ss_cdiv(n_pairs, 256);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == SS_F32)
        wave_minmax_kernel<float><<<blocks, 256, 0, st>>>(reinterpret_cast<const float *>(d_waves), n_win, n_pairs,
                                                         reinterpret_cast<float *>(d_min), reinterpret_cast<float *>(d_p2p));
    else
        wave_minmax_kernel<double><<<blocks, 256, 0, st>>>(reinterpret_cast<const double *>(d_waves), n_win, n_pairs,
                                                          reinterpret_cast<double *>(d_min), reinterpret_cast<double *>(d_p2p));
    SS_LAUNCH_CHECK("wave_minmax_kernel");
    return SS_OK;
}

extern "C" size_t ss_normalize_workspace_bytes(int n_cols)
{
    return n_cols > 0 ? ((size_t)NM_PARTS * 2 * n_cols + 2 * (size_t)n_cols) * 8 : 0;
}

extern "C" int ss_normalize_columns(double *d_x, int64_t n_rows, int n_cols, void *d_ws,
                                    size_t ws_bytes, void *stream)
{
    if (n_rows == 0 || n_cols == 0) return SS_OK;
    SS_REQUIRE(d_x && d_ws, SS_ERR_ARG, "normalize: null pointer");
    SS_REQUIRE(n_rows > 0 && n_cols > 0 && n_cols <= 2048, SS_ERR_UNSUPPORTED,
               "normalize: %d columns outside [1, 2048]", n_cols);
    SS_REQUIRE(ws_bytes >= ss_normalize_workspace_bytes(n_cols), SS_ERR_WORKSPACE,
               "normalize: workspace too small");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    long long *part = reinterpret_cast<long long *>(d_ws);
    double *mn = reinterpret_cast<double *>(part + (size_t)NM_PARTS * 2 * n_cols);
    double *range = mn + n_cols;
    int parts = (int)(n_rows < NM_PARTS ? n_rows : NM_PARTS);
    colminmax_partial_kernel<<<parts, NM_THREADS, (size_t)2 * n_cols * 8, st>>>(d_x, n_rows, n_cols, part);
    SS_LAUNCH_CHECK("colminmax_partial_kernel");
    colminmax_final_kernel<<<(n_cols + 127) / 128, 128, 0, st>>>(part, parts, n_cols, mn, range);
    SS_LAUNCH_CHECK("colminmax_final_kernel");
    const int64_t n_elems = n_rows * n_cols;
    normalize_apply_kernel<<<(unsigned)ss_cdiv(n_elems, 256), 256, 0, st>>>(d_x, n_elems, n_cols, mn, range);
    SS_LAUNCH_CHECK("normalize_apply_kernel");
    return SS_OK;
}
