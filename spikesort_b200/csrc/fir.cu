// "Next" row, SURVEY.md 8f rank 4: zero-phase FIR filter (FilterFir).
//
// Replaces FilterFir.__call__ -> scipy.signal.filtfilt(b, [1], x) (reference
// src/spike_sort/core/filters.py:41-74) inside the (contact, chunk) loop of filter_proxy
// (filters.py:135-141).
//
// filtfilt pads every chunk with 3*K samples of odd extension (K = len(b)) and starts both
// passes from the steady state of a constant input.  For an FIR filter the state only
// reaches K-1 samples back, so inside the part of the output that filtfilt keeps neither
// initial condition is visible: with ext = odd extension (in the INPUT dtype, as NumPy
// computes it; int16 wraps around),
//     y1[m] = sum_k b[k] ext[m-k],   out[o] = sum_k b[k] y1[o + 3K + k],
// and every index touched lies inside ext.  No recurrence: a two-stage stencil.  One CTA
// produces FIR_T outputs of one (contact, chunk) unit: the ext window (FIR_T + 2K - 2
// values) is staged in shared memory as float64, y1 over FIR_T + K - 1 positions goes to
// shared memory, then the backward stage; each thread keeps FIR_R consecutive outputs in
// registers so that every shared-memory load feeds FIR_R FMAs.  HBM traffic is the
// algorithmic minimum, s_in + 8 bytes per sample; the kernel is FP64-bound (2K FMA/sample).
#include "common.cuh"

namespace {

constexpr int FIR_T = 1024;
constexpr int FIR_THREADS = 256;
constexpr int FIR_R = 4;
constexpr int FIR_MAX_TAPS = 2048;

// ext[e] of the chunk x[0..L): scipy.signal._arraytools.odd_ext with n = pad
template <typename T>
__device__ __forceinline__ double fir_ext(const T *__restrict__ x, int64_t L, int64_t pad, int64_t e);

template <>
__device__ __forceinline__ double fir_ext<int16_t>(const int16_t *__restrict__ x, int64_t L, int64_t pad, int64_t e)
{
    if (e < pad) return (double)(int16_t)((int)(int16_t)(2 * (int)__ldg(x)) - (int)__ldg(x + (pad - e)));
    if (e < pad + L) return (double)__ldg(x + (e - pad));
    return (double)(int16_t)((int)(int16_t)(2 * (int)__ldg(x + L - 1)) - (int)__ldg(x + (L - 2 - (e - pad - L))));
}
template <>
__device__ __forceinline__ double fir_ext<float>(const float *__restrict__ x, int64_t L, int64_t pad, int64_t e)
{
    if (e < pad) return (double)__fsub_rn(__fmul_rn(2.0f, __ldg(x)), __ldg(x + (pad - e)));
    if (e < pad + L) return (double)__ldg(x + (e - pad));
    return (double)__fsub_rn(__fmul_rn(2.0f, __ldg(x + L - 1)), __ldg(x + (L - 2 - (e - pad - L))));
}
template <>
__device__ __forceinline__ double fir_ext<double>(const double *__restrict__ x, int64_t L, int64_t pad, int64_t e)
{
    if (e < pad) return __dsub_rn(__dmul_rn(2.0, __ldg(x)), __ldg(x + (pad - e)));
    if (e < pad + L) return __ldg(x + (e - pad));
    return __dsub_rn(__dmul_rn(2.0, __ldg(x + L - 1)), __ldg(x + (L - 2 - (e - pad - L))));
}

// dst[i] = sum_k b[k] * src[i + off(k)] for i in [0, n): off(k) = K-1-k (forward stage,
// causal) or k (backward stage, anti-causal); FIR_R consecutive i per thread
template <bool FORWARD>
__device__ __forceinline__ void fir_stage(const double *__restrict__ bs, int K,
                                          const double *__restrict__ src, double *dst, int n)
{
    for (int i0 = threadIdx.x * FIR_R; i0 < n; i0 += FIR_THREADS * FIR_R) {
        double acc[FIR_R], win[FIR_R];
#pragma unroll
        for (int r = 0; r < FIR_R; ++r) acc[r] = 0.0;
        if (FORWARD) {
            // tap k multiplies src[i + K-1-k]: walk k downwards so the window slides forward
#pragma unroll
            for (int r = 0; r < FIR_R - 1; ++r) win[r + 1] = src[i0 + r];
            for (int k = K - 1; k >= 0; --k) {
#pragma unroll
                for (int r = 0; r < FIR_R - 1; ++r) win[r] = win[r + 1];
                win[FIR_R - 1] = src[i0 + FIR_R - 1 + (K - 1 - k)];
                const double bk = bs[k];
#pragma unroll
                for (int r = 0; r < FIR_R; ++r) acc[r] = fma(bk, win[r], acc[r]);
            }
        } else {
#pragma unroll
            for (int r = 0; r < FIR_R - 1; ++r) win[r + 1] = src[i0 + r];
            for (int k = 0; k < K; ++k) {
#pragma unroll
                for (int r = 0; r < FIR_R - 1; ++r) win[r] = win[r + 1];
                win[FIR_R - 1] = src[i0 + FIR_R - 1 + k];
                const double bk = bs[k];
#pragma unroll
                for (int r = 0; r < FIR_R; ++r) acc[r] = fma(bk, win[r], acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < FIR_R; ++r)
            if (i0 + r < n) dst[i0 + r] = acc[r];
    }
}

// grid (tiles per chunk, chunks, contacts)
template <typename T>
__global__ void __launch_bounds__(FIR_THREADS)
fir_filtfilt_kernel(const T *__restrict__ x, int64_t n_samples, int64_t ld_in, int64_t chunksize,
                    const double *__restrict__ b, int K, double *__restrict__ out, int64_t ld_out)
{
    extern __shared__ __align__(16) double sm[];
    const int n_es = FIR_T + 2 * K - 2 + FIR_R, n_ys = FIR_T + K - 1 + FIR_R;   // + slack for the register window
    double *bs = sm, *es = bs + K, *ys = es + n_es, *os = ys + n_ys;
    const int64_t cs0 = (int64_t)blockIdx.y * chunksize;
    const int64_t L = min(chunksize, n_samples - cs0);
    const int64_t o0 = (int64_t)blockIdx.x * FIR_T;
    if (o0 >= L) return;
    const int n_out = (int)min((int64_t)FIR_T, L - o0);
    const int64_t pad = 3 * (int64_t)K, Lx = L + 2 * pad;
    const T *row = x + (int64_t)blockIdx.z * ld_in + cs0;
    for (int i = threadIdx.x; i < K; i += FIR_THREADS) bs[i] = b[i];
    const int64_t e_lo = o0 + pad - (K - 1);                 // >= 2K + 1 > 0
    for (int i = threadIdx.x; i < n_es; i += FIR_THREADS) {
        const int64_t e = e_lo + i;
        es[i] = e < Lx ? fir_ext<T>(row, L, pad, e) : 0.0;   // beyond Lx only feeds unused slack
    }
    __syncthreads();
    fir_stage<true>(bs, K, es, ys, n_out + K - 1);           // y1 at ext index o0 + pad + i
    for (int i = n_out + K - 1 + threadIdx.x; i < n_ys; i += FIR_THREADS) ys[i] = 0.0;
    __syncthreads();
    fir_stage<false>(bs, K, ys, os, n_out);
    __syncthreads();
    double *orow = out + (int64_t)blockIdx.z * ld_out + cs0 + o0;
    for (int i = threadIdx.x; i < n_out; i += FIR_THREADS) __stcs(orow + i, os[i]);
}

}  // namespace

extern "C" int ss_filtfilt_fir(const void *d_x, int dtype, int64_t n_contacts, int64_t n_samples,
                               int64_t ld_in, const double *d_b, int n_taps, int64_t chunksize,
                               double *d_out, int64_t ld_out, void *stream)
{
    SS_REQUIRE(d_x && d_b && d_out, SS_ERR_ARG, "filtfilt_fir: null pointer");
    SS_REQUIRE(dtype >= SS_I16 && dtype <= SS_F64, SS_ERR_ARG, "filtfilt_fir: bad dtype %d", dtype);
    SS_REQUIRE(n_contacts >= 1 && n_contacts <= 65535 && n_samples >= 1 && chunksize >= 1, SS_ERR_ARG,
               "filtfilt_fir: bad shape");
    SS_REQUIRE(n_taps >= 1 && n_taps <= FIR_MAX_TAPS, SS_ERR_UNSUPPORTED,
               "filtfilt_fir: %d taps outside [1, %d]", n_taps, FIR_MAX_TAPS);
    const int64_t n_chunks = ss_cdiv(n_samples, chunksize);
    const int64_t len_last = n_samples - (n_chunks - 1) * chunksize;
    const int64_t shortest = n_chunks > 1 && chunksize < len_last ? chunksize : len_last;
    SS_REQUIRE(shortest > 3 * (int64_t)n_taps, SS_ERR_PADLEN,
               "The length of the input vector x must be greater than padlen, which is %d.", 3 * n_taps);
    SS_REQUIRE(n_chunks <= 65535, SS_ERR_ARG, "filtfilt_fir: too many chunks");
    const int64_t longest = n_chunks > 1 ? chunksize : len_last;
    const dim3 grid((unsigned)ss_cdiv(longest, FIR_T), (unsigned)n_chunks, (unsigned)n_contacts);
    const size_t smem = ((size_t)n_taps + (FIR_T + 2 * n_taps - 2 + FIR_R) + (FIR_T + n_taps - 1 + FIR_R) +
                         FIR_T) * sizeof(double);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define SS_FIR_LAUNCH(T)                                                                          \
    do {                                                                                          \
        SS_CUDA_CHECK(cudaFuncSetAttribute(fir_filtfilt_kernel<T>,                                \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        fir_filtfilt_kernel<T><<<grid, FIR_THREADS, smem, st>>>(                                  \
            reinterpret_cast<const T *>(d_x), n_samples, ld_in, chunksize, d_b, n_taps, d_out, ld_out); \
    } while (0)
    if (dtype == SS_I16) SS_FIR_LAUNCH(int16_t);
    else if (dtype == SS_F32) SS_FIR_LAUNCH(float);
    else SS_FIR_LAUNCH(double);
#undef SS_FIR_LAUNCH
    SS_LAUNCH_CHECK("fir_filtfilt_kernel");
    return SS_OK;
}
