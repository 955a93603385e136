// Shared helpers for the spikesort_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include "spikesort_b200.h"

void ss_set_error(const char *fmt, ...);
void ss_note_launch(void);        // one more kernel of the library launched (ss_launch_count)

#define SS_CUDA_CHECK(expr)                                                          \
    do {                                                                             \
        cudaError_t e__ = (expr);                                                    \
        if (e__ != cudaSuccess) {                                                    \
            ss_set_error("%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                         __LINE__);                                                  \
            return SS_ERR_CUDA;                                                      \
        }                                                                            \
    } while (0)

#define SS_LAUNCH_CHECK(name)                                                        \
    do {                                                                             \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess) {                                                    \
            ss_set_error("launch of %s: %s", name, cudaGetErrorString(e__));         \
            return SS_ERR_CUDA;                                                      \
        }                                                                            \
        ss_note_launch();                                                            \
    } while (0)

#define SS_REQUIRE(cond, code, ...)                                                  \
    do {                                                                             \
        if (!(cond)) {                                                               \
            ss_set_error(__VA_ARGS__);                                               \
            return code;                                                             \
        }                                                                            \
    } while (0)

static inline size_t ss_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int64_t ss_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
__device__ __forceinline__ int64_t ss_cdiv_dev(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr unsigned SS_FULL = 0xffffffffu;
constexpr int SS_NUM_SMS = 148;   // B200

// One recording element as float64 (exact for all three dtypes).
__device__ __forceinline__ double ss_load(const void *base, int dtype, int64_t i)
{
    if (dtype == SS_I16) return (double)__ldg(reinterpret_cast<const int16_t *>(base) + i);
    if (dtype == SS_F32) return (double)__ldg(reinterpret_cast<const float *>(base) + i);
    return __ldg(reinterpret_cast<const double *>(base) + i);
}

// The same element rounded to float32 (what extract_spikes stores, extract.py:157-163).
__device__ __forceinline__ float ss_load_f32(const void *base, int dtype, int64_t i)
{
    if (dtype == SS_I16) return (float)__ldg(reinterpret_cast<const int16_t *>(base) + i);
    if (dtype == SS_F32) return __ldg(reinterpret_cast<const float *>(base) + i);
    return __double2float_rn(__ldg(reinterpret_cast<const double *>(base) + i));
}

__device__ __forceinline__ double ss_shfl_up(double v, int d)
{
    return __shfl_up_sync(SS_FULL, v, d);
}
__device__ __forceinline__ double ss_shfl_down(double v, int d)
{
    return __shfl_down_sync(SS_FULL, v, d);
}
__device__ __forceinline__ double ss_shfl(double v, int src)
{
    return __shfl_sync(SS_FULL, v, src);
}

// Exact element -> float64 conversions for the streaming kernels.  int16 goes through the
// 2^52 trick (one integer op + one full-rate DADD) because I2F.F64 runs at a quarter of
// the FP64 rate: bits(2^52 + 2^31 + v) = 0x43300000'80000000 ^ (uint32)v  for |v| < 2^31.
__device__ __forceinline__ double ss_to_f64(int16_t v)
{
    const double m = __hiloint2double(0x43300000, (int)(0x80000000u ^ (unsigned)(int)v));
    return m - 4503601774854144.0;                       // 2^52 + 2^31
}
__device__ __forceinline__ double ss_to_f64(float v) { return (double)v; }
__device__ __forceinline__ double ss_to_f64(double v) { return v; }

// Time shards: a window sample at local index i that lies beyond the neighbour samples stored
// around the shard (but inside the recording: a side without a neighbour has halo 0) cannot be
// read; the kernels substitute 0 and count the event so that the host can redo the call
// with a wider halo (the reference, one process, would simply read the sample).
__device__ __forceinline__ bool ss_beyond_halo(int64_t i, int64_t n_samples, int64_t halo_before,
                                               int64_t halo_after)
{
    return (halo_before > 0 && i < -halo_before) || (halo_after > 0 && i >= n_samples + halo_after);
}

static inline int ss_elem_size(int dtype) { return dtype == SS_I16 ? 2 : dtype == SS_F32 ? 4 : 8; }
