// Error reporting and version of the spikesort_b200 C ABI.
#include "common.cuh"
#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}

void ss_note_launch(void) { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" int64_t ss_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

void ss_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int ss_version(void) { return SS_VERSION; }
extern "C" const char *ss_last_error(void) { return g_err; }

// Strided (contacts x time-range) transfer between a pinned host recording and HBM in ONE
// asynchronous DMA: what filter_proxy's per-chunk slicing data[i, j*cs:stop]
// (filters.py:139-141) becomes when a time slab of every contact is uploaded at once.
extern "C" int ss_copy_2d_async(void *dst, size_t dst_pitch, const void *src, size_t src_pitch,
                                size_t width_bytes, size_t rows, int to_device, void *stream)
{
    SS_REQUIRE(dst && src, SS_ERR_ARG, "copy_2d: null pointer");
    SS_REQUIRE(dst_pitch >= width_bytes && src_pitch >= width_bytes, SS_ERR_ARG, "copy_2d: pitch < width");
    if (width_bytes == 0 || rows == 0) return SS_OK;
    SS_CUDA_CHECK(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, rows,
                                    to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost,
                                    reinterpret_cast<cudaStream_t>(stream)));
    return SS_OK;
}
