// Read-bandwidth probe for the access pattern of filt_summary_kernel (sweep 1): every warp walks
// its own contiguous 32 KB (32 tiles of 1 KB) in groups of 4 KB, 16 warps per SM, either with
// 32-bit loads (lane l takes word 32 k + l: 128 B per request) or with 128-bit loads (512 B per
// request); the loaded words are only XOR-ed.  nvcc -arch=sm_100a -O3 -o read_pattern read_pattern.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int VEC, int R>
__global__ void __launch_bounds__(256, 2) reader(const uint32_t *__restrict__ x, size_t n_words, uint32_t *out, int tpw)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t first = ((size_t)blockIdx.x * 8 + warp) * tpw * 256;      // 256 words per tile
    uint32_t acc = 0;
    for (int it = 0; it < tpw; it += R) {
        const size_t base = first + (size_t)it * 256;
        if (base + R * 256 > n_words) break;
        if (VEC == 1) {
            uint32_t raw[R * 8];
#pragma unroll
            for (int i = 0; i < R * 8; ++i) raw[i] = __ldg(x + base + 32 * i + lane);
#pragma unroll
            for (int i = 0; i < R * 8; ++i) acc ^= raw[i];
        } else {
            uint4 raw[R * 2];
#pragma unroll
            for (int i = 0; i < R * 2; ++i) raw[i] = __ldg(reinterpret_cast<const uint4 *>(x + base) + 32 * i + lane);
#pragma unroll
            for (int i = 0; i < R * 2; ++i) acc ^= raw[i].x ^ raw[i].y ^ raw[i].z ^ raw[i].w;
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

int main()
{
    const size_t bytes = 1152000000ull, n_words = bytes / 4;
    uint32_t *x, *out, *flush;
    cudaMalloc(&x, bytes); cudaMalloc(&out, 4); cudaMalloc(&flush, 512u << 20);
    cudaMemset(x, 1, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int tpw = 32;
    const unsigned blocks = (unsigned)((n_words / 256 + 8 * tpw - 1) / (8 * tpw));
    for (int variant = 0; variant < 4; ++variant) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaMemsetAsync(flush, rep, 512u << 20);
            cudaEventRecord(e0);
            if (variant == 0) reader<1, 4><<<blocks, 256>>>(x, n_words, out, tpw);
            if (variant == 1) reader<4, 4><<<blocks, 256>>>(x, n_words, out, tpw);
            if (variant == 2) reader<1, 8><<<blocks, 256>>>(x, n_words, out, tpw);
            if (variant == 3) reader<4, 8><<<blocks, 256>>>(x, n_words, out, tpw);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        const char *names[4] = {"LDG.32  x 32 per group of 4 tiles", "LDG.128 x 8  per group of 4 tiles",
                                "LDG.32  x 64 per group of 8 tiles", "LDG.128 x 16 per group of 8 tiles"};
        printf("%s: %.1f us  %.2f TB/s (%s)\n", names[variant], best * 1e3, bytes / (best * 1e-3) / 1e12,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
