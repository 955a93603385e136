// Cross-file internals of libspikesort_b200 (not part of the C ABI).
#pragma once
#include "common.cuh"

namespace ssi {

// layout of the detection workspace (shared by detect.cu and the fused filter path)
constexpr int DT_WARPS = 8;
constexpr int DT_WORDS_PER_WARP = 32;                        // 32 words x 32 samples
constexpr int DT_BLOCK_WORDS = DT_WARPS * DT_WORDS_PER_WARP; // 256 words
constexpr int DT_BLOCK_SAMPLES = DT_BLOCK_WORDS * 32;        // 8192 samples

struct DetLayout {
    int64_t words_per_row, blocks_per_row, n_blocks;
    size_t off_mask, off_counts, off_offsets, off_scratch, total;
};

constexpr int TH_PARTS = 64;                                  // partial sums per row (threshold)

inline DetLayout det_layout(int64_t n_rows, int64_t n_samples)
{
    DetLayout L;
    L.blocks_per_row = ss_cdiv(n_samples, DT_BLOCK_SAMPLES);
    L.words_per_row = L.blocks_per_row * DT_BLOCK_WORDS;
    L.n_blocks = L.blocks_per_row * n_rows;
    L.off_mask = 0;
    L.off_counts = ss_align_up((size_t)L.words_per_row * n_rows * 4, 256);
    L.off_offsets = ss_align_up(L.off_counts + (size_t)L.n_blocks * 4, 256);
    L.off_scratch = ss_align_up(L.off_offsets + (size_t)(L.n_blocks + 1) * 8, 256);
    L.total = ss_align_up(L.off_scratch + (size_t)n_rows * TH_PARTS * 2 * sizeof(double), 256);
    return L;
}

// crossing rule of extract.py:86-91; comparisons in float64
__device__ __forceinline__ bool crossing(double a, double b, double thr, int falling)
{
    return falling ? (a > thr && b < thr) : (a < thr && b > thr);
}

// detect.cu: mark crossings of the first n_valid samples of every row into a mask
// laid out for rows of n_samples (plain word stores; counts are NOT final)
int mark_head(const void *d_x, int dtype, int64_t n_rows, int64_t n_valid, int64_t n_samples,
              int64_t ld, const double *d_thr, int falling, void *d_ws, cudaStream_t st);
// detect.cu: counts per block from the mask (popc) + exclusive scan -> offsets
int count_and_scan(int64_t n_rows, int64_t n_samples, void *d_ws, int64_t *d_offsets,
                   cudaStream_t st);

// pca_tc.cu: Gram matrices on the tensor cores (tcgen05, 3xTF32); partials in the same
// workspace layout as the float64 kernel of pca.cu
bool pca_gram_tc_eligible(int n_win, int64_t n_spk, int n_c, bool offsets_mode);
int pca_gram_tc(const float *d_waves, int n_win, int64_t n_spk, int n_c, const int64_t *d_offsets,
                int64_t n_groups, const double *d_ref, double *ws_gram, double *ws_sum, int parts,
                int nwp, cudaStream_t st);

}  // namespace ssi
