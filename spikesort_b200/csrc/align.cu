// K4 - peak alignment of spike times and duplicate removal.
//
// Replaces, in reference src/spike_sort/core/extract.py:
//   :235-265  the while-loop of align_spikes (resample == 1)
//   :277-288  remove_doubles
//
// Alignment: every spike is independent (the reference's "queue" is only a way
// to vectorise), so one thread walks one spike to its fixed point: it scans the
// sp_win neighbourhood of the current position, rounded to float32 as
// extract_spikes does (extract.py:157-163); the strict comparison in scan order is
// NumPy's first-occurrence arg-max/min.  Every
// float64 expression is evaluated in the reference's order with explicit
// IEEE round-to-nearest intrinsics (no FMA contraction), the sample index by a
// truncating cast, so the spike times are bit-identical.
#include "common.cuh"
#include "scan.cuh"

namespace {

constexpr int AL_WARPS = 8;
constexpr int RD_BLOCK = 1024;          // elements per block in remove_doubles
constexpr int AL_NONE = 0x7fffffff;

__device__ __forceinline__ int64_t find_row(const int64_t *__restrict__ offsets, int64_t n_rows,
                                            int64_t s)
{
    int64_t lo = 0, hi = n_rows;        // largest r < n_rows with offsets[r] <= s
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid) <= s) lo = mid; else hi = mid;
    }
    return lo;
}

// One THREAD per spike.  (A warp per spike - lanes along the window, shuffle arg-max - made
// every lane repeat the scalar float64 work: two IEEE divisions and the index conversion per
// pass; ncu showed the kernel issue-bound on exactly that, 70 us for 166 k spikes.  Here the
// scalar work is done once per spike, the n_win loads of a thread are independent and hit the
// 2-3 cache lines of its window, and the first extremum falls out of the scan order.)
__global__ void __launch_bounds__(32 * AL_WARPS)
align_kernel(const void *__restrict__ x, int dtype, int64_t n_samples, int64_t ld, double FS,
             const int32_t *__restrict__ rows, int64_t n_rows, const int64_t *__restrict__ offsets,
             double sp_win0, int win0, int win1, double t_min, double t_max, double lo, double hi,
             int use_min, double *__restrict__ spt, int64_t n_spk, int max_iter,
             int64_t index_offset, int64_t halo_before, int64_t halo_after,
             int32_t *__restrict__ halo_miss)
{
    const int64_t s = (int64_t)blockIdx.x * (32 * AL_WARPS) + threadIdx.x;
    if (s >= n_spk || (offsets && s >= __ldg(offsets + n_rows))) return;   // spare capacity
    int64_t r = n_rows > 1 ? find_row(offsets, n_rows, s) : 0;
    if (rows) r = rows[r];
    const int64_t row0 = r * ld;
    const int n_win = win1 - win0;
    double t = spt[s];
    bool moved = false, missed = false;
    for (int it = 0; it < max_iter; ++it) {
        if (!(t >= t_min && t <= t_max)) break;                 // filter_spt, extract.py:108-110
        const int centre = __double2int_rz(__dmul_rn(__ddiv_rn(t, 1000.0), FS));   // :150
        const int64_t first = (int64_t)centre + win0 - index_offset;   // local index
        float best = 0.f;
        int best_k = 0;
        if (first >= -halo_before && first + n_win <= n_samples + halo_after) {
            // whole window inside the stored samples (all spikes but those at the very ends)
            const int64_t base = row0 + first;
            if (dtype == SS_F64) {
                const double *p = reinterpret_cast<const double *>(x) + base;
                for (int k = 0; k < n_win; ++k) {
                    float v = __double2float_rn(__ldg(p + k));
                    if (use_min) v = -v;                        // argmin(v) == argmax(-v), ties alike
                    if (k == 0 || v > best) { best = v; best_k = k; }
                }
            } else {
                for (int k = 0; k < n_win; ++k) {
                    float v = ss_load_f32(x, dtype, base + k);
                    if (use_min) v = -v;
                    if (k == 0 || v > best) { best = v; best_k = k; }
                }
            }
        } else {
            for (int k = 0; k < n_win; ++k) {
                const int64_t i = first + k;
                float v = 0.f;
                if (i >= -halo_before && i < n_samples + halo_after) v = ss_load_f32(x, dtype, row0 + i);
                else missed = missed || ss_beyond_halo(i, n_samples, halo_before, halo_after);
                if (use_min) v = -v;
                if (k == 0 || v > best) { best = v; best_k = k; }
            }
        }
        // time[k] = k*1000.0/FS + sp_win[0]  (extract.py:152), spt += shift (:258)
        const double shift = __dadd_rn(__ddiv_rn(__dmul_rn((double)best_k, 1000.0), FS), sp_win0);
        t = __dadd_rn(t, shift);
        moved = true;
        if (!(shift < lo || shift > hi)) break;                 // :263-264
    }
    if (moved) spt[s] = t;
    if (missed && halo_miss) atomicAdd(halo_miss, 1);
}

// keep[k] = first of its row, or (int64)((t[k]-t[k-1])/tol) > 1
__global__ void __launch_bounds__(RD_BLOCK)
rd_mark_kernel(const double *__restrict__ spt, int64_t n_spk, const int64_t *__restrict__ offsets,
               int64_t n_rows, double tol, const double *__restrict__ prev_last,
               uint8_t *__restrict__ keep, int32_t *__restrict__ counts)
{
    __shared__ int warp_cnt[RD_BLOCK / 32];
    const int64_t k = (int64_t)blockIdx.x * RD_BLOCK + threadIdx.x;
    bool kp = false;
    if (k < n_spk && k >= __ldg(offsets + n_rows)) keep[k] = 0;            // spare capacity
    if (k < n_spk && k < __ldg(offsets + n_rows)) {
        const int64_t r = n_rows > 1 ? find_row(offsets, n_rows, k) : 0;
        if (k == __ldg(offsets + r)) {
            // first of its row: kept, unless the previous time shard ended within tol
            // (prev_last[r] = last aligned time of the row there, NaN if none)
            kp = true;
            if (prev_last) {
                const double pl = prev_last[r];
                if (pl == pl) kp = __double2ll_rz(__ddiv_rn(__dsub_rn(spt[k], pl), tol)) > 1;
            }
        } else {
            const double isi = __dsub_rn(spt[k], spt[k - 1]);
            kp = __double2ll_rz(__ddiv_rn(isi, tol)) > 1;
        }
        keep[k] = kp ? 1 : 0;
    }
    const unsigned m = __ballot_sync(SS_FULL, kp);
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < RD_BLOCK / 32; ++w) s += warp_cnt[w];
        counts[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(RD_BLOCK)
rd_emit_kernel(const double *__restrict__ spt, int64_t n_spk, const uint8_t *__restrict__ keep,
               const int64_t *__restrict__ block_off, double *__restrict__ out, int64_t cap)
{
    __shared__ int warp_cnt[RD_BLOCK / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t k = (int64_t)blockIdx.x * RD_BLOCK + threadIdx.x;
    const bool kp = k < n_spk && keep[k];
    const unsigned m = __ballot_sync(SS_FULL, kp);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += warp_cnt[w];
    if (kp) {
        const int64_t pos = block_off[blockIdx.x] + base + __popc(m & ((1u << lane) - 1u));
        if (pos < cap) out[pos] = spt[k];
    }
}

// new row offsets: kept elements before each old row start (one warp per row)
__global__ void rd_rows_kernel(const uint8_t *__restrict__ keep,
                               const int64_t *__restrict__ block_off,
                               const int64_t *__restrict__ offsets, int64_t n_rows,
                               int64_t n_blocks, int64_t n_spk, int64_t *__restrict__ offsets_out)
{
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r > n_rows) return;
    const int64_t e = offsets[r];
    int64_t cnt;
    if (e >= n_spk) {
        cnt = block_off[n_blocks];
    } else {
        const int64_t blk = e / RD_BLOCK;
        int c = 0;
        for (int64_t i = blk * RD_BLOCK + lane; i < e; i += 32) c += keep[i];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(SS_FULL, c, d);
        cnt = block_off[blk] + c;
    }
    if (lane == 0) offsets_out[r] = cnt;
}

struct RdLayout {
    int64_t n_blocks;
    size_t off_keep, off_counts, off_offsets, total;
};

RdLayout rd_layout(int64_t n_spk)
{
    RdLayout L;
    const int64_t n = n_spk > 0 ? n_spk : 1;
    L.n_blocks = ss_cdiv(n, RD_BLOCK);
    L.off_keep = 0;
    L.off_counts = ss_align_up((size_t)n, 256);
    L.off_offsets = ss_align_up(L.off_counts + (size_t)L.n_blocks * 4, 256);
    L.total = ss_align_up(L.off_offsets + (size_t)(L.n_blocks + 1) * 8, 256);
    return L;
}

}  // namespace

extern "C" int ss_align(const void *d_x, int dtype, int64_t n_samples, int64_t ld, double FS,
                        const int32_t *d_rows, int64_t n_rows, const int64_t *d_offsets,
                        double sp_win0, int win0, int win1, double t_min, double t_max, double lo,
                        double hi, int use_min, double *d_spt, int64_t n_spk, int max_iter,
                        int64_t index_offset, int64_t halo_before, int64_t halo_after,
                        int32_t *d_halo_miss,
                        void *stream)
{
    if (n_spk == 0) return SS_OK;
    SS_REQUIRE(d_x && d_spt, SS_ERR_ARG, "align: null pointer");
    SS_REQUIRE(dtype >= SS_I16 && dtype <= SS_F64, SS_ERR_ARG, "align: bad dtype %d", dtype);
    SS_REQUIRE(n_rows >= 1 && (n_rows == 1 || d_offsets), SS_ERR_ARG,
               "align: offsets required for %lld rows", (long long)n_rows);
    SS_REQUIRE(win1 > win0, SS_ERR_ARG, "align: empty window [%d, %d)", win0, win1);
    SS_REQUIRE(win1 - win0 <= ss_max_window(), SS_ERR_UNSUPPORTED,
               "align: window of %d samples > %d", win1 - win0, ss_max_window());
    SS_REQUIRE(n_spk > 0 && max_iter > 0, SS_ERR_ARG, "align: bad n_spk/max_iter");
    SS_REQUIRE(halo_before >= 0 && halo_after >= 0, SS_ERR_ARG, "align: negative halo");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    align_kernel<<<(unsigned)ss_cdiv(n_spk, 32 * AL_WARPS), 32 * AL_WARPS, 0, st>>>(
        d_x, dtype, n_samples, ld, FS, d_rows, n_rows, d_offsets, sp_win0, win0, win1, t_min,
        t_max, lo, hi, use_min, d_spt, n_spk, max_iter, index_offset, halo_before, halo_after,
        d_halo_miss);
    SS_LAUNCH_CHECK("align_kernel");
    return SS_OK;
}

extern "C" size_t ss_remove_doubles_workspace_bytes(int64_t n_spk)
{
    return rd_layout(n_spk).total;
}

extern "C" int ss_remove_doubles_mark(const double *d_spt, int64_t n_spk, const int64_t *d_offsets,
                                      int64_t n_rows, double tol, const double *d_prev_last,
                                      void *d_ws, size_t ws_bytes, int64_t *d_offsets_out,
                                      void *stream)
{
    SS_REQUIRE(d_offsets && d_offsets_out && d_ws, SS_ERR_ARG, "remove_doubles: null pointer");
    SS_REQUIRE(n_spk >= 0 && n_rows >= 1, SS_ERR_ARG, "remove_doubles: bad shape");
    const RdLayout L = rd_layout(n_spk);
    SS_REQUIRE(ws_bytes >= L.total, SS_ERR_WORKSPACE, "remove_doubles: workspace too small");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    char *ws = reinterpret_cast<char *>(d_ws);
    uint8_t *keep = reinterpret_cast<uint8_t *>(ws + L.off_keep);
    int32_t *counts = reinterpret_cast<int32_t *>(ws + L.off_counts);
    int64_t *boff = reinterpret_cast<int64_t *>(ws + L.off_offsets);
    rd_mark_kernel<<<(unsigned)L.n_blocks, RD_BLOCK, 0, st>>>(d_spt, n_spk, d_offsets, n_rows, tol,
                                                             d_prev_last, keep, counts);
    SS_LAUNCH_CHECK("rd_mark_kernel");
    ssb::scan_counts_kernel<<<1, ssb::SCAN_THREADS, 0, st>>>(counts, L.n_blocks, boff, 0, 0,
                                                             nullptr);
    SS_LAUNCH_CHECK("scan_counts_kernel");
    rd_rows_kernel<<<(unsigned)ss_cdiv(n_rows + 1, 4), 128, 0, st>>>(keep, boff, d_offsets, n_rows,
                                                                    L.n_blocks, n_spk,
                                                                    d_offsets_out);
    SS_LAUNCH_CHECK("rd_rows_kernel");
    return SS_OK;
}

extern "C" int ss_remove_doubles_emit(const double *d_spt, int64_t n_spk, const void *d_ws,
                                      size_t ws_bytes, double *d_out, int64_t cap, void *stream)
{
    if (n_spk == 0 || cap <= 0) return SS_OK;
    SS_REQUIRE(d_spt && d_ws && d_out, SS_ERR_ARG, "remove_doubles_emit: null pointer");
    const RdLayout L = rd_layout(n_spk);
    SS_REQUIRE(ws_bytes >= L.total, SS_ERR_WORKSPACE, "remove_doubles_emit: workspace too small");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const char *ws = reinterpret_cast<const char *>(d_ws);
    rd_emit_kernel<<<(unsigned)L.n_blocks, RD_BLOCK, 0, st>>>(
        d_spt, n_spk, reinterpret_cast<const uint8_t *>(ws + L.off_keep),
        reinterpret_cast<const int64_t *>(ws + L.off_offsets), d_out, cap);
    SS_LAUNCH_CHECK("rd_emit_kernel");
    return SS_OK;
}
