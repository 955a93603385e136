// Exchange steps of the time-sharded chain over peer memory (NVLink 5 / NVSwitch).
//
// A recording sharded by time in multiples of `chunksize` needs no communication for the
// filter (reference filters.py:137-141: every (contact, chunk) is independent).  What does
// cross a shard boundary is tiny - thresholds, 4*n_win filtered samples per
// contact, one spike time per contact, the per-contact Gram partials - so the cost of an
// exchange is latency, not bandwidth.  Every rank therefore owns a small ARENA in
// peer-addressable memory (torch symmetric memory, cudaIpc or plain device memory when all
// shards live on one GPU); the kernels below STORE their payload straight into the
// neighbours' arenas over NVLink and a one-CTA barrier kernel (release/acquire flags with
// monotonically increasing epochs, bounded spin) orders producer and consumer.  There is
// no NCCL call, no staging copy and no host synchronisation on this path.
//
// Arena layout (ss_shard_arena_offsets), all float64 unless noted:
//   flags      uint64[SS_MAX_PEERS]   epoch written by every peer
//   thr        [n_c]                  thresholds (from the rank owning the first 10 s)
//   from_left  [n_c][H]               last H filtered samples of the left neighbour
//   from_right [n_c][H]               first H filtered samples of the right neighbour
//   prev_last  [n_c]                  last aligned spike time of the left neighbour (NaN: none)
//   slots      [world][n_c*n_win^2 + n_c*n_win + n_c + 1]   Gram | sums | counts | status of every rank
#include "common.cuh"
#include "internal.cuh"

namespace {

constexpr int XC_THREADS = 256;

struct PeerPtrs { void *p[SS_MAX_PEERS]; };

struct ArenaLayout { size_t flags, thr, from_left, from_right, prev_last, slots, slot_stride, total; };

ArenaLayout arena_layout(int64_t n_c, int64_t H, int64_t n_win, int64_t world)
{
    ArenaLayout L;
    size_t o = 0;
    L.flags = o;       o = ss_align_up(o + SS_MAX_PEERS * sizeof(uint64_t), 256);
    L.thr = o;         o = ss_align_up(o + (size_t)n_c * 8, 256);
    L.from_left = o;   o = ss_align_up(o + (size_t)n_c * H * 8, 256);
    L.from_right = o;  o = ss_align_up(o + (size_t)n_c * H * 8, 256);
    L.prev_last = o;   o = ss_align_up(o + (size_t)n_c * 8, 256);
    L.slot_stride = ss_align_up(((size_t)n_c * (n_win * n_win + n_win + 1) + 1) * 8, 256);   // + status
    L.slots = o;       o += L.slot_stride * (size_t)world;
    L.total = o;
    return L;
}

__device__ __forceinline__ void st_release_sys(uint64_t *p, uint64_t v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t *p)
{
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t global_ns()
{
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Thread i: tell peer i that this rank reached `epoch`, then wait until peer i did.
// Kernel boundaries order the earlier peer stores of this stream before the release store.
__global__ void peer_barrier_kernel(PeerPtrs flags, int rank, int world, uint64_t epoch,
                                    uint64_t timeout_ns)
{
    const int i = threadIdx.x;
    if (i >= world || i == rank) return;
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint64_t *>(flags.p[i]) + rank, epoch);
    const uint64_t *mine = reinterpret_cast<const uint64_t *>(flags.p[rank]) + i;
    const uint64_t t0 = global_ns();
    while (ld_acquire_sys(mine) < epoch) {
        if (global_ns() - t0 > timeout_ns) __trap();      // never hang the GPU
        __nanosleep(64);
    }
}

// thresholds of the rank owning the first 10 s -> every arena (its own included)
__global__ void put_thresholds_kernel(const double *__restrict__ thr, int n_c, PeerPtrs dst, int world)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_c * world) return;
    const int p = i / n_c, c = i - p * n_c;
    reinterpret_cast<double *>(dst.p[p])[c] = thr[c];
}

// first H samples of every contact -> left neighbour's from_right; last H -> right neighbour's
// from_left.  grid (n_c), coalesced along time.
__global__ void __launch_bounds__(XC_THREADS)
put_halos_kernel(const double *__restrict__ sig, int64_t ld, int64_t n, int H,
                 double *__restrict__ left_from_right, double *__restrict__ right_from_left)
{
    const int64_t c = blockIdx.x;
    const double *row = sig + c * ld;
    for (int k = threadIdx.x; k < H; k += XC_THREADS) {
        if (left_from_right) left_from_right[c * H + k] = row[k];
        if (right_from_left) right_from_left[c * H + k] = row[n - H + k];
    }
}

// neighbours' samples into the margins of the shard's own buffer; flag of the crossing formed
// by the shard's last sample and the right neighbour's first (extract.py:86-91)
__global__ void __launch_bounds__(XC_THREADS)
install_halos_kernel(double *__restrict__ sig, int64_t ld, int64_t n, int H,
                     const double *__restrict__ from_left, const double *__restrict__ from_right,
                     const double *__restrict__ thr, int falling, int32_t *__restrict__ flags)
{
    const int64_t c = blockIdx.x;
    double *row = sig + c * ld;
    for (int k = threadIdx.x; k < H; k += XC_THREADS) {
        if (from_left) row[k - H] = from_left[c * H + k];
        if (from_right) row[n + k] = from_right[c * H + k];
    }
    if (threadIdx.x == 0)
        flags[c] = from_right ? (int)ssi::crossing(row[n - 1], from_right[c * H], thr[c], falling) : 0;
}

__device__ __forceinline__ int64_t xc_find_row(const int64_t *__restrict__ offsets, int64_t n_rows, int64_t s)
{
    int64_t lo = 0, hi = n_rows;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid) <= s) lo = mid; else hi = mid;
    }
    return lo;
}

// spt_out = spt_in with time t_last appended to every flagged row (it is the largest index
// of the shard, so the order is kept); every CTA rebuilds the exclusive scan of the flags.
__global__ void __launch_bounds__(XC_THREADS)
append_boundary_kernel(const double *__restrict__ spt_in, const int64_t *__restrict__ off_in, int n_c,
                       const int32_t *__restrict__ flags, double t_last,
                       double *__restrict__ spt_out, int64_t *__restrict__ off_out, int64_t cap_in)
{
    extern __shared__ int shift[];                        // [n_c + 1] exclusive scan of flags
    __shared__ int warp_tot[XC_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n_c + XC_THREADS - 1) / XC_THREADS;
    const int lo = min(tid * per, n_c), hi = min(lo + per, n_c);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += flags[i];
    int inc = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(SS_FULL, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    int run = base + inc - s;
    for (int i = lo; i < hi; ++i) { shift[i] = run; run += flags[i]; }
    if (tid == XC_THREADS - 1) shift[n_c] = run;
    __syncthreads();

    const int64_t k = (int64_t)blockIdx.x * XC_THREADS + tid;
    // off_in[n_c] > cap_in: the capacity-sized input list was truncated (the caller learns it
    // from the offsets and redoes the step); stay inside both buffers meanwhile
    const int64_t total = min(off_in[n_c], cap_in);
    if (k < total) {
        const int64_t r = xc_find_row(off_in, n_c, k);
        spt_out[k + shift[r]] = spt_in[k];
    }
    if (k < n_c) {
        off_out[k] = off_in[k] + shift[k];
        const int64_t pos = off_in[k + 1] + shift[k];
        if (flags[k] && pos < cap_in + n_c) spt_out[pos] = t_last;
    }
    if (k == 0) off_out[n_c] = total + shift[n_c];
}

// last aligned time of every contact -> right neighbour's prev_last (NaN: no spike here)
__global__ void put_last_kernel(const double *__restrict__ spt, const int64_t *__restrict__ offsets,
                                int n_c, double *__restrict__ right_prev_last)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_c) return;
    const int64_t lo = offsets[c], hi = offsets[c + 1];
    right_prev_last[c] = hi > lo ? spt[hi - 1] : __longlong_as_double(0x7ff8000000000000LL);
}

// Gram / sums of (x - ref) re-expressed for the zero reference,
//   sum x x^T = G + S r^T + r S^T + n r r^T,   sum x = S + n r,
// so that the partials of different shards (each with its own reference) simply add up;
// written into slot `rank` of EVERY arena.  grid (n_c).
__global__ void __launch_bounds__(XC_THREADS)
gram_put_kernel(const double *__restrict__ gram, const double *__restrict__ ssum,
                const double *__restrict__ ref, const int64_t *__restrict__ offsets, int n_c,
                int n_win, PeerPtrs slots, int world, const int64_t *__restrict__ status,
                const int32_t *__restrict__ halo_miss)
{
    const int64_t c = blockIdx.x;
    if (c == 0 && threadIdx.x == 0) {
        // status word of the step: 1 = a capacity-sized list overflowed, 2 = a window left the
        // halo (SS_SHARD_OVERFLOW / SS_SHARD_HALO_MISS); every rank sees the maximum
        double st = (status && *status) ? 1.0 : 0.0;
        if (halo_miss && *halo_miss) st = 2.0;
        const size_t at = (size_t)n_c * (n_win * n_win + n_win + 1);
        for (int p = 0; p < world; ++p) reinterpret_cast<double *>(slots.p[p])[at] = st;
    }
    const int nn = n_win * n_win;
    const double n = (double)(offsets[c + 1] - offsets[c]);
    const double *G = gram + c * nn, *S = ssum + c * n_win, *r = ref + c * n_win;
    for (int e = threadIdx.x; e < nn + n_win + 1; e += XC_THREADS) {
        double v;
        size_t at;
        if (e < nn) {
            const int i = e / n_win, j = e - i * n_win;
            v = G[e] + S[i] * r[j] + r[i] * S[j] + n * r[i] * r[j];
            at = (size_t)c * nn + e;
        } else if (e < nn + n_win) {
            const int i = e - nn;
            v = S[i] + n * r[i];
            at = (size_t)n_c * nn + (size_t)c * n_win + i;
        } else {
            v = n;
            at = (size_t)n_c * (nn + n_win) + c;
        }
        for (int p = 0; p < world; ++p) reinterpret_cast<double *>(slots.p[p])[at] = v;
    }
}

// sum of the slots in rank order (every rank gets bit-identical results)
__global__ void __launch_bounds__(XC_THREADS)
gram_reduce_kernel(const double *__restrict__ slots, size_t slot_stride_f64, int world, int n_c,
                   int n_win, double *__restrict__ gram, double *__restrict__ ssum,
                   int64_t *__restrict__ counts, int64_t *__restrict__ all_counts,
                   int64_t *__restrict__ status_max)
{
    const int64_t e = (int64_t)blockIdx.x * XC_THREADS + threadIdx.x;
    const int64_t ng = (int64_t)n_c * n_win * n_win, ns = (int64_t)n_c * n_win;
    if (e == ng + ns + n_c) {                              // status word: maximum over the ranks
        double m = 0.0;
        for (int p = 0; p < world; ++p) m = fmax(m, slots[(size_t)p * slot_stride_f64 + e]);
        if (status_max) *status_max = __double2ll_rn(m);
        return;
    }
    if (e >= ng + ns + n_c) return;
    double acc = 0.0;
    for (int p = 0; p < world; ++p) {
        const double v = slots[(size_t)p * slot_stride_f64 + e];
        acc += v;
        if (e >= ng + ns) all_counts[(int64_t)p * n_c + (e - ng - ns)] = __double2ll_rn(v);
    }
    if (e < ng) gram[e] = acc;
    else if (e < ng + ns) ssum[e - ng] = acc;
    else counts[e - ng - ns] = __double2ll_rn(acc);
}

// reference waveform of every group for the Gram shift: the group's first spike
template <typename T>
__global__ void first_wave_kernel(const T *__restrict__ waves, int n_win, int64_t n_spk, int n_c,
                                  const int64_t *__restrict__ offsets, int64_t n_groups,
                                  double *__restrict__ ref)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_groups * n_win) return;
    const int64_t g = i / n_win;
    const int w = (int)(i - g * n_win);
    double v = 0.0;
    if (n_spk > 0) {
        if (offsets) v = (double)waves[((int64_t)w * n_spk + min(offsets[g], n_spk - 1)) * n_c];
        else v = (double)waves[(int64_t)w * n_spk * n_c + g];
    }
    ref[i] = v;
}

// Results of one time shard (contact-major inside the shard) into the layout the host wants
// for the WHOLE recording: every list contact-major across shards, waveforms as one
// contiguous (n_win, n_tot[c], 1) block per contact - so that the host side of
// extract_spikes' result (extract.py:154-177) is a plain slice per contact.
// blockIdx.y == 0: times / valid flags / scores; blockIdx.y == 1 + w: waveform row w.
__global__ void __launch_bounds__(XC_THREADS)
merge_shard_kernel(const double *__restrict__ spt, const uint8_t *__restrict__ valid,
                   const double *__restrict__ scores, int ncomps, const float *__restrict__ waves,
                   int64_t n_src, const int64_t *__restrict__ src_off, int n_c,
                   const int64_t *__restrict__ dst_row_base, const int64_t *__restrict__ block_base,
                   const int64_t *__restrict__ n_tot, double *__restrict__ spt_out,
                   uint8_t *__restrict__ valid_out, double *__restrict__ scores_out,
                   float *__restrict__ waves_out, int n_win, int64_t wave_stride)
{
    const int64_t k = (int64_t)blockIdx.x * XC_THREADS + threadIdx.x;
    if (k >= n_src || k >= __ldg(src_off + n_c)) return;     // spare capacity
    const int64_t r = xc_find_row(src_off, n_c, k);
    const int64_t p = dst_row_base[r] + (k - src_off[r]);
    if (blockIdx.y == 0) {
        if (spt_out) spt_out[p] = spt[k];
        if (valid_out) valid_out[p] = valid[k];
        if (scores_out)
            for (int j = 0; j < ncomps; ++j) scores_out[p * ncomps + j] = scores[k * ncomps + j];
    } else {
        const int64_t w = blockIdx.y - 1, bb = block_base[r];
        waves_out[bb * n_win + w * n_tot[r] + (p - bb)] = waves[w * wave_stride + k];
    }
}

PeerPtrs peer_ptrs(void *const *bases, int world, size_t offset)
{
    PeerPtrs P;
    for (int i = 0; i < SS_MAX_PEERS; ++i)
        P.p[i] = i < world && bases[i] ? reinterpret_cast<char *>(bases[i]) + offset : nullptr;
    return P;
}

}  // namespace

extern "C" size_t ss_shard_arena_bytes(int n_c, int H, int n_win, int world)
{
    return arena_layout(n_c, H, n_win, world).total;
}

extern "C" int ss_shard_arena_offsets(int n_c, int H, int n_win, int world, size_t *out7)
{
    SS_REQUIRE(out7 && n_c >= 1 && H >= 0 && n_win >= 1 && world >= 1 && world <= SS_MAX_PEERS,
               SS_ERR_ARG, "arena: bad shape (world <= %d)", SS_MAX_PEERS);
    const ArenaLayout L = arena_layout(n_c, H, n_win, world);
    out7[0] = L.flags; out7[1] = L.thr; out7[2] = L.from_left; out7[3] = L.from_right;
    out7[4] = L.prev_last; out7[5] = L.slots; out7[6] = L.slot_stride;
    return SS_OK;
}

extern "C" int ss_peer_barrier(void *const *arenas, int rank, int world, uint64_t epoch,
                               double timeout_s, void *stream)
{
    SS_REQUIRE(arenas && world >= 1 && world <= SS_MAX_PEERS && rank >= 0 && rank < world,
               SS_ERR_ARG, "peer_barrier: bad rank/world %d/%d", rank, world);
    if (world == 1) return SS_OK;
    const ArenaLayout L = arena_layout(1, 0, 1, world);            // flags sit first
    peer_barrier_kernel<<<1, SS_MAX_PEERS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        peer_ptrs(arenas, world, L.flags), rank, world, epoch,
        (uint64_t)((timeout_s > 0 ? timeout_s : 10.0) * 1e9));
    SS_LAUNCH_CHECK("peer_barrier_kernel");
    return SS_OK;
}

extern "C" int ss_shard_put_thresholds(const double *d_thr, int n_c, int H, int n_win,
                                       void *const *arenas, int world, void *stream)
{
    SS_REQUIRE(d_thr && arenas && world >= 1 && world <= SS_MAX_PEERS, SS_ERR_ARG, "put_thresholds: bad argument");
    const ArenaLayout L = arena_layout(n_c, H, n_win, world);
    put_thresholds_kernel<<<(unsigned)ss_cdiv((int64_t)n_c * world, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        d_thr, n_c, peer_ptrs(arenas, world, L.thr), world);
    SS_LAUNCH_CHECK("put_thresholds_kernel");
    return SS_OK;
}

extern "C" int ss_shard_put_halos(const double *d_sig, int64_t ld, int n_c, int64_t n, int H, int n_win,
                                  void *const *arenas, int rank, int world, void *stream)
{
    SS_REQUIRE(d_sig && arenas && rank >= 0 && rank < world && world <= SS_MAX_PEERS, SS_ERR_ARG, "put_halos: bad argument");
    SS_REQUIRE(H >= 0 && H <= n, SS_ERR_ARG, "put_halos: halo %d larger than the shard", H);
    if (H == 0 || world == 1) return SS_OK;
    const ArenaLayout L = arena_layout(n_c, H, n_win, world);
    double *lfr = rank > 0 ? reinterpret_cast<double *>(reinterpret_cast<char *>(arenas[rank - 1]) + L.from_right) : nullptr;
    double *rfl = rank < world - 1 ? reinterpret_cast<double *>(reinterpret_cast<char *>(arenas[rank + 1]) + L.from_left) : nullptr;
    put_halos_kernel<<<n_c, XC_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_sig, ld, n, H, lfr, rfl);
    SS_LAUNCH_CHECK("put_halos_kernel");
    return SS_OK;
}

extern "C" int ss_shard_install_halos(double *d_sig, int64_t ld, int n_c, int64_t n, int H, int n_win,
                                      const void *arena, int rank, int world, const double *d_thr,
                                      int falling, int32_t *d_flags, void *stream)
{
    SS_REQUIRE(d_sig && arena && d_thr && d_flags && rank >= 0 && rank < world, SS_ERR_ARG, "install_halos: bad argument");
    const ArenaLayout L = arena_layout(n_c, H, n_win, world);
    const char *a = reinterpret_cast<const char *>(arena);
    const double *fl = rank > 0 && H > 0 ? reinterpret_cast<const double *>(a + L.from_left) : nullptr;
    const double *fr = rank < world - 1 && H > 0 ? reinterpret_cast<const double *>(a + L.from_right) : nullptr;
    install_halos_kernel<<<n_c, XC_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        d_sig, ld, n, H, fl, fr, d_thr, falling, d_flags);
    SS_LAUNCH_CHECK("install_halos_kernel");
    return SS_OK;
}

extern "C" int ss_shard_append_boundary(const double *d_spt_in, const int64_t *d_offsets_in, int n_c,
                                        int64_t n_spk_in, const int32_t *d_flags, double t_last,
                                        double *d_spt_out, int64_t *d_offsets_out, void *stream)
{
    SS_REQUIRE(d_offsets_in && d_flags && d_spt_out && d_offsets_out && (d_spt_in || n_spk_in == 0),
               SS_ERR_ARG, "append_boundary: null pointer");
    SS_REQUIRE(n_c >= 1 && n_c <= 12000, SS_ERR_UNSUPPORTED, "append_boundary: %d contacts > 12000", n_c);
    const int64_t work = n_spk_in > n_c + 1 ? n_spk_in : n_c + 1;
    append_boundary_kernel<<<(unsigned)ss_cdiv(work, XC_THREADS), XC_THREADS, (n_c + 1) * sizeof(int),
                             reinterpret_cast<cudaStream_t>(stream)>>>(
        d_spt_in, d_offsets_in, n_c, d_flags, t_last, d_spt_out, d_offsets_out, n_spk_in);
    SS_LAUNCH_CHECK("append_boundary_kernel");
    return SS_OK;
}

extern "C" int ss_shard_put_last(const double *d_spt, const int64_t *d_offsets, int n_c, int H, int n_win,
                                 void *const *arenas, int rank, int world, void *stream)
{
    SS_REQUIRE(d_offsets && arenas && rank >= 0 && rank < world && world <= SS_MAX_PEERS, SS_ERR_ARG, "put_last: bad argument");
    if (rank == world - 1) return SS_OK;
    const ArenaLayout L = arena_layout(n_c, H, n_win, world);
    double *dst = reinterpret_cast<double *>(reinterpret_cast<char *>(arenas[rank + 1]) + L.prev_last);
    put_last_kernel<<<(unsigned)ss_cdiv(n_c, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_spt, d_offsets, n_c, dst);
    SS_LAUNCH_CHECK("put_last_kernel");
    return SS_OK;
}

extern "C" int ss_shard_gram_put(const double *d_gram, const double *d_sum, const double *d_ref,
                                 const int64_t *d_offsets, int n_c, int H, int n_win,
                                 void *const *arenas, int rank, int world,
                                 const int64_t *d_status, const int32_t *d_halo_miss, void *stream)
{
    SS_REQUIRE(d_gram && d_sum && d_ref && d_offsets && arenas && rank >= 0 && rank < world &&
               world <= SS_MAX_PEERS, SS_ERR_ARG, "gram_put: bad argument");
    const ArenaLayout L = arena_layout(n_c, H, n_win, world);
    gram_put_kernel<<<n_c, XC_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        d_gram, d_sum, d_ref, d_offsets, n_c, n_win,
        peer_ptrs(arenas, world, L.slots + L.slot_stride * (size_t)rank), world, d_status, d_halo_miss);
    SS_LAUNCH_CHECK("gram_put_kernel");
    return SS_OK;
}

extern "C" int ss_shard_gram_reduce(const void *arena, int n_c, int H, int n_win, int world,
                                    double *d_gram, double *d_sum, int64_t *d_counts,
                                    int64_t *d_all_counts, int64_t *d_status_max, void *stream)
{
    SS_REQUIRE(arena && d_gram && d_sum && d_counts && d_all_counts, SS_ERR_ARG, "gram_reduce: null pointer");
    const ArenaLayout L = arena_layout(n_c, H, n_win, world);
    const int64_t total = (int64_t)n_c * (n_win * n_win + n_win + 1) + 1;
    gram_reduce_kernel<<<(unsigned)ss_cdiv(total, XC_THREADS), XC_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const double *>(reinterpret_cast<const char *>(arena) + L.slots),
        L.slot_stride / sizeof(double), world, n_c, n_win, d_gram, d_sum, d_counts, d_all_counts,
        d_status_max);
    SS_LAUNCH_CHECK("gram_reduce_kernel");
    return SS_OK;
}

extern "C" int ss_pca_first_wave(const void *d_waves, int dtype, int n_win, int64_t n_spk, int n_c,
                                 const int64_t *d_offsets, int64_t n_groups, double *d_ref,
                                 void *stream)
{
    SS_REQUIRE(d_ref && (d_waves || n_spk == 0) && n_groups >= 1 && n_win >= 1, SS_ERR_ARG, "first_wave: bad argument");
    SS_REQUIRE(dtype >= SS_I16 && dtype <= SS_F64, SS_ERR_ARG, "first_wave: bad dtype %d", dtype);
    const unsigned blocks = (unsigned)ss_cdiv(n_groups * n_win, 256);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == SS_F32)
        first_wave_kernel<float><<<blocks, 256, 0, st>>>(reinterpret_cast<const float *>(d_waves), n_win, n_spk, n_c, d_offsets, n_groups, d_ref);
    else if (dtype == SS_F64)
        first_wave_kernel<double><<<blocks, 256, 0, st>>>(reinterpret_cast<const double *>(d_waves), n_win, n_spk, n_c, d_offsets, n_groups, d_ref);
    else
        first_wave_kernel<int16_t><<<blocks, 256, 0, st>>>(reinterpret_cast<const int16_t *>(d_waves), n_win, n_spk, n_c, d_offsets, n_groups, d_ref);
    SS_LAUNCH_CHECK("first_wave_kernel");
    return SS_OK;
}

extern "C" int ss_merge_shard(const double *d_spt, const uint8_t *d_valid, const double *d_scores,
                              int ncomps, const float *d_waves, int n_win, int64_t wave_stride,
                              int64_t n_src, const int64_t *d_src_offsets, int n_c, const int64_t *d_dst_row_base,
                              const int64_t *d_block_base, const int64_t *d_n_tot,
                              double *d_spt_out, uint8_t *d_valid_out, double *d_scores_out,
                              float *d_waves_out, void *stream)
{
    if (n_src == 0) return SS_OK;
    SS_REQUIRE(d_src_offsets && d_dst_row_base && n_c >= 1, SS_ERR_ARG, "merge_shard: null layout");
    SS_REQUIRE((!d_spt_out || d_spt) && (!d_valid_out || d_valid) && (!d_scores_out || d_scores) &&
               (!d_waves_out || (d_waves && d_block_base && d_n_tot)), SS_ERR_ARG,
               "merge_shard: output without input");
    SS_REQUIRE(n_win >= 0 && n_win < 65535 && ncomps >= 0, SS_ERR_ARG, "merge_shard: bad shape");
    const dim3 grid((unsigned)ss_cdiv(n_src, XC_THREADS), d_waves_out ? 1 + n_win : 1);
    merge_shard_kernel<<<grid, XC_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        d_spt, d_valid, d_scores, ncomps, d_waves, n_src, d_src_offsets, n_c, d_dst_row_base,
        d_block_base, d_n_tot, d_spt_out, d_valid_out, d_scores_out, d_waves_out, n_win, wave_stride);
    SS_LAUNCH_CHECK("merge_shard_kernel");
    return SS_OK;
}
