// K6 on the 5th-generation tensor cores: per-contact Gram matrices of spike waveforms
// in 3xTF32 (tcgen05.mma.kind::tf32, accumulators in TMEM, float64 across chunks).
//
// Replaces the contraction inside np.cov, reference src/spike_sort/core/features.py:225
// (K = np.cov(data), data = (n_win, n_spikes) per contact).
//
// Shape of the problem: M = N = n_win (25..31) is tiny, K = n_spikes is huge, so the
// kernel is HBM-bound; the tensor cores are used to keep the arithmetic off the critical
// path (the float64 CUDA-core kernel in pca.cu is FP64-pipe bound on 10^7 spikes).
//
//  * up to 4 contacts share one 128 x 128 tile: contact p owns rows/columns
//    [32p, 32p + n_win) and row 32p + n_win is a row of ones, so column 32p + n_win of
//    the product holds the row sums needed for the mean (the off-diagonal blocks are
//    cross-contact products nobody reads);
//  * raw float32 rows (one (time step, K-block) = 32 spikes x n_c contacts, contiguous in
//    the reference's (n_win, n_spk, n_c) layout) arrive in shared memory by 1-D bulk
//    async copies (cp.async.bulk + mbarrier complete_tx), 128 spikes per copy, 2 stages;
//  * all threads convert a raw stage into the canonical K-major SWIZZLE_128B UMMA layout:
//    d = x - ref, hi = tf32(d), lo = d - hi;
//  * one thread issues, per 8 spikes, hi.hi^T + hi.lo^T + lo.hi^T as three
//    128x128x8 tcgen05.mma into the same TMEM accumulator and commits to the mbarrier
//    that frees the canonical stage (2 stages: conversion of block b+1 overlaps the
//    MMAs of block b);
//  * every 256 spikes the float32 accumulators are read back (tcgen05.ld 32x32b.x32,
//    warp p reads contact p's diagonal block) and added to float64 registers.
// Every mbarrier wait is bounded: a protocol error traps instead of hanging the GPU.
#include "common.cuh"
#include "internal.cuh"

namespace {

constexpr int TC_KB = 32;                  // spikes per K-block
constexpr int TC_RW = 32;                  // rows per contact in the tile (n_win + 1 <= 32)
constexpr int TC_R = 2;                    // raw stages
constexpr int TC_RB = 4;                   // K-blocks per raw stage: one bulk copy moves a row of
                                           // 128 spikes (2 KB at 4 contacts); 512-byte copies cost
                                           // ~10 % more (copy-engine overhead per request)
constexpr int TC_C = 2;                    // canonical (hi/lo) stages
constexpr int TC_CHUNK = 8;                // K-blocks per float32 accumulation chunk (256 spikes:
                                           // the TMEM accumulation truncates, so chunks stay short)
constexpr int TC_THREADS = 256;
constexpr int TC_CANON_BYTES = (TC_KB / 4) * 16 * 128;   // one 128-row x 32-spike matrix: 16 KB
// canonical K-major SWIZZLE_128B layout: one tile row = 32 spikes = 128 B, 8-row groups of
// 1024 B, the 16-byte chunk c of row r stored at chunk c ^ (r % 8).  (The no-swizzle
// "interleave" layout ran the MMAs at ~1/4 rate: 86 % tensor-pipe-active cycles at 18 % of
// peak in ncu - operand fetch bank conflicts.)
constexpr uint32_t TC_SBO = 1024;          // M-direction stride between 8-row groups (bytes)
constexpr int TC_TMEM_COLS = 256;          // two 128-column accumulators (chunks alternate)

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    for (long long spin = 0; spin < (1LL << 28); ++spin) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(addr), "r"(parity)
                     : "memory");
        if (ok) return;
    }
    __trap();                               // never hang the device on a protocol bug
}

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    // K-major, SWIZZLE_128B: ((8, n), 2) : ((128 B, SBO), 16 B); version 1 (Blackwell)
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) |
           ((uint64_t)(TC_SBO >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                          uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0), "r"(0),
                   "r"(0), "r"(0)
                 : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(bar))
                 : "memory");
}

// round-to-nearest (ties away) to tf32's 10-bit mantissa with integer ops: cvt.rna.tf32.f32
// runs on the quarter-rate XU pipe, which ncu showed saturated (107 %) by the conversion
__device__ __forceinline__ float tf32_round(float v)
{
    return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
}

// grid (parts, n_packs): contact mode n_packs == 1 and the CTA holds all n_c <= 4
// contacts; offsets mode (n_c == 1) one group per blockIdx.y.
__global__ void __launch_bounds__(TC_THREADS, 1)
pca_gram_tc_kernel(const float *__restrict__ waves, int n_win, int64_t n_spk, int n_c,
                   const int64_t *__restrict__ offsets, const double *__restrict__ ref,
                   double *__restrict__ ws_gram, double *__restrict__ ws_sum, int nwp)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // the swizzled tiles must start on a 1024-byte boundary of the shared window
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int parts = gridDim.x, part = blockIdx.x;
    const int P = offsets ? 1 : n_c;                       // contacts in the tile
    const int64_t g0 = offsets ? blockIdx.y : 0;           // first group of this CTA
    const int raw_row = TC_RB * TC_KB * n_c * 4 + 16;      // bytes, 16 B pad: conflict-free reads

    unsigned char *canon_hi = smem;                                    // [C][16 KB]
    unsigned char *canon_lo = canon_hi + TC_C * TC_CANON_BYTES;        // [C][16 KB]
    unsigned char *raw = canon_lo + TC_C * TC_CANON_BYTES;             // [R][n_win][raw_row]
    const int raw_stage = ((n_win * raw_row) + 127) & ~127;
    uint64_t *bars = reinterpret_cast<uint64_t *>(raw + TC_R * raw_stage);
    uint64_t *raw_full = bars, *canon_empty = bars + TC_R, *accum_full = bars + TC_R + TC_C;   // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + TC_R + TC_C + 2);
    float *ref_s = reinterpret_cast<float *>(tmem_slot + 4);           // [P][TC_RW]

    // spike range of this CTA
    int64_t lo_all, hi_all;
    if (offsets) { lo_all = offsets[g0]; hi_all = offsets[g0 + 1]; }
    else { lo_all = 0; hi_all = n_spk; }
    int64_t per = (hi_all - lo_all + parts - 1) / parts;
    per = (per + TC_KB - 1) / TC_KB * TC_KB;
    const int64_t lo = min(lo_all + (int64_t)part * per, hi_all), hi = min(lo + per, hi_all);
    const int64_t b0 = lo & ~(int64_t)3;                   // 16-byte aligned block origin
    const int nblk = (int)((hi - b0 + TC_KB - 1) / TC_KB);

    // ---- one-time setup
    for (int i = tid; i < 2 * TC_C * TC_CANON_BYTES / 16; i += TC_THREADS)
        reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < P * TC_RW; i += TC_THREADS) {
        const int p = i / TC_RW, w = i - p * TC_RW;
        ref_s[i] = w < n_win ? (float)ref[(g0 + p) * n_win + w] : 0.f;
    }
    if (tid == 0) {
        for (int i = 0; i < TC_R; ++i) mbar_init(raw_full + i, 1);
        for (int i = 0; i < TC_C; ++i) mbar_init(canon_empty + i, 1);
        mbar_init(accum_full, 1);
        mbar_init(accum_full + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // instruction descriptor: D f32, A/B tf32, both K-major, N = 128, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

    double acc[TC_RW];
#pragma unroll
    for (int j = 0; j < TC_RW; ++j) acc[j] = 0.0;

    // warp 7 issues the bulk copies of a block: lane 0 arms the barrier with the byte count,
    // lane w copies row w (one elected thread issuing all rows serialises ~25 copies per block)
    auto issue_load = [&](int sb) {                        // super-block = TC_RB K-blocks
        const int st = sb % TC_R;
        const int64_t s0 = b0 + (int64_t)sb * (TC_RB * TC_KB);
        const int64_t nsp = min((int64_t)(TC_RB * TC_KB), n_spk - s0);
        const uint32_t bytes = (uint32_t)(nsp * n_c * 4);
        const uint32_t bar = smem_u32(raw_full + st);
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                         ::"r"(bar), "r"(bytes * (uint32_t)n_win)
                         : "memory");
        __syncwarp();
        if (lane < n_win) {
            const float *src = waves + ((int64_t)lane * n_spk + s0) * n_c;
            const uint32_t dst = smem_u32(raw + st * raw_stage + lane * raw_row);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
                         "[%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                         : "memory");
        }
    };

    const int nsb = (nblk + TC_RB - 1) / TC_RB;
    if (warp == TC_THREADS / 32 - 1)
        for (int sb = 0; sb < TC_R - 1 && sb < nsb; ++sb) issue_load(sb);

    // every thread converts the same <= 4 (row, 4-spike group) items of every block: decode once
    const int WG = (n_win + 1 + 7) / 8;                    // 8-row groups per contact (incl. ones row)
    const int n_items = 8 * P * WG * (TC_KB / 4);          // <= 1024
    constexpr int TC_IT = 4;
    int it_raw[TC_IT], it_can[TC_IT], it_kind[TC_IT], it_s[TC_IT];     // kind 0 skip, 1 data, 2 ones
    float it_ref[TC_IT];
#pragma unroll
    for (int k = 0; k < TC_IT; ++k) {
        const int id = tid + k * TC_THREADS;
        it_kind[k] = 0; it_raw[k] = 0; it_can[k] = 0; it_s[k] = 0; it_ref[k] = 0.f;
        if (id < n_items) {
            const int i0 = id & 7;
            int rest = id >> 3;
            const int p = rest % P;
            rest /= P;
            const int wq = rest % WG, kc = rest / WG;
            const int w = wq * 8 + i0;
            if (w <= n_win) {
                const int r = p * TC_RW + w;
                it_kind[k] = w < n_win ? 1 : 2;
                it_raw[k] = w * raw_row + (4 * kc * n_c + p) * 4;
                it_can[k] = (r >> 3) * (int)TC_SBO + (r & 7) * 128 + ((kc ^ (r & 7)) << 4);
                it_s[k] = 4 * kc;
                it_ref[k] = ref_s[p * TC_RW + w];
            }
        }
    }
    const int cstride = n_c * 4;                           // bytes between consecutive spikes of a row
    for (int b = 0; b < nblk; ++b) {
        const int sb = b / TC_RB;
        if (warp == TC_THREADS / 32 - 1 && b % TC_RB == 0 && sb + TC_R - 1 < nsb)
            issue_load(sb + TC_R - 1);
        mbar_wait(raw_full + sb % TC_R, (uint32_t)((sb / TC_R) & 1));
        if (b >= TC_C) mbar_wait(canon_empty + b % TC_C, (uint32_t)(((b / TC_C) - 1) & 1));

        // ---- raw float32 -> canonical hi / lo tf32 tiles
        const unsigned char *rs = raw + (sb % TC_R) * raw_stage + (b % TC_RB) * (TC_KB * n_c * 4);
        unsigned char *ch = canon_hi + (b % TC_C) * TC_CANON_BYTES;
        unsigned char *cl = canon_lo + (b % TC_C) * TC_CANON_BYTES;
        const int64_t s0 = b0 + (int64_t)b * TC_KB;
        const bool whole = s0 >= lo && s0 + TC_KB <= hi;   // no masking needed
        if (n_c == 4 && !offsets) {
            // 4-contact fast path: thread (w = tid%32, kc = tid/32) converts 4 spikes of row w
            // for all 4 contacts: four 128-bit raw reads, eight 128-bit canonical stores
            const int w = tid & 31, kc = tid >> 5;
            if (w <= n_win) {
                float d[4][4];                             // [spike j][contact p]
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (w < n_win) {
                        const float4 q = *reinterpret_cast<const float4 *>(rs + w * raw_row + (4 * kc + j) * 16);
                        d[j][0] = q.x - ref_s[0 * TC_RW + w];
                        d[j][1] = q.y - ref_s[1 * TC_RW + w];
                        d[j][2] = q.z - ref_s[2 * TC_RW + w];
                        d[j][3] = q.w - ref_s[3 * TC_RW + w];
                    } else {
                        d[j][0] = d[j][1] = d[j][2] = d[j][3] = 1.0f;
                    }
                    if (!whole) {
                        const int64_t sp = s0 + 4 * kc + j;
                        if (sp < lo || sp >= hi) d[j][0] = d[j][1] = d[j][2] = d[j][3] = 0.f;
                    }
                }
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    float h[4], l[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        h[j] = tf32_round(d[j][p]);
                        l[j] = d[j][p] - h[j];          // the MMA truncates to tf32 itself
                    }
                    const int r = p * TC_RW + w;
                    const int off = (r >> 3) * (int)TC_SBO + (r & 7) * 128 + ((kc ^ (r & 7)) << 4);
                    *reinterpret_cast<float4 *>(ch + off) = make_float4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<float4 *>(cl + off) = make_float4(l[0], l[1], l[2], l[3]);
                }
            }
        } else
#pragma unroll
        for (int k = 0; k < TC_IT; ++k) {
            if (it_kind[k] == 0) continue;
            float d[4];
            if (it_kind[k] == 1) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    d[j] = *reinterpret_cast<const float *>(rs + it_raw[k] + j * cstride) - it_ref[k];
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) d[j] = 1.0f;
            }
            if (!whole) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int64_t sp = s0 + it_s[k] + j;
                    if (sp < lo || sp >= hi) d[j] = 0.f;
                }
            }
            float h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                h[j] = tf32_round(d[j]);
                l[j] = d[j] - h[j];                 // the MMA truncates to tf32 itself
            }
            *reinterpret_cast<float4 *>(ch + it_can[k]) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4 *>(cl + it_can[k]) = make_float4(l[0], l[1], l[2], l[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();

        const bool chunk_end = (b % TC_CHUNK == TC_CHUNK - 1) || b == nblk - 1;
        const int ci = b / TC_CHUNK;                       // chunk index; accumulator ci & 1
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ah = smem_u32(ch), al = smem_u32(cl);
            const uint32_t td = tmem_base + (uint32_t)((ci & 1) * 128);
#pragma unroll
            for (int k = 0; k < TC_KB / 8; ++k) {
                const uint64_t dh = umma_desc(ah + k * 32), dl = umma_desc(al + k * 32);
                const uint32_t first = (b % TC_CHUNK == 0 && k == 0) ? 0u : 1u;
                umma_tf32(td, dh, dh, idesc, first);
                umma_tf32(td, dh, dl, idesc, 1u);
                umma_tf32(td, dl, dh, idesc, 1u);
            }
            umma_commit(canon_empty + b % TC_C);
            if (chunk_end) umma_commit(accum_full + (ci & 1));
        }
        // the epilogue of a chunk runs one block LATER (or at the very end): its MMAs have
        // drained by then, and the next chunk accumulates in the other TMEM buffer meanwhile
        const bool last = b == nblk - 1;
        for (int pass = 0; pass < 2; ++pass) {
            int ce;                                        // chunk whose accumulator is read now
            if (pass == 0) { if (!(b % TC_CHUNK == 0 && b > 0)) continue; ce = ci - 1; }
            else { if (!last) continue; ce = ci; }
            mbar_wait(accum_full + (ce & 1), (uint32_t)((ce >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (warp < P) {
                uint32_t v[TC_RW];
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) +
                                       (uint32_t)((ce & 1) * 128 + warp * TC_RW);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                             "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]),
                               "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                               "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
                               "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                               "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                               "=r"(v[30]), "=r"(v[31])
                             : "r"(taddr)
                             : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < TC_RW; ++j) acc[j] += (double)__uint_as_float(v[j]);
            }
            // the buffer is overwritten (accumulate = 0) by chunk ce + 2, whose first MMAs
            // are issued after at least one more __syncthreads: order the reads before it
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
    }

    // ---- partial results of this (group, part): row `lane` of contact `warp`
    if (warp < P && lane < n_win) {
        const int64_t slot = (g0 + warp) * parts + part;
        double *og = ws_gram + slot * nwp * nwp + (int64_t)lane * nwp;
#pragma unroll
        for (int j = 0; j < TC_RW; ++j)
            if (j < n_win) og[j] = acc[j];
        double sv = 0.0;
#pragma unroll
        for (int j = 0; j < TC_RW; ++j)
            if (j == n_win) sv = acc[j];
        ws_sum[slot * nwp + lane] = sv;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                     ::"r"(tmem_base), "r"(TC_TMEM_COLS));
}

}  // namespace

bool ssi::pca_gram_tc_eligible(int n_win, int64_t n_spk, int n_c, bool offsets_mode)
{
    if (n_win + 1 > TC_RW) return false;
    if (offsets_mode ? n_c != 1 : n_c > 4) return false;
    if ((n_spk * n_c) % 4 != 0) return false;              // 16-byte aligned rows for the bulk copies
    return true;
}

int ssi::pca_gram_tc(const float *d_waves, int n_win, int64_t n_spk, int n_c,
                     const int64_t *d_offsets, int64_t n_groups, const double *d_ref,
                     double *ws_gram, double *ws_sum, int parts, int nwp, cudaStream_t st)
{
    const int raw_row = TC_RB * TC_KB * n_c * 4 + 16;
    const int raw_stage = ((n_win * raw_row) + 127) & ~127;
    const size_t smem = (size_t)2 * TC_C * TC_CANON_BYTES + (size_t)TC_R * raw_stage +
                        (TC_R + TC_C + 2) * 8 + 16 + (size_t)4 * TC_RW * 4 + 64 + 1024;
    SS_CUDA_CHECK(cudaFuncSetAttribute(pca_gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    const dim3 grid((unsigned)parts, (unsigned)(d_offsets ? n_groups : 1));
    pca_gram_tc_kernel<<<grid, TC_THREADS, smem, st>>>(d_waves, n_win, n_spk, n_c, d_offsets, d_ref,
                                                      ws_gram, ws_sum, nwp);
    SS_LAUNCH_CHECK("pca_gram_tc_kernel");
    return SS_OK;
}
