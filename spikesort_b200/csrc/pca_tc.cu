// K6 on the 5th-generation tensor cores: per-contact Gram matrices of spike waveforms
// in split-TF32 (tcgen05.mma.kind::tf32, accumulators in TMEM, float64 across blocks).
//
// Replaces the contraction inside np.cov, reference src/spike_sort/core/features.py:225
// (K = np.cov(data), data = (n_win, n_spikes) per contact).
//
// Shape of the problem: M = N = n_win (25..31) is tiny, K = n_spikes is huge, so the
// kernel should be HBM-bound; the tensor cores keep the arithmetic off the critical path
// (the float64 CUDA-core kernel in pca.cu is FP64-pipe bound on 10^7 spikes).
//
//  * up to 4 contacts share one 128 x 128 tile: contact p owns rows/columns
//    [32p, 32p + n_win) and row 32p + n_win is a row of ones, so that column of the product
//    holds the row sums needed for the mean (the off-diagonal blocks are cross-contact
//    products nobody reads; with P < 4 contacts the MMA's N shrinks to 32 P);
//  * d = x - ref is split as d = H + L, H = tf32(d) (round to nearest), L = d - H (exact in
//    float32).  With T = sum H (H + 2L)^T the Gram is G = (T + T^T) / 2 up to the L L^T
//    term (2^-22 relative): TWO MMAs per 8 spikes (H.H^T and H.(2L)^T into the same TMEM
//    accumulator) instead of the three of the classical 3xTF32, and the symmetrisation makes
//    the ones-row trick deliver sum(H + L) as well;
//  * the float32 TMEM accumulation truncates, a systematic shrink of ~2e-8 per accumulation:
//    the accumulator is read back every 32 spikes (8 accumulations, two TMEM buffers
//    alternate) into float64 registers - measured Gram error ~2e-7 (256-spike chunks: 2e-6);
//  * warp-specialised, mbarrier-synchronised, no block-wide barrier in the loop:
//      warp 12     producer: 1-D bulk async copies (cp.async.bulk + complete_tx) of the raw
//                  float32 rows, 64 spikes x n_c contacts per copy, TC_R stages;
//      warps 0-7   converters: raw stage -> canonical K-major SWIZZLE_128B tiles of H and 2L
//                  (TC_C stages), 128-bit shared-memory reads and writes;
//      warp 13     one thread issues the MMAs and commits to the barriers that free the
//                  canonical stage and publish the accumulator;
//      warps 8-11  epilogue: warp p reads contact p's 32 x 32 diagonal block from TMEM
//                  (tcgen05.ld 32x32b.x32) and adds it to float64 registers.
// Every mbarrier wait is bounded: a protocol error traps instead of hanging the GPU.
#include "common.cuh"
#include "internal.cuh"

namespace {

constexpr int TC_KB = 32;                  // spikes per K-block (= one accumulation chunk)
constexpr int TC_RW = 32;                  // rows per contact in the tile (n_win + 1 <= 32)
#ifndef TC_R_STAGES
#define TC_R_STAGES 4
#define TC_C_STAGES 3
#endif
constexpr int TC_R = TC_R_STAGES;          // raw stages
constexpr int TC_RB = 2;                   // K-blocks per raw stage: one bulk copy moves a row of
                                           // 64 spikes (1 KB at 4 contacts)
#ifndef TC_PF_STAGES
#define TC_PF_STAGES 0                     // L2 prefetch distance of the producer in raw stages; measured
#endif                                     // SLOWER with 12 (2.07 vs 1.64 ms on 10^7 spikes): off
constexpr int TC_PF = TC_PF_STAGES;
constexpr int TC_C = TC_C_STAGES;          // canonical (H / 2L) stages
constexpr int TC_NCONV = 8;                // converter warps (0..7): ~400 instructions per 32-spike block
                                           // and warp - with four of them (one per scheduler, nothing to
                                           // hide their latencies behind) they were the critical path
constexpr int TC_EPI0 = 8;                 // epilogue warps 8..11 (warp % 4 = TMEM lane quadrant)
constexpr int TC_PRODUCER = 12, TC_ISSUER = 13;
constexpr int TC_THREADS = 448;
constexpr int TC_CANON_BYTES = (TC_KB / 4) * 16 * 128;   // one 128-row x 32-spike matrix: 16 KB
// canonical K-major SWIZZLE_128B layout: one tile row = 32 spikes = 128 B, 8-row groups of
// 1024 B, the 16-byte chunk c of row r stored at chunk c ^ (r % 8).  (The no-swizzle
// "interleave" layout ran the MMAs at ~1/4 rate: 86 % tensor-pipe-active cycles at 18 % of
// peak in ncu - operand fetch bank conflicts.)
constexpr uint32_t TC_SBO = 1024;          // M-direction stride between 8-row groups (bytes)
constexpr int TC_TMEM_COLS = 256;          // two 128-column accumulators (blocks alternate)

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

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

    unsigned char *canon_h = smem;                                     // [C][16 KB]  H
    unsigned char *canon_l = canon_h + TC_C * TC_CANON_BYTES;          // [C][16 KB]  2L
    unsigned char *raw = canon_l + TC_C * TC_CANON_BYTES;              // [R][n_win][raw_row]
    const int raw_stage = ((n_win * raw_row) + 127) & ~127;
    uint64_t *bars = reinterpret_cast<uint64_t *>(raw + TC_R * raw_stage);
    uint64_t *raw_full = bars, *raw_empty = raw_full + TC_R, *canon_full = raw_empty + TC_R,
             *canon_empty = canon_full + TC_C, *accum_full = canon_empty + TC_C,
             *accum_empty = accum_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_empty + 2);
    float *ref_s = reinterpret_cast<float *>(tmem_slot + 4);           // [P][TC_RW]
    double *sym_s = reinterpret_cast<double *>(canon_h);               // epilogue scratch (tiles are dead)

    // spike range of this CTA
    int64_t lo_all, hi_all;
    if (offsets) { lo_all = offsets[g0]; hi_all = offsets[g0 + 1]; }
    else { lo_all = 0; hi_all = n_spk; }
    int64_t per = (hi_all - lo_all + parts - 1) / parts;
    per = (per + TC_KB - 1) / TC_KB * TC_KB;
    const int64_t lo = min(lo_all + (int64_t)part * per, hi_all), hi = min(lo + per, hi_all);
    const int64_t b0 = lo & ~(int64_t)3;                   // 16-byte aligned block origin
    const int nblk = (int)((hi - b0 + TC_KB - 1) / TC_KB);
    const int nsb = (nblk + TC_RB - 1) / TC_RB;

    // ---- one-time setup
    for (int i = tid; i < 2 * TC_C * TC_CANON_BYTES / 16; i += TC_THREADS)
        reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < P * TC_RW; i += TC_THREADS) {
        const int p = i / TC_RW, w = i - p * TC_RW;
        ref_s[i] = w < n_win ? (float)ref[(g0 + p) * n_win + w] : 0.f;
    }
    if (tid == 0) {
        for (int i = 0; i < TC_R; ++i) { mbar_init(raw_full + i, 1); mbar_init(raw_empty + i, TC_NCONV); }
        for (int i = 0; i < TC_C; ++i) { mbar_init(canon_full + i, TC_NCONV); mbar_init(canon_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(accum_full + i, 1); mbar_init(accum_empty + i, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    // the zero fill above was written by the generic proxy; the MMAs read through the async one
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    double acc[TC_RW];
#pragma unroll
    for (int j = 0; j < TC_RW; ++j) acc[j] = 0.0;

    if (warp == TC_PRODUCER) {
        // ---- raw float32 rows -> shared memory: lane 0 arms the barrier with the byte count,
        // lane w copies row w (one elected thread issuing all rows serialises ~25 copies)
        // The shared-memory stages hold ~100 KB of copies in flight per SM - not enough to cover
        // the DRAM latency of 25 row streams that lie 160 MB apart.  L2 prefetches TC_PF
        // super-blocks ahead (no shared memory needed) turn the staged copies into L2 hits.
        auto prefetch = [&](int sb) {
            const int64_t s0 = b0 + (int64_t)sb * (TC_RB * TC_KB);
            const int64_t nsp = min((int64_t)(TC_RB * TC_KB), n_spk - s0);
            if (lane < n_win && nsp > 0) {
                const float *src = waves + ((int64_t)lane * n_spk + s0) * n_c;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
                             ::"l"(src), "r"((uint32_t)(nsp * n_c * 4)) : "memory");
            }
        };
        if (TC_PF > 0)
            for (int sb = 0; sb < TC_PF && sb < nsb; ++sb) prefetch(sb);
        for (int sb = 0; sb < nsb; ++sb) {
            const int st = sb % TC_R;
            if (TC_PF > 0 && sb + TC_PF < nsb) prefetch(sb + TC_PF);
            mbar_wait(raw_empty + st, (uint32_t)(((sb / TC_R) & 1) ^ 1));
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
        }
    } else if (warp < TC_NCONV) {
        // ---- raw float32 -> canonical H / 2L tiles
        const int ctid = tid;                              // 0 .. 128 * TC_NCONV / 4 - 1
        constexpr int CT = 32 * TC_NCONV;
        // generic path: every thread converts the same (row, 4-spike group) items of every block
        const int WG = (n_win + 1 + 7) / 8;                // 8-row groups per contact (incl. ones row)
        const int n_items = 8 * P * WG * (TC_KB / 4);      // <= 1024
        constexpr int TC_IT = 1024 / CT;
        int it_raw[TC_IT], it_can[TC_IT], it_kind[TC_IT], it_s[TC_IT];     // kind 0 skip, 1 data, 2 ones
        float it_ref[TC_IT];
        const bool fast4 = n_c == 4 && !offsets;
        if (!fast4) {
#pragma unroll
            for (int k = 0; k < TC_IT; ++k) {
                const int id = ctid + k * CT;
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
        }
        const int cstride = n_c * 4;                       // bytes between consecutive spikes of a row
        // 4-contact fast path: row w = lane, K-groups kc = warp + TC_NCONV * half
        int f4_raw[8 / TC_NCONV], f4_can[8 / TC_NCONV][4];
        float f4_ref[4];
#pragma unroll
        for (int half = 0; half < 8 / TC_NCONV; ++half) {
            const int kc = warp + half * TC_NCONV;
            f4_raw[half] = lane * raw_row + 4 * kc * 16;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int r = p * TC_RW + lane;
                f4_can[half][p] = (r >> 3) * (int)TC_SBO + (r & 7) * 128 + ((kc ^ (r & 7)) << 4);
            }
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) f4_ref[p] = (fast4 && lane < n_win) ? ref_s[p * TC_RW + lane] : 0.f;
        for (int b = 0; b < nblk; ++b) {
            const int sb = b / TC_RB, st = sb % TC_R, cs = b % TC_C;
            if (b % TC_RB == 0) mbar_wait(raw_full + st, (uint32_t)((sb / TC_R) & 1));
            mbar_wait(canon_empty + cs, (uint32_t)(((b / TC_C) & 1) ^ 1));
            const unsigned char *rs = raw + st * raw_stage + (b % TC_RB) * (TC_KB * n_c * 4);
            unsigned char *ch = canon_h + cs * TC_CANON_BYTES;
            unsigned char *cl = canon_l + cs * TC_CANON_BYTES;
            const int64_t s0 = b0 + (int64_t)b * TC_KB;
            const bool whole = s0 >= lo && s0 + TC_KB <= hi;   // no masking needed
            if (fast4 && whole) {
                // the common block: no masking, offsets and references hoisted out of the loop.
                // thread (w = lane, kc) converts 4 spikes of row w for all 4 contacts: four
                // 128-bit raw reads, eight 128-bit canonical stores; TC_NCONV warps x 2 passes
                if (lane < n_win) {
#pragma unroll
                    for (int half = 0; half < 8 / TC_NCONV; ++half) {
                        const float4 *src = reinterpret_cast<const float4 *>(rs + f4_raw[half]);
                        const float4 q0 = src[0], q1 = src[1], q2 = src[2], q3 = src[3];
                        const float dd[4][4] = {{q0.x, q1.x, q2.x, q3.x}, {q0.y, q1.y, q2.y, q3.y},
                                                {q0.z, q1.z, q2.z, q3.z}, {q0.w, q1.w, q2.w, q3.w}};
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            float h[4], l[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float d = dd[p][j] - f4_ref[p];
                                h[j] = tf32_round(d);
                                l[j] = 2.0f * (d - h[j]);          // exact; the MMA truncates to tf32
                            }
                            *reinterpret_cast<float4 *>(ch + f4_can[half][p]) = make_float4(h[0], h[1], h[2], h[3]);
                            *reinterpret_cast<float4 *>(cl + f4_can[half][p]) = make_float4(l[0], l[1], l[2], l[3]);
                        }
                    }
                } else if (lane == n_win) {
#pragma unroll
                    for (int half = 0; half < 8 / TC_NCONV; ++half)
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            *reinterpret_cast<float4 *>(ch + f4_can[half][p]) = make_float4(1.f, 1.f, 1.f, 1.f);
                            *reinterpret_cast<float4 *>(cl + f4_can[half][p]) = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                }
            } else if (fast4) {
                // thread (w = lane, kc) converts 4 spikes of row w for all 4 contacts: four
                // 128-bit raw reads, eight 128-bit canonical stores; TC_NCONV warps x 2 passes
                const int w = lane;
                if (w <= n_win) {
#pragma unroll
                    for (int half = 0; half < 8 / TC_NCONV; ++half) {
                        const int kc = warp + half * TC_NCONV;
                        float d[4][4];                         // [spike j][contact p]
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
                                l[j] = 2.0f * (d[j][p] - h[j]);    // exact; the MMA truncates to tf32
                            }
                            const int r = p * TC_RW + w;
                            const int off = (r >> 3) * (int)TC_SBO + (r & 7) * 128 + ((kc ^ (r & 7)) << 4);
                            *reinterpret_cast<float4 *>(ch + off) = make_float4(h[0], h[1], h[2], h[3]);
                            *reinterpret_cast<float4 *>(cl + off) = make_float4(l[0], l[1], l[2], l[3]);
                        }
                    }
                }
            } else {
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
                        l[j] = 2.0f * (d[j] - h[j]);
                    }
                    *reinterpret_cast<float4 *>(ch + it_can[k]) = make_float4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<float4 *>(cl + it_can[k]) = make_float4(l[0], l[1], l[2], l[3]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(canon_full + cs);
                if (b % TC_RB == TC_RB - 1 || b == nblk - 1) mbar_arrive(raw_empty + st);
            }
        }
    } else if (warp == TC_ISSUER) {
        // ---- one thread issues all MMAs
        if (lane == 0) {
            // instruction descriptor: D f32, A/B tf32, both K-major, N = 32 P, M = 128
            const uint32_t Ncols = 32u * (uint32_t)P;
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((Ncols >> 3) << 17) | ((128u >> 4) << 24);
            for (int b = 0; b < nblk; ++b) {
                const int cs = b % TC_C, ab = b & 1;
                mbar_wait(canon_full + cs, (uint32_t)((b / TC_C) & 1));
                mbar_wait(accum_empty + ab, (uint32_t)(((b >> 1) & 1) ^ 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t ah = smem_u32(canon_h + cs * TC_CANON_BYTES);
                const uint32_t al = smem_u32(canon_l + cs * TC_CANON_BYTES);
                const uint32_t td = tmem_base + (uint32_t)(ab * 128);
#pragma unroll
                for (int k = 0; k < TC_KB / 8; ++k) {
                    const uint64_t dh = umma_desc(ah + k * 32), dl = umma_desc(al + k * 32);
                    umma_tf32(td, dh, dh, idesc, k == 0 ? 0u : 1u);     // H H^T
                    umma_tf32(td, dh, dl, idesc, 1u);                   // H (2L)^T
                }
                umma_commit(canon_empty + cs);
                umma_commit(accum_full + ab);
            }
        }
    } else if (warp >= TC_EPI0 && warp < TC_EPI0 + 4) {
        // ---- accumulators -> float64: warp p reads contact p's diagonal block
        const int p = warp - TC_EPI0;
        for (int b = 0; b < nblk; ++b) {
            const int ab = b & 1;
            mbar_wait(accum_full + ab, (uint32_t)((b >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (p < P) {
                uint32_t v[TC_RW];
                const uint32_t taddr = tmem_base + ((uint32_t)(p * 32) << 16) +
                                       (uint32_t)(ab * 128 + p * TC_RW);
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
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(accum_empty + ab);      // the MMAs may overwrite it now
#pragma unroll
                for (int j = 0; j < TC_RW; ++j) acc[j] += (double)__uint_as_float(v[j]);
            } else {
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(accum_empty + ab);
            }
        }
    }

    // ---- partial results of this (group, part): G = (T + T^T) / 2 per contact, through shared
    // memory (the tiles are dead once every MMA has been committed and read back)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp >= TC_EPI0 && warp < TC_EPI0 + 4 && warp - TC_EPI0 < P) {
        const int p = warp - TC_EPI0;
        double *T = sym_s + p * (TC_RW * (TC_RW + 1));                // [32][33]
#pragma unroll
        for (int j = 0; j < TC_RW; ++j) T[lane * (TC_RW + 1) + j] = acc[j];
        __syncwarp();
        if (lane < n_win) {
            const int64_t slot = (g0 + p) * parts + part;
            double *og = ws_gram + slot * nwp * nwp + (int64_t)lane * nwp;
            for (int j = 0; j < n_win; ++j)
                og[j] = 0.5 * (T[lane * (TC_RW + 1) + j] + T[j * (TC_RW + 1) + lane]);
            ws_sum[slot * nwp + lane] = 0.5 * (T[lane * (TC_RW + 1) + n_win] + T[n_win * (TC_RW + 1) + lane]);
        }
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
                        (2 * TC_R + 2 * TC_C + 4) * 8 + 16 + (size_t)4 * TC_RW * 4 + 64 + 1024;
    SS_REQUIRE(smem <= 232448, SS_ERR_UNSUPPORTED, "pca_gram_tc: %zu bytes of shared memory", smem);
    SS_CUDA_CHECK(cudaFuncSetAttribute(pca_gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    const dim3 grid((unsigned)parts, (unsigned)(d_offsets ? n_groups : 1));
    pca_gram_tc_kernel<<<grid, TC_THREADS, smem, st>>>(d_waves, n_win, n_spk, n_c, d_offsets, d_ref,
                                                      ws_gram, ws_sum, nwp);
    SS_LAUNCH_CHECK("pca_gram_tc_kernel");
    return SS_OK;
}
