// K6-K8 - per-contact PCA of spike waveforms.
//
// Replaces PCA / fetPCA, reference src/spike_sort/core/features.py:224-231 and
// 326-344:  K = np.cov(data) (ddof = 1); eig; sort descending; |evals|;
// score = evecs[:, :ncomps].T @ data / sqrt(evals[:ncomps])  (UN-centred data).
//
//   pca_gram_kernel    : float64 Gram matrix and column sums of the waveforms
//                        shifted by a reference waveform (kills the cancellation
//                        in sum(x x^T) - n mu mu^T for peak-aligned spikes);
//                        4x4 register tiles over the upper triangle, spikes staged
//                        in shared memory; per-(group, part) partials reduced in a
//                        fixed order (deterministic, and additive across time
//                        shards -> the quantity the multi-GPU path all-reduces);
//   pca_eig_kernel     : covariance from Gram/sums/count, parallel-ordering
//                        (round-robin) cyclic Jacobi in float64, one CTA per group;
//   pca_project_kernel : scores, one thread per (spike, group).
#include "common.cuh"
#include "internal.cuh"
#include <cstdlib>

namespace {

constexpr int PC_THREADS = 256;
constexpr int PC_SC = 64;               // spikes per staged chunk
constexpr int PC_MAX_WIN = 112;             // eig: 2 * 112 * 113 * 8 B = 202 KB of shared memory
constexpr int PC_MAX_TILES = 2;         // 4x4 tiles per thread (28*29/2 = 406 <= 512)
constexpr int PC_MAX_COMPS = 8;

__host__ __device__ inline int pc_pad4(int n) { return (n + 3) & ~3; }

static int pc_parts(int64_t n_groups)
{
    int64_t p = 592 / n_groups;
    if (p < 1) p = 1;
    if (p > 148) p = 148;
    return (int)p;
}

// grid (parts, n_groups); T = element type of the waveform block (the reference promotes any
// dtype to float64, features.py:224: float32 from extract_spikes, float64 from resample_spikes)
template <typename T>
__global__ void __launch_bounds__(PC_THREADS)
pca_gram_kernel(const T *__restrict__ waves, int n_win, int64_t n_spk, int n_c,
                const int64_t *__restrict__ offsets, const double *__restrict__ ref,
                double *__restrict__ ws_gram, double *__restrict__ ws_sum)
{
    extern __shared__ __align__(16) double sh[];         // [PC_SC][RS]
    __shared__ double s_ref[PC_MAX_WIN + 4];
    __shared__ unsigned char s_ti[PC_THREADS * PC_MAX_TILES], s_tj[PC_THREADS * PC_MAX_TILES];
    const int nwp = pc_pad4(n_win), RS = nwp + 2, ntile = nwp / 4;
    const int n_tiles = ntile * (ntile + 1) / 2;
    const int tid = threadIdx.x;
    const int64_t g = blockIdx.y;
    const int parts = gridDim.x, part = blockIdx.x;

    for (int i = tid; i < nwp; i += PC_THREADS) s_ref[i] = i < n_win ? ref[g * n_win + i] : 0.0;
    if (tid == 0) {                                       // enumerate upper-triangle tiles row-major
        int k = 0;
        for (int a = 0; a < ntile; ++a)
            for (int b = a; b < ntile; ++b) { s_ti[k] = (unsigned char)a; s_tj[k] = (unsigned char)b; ++k; }
    }
    int64_t lo_all, hi_all, c;
    if (offsets) { lo_all = offsets[g]; hi_all = offsets[g + 1]; c = 0; }
    else { lo_all = 0; hi_all = n_spk; c = g; }
    const int64_t cnt = hi_all - lo_all;
    const int64_t per = (cnt + parts - 1) / parts;
    const int64_t lo = min(lo_all + (int64_t)part * per, hi_all), hi = min(lo + per, hi_all);
    __syncthreads();

    double acc[PC_MAX_TILES][4][4];
#pragma unroll
    for (int t = 0; t < PC_MAX_TILES; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[t][i][j] = 0.0;
    double sacc = 0.0;
    int my_ti[PC_MAX_TILES], my_tj[PC_MAX_TILES];
    bool my_ok[PC_MAX_TILES];
#pragma unroll
    for (int t = 0; t < PC_MAX_TILES; ++t) {
        const int k = tid + t * PC_THREADS;
        my_ok[t] = k < n_tiles;
        my_ti[t] = my_ok[t] ? s_ti[k] : 0;
        my_tj[t] = my_ok[t] ? s_tj[k] : 0;
    }

    for (int64_t base = lo; base < hi; base += PC_SC) {
        const int ns = (int)min((int64_t)PC_SC, hi - base);
        // stage: threads along spikes (global stride n_c floats), shifted to float64
        for (int idx = tid; idx < PC_SC * nwp; idx += PC_THREADS) {
            const int w = idx / PC_SC, sl = idx - w * PC_SC;
            double v = 0.0;
            if (sl < ns && w < n_win)
                v = ss_to_f64(__ldg(waves + ((int64_t)w * n_spk + base + sl) * n_c + c)) - s_ref[w];
            sh[sl * RS + w] = v;
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < PC_MAX_TILES; ++t) {
            if (my_ok[t]) {
                const double *pa = sh + 4 * my_ti[t], *pb = sh + 4 * my_tj[t];
#pragma unroll 4
                for (int s = 0; s < ns; ++s) {
                    const double2 a01 = *reinterpret_cast<const double2 *>(pa + s * RS);
                    const double2 a23 = *reinterpret_cast<const double2 *>(pa + s * RS + 2);
                    const double2 b01 = *reinterpret_cast<const double2 *>(pb + s * RS);
                    const double2 b23 = *reinterpret_cast<const double2 *>(pb + s * RS + 2);
                    const double a[4] = {a01.x, a01.y, a23.x, a23.y};
                    const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[t][i][j] += a[i] * b[j];
                }
            }
        }
        if (tid < nwp)
            for (int s = 0; s < ns; ++s) sacc += sh[s * RS + tid];
        __syncthreads();
    }

    double *og = ws_gram + ((int64_t)g * parts + part) * nwp * nwp;
#pragma unroll
    for (int t = 0; t < PC_MAX_TILES; ++t)
        if (my_ok[t])
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    og[(4 * my_ti[t] + i) * nwp + 4 * my_tj[t] + j] = acc[t][i][j];
    if (tid < nwp) ws_sum[((int64_t)g * parts + part) * nwp + tid] = sacc;
}

// grid (n_groups): sums the partials in part order, mirrors the upper triangle
__global__ void __launch_bounds__(PC_THREADS)
pca_gram_reduce_kernel(const double *__restrict__ ws_gram, const double *__restrict__ ws_sum,
                       int n_win, int parts, double *__restrict__ gram, double *__restrict__ sum)
{
    const int nwp = pc_pad4(n_win);
    const int64_t g = blockIdx.x;
    for (int idx = threadIdx.x; idx < n_win * n_win; idx += PC_THREADS) {
        const int i = idx / n_win, j = idx - i * n_win;
        const int a = i <= j ? i : j, b = i <= j ? j : i;
        double s = 0.0;
        for (int p = 0; p < parts; ++p) s += ws_gram[((int64_t)g * parts + p) * nwp * nwp + a * nwp + b];
        gram[(int64_t)g * n_win * n_win + idx] = s;
    }
    for (int i = threadIdx.x; i < n_win; i += PC_THREADS) {
        double s = 0.0;
        for (int p = 0; p < parts; ++p) s += ws_sum[((int64_t)g * parts + p) * nwp + i];
        sum[(int64_t)g * n_win + i] = s;
    }
}

// One CTA per group.  smem: A[ne*ld], V[ne*ld] (row stride ld odd: column sweeps are
// bank-conflict free), rotation (c, s) per pair, reduction scratch.  A round is two phases:
// the ne/2 disjoint rotations, then A by 2x2 blocks (thread = (row pair, column pair)) and
// the columns of V (warp = pair, lane = row) - no integer division in the sweeps.
__global__ void __launch_bounds__(PC_THREADS)
pca_eig_kernel(const double *__restrict__ gram, const double *__restrict__ sum,
               const int64_t *__restrict__ counts, int n, double *__restrict__ evals,
               double *__restrict__ evecs)
{
    extern __shared__ double sm[];
    const int ne = (n + 1) & ~1, ld = ne | 1;
    double *A = sm, *V = A + ne * ld, *rc = V + ne * ld, *rs = rc + ne / 2, *red = rs + ne / 2;
    __shared__ int s_p[PC_MAX_WIN / 2 + 1], s_q[PC_MAX_WIN / 2 + 1];
    __shared__ int s_done;
    __shared__ double s_prev_off;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    if (tid == 0) s_prev_off = 1e300;
    const int64_t g = blockIdx.x;
    const double cnt = (double)counts[g];
    for (int i = warp; i < ne; i += nwarps)
        for (int j = lane; j < ne; j += 32) {
            double a = 0.0;
            if (i < n && j < n)
                a = (gram[(g * n + i) * n + j] - sum[g * n + i] * sum[g * n + j] / cnt) / (cnt - 1.0);
            A[i * ld + j] = a;
            V[i * ld + j] = i == j ? 1.0 : 0.0;
        }
    __syncthreads();
    const int half = ne / 2;
    for (int sweep = 0; sweep < 60; ++sweep) {
        // convergence: off-diagonal energy against diagonal energy
        double off = 0.0, dg = 0.0;
        for (int i = warp; i < n; i += nwarps)
            for (int j = lane; j < n; j += 32) {
                const double a = A[i * ld + j];
                if (i == j) dg += a * a; else off += a * a;
            }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            off += __shfl_xor_sync(SS_FULL, off, d);
            dg += __shfl_xor_sync(SS_FULL, dg, d);
        }
        if (lane == 0) { red[2 * warp] = off; red[2 * warp + 1] = dg; }
        __syncthreads();
        if (tid == 0) {
            double o = 0.0, d2 = 0.0;
            for (int w = 0; w < nwarps; ++w) { o += red[2 * w]; d2 += red[2 * w + 1]; }
            // converged, or sitting on the round-off floor: for some matrices the off-diagonal
            // energy stalls just above 1e-30 of the diagonal's and the sweeps would run to
            // the limit (seen as a 1.5 ms eigensolve at 4 GPUs) without changing anything
            s_done = !(o > 1e-30 * d2) || (o <= 1e-24 * d2 && o >= 0.25 * s_prev_off);   // also stops on NaN
            s_prev_off = o;
        }
        __syncthreads();
        if (s_done) break;
        for (int rd = 0; rd < ne - 1; ++rd) {
            for (int k = tid; k < half; k += blockDim.x) {
                int p, q;
                if (k == 0) { p = ne - 1; q = rd; }
                else {
                    p = rd + k; if (p >= ne - 1) p -= ne - 1;
                    q = rd - k; if (q < 0) q += ne - 1;
                }
                if (p > q) { const int t = p; p = q; q = t; }
                double c = 1.0, s = 0.0;
                if (q < n) {
                    // rotation that zeroes A[p][q] (the smaller angle), without divisions:
                    // cos 2t = |alpha|/hyp, c = sqrt((1 + cos 2t)/2), s = sgn(alpha) beta/(2 c hyp)
                    const double apq = A[p * ld + q];
                    if (apq != 0.0) {
                        const double alpha = A[q * ld + q] - A[p * ld + p], beta = 2.0 * apq;
                        const double h2 = alpha * alpha + beta * beta;
                        if (h2 > 0.0 && h2 < 1e300) {
                            const double ih = rsqrt(h2);
                            const double c2 = 0.5 + 0.5 * fabs(alpha) * ih;
                            const double ic = rsqrt(c2);
                            c = c2 * ic;
                            s = (alpha >= 0.0 ? 0.5 : -0.5) * beta * ih * ic;
                        }
                    }
                }
                s_p[k] = p; s_q[k] = q; rc[k] = c; rs[k] = s;
            }
            __syncthreads();
            // A <- J^T A J by 2x2 blocks: the block of row pair bi and column pair bj only
            // needs the two rotations, so one pass (and one barrier) updates all of A;
            // column rotation first, then row rotation (B' = J1^T (B J2)).
            for (int bi = tid >> 4; bi < half; bi += PC_THREADS / 16) {
                const int p1 = s_p[bi], q1 = s_q[bi];
                const double c1 = rc[bi], s1 = rs[bi];
                for (int bj = tid & 15; bj < half; bj += 16) {
                    const double c2 = rc[bj], s2 = rs[bj];
                    if (s1 == 0.0 && s2 == 0.0) continue;
                    const int p2 = s_p[bj], q2 = s_q[bj];
                    const double b00 = A[p1 * ld + p2], b01 = A[p1 * ld + q2];
                    const double b10 = A[q1 * ld + p2], b11 = A[q1 * ld + q2];
                    const double t00 = c2 * b00 - s2 * b01, t01 = s2 * b00 + c2 * b01;
                    const double t10 = c2 * b10 - s2 * b11, t11 = s2 * b10 + c2 * b11;
                    A[p1 * ld + p2] = c1 * t00 - s1 * t10;
                    A[p1 * ld + q2] = c1 * t01 - s1 * t11;
                    A[q1 * ld + p2] = s1 * t00 + c1 * t10;
                    A[q1 * ld + q2] = s1 * t01 + c1 * t11;
                }
            }
            // V <- V J (columns only)
            for (int k = warp; k < half; k += nwarps) {
                const double c = rc[k], s = rs[k];
                if (s != 0.0) {
                    const int p = s_p[k], q = s_q[k];
                    for (int i = lane; i < ne; i += 32) {
                        const double vip = V[i * ld + p], viq = V[i * ld + q];
                        V[i * ld + p] = c * vip - s * viq;
                        V[i * ld + q] = s * vip + c * viq;
                    }
                }
            }
            __syncthreads();
        }
    }
    // rank by signed eigenvalue, descending (np.argsort(evals)[::-1]), then |.|
    for (int k = tid; k < n; k += blockDim.x) {
        const double lk = A[k * ld + k];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const double lj = A[j * ld + j];
            if (lj > lk || (lj == lk && j > k)) ++rank;
        }
        evals[g * n + rank] = fabs(lk);
        for (int i = 0; i < n; ++i) evecs[(g * n + i) * n + rank] = V[i * ld + k];
    }
}

// ---- top-k eigenpairs only (what fetPCA needs: the scores of the first ncomps components) ----
// pca_eig_kernel's Jacobi sweeps are latency-bound (about 230 rounds of ~1 us for n = 30).  When
// only the k leading pairs are wanted the classical dense path is ~8x shorter: Householder
// tridiagonalisation (n - 2 steps), Sturm-sequence MULTISECTION for the k largest eigenvalues
// (32/k shifts per pass in the lanes of the warp, division-free determinant recurrence on the
// scaled matrix), inverse iteration with a pivoted tridiagonal LU (reciprocal pivots, three
// iterations, Gram-Schmidt inside clusters as LAPACK's dstein does), back-transformation by
// the stored reflectors.  One WARP per group, everything in shared memory, __syncwarp only.
// tools/proto_eig_top.py is the NumPy statement of the same operation order.
constexpr int ET_MAXK = PC_MAX_COMPS;

__device__ __forceinline__ double et_sum(double v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(SS_FULL, v, d);
    return v;
}
__device__ __forceinline__ double et_max(double v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmax(v, __shfl_xor_sync(SS_FULL, v, d));
    return v;
}
__device__ __forceinline__ double et_min(double v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmin(v, __shfl_xor_sync(SS_FULL, v, d));
    return v;
}

__host__ __device__ inline size_t et_smem_doubles(int n, int k)
{
    return (size_t)n * (n | 1) + 6 * (size_t)n + 2 * ET_MAXK + (size_t)k * n * 5 + (size_t)(k * n + 1) / 2 + 2 + 16;
}

__global__ void __launch_bounds__(32)
pca_eig_top_kernel(const double *__restrict__ gram, const double *__restrict__ sum,
                   const int64_t *__restrict__ counts, int n, int k, double *__restrict__ evals,
                   double *__restrict__ evecs)
{
    extern __shared__ double sm[];
    const int ld = n | 1, lane = threadIdx.x;
    double *A = sm, *d = A + n * ld, *e = d + n, *e2 = e + n, *tau = e2 + n, *vs = tau + n, *ws = vs + n;
    double *lam = ws + n, *lamu = lam + ET_MAXK, *Z = lamu + ET_MAXK, *LU = Z + k * n;
    int *piv = reinterpret_cast<int *>(LU + (size_t)k * 4 * n);
    // Sturm recurrence operands, padded to a multiple of 8 rows: ds[i] = d[i], es2[i] = e[i-1]^2
    // (they alias the LU area, which is only filled after the multisection)
    const int n8 = (n + 7) & ~7;
    double *ds = LU, *es2 = LU + n8;
    const int64_t g = blockIdx.x;
    const double cnt = (double)counts[g];
    {
        const double icnt = 1.0 / cnt, idof = 1.0 / (cnt - 1.0);
        const double *G = gram + g * n * n, *S = sum + g * n;
#pragma unroll 4
        for (int idx = lane; idx < n * n; idx += 32) {
            const int i = idx / n, j = idx - i * n;
            A[i * ld + j] = (G[idx] - S[i] * S[j] * icnt) * idof;
        }
    }
    __syncwarp();

    // 1. A = Q T Q^T; reflector kk: v = (0.., 1, A[kk+2.., kk]), kept in column kk of A.
    //    The single warp lives on instruction-level parallelism: columns are taken eight at a
    //    time through registers (all loads of a batch before its stores: the compiler cannot
    //    prove that the row, v and w arrays do not alias)
    for (int kk = 0; kk < n - 2; ++kk) {
        const double alpha = A[(kk + 1) * ld + kk];
        double x2 = 0.0;
        for (int i = kk + 2 + lane; i < n; i += 32) { const double a = A[i * ld + kk]; x2 += a * a; }
        x2 = et_sum(x2);
        if (lane == 0) d[kk] = A[kk * ld + kk];
        if (!(x2 > 0.0)) {                                   // nothing to annihilate (or NaN)
            if (lane == 0) { tau[kk] = 0.0; e[kk] = alpha; }
            __syncwarp();
            continue;
        }
        const double h2 = alpha * alpha + x2, rh = rsqrt(h2);
        const double beta = -copysign(h2 * rh, alpha);
        const double tk = (alpha - beta) * copysign(rh, alpha), sc = 1.0 / (alpha - beta);
        for (int i = lane; i < n; i += 32) {
            double v = 0.0;
            if (i == kk + 1) v = 1.0;
            else if (i > kk + 1) { v = A[i * ld + kk] * sc; A[i * ld + kk] = v; }
            vs[i] = v;
        }
        if (lane == 0) { tau[kk] = tk; e[kk] = beta; }
        __syncwarp();
        // p = tau A v (trailing block), w = p - (tau/2)(p.v) v
        double pv = 0.0;
        for (int i = kk + 1 + lane; i < n; i += 32) {
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            const double *row = A + i * ld;
            for (int j = kk + 1; j < n; j += 8) {
                double r[8], v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const bool ok = j + u < n;
                    r[u] = ok ? row[j + u] : 0.0;
                    v[u] = ok ? vs[j + u] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) acc[u & 3] = fma(r[u], v[u], acc[u & 3]);
            }
            const double pi = tk * ((acc[0] + acc[1]) + (acc[2] + acc[3]));
            ws[i] = pi;
            pv += pi * vs[i];
        }
        pv = et_sum(pv);
        __syncwarp();
        const double hk = 0.5 * tk * pv;
        for (int i = kk + 1 + lane; i < n; i += 32) ws[i] -= hk * vs[i];
        __syncwarp();
        for (int i = kk + 1 + lane; i < n; i += 32) {
            double *row = A + i * ld;
            const double vi = vs[i], wi = ws[i];
            for (int j = kk + 1; j < n; j += 8) {
                double r[8], v[8], w[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const bool ok = j + u < n;
                    r[u] = ok ? row[j + u] : 0.0;
                    v[u] = ok ? vs[j + u] : 0.0;
                    w[u] = ok ? ws[j + u] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) r[u] -= vi * w[u] + wi * v[u];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (j + u < n) row[j + u] = r[u];
            }
        }
        __syncwarp();
    }
    if (lane == 0) {
        if (n >= 2) { d[n - 2] = A[(n - 2) * ld + n - 2]; e[n - 2] = A[(n - 1) * ld + n - 2]; }
        d[n - 1] = A[(n - 1) * ld + n - 1];
    }
    __syncwarp();

    // 2. scale T to unit size; Gershgorin interval
    double mx = 0.0;
    for (int i = lane; i < n; i += 32) mx = fmax(mx, fmax(fabs(d[i]), i < n - 1 ? fabs(e[i]) : 0.0));
    mx = et_max(mx);
    const double scale = mx > 0.0 ? mx : 1.0, iscale = 1.0 / scale;
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        d[i] *= iscale;
        const double ei = i < n - 1 ? e[i] * iscale : 0.0;
        e[i] = ei;
        e2[i] = ei * ei;
    }
    __syncwarp();
    double glo = 1e300, ghi = -1e300;
    for (int i = lane; i < n; i += 32) {
        const double r = (i > 0 ? fabs(e[i - 1]) : 0.0) + (i < n - 1 ? fabs(e[i]) : 0.0);
        glo = fmin(glo, d[i] - r);
        ghi = fmax(ghi, d[i] + r);
    }
    glo = et_min(glo);
    ghi = et_max(ghi);
    for (int i = lane; i < n8; i += 32) {
        ds[i] = i < n ? d[i] : 0.0;
        es2[i] = (i > 0 && i < n) ? e2[i - 1] : 0.0;
    }
    __syncwarp();
    {
        const double tn = fmax(fabs(glo), fabs(ghi)), pad = 2.0 * 2.220446049250313e-16 * tn * n + 2e-300;
        glo -= pad;
        ghi += pad;
    }
    // 3. multisection: lane group m (L lanes) brackets the m-th largest eigenvalue.  Sturm count
    //    = sign changes of the leading-minor determinants p_i = (d_i - x) p_{i-1} - e_{i-1}^2
    //    p_{i-2}: one dependent FMA per row, no division; the pair (p_i, p_{i-1}) is rescaled
    //    every eight rows (|d - x| <= 4 and the pair cannot shrink below 1e-136 per block)
    int kp = 1;
    while (kp < k) kp <<= 1;
    const int L = 32 / kp;
    const int m_of = min(lane / L, k - 1), l_of = lane % L, base = (lane / L) * L;
    const unsigned gmask = L == 32 ? 0xffffffffu : ((1u << L) - 1u);
    int passes = 2;
    { double w = 1.0; while (w < 1.5e17) { w *= (double)(L + 1); ++passes; } }
    {
        const int jt = n - 1 - m_of;                          // ascending index wanted
        double lo = glo, hi = ghi;
        for (int ps = 0; ps < passes; ++ps) {
            const double x = lo + (hi - lo) * ((double)(l_of + 1) / (double)(L + 1));
            int c = 0;
            double p2 = 0.0, p1 = 1.0;
            // ds / es2 are padded to a multiple of 8 rows (zeros): the recurrence may run on,
            // only the count is masked
            for (int i0 = 0; i0 < n; i0 += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const double t = es2[i0 + u] * p2;
                    double p = fma(ds[i0 + u] - x, p1, -t);
                    if (p == 0.0) p = -copysign(1e-300, p1);
                    c += (i0 + u < n) & ((unsigned)(__double2hiint(p) ^ __double2hiint(p1)) >> 31);
                    p2 = p1;
                    p1 = p;
                }
                if (fabs(p1) < 1e-120 && fabs(p2) < 1e-120) { p1 *= 1e120; p2 *= 1e120; }
                else if (fabs(p1) > 1e120) { p1 *= 1e-120; p2 *= 1e-120; }
            }
            const unsigned bal = (__ballot_sync(SS_FULL, c > jt) >> base) & gmask;
            const int f = bal ? __ffs(bal) - 1 : L;
            const double xl = __shfl_sync(SS_FULL, x, base + (f > 0 ? f - 1 : 0));
            const double xh = __shfl_sync(SS_FULL, x, base + (f < L ? f : 0));
            if (f > 0) lo = xl;
            if (f < L) hi = xh;
        }
        if (l_of == 0 && lane / L < k) lam[m_of] = 0.5 * (lo + hi);
    }
    __syncwarp();
    // 4. inverse iteration on the scaled T
    const double tol = 2.220446049250313e-16;                 // eps * |T|, |T| = 1
    if (lane == 0)
        for (int m = 0; m < k; ++m) {
            double lm = lam[m];
            if (m > 0 && fabs(lam[m] - lam[m - 1]) <= 10.0 * tol) lm = lamu[m - 1] - 10.0 * tol;
            lamu[m] = lm;
        }
    __syncwarp();
    if (lane < k) {
        // T - lam I = P L U (partial pivoting); ra = 1 / pivots (tiny ones perturbed to eps as
        // dlagts job -1 does), b and d2 = super-diagonals of U, c = multipliers
        double *ra = LU + (size_t)lane * 4 * n, *b = ra + n, *c = b + n, *d2 = c + n;
        int *pv = piv + (size_t)lane * n;
        const double lm = lamu[lane];
        double ak = d[0] - lm;                                // pivot candidate of row q
        double bq = e[0];                                     // b[q] (possibly modified by row q-1)
        for (int q = 0; q < n - 1; ++q) {
            const double cq = e[q], an0 = d[q + 1] - lm, bn0 = e[q + 1];   // bn0 = 0 at q = n-2
            double piv_v, mult, anext, bnext = bn0, bstore, d2q = 0.0;
            int swapped;
            if (fabs(ak) >= fabs(cq)) {
                swapped = 0;
                piv_v = ak;
                if (fabs(piv_v) < tol) piv_v = piv_v < 0.0 ? -tol : tol;
                const double r = 1.0 / piv_v;
                mult = ak != 0.0 ? cq * r : 0.0;
                anext = an0 - mult * bq;
                bstore = bq;
                ra[q] = r;
            } else {
                swapped = 1;
                piv_v = cq;                                   // |cq| > |ak| >= 0: not tiny unless ak is
                if (fabs(piv_v) < tol) piv_v = piv_v < 0.0 ? -tol : tol;
                const double r = 1.0 / piv_v;
                mult = ak * r;
                anext = bq - mult * an0;
                d2q = bn0;
                bnext = -mult * bn0;
                bstore = an0;
                ra[q] = r;
            }
            pv[q] = swapped;
            c[q] = mult;
            b[q] = bstore;
            d2[q] = d2q;
            ak = anext;
            bq = bnext;
        }
        if (fabs(ak) < tol) ak = ak < 0.0 ? -tol : tol;
        ra[n - 1] = 1.0 / ak;
        b[n - 1] = 0.0;
        d2[n - 1] = 0.0;
        // start vectors: a fixed hash of (row, eigenvalue index) - different per eigenvalue, so
        // that a cluster (or a diagonal T) still yields independent vectors (dstein: random)
        double *z = Z + (size_t)lane * n;
        for (int i = 0; i < n; ++i) {
            unsigned h = (unsigned)(i + 1) * 2654435761u ^ (unsigned)(lane + 1) * 2246822519u;
            h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
            z[i] = (double)(h >> 8) * (1.0 / 16777216.0) + 0.25;
        }
    }
    __syncwarp();
    for (int it = 0; it < 3; ++it) {
        if (lane < k) {
            const double *ra = LU + (size_t)lane * 4 * n, *b = ra + n, *c = b + n, *d2 = c + n;
            const int *pv = piv + (size_t)lane * n;
            double *z = Z + (size_t)lane * n;
            double zq = z[0];                                 // carried in a register
            for (int q = 0; q < n - 1; ++q) {
                const double zn = z[q + 1], cq = c[q];
                if (pv[q] == 0) { z[q] = zq; zq = zn - cq * zq; }
                else { z[q] = zn; zq = zq - cq * zn; }
            }
            z[n - 1] = zq;
            double big = 0.0, y1 = 0.0, y2 = 0.0;             // z[q+1], z[q+2] of the back substitution
            for (int q = n - 1; q >= 0; --q) {
                const double t = (z[q] - b[q] * y1 - d2[q] * y2) * ra[q];
                z[q] = t;
                big = fmax(big, fabs(t));
                y2 = y1;
                y1 = t;
            }
            const double ib = big > 0.0 ? 1.0 / big : 1.0;
            for (int q = 0; q < n; ++q) z[q] *= ib;
        }
        __syncwarp();
        for (int m = 0; m < k; ++m) {                         // Gram-Schmidt inside clusters, normalise
            double *zm = Z + (size_t)m * n;
            for (int j = 0; j < m; ++j) {
                if (fabs(lam[m] - lam[j]) < 1e-3) {
                    const double *zj = Z + (size_t)j * n;
                    double dot = 0.0;
                    for (int i = lane; i < n; i += 32) dot += zm[i] * zj[i];
                    dot = et_sum(dot);
                    for (int i = lane; i < n; i += 32) zm[i] -= dot * zj[i];
                    __syncwarp();
                }
            }
            double nn = 0.0;
            for (int i = lane; i < n; i += 32) nn += zm[i] * zm[i];
            nn = et_sum(nn);
            const double inn = nn > 0.0 ? rsqrt(nn) : 0.0;
            for (int i = lane; i < n; i += 32) zm[i] *= inn;
            __syncwarp();
        }
    }
    // 5. back-transformation u = H_0 H_1 ... H_{n-3} z, two vectors per pass (their reductions
    //    interleave)
    for (int m0 = 0; m0 < k; m0 += 2) {
        double *za = Z + (size_t)m0 * n, *zb = Z + (size_t)(m0 + 1 < k ? m0 + 1 : m0) * n;
        const bool two = m0 + 1 < k;
        for (int kk = n - 3; kk >= 0; --kk) {
            const double tk = tau[kk];
            if (tk == 0.0) continue;
            double sa = 0.0, sb = 0.0;
            for (int i = kk + 1 + lane; i < n; i += 32) {
                const double v = i == kk + 1 ? 1.0 : A[i * ld + kk];
                sa = fma(v, za[i], sa);
                sb = fma(v, zb[i], sb);
            }
#pragma unroll
            for (int dd = 16; dd > 0; dd >>= 1) {
                sa += __shfl_xor_sync(SS_FULL, sa, dd);
                sb += __shfl_xor_sync(SS_FULL, sb, dd);
            }
            sa *= tk;
            sb *= tk;
            for (int i = kk + 1 + lane; i < n; i += 32) {
                const double v = i == kk + 1 ? 1.0 : A[i * ld + kk];
                const double a = za[i] - sa * v, b = zb[i] - sb * v;
                za[i] = a;
                if (two) zb[i] = b;
            }
            __syncwarp();
        }
    }
    // 6. columns 0..k-1 hold the leading pairs (descending); the rest is zero
    for (int j = lane; j < n; j += 32) evals[g * n + j] = j < k ? fabs(lam[j] * scale) : 0.0;
#pragma unroll 4
    for (int idx = lane; idx < n * n; idx += 32) {
        const int i = idx / n, j = idx - i * n;
        evecs[g * n * n + idx] = j < k ? Z[(size_t)j * n + i] : 0.0;
    }
}

// scores: thread per (spike, contact).  NC = ncomps at compile time (0: run-time value, any
// count up to PC_MAX_COMPS); the leading eigenvector columns (pre-divided by sqrt(eval)) sit in
// shared memory, n_win x NC per group of the block.
template <typename T, int NC>
__global__ void __launch_bounds__(256)
pca_project_kernel(const T *__restrict__ waves, int n_win, int64_t n_spk, int n_c,
                   const int64_t *__restrict__ offsets, int64_t n_groups,
                   const double *__restrict__ evals, const double *__restrict__ evecs, int ncomps_rt,
                   double *__restrict__ scores)
{
    const int ncomps = NC > 0 ? NC : ncomps_rt;
    constexpr int NCM = NC > 0 ? NC : PC_MAX_COMPS;
    extern __shared__ double pj_v[];                      // [groups of the block][n_win][ncomps]
    const int64_t idx0 = (int64_t)blockIdx.x * blockDim.x;
    const int64_t idx = idx0 + threadIdx.x;               // (s, c) flattened
    // groups touched by this block: contact mode all n_c; offsets mode [g_lo, g_hi]
    int64_t g_lo = 0, g_cnt = n_c;
    if (offsets) {
        auto group_of = [&](int64_t sp) {
            int64_t lo = 0, hi = n_groups;
            while (hi - lo > 1) {
                const int64_t mid = (lo + hi) >> 1;
                if (__ldg(offsets + mid) <= sp) lo = mid; else hi = mid;
            }
            return lo;
        };
        const int64_t s_last = min(idx0 + (int64_t)blockDim.x, n_spk) - 1;
        g_lo = group_of(idx0);
        g_cnt = group_of(s_last) - g_lo + 1;
    }
    const bool staged = g_cnt <= 8;                        // else straight from global memory
    if (staged) {
        const int per = n_win * ncomps;
        for (int e = threadIdx.x; e < (int)g_cnt * per; e += blockDim.x) {
            const int gl = e / per, rem = e - gl * per, w = rem / ncomps, j = rem - w * ncomps;
            const int64_t g = g_lo + gl;
            pj_v[e] = __ldg(evecs + (g * n_win + w) * n_win + j) / sqrt(__ldg(evals + g * n_win + j));
        }
    }
    __syncthreads();
    if (idx >= n_spk * n_c) return;
    const int64_t s = idx / n_c;
    const int c = (int)(idx - s * n_c);
    int64_t g = c;
    if (offsets) {
        int64_t lo = 0, hi = n_groups;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(offsets + mid) <= s) lo = mid; else hi = mid;
        }
        g = lo;
        if (s >= __ldg(offsets + n_groups)) return;          // spare capacity, not a spike
    }
    double acc[NCM];
#pragma unroll
    for (int j = 0; j < NCM; ++j) acc[j] = 0.0;
    const int64_t step = n_spk * (int64_t)n_c;
    double *o = scores + idx * ncomps;
    // the n_win loads of a thread are 25-31 independent streams n_spk*n_c elements apart: issue
    // them eight at a time (one at a time the kernel ran at 1.5 TB/s on 10^7 spikes - one
    // 128-byte request per warp in flight)
    constexpr int PU = 16;
    if (staged) {
        const double *V = pj_v + (g - g_lo) * n_win * ncomps;
        for (int w0 = 0; w0 < n_win; w0 += PU) {
            T raw[PU];
#pragma unroll
            for (int u = 0; u < PU; ++u)
                raw[u] = w0 + u < n_win ? __ldg(waves + (int64_t)(w0 + u) * step + idx) : T(0);
#pragma unroll
            for (int u = 0; u < PU; ++u) {
                if (w0 + u < n_win) {
                    const double xv = ss_to_f64(raw[u]);
#pragma unroll
                    for (int j = 0; j < NCM; ++j)
                        if (j < ncomps) acc[j] = fma(V[(w0 + u) * ncomps + j], xv, acc[j]);
                }
            }
        }
        // (v / sqrt(eval)) . x instead of (v . x) / sqrt(eval): one rounding apart
#pragma unroll
        for (int j = 0; j < NCM; ++j)
            if (j < ncomps) o[j] = acc[j];
        return;
    }
    const double *V = evecs + g * n_win * n_win;
    for (int w = 0; w < n_win; ++w) {
        const double xv = ss_to_f64(__ldg(waves + (int64_t)w * step + idx));
#pragma unroll
        for (int j = 0; j < NCM; ++j)
            if (j < ncomps) acc[j] += __ldg(V + w * n_win + j) * xv;
    }
#pragma unroll
    for (int j = 0; j < NCM; ++j)
        if (j < ncomps) o[j] = acc[j] / sqrt(__ldg(evals + g * n_win + j));
}

}  // namespace

extern "C" int ss_max_pca_window(void) { return PC_MAX_WIN; }

extern "C" size_t ss_pca_workspace_bytes(int n_win, int64_t n_groups)
{
    if (n_win <= 0 || n_win > PC_MAX_WIN || n_groups <= 0) return 0;
    const int nwp = pc_pad4(n_win);
    return (size_t)n_groups * pc_parts(n_groups) * ((size_t)nwp * nwp + nwp) * sizeof(double);
}

extern "C" int ss_pca_gram(const void *d_waves, int dtype, int n_win, int64_t n_spk, int n_c,
                           const int64_t *d_offsets, int64_t n_groups, const double *d_ref,
                           double *d_gram, double *d_sum, int mode, void *d_ws, size_t ws_bytes,
                           void *stream)
{
    SS_REQUIRE((d_waves || n_spk == 0) && d_ref && d_gram && d_sum && d_ws, SS_ERR_ARG, "pca_gram: null pointer");
    SS_REQUIRE(dtype >= SS_I16 && dtype <= SS_F64, SS_ERR_ARG, "pca_gram: bad dtype %d", dtype);
    SS_REQUIRE(mode >= SS_GRAM_AUTO && mode <= SS_GRAM_TC, SS_ERR_ARG, "pca_gram: bad mode %d", mode);
    SS_REQUIRE(n_win >= 1 && n_win <= PC_MAX_WIN, SS_ERR_UNSUPPORTED, "pca_gram: n_win %d outside [1, %d]", n_win, PC_MAX_WIN);
    SS_REQUIRE(n_spk >= 0 && n_c >= 1 && n_groups >= 1 && n_groups <= 65535, SS_ERR_ARG, "pca_gram: bad shape");
    SS_REQUIRE(d_offsets ? n_c == 1 : n_groups == n_c, SS_ERR_ARG, "pca_gram: group layout mismatch");
    SS_REQUIRE(ws_bytes >= ss_pca_workspace_bytes(n_win, n_groups), SS_ERR_WORKSPACE, "pca_gram: workspace too small");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int parts = pc_parts(n_groups), nwp = pc_pad4(n_win);
    double *ws_gram = reinterpret_cast<double *>(d_ws);
    double *ws_sum = ws_gram + (size_t)n_groups * parts * nwp * nwp;
    const size_t smem = (size_t)PC_SC * (nwp + 2) * sizeof(double);
    const int smem_max = (int)((size_t)PC_SC * (PC_MAX_WIN + 2) * sizeof(double));
    SS_CUDA_CHECK(cudaFuncSetAttribute(pca_gram_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    SS_CUDA_CHECK(cudaFuncSetAttribute(pca_gram_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    SS_CUDA_CHECK(cudaFuncSetAttribute(pca_gram_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    // Large float32 problems go to the tensor cores (3xTF32 with short float32 accumulation
    // chunks, ~1e-7 relative on the covariance); small ones and float64 / int16 blocks stay on
    // the float64 CUDA-core kernel (exact products, ~1e-15).  The choice depends on the SHAPE
    // of the call only (n_spk is the capacity in the speculative chain), so ss_pca_gram_mode
    // lets the caller pin it: the chain pins the float64 kernel unless told otherwise.
    // SS_PCA_TC=0 / 1 forces the choice for A/B runs.
    const char *force = getenv("SS_PCA_TC");
    const int64_t per_group = d_offsets ? n_spk / (n_groups > 0 ? n_groups : 1) : n_spk;
    const bool tc_ok = dtype == SS_F32 && ssi::pca_gram_tc_eligible(n_win, n_spk, n_c, d_offsets != nullptr) &&
                       (reinterpret_cast<uintptr_t>(d_waves) % 16 == 0);
    bool use_tc = tc_ok && (mode == SS_GRAM_TC || (mode == SS_GRAM_AUTO && per_group >= 65536));
    if (force && force[0] == '0') use_tc = false;
    if (force && force[0] == '1') use_tc = tc_ok;
    if (use_tc) {
        const int rc = ssi::pca_gram_tc(reinterpret_cast<const float *>(d_waves), n_win, n_spk, n_c,
                                        d_offsets, n_groups, d_ref, ws_gram, ws_sum, parts, nwp, st);
        if (rc) return rc;
    } else {
        // unused upper-triangle tiles of partials are never read; no memset needed
        const dim3 grid(parts, (unsigned)n_groups);
        if (dtype == SS_F32)
            pca_gram_kernel<float><<<grid, PC_THREADS, smem, st>>>(
                reinterpret_cast<const float *>(d_waves), n_win, n_spk, n_c, d_offsets, d_ref, ws_gram, ws_sum);
        else if (dtype == SS_F64)
            pca_gram_kernel<double><<<grid, PC_THREADS, smem, st>>>(
                reinterpret_cast<const double *>(d_waves), n_win, n_spk, n_c, d_offsets, d_ref, ws_gram, ws_sum);
        else
            pca_gram_kernel<int16_t><<<grid, PC_THREADS, smem, st>>>(
                reinterpret_cast<const int16_t *>(d_waves), n_win, n_spk, n_c, d_offsets, d_ref, ws_gram, ws_sum);
        SS_LAUNCH_CHECK("pca_gram_kernel");
    }
    pca_gram_reduce_kernel<<<(unsigned)n_groups, PC_THREADS, 0, st>>>(ws_gram, ws_sum, n_win, parts, d_gram, d_sum);
    SS_LAUNCH_CHECK("pca_gram_reduce_kernel");
    return SS_OK;
}

extern "C" int ss_pca_eig(const double *d_gram, const double *d_sum, const int64_t *d_counts,
                          int n_win, int64_t n_groups, double *d_evals, double *d_evecs, void *stream)
{
    SS_REQUIRE(d_gram && d_sum && d_counts && d_evals && d_evecs, SS_ERR_ARG, "pca_eig: null pointer");
    SS_REQUIRE(n_win >= 1 && n_win <= PC_MAX_WIN, SS_ERR_UNSUPPORTED, "pca_eig: n_win %d outside [1, %d]", n_win, PC_MAX_WIN);
    SS_REQUIRE(n_groups >= 1, SS_ERR_ARG, "pca_eig: bad shape");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int ne = (n_win + 1) & ~1;
    const size_t smem = ((size_t)2 * ne * (ne | 1) + ne + 2 * (PC_THREADS / 32) + 8) * sizeof(double);
    const int ne_max = (PC_MAX_WIN + 1) & ~1;
    SS_CUDA_CHECK(cudaFuncSetAttribute(pca_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(((size_t)2 * ne_max * (ne_max | 1) + ne_max + 2 * (PC_THREADS / 32) + 8) * sizeof(double))));
    const int eig_threads = PC_THREADS;
    pca_eig_kernel<<<(unsigned)n_groups, eig_threads, smem, st>>>(d_gram, d_sum, d_counts, n_win, d_evals, d_evecs);
    SS_LAUNCH_CHECK("pca_eig_kernel");
    return SS_OK;
}

extern "C" int ss_pca_eig_top(const double *d_gram, const double *d_sum, const int64_t *d_counts,
                              int n_win, int64_t n_groups, int ncomps, double *d_evals,
                              double *d_evecs, void *stream)
{
    SS_REQUIRE(d_gram && d_sum && d_counts && d_evals && d_evecs, SS_ERR_ARG, "pca_eig_top: null pointer");
    SS_REQUIRE(n_win >= 1 && n_win <= PC_MAX_WIN, SS_ERR_UNSUPPORTED, "pca_eig_top: n_win %d outside [1, %d]", n_win, PC_MAX_WIN);
    SS_REQUIRE(ncomps >= 1 && ncomps <= ET_MAXK && ncomps <= n_win, SS_ERR_UNSUPPORTED,
               "pca_eig_top: ncomps %d outside [1, min(%d, n_win)]", ncomps, ET_MAXK);
    SS_REQUIRE(n_groups >= 1, SS_ERR_ARG, "pca_eig_top: bad shape");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = et_smem_doubles(n_win, ncomps) * sizeof(double);
    SS_CUDA_CHECK(cudaFuncSetAttribute(pca_eig_top_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(et_smem_doubles(PC_MAX_WIN, ET_MAXK) * sizeof(double))));
    pca_eig_top_kernel<<<(unsigned)n_groups, 32, smem, st>>>(d_gram, d_sum, d_counts, n_win, ncomps, d_evals, d_evecs);
    SS_LAUNCH_CHECK("pca_eig_top_kernel");
    return SS_OK;
}

// K8 for big float32 blocks in contact mode (fetPCA on (n_win, n_spk, n_c) blocks): the scalar
// kernel above keeps 16 x 128 B per warp in flight and was latency-bound on 10^7 spikes (ncu:
// long_scoreboard 12.8 per issue, 42-54 % of the HBM peak).  Here a thread owns FOUR consecutive
// (spike, contact) elements: one 128-bit load per time row (512 B per warp and request, PV rows
// in flight = 4 KB per warp), and its 4 x ncomps scores leave as 128-bit stores.  Same products
// in the same order as the scalar kernel (bit-identical scores).  Needs n_spk * n_c % 4 == 0 (row
// starts 16-byte aligned) and 16-byte aligned pointers; the launcher checks.
#ifndef PJ_PV
#define PJ_PV 4                            // time rows in flight per thread (A/B on B200, 10^7 spikes: 4 rows x 4 CTAs 0.79 ms, 6 x 3 0.80, 8 x 2 0.94, 12 x 2 0.98)
#define PJ_CTAS 4                          // resident CTAs the kernel is compiled for
#endif
template <int NC>
__global__ void __launch_bounds__(256, PJ_CTAS)
pca_project_vec4_kernel(const float *__restrict__ waves, int n_win, int64_t total, int n_c,
                        const double *__restrict__ evals, const double *__restrict__ evecs,
                        double *__restrict__ scores)
{
    extern __shared__ double pj_v[];                      // [n_c][n_win][NC]
    for (int e = threadIdx.x; e < n_c * n_win * NC; e += blockDim.x) {
        const int g = e / (n_win * NC), rem = e - g * (n_win * NC), w = rem / NC, j = rem - w * NC;
        pj_v[e] = __ldg(evecs + ((int64_t)g * n_win + w) * n_win + j) / sqrt(__ldg(evals + (int64_t)g * n_win + j));
    }
    __syncthreads();
    const int64_t e0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (e0 >= total) return;
    const double *V[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) V[e] = pj_v + (int)((e0 + e) % n_c) * n_win * NC;
    double acc[4][NC];
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int j = 0; j < NC; ++j) acc[e][j] = 0.0;
    constexpr int PV = PJ_PV;
    const float4 *src = reinterpret_cast<const float4 *>(waves + e0);
    const int64_t step4 = total / 4;
    for (int w0 = 0; w0 < n_win; w0 += PV) {
        float4 raw[PV];
#pragma unroll
        for (int u = 0; u < PV; ++u)
            raw[u] = w0 + u < n_win ? __ldg(src + (int64_t)(w0 + u) * step4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < PV; ++u) {
            if (w0 + u < n_win) {
                const double xv[4] = {(double)raw[u].x, (double)raw[u].y, (double)raw[u].z, (double)raw[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int j = 0; j < NC; ++j) acc[e][j] = fma(V[e][(w0 + u) * NC + j], xv[e], acc[e][j]);
            }
        }
    }
    double *o = scores + e0 * NC;
    if (NC == 2) {
#pragma unroll
        for (int e = 0; e < 4; ++e) reinterpret_cast<double2 *>(o)[e] = make_double2(acc[e][0], acc[e][1]);
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int j = 0; j < NC; ++j) o[e * NC + j] = acc[e][j];
    }
}

extern "C" int ss_pca_project(const void *d_waves, int dtype, int n_win, int64_t n_spk, int n_c,
                              const int64_t *d_offsets, int64_t n_groups, const double *d_evals,
                              const double *d_evecs, int ncomps, double *d_scores, void *stream)
{
    if (n_spk == 0) return SS_OK;
    SS_REQUIRE(d_waves && d_evals && d_evecs && d_scores, SS_ERR_ARG, "pca_project: null pointer");
    SS_REQUIRE(n_win >= 1 && n_win <= PC_MAX_WIN, SS_ERR_UNSUPPORTED, "pca_project: n_win %d outside [1, %d]", n_win, PC_MAX_WIN);
    SS_REQUIRE(ncomps >= 1 && ncomps <= PC_MAX_COMPS && ncomps <= n_win, SS_ERR_UNSUPPORTED, "pca_project: ncomps %d outside [1, %d]", ncomps, PC_MAX_COMPS);
    SS_REQUIRE(d_offsets ? n_c == 1 : n_groups == n_c, SS_ERR_ARG, "pca_project: group layout mismatch");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t total = n_spk * n_c;
    SS_REQUIRE(dtype >= SS_I16 && dtype <= SS_F64, SS_ERR_ARG, "pca_project: bad dtype %d", dtype);
    // big float32 blocks in contact mode: four elements per thread, 128-bit loads and stores
    if (dtype == SS_F32 && !d_offsets && ncomps <= 3 && n_c <= 8 && total >= (1 << 20) && total % 4 == 0 &&
        (reinterpret_cast<uintptr_t>(d_waves) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_scores) & 15) == 0) {
        const unsigned vblocks = (unsigned)ss_cdiv(total / 4, 256);
        const size_t vsmem = (size_t)n_c * n_win * ncomps * sizeof(double);
        const float *wv = reinterpret_cast<const float *>(d_waves);
        if (ncomps == 1)
            pca_project_vec4_kernel<1><<<vblocks, 256, vsmem, st>>>(wv, n_win, total, n_c, d_evals, d_evecs, d_scores);
        else if (ncomps == 2)
            pca_project_vec4_kernel<2><<<vblocks, 256, vsmem, st>>>(wv, n_win, total, n_c, d_evals, d_evecs, d_scores);
        else
            pca_project_vec4_kernel<3><<<vblocks, 256, vsmem, st>>>(wv, n_win, total, n_c, d_evals, d_evecs, d_scores);
        SS_LAUNCH_CHECK("pca_project_vec4_kernel");
        return SS_OK;
    }
    const unsigned blocks = (unsigned)ss_cdiv(total, 256);
    const size_t smem = (size_t)8 * n_win * ncomps * sizeof(double);
#define SS_PROJECT(T_, NC_)                                                                        \
    pca_project_kernel<T_, NC_><<<blocks, 256, smem, st>>>(reinterpret_cast<const T_ *>(d_waves), n_win, \
        n_spk, n_c, d_offsets, n_groups, d_evals, d_evecs, ncomps, d_scores)
#define SS_PROJECT_T(T_)                                                                           \
    do {                                                                                           \
        if (ncomps == 1) SS_PROJECT(T_, 1);                                                        \
        else if (ncomps == 2) SS_PROJECT(T_, 2);                                                   \
        else if (ncomps == 3) SS_PROJECT(T_, 3);                                                   \
        else SS_PROJECT(T_, 0);                                                                    \
    } while (0)
    if (dtype == SS_F32) SS_PROJECT_T(float);
    else if (dtype == SS_F64) SS_PROJECT_T(double);
    else SS_PROJECT_T(int16_t);
#undef SS_PROJECT_T
#undef SS_PROJECT
    SS_LAUNCH_CHECK("pca_project_kernel");
    return SS_OK;
}
