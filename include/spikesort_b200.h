/*
 * spikesort_b200 - C ABI of the B200-native SpikeSort front end
 * (filter -> threshold -> detect -> align/extract -> PCA).
 *
 * This is the drop-in boundary for the hot path of btel/SpikeSort
 * (reference paths relative to src/spike_sort/core/).  The reference has no
 * FFI of its own: its hot path is NumPy/SciPy calls inside
 *   filters.py:99-143   Filter.__call__ / filter_proxy
 *   extract.py:43-94    detect_spikes
 *   extract.py:97-184   filter_spt / extract_spikes
 *   extract.py:208-288  align_spikes / remove_doubles
 *   features.py:200-232 PCA, features.py:310-346 fetPCA
 * Each entry point below names the lines it replaces.  A ctypes binding (what a
 * SpikeSort maintainer would add) is shown in INTEGRATION.md and implemented in
 * spikesort_b200/_lib.py.
 *
 * Conventions
 *   - plain C: pointers and sizes only, no C++/torch types;
 *   - every pointer named d_* is DEVICE memory owned by the caller, everything
 *     else is host memory; the library never allocates device memory, with one
 *     exception: the weight table of each distinct filter (<= 100 KB) is cached on every
 *     device that uses it, for the life of the process;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); all
 *     work is enqueued on it and NO call synchronises the host;
 *   - return value: 0 on success, negative ss_status on failure, message via
 *     ss_last_error() (thread-local);
 *   - recordings are row-major (n_contacts, n_samples) with a row stride `ld`
 *     in elements, dtype one of ss_dtype;
 *   - several contacts ("rows") are processed by one call; per-row results are
 *     concatenated and delimited by an offsets array of n_rows+1 int64.
 */
#ifndef SPIKESORT_B200_H
#define SPIKESORT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SS_VERSION 100            /* 0.1.0 */

typedef enum { SS_I16 = 0, SS_F32 = 1, SS_F64 = 2 } ss_dtype;
typedef enum { SS_GRAM_AUTO = 0, SS_GRAM_F64 = 1, SS_GRAM_TC = 2 } ss_gram_mode;

typedef enum {
    SS_OK = 0,
    SS_ERR_ARG = -1,              /* bad argument                                        */
    SS_ERR_CUDA = -2,             /* CUDA runtime error (message has the detail)         */
    SS_ERR_PADLEN = -3,           /* a chunk is not longer than padlen (SciPy ValueError) */
    SS_ERR_WORKSPACE = -4,        /* workspace too small                                 */
    SS_ERR_UNSUPPORTED = -5       /* filter order / window size outside compiled range   */
} ss_status;

int         ss_version(void);
const char *ss_last_error(void);
/* Kernels of this library launched by the process so far (every launch site counts itself):
 * bench.py reports the difference over its timed region as gpu_launches. */
int64_t ss_launch_count(void);

/* Time slab of every contact between a (pinned) host recording and HBM as one asynchronous
 * 2-D DMA (rows = contacts, pitches in bytes); to_device 1 = host -> device, 0 = back. */
int ss_copy_2d_async(void *dst, size_t dst_pitch, const void *src, size_t src_pitch,
                     size_t width_bytes, size_t rows, int to_device, void *stream);
/* limits compiled into the library */
int         ss_max_filter_order(void);      /* total number of delay states     */
int         ss_max_window(void);            /* largest n_win for align/extract  */
int         ss_max_pca_window(void);        /* largest n_win for the PCA kernels */

/* ------------------------------------------------------------------ filter
 * Replaces Filter.__call__ -> scipy.signal.filtfilt (filters.py:99-101) looped
 * over (contact, chunk) by filter_proxy (filters.py:135-141): zero-phase IIR,
 * odd extension of `padlen` samples at both ends of EVERY chunk, steady-state
 * initial conditions, float64 output.
 *
 * The filter is given as a cascade of `n_sections` second-order sections in
 * SciPy's sos layout [b0 b1 b2 a0 a1 a2] (a0 must be 1; a first-order section
 * has b2 = a2 = 0 and must come last).  The Python host derives the sections
 * from the reference's (b, a) coefficients (spikesort_b200/_design.py).
 * padlen is the reference's 3*max(len(a), len(b)).
 *
 * Which kernels run is a property of the call (filter, dtype, geometry), never of the data.
 * Two environment switches exist for tests and A/B runs, read at every call:
 *   SS_APPLY_TMA = 0   second sweep of orders 1-2 on int16 through the element kernel only
 *                  1   (default) through the copy engine where the output has a tensor map
 *                  2   as 1, but SS_ERR_UNSUPPORTED instead of falling back
 *   SS_FILT_SPAN = 1   single-pass span kernel for filters that decay within one tile
 * All of them produce the same values (the copy-engine and element sweeps bit for bit).
 */
size_t ss_filtfilt_workspace_bytes(int64_t n_contacts, int64_t n_samples,
                                   int64_t chunksize, int padlen, int n_sections);
int ss_filtfilt_sos(const void *d_x, int dtype, int64_t n_contacts, int64_t n_samples,
                    int64_t ld_in, const double *sos, int n_sections, int padlen,
                    int64_t chunksize, double *d_out, int64_t ld_out,
                    void *d_ws, size_t ws_bytes, void *stream);

/* Fused K1 + K2 + K3(pass 1): the same filter, with the threshold crossings of the
 * float64 output marked in the epilogue of the filter's second sweep (no second pass
 * over the filtered signal).  Equivalent to ss_filtfilt_sos, [ss_var_threshold],
 * ss_detect_mark on the output - same d_out, same d_thr, same workspace contents and
 * d_offsets - so ss_detect_emit follows unchanged.
 *   auto_thr != 0 : d_thr[c] = factor*sqrt(var(out[c, :n0])) is computed (extract.py:79):
 *                   the chunks holding the first n0 samples are filtered first;
 *   auto_thr == 0 : d_thr is an input.
 * d_det_ws: ss_detect_workspace_bytes(n_contacts, n_samples).
 */
int ss_filtfilt_detect_sos(const void *d_x, int dtype, int64_t n_contacts, int64_t n_samples,
                           int64_t ld_in, const double *sos, int n_sections, int padlen,
                           int64_t chunksize, double *d_out, int64_t ld_out,
                           void *d_ws, size_t ws_bytes,
                           int auto_thr, int64_t n0, double factor, double *d_thr, int falling,
                           void *d_det_ws, size_t det_ws_bytes, int64_t *d_offsets,
                           int prepared, void *stream);

/* Split form of the above for time shards (the thresholds come from another rank):
 *   ss_filtfilt_prepare_sos        : sweep 1 + tile scan of every chunk into d_ws (needs
 *                                    no threshold);
 *   ss_filtfilt_head_threshold_sos : with d_ws prepared, filters the chunks holding the
 *                                    first n0 samples into d_out and computes d_thr;
 *   ss_filtfilt_detect_sos with prepared != 0 skips sweep 1.
 */
int ss_filtfilt_prepare_sos(const void *d_x, int dtype, int64_t n_contacts, int64_t n_samples,
                            int64_t ld_in, const double *sos, int n_sections, int padlen,
                            int64_t chunksize, void *d_ws, size_t ws_bytes, void *stream);
int ss_filtfilt_head_threshold_sos(const void *d_x, int dtype, int64_t n_contacts,
                                   int64_t n_samples, int64_t ld_in, const double *sos,
                                   int n_sections, int padlen, int64_t chunksize, double *d_out,
                                   int64_t ld_out, void *d_ws, size_t ws_bytes, int64_t n0,
                                   double factor, double *d_thr, void *d_thr_ws, size_t thr_ws_bytes,
                                   void *stream);

/* --------------------------------------------------------------- threshold
 * Replaces extract.py:79: thr[r] = factor * sqrt(var(x[r, :n0])) (population
 * variance, float64).  `factor` carries the sign (negative for falling edges,
 * extract.py:80-81).  d_ws: at least ss_threshold_workspace_bytes(n_rows).
 */
size_t ss_threshold_workspace_bytes(int64_t n_rows);
int ss_var_threshold(const void *d_x, int dtype, int64_t n_rows, int64_t ld, int64_t n0,
                     double factor, double *d_thr, void *d_ws, size_t ws_bytes, void *stream);

/* ------------------------------------------------------------------ detect
 * Replaces extract.py:86-93.  Row r has a crossing at i when
 *   rising : x[r,i] < thr[r] && x[r,i+1] > thr[r]      (falling: swapped)
 * and reports t = i*1000.0/FS (float64 ms), ascending per row.
 * ss_detect_mark   : one pass over the data -> crossing bitmask + counts in the
 *                    workspace, and d_offsets[n_rows+1] (exclusive prefix sum of
 *                    the per-row counts; last entry = total).
 * ss_detect_emit   : writes the total times into d_spt (capacity `cap`); reads
 *                    only the workspace.  The caller reads d_offsets in between
 *                    to size d_spt.  index_offset = index of x[.,0] in the whole
 *                    recording (0 unless x is a time shard, see "time shards").
 */
size_t ss_detect_workspace_bytes(int64_t n_rows, int64_t n_samples);
int ss_detect_mark(const void *d_x, int dtype, int64_t n_rows, int64_t n_samples, int64_t ld,
                   const double *d_thr, int falling, void *d_ws, size_t ws_bytes,
                   int64_t *d_offsets, void *stream);
int ss_detect_emit(int64_t n_rows, int64_t n_samples, double FS, int64_t index_offset,
                   const void *d_ws, size_t ws_bytes, double *d_spt, int64_t cap, void *stream);

/* -------------------------------------------------------------- time shards
 * A recording sharded by time across GPUs keeps GLOBAL spike times: the kernels that
 * turn indices into times or back take
 *   index_offset : index in the whole recording of the shard's sample x[., 0];
 *   halo_before / halo_after : number of valid samples stored immediately before
 *                  x[., 0] / after x[., n_samples-1] in the same rows (neighbour
 *                  shards' filtered samples), readable by windows that cross the edge.
 * t_min / t_max passed to align / extract are then those of the WHOLE recording.
 * All three are 0 for an unsharded recording.
 *   d_halo_miss : optional device int32 counter (NULL: none).  A window sample that lies
 *                  inside the recording but beyond the stored neighbour samples (an
 *                  alignment walk longer than the halo) cannot be read; the kernels use 0
 *                  for it and ADD the number of affected spikes (resampled alignment:
 *                  samples) here.  The counter is never reset by the library; a non-zero
 *                  value means the shard's result differs from the unsharded one and the
 *                  call must be repeated with a wider halo.
 */

/* ------------------------------------------------------------------- align
 * Replaces the loop of align_spikes (extract.py:235-265), resample == 1.
 * Spikes d_spt[d_offsets[r] .. d_offsets[r+1]) are aligned on row
 * d_rows[r] of the recording (d_rows NULL: row r).  For each spike, while
 * t_min <= t <= t_max: centre = (int32)(t/1000.0*FS); find the FIRST max (or
 * min) of float32(x[centre+win0 .. centre+win1)); t += k*1000.0/FS + sp_win0;
 * stop unless that shift is < lo or > hi.  All doubles are computed by the
 * host exactly as NumPy does (extract.py:108-109,151-152,229,263-264).
 */
int ss_align(const void *d_x, int dtype, int64_t n_samples, int64_t ld, double FS,
             const int32_t *d_rows, int64_t n_rows, const int64_t *d_offsets,
             double sp_win0, int win0, int win1, double t_min, double t_max,
             double lo, double hi, int use_min, double *d_spt, int64_t n_spk,
             int max_iter, int64_t index_offset, int64_t halo_before, int64_t halo_after,
             int32_t *d_halo_miss, void *stream);

/* Replaces remove_doubles (extract.py:277-288) per row: keeps the first time
 * of each row and every time with (int64)((t[k]-t[k-1])/tol) > 1.
 * mark: flags + d_offsets_out[n_rows+1]; emit: compacted times into d_out.
 * d_prev_last (may be NULL): per row, the last ALIGNED time of the previous time
 * shard (NaN: none) - the first time of a row is then compared with it, as np.diff
 * does on the unsharded list.
 */
size_t ss_remove_doubles_workspace_bytes(int64_t n_spk);
int ss_remove_doubles_mark(const double *d_spt, int64_t n_spk, const int64_t *d_offsets,
                           int64_t n_rows, double tol, const double *d_prev_last,
                           void *d_ws, size_t ws_bytes,
                           int64_t *d_offsets_out, void *stream);
int ss_remove_doubles_emit(const double *d_spt, int64_t n_spk, const void *d_ws,
                           size_t ws_bytes, double *d_out, int64_t cap, void *stream);

/* ----------------------------------------------------------------- extract
 * Replaces extract_spikes (extract.py:146-177), resample == 1.
 * Output d_waves is float32 (n_win, n_spk, n_c) C-contiguous, n_win = win1-win0.
 *   d_contacts != NULL : every spike is cut on the n_c listed contacts
 *                        (the reference's `contacts` argument);
 *   d_contacts == NULL : n_c must be 1 and spike s is cut on row
 *                        d_rows[r] (NULL: r) for d_offsets[r] <= s < d_offsets[r+1]
 *                        (the per-contact chain).
 * Windows leaving [0, n_samples) are clipped and zero-filled.
 * d_valid[s] = (t_min <= t <= t_max) (extract.py:147-148,174-177);
 * *d_n_invalid counts the invalid ones (0 -> the reference omits 'is_valid').
 */
int ss_extract(const void *d_x, int dtype, int64_t n_samples, int64_t ld, double FS,
               const int32_t *d_contacts, int n_c,
               const int32_t *d_rows, int64_t n_rows, const int64_t *d_offsets,
               int win0, int win1, double t_min, double t_max,
               const double *d_spt, int64_t n_spk,
               float *d_waves, uint8_t *d_valid, int32_t *d_n_invalid,
               int64_t index_offset, int64_t halo_before, int64_t halo_after,
               int32_t *d_halo_miss, void *stream);

/* --------------------------------------------------------------------- PCA
 * Replaces PCA / fetPCA (features.py:224-231, 326-344) on a waveform block
 * (n_win, n_spk, n_c) of element type `dtype` (ss_dtype): float32 from extract_spikes,
 * float64 from resample_spikes / extract_spikes(resample != 1), int16 exact - the reference
 * promotes every dtype to float64 (features.py:224) and so do the kernels.
 * A "group" is one PCA problem:
 *   d_offsets == NULL : n_groups = n_c, group g = contact g over all spikes
 *                       (fetPCA's loop, features.py:335-336);
 *   d_offsets != NULL : n_c must be 1, group g = spikes [d_offsets[g], d_offsets[g+1]).
 * ss_pca_gram    : d_gram[g] = sum_s (x_s - ref_g)(x_s - ref_g)^T (n_win x n_win,
 *                  float64), d_sum[g] = sum_s (x_s - ref_g).  d_ref[g][n_win] is an
 *                  INPUT: any waveform close to the mean (the host passes the first
 *                  spike of the group); it only conditions the arithmetic, the
 *                  covariance does not depend on it.  Partial Grams of time shards
 *                  computed with the same ref add up: this is what the multi-GPU
 *                  path all-reduces.  `mode` (ss_gram_mode) picks the arithmetic:
 *                  SS_GRAM_F64 exact float64 products on the CUDA cores; SS_GRAM_TC
 *                  3xTF32 on the tensor cores (tcgen05; float32 blocks with n_win <= 31
 *                  and <= 4 contacts, ~1e-7 relative on the covariance; silently
 *                  SS_GRAM_F64 when the block is not eligible); SS_GRAM_AUTO = TC from
 *                  65536 spikes per group on.  The choice depends on the call's shape
 *                  only, never on the data.
 * ss_pca_eig     : covariance (ddof=1) from gram/sum/count, symmetric Jacobi
 *                  eigensolve, eigenvalues |.| sorted descending, eigenvectors
 *                  in columns (features.py:225-229).  d_counts[g] = spikes per group.
 * ss_pca_eig_top : the same covariance, but only the ncomps LEADING eigenpairs (all that
 *                  fetPCA's scores need, features.py:230-231,335-336): Householder
 *                  tridiagonalisation, Sturm multisection, inverse iteration, one warp per
 *                  group.  Columns >= ncomps of d_evecs / entries >= ncomps of d_evals are
 *                  zero.  PCA()'s full (evals, evecs) still come from ss_pca_eig.
 * ss_pca_project : d_scores[s, g*ncomps + j] = (v_j . x_s) / sqrt(eval_j) on the
 *                  UN-centred data (features.py:230-231); (n_spk, n_groups*ncomps)
 *                  float64 when d_offsets == NULL, else (n_spk, ncomps).
 */
size_t ss_pca_workspace_bytes(int n_win, int64_t n_groups);
int ss_pca_gram(const void *d_waves, int dtype, int n_win, int64_t n_spk, int n_c,
                const int64_t *d_offsets, int64_t n_groups, const double *d_ref,
                double *d_gram, double *d_sum, int mode,
                void *d_ws, size_t ws_bytes, void *stream);
int ss_pca_eig(const double *d_gram, const double *d_sum,
               const int64_t *d_counts, int n_win, int64_t n_groups,
               double *d_evals, double *d_evecs, void *stream);
int ss_pca_eig_top(const double *d_gram, const double *d_sum,
                   const int64_t *d_counts, int n_win, int64_t n_groups, int ncomps,
                   double *d_evals, double *d_evecs, void *stream);
int ss_pca_project(const void *d_waves, int dtype, int n_win, int64_t n_spk, int n_c,
                   const int64_t *d_offsets, int64_t n_groups,
                   const double *d_evals, const double *d_evecs, int ncomps,
                   double *d_scores, void *stream);

/* ------------------------------------------------------- cheap features (8f)
 * Replaces fetP2P / fetMin (features.py:473-474, 483-484) on a float32 or float64
 * waveform block (n_win, n_spk, n_c): d_min[s, c] = min_w, d_p2p[s, c] = max_w - min_w, in
 * the block's own dtype (as np.min / np.ptp).  Either output may be NULL.
 * ss_normalize_columns replaces normalize (features.py:191-193) on a row-major float64
 * (n_rows, n_cols) matrix, in place: (x - colmin) / (colmax - colmin).  Both bit-exact.
 */
int ss_wave_minmax(const void *d_waves, int dtype, int n_win, int64_t n_spk, int n_c,
                   void *d_min, void *d_p2p, void *stream);
size_t ss_normalize_workspace_bytes(int n_cols);
int ss_normalize_columns(double *d_x, int64_t n_rows, int n_cols, void *d_ws, size_t ws_bytes,
                         void *stream);

/* ------------------------------------------------------- spline upsampling (8f)
 * Replaces resample_spikes (extract.py:187-205) and the resample != 1 branch of
 * align_spikes (extract.py:245-258): splrep(time, wave, s=0) + splev per waveform, i.e.
 * FITPACK's fpcurf/fpback/splev.  The waveform-independent part is a "plan" built by the
 * host in FITPACK's operation order (spikesort_b200/_spline.py), m = n_win:
 *   d_plan_f64 = [rot_cos (4m) | rot_sin (4m) | tri (4m, row-major m x 4) | ev_h (4 n_out)]
 *   d_plan_i32 = [rot_j (4m, -1 = no rotation) | ev_l (n_out)]
 * rot_*[4 it + q] is the q-th Givens rotation of data row `it` against coefficient rot_j;
 * tri the triangularised banded matrix; ev_l / ev_h the first coefficient and the four
 * B-spline values of every new abscissa.  The kernels replay the right-hand-side
 * operations unfused and in order: results are bit-identical to SciPy's.
 *   ss_resample        : d_out (n_out, n_spk, n_c) float64 from d_waves (n_win, n_spk, n_c)
 *                        of element type `dtype`;
 *   ss_align_resampled : as ss_align, the FIRST extremum searched on the upsampled
 *                        float64 window, shift = d_rtime[k] (the n_out new times, ms).
 */
int ss_resample(const void *d_waves, int dtype, int n_win, int64_t n_spk, int n_c,
                const double *d_plan_f64, const int32_t *d_plan_i32, int n_out,
                double *d_out, void *stream);
int ss_align_resampled(const void *d_x, int dtype, int64_t n_samples, int64_t ld, double FS,
                       const int32_t *d_rows, int64_t n_rows, const int64_t *d_offsets,
                       int win0, int win1, double t_min, double t_max, double lo, double hi,
                       int use_min, const double *d_plan_f64, const int32_t *d_plan_i32,
                       int n_out, const double *d_rtime, double *d_spt, int64_t n_spk,
                       int max_iter, int64_t index_offset, int64_t halo_before,
                       int64_t halo_after, int32_t *d_halo_miss, void *stream);

/* ------------------------------------------------------- zero-phase FIR (8f)
 * Replaces FilterFir.__call__ = scipy.signal.filtfilt(b, [1], x) (filters.py:41-74) for
 * every (contact, chunk) unit of filter_proxy (filters.py:135-141); b = remez taps designed
 * on the host.  Same chunking, odd padding (3*n_taps, in the input dtype) and error
 * (SS_ERR_PADLEN when a chunk is not longer than the padding) as ss_filtfilt_sos.
 */
int ss_filtfilt_fir(const void *d_x, int dtype, int64_t n_contacts, int64_t n_samples,
                    int64_t ld_in, const double *d_b, int n_taps, int64_t chunksize,
                    double *d_out, int64_t ld_out, void *stream);

/* ss_filtfilt_detect_sos for one time shard with the halo exchange FUSED into sweep 2: the
 * tiles that produce the shard's first / last H filtered samples of every contact store them,
 * besides d_out, straight into the left neighbour's `from_right` / the right neighbour's
 * `from_left` arena region (peer memory; NULL = no neighbour on that side).  Follow with
 * ss_peer_barrier; ss_shard_put_halos is then not needed. */
int ss_filtfilt_detect_sos_peer(const void *d_x, int dtype, int64_t n_contacts, int64_t n_samples,
                                int64_t ld_in, const double *sos, int n_sections, int padlen,
                                int64_t chunksize, double *d_out, int64_t ld_out, void *d_ws,
                                size_t ws_bytes, int auto_thr, int64_t n0, double factor,
                                double *d_thr, int falling, void *d_det_ws, size_t det_ws_bytes,
                                int64_t *d_offsets, int prepared, double *d_left_from_right,
                                double *d_right_from_left, int H, void *stream);

/* ------------------------------------------- time shards over peer memory (8e)
 * Exchange steps of a recording sharded by time across GPUs (one process per GPU).  The
 * reference has no such step: its filter_proxy walks (contact, chunk) units in one process
 * (filters.py:135-141) and detect/align/extract see the whole recording.  What crosses a
 * shard boundary (extract.py:79,86-91,160-163,245,281; features.py:224-231) travels through
 * one small ARENA per rank in peer-addressable device memory: `arenas` is a HOST array of
 * `world` device pointers, arenas[r] = base of rank r's arena as seen from THIS process
 * (torch symmetric memory, cudaIpc, or plain allocations when all shards sit on one GPU).
 * Producers store directly into the peers' arenas (NVLink), ss_peer_barrier orders them
 * with the consumers; nothing here synchronises with the host.
 *   ss_shard_arena_bytes / _offsets : size and layout {flags, thr, from_left, from_right,
 *                                     prev_last, slots, slot_stride}; zero the arena once.
 *   ss_peer_barrier        : all ranks reached `epoch` (1, 2, 3, ... per arena set) and the
 *                            peer stores issued before it on this stream are visible;
 *                            bounded spin (timeout_s, default 10 s), traps instead of hanging.
 *   ss_shard_put_thresholds: thresholds of the rank owning the first 10 s -> every arena.
 *   ss_shard_put_halos     : first/last H filtered samples of every contact -> neighbours.
 *   ss_shard_install_halos : neighbours' samples into the margins d_sig[-H..0) and
 *                            d_sig[n..n+H) of the shard's buffer; d_flags[c] = 1 where the
 *                            shard's last sample and the next shard's first form a crossing.
 *   ss_shard_append_boundary: appends t_last to the spike list of every flagged contact
 *                            (d_spt_out needs room for n_spk_in + n_c times).
 *   ss_shard_put_last      : last aligned time per contact (NaN: none) -> right neighbour,
 *                            read there as d_prev_last of ss_remove_doubles_mark.
 *   ss_shard_gram_put      : Gram/sums of ss_pca_gram re-expressed for the zero reference
 *                            -> slot `rank` of every arena;
 *   ss_shard_gram_reduce   : sum of the slots in rank order (bit-identical on every rank),
 *                            per-contact counts of every rank in d_all_counts (world, n_c).
 *                            A status word travels with every partial: SS_SHARD_OVERFLOW (1)
 *                            when *d_status (may be NULL) is non-zero - a capacity-sized list
 *                            overflowed - and SS_SHARD_HALO_MISS (2) when *d_halo_miss (may
 *                            be NULL; the counter of ss_align / ss_extract) is non-zero - a
 *                            window left the halo.  *d_status_max is the maximum over the
 *                            ranks: how all ranks learn that the step must be repeated
 *                            (exact sizes / a wider halo).
 * Kernels that take (d_spt, n_spk, d_offsets) ignore entries at or beyond
 * d_offsets[n_rows], so a list with spare capacity can be passed without a host sync.
 */
#define SS_MAX_PEERS 16
#define SS_SHARD_OVERFLOW 1
#define SS_SHARD_HALO_MISS 2
size_t ss_shard_arena_bytes(int n_c, int H, int n_win, int world);
int ss_shard_arena_offsets(int n_c, int H, int n_win, int world, size_t *out7);
int ss_peer_barrier(void *const *arenas, int rank, int world, uint64_t epoch, double timeout_s,
                    void *stream);
int ss_shard_put_thresholds(const double *d_thr, int n_c, int H, int n_win, void *const *arenas,
                            int world, void *stream);
int ss_shard_put_halos(const double *d_sig, int64_t ld, int n_c, int64_t n, int H, int n_win,
                       void *const *arenas, int rank, int world, void *stream);
int ss_shard_install_halos(double *d_sig, int64_t ld, int n_c, int64_t n, int H, int n_win,
                           const void *arena, int rank, int world, const double *d_thr,
                           int falling, int32_t *d_flags, void *stream);
int ss_shard_append_boundary(const double *d_spt_in, const int64_t *d_offsets_in, int n_c,
                             int64_t n_spk_in, const int32_t *d_flags, double t_last,
                             double *d_spt_out, int64_t *d_offsets_out, void *stream);
int ss_shard_put_last(const double *d_spt, const int64_t *d_offsets, int n_c, int H, int n_win,
                      void *const *arenas, int rank, int world, void *stream);
int ss_shard_gram_put(const double *d_gram, const double *d_sum, const double *d_ref,
                      const int64_t *d_offsets, int n_c, int H, int n_win, void *const *arenas,
                      int rank, int world, const int64_t *d_status, const int32_t *d_halo_miss,
                      void *stream);
int ss_shard_gram_reduce(const void *arena, int n_c, int H, int n_win, int world, double *d_gram,
                         double *d_sum, int64_t *d_counts, int64_t *d_all_counts,
                         int64_t *d_status_max, void *stream);
/* Results of one time shard into the whole recording's host layout: lists contact-major
 * across shards (position d_dst_row_base[c] + rank inside the shard's row), waveforms as one
 * contiguous (n_win, d_n_tot[c], 1) block per contact starting at d_block_base[c]*n_win.
 * Any input/output pair may be NULL; wave_stride = row stride of d_waves in elements. */
int ss_merge_shard(const double *d_spt, const uint8_t *d_valid, const double *d_scores,
                   int ncomps, const float *d_waves, int n_win, int64_t wave_stride, int64_t n_src,
                   const int64_t *d_src_offsets, int n_c, const int64_t *d_dst_row_base,
                   const int64_t *d_block_base, const int64_t *d_n_tot, double *d_spt_out,
                   uint8_t *d_valid_out, double *d_scores_out, float *d_waves_out, void *stream);
/* reference waveform for the Gram shift of ss_pca_gram: the first spike of every group */
int ss_pca_first_wave(const void *d_waves, int dtype, int n_win, int64_t n_spk, int n_c,
                      const int64_t *d_offsets, int64_t n_groups, double *d_ref, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SPIKESORT_B200_H */
