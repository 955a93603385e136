/*
 * TEST INFRASTRUCTURE (oracle) - plain-C restatement of the zero-phase filter
 * the reference applies to every (contact, chunk) unit:
 *
 *   reference call sites : src/spike_sort/core/filters.py:38,74,101
 *                          (signal.filtfilt(b, a, x)), looped by
 *                          filter_proxy, filters.py:135-141
 *   algorithm            : SciPy (third-party, not in /root/reference; the
 *                          reference pins only scipy>=0.9.0, requirements.txt:2;
 *                          this image has SciPy 1.18.1):
 *                          scipy.signal.filtfilt(method='pad', padtype='odd',
 *                          padlen=3*max(len(a),len(b))), lfilter_zi, and the
 *                          Direct-Form-II-transposed loop of lfilter.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * arm may link or call this file.  It is pinned against scipy.signal.filtfilt
 * in tests/test_oracle.py (bit-exact on this image's SciPy build).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction so
 * the rounding sequence is the same as SciPy's generic x86-64 build).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { ORC_I16 = 0, ORC_F32 = 1, ORC_F64 = 2 };

/* One DF2T pass, float64, in place over y[0..n) with stride +-1 handled by the
 * caller through `step`.  State z has ntaps-1 entries and is updated.
 * Operation order follows scipy's lfilter inner loop:
 *   y = z[0] + b[0]*x;  z[k] = z[k+1] + x*b[k+1] - y*a[k+1];  z[last] = x*b[n] - y*a[n] */
static void orc_df2t(const double *b, const double *a, int ntaps, double *z,
                     double *v, long n, long step)
{
    long t;
    int k, ns = ntaps - 1;
    if (ns == 0) {
        for (t = 0; t < n; ++t, v += step) *v = *v * b[0];
        return;
    }
    for (t = 0; t < n; ++t, v += step) {
        double x = *v;
        double y = z[0] + b[0] * x;
        for (k = 0; k < ns - 1; ++k)
            z[k] = z[k + 1] + x * b[k + 1] - y * a[k + 1];
        z[ns - 1] = x * b[ns] - y * a[ns];
        *v = y;
    }
}

static double orc_get(const void *x, int dtype, long i)
{
    switch (dtype) {
    case ORC_I16: return (double)((const int16_t *)x)[i];
    case ORC_F32: return (double)((const float *)x)[i];
    default:      return ((const double *)x)[i];
    }
}

/* odd extension value 2*end - inner computed IN THE INPUT DTYPE, as NumPy does
 * in scipy.signal._arraytools.odd_ext (int16 wraps, float32 rounds). */
static double orc_odd(const void *x, int dtype, long iend, long iin)
{
    switch (dtype) {
    case ORC_I16: {
        const int16_t *p = (const int16_t *)x;
        int16_t two = (int16_t)(2 * (int32_t)p[iend]);          /* wraps */
        return (double)(int16_t)((int32_t)two - (int32_t)p[iin]);
    }
    case ORC_F32: {
        const float *p = (const float *)x;
        volatile float two = 2.0f * p[iend];
        volatile float r = two - p[iin];
        return (double)r;
    }
    default: {
        const double *p = (const double *)x;
        return 2.0 * p[iend] - p[iin];
    }
    }
}

/* filtfilt of one 1-D chunk.  b, a: ntaps each (pad the shorter with zeros),
 * any a[0] != 0; zi = lfilter_zi(b, a) (ntaps-1 values, solved by the caller
 * with the same LAPACK routine SciPy uses so that results are bit-identical).  Returns 0, or -3 if n <= padlen (SciPy raises ValueError). */
int orc_filtfilt(const double *b_in, const double *a_in, const double *zi, int ntaps,
                 const void *x, int dtype, long n, double *out)
{
    long edge = 3L * ntaps, next, i;
    int k, ns = ntaps - 1;
    double *b, *a, *z, *ext;
    if (n <= edge) return -3;
    b = (double *)malloc(sizeof(double) * ntaps * 3);
    if (!b) return -1;
    a = b + ntaps; z = a + ntaps;
    for (k = 0; k < ntaps; ++k) { b[k] = b_in[k] / a_in[0]; a[k] = a_in[k] / a_in[0]; }
    next = n + 2 * edge;
    ext = (double *)malloc(sizeof(double) * next);
    if (!ext) { free(b); return -1; }
    for (i = 0; i < edge; ++i) ext[i] = orc_odd(x, dtype, 0, edge - i);
    for (i = 0; i < n; ++i) ext[edge + i] = orc_get(x, dtype, i);
    for (i = 0; i < edge; ++i) ext[edge + n + i] = orc_odd(x, dtype, n - 1, n - 2 - i);
    /* forward, zi * ext[0] */
    for (k = 0; k < ns; ++k) z[k] = zi[k] * ext[0];
    orc_df2t(b, a, ntaps, z, ext, next, 1);
    /* backward over the forward output, zi * y[-1] */
    for (k = 0; k < ns; ++k) z[k] = zi[k] * ext[next - 1];
    orc_df2t(b, a, ntaps, z, ext + next - 1, next, -1);
    memcpy(out, ext + edge, sizeof(double) * n);
    free(ext);
    free(b);
    return 0;
}

/* filter_proxy's double loop (filters.py:135-141) over a C-contiguous
 * (n_contacts, n_samples) block with row stride ld (elements). */
int orc_filter_proxy(const double *b, const double *a, const double *zi, int ntaps,
                     const void *x, int dtype, long n_contacts, long n_samples,
                     long ld, long chunksize, double *out, long ld_out)
{
    static const int esz[3] = { 2, 4, 8 };
    long c, j0;
    for (c = 0; c < n_contacts; ++c)
        for (j0 = 0; j0 < n_samples; j0 += chunksize) {
            long len = n_samples - j0 < chunksize ? n_samples - j0 : chunksize;
            const char *src = (const char *)x + (size_t)(c * ld + j0) * esz[dtype];
            int rc = orc_filtfilt(b, a, zi, ntaps, src, dtype, len, out + c * ld_out + j0);
            if (rc) return rc;
        }
    return 0;
}
