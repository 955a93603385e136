"""Host side of the spline upsampling kernels (``csrc/resample.cu``).

``resample_spikes`` (reference ``core/extract.py:187-205``) calls, for every (spike,
contact) waveform, ``scipy.interpolate.splrep(time, wave, s=0)`` and ``splev``.  With
``s=0`` FITPACK's ``curfit`` (``fpcurf.f``) builds the interpolating cubic spline on the
not-a-knot knot vector ``[x0]*4 + x[2:-2] + [x_end]*4``:

1. every data point adds one row (the 4 non-zero B-splines at ``x[it]``, ``fpbspl``) to a
   banded observation matrix, which is rotated into upper-triangular form by Givens
   rotations (``fpgivs``); the SAME rotations are applied to the right-hand side
   (``fprota(cos, sin, yi, z(j))``);
2. back substitution (``fpback``) yields the B-spline coefficients ``c``;
3. ``splev`` evaluates ``sum_j c[l-4+j] * h[j]`` with the 4 non-zero B-splines at the
   new abscissa.

Everything that does not touch the waveform (knots, the rotation angles, the
triangular matrix, the B-spline values at the new abscissae) depends only on the two
time grids.  ``spline_plan`` computes those once on the host, with exactly FITPACK's
operation order in IEEE float64, and the kernels replay the data-dependent operations
(2 multiplies + 1 add per rotation output, the back substitution, the 4-term sums) with
unfused float64 instructions in the same order - the upsampled waveforms are therefore
BIT-IDENTICAL to SciPy's (``tests/test_design.py::test_spline_plan_bit_exact_vs_scipy``).

The third-party algorithm restated here is FITPACK (P. Dierckx) as shipped in SciPy
(reference pin ``scipy>=0.9.0``, ``requirements.txt:2``; this image: 1.18.1), routines
``fpcurf`` (interpolation branch), ``fpbspl``, ``fpgivs``, ``fprota``, ``fpback``, ``splev``.
"""
from math import sqrt

import numpy as np

K = 3            # cubic, the splrep default the reference uses
K1 = K + 1

_CACHE = {}


def _fpbspl(t, x, l):
    """The K1 non-zero B-splines of degree K at t[l-1] <= x < t[l] (l 1-based), fpbspl.f."""
    h = [0.0] * (K + 2)
    hh = [0.0] * K1
    h[0] = 1.0
    for j in range(1, K + 1):
        for i in range(j):
            hh[i] = h[i]
        h[0] = 0.0
        for i in range(1, j + 1):
            li = l + i
            lj = li - j
            if t[li - 1] == t[lj - 1]:
                h[i] = 0.0
                continue
            f = hh[i - 1] / (t[li - 1] - t[lj - 1])
            h[i - 1] = h[i - 1] + f * (t[li - 1] - x)
            h[i] = f * (x - t[lj - 1])
    return h[:K1]


class SplinePlan(object):
    """Waveform-independent part of splrep(time, ., s=0) + splev(new_time, .).

    f64 / i32 are the two flat arrays handed to the C ABI (layout in
    include/spikesort_b200.h, "spline upsampling"):
      f64 = [rot_cos (4m) | rot_sin (4m) | tri (4m) | ev_h (4 n_out)]
      i32 = [rot_j (4m, -1 = no rotation) | ev_l (n_out)]
    """

    def __init__(self, time, new_time):
        x = [float(v) for v in np.asarray(time, dtype=np.float64)]
        m = len(x)
        if m <= K:
            raise TypeError('m > k must hold')            # splrep's own check
        n = m + K1
        nk1 = m
        t = [x[0]] * K1 + x[2:m - 2] + [x[-1]] * K1      # fpcurf.f, interpolation knots
        assert len(t) == n
        a = [[0.0] * K1 for _ in range(nk1)]
        rot_j = [-1] * (K1 * m)
        rot_c = [0.0] * (K1 * m)
        rot_s = [0.0] * (K1 * m)
        l = K1
        for it in range(m):
            xi = x[it]
            while not (xi < t[l] or l == nk1):
                l += 1
            h = _fpbspl(t, xi, l)
            j = l - K1
            for i in range(1, K1 + 1):
                j += 1
                piv = h[i - 1]
                if piv == 0.0:
                    continue
                ww = a[j - 1][0]                          # fpgivs.f
                store = abs(piv)
                if store >= ww:
                    dd = store * sqrt(1.0 + (ww / piv) ** 2)
                else:
                    dd = ww * sqrt(1.0 + (piv / ww) ** 2)
                cs = ww / dd
                sn = piv / dd
                a[j - 1][0] = dd
                # the kernels keep y and z in the same storage, which needs j <= it (an
                # empty pivot row turns the rotation into an exact row move: cos 0, sin +-1)
                assert j - 1 <= it, "rotation above the diagonal"
                rot_j[K1 * it + i - 1] = j - 1
                rot_c[K1 * it + i - 1] = cs
                rot_s[K1 * it + i - 1] = sn
                i2 = 0
                for i1 in range(i + 1, K1 + 1):           # fprota on the matrix row
                    i2 += 1
                    stor1, stor2 = h[i1 - 1], a[j - 1][i2]
                    a[j - 1][i2] = cs * stor2 + sn * stor1
                    h[i1 - 1] = cs * stor1 - sn * stor2
        ev_l, ev_h = [], []
        l, l1, k2 = K1, K1 + 1, K1 + 1                    # splev.f interval search
        for arg in np.asarray(new_time, dtype=np.float64):
            arg = float(arg)
            while not (arg >= t[l - 1] or l1 == k2):
                l1 = l
                l -= 1
            while not (arg < t[l1 - 1] or l == nk1):
                l = l1
                l1 = l + 1
            ev_l.append(l - K1)
            ev_h.extend(_fpbspl(t, arg, l))
        self.m = m
        self.n_out = len(ev_l)
        self.knots = np.array(t)
        self.new_time = np.ascontiguousarray(new_time, dtype=np.float64)
        tri = [v for row in a for v in row]
        self.f64 = np.array(rot_c + rot_s + tri + ev_h, dtype=np.float64)
        self.i32 = np.array(rot_j + ev_l, dtype=np.int32)

    # -- plain-Python replay of what the kernels do (host-logic test, small cases only)
    def apply(self, y):
        m, n_out = self.m, self.n_out
        rc, rs, tri, eh = (self.f64[:4 * m], self.f64[4 * m:8 * m], self.f64[8 * m:12 * m],
                           self.f64[12 * m:])
        rj, el = self.i32[:4 * m], self.i32[4 * m:]
        z = [0.0] * m
        for it in range(m):
            yi = float(y[it])
            for r in range(4 * it, 4 * it + 4):
                j = rj[r]
                if j < 0:
                    continue
                s1, s2 = yi, z[j]
                z[j] = float(rc[r]) * s2 + float(rs[r]) * s1
                yi = float(rc[r]) * s1 - float(rs[r]) * s2
        for i in range(m - 1, -1, -1):                    # fpback.f, bandwidth 4
            store = z[i]
            for q in range(1, min(K, m - 1 - i) + 1):
                store = store - z[i + q] * float(tri[4 * i + q])
            z[i] = store / float(tri[4 * i])
        out = np.empty(n_out)
        for q in range(n_out):
            sp = 0.0
            for j in range(4):
                sp = sp + z[el[q] + j] * float(eh[4 * q + j])
            out[q] = sp
        return out


def spline_plan(time, FS_new):
    """SplinePlan for extract.py:194 `arange(time[0], time[-1], 1000.0 / FS_new)`, cached."""
    time = np.ascontiguousarray(time, dtype=np.float64)
    key = (time.tobytes(), float(FS_new))
    if key not in _CACHE:
        if len(_CACHE) > 32:
            _CACHE.clear()
        _CACHE[key] = SplinePlan(time, np.arange(time[0], time[-1], 1000.0 / FS_new))
    return _CACHE[key]
