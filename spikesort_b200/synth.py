"""Synthetic extracellular recordings (seeded) for tests, smoke and bench.

The reference ships no data (the tutorial file is a download,
``docs/source/datafiles.rst:17``); SURVEY.md section 8d defines surrogates:
Gaussian noise plus Poisson spike trains of a biphasic template, int16 like the
Bakerlab/HDF5 sources (``src/spike_sort/io/filters.py:72-129``).
"""
import math

import numpy as np


def spike_template(FS, negative=True):
    """Biphasic spike, ~1 ms, peak amplitude 1 (float32 numpy)."""
    n = max(int(round(1.2e-3 * FS)), 8)
    tt = np.arange(n) / FS * 1000.0                    # ms
    main = np.exp(-0.5 * ((tt - 0.3) / 0.08) ** 2)
    rebound = 0.35 * np.exp(-0.5 * ((tt - 0.65) / 0.18) ** 2)
    w = (main - rebound).astype(np.float32)
    w /= np.abs(w).max()
    return -w if negative else w


def make_recording(n_contacts, n_samples, FS, seed, device="cpu", rate_hz=10.0, noise=10.0,
                   amp=(80.0, 200.0), negative=True, dtype="int16", drift_amp=0.0,
                   burst=False, block=1 << 24):
    """Returns a torch tensor (n_contacts, n_samples) of `dtype` on `device`.

    rate_hz: mean spike rate per contact; burst=True packs spikes in bursts with
    1-3 ms inter-spike intervals and decreasing amplitude (overlapping windows)."""
    import torch
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    out_dtype = {"int16": torch.int16, "float32": torch.float32, "float64": torch.float64}[dtype]
    out = torch.empty((n_contacts, n_samples), dtype=out_dtype, device=dev)
    tpl = torch.from_numpy(spike_template(FS, negative)).to(dev)
    L = tpl.numel()
    dur_s = n_samples / float(FS)
    for c in range(n_contacts):
        row = torch.empty(n_samples, dtype=torch.float32, device=dev)
        for lo in range(0, n_samples, block):
            hi = min(lo + block, n_samples)
            row[lo:hi] = torch.randn(hi - lo, generator=g, device=dev, dtype=torch.float32) * noise
        if drift_amp:
            tt = torch.arange(n_samples, device=dev, dtype=torch.float32) / float(FS)
            row += drift_amp * torch.sin(2 * math.pi * 5.0 * tt + c)
        n_ev = int(rate_hz * dur_s)
        if n_ev > 0 and n_samples > 4 * L:
            if burst:
                n_b = max(n_ev // 5, 1)
                starts = torch.randint(L, n_samples - 16 * L, (n_b,), generator=g, device=dev)
                isi = torch.randint(int(1e-3 * FS), int(3e-3 * FS) + 1, (n_b, 5), generator=g,
                                    device=dev)
                pos = (starts[:, None] + torch.cumsum(isi, 1) - isi[:, :1]).reshape(-1)
                dec = (0.85 ** torch.arange(5, device=dev, dtype=torch.float32))[None, :]
                a = (torch.rand(n_b, 1, generator=g, device=dev) * (amp[1] - amp[0]) + amp[0])
                a = (a * dec).reshape(-1)
            else:
                pos = torch.randint(L, n_samples - 2 * L, (n_ev,), generator=g, device=dev)
                a = torch.rand(n_ev, generator=g, device=dev) * (amp[1] - amp[0]) + amp[0]
            idx = (pos[:, None] + torch.arange(L, device=dev)[None, :]).reshape(-1)
            val = (a[:, None] * tpl[None, :]).reshape(-1)
            row.index_add_(0, idx, val)
        if out_dtype == torch.int16:
            out[c] = row.round().clamp_(-32768, 32767).to(torch.int16)
        else:
            out[c] = row.to(out_dtype)
    return out


def make_waveforms(n_win, n_spk, n_c, seed, device="cpu", n_templates=5, noise=0.3):
    """(n_win, n_spk, n_c) float32 waveforms drawn from a few templates + noise
    (config C5: fetPCA-only on pre-extracted waveforms)."""
    import torch
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    tt = torch.linspace(0, 1, n_win, device=dev)
    tpls = torch.stack([torch.sin(2 * math.pi * (k + 1) * 0.5 * tt) * math.exp(-0.2 * k) *
                        torch.exp(-((tt - 0.3 - 0.05 * k) / 0.25) ** 2) for k in range(n_templates)])
    gains = torch.rand(n_templates, n_c, generator=g, device=dev) + 0.5
    out = torch.empty((n_win, n_spk, n_c), dtype=torch.float32, device=dev)
    step = 1 << 20
    for lo in range(0, n_spk, step):
        hi = min(lo + step, n_spk)
        lab = torch.randint(0, n_templates, (hi - lo,), generator=g, device=dev)
        w = tpls[lab].t()[:, :, None] * gains[lab][None, :, :] * 5.0
        w = w + torch.randn(w.shape, generator=g, device=dev) * noise
        out[:, lo:hi, :] = w
    return out
