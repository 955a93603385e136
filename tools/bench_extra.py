"""bench.py --workload c1 | c4 | c5: the configs of BASELINE.json that are not the driver's
bench line (configs[1] is), each with per-stage times, algorithmic / moved bytes against the
measured HBM peak, a parity block against the CPU oracle and the CPU baseline timed beside it.

  c1  configs[0]: tutorial tetrode, 4 x 23 512 500 int16 @ 25 kHz, the documented flow
      (examples/sorting/cluster_auto.py:24-33 after fltLinearIIR): detect on contact 3,
      align_spikes(type='max', resample=10), extract 'all', combine(fetP2P, fetPCA) - through
      the drop-in functions themselves (spikesort_b200.{filters,extract,features}).
  c4  configs[3]: 384 contacts x 60 s @ 30 kHz, >= 200 Hz/contact in bursts - compaction,
      alignment re-queues, duplicate removal, Gram.
  c5  configs[4]: fetPCA only on 10 M pre-extracted 4-contact waveforms (25 x 1e7 x 4 float32):
      tensor-core Gram + eigensolve + projection; 8.64 GB algorithmic.

One JSON line per run (rank 0, one GPU); `metric`/`unit` stay BASELINE.json's except for c5,
whose natural unit (waveform samples/s) is stated in the line.
"""
import json
import os
import statistics
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "channel-samples/sec filter->detect->extract->PCA"
UNIT = "channel-samples/s"


def _peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class _Stages(object):
    """CUDA-event stage timer on the current stream (the one the library launches on)."""

    def __init__(self, torch):
        self.t, self.ev, self.names = torch, [], []

    def mark(self, name):
        e = self.t.cuda.Event(enable_timing=True)
        e.record()
        self.ev.append(e)
        self.names.append(name)

    def ms(self):
        self.t.cuda.synchronize()
        return {self.names[i]: self.ev[i - 1].elapsed_time(self.ev[i]) for i in range(1, len(self.ev))}


def _mean_stage(dicts):
    return {k: statistics.mean(d[k] for d in dicts) for k in dicts[0]}


def _clock_sampler(index):
    import bench
    return bench.ClockSampler(index)


# --------------------------------------------------------------------------- c1
def run_c1(args):
    import torch
    from oracle import port
    import spikesort_b200 as ss
    from spikesort_b200 import _lib, extract, features, filters, synth
    _lib.load()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    FS, n_c, n = 25000, 4, 23512500
    sp_win = [-0.2, 0.8]
    x = synth.make_recording(n_c, n, FS, seed=1221, device=dev, rate_hz=18.0, noise=30.0,
                             amp=(150.0, 400.0), negative=False, drift_amp=100.0)
    torch.cuda.synchronize()
    raw = {'data': x, 'FS': FS, 'n_contacts': n_c}

    def flow(raw_dict, st=None):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            f = filters.fltLinearIIR(raw_dict, 800., 100.)
            st and st.mark("filter")
            spt = extract.detect_spikes(f, contact=3, thresh='auto')
            st and st.mark("detect")
            spa = extract.align_spikes(f, spt, sp_win, type="max", resample=10, contact=3)
            st and st.mark("align_rs10")
            wv = extract.extract_spikes(f, spa, sp_win)
            st and st.mark("extract")
            feat = features.combine((features.fetP2P(wv), features.fetPCA(wv)), norm=True)
            st and st.mark("features")
        return spt, spa, wv, feat

    for _ in range(max(args.warmup, 3)):
        out = flow(raw)
    torch.cuda.synchronize()
    sampler = _clock_sampler(0)
    launches0 = _lib.load().ss_launch_count()
    stage_ms, t0 = [], time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        st = _Stages(torch)
        st.mark("start")
        out = flow(raw, st)
        stage_ms.append(st)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / args.steps
    ms_step = e0.elapsed_time(e1) / args.steps
    launches = int(_lib.load().ss_launch_count() - launches0)
    clocks = sampler.stop()
    stages = _mean_stage([s.ms() for s in stage_ms])
    spt, spa, wv, feat = out
    n_det, n_spk, n_win = len(spt['data']), len(spa['data']), wv['data'].shape[0]

    # e2e: the same flow from a pinned HOST recording (H2D of the raw int16 inside)
    host = torch.empty((n_c, n), dtype=torch.int16).pin_memory()
    host.copy_(x)
    raw_h = {'data': host, 'FS': FS, 'n_contacts': n_c}
    flow(raw_h)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        o2 = flow(raw_h)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / 3
    d2h = o2[0]['data'].nbytes + o2[1]['data'].nbytes + o2[2]['data'].nbytes + o2[3]['data'].nbytes

    # CPU oracle on the whole recording: baseline + parity
    xs = x.cpu().numpy()
    raw_o = {'data': xs, 'FS': FS, 'n_contacts': n_c}
    port.build()
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fo = port.filter_proxy(raw_o, port.Filter(800., 100.))
        t_f = time.perf_counter()
        spt_o = port.detect_spikes(fo, contact=3, thresh='auto')
        spa_o = port.align_spikes(fo, spt_o, sp_win, type="max", resample=10, contact=3)
        wv_o = port.extract_spikes(fo, spa_o, sp_win)
        feat_o = port.combine((port.fetP2P(wv_o), port.fetPCA(wv_o)), norm=True)
    cpu_s = time.perf_counter() - t0
    same_det = len(spt_o['data']) == n_det and np.array_equal(spt_o['data'], spt['data'])
    same_al = len(spa_o['data']) == n_spk and np.array_equal(spa_o['data'], spa['data'])
    mism = 0 if same_al else len(np.setxor1d(spa_o['data'], spa['data']))
    waves_equal = same_al and np.array_equal(wv_o['data'], wv['data'])
    feat_err = None
    if same_al:
        a, b = feat['data'], feat_o['data']
        # normalised columns: a mirrored principal axis maps column c to 1 - c
        feat_err = float(max(min(np.abs(a[:, k] - b[:, k]).max(), np.abs(1.0 - a[:, k] - b[:, k]).max())
                             for k in range(a.shape[1])))
    peak, peak_src = _peak()
    alg = n_c * n * (2 + 8) + min(n, 10 * FS) * 8 + n * 8 + n_det * 8 + n_det * (n_win * 8 + 16) + \
        n_spk * n_win * n_c * 12 + n_spk * 8 + 2 * n_spk * n_win * n_c * 4 + n_spk * n_c * 3 * 8
    moved = n_c * n * (2 + 8) + min(n, 10 * FS) * 8 + n * 8          # filter in/out + threshold + detect read
    line = {
        "metric": METRIC, "value": n_c * n / (ms_step / 1000.0), "unit": UNIT, "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "configs[0]: tutorial tetrode (4 x 23,512,500 int16 @ 25 kHz, seed 1221): "
                               "fltLinearIIR(800,100) -> detect_spikes(contact=3, 'auto') -> "
                               "align_spikes(type='max', resample=10) -> extract_spikes('all') -> "
                               "combine(fetP2P, fetPCA), through the drop-in functions "
                               "(cluster_auto.py:24-33)",
                   "n_contacts": n_c, "n_samples": n, "FS": FS, "sp_win": sp_win,
                   "spikes_detected": n_det, "spikes_kept": n_spk,
                   "l2": "inputs larger than L2 (188 MB in, 752 MB filtered per step)"},
        "clocks": clocks, "stage_ms": stages, "host_wall_ms_per_step": wall * 1000.0,
        "e2e": {"value": n_c * n / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(host.numel() * 2),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1000.0, "steps": 3,
                "api": "filters.fltLinearIIR, extract.detect_spikes / align_spikes / extract_spikes, "
                       "features.combine((fetP2P, fetPCA)) on a pinned host recording"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "whole flow (the filter is ~all of the bytes)",
                     "algorithmic_bytes_per_launch": alg, "achieved": alg / (ms_step / 1e3) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": alg / (ms_step / 1e3) / 1e9 / peak,
                     "moved_bytes_estimate": moved, "frac_moved": moved / (ms_step / 1e3) / 1e9 / peak,
                     "filter_stage": {"ms": stages["filter"],
                                      "frac_moved": n_c * n * 10 / (stages["filter"] / 1e3) / 1e9 / peak},
                     "traffic": None, "peak_source": peak_src,
                     "note": "the flow is five library calls with their own host synchronisations "
                             "(sizes of the results): launch / sync latency, not bandwidth, bounds the "
                             "stages behind the filter"},
        "cpu_baseline": {"value": n_c * n / cpu_s, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": "the whole recording, oracle/port.py (C filtfilt + NumPy + SciPy "
                                   "splines), 1 core, %.1f s (filter %.1f s)" % (cpu_s, t_f - t0)},
        "parity": {"detected_identical": bool(same_det), "aligned_identical": bool(same_al),
                   "spike_times_compared": len(spa_o['data']), "spike_time_mismatches": int(mism),
                   "waveforms_identical": bool(waves_equal), "features_max_abs_err": feat_err,
                   "against": "oracle/port.py on the whole recording"},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- c4
def run_c4(args):
    import torch
    import bench
    from oracle import port
    from spikesort_b200 import _lib, _ops, filters, pipeline, synth
    _lib.load()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    FS, n_c, n, chunk = 30000, 384, 1800000, 600000
    sp_win = [-0.2, 0.8]
    x = synth.make_recording(n_c, n, FS, seed=4, device=dev, rate_hz=220.0, noise=10.0, burst=True)
    torch.cuda.synchronize()
    flt = filters.Filter(800., 100.)
    kw = dict(filt=flt, sp_win=sp_win, edge='falling', align_type='min', ncomps=2, chunksize=chunk)

    names = ["filtfilt_detect", "align", "remove_doubles", "extract", "pca_gram", "pca_eig", "pca_project"]
    originals = {k: getattr(_ops, k) for k in names}
    acc = {k: [] for k in names}

    def timed(name):
        fn = originals[name]

        def wrapper(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            acc[name].append((e0, e1))
            return out
        return wrapper

    pending = {"res": None}

    def step():
        r = pipeline.run_chain(x, FS, lazy=True, **kw)
        prev, pending["res"] = pending["res"], r
        if prev is not None:
            prev.finalize()
        return r

    for _ in range(max(args.warmup, 3)):
        res = step()
    pending["res"].finalize()
    pending["res"] = None
    torch.cuda.synchronize()
    for k in names:
        setattr(_ops, k, timed(k))
    sampler = _clock_sampler(0)
    launches0 = _lib.load().ss_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = step()
    pending["res"].finalize()
    e1.record()
    torch.cuda.synchronize()
    for k in names:
        setattr(_ops, k, originals[k])
    ms_step = e0.elapsed_time(e1) / args.steps
    launches = int(_lib.load().ss_launch_count() - launches0)
    clocks = sampler.stop()
    stages = {k: statistics.mean(a.elapsed_time(b) for a, b in v) for k, v in acc.items() if v}
    n_det, n_spk = int(res.spt_detected.numel()), int(res.spt.numel())
    n_win = int(res.waves.shape[0])
    peak, peak_src = _peak()
    alg = n_c * n * 18 + n_c * min(n, 10 * FS) * 8 + n_det * 8 + n_det * (n_win * 8 + 16) + \
        n_spk * (n_win * 12 + 8) + n_spk * (n_win * 8 + 16)
    moved = n_c * n * 10 + n_c * min(n, 10 * FS) * 8 + n_det * (n_win * 8 + 16) + \
        n_spk * (n_win * 12 + 8) + n_spk * (n_win * 8 + 16)

    # CPU oracle on a subset of the contacts: baseline + parity
    port.build()
    sub = list(range(0, n_c, 48))                                  # 8 contacts
    xs = x[sub].cpu().numpy()
    keep = {}
    bench.FS, bench.SP_WIN, bench.CHUNK, bench.FILTER_ARGS = FS, sp_win, chunk, (800., 100.)
    cpu_s, _ = bench.cpu_chain(xs, None, keep=keep)
    mism = total = 0
    worst = 0.0
    # the GPU's PCA basis per contact only depends on that contact: comparable one by one
    for k, c in enumerate(sub):
        g = res.per_contact(c)
        o_spt, o_sc = keep[k]
        total += len(o_spt)
        if len(o_spt) == len(g['spt']) and np.array_equal(o_spt, g['spt']):
            if o_sc is not None:
                sgn = np.sign(np.sum(g['scores'] * o_sc, axis=0))
                worst = max(worst, float(np.abs(g['scores'] * sgn - o_sc).max() / np.abs(o_sc).max()))
        else:
            mism += len(np.setxor1d(o_spt, g['spt']))
    line = {
        "metric": METRIC, "value": n_c * n / (ms_step / 1000.0), "unit": UNIT, "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "configs[3]: burst stress, 384 contacts x 60 s @ 30 kHz int16 "
                               "(384 x 1,800,000), >= 200 Hz/contact in bursts (ISI 1-3 ms, decreasing "
                               "amplitude), fltLinearIIR(800,100) -> auto threshold -> detect(falling) -> "
                               "align(min) -> remove doubles -> extract -> fetPCA(2), every contact",
                   "n_contacts": n_c, "n_samples": n, "FS": FS, "sp_win": sp_win, "chunksize": chunk,
                   "spike_rate_hz_per_contact_detected": n_det / n_c / (n / FS),
                   "spikes_detected": n_det, "spikes_kept": n_spk,
                   "l2": "inputs larger than L2 (1.38 GB in, 5.5 GB filtered per step)"},
        "clocks": clocks, "stage_ms": stages, "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "whole chain", "algorithmic_bytes_per_launch": alg,
                     "achieved": alg / (ms_step / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg / (ms_step / 1e3) / 1e9 / peak, "moved_bytes_estimate": moved,
                     "frac_moved": moved / (ms_step / 1e3) / 1e9 / peak, "traffic": None,
                     "peak_source": peak_src},
        "cpu_baseline": {"value": len(sub) * n / cpu_s, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": "%d of the %d contacts (every 48th), all %d samples, oracle/port.py "
                                   "chain, 1 core, %.1f s" % (len(sub), n_c, n, cpu_s)},
        "parity": {"contacts_compared": len(sub), "spike_times_compared": total,
                   "spike_time_mismatches": int(mism), "pca_score_max_rel_err": worst,
                   "against": "oracle/port.py chain on the same contacts"},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- c5
def run_c5(args):
    import torch
    from oracle import port
    from spikesort_b200 import _lib, _ops, features, synth
    _lib.load()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    n_win, n_spk, n_c, ncomps = 25, int(os.environ.get("SS_C5_SPIKES", 10000000)), 4, 2
    w = synth.make_waveforms(n_win, n_spk, n_c, seed=5, device=dev) + 1.0
    torch.cuda.synchronize()

    def device_fetpca(st=None):
        gram, ssum, _, counts = _ops.pca_gram(w)                   # 'auto': tensor cores at this size
        st and st.mark("gram")
        evals, evecs = _ops.pca_eig(gram, ssum, counts, top=ncomps)
        st and st.mark("eig")
        sc = _ops.pca_project(w, evals, evecs, ncomps)
        st and st.mark("project")
        return sc

    for _ in range(max(args.warmup, 3)):
        sc = device_fetpca()
    torch.cuda.synchronize()
    from bench import ClockSampler
    sampler = ClockSampler(0)
    launches0 = _lib.load().ss_launch_count()
    sts = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        st = _Stages(torch)
        st.mark("start")
        sc = device_fetpca(st)
        sts.append(st)
    e1.record()
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / args.steps
    launches = int(_lib.load().ss_launch_count() - launches0)
    clocks = sampler.stop()
    stages = _mean_stage([s.ms() for s in sts])
    peak, peak_src = _peak()
    in_bytes = n_win * n_spk * n_c * 4
    alg = 2 * in_bytes + n_spk * n_c * ncomps * 8
    samples = n_win * n_spk * n_c

    # float64 CUDA-core Gram on a bounded part as the accuracy reference of the tensor-core one
    n_chk = min(n_spk, 2000000)
    wc = w[:, :n_chk, :].contiguous()
    g_tc, s_tc, ref, cnt = _ops.pca_gram(wc, mode="tc")
    g_64, s_64, _, _ = _ops.pca_gram(wc, ref=ref, mode="f64")
    gram_err = float((g_tc - g_64).abs().max() / g_64.abs().max())
    del wc

    # e2e: public features.fetPCA on a pinned host block
    e2e = None
    if not args.no_e2e:
        host = torch.empty(w.shape, dtype=torch.float32).pin_memory()
        host.copy_(w)
        spd = {'data': host, 'time': np.arange(n_win) / 25.0 - 0.2, 'FS': 25000}
        features.fetPCA(spd, ncomps=ncomps)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = features.fetPCA(spd, ncomps=ncomps)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        e2e = {"value": samples / e2e_s, "unit": "waveform-samples/s", "h2d_bytes_per_step": int(in_bytes),
               "d2h_bytes_per_step": int(out['data'].nbytes), "ms_per_step": e2e_s * 1e3, "steps": 1,
               "api": "features.fetPCA(sp_waves_dict, ncomps=2) on a pinned host block"}
        del host

    # CPU baseline + parity on a bounded sample (np.cov + eig of the oracle)
    cpu = parity = None
    if not args.no_cpu_baseline:
        n_cpu = min(n_spk, 1000000)
        hw = w[:, :n_cpu, :].cpu().numpy()
        t0 = time.perf_counter()
        ref_f = port.fetPCA({'data': hw}, ncomps=ncomps)
        cpu_s = time.perf_counter() - t0
        got = features.fetPCA({'data': hw}, ncomps=ncomps)
        a, b = got['data'], ref_f['data']
        sgn = np.sign(np.sum(a * b, axis=0))
        parity = {"spikes_compared": n_cpu, "pca_score_max_rel_err": float(np.abs(a * sgn - b).max() / np.abs(b).max()),
                  "gram_tc_vs_f64_max_rel_err": gram_err,
                  "against": "oracle/port.py fetPCA (np.cov + np.linalg.eig) on the first %d spikes; the "
                             "tensor-core Gram against the float64 CUDA-core Gram on %d spikes" % (n_cpu, n_chk)}
        cpu = {"value": n_win * n_cpu * n_c / cpu_s, "unit": "waveform-samples/s", "cores": os.cpu_count(),
               "kind": "port", "sample": "first %d of the %d spikes, oracle fetPCA, NumPy/BLAS default "
                                         "threads, %.2f s" % (n_cpu, n_spk, cpu_s)}
    line = {
        "metric": "waveform-samples/sec fetPCA (configs[4]; channel-samples do not exist here)",
        "value": samples / (ms_step / 1e3), "unit": "waveform-samples/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32x2+f64",
        "data": "synthetic",
        "config": {"workload": "configs[4]: fetPCA only on %d pre-extracted 4-contact waveforms "
                               "(%d x %d x %d float32, %.2f GB): Gram (tcgen05 split-TF32) + leading "
                               "eigenpairs + projection (ncomps=2)" % (n_spk, n_win, n_spk, n_c, in_bytes / 1e9),
                   "n_win": n_win, "n_spikes": n_spk, "n_contacts": n_c, "ncomps": ncomps,
                   "l2": "inputs larger than L2 (%.2f GB read twice per step)" % (in_bytes / 1e9)},
        "clocks": clocks, "stage_ms": stages, "e2e": e2e, "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "pca_gram_tc_kernel (Gram pass)",
                     "algorithmic_bytes_per_launch": in_bytes, "ms_per_launch": stages["gram"],
                     "achieved": in_bytes / (stages["gram"] / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": in_bytes / (stages["gram"] / 1e3) / 1e9 / peak, "traffic": None,
                     "peak_source": peak_src,
                     "project": {"algorithmic_bytes": in_bytes + n_spk * n_c * ncomps * 8,
                                 "ms": stages["project"],
                                 "frac": (in_bytes + n_spk * n_c * ncomps * 8) / (stages["project"] / 1e3) / 1e9 / peak},
                     "whole_fetpca": {"algorithmic_bytes": alg, "ms": ms_step,
                                      "frac": alg / (ms_step / 1e3) / 1e9 / peak}},
        "cpu_baseline": cpu, "parity": parity,
    }
    print(json.dumps(line))


def main(workload, args):
    {"c1": run_c1, "c4": run_c4, "c5": run_c5}[workload](args)
