#!/usr/bin/env python
"""bench.py - channel-samples/s of the SpikeSort front end on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): channel-samples/sec of filter -> detect -> extract -> PCA.
Workload at one GPU: configs[1], a 32-channel 30 kHz 10-minute synthetic int16
recording (32 x 18 000 000).  A step = one pass of the whole chain
(fltLinearIIR(800, 100) -> 'auto' threshold -> detect -> align -> remove doubles ->
extract -> per-contact PCA, all 32 contacts) over the recording.

  value      : chain throughput with the recording resident in HBM (CUDA events)
  e2e        : the same chain through the spike_sort.core-compatible public API with
               HOST buffers (pinned raw recording in, NumPy results out)
  roofline   : the fused filter + auto-threshold + detect stage (ss_filtfilt_detect_sos +
               ss_detect_emit, ~80 % of the step) against the measured HBM peak
  parity     : end-to-end spike-time mismatches / PCA error against the CPU oracle chain
  cpu_baseline: the oracle port (CPU restatement of the reference, oracle/) timed on a
               bounded sample of the same workload, on rank 0

With N > 1 every rank owns one recording-sized time shard (weak scaling): thresholds come
from the rank owning the first 10 s, window halos and last spike times from the neighbours,
per-contact Gram partials from everyone - stored by the library's own kernels straight into
the peers' arenas over NVLink (spikesort_b200/dist.py, csrc/exchange.cu; SS_TRANSPORT=nccl
selects torch.distributed collectives instead).

--impl reference times the CPU oracle port itself (all host cores, process pool over
contacts) on bounded samples of the same workload: the reference is pure Python and
cannot travel, the port restates it (oracle/port.py).
"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np                                             # noqa: E402

METRIC = "channel-samples/sec filter->detect->extract->PCA"
UNIT = "channel-samples/s"
FS = 30000
N_CONTACTS = 32
N_SAMPLES = 18000000
SP_WIN = [-0.2, 0.8]
CHUNK = 1000000
SEED = 2
RATE_HZ = 10.0
SM_COUNT = 148                      # B200
WORKLOAD = "c2"
TRAFFIC_SCAN_C2 = 7.154e9            # two-sweep scan (default): x read twice + float64 written once
TRAFFIC_SPAN_C2 = 6.18e9            # SS_FILT_SPAN=1: x read once
FILTER_ARGS = (800., 100.)          # fltLinearIIR(800, 100): order-1 Butterworth high-pass
CONFIG = {
    "workload": "configs[1]: 32-channel 30 kHz 10-minute synthetic int16 recording "
                "(32 x 18,000,000), fltLinearIIR(800,100) -> auto threshold -> detect(falling) "
                "-> align(min) -> extract -> fetPCA(ncomps=2), every contact",
    "n_contacts": N_CONTACTS, "n_samples": N_SAMPLES, "FS": FS, "sp_win": SP_WIN,
    "chunksize": CHUNK, "spike_rate_hz": RATE_HZ, "seed": SEED,
    "l2": "inputs larger than L2 (1.15 GB in, 4.6 GB filtered per step)",
}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ----------------------------------------------------------------- CPU arms
def _cpu_chain_one(args):
    """One contact of the oracle chain (runs in a worker process)."""
    import warnings
    from oracle import port
    x, c = args
    raw = {'data': x, 'FS': FS, 'n_contacts': 1}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, res = port.chain_per_contact(raw, port.Filter(*FILTER_ARGS), SP_WIN, edge='falling',
                                        align_type='min', ncomps=2, chunksize=CHUNK)
    return c, len(res[0][1]['data']), res[0][1]['data'], res[0][3]['data'] if res[0][3] else None


def cpu_chain(x, pool=None, keep=None):
    """Oracle chain over all contacts of host array x; returns (seconds, n_spikes).
    keep: dict filled with {contact: (aligned spike times, PCA scores)} of the oracle."""
    t0 = time.perf_counter()
    jobs = [(x[c:c + 1], c) for c in range(x.shape[0])]
    if pool is None:
        out = [_cpu_chain_one(j) for j in jobs]
    else:
        out = list(pool.map(_cpu_chain_one, jobs))
    dt = time.perf_counter() - t0
    if keep is not None:
        for c, _, spt, sc in out:
            keep[c] = (spt, sc)
    return dt, sum(o[1] for o in out)


def _shim_chain_one(args):
    """One contact through the UNMODIFIED reference files (oracle/ref_shim.py loads them from
    /root/reference with the py3 shims): filter_proxy -> detect -> align -> extract -> fetPCA."""
    import warnings
    from oracle import ref_shim
    x, c = args
    flt, ext, fet = ref_shim.filters(), ref_shim.load("extract"), ref_shim.load("features")
    raw = {'data': x, 'FS': FS, 'n_contacts': 1}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        f = flt.filter_proxy(raw, flt.Filter(*FILTER_ARGS), CHUNK)
        spt = ext.detect_spikes(f, thresh='auto', edge='falling', contact=0)
        spa = ext.align_spikes(f, spt, SP_WIN, type='min', contact=0)
        wv = ext.extract_spikes(f, spa, SP_WIN, contacts=0)
        ft = fet.fetPCA(wv, ncomps=2) if len(spa['data']) > 1 else None
    return c, len(spa['data']), spa['data'], ft['data'] if ft else None


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.
    kind "port": oracle/port.py (C filtfilt + NumPy), a process pool over contacts, every step
    the WHOLE configured recording (the config printed is the config run).  kind "shim"
    (--ref-kind shim, only where /root/reference exists, i.e. not on the GPU box): the
    unmodified reference files themselves on a stated, bounded part of the recording."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import port
    from spikesort_b200 import synth
    port.build()
    cores = os.cpu_count() or 1
    shim = args.ref_kind == "shim"
    if shim and not os.path.isdir("/root/reference"):
        print(json.dumps({"impl": "reference", "unavailable": "kind=shim needs /root/reference "
                          "(absent on the GPU box); the default kind=port runs everywhere"}))
        return
    n_sample = 3000000 if shim else N_SAMPLES
    x = synth.make_recording(N_CONTACTS, N_SAMPLES, FS, seed=SEED, rate_hz=RATE_HZ,
                             noise=10.0).numpy()[:, :n_sample]
    worker = _shim_chain_one if shim else _cpu_chain_one
    ctx = mp.get_context("fork")
    n_proc = min(cores, N_CONTACTS)
    with ctx.Pool(n_proc) as pool:
        for _ in range(max(args.warmup, 1)):
            list(pool.map(worker, [(x[c:c + 1, :300000], c) for c in range(N_CONTACTS)]))
        times = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            list(pool.map(worker, [(x[c:c + 1], c) for c in range(N_CONTACTS)]))
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = N_CONTACTS * n_sample * args.steps / total
    if shim:
        sample = ("UNMODIFIED reference files (oracle/ref_shim.py): %d contacts x the first %d of the "
                  "%d samples per step, process pool over contacts" % (N_CONTACTS, n_sample, N_SAMPLES))
    else:
        sample = ("all %d samples of all %d contacts per step (the whole configured recording), "
                  "oracle/port.py chain, process pool over contacts" % (n_sample, N_CONTACTS))
    config = dict(CONFIG) if not shim else dict(CONFIG, n_samples=n_sample,
                                                note="kind=shim runs a part of the recording")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_proc,
                         "kind": "shim" if shim else "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ GPU arm
class ClockSampler(object):
    """Samples SM clock and clock-event reasons of one GPU DURING the timed region
    (NVML from a background thread, ~2 ms period)."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20,
               "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        import threading
        self.samples, self.bits, self.max_mhz = [], 0, None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
            return
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        if self._thread is None:
            return out
        self._stop.set()
        self._thread.join(timeout=2)
        if self.samples:
            out["sm_mhz"] = statistics.median(self.samples)
            out["samples"] = len(self.samples)
        out["reasons"] = sorted(k for k, b in self.REASONS.items() if self.bits & b)
        return out


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from spikesort_b200 import _lib, _ops, filters, pipeline, synth, extract, features
    from spikesort_b200 import dist as ssdist

    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    local = env_int("LOCAL_RANK", 0)
    _lib.load()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    flt = filters.Filter(*FILTER_ARGS)
    # shard r of a 10*world-minute recording: same statistics, seed + rank
    x = synth.make_recording(N_CONTACTS, N_SAMPLES, FS, seed=SEED + rank, device=dev,
                             rate_hz=RATE_HZ, noise=10.0)
    torch.cuda.synchronize()

    last = {"res": None}

    def step():
        # lazy: the chain of this step is queued before the previous step's counts are read
        # (finalize(): the one host sync per step), so the GPU does not idle while the host gets
        # from that synchronisation to the next step's first launch - single GPU and sharded alike
        # Consecutive (independent) steps alternate between two streams (SS_BENCH_STREAMS=1: one),
        # so that the latency-bound tail of one step (align ... PCA, ~0.3 ms of small kernels)
        # runs under the next step's filter; sharded steps also alternate between two arena sets
        slot = nstep["i"] % len(streams) if streams else 0
        st = streams[slot] if streams else None
        nstep["i"] += 1
        if world > 1:
            if st is not None:
                with torch.cuda.stream(st):
                    r = ssdist.run_chain_sharded(x, FS, flt, SP_WIN, edge='falling', align_type='min',
                                                 ncomps=2, chunksize=CHUNK, lazy=True, slot=slot)
                r._bench_stream = st
            else:
                r = ssdist.run_chain_sharded(x, FS, flt, SP_WIN, edge='falling', align_type='min',
                                             ncomps=2, chunksize=CHUNK, lazy=True)
        else:
            if st is not None:
                with torch.cuda.stream(st):
                    r = pipeline.run_chain(x, FS, filt=flt, sp_win=SP_WIN, edge='falling',
                                           align_type='min', ncomps=2, chunksize=CHUNK, lazy=True)
                r._bench_stream = st
            else:
                r = pipeline.run_chain(x, FS, filt=flt, sp_win=SP_WIN, edge='falling',
                                       align_type='min', ncomps=2, chunksize=CHUNK, lazy=True)
        prev, pending["res"] = pending["res"], r
        if prev is not None:
            last["res"] = finalize(prev)
        return r

    def finalize(r):
        st = getattr(r, "_bench_stream", None)
        if st is None:
            return r.finalize()
        with torch.cuda.stream(st):
            return r.finalize()

    pending = {"res": None}
    nstep = {"i": 0}
    n_streams = env_int("SS_BENCH_STREAMS", 2)
    streams = [torch.cuda.Stream() for _ in range(n_streams)] if n_streams > 1 else []

    def finish_last():
        # inside the timed region: the last step's counts are read (and its arrays trimmed) too
        if pending["res"] is not None:
            last["res"] = finalize(pending["res"])
            pending["res"] = None
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    res = None
    # every stream needs its own warm-up (the caching allocator keeps one pool per stream: the
    # first steps on a stream allocate their 5 GB of buffers with cudaMalloc)
    args.warmup = max(args.warmup, 5 * max(len(streams), 1))
    for _ in range(args.warmup):
        res = step()
    finish_last()
    barrier()

    # filter-stage events (the dominant kernels) + whole-step events, all on the
    # current stream the library launches on
    sos, padlen = flt.sections(FS)
    f_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            for _ in range(args.steps)]
    orig_filtfilt, orig_prepare = _ops.filtfilt_detect, _ops.filtfilt_prepare
    idx = {"i": 0, "open": False}

    def timed_prepare(*a, **k):
        # time-sharded path: sweep 1 is a separate call, the stage starts here
        i = idx["i"]
        if i < len(f_ev):
            f_ev[i][0].record()
            idx["open"] = True
        return orig_prepare(*a, **k)

    def timed_filtfilt(*a, **k):
        i = idx["i"]
        if i < len(f_ev) and not idx["open"]:
            f_ev[i][0].record()
        out = orig_filtfilt(*a, **k)
        if i < len(f_ev):
            f_ev[i][1].record()
            idx["i"] += 1
            idx["open"] = False
        return out

    _ops.filtfilt_prepare = timed_prepare
    _ops.filtfilt_detect = timed_filtfilt
    p_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            for _ in range(args.steps)]
    orig_gram, orig_project = _ops.pca_gram, _ops.pca_project
    pidx = {"i": 0}

    def timed_gram(*a, **k):
        if pidx["i"] < len(p_ev):
            p_ev[pidx["i"]][0].record()
        return orig_gram(*a, **k)

    def timed_project(*a, **k):
        out = orig_project(*a, **k)
        if pidx["i"] < len(p_ev):
            p_ev[pidx["i"]][1].record()
            pidx["i"] += 1
        return out

    _ops.pca_gram, _ops.pca_project = timed_gram, timed_project
    sampler = ClockSampler(local)
    launches0 = _lib.load().ss_launch_count()              # counted by the library at every launch site
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for st in streams:
        st.wait_event(ev0)
    for _ in range(args.steps):
        res = step()
    finish_last()
    ev1.record()
    barrier()
    launches = int(_lib.load().ss_launch_count() - launches0)
    clocks = sampler.stop()
    ms_pca = [a.elapsed_time(b) for a, b in p_ev[:pidx["i"]]]
    ms_total = ev0.elapsed_time(ev1)
    ms_filter = [a.elapsed_time(b) for a, b in f_ev[:idx["i"]]]
    # the same stage with nothing else on the GPU: a few more steps on ONE stream, after the timed
    # region (in the timed region the other stream's align ... PCA kernels run under the filter)
    ms_filter_alone, ms_step_alone = [], None
    if streams:
        n_iso = 10                                         # the first 4 re-warm the stream
        f_ev[:] = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                   for _ in range(n_iso)]
        idx["i"], idx["open"] = 0, False
        p_ev[:] = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                   for _ in range(n_iso)]
        pidx["i"] = 0
        saved, streams[:] = list(streams), streams[:1]     # (its allocator pool is warm)
        iso0, iso1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for k in range(n_iso):
            if k == 4:
                with torch.cuda.stream(streams[0]):
                    iso0.record()
            step()
        finish_last()
        with torch.cuda.stream(streams[0]):
            iso1.record()
        barrier()
        streams[:] = saved
        ms_filter_alone = [a.elapsed_time(b) for a, b in f_ev[:idx["i"]]][4:]
        ms_step_alone = iso0.elapsed_time(iso1) / (n_iso - 4)
        ms_pca = [a.elapsed_time(b) for a, b in p_ev[:pidx["i"]]][4:]     # (un-overlapped)
    _ops.filtfilt_detect, _ops.filtfilt_prepare = orig_filtfilt, orig_prepare
    _ops.pca_gram, _ops.pca_project = orig_gram, orig_project
    res = last["res"]
    n_spk = int(res.spt.numel())
    n_det = int(res.spt_detected.numel())

    tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total = float(tt.item())
    ms_step = ms_total / args.steps
    value = world * N_CONTACTS * N_SAMPLES / (ms_step / 1000.0)

    # ---- e2e: public API, host buffers in, NumPy out (rank-local, then max over ranks)
    if args.no_e2e:
        if rank == 0:
            fm = statistics.mean(ms_filter) if ms_filter else float("nan")
            ab = N_CONTACTS * N_SAMPLES * (2 + 8) + N_CONTACTS * min(N_SAMPLES, 10 * FS) * 8 + \
                N_CONTACTS * N_SAMPLES * 8 + n_det * 8
            print(json.dumps({"profiling_run": True, "n_gpus": world, "ms_per_step": ms_step,
                              "value": value, "workload": CONFIG["workload"],
                              "ms_filter_stage": fm,
                              "filter_stage_frac_of_hbm_peak": ab / (fm / 1000.0) / 1e9 / measured_peak()[0],
                              "ms_pca": statistics.mean(ms_pca) if ms_pca else None,
                              "spikes_detected_rank0": n_det, "spikes_kept_rank0": n_spk,
                              "clocks": clocks, "gpu_launches": launches}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    e2e_steps = max(1, min(args.steps, 3))
    host = torch.empty((N_CONTACTS, N_SAMPLES), dtype=torch.int16).pin_memory()
    host.copy_(x)
    torch.cuda.synchronize()
    raw = {'data': host, 'FS': FS, 'n_contacts': N_CONTACTS}
    import warnings

    import spikesort_b200

    def e2e_step():
        # ONE public call for the whole recording: pinned host recording in (the H2D copy of
        # all 1.15 GB is inside), NumPy results out (spike times, waveforms, PCA scores)
        if world > 1:
            # this rank's shard of the (world x 10 min) recording, exchanges over NVLink
            out = ssdist.sort_frontend_sharded(raw, flt, SP_WIN, edge='falling', align_type='min',
                                               thresh='auto', ncomps=2, chunksize=CHUNK)
        else:
            out = spikesort_b200.sort_frontend(raw, filt=flt, sp_win=SP_WIN, edge='falling',
                                               align_type='min', thresh='auto', ncomps=2,
                                               chunksize=CHUNK)
        d2h = 0
        for c in range(N_CONTACTS):
            d2h += out['spt'][c]['data'].nbytes + out['spt_aligned'][c]['data'].nbytes + \
                out['sp_waves'][c]['data'].nbytes + out['features'][c]['data'].nbytes + 8
        return d2h

    def e2e_step_per_contact():
        # the same work through the reference's per-contact functions (32 x 4 calls)
        d2h = 0
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            f = filters.fltLinearIIR(raw, *FILTER_ARGS)        # H2D of the raw recording inside
            for c in range(N_CONTACTS):
                spt = extract.detect_spikes(f, thresh='auto', edge='falling', contact=c)
                spa = extract.align_spikes(f, spt, SP_WIN, type='min', contact=c)
                wv = extract.extract_spikes(f, spa, SP_WIN, contacts=c)
                ft = features.fetPCA(wv, ncomps=2)
                d2h += spt['data'].nbytes + spa['data'].nbytes + wv['data'].nbytes + \
                    ft['data'].nbytes + 8
        return d2h

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(e2e_steps):
        d2h = e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e_value = world * N_CONTACTS * N_SAMPLES / e2e_s
    e2e_step_per_contact()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_step_per_contact()
    torch.cuda.synchronize()
    e2e_pc_s = time.perf_counter() - t0

    # ---- roofline of the filter stage
    peak, peak_src = measured_peak()
    filt_ms = statistics.mean(ms_filter) if ms_filter else float("nan")
    # filter C*N*(s_in+8) + auto-threshold min(N,10*FS)*8 per contact +
    # detect N*8 + n_spk*8 per contact (the fused call covers all three stages)
    alg_bytes = N_CONTACTS * N_SAMPLES * (2 + 8) + N_CONTACTS * min(N_SAMPLES, 10 * FS) * 8 + \
        N_CONTACTS * N_SAMPLES * 8 + n_det * 8
    # The stage is timed twice with CUDA events on the launching stream: inside the timed region
    # (where consecutive steps alternate between two streams and the OTHER step's align / extract /
    # PCA kernels run under the filter: the interval is longer than the kernels need and adds up
    # to more than the step), and - `roofline` proper - on one stream right after it, same
    # buffers, same launches, nothing else on the GPU: the number that can be compared with the
    # serialised ncu launch list (profiles/r02_launches_c2_tma.txt).
    region_ms = filt_ms
    if ms_filter_alone:
        filt_ms = statistics.mean(ms_filter_alone)
    step_alone = ms_step_alone if ms_filter_alone else ms_step
    achieved = alg_bytes / (filt_ms / 1000.0) / 1e9
    # bytes the kernels of this stage really move: dram__bytes_read.sum + dram__bytes_write.sum
    # summed over the stage's launches of one step, from the ncu launch list of this command in
    # the final state of the round (profiles/r02_launches_c2*.csv; captured for configs[1] only)
    span = os.environ.get("SS_FILT_SPAN", "0")[:1] == "1"
    traffic = {False: TRAFFIC_SCAN_C2, True: TRAFFIC_SPAN_C2}[span] if WORKLOAD == "c2" else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": traffic,
                "traffic_unit": "bytes per launch (ncu launch list, profiles/r02_launches_c2%s.csv)"
                                % ("_span" if span else "_tma"),
                "frac_moved": (traffic / (filt_ms / 1000.0) / 1e9 / peak) if traffic else None,
                "kernel": ("ss_filtfilt_detect_sos + ss_detect_emit = filter (%s, crossing marks fused) + "
                           "auto threshold + count/scan/emit: the filter, threshold and detect stages "
                           "of the chain" % ("single-pass span kernel" if span else
                                             "summary sweep, tile scan, apply sweep")),
                "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": filt_ms,
                "share_of_step": filt_ms / step_alone, "peak_source": peak_src,
                "measured": ("CUDA events around the stage's launches on ONE stream, %d steps right "
                             "after the timed region (single-stream step: single_stream_ms_per_step); "
                             "in_timed_region = the same events inside the timed region, where two "
                             "streams overlap consecutive steps" % len(ms_filter_alone))
                if ms_filter_alone else "CUDA events around the stage's launches inside the timed region",
                "in_timed_region": {"ms_per_launch": region_ms, "streams": max(len(streams), 1),
                                    "frac": alg_bytes / (region_ms / 1000.0) / 1e9 / peak,
                                    "share_of_step": region_ms / ms_step},
                "single_stream_ms_per_step": step_alone,
                "note": "frac counts SURVEY 8(d)'s algorithmic bytes (incl. the detect pass over the "
                        "filtered signal, which the fused marks never make: it may exceed the bytes "
                        "moved); frac_moved counts the DRAM bytes ncu measured for the same launches"}
    # FP64 side of the same stage: FMAs per sample of the two sweeps - per
    # second-order section and direction 5 (recurrence) + 2 (correction) + 20/16 (lane scan),
    # first-order 3 + 1 + 5/16, plus the 2*NS+1 dot products of sweep 1.  Order 1 is far from
    # this ceiling (HBM-bound); the order-8 band-pass (--workload c2-bp8) is bound by it.
    sos_b, _ = flt.sections(FS)
    n_fo = int(sum(1 for r in sos_b if r[2] == 0.0 and r[5] == 0.0))
    n_so = len(sos_b) - n_fo
    fma_per_sample = 2 * (n_so * 8.25 + n_fo * 4.3125) + (2 * (2 * n_so + n_fo) + 1)
    sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
    fp64_peak = SM_COUNT * 64 * sm_mhz * 1e6 / 1e12                 # 64 FP64 FMA lanes per SM and clock
    fp64_ach = N_CONTACTS * N_SAMPLES * fma_per_sample / (filt_ms / 1000.0) / 1e12
    roofline["fp64"] = {"fma_per_sample": fma_per_sample, "achieved": fp64_ach, "peak": fp64_peak,
                        "unit": "TFMA/s", "frac": fp64_ach / fp64_peak,
                        "peak_source": "148 SMs x 64 FP64 FMA/clk x the SM clock sampled during the run"}
    if fp64_ach / fp64_peak > achieved / peak:
        roofline["bound"] = "fp64"
        roofline["note"] = ("this filter order is bound by the FP64 pipe, not by HBM: frac/achieved/peak "
                            "above are the HBM view of the same launch, roofline.fp64 the binding one")
    # whole-chain roofline: 18 B + spikes
    n_win = int(res.waves.shape[0])
    chain_bytes = N_CONTACTS * N_SAMPLES * 18 + n_det * 8 + n_det * (n_win * 8 + 16) + \
        n_spk * (n_win * 12 + 8) + n_spk * (n_win * 8 + 16)
    chain = {"algorithmic_bytes_per_step": chain_bytes,
             "achieved_GBs": chain_bytes / (ms_step / 1000.0) / 1e9,
             "frac": chain_bytes / (ms_step / 1000.0) / 1e9 / peak}
    if ms_pca:
        # north_star's target is stated on the filter -> detect -> extract chain; with two
        # streams the PCA of one step runs under the next step's filter, so the split is taken
        # from the single-stream steps measured right after the timed region
        pca_ms = statistics.mean(ms_pca)
        pca_bytes = n_spk * (n_win * 8 + 16)
        chain["pca_ms"] = pca_ms
        chain["single_stream_ms_per_step"] = step_alone
        chain["filter_detect_extract"] = {
            "ms": step_alone - pca_ms, "algorithmic_bytes": chain_bytes - pca_bytes,
            "frac": (chain_bytes - pca_bytes) / ((step_alone - pca_ms) / 1000.0) / 1e9 / peak}

    # ---- N > 1: every rank's results of a few contacts to rank 0 (outside the timed region)
    PAR_C = 4                                             # contacts compared at N > 1
    gathered = None
    if world > 1 and not args.no_cpu_baseline:
        mine = []
        for c in range(PAR_C):
            g = res.per_contact(c)
            mine.append((g['spt'], g.get('scores')))
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(mine, gathered, dst=0)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline on a bounded sample (oracle port, 1 core as the reference runs)
    cpu = parity = None
    if not args.no_cpu_baseline:
        from oracle import port
        port.build()
        n_sample = N_SAMPLES                      # the whole recording: ~8 s on one core
        xs = x[:, :n_sample].cpu().numpy()
        oracle_out = {}
        dt, _ = cpu_chain(xs, None, keep=oracle_out)
        cpu = {"value": N_CONTACTS * n_sample / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "all %d samples (600 s) of all %d contacts of rank 0's recording, "
                         "oracle/port.py chain (C filtfilt + NumPy), 1 core as the reference "
                         "runs, %.1f s" % (n_sample, N_CONTACTS, dt)}
        # north_star: end-to-end spike-time mismatches are counted and reported - the GPU
        # chain (own filter) against the CPU oracle chain on the same recording
        mism, total, worst = 0, 0, 0.0
        if world == 1:
            for c in range(N_CONTACTS):
                g = res.per_contact(c)
                o_spt, o_sc = oracle_out[c]
                total += len(o_spt)
                if len(o_spt) == len(g['spt']) and np.array_equal(o_spt, g['spt']):
                    if o_sc is not None and 'scores' in g:
                        sgn = np.sign(np.sum(g['scores'] * o_sc, axis=0))
                        worst = max(worst, float(np.abs(g['scores'] * sgn - o_sc).max() /
                                                 np.abs(o_sc).max()))
                else:
                    mism += len(np.setxor1d(o_spt, g['spt']))
            against = "oracle/port.py chain on the whole recording, every contact"
        else:
            # the WHOLE sharded recording (all ranks' shards in time order, regenerated here from
            # their seeds) through the oracle for the first PAR_C contacts; every rank's spike
            # times and the scores in the merged PCA basis against it
            whole = np.concatenate(
                [synth.make_recording(N_CONTACTS, N_SAMPLES, FS, seed=SEED + r, device=dev,
                                      rate_hz=RATE_HZ, noise=10.0)[:PAR_C].cpu().numpy()
                 for r in range(world)], axis=1)
            o_all = {}
            cpu_chain(whole, None, keep=o_all)
            for c in range(PAR_C):
                g_spt = np.concatenate([gathered[r][c][0] for r in range(world)])
                parts = [gathered[r][c][1] for r in range(world) if gathered[r][c][1] is not None and
                         len(gathered[r][c][0])]
                g_sc = np.concatenate(parts, axis=0) if parts else None
                o_spt, o_sc = o_all[c]
                total += len(o_spt)
                if len(o_spt) == len(g_spt) and np.array_equal(o_spt, g_spt):
                    if o_sc is not None and g_sc is not None:
                        sgn = np.sign(np.sum(g_sc * o_sc, axis=0))
                        worst = max(worst, float(np.abs(g_sc * sgn - o_sc).max() / np.abs(o_sc).max()))
                else:
                    mism += len(np.setxor1d(o_spt, g_spt))
            against = ("oracle/port.py chain on the whole %d-shard recording (%d x %d samples), "
                       "contacts 0..%d, all ranks' spike times and merged-basis scores"
                       % (world, PAR_C, world * N_SAMPLES, PAR_C - 1))
        parity = {"spike_times_compared": total, "spike_time_mismatches": mism,
                  "pca_score_max_rel_err": worst, "against": against}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(CONFIG, parallelism=("time-sharded x%d, exchanges over %s" % (
                           world, "NVLink peer memory (own kernels)"
                           if os.environ.get("SS_TRANSPORT", "peer") == "peer" else "NCCL")
                           if world > 1 else "1 GPU"),
                       spikes_detected_rank0=n_det, spikes_kept_rank0=n_spk),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": int(host.numel() * 2), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "ms_per_step": e2e_s * 1000.0,
                "api": ("spikesort_b200.dist.sort_frontend_sharded(raw_shard_dict, filt, ...) on every "
                        "rank: pinned host shard in, sharded chain with NVLink exchanges, "
                        "per-contact NumPy dicts out" if world > 1 else
                        "spikesort_b200.sort_frontend(raw_dict, filt, ...): one call, pinned host "
                        "recording in, per-contact NumPy dicts (spt, sp_waves, features) out"),
                "per_contact_api": {
                    "value": N_CONTACTS * N_SAMPLES / e2e_pc_s, "ms_per_step": e2e_pc_s * 1000.0,
                    "api": "filters.fltLinearIIR + 32 x (extract.detect_spikes, align_spikes, "
                           "extract_spikes, features.fetPCA), rank-local"}},
        "gpu_launches": launches,
        "roofline": roofline, "chain_roofline": chain, "cpu_baseline": cpu,
        "parity": parity,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def select_workload(name):
    """--workload c3: one GPU's time shard of configs[2] (Neuropixels scale, 384 contacts x
    1 h at 30 kHz sharded over 8 GPUs = 384 x 13,500,000 per GPU).  Device numbers only
    (a 10.4 GB pinned recording per rank is not staged): not the driver's bench line."""
    global N_CONTACTS, N_SAMPLES, SEED, CONFIG, FILTER_ARGS, WORKLOAD
    WORKLOAD = name
    if name == "c2":
        return
    if name == "c2-bp8":
        # configs[1] with the band-pass SURVEY.md 8d names next to the order-1 high-pass:
        # Filter([300, 6000], [100, 8000]) -> Butterworth order 8 (4 biquads); extra data point
        FILTER_ARGS = (np.array([300., 6000.]), np.array([100., 8000.]))
        CONFIG = dict(CONFIG, workload=CONFIG["workload"].replace(
            "fltLinearIIR(800,100)", "Filter([300,6000],[100,8000]) (order-8 band-pass)"))
        return
    N_CONTACTS, N_SAMPLES, SEED = 384, 13500000, 3
    CONFIG = dict(CONFIG, n_contacts=N_CONTACTS, n_samples=N_SAMPLES, seed=SEED,
                  workload="configs[2] shard: 384-channel 30 kHz recording, 450 s per GPU "
                           "(1 h over 8 GPUs), int16 (384 x 13,500,000 per GPU), same chain",
                  l2="inputs larger than L2 (10.4 GB in, 41.5 GB filtered per step and GPU)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c2-bp8", "c1", "c4", "c5"])
    ap.add_argument("--ref-kind", default="port", choices=["port", "shim"],
                    help="--impl reference: the oracle port (default, runs anywhere) or the "
                         "unmodified reference files through oracle/ref_shim.py")
    args = ap.parse_args()
    if args.workload in ("c1", "c4", "c5"):
        # the configs that are not the driver's bench line: tools/bench_extra.py
        from tools import bench_extra
        bench_extra.main(args.workload, args)
        return
    select_workload(args.workload)
    if args.workload == "c3":
        args.no_e2e = args.no_cpu_baseline = True
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
