"""Generates the golden vectors in tests/golden/ by running the UNMODIFIED reference
files (``/root/reference/src/spike_sort/core/{filters,extract,features}.py``) through
``oracle/ref_shim.py``.  Runs only where the reference tree exists (the build
container); the resulting .npz files are committed and are what the CPU and GPU
tests compare against on any machine.

    python tests/golden/make_golden.py

Inputs are regenerated from the seeds stored in each file, and stored too where
they are small, so the fixtures are self-contained.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim                                   # noqa: E402
from spikesort_b200 import synth                              # noqa: E402


def sine_fixture():
    """The analytic signal of reference tests/test_core.py:45-53."""
    n_spikes, FS = 100, 25E3
    period = 1000.0 / FS * 100
    time = np.arange(0, period * n_spikes, 1000.0 / FS)
    spikes = np.sin(2 * np.pi / period * time)[np.newaxis, :]
    return {"data": spikes, "n_contacts": 1, "FS": FS}, period, time


def main():
    rf, re_, rfe = ref_shim.filters(), ref_shim.extract(), ref_shim.features()
    out = {}

    # --- doctest known answers, docs/source/datastructures.rst:98-109,163-199
    raw = {'data': np.array([[0, 1, 0, 0, 0, 1]]), 'FS': 10000, 'n_contacts': 1}
    spt = re_.detect_spikes(raw, thresh=0.8)
    out['doc_detect_spt'] = spt['data']
    raw2 = {'data': np.array([[0, 1, 1, 0, 0, 0, 1, -1, 0, 0, 0]]), 'FS': 10000, 'n_contacts': 1}
    w = re_.extract_spikes(raw2, {'data': np.array([0.15, 0.65, 1])}, [0, 0.4])
    out['doc_extract_in'] = raw2['data']
    out['doc_extract_waves'] = w['data']
    out['doc_extract_time'] = w['time']
    out['doc_extract_valid'] = w.get('is_valid', np.ones(3, bool))

    # --- sine signal: detect / align / extract, tests/test_core.py:55-198
    sp, period, time = sine_fixture()
    d = re_.detect_spikes(sp, thresh=0.5)
    out['sine_detect'] = d['data']
    d_f = re_.detect_spikes(sp, thresh=-0.5, edge='falling')
    out['sine_detect_falling'] = d_f['data']
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        sp_win = [-period / 6., period / 3.]
        a = re_.align_spikes(sp, d, sp_win, type='max')
        out['sine_align_max'] = a['data']
        a2 = re_.align_spikes(sp, d, [-period / 24, period / 24], type='max')   # short window
        out['sine_align_short'] = a2['data']
        dbl = {'data': np.sort(np.concatenate((d['data'], d['data'] + 0.2)))}
        a3 = re_.align_spikes(sp, dbl, sp_win, type='max', remove=True)
        out['sine_align_doubles_in'] = dbl['data']
        out['sine_align_doubles'] = a3['data']
        a4 = re_.align_spikes(sp, d, sp_win, type='min')
        out['sine_align_min'] = a4['data']
    out['sine_sp_win'] = np.array(sp_win)
    ws = re_.extract_spikes(sp, a, sp_win)
    out['sine_extract'] = ws['data']
    out['sine_extract_time'] = ws['time']
    edge_spt = {'data': np.array([0.02, 1.0, time[-1] - 0.01, time[-1] + 5.0])}
    we = re_.extract_spikes(sp, edge_spt, [-0.5, 1.0])
    out['sine_edge_spt'] = edge_spt['data']
    out['sine_edge_waves'] = we['data']
    out['sine_edge_valid'] = we['is_valid']
    rd_in = np.array([1.0, 1.02, 1.05, 3.0, 3.0799999, 3.08, 3.0800001, 2.0, 9.0])
    out['rd_in'] = rd_in
    out['rd_out'] = re_.remove_doubles({'data': rd_in}, 1000.0 / 25000.0)['data']

    # --- ellip high-pass + detect on the shifted sine: exactly 100 crossings (test_core.py:33-43)
    sp2, period, _ = sine_fixture()
    sp2['data'] = sp2['data'] + 2
    sp_freq = 1000.0 / period
    filt = rf.fltLinearIIR(sp2, sp_freq * 0.5, sp_freq * 0.4, 1, 10, 'ellip')
    out['sine_hp_filtered'] = np.asarray(filt['data'])
    out['sine_hp_detect'] = re_.detect_spikes(
        {'data': np.asarray(filt['data']), 'FS': sp2['FS'], 'n_contacts': 1}, thresh=0.5)['data']

    # --- seeded synthetic tetrode chain (int16, 25 kHz), multi-chunk filter
    FS = 25000
    x = synth.make_recording(2, 60000, FS, seed=1221, rate_hz=40.0, noise=12.0,
                             amp=(120.0, 300.0), negative=False, drift_amp=60.0).numpy()
    out['chain_seed'] = np.array([1221])
    out['chain_raw'] = x
    raw = {'data': x, 'FS': FS, 'n_contacts': 2}
    f = rf.filter_proxy(raw, rf.Filter(800., 100.), chunksize=25000)   # 3 chunks, last short
    fd = {'data': np.asarray(f['data']), 'FS': FS, 'n_contacts': 2}
    out['chain_filtered'] = fd['data']
    det = re_.detect_spikes(fd, contact=1, thresh='auto')
    out['chain_thresh'] = np.array([det['thresh']])
    out['chain_detect'] = det['data']
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        al = re_.align_spikes(fd, det, [-0.2, 0.8], type='max', contact=1)
    out['chain_align'] = al['data']
    wv = re_.extract_spikes(fd, al, [-0.2, 0.8])
    out['chain_waves'] = wv['data']
    pc = rfe.fetPCA(wv, ncomps=2)
    out['chain_pca'] = pc['data']
    evals, evecs, score = rfe.PCA(wv['data'][:, :, 0], 3)
    out['chain_pca0_evals'] = evals
    out['chain_pca0_evecs'] = evecs
    out['chain_pca0_score'] = score

    # --- cheap features and combine/normalize on the same waveforms (features.py:98-197,445-489)
    p2p = rfe.fetP2P(wv)
    out['chain_p2p'] = p2p['data']
    out['chain_min'] = rfe.fetMin(wv)['data']
    comb = rfe.combine((rfe.fetP2P(wv), rfe.fetPCA(wv, ncomps=2)), norm=True,
                       feat_method_names=['P2P', 'PCA'])
    out['chain_combined'] = comb['data']
    out['chain_combined_names'] = np.array(comb['names'])

    # --- spline upsampling (extract.py:179-205,245-258): the tutorial's resample=10
    # alignment, the deprecated extract(resample=) branch and resample_spikes itself
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        al_rs = re_.align_spikes(fd, det, [-0.2, 0.8], type='max', contact=1, resample=10)
        out['chain_align_rs10'] = al_rs['data']
        al_rs_min = re_.align_spikes(fd, det, [-0.2, 0.8], type='min', contact=0, resample=4,
                                     remove=False)
        out['chain_align_rs4_min'] = al_rs_min['data']
        wv_rs = re_.extract_spikes(fd, al_rs, [-0.2, 0.8], resample=3)
    out['chain_waves_rs3'] = wv_rs['data']
    out['chain_waves_rs3_time'] = wv_rs['time']
    out['chain_waves_rs3_keys'] = np.array(sorted(wv_rs.keys()))
    sp3, period, _ = sine_fixture()                      # tests/test_core.py:138-146
    zc = period * np.arange(100) + 1000.0 / sp3['FS'] / 2.0
    sw = re_.extract_spikes(sp3, {'data': zc}, [0, period])
    sr = re_.resample_spikes(sw, sp3['FS'] * 2)
    out['sine_resample_spt'] = zc
    out['sine_resample'] = sr['data']
    out['sine_resample_time'] = sr['time']

    # --- FilterFir (remez, filters.py:41-74) and ZeroPhaseFilter (filters.py:10-38) through
    # filter_proxy, int16 input, three chunks with a short last one
    xf = synth.make_recording(2, 14000, FS, seed=99, rate_hz=40.0, noise=12.0,
                              amp=(120.0, 300.0), negative=True, drift_amp=60.0).numpy()
    rawf = {'data': xf, 'FS': FS, 'n_contacts': 2}
    out['fir_raw'] = xf
    out['fir_hp_filtered'] = np.asarray(rf.filter_proxy(rawf, rf.FilterFir(300., 100., 101),
                                                        chunksize=6000)['data'])
    out['fir_lp_filtered'] = np.asarray(rf.filter_proxy(rawf, rf.FilterFir(3000., 5000., 64),
                                                        chunksize=6000)['data'])
    out['zpf_filtered'] = np.asarray(rf.filter_proxy(rawf, rf.ZeroPhaseFilter('ellip', [300., 3000.]),
                                                     chunksize=6000)['data'])

    # --- band-pass (order 8) on float32 input, single chunk
    x32 = synth.make_recording(1, 20000, 30000, seed=7, rate_hz=30.0, noise=8.0,
                               dtype="float32").numpy()
    raw = {'data': x32, 'FS': 30000, 'n_contacts': 1}
    fbp = rf.filter_proxy(raw, rf.Filter(np.array([300., 6000.]), np.array([100., 8000.])))
    out['bp_raw'] = x32
    out['bp_filtered'] = np.asarray(fbp['data'])

    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%.1f KiB, %d arrays)" % (path, os.path.getsize(path) / 1024.0, len(out)))


if __name__ == "__main__":
    main()
