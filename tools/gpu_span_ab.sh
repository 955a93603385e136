#!/bin/bash
# span kernel on the GPU box: parity, bench, ncu full capture of the span kernel; TC Gram tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x -k "not tc_gram and not tc_path" > gpurun_out/pytest_all.log 2>&1; tail -3 gpurun_out/pytest_all.log
timeout 300 python -m pytest tests/test_gpu_pca_tc.py -m gpu -q --timeout 200 -s > gpurun_out/pytest_tc.log 2>&1; tail -15 gpurun_out/pytest_tc.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_span.json 2> gpurun_out/bench_span.err || tail -5 gpurun_out/bench_span.err
python - <<'PY'
import json
for f in ("bench_span",):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.4g  ms/step %.3f  filter ms %.3f  frac %.3f  chain frac %.3f  e2e %.4g (%.1f ms)  launches %d" % (
            d["value"], d["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"],
            d["chain_roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"]))
    except Exception as e:
        print(f, "parse failed", e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filt_span -s 4 -c 2 -o gpurun_out/ncu_span -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -5
