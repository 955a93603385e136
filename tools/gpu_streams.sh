#!/bin/bash
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 400 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.4g ms/step %.3f steps %d warmup %d" % (d["value"], d["ms_per_step"], d["steps"], d["warmup"]))
print("roofline ms %.3f frac %.3f frac_moved %.3f share %.3f single-stream step %.3f | in region %s" % (r["ms_per_launch"], r["frac"], r["frac_moved"], r["share_of_step"], r["single_stream_ms_per_step"], r["in_timed_region"]))
print("chain", d["chain_roofline"])
print("parity", d["parity"]["spike_time_mismatches"], d["parity"]["pca_score_max_rel_err"], "e2e %.1f ms per-contact %.1f ms" % (d["e2e"]["ms_per_step"], d["e2e"]["per_contact_api"]["ms_per_step"]), "launches", d["gpu_launches"], d["clocks"])
PY
