#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2_final.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],1), round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step_min_median_max"], "frac", round(d["roofline"]["frac"],3))
print(d["parity"])
print(d["cpu_baseline"]["value"], d["incumbent_hf_gpu"]["value"])
PY
tail -3 gpurun_out/bench_r2_final.err
