#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/sweep_decode.py --configs "chains=2,xattn=0;chains=2,xattn=1;chains=1,xattn=0;chains=3,xattn=1" > gpurun_out/sweep_union_full.log 2>&1
mv gpurun_out/sweep_decode.json gpurun_out/sweep_union_full.json
timeout 600 python tools/sweep_decode.py --lengths uniform --configs "chains=2,xattn=0;chains=2,xattn=1" > gpurun_out/sweep_union_uniform.log 2>&1
mv gpurun_out/sweep_decode.json gpurun_out/sweep_union_uniform.json
timeout 600 python tools/sweep_decode.py --lengths alpaca --configs "chains=2,xattn=0;chains=2,xattn=1;chains=2,xattn=2" > gpurun_out/sweep_union_alpaca.log 2>&1
mv gpurun_out/sweep_decode.json gpurun_out/sweep_union_alpaca.json
for f in full uniform alpaca; do echo == $f; python - <<PY
import json
for r in json.load(open("gpurun_out/sweep_union_$f.json")):
    print({k: (round(v,2) if isinstance(v,float) else v) for k,v in r.items() if k not in ("decode_ms_all","launches")})
PY
done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all_r2b.log 2>&1; tail -5 gpurun_out/pytest_all_r2b.log
