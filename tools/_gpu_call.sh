#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "norm_tail or golden_generate or run_to_run or slot_pool" > gpurun_out/pytest_normtail.log 2>&1; tail -4 gpurun_out/pytest_normtail.log
timeout 400 python tools/sweep_decode.py --no-profile --configs "norm_tail=1;norm_tail=0;norm_tail=1,xattn=0;norm_tail=0,xattn=0" > gpurun_out/sweep_normtail_full.log 2>&1
timeout 400 python tools/sweep_decode.py --no-profile --lengths alpaca --configs "norm_tail=1;norm_tail=0" > gpurun_out/sweep_normtail_alpaca.log 2>&1
for f in full alpaca; do echo == $f; grep '^{"config"' gpurun_out/sweep_normtail_$f.log | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['config'], 'decode_ms', [round(x,2) for x in r['decode_ms_all']], 'launches', r['launches'], 'tokens equal first', r['tokens_equal_first_config'])"; grep -i "error" gpurun_out/sweep_normtail_$f.log | head -3; done
