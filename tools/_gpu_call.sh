#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "geglu or gemm" > gpurun_out/pytest_geglu.log 2>&1; tail -4 gpurun_out/pytest_geglu.log
timeout 600 python tools/sweep_decode.py --no-profile --configs "geglu_pairwise=1;geglu_pairwise=0;geglu_pairwise=1" > gpurun_out/sweep_geglu.log 2>&1
grep '^{"config"' gpurun_out/sweep_geglu.log | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['config'], 'encoder_ms', round(r['encoder_ms'],2), 'decode_ms', round(r['decode_ms'],2), 'tokens equal first', r['tokens_equal_first_config'])"
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "golden or headline" > gpurun_out/pytest_geglu_model.log 2>&1; tail -3 gpurun_out/pytest_geglu_model.log
