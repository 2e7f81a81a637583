#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_determinism.py "xattn=1;xattn=1,chains=3;xattn=1,xattn_stages=2;xattn=1,sk_stages64=2,sk_stages128=2;xattn=0" > gpurun_out/diag_det5.log 2> gpurun_out/diag_det5.err
cut -c1-200 gpurun_out/diag_det5.log
timeout 600 python tools/sweep_decode.py --configs "chains=2,xattn=0;chains=2,xattn=1;chains=2,xattn=1,xattn_stages=6;chains=2,xattn=1,xattn_stages=4" > gpurun_out/sweep_fix_full.log 2>&1
mv gpurun_out/sweep_decode.json gpurun_out/sweep_fix_full.json
timeout 600 python tools/sweep_decode.py --lengths alpaca --configs "chains=2,xattn=0;chains=2,xattn=1" > gpurun_out/sweep_fix_alpaca.log 2>&1
mv gpurun_out/sweep_decode.json gpurun_out/sweep_fix_alpaca.json
tail -6 gpurun_out/sweep_fix_full.log; tail -3 gpurun_out/sweep_fix_alpaca.log
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -x -q -k "attn or determinism or cross or slot_pool" > gpurun_out/pytest_fix.log 2>&1; tail -5 gpurun_out/pytest_fix.log
