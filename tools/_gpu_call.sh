#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/sweep_sk.py "64,2,64,4,128,2,64,4;64,2,64,8,128,2,64,8;64,4,64,4,128,2,64,4;64,2,64,4,128,4,64,4;64,2,64,4,64,2,64,4;128,2,64,4,128,2,64,4;64,2,64,2,128,2,64,4;64,2,64,4,128,2,64,8;64,2,64,4,128,2,64,4" > gpurun_out/sweep_sk_full.log 2>&1
grep '^{"sk"' gpurun_out/sweep_sk_full.log; grep -i "error\|Traceback" gpurun_out/sweep_sk_full.log | head -3
