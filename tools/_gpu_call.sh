#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/trace_decode.py r2_final > gpurun_out/trace_r2_final.out 2>&1
timeout 120 python tools/analyze_trace.py gpurun_out/trace_r2_final.json > gpurun_out/trace_r2_final_summary.txt 2>&1; head -60 gpurun_out/trace_r2_final_summary.txt
B200T5_XATTN=ldg timeout 600 python tools/trace_decode.py r2_final_ldg > gpurun_out/trace_r2_final_ldg.out 2>&1
timeout 120 python tools/analyze_trace.py gpurun_out/trace_r2_final_ldg.json > gpurun_out/trace_r2_final_ldg_summary.txt 2>&1; head -40 gpurun_out/trace_r2_final_ldg_summary.txt
