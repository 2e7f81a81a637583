#!/bin/bash
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_r2_memcheck_final.log python tools/sanitize_kernels.py attention model > gpurun_out/sanitize_final.out 2>&1
tail -3 gpurun_out/sanitize_final.out; tail -4 gpurun_out/sanitizer_r2_memcheck_final.log
