#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-rows 0 --hf-gpu-batches 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/ncu_launch.log 2>&1
wc -l gpurun_out/launches_r2.csv
python tools/summarize_launches.py gpurun_out/launches_r2.csv > gpurun_out/launches_r2.md 2>&1; head -40 gpurun_out/launches_r2.md
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_cross_stream_kernel -s 40 -c 2 -f -o gpurun_out/prof_xs_r2_final $B > gpurun_out/ncu_xs.log 2>&1
ncu -i gpurun_out/prof_xs_r2_final.ncu-rep --page raw --csv > gpurun_out/prof_xs_r2_final_raw.csv 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:encoder_attn_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_encattn_r2 $B > gpurun_out/ncu_enc.log 2>&1
ncu -i gpurun_out/prof_encattn_r2.ncu-rep --page raw --csv > gpurun_out/prof_encattn_r2_raw.csv 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/*_raw.csv | tail -6
