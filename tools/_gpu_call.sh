#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --model flan-t5-large --steps 2 --warmup 3 --no-cpu-baseline --hf-gpu-batches 0 --parity-rows 0 > gpurun_out/bench_r2_large_n8.json 2> gpurun_out/bench_r2_large_n8.err; head -c 500 gpurun_out/bench_r2_large_n8.json; echo
timeout 900 python tools/bench_pool.py --model flan-t5-large --weak --workers 1,8 --n 4096 --reps 1 --tag r2_large_g8 > gpurun_out/pool_r2_large_g8.log 2>&1; grep '^{"workers"' gpurun_out/pool_r2_large_g8.log | cut -c1-330; tail -2 gpurun_out/pool_r2_large_g8.log | cut -c1-200
