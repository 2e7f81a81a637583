#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_model_gpu.py -m gpu -q -k "slot_pool or run_to_run or retired" --durations=8 2>&1 | tail -14
