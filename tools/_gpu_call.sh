#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/profile_pool_block.py > gpurun_out/profile_pool_block.log 2>&1; tail -4 gpurun_out/profile_pool_block.log
