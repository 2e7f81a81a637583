#!/bin/bash
bash tools/final_round_check.sh > gpurun_out/final_check.log 2>&1; tail -30 gpurun_out/final_check.log
