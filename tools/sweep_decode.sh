#!/bin/bash
# Usage: tools/sweep_decode.sh "ENV=.. ENV=.." ...   -> one line per configuration (ms/step, decode ms, encoder ms)
# Runs the full-size bench (FLAN-T5-base, 256 x 512 -> 128) without the CPU leg.
for cfg in "$@"; do
  out=$(env $cfg timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],1), round(d['decode_loop']['ms'],1), round(d['encoder']['ms'],1), round(d['value']))" 2>&1 | tail -1)
  echo "$cfg => $out"
done
