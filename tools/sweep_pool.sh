for cfg in "32 8" "8 8" "64 8" "32 4" "16 4" "32 2" "32 16"; do set -- $cfg; timeout 120 python tools/bench_stream.py --n 2048 --only-pool --admit $1 --poll $2 2>/dev/null | tail -1; done
