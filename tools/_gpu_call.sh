mkdir -p gpurun_out
rm -f gpurun_out/parity_headline.jsonl
timeout 700 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "PYTEST rc=$?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 400 python bench.py --steps 5 > gpurun_out/bench_r2_b.json 2>gpurun_out/bench_r2_b.err; echo "BENCH rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_b.json')); print({k:d[k] for k in ('ms_per_step','value')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['frac_isolated_chain_rows'], d['decode_loop']['ms'], d['encoder']['ms'], d.get('parity'), d.get('incumbent_hf_gpu',{}).get('value'), d.get('cpu_baseline',{}).get('value'))"
for X in stream ldg; do B200T5_XATTN=$X timeout 200 python bench.py --steps 5 --lengths alpaca --no-cpu-baseline --hf-gpu-batches 0 --parity-rows 0 > gpurun_out/bench_alpaca_$X.json 2>gpurun_out/bench_alpaca_$X.err; python -c "
import json; d=json.load(open('gpurun_out/bench_alpaca_$X.json')); print('ALPACA $X', d['ms_per_step'], d['decode_loop']['ms'], d['encoder']['ms'])"; done
for X in stream ldg; do for O in 1 0; do B200T5_XATTN=$X B200T5_ADMIT_OVERLAP=$O timeout 200 python tools/bench_stream.py --n 4096 --lengths full > gpurun_out/stream_full_${X}_$O.json 2>gpurun_out/stream_err.log; echo "STREAM full $X overlap=$O"; cut -c1-600 gpurun_out/stream_full_${X}_$O.json; done; done
B200T5_ADMIT_OVERLAP=1 timeout 200 python tools/bench_stream.py --n 4096 --lengths alpaca > gpurun_out/stream_alpaca_1.json 2>>gpurun_out/stream_err.log; echo "STREAM alpaca overlap=1"; cut -c1-600 gpurun_out/stream_alpaca_1.json
B200T5_ADMIT_OVERLAP=0 timeout 200 python tools/bench_stream.py --n 4096 --lengths alpaca > gpurun_out/stream_alpaca_0.json 2>>gpurun_out/stream_err.log; echo "STREAM alpaca overlap=0"; cut -c1-600 gpurun_out/stream_alpaca_0.json
