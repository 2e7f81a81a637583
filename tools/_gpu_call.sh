mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "encoder_attn or argmax" > gpurun_out/k1.log 2>&1; echo "K1 rc=$?"; tail -3 gpurun_out/k1.log | cut -c1-300
DECODE_OPTS="chains=1,xattn=1" DECODE_T=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_cross_stream_kernel -s 30 -c 2 -o gpurun_out/prof_xs_r2 python tools/decode_once.py > gpurun_out/ncu_xs.log 2>&1; echo "NCU rc=$?"; tail -3 gpurun_out/ncu_xs.log | cut -c1-300
timeout 200 python tools/sweep_decode.py --configs "chains=2;chains=1" --reps 3 > gpurun_out/sweep2.log 2>gpurun_out/sweep2.err; echo "SWEEP rc=$?"; cut -c1-330 gpurun_out/sweep2.log
rm -f gpurun_out/parity_headline.jsonl
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "PYTEST rc=$?"; tail -12 gpurun_out/pytest_gpu.log | cut -c1-400
