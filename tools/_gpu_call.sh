mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "cross_attn or argmax" > gpurun_out/k1.log 2>&1; echo "K1 rc=$?"; tail -4 gpurun_out/k1.log | cut -c1-400
timeout 400 python tools/sweep_decode.py --configs "chains=2;chains=1,xattn=1;chains=2,xattn=1;chains=3,xattn=1;chains=2,xattn=1,xattn_stages=4;chains=2,xattn=1,xattn_stages=7,sk_stages64=3,sk_stages128=2;chains=2,xattn=1,xattn_late_pdl=0;chains=3,xattn=1,xattn_stages=4;chains=4,xattn=1,xattn_stages=4" --reps 3 > gpurun_out/sweep3.log 2>gpurun_out/sweep3.err; echo "SWEEP rc=$?"; cut -c1-330 gpurun_out/sweep3.log; tail -3 gpurun_out/sweep3.err
python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
from anyscale_workshop_nyc_2023_b200.modeling import B200T5ForConditionalGeneration
from anyscale_workshop_nyc_2023_b200.synth import SPECS, synthetic_token_batch
from anyscale_workshop_nyc_2023_b200.workload import checkpoint_dir
m = B200T5ForConditionalGeneration.from_pretrained(checkpoint_dir("flan-t5-base", 0))
ids, mask = synthetic_token_batch(256, 512, 32128, seed=1, lengths="full")
ids, mask = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
for x in (0, 1):
    m.set_option("xattn", x)
    m.generate(input_ids=ids, attention_mask=mask, max_new_tokens=8, min_new_tokens=8)
    for rows in (0, 128, 64):
        r = m.bench_cross_attention(reps=5, rows_per_launch=rows)
        print("ISOLATED xattn=%d rows=%d us=%.1f GB/s=%.0f frac=%.3f" % (x, rows, r["ms_per_launch"] * 1e3, r["bytes_per_launch"] / r["ms_per_launch"] / 1e6, r["bytes_per_launch"] / r["ms_per_launch"] / 1e6 / 6572.2))
PY
rm -f gpurun_out/parity_headline.jsonl
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "PYTEST rc=$?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-400
