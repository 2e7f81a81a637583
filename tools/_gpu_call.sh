mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py -x -q -s -k "cross_attn" > gpurun_out/k1.log 2>&1; echo "K1 rc=$?"; grep "stream vs" gpurun_out/k1.log | sort | uniq -c | sort -rn | head -20; tail -3 gpurun_out/k1.log | cut -c1-300
python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
from anyscale_workshop_nyc_2023_b200.modeling import B200T5ForConditionalGeneration
from anyscale_workshop_nyc_2023_b200.synth import SPECS, synthetic_token_batch
from anyscale_workshop_nyc_2023_b200.workload import checkpoint_dir
m = B200T5ForConditionalGeneration.from_pretrained(checkpoint_dir("flan-t5-base", 0))
ids, mask = synthetic_token_batch(256, 512, 32128, seed=1, lengths="full")
ids, mask = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
m.set_option("xattn", 1)
for pf in (0, 1):
  for st in (3, 5, 7):
    m.set_option("xattn_l2pf", pf); m.set_option("xattn_stages", st)
    m.generate(input_ids=ids, attention_mask=mask, max_new_tokens=8, min_new_tokens=8)
    for rows in (0, 128, 86, 64):
        r = m.bench_cross_attention(reps=5, rows_per_launch=rows)
        print("ISOLATED l2pf=%d stages=%d rows=%d us=%.1f frac=%.3f" % (pf, st, rows, r["ms_per_launch"] * 1e3, r["bytes_per_launch"] / r["ms_per_launch"] / 1e6 / 6572.2), flush=True)
PY
timeout 500 python tools/sweep_decode.py --configs "chains=2,xattn=1,xattn_serialize=1;chains=2,xattn=1,xattn_serialize=1,xattn_l2pf=0;chains=3,xattn=1,xattn_serialize=1;chains=4,xattn=1,xattn_serialize=1;chains=2,xattn=1,xattn_serialize=1,xattn_stages=4;chains=3,xattn=1,xattn_serialize=1,xattn_stages=4;chains=2,xattn=1,xattn_serialize=1,xattn_stages=3;chains=2,xattn=1;chains=3,xattn=1,xattn_serialize=1,xattn_stages=3;chains=3,xattn=1" --reps 3 > gpurun_out/sweep5.log 2>gpurun_out/sweep5.err; echo "SWEEP rc=$?"; cut -c1-300 gpurun_out/sweep5.log; tail -3 gpurun_out/sweep5.err
