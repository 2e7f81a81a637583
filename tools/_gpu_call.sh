#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "encoder" > gpurun_out/pytest_encattn.log 2>&1; tail -3 gpurun_out/pytest_encattn.log
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "golden" > gpurun_out/pytest_encattn_model.log 2>&1; tail -3 gpurun_out/pytest_encattn_model.log
timeout 400 python tools/sweep_decode.py --no-profile --reps 4 --configs "chains=0" > gpurun_out/sweep_encattn.log 2>&1
grep '^{"config"' gpurun_out/sweep_encattn.log | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['config'], 'encoder_ms', round(r['encoder_ms'],3), 'decode_ms', round(r['decode_ms'],2))"
for f in fp16; do timeout 300 python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
from anyscale_workshop_nyc_2023_b200.modeling import B200T5ForConditionalGeneration
from anyscale_workshop_nyc_2023_b200.synth import SPECS, synthetic_token_batch
from anyscale_workshop_nyc_2023_b200.workload import checkpoint_dir
m = B200T5ForConditionalGeneration.from_pretrained(checkpoint_dir("flan-t5-base", 0), torch_dtype=torch.float16)
ids, mask = synthetic_token_batch(256, 512, 32128, seed=1, lengths="full")
ids, mask = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
enc = []
for _ in range(4):
    m.generate(input_ids=ids, attention_mask=mask, max_new_tokens=8, min_new_tokens=8); enc.append(m.stats()["encoder_ms"])
print("fp16 encoder_ms", [round(x, 2) for x in enc])
PY
done
