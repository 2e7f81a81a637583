"""One encoder pass of FLAN-T5-base at the bench shape (for ncu captures of the encoder kernels)."""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from anyscale_workshop_nyc_2023_b200.modeling import B200T5ForConditionalGeneration  # noqa: E402
from anyscale_workshop_nyc_2023_b200.synth import SPECS, synthetic_token_batch  # noqa: E402
from anyscale_workshop_nyc_2023_b200.workload import checkpoint_dir  # noqa: E402

model_name = os.environ.get("TRACE_MODEL", "flan-t5-base")
B, S = int(os.environ.get("TRACE_B", 256)), int(os.environ.get("TRACE_S", 512))
spec = SPECS[model_name]
dtype = torch.float16 if os.environ.get("TRACE_DTYPE", "bf16") == "fp16" else torch.bfloat16
model = B200T5ForConditionalGeneration.from_pretrained(checkpoint_dir(model_name, 0), torch_dtype=dtype)
ids, mask = synthetic_token_batch(B, S, spec.vocab_size, seed=1, lengths="full")
for _ in range(int(os.environ.get("TRACE_REPS", 2))):
    out = model.encode(ids, mask)
torch.cuda.synchronize()
print("encoded", tuple(out.shape))
