"""A few decode steps of the headline configuration, for `ncu` captures of the decode kernels
(`ncu --set full -k regex:attn_cross_stream_kernel -s <skip> -c 2 python tools/decode_once.py`):
FLAN-T5-base, batch 256, 512-token prompts, DECODE_T (default 16) forced steps, two calls (the second is the warm one).
DECODE_OPTS="chains=1,xattn=0" sets library options first."""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from anyscale_workshop_nyc_2023_b200.modeling import B200T5ForConditionalGeneration  # noqa: E402
from anyscale_workshop_nyc_2023_b200.synth import SPECS, synthetic_token_batch  # noqa: E402
from anyscale_workshop_nyc_2023_b200.workload import checkpoint_dir  # noqa: E402


def main():
    name = os.environ.get("DECODE_MODEL", "flan-t5-base")
    B, S, T = int(os.environ.get("DECODE_B", 256)), int(os.environ.get("DECODE_S", 512)), int(os.environ.get("DECODE_T", 16))
    model = B200T5ForConditionalGeneration.from_pretrained(checkpoint_dir(name, 0))
    for kv in filter(None, os.environ.get("DECODE_OPTS", "").split(",")):
        k, v = kv.split("=")
        model.set_option(k, int(v))
    ids, mask = synthetic_token_batch(B, S, SPECS[name].vocab_size, seed=1, lengths="full")
    ids, mask = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    for _ in range(2):
        model.generate(input_ids=ids, attention_mask=mask, max_new_tokens=T, min_new_tokens=T)
    torch.cuda.synchronize()
    print(model.stats())


if __name__ == "__main__":
    main()
