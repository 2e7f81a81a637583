"""Split-K tile / split choices of the decode GEMMs (B200T5_SK is read when a handle is created: one model load per
configuration). Prints the best decode-loop time of three forced-length generate calls per configuration.
    python tools/sweep_sk.py "64,2,64,4,128,2,64,4;64,2,64,8,128,2,64,8" [--lengths full]"""
import gc
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from anyscale_workshop_nyc_2023_b200.modeling import B200T5ForConditionalGeneration  # noqa: E402
from anyscale_workshop_nyc_2023_b200.synth import SPECS, synthetic_token_batch  # noqa: E402
from anyscale_workshop_nyc_2023_b200.workload import checkpoint_dir  # noqa: E402

lengths = sys.argv[sys.argv.index("--lengths") + 1] if "--lengths" in sys.argv else "full"
ids, mask = synthetic_token_batch(256, 512, SPECS["flan-t5-base"].vocab_size, seed=1, lengths=lengths)
ids, mask = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
base = None
for cfg in sys.argv[1].split(";"):
    os.environ["B200T5_SK"] = cfg
    model = B200T5ForConditionalGeneration.from_pretrained(checkpoint_dir("flan-t5-base", 0))
    dec = []
    for _ in range(4):
        out = model.generate(input_ids=ids, attention_mask=mask, max_new_tokens=128, min_new_tokens=128)
        dec.append(model.stats()["decode_ms"])
    toks = out.cpu()
    if base is None:
        base = toks
    print(json.dumps({"sk": cfg, "decode_ms": min(dec[1:]), "all": [round(x, 2) for x in dec[1:]], "tokens_equal_first": bool(torch.equal(toks, base))}), flush=True)
    del model
    gc.collect()
    torch.cuda.empty_cache()
