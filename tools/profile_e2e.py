"""Where the host side of one `_predict_numpy` call goes (bench.py's e2e leg): stage timings with a synchronise after each.
    python tools/profile_e2e.py [--batch 256]"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np
import pandas as pd
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from anyscale_workshop_nyc_2023_b200.synth import SPECS, synthetic_token_batch  # noqa: E402
from anyscale_workshop_nyc_2023_b200.workload import checkpoint_dir, make_batch_predictor  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
a = ap.parse_args()
spec = SPECS["flan-t5-base"]
bp = make_batch_predictor(checkpoint_dir("flan-t5-base", 0), device_map="auto", torch_dtype=torch.bfloat16)
ids, mask = synthetic_token_batch(a.batch, 512, spec.vocab_size, seed=1, lengths="full")
import anyscale_workshop_nyc_2023_b200.rayshim as rayshim  # noqa: E402

ds = rayshim.data.from_numpy({"input_ids": ids[:8], "attention_mask": mask[:8]})
bp.predict(ds, batch_size=8, num_gpus_per_worker=1, max_new_tokens=4)
pred = bp._worker.predictor
model, tok = pred.model, pred.tokenizer
kw = dict(max_new_tokens=128, min_new_tokens=128)
batch = {"input_ids": ids, "attention_mask": mask, "labels": ids}
for _ in range(2):
    pred._predict_numpy(batch, **kw)
sync = torch.cuda.synchronize
for rep in range(3):
    t = [time.perf_counter()]
    tens = {k: pred._to_device(k, v) for k, v in batch.items() if k != "labels"}
    sync(); t.append(time.perf_counter())
    out = model.generate(**tens, **kw)
    sync(); t.append(time.perf_counter())
    host = out.cpu()
    t.append(time.perf_counter())
    texts = tok.batch_decode(host, skip_special_tokens=True)
    t.append(time.perf_counter())
    df = pd.DataFrame(texts, columns=["generated_output"])
    t.append(time.perf_counter())
    st = model.stats()
    t0 = time.perf_counter(); pred._predict_numpy(batch, **kw); sync(); whole = time.perf_counter() - t0
    names = ["to_device", "generate", "d2h", "batch_decode", "dataframe"]
    print({n: round(1e3 * (t[i + 1] - t[i]), 2) for i, n in enumerate(names)}, "library enc+dec ms", round(st["encoder_ms"] + st["decode_ms"], 2),
          "whole _predict_numpy ms", round(1e3 * whole, 2), flush=True)
