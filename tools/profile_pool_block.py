"""Where one 4096-row block of the product path spends its time on one GPU (strings -> tokenise -> _predict_numpy
(slot pool, natural EOS) -> strings): stage timings, sequential, no overlap.
    python tools/profile_pool_block.py [--rows 4096]"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np
import pandas as pd
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from anyscale_workshop_nyc_2023_b200 import rayshim  # noqa: E402

rayshim.install()
from anyscale_workshop_nyc_2023_b200.preprocess import make_preprocess_function  # noqa: E402
from anyscale_workshop_nyc_2023_b200.rayshim.data import BatchMapper  # noqa: E402
from anyscale_workshop_nyc_2023_b200.synth import synthetic_alpaca_rows  # noqa: E402
from anyscale_workshop_nyc_2023_b200.workload import checkpoint_dir, make_batch_predictor  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=4096)
a = ap.parse_args()
ckpt = checkpoint_dir("flan-t5-base", 0)
raw = pd.DataFrame(synthetic_alpaca_rows(a.rows))
prep = BatchMapper(make_preprocess_function(str(ckpt)), batch_format="pandas", batch_size=4096)
bp = make_batch_predictor(ckpt, device_map="auto", torch_dtype=torch.bfloat16)
ds = rayshim.data.from_numpy({"input_ids": np.ones((8, 16), dtype=np.int64), "attention_mask": np.ones((8, 16), dtype=np.int64)})
bp.predict(ds, batch_size=8, num_gpus_per_worker=1, max_new_tokens=4)
pred = bp._worker.predictor
model, tok = pred.model, pred.tokenizer


def T(f, *args, **kw):
    t = time.perf_counter()
    r = f(*args, **kw)
    torch.cuda.synchronize()
    return r, 1e3 * (time.perf_counter() - t)


for rep in range(3):
    block, t_tok = T(prep.transform_batch, raw)
    ids, mask = block["input_ids"], block["attention_mask"]
    (out, lens), t_gen = T(model.generate_stream, ids, mask, max_new_tokens=128)
    st = model.stats()
    texts, t_dec = T(tok.batch_decode, out, skip_special_tokens=True)
    df, t_df = T(pd.DataFrame, texts, columns=["generated_output"])
    _, t_all = T(pred._predict_numpy, block, max_new_tokens=128)
    print({"tokenise_ms": round(t_tok), "generate_stream_ms": round(t_gen), "lib_encoder_ms": round(st["encoder_ms"], 1), "lib_decode_ms": round(st["decode_ms"], 1),
           "steps": st["decode_steps"], "launches": st["kernel_launches"], "batch_decode_ms": round(t_dec, 1), "dataframe_ms": round(t_df, 1),
           "_predict_numpy_ms": round(t_all), "valid_tokens": int(mask.sum()), "generated_tokens": int(lens.sum())}, flush=True)
