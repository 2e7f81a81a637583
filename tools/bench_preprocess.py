"""Tokenisation stage throughput on the host: reference style (tokenizer re-created per batch,
padding="max_length" in Python, JOB/utils.py:20-31), the cached mirror, and the lean path of preprocess.py.
CPU only. Usage: python tools/bench_preprocess.py [rows]"""
import sys
import time
from pathlib import Path

import pandas as pd

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from anyscale_workshop_nyc_2023_b200 import preprocess  # noqa: E402
from anyscale_workshop_nyc_2023_b200.synth import synthetic_alpaca_rows  # noqa: E402
from anyscale_workshop_nyc_2023_b200.workload import ASSETS  # noqa: E402


def main(n):
    from transformers import T5Tokenizer

    tokdir = str(ASSETS / "tokenizer")
    df = pd.DataFrame(synthetic_alpaca_rows(n, seed=3))

    def reference_style(batch):
        tokenizer = T5Tokenizer.from_pretrained(tokdir)
        enc = tokenizer(list(batch["instruction"]), list(batch["input"]), padding="max_length", truncation=True, return_tensors="np")
        enc["labels"] = enc["input_ids"].copy()
        return dict(enc)

    arms = {"reference style (JOB/utils.py)": reference_style,
            "cached tokenizer": preprocess.make_preprocess_function(tokdir, lean=False),
            "lean (default)": preprocess.make_preprocess_function(tokdir, lean=True)}
    for name, fn in arms.items():
        fn(df.iloc[:32])
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            out = fn(df)
            best = min(best, time.perf_counter() - t0)
        print(f"{name:34s} {n / best:10.0f} prompts/s  ({best * 1e3:.0f} ms for {n} rows, ids {out['input_ids'].shape})")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 4096)
