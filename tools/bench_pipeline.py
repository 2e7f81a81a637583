"""End to end from raw strings to generated strings through the notebook's API, the CPU stage (tokenisation) either
materialised first (AIR's separate CPU stage) or streamed under the GPU stage (rayshim BatchPredictor, default):

  python tools/bench_pipeline.py [--model flan-t5-base] [--n 4096] [--batch 256|4096] [--new 128]

instruction/input strings -> BatchMapper(preprocess_function) -> BatchPredictor.predict -> DataFrame[generated_output].
Natural EOS. Prints one JSON line (wall clock, second of two runs per variant)."""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from anyscale_workshop_nyc_2023_b200 import rayshim  # noqa: E402
from anyscale_workshop_nyc_2023_b200.preprocess import make_preprocess_function  # noqa: E402
from anyscale_workshop_nyc_2023_b200.synth import synthetic_alpaca_rows  # noqa: E402
from anyscale_workshop_nyc_2023_b200.workload import checkpoint_dir, make_batch_predictor  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="flan-t5-base")
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--new", type=int, default=128)
    ap.add_argument("--reference-tokenisation", action="store_true", help="per-row padding as JOB/utils.py does (lean=False)")
    a = ap.parse_args()
    ckpt = checkpoint_dir(a.model, 0)
    ds = rayshim.data.from_huggingface(synthetic_alpaca_rows(a.n))
    prep = rayshim.data.BatchMapper(make_preprocess_function(str(ckpt), lean=not a.reference_tokenisation), batch_format="pandas")
    bp = make_batch_predictor(ckpt, preprocessor=prep, device_map="auto", torch_dtype=torch.bfloat16)
    res = {}
    outs = {}
    for name, streamed in (("materialised", False), ("streamed", True)):
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = bp.predict(ds, batch_size=a.batch, num_gpus_per_worker=1, max_scoring_workers=1,
                             pipeline_cpu_stage=streamed, max_new_tokens=a.new).to_pandas()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        res[name] = dt
        outs[name] = out["generated_output"].tolist()
    t0 = time.perf_counter()
    prep.transform(ds)
    tok_s = time.perf_counter() - t0
    line = {"workload": f"{a.model}, {a.n} alpaca-style prompts (strings), batch_size={a.batch}, max_new_tokens={a.new}, natural EOS",
            "tokenisation_alone_s": round(tok_s, 4), "identical_strings": outs["materialised"] == outs["streamed"]}
    for k, v in res.items():
        line[k] = {"seconds": round(v, 4), "prompts_per_s": round(a.n / v, 1)}
    line["speedup"] = round(res["materialised"] / res["streamed"], 3)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
