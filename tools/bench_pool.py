"""The product's multi-GPU path, measured as a user drives it (BASELINE configs[2] / configs[3]): Alpaca-style STRINGS ->
BatchMapper(preprocess_function) -> BatchPredictor.predict(ds, num_gpus_per_worker=1, batch_size=..., max_new_tokens=128)
over one scoring process per GPU (rayshim/pool.py), natural EOS, exactly the calls of
NLP_workloads/Anyscale_job/flan-t5-batch-inference.py:119-138 (predict -> to_pandas -> join).

    python tools/bench_pool.py --workers 1,8 [--n 65536] [--model flan-t5-base] [--batch-size 4096]

For every worker count: a cold predict (spawns the pool: model load + graph capture) and warm ones (the pool stays
alive between predict calls); prompts/s and generated tokens/s from wall-clock around predict(); the output rows of
every count are compared with the first count's, row for row. One JSON line; also written to gpurun_out/pool_<tag>.json.
(`generated_tokens_per_s_est` re-tokenises the output strings: a lower bound that is far off for random-weight checkpoints,
whose token ids do not round-trip through text; tools/profile_pool_block.py reports the library's own count.)"""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from anyscale_workshop_nyc_2023_b200 import rayshim  # noqa: E402

rayshim.install()

from ray.data.preprocessors import BatchMapper  # noqa: E402
from ray.train.batch_predictor import BatchPredictor  # noqa: E402
from ray.train.huggingface import HuggingFaceCheckpoint  # noqa: E402
from transformers import T5Tokenizer  # noqa: E402

from anyscale_workshop_nyc_2023_b200.modeling import B200T5ForConditionalGeneration  # noqa: E402
from anyscale_workshop_nyc_2023_b200.predictor import HuggingFaceModelPredictor  # noqa: E402
from anyscale_workshop_nyc_2023_b200.preprocess import make_preprocess_function  # noqa: E402
from anyscale_workshop_nyc_2023_b200.synth import synthetic_alpaca_rows  # noqa: E402
from anyscale_workshop_nyc_2023_b200.workload import checkpoint_dir  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="flan-t5-base")
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--workers", default="1,2", help="comma-separated scoring-worker counts (one GPU each)")
    ap.add_argument("--batch-size", type=int, default=4096)
    ap.add_argument("--max-new-tokens", type=int, default=128)
    ap.add_argument("--reps", type=int, default=2, help="warm predict() calls per worker count")
    ap.add_argument("--tag", default="base")
    ap.add_argument("--weak", action="store_true", help="--n prompts PER WORKER (the dataset grows with the pool, as bench.py's per-rank work does)")
    a = ap.parse_args()
    ckpt = checkpoint_dir(a.model, seed=0)
    counts = [int(x) for x in a.workers.split(",")]
    all_rows = synthetic_alpaca_rows(a.n * (max(counts) if a.weak else 1))
    prep = BatchMapper(make_preprocess_function(str(ckpt)), batch_format="pandas", batch_size=4096)
    tok = T5Tokenizer.from_pretrained(str(ckpt))
    res = {"model": a.model, "prompts": a.n, "per_worker": bool(a.weak), "batch_size": a.batch_size, "max_new_tokens": a.max_new_tokens,
           "gpus_visible": torch.cuda.device_count(), "runs": []}
    first = None
    for n_workers in counts:
        n = a.n * n_workers if a.weak else a.n
        ds = rayshim.data.from_huggingface({k: v[:n] for k, v in all_rows.items()})  # (sequential generator: a prefix)
        checkpoint = HuggingFaceCheckpoint.from_directory(str(ckpt))
        checkpoint.set_preprocessor(prep)
        bp = BatchPredictor.from_checkpoint(checkpoint=checkpoint, predictor_cls=HuggingFaceModelPredictor,
                                            model_cls=B200T5ForConditionalGeneration, tokenizer=T5Tokenizer, use_gpu=True,
                                            device_map="auto", torch_dtype=torch.bfloat16)
        kw = dict(num_gpus_per_worker=1, batch_size=a.batch_size, max_new_tokens=a.max_new_tokens,
                  min_scoring_workers=n_workers, max_scoring_workers=n_workers)
        times = []
        for rep in range(a.reps + 1):
            t0 = time.perf_counter()
            prediction = bp.predict(ds, **kw)
            prediction_pd = prediction.to_pandas()
            times.append(time.perf_counter() - t0)
        joined = ds.to_pandas().join(prediction_pd, how="inner")
        assert len(joined) == n
        texts = prediction_pd["generated_output"].tolist()
        if first is None:
            first = texts
        warm = min(times[1:])
        gen_tokens = sum(len(t) for t in tok(texts[: min(n, 2048)], add_special_tokens=True)["input_ids"]) * (n / min(n, 2048))
        common = min(len(first), len(texts))
        res["runs"].append({"workers": n_workers, "prompts": n, "cold_s": times[0], "warm_s": warm, "warm_all_s": times[1:],
                            "prompts_per_s": n / warm, "generated_tokens_per_s_est": gen_tokens / warm,
                            "rows_equal_first_run": texts[:common] == first[:common], "rows_compared": common})
        print(json.dumps(res["runs"][-1]), file=sys.stderr, flush=True)
        bp.shutdown()
    base = res["runs"][0]
    for r in res["runs"]:
        r["scaling_vs_first"] = r["prompts_per_s"] / base["prompts_per_s"]
        r["efficiency_vs_first"] = r["scaling_vs_first"] / (r["workers"] / base["workers"])
    print(json.dumps(res))
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / f"pool_{a.tag}.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
