"""CPU-only: the Ray-compatible shim and the predictor/preprocess mirrors, driven exactly as the
notebook drives them (BASELINE config 1: CPU map_batches path with the HF model)."""
import numpy as np
import pandas as pd
import pytest
import torch

from anyscale_workshop_nyc_2023_b200 import rayshim
from anyscale_workshop_nyc_2023_b200.parallel import restore_order, shard_block_indices
from anyscale_workshop_nyc_2023_b200.preprocess import make_preprocess_function
from anyscale_workshop_nyc_2023_b200.synth import SPECS, synthetic_alpaca_rows
from anyscale_workshop_nyc_2023_b200.workload import ASSETS, checkpoint_dir, make_batch_predictor


class HFOnCpu:
    """model_cls stand-in with from_pretrained(dir, **kw): the dependency's own model on CPU."""

    @staticmethod
    def from_pretrained(path, **kw):
        from oracle.hf_anchor import load_hf_model

        return load_hf_model(path, dtype=torch.float32, device="cpu")


def test_dataset_ops_preserve_order():
    ds = rayshim.data.from_huggingface(synthetic_alpaca_rows(23))
    assert ds.count() == 23 and set(ds.columns()) == {"instruction", "input", "output", "text"}
    lim = ds.limit(10)
    assert lim.count() == 10
    df = lim.to_pandas()
    assert df["instruction"].tolist() == synthetic_alpaca_rows(23)["instruction"][:10]
    doubled = lim.map_batches(lambda b: pd.DataFrame({"n": b["instruction"].str.len()}), batch_size=3)
    assert doubled.to_pandas()["n"].tolist() == [len(s) for s in df["instruction"]]
    assert len(lim.take(4)) == 4
    both = rayshim.data.from_huggingface({"train": synthetic_alpaca_rows(5), "test": synthetic_alpaca_rows(4)})
    assert both["train"].count() == 5 and both["test"].count() == 4


def test_preprocess_pair_template_and_padding():
    fn = make_preprocess_function(str(ASSETS / "tokenizer"))
    out = fn(pd.DataFrame({"instruction": ["Describe the water cycle.", "vi"], "input": ["2, 4, 8", ""]}))
    assert set(out) == {"input_ids", "attention_mask", "labels"}
    assert out["input_ids"].shape == (2, 512) and out["input_ids"].dtype == np.int64
    assert (out["labels"] == out["input_ids"]).all()
    row = out["input_ids"][1]
    n = int(out["attention_mask"][1].sum())
    assert row[n - 1] == 1 and row[n - 2] == 1  # "A </s> B </s>" with an empty B still has two EOS
    assert (row[n:] == 0).all()


def test_install_registers_ray_modules():
    assert rayshim.install() or True
    import ray
    from ray.data.preprocessors import BatchMapper  # noqa: F401
    from ray.train.batch_predictor import BatchPredictor  # noqa: F401
    from ray.train.predictor import Predictor  # noqa: F401

    assert getattr(ray, "__b200_shim__", False)
    with pytest.raises(NotImplementedError):
        from ray.train.huggingface import HuggingFaceTrainer

        HuggingFaceTrainer()


def test_notebook_flow_on_cpu_config1():
    """from_checkpoint -> predict -> to_pandas -> join, as notebook :875-934, tiny model on CPU."""
    from ray.data.preprocessors import BatchMapper

    ckpt = checkpoint_dir("tiny", seed=1)
    rows = synthetic_alpaca_rows(12)
    validation = rayshim.data.from_huggingface(rows).limit(10)
    prep = BatchMapper(make_preprocess_function(str(ckpt), max_length=32), batch_format="pandas", batch_size=4096)
    bp = make_batch_predictor(ckpt, model_cls=HFOnCpu, preprocessor=prep)
    prediction = bp.predict(validation, batch_size=4, max_new_tokens=6)
    input_pd, pred_pd = validation.to_pandas(), prediction.to_pandas()
    joined = input_pd.join(pred_pd, how="inner")
    assert len(joined) == 10 and "generated_output" in joined.columns
    assert all(isinstance(s, str) for s in joined["generated_output"])
    # row alignment: predicting one row alone gives the same text as the batched run
    single = bp.predict(validation.limit(1), batch_size=4, max_new_tokens=6).to_pandas()
    assert single["generated_output"][0] == pred_pd["generated_output"][0]


def test_predictor_matches_direct_generate():
    from anyscale_workshop_nyc_2023_b200.predictor import HuggingFaceModelPredictor
    from anyscale_workshop_nyc_2023_b200.synth import synthetic_token_batch
    from oracle.hf_anchor import hf_generate, load_hf_model
    from transformers import T5Tokenizer

    ckpt = checkpoint_dir("tiny", seed=1)
    model = load_hf_model(ckpt)
    tok = T5Tokenizer.from_pretrained(str(ckpt))
    ids, mask = synthetic_token_batch(5, 16, SPECS["tiny"].vocab_size, seed=3, lengths="uniform")
    p = HuggingFaceModelPredictor(model, tokenizer=tok)
    df = p.predict(pd.DataFrame({"input_ids": list(ids), "attention_mask": list(mask), "labels": list(ids)}), max_new_tokens=5)
    want = tok.batch_decode(hf_generate(model, ids, mask, 5), skip_special_tokens=True)
    assert df["generated_output"].tolist() == want
    only = p._predict_numpy({"input_ids": ids, "attention_mask": mask, "junk": ids}, feature_columns=["input_ids", "attention_mask"], max_new_tokens=5)
    assert only["generated_output"].tolist() == want


def test_sharding_roundtrip():
    for n, w in [(10, 1), (10, 3), (7, 8), (0, 2)]:
        per_rank = [[f"b{i}" for i in shard_block_indices(n, r, w)] for r in range(w)]
        assert restore_order(per_rank, n) == [f"b{i}" for i in range(n)]
    with pytest.raises(ValueError):
        shard_block_indices(4, 2, 2)


def test_lean_preprocess_equals_reference_style_tokenisation():
    """The lean tokenisation path (no per-row Python padding) must produce exactly the arrays of the reference's
    `tokenizer(a, b, padding="max_length", truncation=True, return_tensors="np")` (JOB/utils.py:23-31), including
    truncated pairs, empty strings and unicode."""
    import pandas as pd

    from anyscale_workshop_nyc_2023_b200 import preprocess
    from anyscale_workshop_nyc_2023_b200.synth import synthetic_alpaca_rows
    from anyscale_workshop_nyc_2023_b200.workload import ASSETS

    tokdir = str(ASSETS / "tokenizer")
    rows = pd.DataFrame(synthetic_alpaca_rows(200, seed=11))
    extra = pd.DataFrame({
        "instruction": ["", "x", "word " * 700, "short", "naive cafe \u00e9\u00e8 \u4f60\u597d"],
        "input": ["", "", "tail " * 50, "other " * 900, "\t tabs  and   spaces \n"],
    })
    batch = pd.concat([rows[["instruction", "input"]], extra], ignore_index=True)
    lean = preprocess.make_preprocess_function(tokdir, lean=True)(batch)
    ref = preprocess.make_preprocess_function(tokdir, lean=False)(batch)
    for k in ("input_ids", "attention_mask", "labels"):
        assert lean[k].dtype == ref[k].dtype == np.int64 and lean[k].shape == ref[k].shape
        assert np.array_equal(lean[k], ref[k]), k
    assert lean["attention_mask"][-3].sum() == lean["input_ids"].shape[1]  # a truncated pair fills the row


def test_streamed_cpu_stage_prefetch_keeps_order_and_propagates_errors():
    """rayshim.train._prefetch: the tokenisation of block i+1 runs on a producer thread while block i is being
    scored; results arrive in order, a producer exception surfaces in the consumer, an abandoned consumer does not
    leave the producer blocked."""
    import threading
    import time

    from anyscale_workshop_nyc_2023_b200.rayshim.train import _prefetch

    seen_threads = set()

    def slow_square(x):
        seen_threads.add(threading.current_thread().name)
        time.sleep(0.01)
        return x * x

    t0 = time.perf_counter()
    got = []
    for v in _prefetch(list(range(12)), slow_square):
        time.sleep(0.01)  # the "GPU" stage
        got.append(v)
    dt = time.perf_counter() - t0
    assert got == [i * i for i in range(12)]
    assert seen_threads == {"b200t5-cpu-stage"}
    assert dt < 0.22  # overlapped: ~12 x 0.01 + one stage of latency, not 12 x 0.02

    def boom(x):
        if x == 3:
            raise ValueError("tokeniser failed")
        return x

    with pytest.raises(ValueError, match="tokeniser failed"):
        list(_prefetch(list(range(6)), boom))
    it = _prefetch(list(range(100)), slow_square)
    assert next(it) == 0
    it.close()  # consumer walks away: the producer must stop instead of blocking on a full queue
    time.sleep(0.3)
    assert not any(t.name == "b200t5-cpu-stage" and t.is_alive() for t in threading.enumerate())


def test_batch_predictor_streamed_and_materialised_cpu_stage_agree():
    """BatchPredictor.predict with a GPU stage requested: the tokenisation runs either as a materialised CPU stage
    (AIR's behaviour) or streamed block by block under the scoring loop; same rows, same order, and the predictor
    never sees a preprocessor of its own in either case."""
    from anyscale_workshop_nyc_2023_b200.rayshim.train import BatchPredictor, Predictor

    calls = []

    class UpperPredictor(Predictor):
        @classmethod
        def from_checkpoint(cls, checkpoint, use_gpu=False, **kw):
            p = cls(preprocessor=checkpoint.get_preprocessor())
            p.use_gpu = use_gpu
            return p

        def _predict_pandas(self, data, **kw):
            calls.append(len(data))
            assert self.get_preprocessor() is None  # the CPU stage ran outside the predictor
            return pd.DataFrame({"generated_output": [f"{a}|{b}".upper() for a, b in zip(data["a"], data["n"])]})

    class Ckpt:
        def __init__(self, prep):
            self._prep = prep

        def get_preprocessor(self):
            return self._prep

    prep = rayshim.data.BatchMapper(lambda b: pd.DataFrame({"a": b["instruction"].str[:5], "n": b["instruction"].str.len()}),
                                    batch_format="pandas")
    ds = rayshim.data.from_huggingface(synthetic_alpaca_rows(37))
    bp = BatchPredictor.from_checkpoint(Ckpt(prep), UpperPredictor)
    outs = {}
    for streamed in (False, True):
        calls.clear()
        outs[streamed] = bp.predict(ds, batch_size=8, num_gpus_per_worker=1, pipeline_cpu_stage=streamed).to_pandas()
        assert calls == [8, 8, 8, 8, 5]
    assert outs[True]["generated_output"].tolist() == outs[False]["generated_output"].tolist()
    want = [f"{s[:5]}|{len(s)}".upper() for s in synthetic_alpaca_rows(37)["instruction"]]
    assert outs[True]["generated_output"].tolist() == want


def test_example_script_reference_leg_runs_on_cpu():
    """examples/batch_inference.py is the notebook's flow as a script; its `--model-cls hf` leg (the dependency's own
    model through the same shim, predictor and preprocess function) must run without a GPU."""
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    proc = subprocess.run([sys.executable, str(root / "examples" / "batch_inference.py"), "--model-cls", "hf", "--n", "5",
                           "--max-new-tokens", "4", "--batch-size", "2"], capture_output=True, text=True, timeout=600, cwd=str(root))
    assert proc.returncode == 0, proc.stderr[-2000:]
    assert "5 prompts in" in proc.stdout and "model_cls=HFModelOnCpu" in proc.stdout
