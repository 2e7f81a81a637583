"""CPU-only: the reference's OWN hot-path file, unmodified, driven through the shim.

`NLP_workloads/Anyscale_job/predictor.py:14-106` (HuggingFaceModelPredictor) and `utils.py:6-33`
(preprocess_function) are loaded straight from the reference checkout with importlib - nothing is copied - after
`rayshim.install()` has made `import ray` resolve (predictor.py:7). The checkout exists in the build container only
(the GPU boxes have no /root/reference), so these tests skip there; what they establish carries to the GPU through
two pinned equalities:

    reference predictor + HF model  ==  direct HF generate                    (here)
    reference predictor             ==  this package's mirror, same model      (here)
    mirror + B200 model             ==  B200 generate == HF on the same GPU    (tests/test_model_gpu.py)

Flow reproduced: flan-t5-batch-inference.py:119-138 (from_checkpoint -> predict -> to_pandas -> join).
"""
import numpy as np
import pandas as pd
import pytest
import torch

from anyscale_workshop_nyc_2023_b200 import rayshim, refsource
from anyscale_workshop_nyc_2023_b200.preprocess import make_preprocess_function
from anyscale_workshop_nyc_2023_b200.synth import SPECS, synthetic_alpaca_rows, synthetic_token_batch
from anyscale_workshop_nyc_2023_b200.workload import checkpoint_dir

pytestmark = pytest.mark.skipif(refsource.reference_root() is None, reason="the reference checkout is not present on this machine")


class HFOnCpu:
    """model_cls stand-in with from_pretrained(dir, **kw): the dependency's own model on CPU (what
    `T5ForConditionalGeneration` is in the reference script, minus the hub download)."""

    @staticmethod
    def from_pretrained(path, **kw):
        from oracle.hf_anchor import load_hf_model

        assert kw.get("torch_dtype") is torch.float16 and kw.get("device_map") == "auto"  # forwarded untouched (NB:881-882)
        return load_hf_model(path, dtype=torch.float32, device="cpu")


def test_reference_predictor_file_is_loaded_unmodified():
    mod = refsource.load_reference_predictor_module()
    src = (refsource.reference_root() / "NLP_workloads" / "Anyscale_job" / "predictor.py").read_text()
    assert mod.__file__.startswith(str(refsource.reference_root()))
    assert "class HuggingFaceModelPredictor(Predictor)" in src and "from ray.train.predictor import Predictor" in src
    from anyscale_workshop_nyc_2023_b200.rayshim.train import Predictor

    assert issubclass(mod.HuggingFaceModelPredictor, Predictor)  # its base class is the shim's


def test_reference_script_flow_with_the_unmodified_predictor_class():
    """BatchPredictor.from_checkpoint(checkpoint=..., predictor_cls=<reference class>, model_cls=..., tokenizer=T5Tokenizer,
    use_gpu=..., device_map="auto", torch_dtype=torch.float16) -> predict(ds, num_gpus_per_worker=..., batch_size=...,
    max_new_tokens=...) -> to_pandas -> join, as flan-t5-batch-inference.py:119-138."""
    from ray.data.preprocessors import BatchMapper
    from ray.train.batch_predictor import BatchPredictor
    from transformers import T5Tokenizer

    from anyscale_workshop_nyc_2023_b200.rayshim.train import HuggingFaceCheckpoint
    from oracle.hf_anchor import hf_generate, load_hf_model

    RefPredictor = refsource.load_reference_predictor_module().HuggingFaceModelPredictor
    ckpt = checkpoint_dir("tiny", seed=1)
    use_gpu = False
    validation_dataset = rayshim.data.from_huggingface(synthetic_alpaca_rows(11)).limit(10)
    fn = make_preprocess_function(str(ckpt), max_length=32, lean=False)  # the reference's own tokenizer call
    checkpoint = HuggingFaceCheckpoint.from_directory(str(ckpt))
    checkpoint.set_preprocessor(BatchMapper(fn, batch_format="pandas", batch_size=4096))
    predictor = BatchPredictor.from_checkpoint(checkpoint=checkpoint, predictor_cls=RefPredictor, model_cls=HFOnCpu,
                                               tokenizer=T5Tokenizer, use_gpu=use_gpu, device_map="auto", torch_dtype=torch.float16)
    prediction = predictor.predict(validation_dataset, num_gpus_per_worker=int(use_gpu), batch_size=4, max_new_tokens=7)
    input_data_pd = validation_dataset.to_pandas()
    prediction_pd = prediction.to_pandas()
    outputs = input_data_pd.join(prediction_pd, how="inner").head(n=7)
    assert len(outputs) == 7 and "generated_output" in outputs.columns and "instruction" in outputs.columns
    # row for row what the dependency generates directly for the same tokenised prompts
    enc = fn(input_data_pd)
    model = load_hf_model(ckpt)
    tok = T5Tokenizer.from_pretrained(str(ckpt))
    want = []
    for lo in range(0, 10, 4):
        want += tok.batch_decode(hf_generate(model, enc["input_ids"][lo:lo + 4], enc["attention_mask"][lo:lo + 4], 7), skip_special_tokens=True)
    assert prediction_pd["generated_output"].tolist() == want


def test_mirror_predictor_equals_the_reference_predictor():
    """Same model, same inputs, same kwargs through both classes: identical DataFrames - for the dict-of-columns input
    of the hot path, with `labels` present (JOB/utils.py:31), with feature_columns, and for max_length-default calls."""
    from transformers import T5Tokenizer

    from anyscale_workshop_nyc_2023_b200.predictor import HuggingFaceModelPredictor as Mirror
    from oracle.hf_anchor import load_hf_model

    Ref = refsource.load_reference_predictor_module().HuggingFaceModelPredictor
    ckpt = checkpoint_dir("tiny", seed=1)
    model = load_hf_model(ckpt)
    tok = T5Tokenizer.from_pretrained(str(ckpt))
    ids, mask = synthetic_token_batch(6, 20, SPECS["tiny"].vocab_size, seed=17, lengths="uniform")
    ref, mir = Ref(model, tokenizer=tok), Mirror(model, tokenizer=tok)
    cases = [
        ({"input_ids": ids, "attention_mask": mask, "labels": ids.copy()}, dict(max_new_tokens=6)),
        ({"input_ids": ids, "attention_mask": mask, "labels": ids.copy(), "junk": ids}, dict(feature_columns=["input_ids", "attention_mask"], max_new_tokens=4)),
        ({"input_ids": ids, "attention_mask": mask}, dict()),  # GenerationConfig default max_length = 20
        ({"input_ids": ids, "attention_mask": mask}, dict(max_new_tokens=5, min_new_tokens=5)),
    ]
    for data, kw in cases:
        a = ref._predict_numpy({k: v.copy() for k, v in data.items()}, **kw)
        b = mir._predict_numpy({k: v.copy() for k, v in data.items()}, **kw)
        assert list(a.columns) == list(b.columns) == ["generated_output"]
        assert a["generated_output"].tolist() == b["generated_output"].tolist()
    # the classmethod: same constructor contract (tokenizer class resolved through the checkpoint)
    from anyscale_workshop_nyc_2023_b200.rayshim.train import HuggingFaceCheckpoint

    ck = HuggingFaceCheckpoint.from_directory(str(ckpt))
    for cls in (Ref, Mirror):
        p = cls.from_checkpoint(ck, HFOnCpu, tokenizer=T5Tokenizer, use_gpu=False, device_map="auto", torch_dtype=torch.float16)
        assert p.use_gpu is False and p.tokenizer.__class__.__name__ == "T5Tokenizer" and p.get_preprocessor() is None


def test_reference_preprocess_function_equals_the_mirror(monkeypatch):
    """utils.py:6-33 hard-codes `T5Tokenizer.from_pretrained("google/flan-t5-base")` (a hub download); with that one
    call pointed at the local tokenizer files the unmodified function and this package's lean mirror produce identical
    arrays."""
    from transformers import T5Tokenizer

    from anyscale_workshop_nyc_2023_b200.workload import ASSETS

    utils = refsource.load_reference_utils_module()
    real = T5Tokenizer.from_pretrained
    monkeypatch.setattr(utils.T5Tokenizer, "from_pretrained",
                        classmethod(lambda cls, name, *a, **k: real(str(ASSETS / "tokenizer"), *a, **k)))
    batch = pd.DataFrame(synthetic_alpaca_rows(40, seed=5))[["instruction", "input"]]
    ref = utils.preprocess_function(batch)
    mir = make_preprocess_function(str(ASSETS / "tokenizer"))(batch)
    assert set(ref) == set(mir) == {"input_ids", "attention_mask", "labels"}
    for k in ref:
        assert np.array_equal(np.asarray(ref[k]), mir[k]), k
