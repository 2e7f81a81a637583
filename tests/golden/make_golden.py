"""Generates tests/golden/*.npz by running the reference's own dependency (transformers'
T5ForConditionalGeneration.generate, eager attention) on seeded synthetic checkpoints, in the
build container (CPU). Re-run with:  python tests/golden/make_golden.py   (--fp16: only the *_fp16.npz files)
The fixtures pin oracle/t5_oracle.py; they are environment-stamped (torch / transformers versions).
"""
import sys
from pathlib import Path

import numpy as np
import torch
import transformers

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from anyscale_workshop_nyc_2023_b200.synth import SPECS, save_checkpoint, synthetic_token_batch  # noqa: E402
from oracle.hf_anchor import hf_generate, hf_teacher_forced_logits, load_hf_model  # noqa: E402

CASES = [  # name, spec, weight seed, B, S, max_new, input seed, lengths
    ("tiny_a", "tiny", 1, 6, 24, 12, 101, "uniform"),
    ("tiny_full", "tiny", 1, 4, 16, 10, 102, "full"),
    ("mini_a", "mini", 2, 5, 40, 16, 103, "uniform"),
]


def main_fp16():
    """fp16 goldens (the notebook's literal torch_dtype, NB:882) in their own files, so that the fp32/bf16 fixtures
    above stay byte-identical: tests/golden/<case>_fp16.npz."""
    import tempfile

    out_dir = Path(__file__).resolve().parent
    for name, spec_name, wseed, B, S, T, iseed, lengths in CASES:
        spec = SPECS[spec_name]
        ids, mask = synthetic_token_batch(B, S, spec.vocab_size, iseed, lengths)
        with tempfile.TemporaryDirectory() as d:
            save_checkpoint(d, spec, seed=wseed)
            m = load_hf_model(d, dtype=torch.float16)
            assert m.encoder.block[0].layer[1].DenseReluDense.wo.weight.dtype == torch.float32
            toks = hf_generate(m, ids, mask, T)
            res = {"ids": ids, "mask": mask, "tokens_fp16": toks, "forced_fp16": hf_generate(m, ids, mask, T, min_new_tokens=T),
                   "logits_fp16": hf_teacher_forced_logits(m, ids, mask, toks[:, :-1]).astype(np.float32)}
            with torch.no_grad():
                enc = m.encoder(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(mask)).last_hidden_state
            res["enc_fp16"] = enc.float().numpy()
        res["meta"] = np.array([f"spec={spec_name} wseed={wseed} max_new={T} torch={torch.__version__} transformers={transformers.__version__} wo=fp32"])
        np.savez_compressed(out_dir / f"{name}_fp16.npz", **res)
        print(name + "_fp16", {k: v.shape for k, v in res.items() if k != "meta"})


def main():
    import tempfile

    out_dir = Path(__file__).resolve().parent
    torch.manual_seed(0)
    for name, spec_name, wseed, B, S, T, iseed, lengths in CASES:
        spec = SPECS[spec_name]
        ids, mask = synthetic_token_batch(B, S, spec.vocab_size, iseed, lengths)
        with tempfile.TemporaryDirectory() as d:
            save_checkpoint(d, spec, seed=wseed)
            res = {"ids": ids, "mask": mask}
            for tag, dtype in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
                m = load_hf_model(d, dtype=dtype)
                toks = hf_generate(m, ids, mask, T)
                forced = hf_generate(m, ids, mask, T, min_new_tokens=T)
                res[f"tokens_{tag}"] = toks
                res[f"forced_{tag}"] = forced
                res[f"logits_{tag}"] = hf_teacher_forced_logits(m, ids, mask, toks[:, :-1]).astype(np.float32)
                with torch.no_grad():
                    enc = m.encoder(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(mask)).last_hidden_state
                res[f"enc_{tag}"] = enc.float().numpy()
        res["meta"] = np.array([f"spec={spec_name} wseed={wseed} max_new={T} torch={torch.__version__} transformers={transformers.__version__}"])
        np.savez_compressed(out_dir / f"{name}.npz", **res)
        print(name, {k: v.shape for k, v in res.items() if k != "meta"})


if __name__ == "__main__":
    if "--fp16" in sys.argv:
        main_fp16()
    else:
        main()
        main_fp16()
