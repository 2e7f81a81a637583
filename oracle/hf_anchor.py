"""The dependency the reference actually executes (transformers' T5ForConditionalGeneration),
instantiated on a synthetic checkpoint. TEST INFRASTRUCTURE (see oracle/t5_oracle.py header).

transformers 5.5.0 force-ties lm_head to the shared embedding (configuration_t5.py:82-83), while
real FLAN-T5 checkpoints carry a separate lm_head; the head is therefore re-assigned after
construction (SURVEY 8c "oracle caveats"). `tie_word_embeddings=False` in the config keeps
`scale_decoder_outputs` False, as for FLAN-T5.
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import torch

from anyscale_workshop_nyc_2023_b200.synth import load_state_dict_f32


def load_hf_model(ckpt_dir, dtype=torch.float32, device="cpu"):
    from transformers import T5Config, T5ForConditionalGeneration

    ckpt_dir = Path(ckpt_dir)
    cfg = json.loads((ckpt_dir / "config.json").read_text())
    for k in ("architectures", "model_type", "torch_dtype", "dense_act_fn", "is_gated_act"):
        cfg.pop(k, None)
    cfg["tie_word_embeddings"] = False
    config = T5Config(**cfg)
    config._attn_implementation = "eager"
    with torch.device("cpu"):
        model = T5ForConditionalGeneration(config)
    sd = {k: torch.from_numpy(np.array(v)) for k, v in load_state_dict_f32(ckpt_dir).items()}
    lm = sd.pop("lm_head.weight")
    sd["encoder.embed_tokens.weight"] = sd["shared.weight"]
    sd["decoder.embed_tokens.weight"] = sd["shared.weight"]
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [m for m in missing if m != "lm_head.weight"]
    assert not missing and not unexpected, (missing, unexpected)
    model.lm_head.weight = torch.nn.Parameter(lm.clone())  # untie
    assert model.lm_head.weight.data_ptr() != model.shared.weight.data_ptr()
    model = model.to(dtype=dtype, device=device).eval()
    if dtype == torch.float16:
        # what from_pretrained(torch_dtype=float16) does for T5 (`_keep_in_fp32_modules = ["wo"]`): the feed-forward
        # output projection keeps its fp32 checkpoint values
        for name, mod in model.named_modules():
            if name.endswith("DenseReluDense.wo"):
                mod.weight = torch.nn.Parameter(sd[f"{name}.weight"].to(device=device, dtype=torch.float32).clone(), requires_grad=False)
    model.generation_config.decoder_start_token_id = config.decoder_start_token_id
    model.generation_config.eos_token_id = config.eos_token_id
    model.generation_config.pad_token_id = config.pad_token_id
    return model


@torch.no_grad()
def hf_generate(model, ids: np.ndarray, mask: np.ndarray, max_new_tokens: int, min_new_tokens: int = 0) -> np.ndarray:
    dev = model.device
    kw = dict(input_ids=torch.from_numpy(ids).to(dev), attention_mask=torch.from_numpy(mask).to(dev),
              labels=torch.from_numpy(ids).to(dev),  # the reference passes labels too (JOB/utils.py:31); HF drops it
              max_new_tokens=max_new_tokens, do_sample=False, num_beams=1)
    if min_new_tokens:
        kw["min_new_tokens"] = min_new_tokens
    return model.generate(**kw).cpu().numpy()


@torch.no_grad()
def hf_teacher_forced_logits(model, ids: np.ndarray, mask: np.ndarray, decoder_input_ids: np.ndarray) -> np.ndarray:
    dev = model.device
    out = model(input_ids=torch.from_numpy(ids).to(dev), attention_mask=torch.from_numpy(mask).to(dev),
                decoder_input_ids=torch.from_numpy(decoder_input_ids).to(dev), use_cache=False)
    return out.logits.float().cpu().numpy()
