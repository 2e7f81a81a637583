"""Tokenisation stage feeding the path: mirror of `preprocess_function`
(NLP_workloads/Anyscale_job/utils.py:6-33; notebook :260-287).

Same contract - a pandas batch with string columns "instruction" and "input" becomes
{"input_ids", "attention_mask", "labels"} int64 [N, model_max_length] via pair encoding
`A </s> B </s>`, truncation and max-length padding, `labels` a copy of `input_ids` - but the
tokenizer is loaded once per process instead of once per 4096-row batch (the reference
re-instantiates it at JOB/utils.py:20-21), and its location is configurable because
"google/flan-t5-base" cannot be fetched offline.
"""
from __future__ import annotations

import os
from typing import Any, Dict

_TOKENIZERS: Dict[str, Any] = {}
DEFAULT_TOKENIZER = os.environ.get("B200T5_TOKENIZER", "google/flan-t5-base")


def get_tokenizer(name_or_path: str = None):
    from transformers import T5Tokenizer

    key = str(name_or_path or DEFAULT_TOKENIZER)
    tok = _TOKENIZERS.get(key)
    if tok is None:
        tok = T5Tokenizer.from_pretrained(key)
        _TOKENIZERS[key] = tok
    return tok


def make_preprocess_function(tokenizer_name_or_path: str = None, max_length: int = None):
    def preprocess_function(batch: Dict[str, Any]) -> Dict[str, Any]:
        tokenizer = get_tokenizer(tokenizer_name_or_path)
        kw = {} if max_length is None else {"max_length": max_length}
        enc = tokenizer(list(batch["instruction"]), list(batch["input"]), padding="max_length", truncation=True,
                        return_tensors="np", **kw)
        out = dict(enc)
        out["labels"] = out["input_ids"].copy()
        return out

    return preprocess_function


preprocess_function = make_preprocess_function()
