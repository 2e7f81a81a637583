"""Tokenisation stage feeding the path: mirror of `preprocess_function`
(NLP_workloads/Anyscale_job/utils.py:6-33; notebook :260-287).

Same contract - a pandas batch with string columns "instruction" and "input" becomes
{"input_ids", "attention_mask", "labels"} int64 [N, model_max_length] via pair encoding
`A </s> B </s>`, truncation and max-length padding, `labels` a copy of `input_ids` - but the
tokenizer is loaded once per process instead of once per 4096-row batch (the reference
re-instantiates it at JOB/utils.py:20-21), and its location is configurable because
"google/flan-t5-base" cannot be fetched offline.

SURVEY section 8(f) row 2, first step: the encode itself is ~1 % of what `tokenizer(..., padding="max_length",
return_tensors="np")` costs - the rest is Python-side padding of every row to 512 and list -> array conversion.
`lean=True` (default) encodes the pairs without padding through the tokenizer's own backend (same template
`A </s> B </s>`, same longest-first truncation) and builds the padded int64 arrays with two vectorised numpy
operations: identical arrays (tests/test_shim_cpu.py), 3-4x the prompts/s (tools/bench_preprocess.py).
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


def encode_pairs_padded(tokenizer, first, second, max_length: int = None):
    """`tokenizer(first, second, padding="max_length", truncation=True, return_tensors="np")` without the
    per-row Python padding: returns (input_ids, attention_mask), int64 [N, max_length]."""
    from itertools import chain

    import numpy as np

    length = int(max_length or tokenizer.model_max_length)
    enc = tokenizer(list(first), list(second), padding=False, truncation=True, max_length=length)
    rows = enc["input_ids"]
    n = len(rows)
    lens = np.fromiter((len(r) for r in rows), dtype=np.int64, count=n)
    ids = np.full((n, length), tokenizer.pad_token_id, dtype=np.int64)
    keep = np.arange(length)[None, :] < lens[:, None]
    if getattr(tokenizer, "padding_side", "right") != "right":
        keep = keep[:, ::-1]
    ids[keep] = np.fromiter(chain.from_iterable(rows), dtype=np.int64, count=int(lens.sum()))
    return ids, keep.astype(np.int64)


def make_preprocess_function(tokenizer_name_or_path: str = None, max_length: int = None, lean: bool = True):
    def preprocess_function(batch: Dict[str, Any]) -> Dict[str, Any]:
        tokenizer = get_tokenizer(tokenizer_name_or_path)
        if lean:
            ids, mask = encode_pairs_padded(tokenizer, batch["instruction"], batch["input"], max_length)
            return {"input_ids": ids, "attention_mask": mask, "labels": ids.copy()}
        kw = {} if max_length is None else {"max_length": max_length}
        enc = tokenizer(list(batch["instruction"]), list(batch["input"]), padding="max_length", truncation=True,
                        return_tensors="np", **kw)
        out = dict(enc)
        out["labels"] = out["input_ids"].copy()
        return out

    return preprocess_function


preprocess_function = make_preprocess_function()
