"""Trains the synthetic SentencePiece unigram model shipped under
anyscale_workshop_nyc_2023_b200/assets/tokenizer (no FLAN-T5 spiece.model exists offline).
T5 conventions: pad=0, eos=1, unk=2, no bos; `model_max_length` 512 as in FLAN-T5's
tokenizer_config.json, which is what makes `padding="max_length"` pad to 512.
Re-run:  python tools/make_tokenizer.py"""
import io
import json
import sys
from pathlib import Path

import sentencepiece as spm

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from anyscale_workshop_nyc_2023_b200.synth import synthetic_alpaca_rows  # noqa: E402

OUT = ROOT / "anyscale_workshop_nyc_2023_b200" / "assets" / "tokenizer"


def main():
    rows = synthetic_alpaca_rows(4000, seed=1)
    corpus = rows["instruction"] + [x for x in rows["input"] if x] + rows["output"]
    model = io.BytesIO()
    spm.SentencePieceTrainer.train(sentence_iterator=iter(corpus), model_writer=model, vocab_size=320,
                                   model_type="unigram", pad_id=0, eos_id=1, unk_id=2, bos_id=-1,
                                   character_coverage=1.0, hard_vocab_limit=False, num_threads=1)
    OUT.mkdir(parents=True, exist_ok=True)
    (OUT / "spiece.model").write_bytes(model.getvalue())
    cfg = {"model_max_length": 512, "eos_token": "</s>", "pad_token": "<pad>", "unk_token": "<unk>", "extra_ids": 100,
           "tokenizer_class": "T5Tokenizer"}
    (OUT / "tokenizer_config.json").write_text(json.dumps(cfg, indent=1))
    print("wrote", OUT, len(model.getvalue()), "bytes")


if __name__ == "__main__":
    main()
