"""`B200T5ForConditionalGeneration`: the object the reference predictor holds as ``self.model``.

The reference's seam is duck-typed (NLP_workloads/Anyscale_job/predictor.py):
    checkpoint.get_model(model_cls, **get_model_kwargs)   :68   -> model_cls.from_pretrained(dir, **kw)
    self.model.device                                      :98
    self.model.generate(**generate_kwargs) -> LongTensor   :102  (consumed by tokenizer.batch_decode :104)
so passing ``model_cls=B200T5ForConditionalGeneration`` to ``BatchPredictor.from_checkpoint``
(notebook lines 875-883) swaps the Hugging Face eager model for the sm_100a kernels behind
libb200t5.so without touching the predictor.

PyTorch is used for device memory and streams only; all arithmetic runs in the CUDA library.
There is no CPU path: constructing the model without a B200 raises.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import threading
import warnings
from pathlib import Path
from types import SimpleNamespace
from typing import Any, Dict, Optional

import numpy as np
import torch

from . import _lib
from .synth import read_safetensors

_POOL_MAX_S = 512  # the slot pool admits prompts through the packed tcgen05 encoder (csrc: kEncTcMaxS)
_HF_DEFAULT_MAX_LENGTH = 20  # GenerationConfig default the notebook's single-prompt cell relies on (NB:577)

_IGNORED_WEIGHTS = ("decoder.block.0.layer.1.EncDecAttention.relative_attention_bias.weight",)
_ALIASES = ("encoder.embed_tokens.weight", "decoder.embed_tokens.weight")


def _chk(model, rc: int, handle=None) -> None:
    _lib.check(rc, handle, model._lib)


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class B200T5ForConditionalGeneration:
    """FLAN-T5 greedy generation on one B200. API subset of transformers.T5ForConditionalGeneration
    that the workshop's predictor and notebook cells touch."""

    main_input_name = "input_ids"

    def __init__(self, config: Dict[str, Any], device: torch.device, compute_dtype: torch.dtype = torch.bfloat16):
        if device.type != "cuda":
            raise RuntimeError("B200T5ForConditionalGeneration runs on a B200 only; there is no CPU fallback")
        if compute_dtype not in (torch.bfloat16, torch.float16):
            raise ValueError(f"compute dtype must be bfloat16 or float16, got {compute_dtype}")
        # one shared library per numerics contract: bf16 everywhere, or the notebook's literal torch_dtype=float16
        # (NB:882) with transformers' fp32 `wo` / fp32 residual stream
        self._dtype = compute_dtype
        self._lib = _lib.load("fp16" if compute_dtype == torch.float16 else "bf16")
        self._device = device
        self.config = SimpleNamespace(**config)
        self.generation_config = SimpleNamespace(
            max_length=_HF_DEFAULT_MAX_LENGTH,
            eos_token_id=config.get("eos_token_id", 1),
            pad_token_id=config.get("pad_token_id", 0),
            decoder_start_token_id=config.get("decoder_start_token_id", config.get("pad_token_id", 0)),
        )
        ffp = config.get("feed_forward_proj", "gated-gelu")
        cfg = _lib.Config(
            vocab_size=config["vocab_size"], d_model=config["d_model"], d_kv=config["d_kv"], d_ff=config["d_ff"],
            num_heads=config["num_heads"], num_layers=config["num_layers"],
            num_decoder_layers=config.get("num_decoder_layers") or config["num_layers"],
            relative_attention_num_buckets=config.get("relative_attention_num_buckets", 32),
            relative_attention_max_distance=config.get("relative_attention_max_distance", 128),
            layer_norm_epsilon=config.get("layer_norm_epsilon", 1e-6),
            pad_token_id=self.generation_config.pad_token_id,
            eos_token_id=self.generation_config.eos_token_id,
            decoder_start_token_id=self.generation_config.decoder_start_token_id,
            is_gated_gelu=1 if ffp == "gated-gelu" else 0,
            # HF decides "scale decoder outputs" from tie_word_embeddings (configuration_t5.py:82)
            scale_decoder_outputs=0 if config.get("tie_word_embeddings", True) is False else 1,
        )
        h = C.c_void_p()
        index = device.index if device.index is not None else torch.cuda.current_device()
        _chk(self, self._lib.b200t5_create(C.byref(cfg), index, C.byref(h)))
        self._h = h
        self._index = index
        self.last_lengths: Optional[torch.Tensor] = None
        # generate() switches to the continuous-batching path for batches larger than pool_size; that path runs
        # pool_slots decode slots (512 measured best on B200 for FLAN-T5-base: +18 % over 256, profiles/stream_r1.md)
        self.pool_size = int(os.environ.get("B200T5_POOL", "256"))
        # one handle = one execution plan: calls are serialised. Two host threads may alternate on it (the detokenise /
        # DataFrame tail of block i overlaps the GPU part of block i+1: rayshim/train.py:_overlap_tail).
        self._gpu_lock = threading.RLock()
        self.pool_slots = int(os.environ.get("B200T5_POOL_SLOTS", "512"))

    # ------------------------------------------------------------------ loading
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *model_args, device_map=None, torch_dtype=None,
                        dtype=None, device=None, **kwargs) -> "B200T5ForConditionalGeneration":
        """Load a Hugging Face T5 directory (config.json + model.safetensors | pytorch_model.bin).

        `device_map="auto"` / `torch_dtype=` are accepted as the notebook passes them (NB:881-882).
        torch_dtype=bfloat16 (default) and torch_dtype=float16 select the two numerics contracts the library
        implements; float16 is transformers' mode for T5: fp16 weights and activations, `wo` kept in fp32, fp32
        residual stream from the first feed-forward block on. float32 is honoured as "load and round to
        bfloat16" with a warning.
        """
        path = Path(pretrained_model_name_or_path)
        if not (path / "config.json").exists():
            raise FileNotFoundError(f"{path} is not a Hugging Face checkpoint directory (no config.json); "
                                    "hub downloads are not available offline")
        want = dtype if dtype is not None else torch_dtype
        compute = torch.float16 if want in (torch.float16, "float16", "half") else torch.bfloat16
        if want not in (None, torch.bfloat16, "bfloat16", "auto", torch.float16, "float16", "half"):
            warnings.warn(f"B200T5ForConditionalGeneration computes in bfloat16 or float16; requested {want} weights "
                          "are rounded to bfloat16 on load", stacklevel=2)
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: B200T5ForConditionalGeneration has no CPU fallback")
        if device is None:
            if device_map is None or device_map in ("auto", "balanced", "sequential"):
                device = torch.device("cuda", torch.cuda.current_device())  # one replica per process/GPU
            elif isinstance(device_map, dict):
                device = torch.device(next(iter(device_map.values())))
            else:
                device = torch.device(device_map)
        if torch.device(device).type == "cuda" and torch.device(device).index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        device = torch.device(device)
        config = json.loads((path / "config.json").read_text())
        model = cls(config, device, compute)
        model._load_weights(path)
        return model

    def _load_weights(self, path: Path) -> None:
        st = path / "model.safetensors"
        tensors: Dict[str, torch.Tensor] = {}
        if st.exists():
            for name, (dt, shape, arr) in read_safetensors(st).items():
                a = np.array(arr)  # copy out of the memmap
                if dt == "BF16":
                    t = torch.from_numpy(a.view(np.int16)).view(torch.bfloat16)
                else:
                    t = torch.from_numpy(a)
                tensors[name] = t
        elif (path / "pytorch_model.bin").exists():
            tensors = torch.load(path / "pytorch_model.bin", map_location="cpu", weights_only=True)
        else:
            raise FileNotFoundError(f"{path}: neither model.safetensors nor pytorch_model.bin found")
        self.load_state_dict(tensors)

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True):
        codes = {torch.bfloat16: _lib.DTYPE_BF16, torch.float16: _lib.DTYPE_F16, torch.float32: _lib.DTYPE_F32}
        with torch.cuda.device(self._index):
            for name, t in state_dict.items():
                if name in _IGNORED_WEIGHTS or name in _ALIASES:
                    continue
                if t.dtype not in codes:
                    t = t.float()
                dev = t.to(self._device).contiguous()
                shape = (C.c_int64 * dev.dim())(*dev.shape)
                _chk(self, self._lib.b200t5_set_weight(self._h, name.encode(), _ptr(dev), codes[dev.dtype], shape,
                                                       dev.dim()), self._h)
                del dev
            _chk(self, self._lib.b200t5_finalize(self._h), self._h)
        return SimpleNamespace(missing_keys=[], unexpected_keys=[])

    # ------------------------------------------------------------------ nn.Module-ish surface
    @property
    def device(self) -> torch.device:
        return self._device

    @property
    def dtype(self) -> torch.dtype:
        return self._dtype

    def eval(self):
        return self

    def to(self, *args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, (str, torch.device)) and torch.device(a) != self._device and torch.device(a).type != "cuda":
                raise RuntimeError("B200T5ForConditionalGeneration cannot be moved off the GPU")
        return self

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.b200t5_destroy(h)
            except Exception:
                pass
            self._h = None

    # ------------------------------------------------------------------ generation
    def _gen_params(self, max_new_tokens, max_length, min_new_tokens, min_length, eos_token_id, pad_token_id,
                    decoder_start_token_id, poll_interval) -> _lib.GenParams:
        # GenerationMixin._prepare_generated_length (generation/utils.py:1619-1639): the decoder
        # prompt is the single start token, so max_length = max_new_tokens + 1.
        if max_new_tokens is None:
            max_length = self.generation_config.max_length if max_length is None else max_length
            max_new_tokens = int(max_length) - 1
        if max_new_tokens < 1:
            raise ValueError(f"max_new_tokens must be >= 1, got {max_new_tokens}")
        if min_new_tokens is None:
            min_new_tokens = max(int(min_length) - 1, 0) if min_length else 0
        if isinstance(eos_token_id, (list, tuple)):
            if len(eos_token_id) != 1:
                raise NotImplementedError("multiple eos_token_id values are not supported")
            eos_token_id = eos_token_id[0]
        return _lib.GenParams(
            max_new_tokens=int(max_new_tokens), min_new_tokens=int(min(min_new_tokens, max_new_tokens)),
            eos_token_id=-1 if eos_token_id is None else int(eos_token_id),
            pad_token_id=-1 if pad_token_id is None else int(pad_token_id),
            decoder_start_token_id=-1 if decoder_start_token_id is None else int(decoder_start_token_id),
            poll_interval=int(poll_interval),
        )

    @torch.no_grad()
    def generate(self, input_ids=None, attention_mask=None, *, max_new_tokens=None, max_length=None,
                 min_new_tokens=None, min_length=None, eos_token_id=None, pad_token_id=None,
                 decoder_start_token_id=None, do_sample=False, num_beams=1, poll_interval=8,
                 **unused) -> torch.LongTensor:
        """Greedy `generate`: returns int64 [B, 1+T'] on `self.device`, column 0 the decoder start
        token, rows padded after their EOS, T' = steps until every row finished (<= max_new_tokens).
        `labels` (which the reference passes, JOB/utils.py:31) and other HF kwargs that do not
        change greedy decoding are accepted and ignored, as HF itself does (generation/utils.py:583)."""
        if input_ids is None:
            input_ids = unused.pop("inputs", None)
        if input_ids is None:
            raise ValueError("input_ids is required")
        if do_sample or (num_beams is not None and num_beams != 1):
            raise NotImplementedError("only greedy decoding (do_sample=False, num_beams=1) is implemented")
        for k in ("temperature", "top_k", "top_p", "repetition_penalty", "no_repeat_ngram_size", "num_return_sequences"):
            v = unused.get(k)
            if v is not None and not (isinstance(v, (int, float)) and v == 1):
                raise NotImplementedError(f"generate({k}=...) is not supported by the B200 path")
        for k in ("logits_processor", "stopping_criteria", "forced_bos_token_id", "decoder_input_ids", "encoder_outputs"):
            if unused.get(k) is not None:  # (may be tensors: no truth-value tests)
                raise NotImplementedError(f"generate({k}=...) is not supported by the B200 path")
        gp = self._gen_params(max_new_tokens, max_length, min_new_tokens, min_length, eos_token_id, pad_token_id,
                              decoder_start_token_id, poll_interval)
        host = torch.as_tensor(input_ids)
        if host.dim() != 2:
            raise ValueError(f"input_ids must be [batch, seq], got {tuple(host.shape)}")
        if self.takes_host_batches(host.shape[0], host.shape[1]) and host.device.type == "cpu":
            # more rows than one pool of decode slots, still in host memory: the slot pool admits prompts from host
            # buffers as slots free up, so nothing is copied to the device (and back) up front
            return self._generate_pool_from_host(host, attention_mask, gp)
        ids = host.to(device=self._device, dtype=torch.long).contiguous()
        B, S = ids.shape
        if ids.numel():
            lo, hi = torch.aminmax(ids)  # one kernel, one synchronisation
            if bool(((lo < 0) | (hi >= self.config.vocab_size)).item()):
                raise IndexError("input_ids contain token ids outside [0, vocab_size)")
        mask = None
        if attention_mask is not None:
            mask = torch.as_tensor(attention_mask).to(device=self._device, dtype=torch.long).contiguous()
            if mask.shape != ids.shape:
                raise ValueError("attention_mask shape must match input_ids")
        else:
            mask = self._infer_attention_mask(ids, gp)
        if B > self.pool_size:
            if self.takes_host_batches(B, S):
                # more rows than one pool of decode slots: continuous batching, same tokens row for row. The pool
                # admits prompts from host memory as slots free up (its entry point takes host buffers).
                out_np, _ = self.generate_stream(ids.cpu().numpy(), None if mask is None else mask.cpu().numpy(), _gen_params=gp)
                return torch.from_numpy(out_np).to(self._device)
            # prompts the slot pool cannot take (it needs the packed tcgen05 encoder, S <= 512): static batches
            outs = [self._generate_static(ids[lo:lo + self.pool_size], None if mask is None else mask[lo:lo + self.pool_size], gp)
                    for lo in range(0, B, self.pool_size)]
            width = max(o.shape[1] for o in outs)
            pad = gp.pad_token_id if gp.pad_token_id >= 0 else self.generation_config.pad_token_id
            return torch.cat([torch.nn.functional.pad(o, (0, width - o.shape[1]), value=pad) for o in outs], dim=0)
        return self._generate_static(ids, mask, gp)

    def takes_host_batches(self, B: int, S: int) -> bool:
        """True when a [B, S] batch would go through the slot pool, whose entry point takes HOST buffers: a caller that
        still has the batch in host memory (predictor.py) hands it over as it is."""
        return B > self.pool_size and S <= _POOL_MAX_S and os.environ.get("B200T5_STREAM", "1") != "0"

    def _generate_pool_from_host(self, ids: torch.Tensor, attention_mask, gp) -> torch.Tensor:
        ids_np = np.ascontiguousarray(ids.numpy(), dtype=np.int64)
        if ids_np.size and (int(ids_np.min()) < 0 or int(ids_np.max()) >= self.config.vocab_size):
            raise IndexError("input_ids contain token ids outside [0, vocab_size)")
        if attention_mask is not None:
            mask_np = np.ascontiguousarray(torch.as_tensor(attention_mask).cpu().numpy(), dtype=np.int64)
            if mask_np.shape != ids_np.shape:
                raise ValueError("attention_mask shape must match input_ids")
        else:
            mask_np = self._infer_mask_np(ids_np, gp)
        out_np, _ = self.generate_stream(ids_np, mask_np, _gen_params=gp)
        return torch.from_numpy(out_np).to(self._device)

    def _infer_attention_mask(self, ids: torch.Tensor, gp) -> Optional[torch.Tensor]:
        """GenerationMixin._prepare_attention_mask_for_generation (transformers generation/utils.py): without an
        attention_mask the pad positions are masked when the pad token occurs in the inputs and is not the EOS
        token; otherwise every position is attended (None = all ones for the library)."""
        pad = gp.pad_token_id if gp.pad_token_id >= 0 else self.generation_config.pad_token_id
        eos = gp.eos_token_id if gp.eos_token_id >= 0 else self.generation_config.eos_token_id
        if pad is None or pad == eos:
            return None
        is_pad = ids == pad
        if not bool(is_pad.any().item()):
            return None
        return (~is_pad).to(torch.long).contiguous()

    def _generate_static(self, ids: torch.Tensor, mask: Optional[torch.Tensor], gp) -> torch.Tensor:
        B, S = ids.shape
        ids = ids.contiguous()
        mask = None if mask is None else mask.contiguous()
        T = gp.max_new_tokens
        with torch.cuda.device(self._index):
            out = torch.empty((B, T + 1), dtype=torch.long, device=self._device)
            lens = torch.empty((B,), dtype=torch.int32, device=self._device)
            stream = torch.cuda.current_stream(self._device)
            with self._gpu_lock:
                _chk(self, self._lib.b200t5_generate(self._h, _ptr(ids), _ptr(mask), B, S, C.byref(gp), _ptr(out),
                                                     _ptr(lens), C.c_void_p(stream.cuda_stream)), self._h)
                self.last_lengths = lens
                steps = int(lens.max().item())  # synchronises; HF returns exactly the steps it ran
        return out[:, : steps + 1]

    def _infer_mask_np(self, ids: np.ndarray, gp) -> Optional[np.ndarray]:
        pad = gp.pad_token_id if gp.pad_token_id >= 0 else self.generation_config.pad_token_id
        eos = gp.eos_token_id if gp.eos_token_id >= 0 else self.generation_config.eos_token_id
        if pad is None or pad == eos or not (ids == pad).any():
            return None
        return np.ascontiguousarray(ids != pad, dtype=np.int64)

    def generate_host(self, input_ids: np.ndarray, attention_mask: Optional[np.ndarray] = None, **kw):
        """numpy in / numpy out through b200t5_generate_host (the foreign-host entry point):
        H2D copy, generation, D2H copy and synchronisation all happen inside the library."""
        gp = self._gen_params(kw.get("max_new_tokens"), kw.get("max_length"), kw.get("min_new_tokens"),
                              kw.get("min_length"), kw.get("eos_token_id"), kw.get("pad_token_id"),
                              kw.get("decoder_start_token_id"), kw.get("poll_interval", 8))
        ids = np.ascontiguousarray(input_ids, dtype=np.int64)
        B, S = ids.shape
        mask = self._infer_mask_np(ids, gp) if attention_mask is None else np.ascontiguousarray(attention_mask, dtype=np.int64)
        out = np.empty((B, gp.max_new_tokens + 1), dtype=np.int64)
        lens = np.empty((B,), dtype=np.int32)
        with self._gpu_lock:
            _chk(self, self._lib.b200t5_generate_host(
                self._h, ids.ctypes.data_as(C.c_void_p), None if mask is None else mask.ctypes.data_as(C.c_void_p), B, S,
                C.byref(gp), out.ctypes.data_as(C.c_void_p), lens.ctypes.data_as(C.c_void_p)), self._h)
        return out[:, : int(lens.max()) + 1], lens

    def generate_stream(self, input_ids: np.ndarray, attention_mask: Optional[np.ndarray] = None, *, pool: Optional[int] = None,
                        admit_min: int = 0, _gen_params=None, **kw):
        """N prompts through a pool of decode slots (b200t5_generate_stream): a slot whose row has finished is
        refilled with the next prompt, so short answers do not wait for the slowest row of a fixed batch as they do
        when BatchPredictor hands `generate` one batch at a time (NB:908-913 -> JOB/predictor.py:102). Returns
        (int64 [N, 1+T'], int32 lengths [N]) with the rows in input order; every row equals what `generate` returns
        for that prompt."""
        gp = _gen_params or self._gen_params(kw.get("max_new_tokens"), kw.get("max_length"), kw.get("min_new_tokens"),
                                             kw.get("min_length"), kw.get("eos_token_id"), kw.get("pad_token_id"),
                                             kw.get("decoder_start_token_id"), kw.get("poll_interval", 8))
        ids = np.ascontiguousarray(input_ids, dtype=np.int64)
        if ids.ndim != 2:
            raise ValueError(f"input_ids must be [batch, seq], got {ids.shape}")
        N, S = ids.shape
        if ids.size and (int(ids.min()) < 0 or int(ids.max()) >= self.config.vocab_size):
            raise IndexError("input_ids contain token ids outside [0, vocab_size)")
        mask = self._infer_mask_np(ids, gp) if attention_mask is None else np.ascontiguousarray(attention_mask, dtype=np.int64)
        if mask is not None and mask.shape != ids.shape:
            raise ValueError("attention_mask shape must match input_ids")
        out = np.empty((N, gp.max_new_tokens + 1), dtype=np.int64)
        lens = np.empty((N,), dtype=np.int32)
        with self._gpu_lock:
            _chk(self, self._lib.b200t5_generate_stream(
                self._h, ids.ctypes.data_as(C.c_void_p), None if mask is None else mask.ctypes.data_as(C.c_void_p), N, S,
                C.byref(gp), int(pool or self.pool_slots), int(admit_min), out.ctypes.data_as(C.c_void_p),
                lens.ctypes.data_as(C.c_void_p)), self._h)
        self.last_lengths = torch.from_numpy(lens)
        return out[:, : int(lens.max()) + 1], lens

    def stats(self) -> Dict[str, float]:
        s = _lib.Stats()
        _chk(self, self._lib.b200t5_get_stats(self._h, C.byref(s)), self._h)
        return {k: getattr(s, k) for k, _ in _lib.Stats._fields_}

    def bench_cross_attention(self, reps: int = 5, rows_per_launch: int = 0) -> Dict[str, float]:
        """Average launch time of the cross-attention decode kernel ALONE on the last call's KV arena, in launches of
        `rows_per_launch` rows (0 = the whole batch). A microbenchmark; `xattn_profile` is the in-situ figure."""
        ms, nbytes = C.c_float(), C.c_double()
        with torch.cuda.device(self._index):
            stream = torch.cuda.current_stream(self._device)
            _chk(self, self._lib.b200t5_bench_cross_attn(self._h, reps, int(rows_per_launch), C.byref(ms), C.byref(nbytes),
                                                         C.c_void_p(stream.cuda_stream)), self._h)
        return {"ms_per_launch": ms.value, "bytes_per_launch": nbytes.value}

    def set_option(self, name: str, value: int) -> None:
        """Runtime knob of the library (include/b200t5.h: b200t5_set_option); drops the execution plan."""
        _chk(self, self._lib.b200t5_set_option(self._h, name.encode(), int(value)), self._h)

    def xattn_profile(self) -> Dict[str, float]:
        """In-situ timing of the cross-attention launches of the step graph since set_option("profile_xattn", 1): per
        launch, and per layer as the union of the row-chains' (possibly overlapping) launches."""
        us, n, nbytes, busy, lbytes = C.c_double(), C.c_int64(), C.c_double(), C.c_double(), C.c_double()
        _chk(self, self._lib.b200t5_get_xattn_profile(self._h, C.byref(us), C.byref(n), C.byref(nbytes), C.byref(busy),
                                                      C.byref(lbytes)), self._h)
        return {"us_per_launch": us.value, "launches": int(n.value), "bytes_per_launch": nbytes.value,
                "busy_us_per_layer": busy.value, "bytes_per_layer": lbytes.value}

    # ------------------------------------------------------------------ parity hooks (tests)
    @torch.no_grad()
    def encode(self, input_ids, attention_mask=None) -> torch.Tensor:
        ids = torch.as_tensor(input_ids).to(self._device, torch.long).contiguous()
        mask = None if attention_mask is None else torch.as_tensor(attention_mask).to(self._device, torch.long).contiguous()
        B, S = ids.shape
        out = torch.empty((B, S, self.config.d_model), dtype=self._dtype, device=self._device)
        with torch.cuda.device(self._index):
            stream = torch.cuda.current_stream(self._device)
            _chk(self, self._lib.b200t5_encode(self._h, _ptr(ids), _ptr(mask), B, S, _ptr(out),
                                               C.c_void_p(stream.cuda_stream)), self._h)
            torch.cuda.synchronize(self._device)
        return out

    @torch.no_grad()
    def decode_logits(self, input_ids, attention_mask, decoder_input_ids) -> torch.Tensor:
        ids = torch.as_tensor(input_ids).to(self._device, torch.long).contiguous()
        mask = None if attention_mask is None else torch.as_tensor(attention_mask).to(self._device, torch.long).contiguous()
        dec = torch.as_tensor(decoder_input_ids).to(self._device, torch.long).contiguous()
        B, S = ids.shape
        T = dec.shape[1]
        out = torch.empty((B, T, self.config.vocab_size), dtype=torch.float32, device=self._device)
        with torch.cuda.device(self._index):
            stream = torch.cuda.current_stream(self._device)
            _chk(self, self._lib.b200t5_decode_logits(self._h, _ptr(ids), _ptr(mask), B, S, _ptr(dec), T, _ptr(out),
                                                      C.c_void_p(stream.cuda_stream)), self._h)
        return out
