"""`ray.train` subset: Predictor, BatchPredictor, HuggingFaceCheckpoint (batch inference only).

Semantics follow Ray AIR 2.3.1 as the reference uses them (the Ray source is
not available offline, so the behaviours below are the documented ones the notebook relies on):
  * `BatchPredictor.from_checkpoint(checkpoint, predictor_cls, **predictor_kwargs)` stores kwargs;
    `predict(ds, batch_size=..., num_gpus_per_worker=..., **predict_kwargs)` builds, in every
    scoring worker, `predictor_cls.from_checkpoint(checkpoint, **predictor_kwargs)` and calls
    `predictor.predict(batch, **predict_kwargs)` per `batch_size` rows; output row order == input
    row order; `use_gpu=True` is injected when GPUs are requested and the predictor accepts it;
  * the checkpoint's preprocessor runs as its own CPU stage before a GPU scoring stage;
  * `Predictor.predict` hands `_predict_numpy` a dict of numpy columns and returns a DataFrame.
One scoring worker = one process = one GPU (dataset blocks sharded round-robin, no collective).
"""
from __future__ import annotations

import inspect
import os
from pathlib import Path
from typing import Any, Dict, List, Optional, Type

import numpy as np
import pandas as pd

from .data import BatchMapper, Dataset, _concat, _to_block, _to_pandas


class Predictor:
    """Base class of AIR predictors (`ray.train.predictor.Predictor`)."""

    def __init__(self, preprocessor: Optional[Any] = None):
        self._preprocessor = preprocessor

    @classmethod
    def from_checkpoint(cls, checkpoint: Any, **kwargs) -> "Predictor":
        raise NotImplementedError

    def get_preprocessor(self):
        return self._preprocessor

    def set_preprocessor(self, preprocessor) -> None:
        self._preprocessor = preprocessor

    def predict(self, data: Any, **kwargs) -> pd.DataFrame:
        if not isinstance(data, (pd.DataFrame, dict, np.ndarray)):
            raise TypeError(f"unsupported batch type {type(data)}; expected DataFrame, dict or ndarray")
        if self._preprocessor is not None:
            data = self._preprocessor.transform_batch(data)
        has_np = type(self)._predict_numpy is not Predictor._predict_numpy
        has_pd = type(self)._predict_pandas is not Predictor._predict_pandas
        if has_np and not (has_pd and isinstance(data, pd.DataFrame)):
            out = self._predict_numpy(_to_block(data) if not isinstance(data, np.ndarray) else data, **kwargs)
        elif has_pd:
            out = self._predict_pandas(data if isinstance(data, pd.DataFrame) else _to_pandas(_to_block(data)), **kwargs)
        else:
            raise NotImplementedError("a Predictor must implement _predict_numpy or _predict_pandas")
        if isinstance(out, dict):
            out = _to_pandas(_to_block(out))
        elif isinstance(out, np.ndarray):
            out = pd.DataFrame({"predictions": list(out)})
        return out

    def _predict_numpy(self, data, **kwargs):
        raise NotImplementedError

    def _predict_pandas(self, data, **kwargs):
        raise NotImplementedError


class HuggingFaceCheckpoint:
    """Directory in `save_pretrained` format + an optional AIR preprocessor."""

    def __init__(self, path: str, preprocessor: Optional[Any] = None):
        self.path = str(path)
        self._preprocessor = preprocessor

    @classmethod
    def from_directory(cls, path: str) -> "HuggingFaceCheckpoint":
        return cls(path)

    @classmethod
    def from_model(cls, model: Any = None, tokenizer: Any = None, *, path: str, preprocessor: Any = None):
        Path(path).mkdir(parents=True, exist_ok=True)
        if model is not None:
            model.save_pretrained(path)
        if tokenizer is not None:
            tokenizer.save_pretrained(path)
        return cls(path, preprocessor)

    def to_directory(self, path: Optional[str] = None) -> str:
        return self.path

    def get_preprocessor(self):
        return self._preprocessor

    def set_preprocessor(self, preprocessor) -> None:
        self._preprocessor = preprocessor

    def get_model(self, model: Any, **pretrained_model_kwargs) -> Any:
        """Class -> `model.from_pretrained(dir, **kw)` (predictor.py:68); instance -> load the
        checkpoint's weights into it (notebook :554)."""
        if isinstance(model, type):
            return model.from_pretrained(self.path, **pretrained_model_kwargs)
        import torch

        from ..synth import read_safetensors

        st = Path(self.path) / "model.safetensors"
        if st.exists():
            sd = {}
            for name, (dt, _shape, arr) in read_safetensors(st).items():
                a = np.array(arr)
                sd[name] = torch.from_numpy(a.view(np.int16)).view(torch.bfloat16) if dt == "BF16" else torch.from_numpy(a)
        else:
            sd = torch.load(Path(self.path) / "pytorch_model.bin", map_location="cpu", weights_only=True)
        model.load_state_dict(sd, strict=False)
        return model

    def get_tokenizer(self, tokenizer: Type, **kwargs) -> Any:
        return tokenizer.from_pretrained(self.path, **kwargs)


Checkpoint = HuggingFaceCheckpoint


class _ScoringWorker:
    """What runs inside one scoring worker (Ray's ScoringWrapper actor)."""

    def __init__(self, checkpoint, predictor_cls, predictor_kwargs: Dict[str, Any], override_prep: bool,
                 dedicated_process: bool = False):
        self.predictor = predictor_cls.from_checkpoint(checkpoint, **predictor_kwargs)
        if override_prep:
            self.predictor.set_preprocessor(None)
        if dedicated_process or os.environ.get("B200T5_GC_FREEZE") == "1":
            # a long-lived scoring process (one per GPU, rayshim/pool.py): the objects created while loading (model,
            # tokenizer, imported modules) will never be garbage - keep the collector from re-walking them (a full
            # collection otherwise stalls a ~220 ms scoring call by tens of ms every so often). Not done to the caller's
            # own process unless asked for (B200T5_GC_FREEZE=1): it is a process-wide setting.
            import gc

            gc.collect()
            gc.freeze()

    def __call__(self, batch, feature_columns, keep_columns, predict_kwargs) -> pd.DataFrame:
        """`batch`: a pandas batch, or - what the worker-side CPU stage hands over - a block of numpy columns (the
        tokenised [N, 512] arrays go to `_predict_numpy` as they are: no DataFrame of 4096 row objects in between,
        which cost more host time per block than the GPU needs to score it)."""
        if isinstance(batch, dict):
            data = {k: batch[k] for k in feature_columns} if feature_columns else batch
            out = self.predictor.predict(data, **predict_kwargs)
            if keep_columns:
                out = out.copy()
                for c in keep_columns:
                    out[c] = batch[c] if batch[c].ndim == 1 else list(batch[c])
            return out
        data = batch[feature_columns] if feature_columns else batch
        out = self.predictor.predict(data, **predict_kwargs)
        if keep_columns:
            out = out.copy()
            for c in keep_columns:
                out[c] = batch[c].to_numpy() if batch[c].dtype != object else list(batch[c])
        return out


def _visible_gpus() -> int:
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _prefetch(items: List[Any], fn, depth: int = 2):
    """Yield fn(item) in order while a producer thread works `depth` items ahead (the consumer spends its time inside
    the GPU library with the GIL released, so the two overlap). Exceptions of the producer are re-raised here."""
    import queue
    import threading

    q: "queue.Queue" = queue.Queue(maxsize=depth)
    stop = threading.Event()

    def put(x) -> bool:
        while not stop.is_set():
            try:
                q.put(x, timeout=0.1)
                return True
            except queue.Full:
                continue
        return False

    def run():
        try:
            for it in items:
                if not put(("ok", fn(it))):
                    return
            put(("end", None))
        except BaseException as e:  # noqa: BLE001 - handed to the consumer
            put(("err", e))

    t = threading.Thread(target=run, name="b200t5-cpu-stage", daemon=True)
    t.start()
    try:
        while True:
            kind, val = q.get()
            if kind == "end":
                return
            if kind == "err":
                raise val
            yield val
    finally:
        stop.set()
        t.join(timeout=5)


def _model_batch(transformed: Any):
    """Output of a worker-side preprocessor as the scoring stage takes it: numpy columns stay numpy columns."""
    if isinstance(transformed, dict) and all(isinstance(v, np.ndarray) for v in transformed.values()):
        return transformed
    return _to_pandas(_to_block(transformed))


def _overlap_tail(fn, items, enabled: bool = True):
    """Yield fn(item) in input order with TWO calls in flight on two threads. The GPU part of a call is serialised by
    the model's own lock and runs with the GIL released (ctypes), so the host tail of block i (detokenise, DataFrame
    assembly: predictor.py:103-106 of the reference) overlaps the generation of block i+1. `enabled` False (a model
    that does not declare such a lock): plain sequential map."""
    if not enabled:
        for it in items:
            yield fn(it)
        return
    import collections
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(max_workers=2, thread_name_prefix="b200t5-score") as ex:
        pending: "collections.deque" = collections.deque()
        try:
            for it in items:
                pending.append(ex.submit(fn, it))
                if len(pending) == 2:
                    yield pending.popleft().result()
            while pending:
                yield pending.popleft().result()
        finally:
            for f in pending:
                f.cancel()


def _overlap_ok(worker: "_ScoringWorker") -> bool:
    return hasattr(getattr(worker.predictor, "model", None), "_gpu_lock")


class BatchPredictor:
    def __init__(self, checkpoint: Any, predictor_cls: Type[Predictor], **predictor_kwargs: Any):
        self._checkpoint = checkpoint
        self._predictor_cls = predictor_cls
        self._predictor_kwargs = predictor_kwargs
        self._override_preprocessor = None
        self._worker: Optional[_ScoringWorker] = None  # kept warm between predict() calls
        self._worker_key = None
        self._pool = None                              # likewise the one-process-per-GPU pool
        self._pool_key = None

    @classmethod
    def from_checkpoint(cls, checkpoint: Any, predictor_cls: Type[Predictor], **kwargs: Any) -> "BatchPredictor":
        return cls(checkpoint=checkpoint, predictor_cls=predictor_cls, **kwargs)

    def get_preprocessor(self):
        return self._override_preprocessor or self._checkpoint.get_preprocessor()

    def set_preprocessor(self, preprocessor) -> None:
        self._override_preprocessor = preprocessor

    def shutdown(self) -> None:
        """Stop the scoring processes (they otherwise live as long as this BatchPredictor, like Ray's actor pool
        lives for the duration of the job)."""
        if self._pool is not None:
            self._pool.close()
            self._pool = None

    def __del__(self):
        try:
            self.shutdown()
        except Exception:  # noqa: BLE001
            pass

    def predict(self, data: Dataset, *, feature_columns: Optional[List[str]] = None,
                keep_columns: Optional[List[str]] = None, batch_size: int = 4096, min_scoring_workers: int = 1,
                max_scoring_workers: Optional[int] = None, num_cpus_per_worker: Optional[int] = None,
                num_gpus_per_worker: Optional[int] = None, separate_gpu_stage: bool = True,
                ray_remote_args: Optional[Dict[str, Any]] = None, pipeline_cpu_stage: bool = True,
                **predict_kwargs) -> Dataset:
        num_gpus = int(num_gpus_per_worker or 0)
        kwargs = dict(self._predictor_kwargs)
        sig = inspect.signature(self._predictor_cls.from_checkpoint)
        if num_gpus > 0 and "use_gpu" in sig.parameters and "use_gpu" not in kwargs:
            kwargs["use_gpu"] = True
        batches = list(data.iter_batches(batch_size=batch_size, batch_format="pandas"))
        # how many scoring workers: decided BEFORE the CPU stage is placed (it runs inside them)
        n_workers = 1
        if num_gpus > 0:
            visible = _visible_gpus()
            cap = max_scoring_workers or visible
            n_workers = max(min(cap, visible // max(num_gpus, 1), len(batches)), 1)
            n_workers = max(n_workers, min(min_scoring_workers, max(visible // max(num_gpus, 1), 1)))
        prep = self.get_preprocessor()
        # With a GPU stage the preprocessor is a CPU stage of its own, as in AIR - but STREAMED inside each scoring
        # worker: a producer thread tokenises block i+1 while the GPU generates block i (the library call releases
        # the GIL). Tokenisation is row-wise, so the results do not depend on how the rows are blocked.
        # pipeline_cpu_stage=False: materialise the whole tokenised dataset first (what AIR 2.3 does).
        override_prep = prep is not None and num_gpus > 0 and separate_gpu_stage
        worker_prep = None
        if override_prep:
            if pipeline_cpu_stage:
                worker_prep = prep
            else:
                batches = list(prep.transform(data).iter_batches(batch_size=batch_size, batch_format="pandas"))
        key = (id(self._checkpoint), override_prep, repr(sorted(kwargs.items(), key=lambda kv: kv[0])))
        if n_workers <= 1:
            if self._worker is None or self._worker_key != key:
                self._worker = _ScoringWorker(self._checkpoint, self._predictor_cls, kwargs, override_prep)
                self._worker_key = key
            if worker_prep is not None:
                outs = list(_overlap_tail(lambda b: self._worker(b, feature_columns, keep_columns, predict_kwargs),
                                          _prefetch(batches, lambda raw: _model_batch(worker_prep.transform_batch(raw))),
                                          _overlap_ok(self._worker)))
            else:
                outs = list(_overlap_tail(lambda b: self._worker(b, feature_columns, keep_columns, predict_kwargs), batches,
                                          _overlap_ok(self._worker)))
        else:
            from .pool import GpuWorkerPool

            pool_key = key + (n_workers, num_gpus)
            if self._pool is None or self._pool.closed or self._pool_key != pool_key:
                self.shutdown()
                self._pool = GpuWorkerPool(n_workers, self._checkpoint, self._predictor_cls, kwargs, override_prep,
                                           gpus_per_worker=num_gpus)
                self._pool_key = pool_key
            outs = self._pool.map_ordered(batches, feature_columns, keep_columns, predict_kwargs, prep=worker_prep)
        return Dataset([_to_block(o) for o in outs])


# ---- entry points of the reference that are outside the batch-inference path
class _Config:
    def __init__(self, *args, **kwargs):
        self.__dict__.update(kwargs)


RunConfig = ScalingConfig = CheckpointConfig = _Config


class HuggingFaceTrainer:
    def __init__(self, *a, **k):
        raise NotImplementedError(
            "HuggingFaceTrainer (fine-tuning, BASELINE config 5) is outside the batch-inference hot path "
            "this repository implements; see DESIGN.md 'out of scope'")
