"""Minimal Ray-compatible surface for the one path this repository accelerates.

Ray (pinned ray==2.3.1 in the reference's requirements.txt:134) is not installable offline, so
the notebook's calls are served by this shim: `ray.init/shutdown`, `ray.data.from_huggingface /
from_items / from_pandas`, `Dataset.limit / map_batches / to_pandas / take / show / count /
schema`, `ray.data.preprocessors.BatchMapper`, `ray.train.predictor.Predictor`,
`ray.train.batch_predictor.BatchPredictor`, and a `HuggingFaceCheckpoint` with
`get_model / get_tokenizer / get_preprocessor` (call sites:
Anyscale_job/flan-t5-batch-inference.py:28-38,119-138 and notebook :184-216,:296,:875-934).
Only the batch-inference path is implemented: training, tuning and serving entry points raise.

`install()` registers the shim under the module name `ray` ONLY when real Ray is absent.
"""
from __future__ import annotations

import importlib.util
import sys
import types

from . import data, train  # noqa: F401
from .data import Dataset  # noqa: F401

_initialized = False


def init(*args, **kwargs):
    global _initialized
    _initialized = True
    return {"shim": True, "address": None}


def shutdown():
    global _initialized
    _initialized = False


def is_initialized() -> bool:
    return _initialized


def _unsupported(name):
    def f(*a, **k):
        raise NotImplementedError(f"ray.{name} is outside the batch-inference path this repository implements")

    return f


remote = _unsupported("remote")
get = _unsupported("get")
put = _unsupported("put")


def install(force: bool = False) -> bool:
    """Make `import ray` resolve to this shim if Ray is not installed. Returns True if installed."""
    if not force and "ray" in sys.modules and not getattr(sys.modules["ray"], "__b200_shim__", False):
        return False
    if not force and importlib.util.find_spec("ray") is not None and "ray" not in sys.modules:
        return False
    me = sys.modules[__name__]
    me.__b200_shim__ = True
    sys.modules["ray"] = me
    sys.modules["ray.data"] = data
    pre = types.ModuleType("ray.data.preprocessors")
    pre.BatchMapper = data.BatchMapper
    sys.modules["ray.data.preprocessors"] = pre
    data.preprocessors = pre
    sys.modules["ray.train"] = train
    for sub, names in {
        "ray.train.predictor": ["Predictor"],
        "ray.train.batch_predictor": ["BatchPredictor"],
        "ray.train.huggingface": ["HuggingFaceCheckpoint", "HuggingFaceTrainer"],
        "ray.air": ["Checkpoint"],
        "ray.air.checkpoint": ["Checkpoint"],
        "ray.air.config": ["RunConfig", "ScalingConfig", "CheckpointConfig"],
    }.items():
        m = types.ModuleType(sub)
        for n in names:
            setattr(m, n, getattr(train, n))
        sys.modules[sub] = m
    train.predictor = sys.modules["ray.train.predictor"]
    train.batch_predictor = sys.modules["ray.train.batch_predictor"]
    train.huggingface = sys.modules["ray.train.huggingface"]
    me.air = sys.modules["ray.air"]
    me.air.config = sys.modules["ray.air.config"]
    return True
