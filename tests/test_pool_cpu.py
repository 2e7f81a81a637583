"""CPU-only: the one-process-per-GPU scoring pool (rayshim/pool.py) with stand-in predictors - block order, in-worker
CPU stage, persistence across predict() calls, and loud failure when a worker dies natively."""
import os

import pandas as pd
import pytest

from anyscale_workshop_nyc_2023_b200 import rayshim
from anyscale_workshop_nyc_2023_b200.rayshim.pool import GpuWorkerPool, _visible_devices
from anyscale_workshop_nyc_2023_b200.rayshim.train import Predictor


class EchoPredictor(Predictor):
    """Reports which process scored each row; dies without a Python exception when asked to."""

    @classmethod
    def from_checkpoint(cls, checkpoint, use_gpu=False, **kw):
        return cls(preprocessor=checkpoint.get_preprocessor())

    def _predict_pandas(self, data, **kw):
        if (data["text"] == "die").any():
            os._exit(77)  # what a CUDA abort / segfault looks like from the driver: no exception, no message
        if (data["text"] == "raise").any():
            raise ValueError("bad row")
        return pd.DataFrame({"generated_output": [f"{t}:{kw.get('suffix', '')}" for t in data["text"]],
                             "pid": [os.getpid()] * len(data), "gpu": [os.environ.get("CUDA_VISIBLE_DEVICES")] * len(data)})


class Ckpt:
    def __init__(self, prep=None):
        self._prep = prep

    def get_preprocessor(self):
        return self._prep


def _upper(batch):
    return pd.DataFrame({"text": batch["text"].str.upper()})


def _blocks(n_blocks, rows=3):
    return [pd.DataFrame({"text": [f"b{i}r{j}" for j in range(rows)]}) for i in range(n_blocks)]


def test_visible_devices_are_passed_through_as_strings(monkeypatch):
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "GPU-aaaa,GPU-bbbb,MIG-cccc,3")
    assert _visible_devices(3, 1) == ["GPU-aaaa", "GPU-bbbb", "MIG-cccc"]
    assert _visible_devices(2, 2) == ["GPU-aaaa,GPU-bbbb", "MIG-cccc,3"]
    assert _visible_devices(6, 1)[4] == "GPU-aaaa"  # more workers than devices wrap around


@pytest.mark.timeout(300)
def test_pool_order_persistence_and_worker_side_cpu_stage():
    with GpuWorkerPool(2, Ckpt(), EchoPredictor, {}, True) as pool:
        outs = pool.map_ordered(_blocks(7), None, None, {"suffix": "x"})
        assert [o["generated_output"].tolist() for o in outs] == [[f"b{i}r{j}:x" for j in range(3)] for i in range(7)]
        pids = [o["pid"][0] for o in outs]
        assert pids[0] == pids[2] == pids[4] == pids[6] and pids[1] == pids[3] == pids[5] and pids[0] != pids[1]
        assert {o["gpu"][0] for o in outs} == {"0", "1"} or len({o["gpu"][0] for o in outs}) == 2
        # second call: the same processes (model load and graph capture are paid once per pool, not per call),
        # each tokenising its own raw blocks
        prep = rayshim.data.BatchMapper(_upper, batch_format="pandas")
        again = pool.map_ordered(_blocks(5), None, None, {}, prep=prep)
        assert [o["generated_output"][0] for o in again] == [f"B{i}R0:" for i in range(5)]
        assert {o["pid"][0] for o in again} == set(pids)
        # a Python exception in a worker surfaces with its traceback
        bad = _blocks(3)
        bad[1].loc[0, "text"] = "raise"
        with pytest.raises(RuntimeError, match="bad row"):
            pool.map_ordered(bad, None, None, {})


@pytest.mark.timeout(300)
def test_native_death_of_a_worker_is_reported_not_hung():
    pool = GpuWorkerPool(2, Ckpt(), EchoPredictor, {}, True)
    blocks = _blocks(4)
    blocks[3].loc[1, "text"] = "die"
    with pytest.raises(RuntimeError, match=r"died .* without reporting an exception .*77"):
        pool.map_ordered(blocks, None, None, {})
    assert pool.closed
    with pytest.raises(RuntimeError, match="shut down"):
        pool.map_ordered(_blocks(1), None, None, {})
