"""One scoring process per GPU (what Ray's ActorPoolStrategy + num_gpus=1 gives the reference:
`predictor.predict(..., num_gpus_per_worker=int(use_gpu), batch_size=256)`,
NLP_workloads/Anyscale_job/flan-t5-batch-inference.py:129-134; notebook :908-913).

Batches are dealt round-robin to the workers (static sharding, no collective, no GPU<->GPU
traffic); results come back tagged with their index and are re-assembled in input order.
Workers are spawned (not forked) with CUDA_VISIBLE_DEVICES pinned before CUDA initialises.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import traceback
from typing import Any, Dict, List

import cloudpickle


def _worker_main(gpu_index: int, payload: bytes, task_q, result_q) -> None:
    os.environ["CUDA_VISIBLE_DEVICES"] = str(gpu_index)
    try:
        from .train import _ScoringWorker

        checkpoint, predictor_cls, kwargs, override_prep = cloudpickle.loads(payload)
        worker = _ScoringWorker(checkpoint, predictor_cls, kwargs, override_prep)
        result_q.put(("ready", gpu_index, None))
        while True:
            task = task_q.get()
            if task is None:
                break
            idx, blob = task
            batch, feature_columns, keep_columns, predict_kwargs = cloudpickle.loads(blob)
            out = worker(batch, feature_columns, keep_columns, predict_kwargs)
            result_q.put(("ok", idx, cloudpickle.dumps(out)))
    except Exception:  # surface the failure to the driver instead of hanging it
        result_q.put(("error", gpu_index, traceback.format_exc()))


class GpuWorkerPool:
    def __init__(self, n_workers: int, checkpoint: Any, predictor_cls: Any, kwargs: Dict[str, Any], override_prep: bool):
        ctx = mp.get_context("spawn")
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        gpu_ids = [int(x) for x in visible.split(",")] if visible else list(range(n_workers))
        payload = cloudpickle.dumps((checkpoint, predictor_cls, kwargs, override_prep))
        self.result_q = ctx.Queue()
        self.task_qs = [ctx.Queue() for _ in range(n_workers)]
        self.procs = [ctx.Process(target=_worker_main, args=(gpu_ids[i % len(gpu_ids)], payload, self.task_qs[i], self.result_q), daemon=True)
                      for i in range(n_workers)]
        for p in self.procs:
            p.start()
        for _ in self.procs:
            kind, who, info = self.result_q.get()
            if kind == "error":
                self.close()
                raise RuntimeError(f"scoring worker on GPU {who} failed to start:\n{info}")

    def map_ordered(self, batches: List[Any], feature_columns, keep_columns, predict_kwargs) -> List[Any]:
        n = len(self.procs)
        for i, b in enumerate(batches):
            self.task_qs[i % n].put((i, cloudpickle.dumps((b, feature_columns, keep_columns, predict_kwargs))))
        outs: List[Any] = [None] * len(batches)
        for _ in batches:
            kind, idx, blob = self.result_q.get()
            if kind == "error":
                self.close()
                raise RuntimeError(f"scoring worker on GPU {idx} failed:\n{blob}")
            outs[idx] = cloudpickle.loads(blob)
        return outs

    def close(self) -> None:
        for q in self.task_qs:
            try:
                q.put(None)
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=10)
            if p.is_alive():
                p.terminate()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
