"""One scoring process per GPU (what Ray's ActorPoolStrategy + num_gpus=1 gives the reference:
`predictor.predict(..., num_gpus_per_worker=int(use_gpu), batch_size=256)`,
NLP_workloads/Anyscale_job/flan-t5-batch-inference.py:129-134; notebook :908-913).

Blocks are dealt round-robin to the workers (`parallel.shard_block_indices`: static sharding, no collective, no
GPU<->GPU traffic); results come back tagged with their block index and are re-assembled in input order
(`parallel.restore_order`). Workers are spawned (not forked) with CUDA_VISIBLE_DEVICES pinned before CUDA
initialises, and they STAY ALIVE between `predict` calls: the model load and the CUDA-graph capture are paid once
per pool, not once per call.

The CPU stage runs inside the workers: when the checkpoint carries a preprocessor, a worker receives the RAW rows of
its blocks (strings - small messages) and tokenises block i+1 on a producer thread while the GPU generates block i,
instead of the driver process tokenising the whole dataset up front. Tokenised arrays therefore never cross a process
boundary; what crosses it is strings in, strings out. A worker also keeps two blocks in flight on two threads, so the
host tail of block i (detokenise, DataFrame) overlaps the generation of block i+1 (`train._overlap_tail`).

Failure handling: every wait on a worker polls its liveness, so a worker killed by a native failure (CUDA abort,
segfault, OOM kill - which post no Python exception) surfaces as a RuntimeError with its exit code instead of a hang.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import pickle
import queue
import traceback
from typing import Any, Dict, List, Optional

import cloudpickle

from ..parallel import restore_order, shard_block_indices

_POLL_S = 1.0


def _worker_main(visible: str, payload: bytes, task_q, result_q) -> None:
    os.environ["CUDA_VISIBLE_DEVICES"] = visible
    try:
        from .train import _model_batch, _overlap_ok, _overlap_tail, _prefetch, _ScoringWorker

        checkpoint, predictor_cls, kwargs, override_prep = cloudpickle.loads(payload)
        worker = _ScoringWorker(checkpoint, predictor_cls, kwargs, override_prep, dedicated_process=True)
        result_q.put(("ready", visible, None))
        while True:
            task = task_q.get()
            if task is None:
                break
            call_id, blob = task
            items, feature_columns, keep_columns, predict_kwargs, prep = cloudpickle.loads(blob)
            if prep is not None:  # this worker's own CPU stage, one block ahead of its GPU stage
                stream = _prefetch(items, lambda it: (it[0], _model_batch(prep.transform_batch(it[1]))))
            else:
                stream = iter(items)
            # (two blocks in flight: the detokenise tail of block i overlaps the generation of block i+1)
            scored = _overlap_tail(lambda it: (it[0], worker(it[1], feature_columns, keep_columns, predict_kwargs)), stream,
                                   _overlap_ok(worker))
            for idx, out in scored:
                result_q.put(("ok", call_id, idx, pickle.dumps(out, protocol=pickle.HIGHEST_PROTOCOL)))
            result_q.put(("done", call_id, visible, None))
    except BaseException:  # noqa: BLE001 - surface the failure to the driver instead of hanging it
        result_q.put(("error", visible, traceback.format_exc()))


def _visible_devices(n_workers: int, gpus_per_worker: int) -> List[str]:
    """CUDA_VISIBLE_DEVICES value of every worker: entries of the parent's own list (indices, UUIDs or MIG ids, passed
    through as strings), `gpus_per_worker` of them per worker."""
    env = os.environ.get("CUDA_VISIBLE_DEVICES")
    if env:
        ids = [x.strip() for x in env.split(",") if x.strip()]
    else:
        import torch

        ids = [str(i) for i in range(torch.cuda.device_count() if torch.cuda.is_available() else n_workers)]
    g = max(int(gpus_per_worker), 1)
    groups = [ids[i:i + g] for i in range(0, len(ids) - g + 1, g)] or [ids[:g] or ["0"]]
    return [",".join(groups[i % len(groups)]) for i in range(n_workers)]


class GpuWorkerPool:
    def __init__(self, n_workers: int, checkpoint: Any, predictor_cls: Any, kwargs: Dict[str, Any], override_prep: bool,
                 gpus_per_worker: int = 1, start_timeout_s: float = 900.0):
        ctx = mp.get_context("spawn")
        self.visible = _visible_devices(n_workers, gpus_per_worker)
        payload = cloudpickle.dumps((checkpoint, predictor_cls, kwargs, override_prep))
        self.result_q = ctx.Queue()
        self.task_qs = [ctx.Queue() for _ in range(n_workers)]
        self.procs = [ctx.Process(target=_worker_main, args=(self.visible[i], payload, self.task_qs[i], self.result_q), daemon=True)
                      for i in range(n_workers)]
        self._calls = 0
        self.closed = False
        for p in self.procs:
            p.start()
        ready = 0
        waited = 0.0
        while ready < n_workers:
            msg = self._get(waited_s=waited, limit_s=start_timeout_s, what="start")
            if msg is None:
                waited += _POLL_S
                continue
            if msg[0] == "error":
                self.close()
                raise RuntimeError(f"scoring worker on GPU {msg[1]} failed to start:\n{msg[2]}")
            ready += 1

    @property
    def n_workers(self) -> int:
        return len(self.procs)

    def _get(self, waited_s: float, limit_s: Optional[float], what: str):
        """One poll of the result queue; raises if a worker died without posting a result (native failure)."""
        try:
            return self.result_q.get(timeout=_POLL_S)
        except queue.Empty:
            dead = [(self.visible[i], p.exitcode) for i, p in enumerate(self.procs) if not p.is_alive()]
            if dead:
                self.close()
                raise RuntimeError(f"scoring worker(s) died during {what} without reporting an exception "
                                   f"(GPU, exit code): {dead} - a native failure (CUDA abort / segfault / OOM kill)") from None
            if limit_s is not None and waited_s + _POLL_S >= limit_s:
                self.close()
                raise TimeoutError(f"scoring workers did not finish {what} within {limit_s:.0f} s") from None
            return None

    def map_ordered(self, batches: List[Any], feature_columns, keep_columns, predict_kwargs, prep: Any = None,
                    timeout_s: Optional[float] = None) -> List[Any]:
        """Score `batches` (block i on worker i mod N); returns the outputs in input order. `prep`: the AIR
        preprocessor each worker applies to its own raw blocks (None: the blocks are model inputs already)."""
        if self.closed:
            raise RuntimeError("the worker pool has been shut down")
        n = self.n_workers
        self._calls += 1
        call_id = self._calls
        for r in range(n):
            items = [(i, batches[i]) for i in shard_block_indices(len(batches), r, n)]
            self.task_qs[r].put((call_id, cloudpickle.dumps((items, feature_columns, keep_columns, predict_kwargs, prep))))
        per_rank: List[Dict[int, Any]] = [dict() for _ in range(n)]
        done, waited = 0, 0.0
        while done < n:
            msg = self._get(waited_s=waited, limit_s=timeout_s, what="predict")
            if msg is None:
                waited += _POLL_S
                continue
            if msg[0] == "error":
                self.close()
                raise RuntimeError(f"scoring worker on GPU {msg[1]} failed:\n{msg[2]}")
            if msg[1] != call_id:
                continue  # stale message of an aborted call
            if msg[0] == "ok":
                per_rank[msg[2] % n][msg[2]] = pickle.loads(msg[3])
            elif msg[0] == "done":
                done += 1
        return restore_order([[d[i] for i in shard_block_indices(len(batches), r, n)] for r, d in enumerate(per_rank)], len(batches))

    def close(self) -> None:
        if self.closed:
            return
        self.closed = True
        for q in self.task_qs:
            try:
                q.put(None)
            except Exception:  # noqa: BLE001
                pass
        for p in self.procs:
            p.join(timeout=10)
            if p.is_alive():
                p.terminate()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
