"""`ray.data` subset: an eager, block-based Dataset with pandas/numpy batch formats."""
from __future__ import annotations

from typing import Any, Callable, Dict, Iterator, List, Optional, Union

import numpy as np
import pandas as pd

Block = Dict[str, np.ndarray]  # column -> array whose first axis is the row axis


def _col_to_array(col: pd.Series) -> np.ndarray:
    if col.dtype == object and len(col) and isinstance(col.iloc[0], np.ndarray):
        return np.stack(col.to_list())
    return col.to_numpy()


def _to_block(batch: Any) -> Block:
    if isinstance(batch, pd.DataFrame):
        return {c: _col_to_array(batch[c]) for c in batch.columns}
    if isinstance(batch, dict):
        return {k: np.asarray(v) for k, v in batch.items()}
    if isinstance(batch, np.ndarray):
        return {"__value__": batch}
    raise TypeError(f"unsupported batch type {type(batch)}")


def _block_len(b: Block) -> int:
    return 0 if not b else len(next(iter(b.values())))


def _to_pandas(b: Block) -> pd.DataFrame:
    cols = {}
    for k, v in b.items():
        cols[k] = list(v) if v.ndim > 1 else v
    return pd.DataFrame(cols)


def _slice(b: Block, lo: int, hi: int) -> Block:
    return {k: v[lo:hi] for k, v in b.items()}


def _concat(blocks: List[Block]) -> Block:
    blocks = [b for b in blocks if _block_len(b)]
    if not blocks:
        return {}
    return {k: np.concatenate([b[k] for b in blocks], axis=0) for k in blocks[0]}


class Dataset:
    """Rows are kept in order; every transformation executes eagerly and preserves order, which is
    what `input_data_pd.join(prediction_pd, how="inner")` (notebook :934) relies on."""

    def __init__(self, blocks: List[Block]):
        self._blocks = [b for b in blocks if _block_len(b)]

    # ---- construction helpers
    @staticmethod
    def _from_block(b: Block, block_rows: int = 4096) -> "Dataset":
        n = _block_len(b)
        return Dataset([_slice(b, i, min(i + block_rows, n)) for i in range(0, n, block_rows)])

    # ---- inspection
    def count(self) -> int:
        return sum(_block_len(b) for b in self._blocks)

    def num_blocks(self) -> int:
        return len(self._blocks)

    def schema(self):
        if not self._blocks:
            return {}
        return {k: (v.dtype, v.shape[1:]) for k, v in self._blocks[0].items()}

    def columns(self) -> List[str]:
        return list(self._blocks[0]) if self._blocks else []

    def take(self, n: int = 20) -> List[dict]:
        out = []
        for b in self._blocks:
            for i in range(_block_len(b)):
                if len(out) >= n:
                    return out
                out.append({k: v[i] for k, v in b.items()})
        return out

    def take_all(self) -> List[dict]:
        return self.take(self.count())

    def show(self, n: int = 20) -> None:
        for row in self.take(n):
            print(row)

    def to_pandas(self, limit: Optional[int] = None) -> pd.DataFrame:
        df = _to_pandas(_concat(self._blocks)) if self._blocks else pd.DataFrame()
        return df if limit is None else df.head(limit)

    def to_numpy(self) -> Block:
        return _concat(self._blocks)

    def fully_executed(self) -> "Dataset":
        return self

    materialize = fully_executed

    def __repr__(self):
        return f"Dataset(num_blocks={len(self._blocks)}, num_rows={self.count()}, schema={self.schema()})"

    # ---- transformations
    def limit(self, n: int) -> "Dataset":
        out, left = [], n
        for b in self._blocks:
            if left <= 0:
                break
            k = min(left, _block_len(b))
            out.append(_slice(b, 0, k))
            left -= k
        return Dataset(out)

    def repartition(self, num_blocks: int) -> "Dataset":
        whole = _concat(self._blocks)
        n = _block_len(whole)
        edges = np.linspace(0, n, num_blocks + 1).astype(int)
        return Dataset([_slice(whole, edges[i], edges[i + 1]) for i in range(num_blocks)])

    def iter_batches(self, batch_size: Optional[int] = 256, batch_format: str = "pandas") -> Iterator[Any]:
        whole = _concat(self._blocks)
        n = _block_len(whole)
        step = n if batch_size is None else batch_size
        for lo in range(0, n, max(step, 1)):
            b = _slice(whole, lo, min(lo + step, n))
            yield _to_pandas(b) if batch_format == "pandas" else b

    def map_batches(self, fn: Union[Callable, type], *, batch_size: Optional[int] = 4096, batch_format: str = "pandas",
                    compute: Any = None, fn_constructor_args=(), fn_constructor_kwargs=None, fn_kwargs=None,
                    num_gpus: Optional[float] = None, **ray_remote_args) -> "Dataset":
        """`fn` is a function or a callable class (constructed once, like an actor)."""
        if isinstance(fn, type):
            fn = fn(*fn_constructor_args, **(fn_constructor_kwargs or {}))
        kw = fn_kwargs or {}
        out = [_to_block(fn(batch, **kw)) for batch in self.iter_batches(batch_size, batch_format)]
        return Dataset(out)


class BatchMapper:
    """`ray.data.preprocessors.BatchMapper(fn, batch_format="pandas", batch_size=4096)` (notebook :296)."""

    def __init__(self, fn: Callable, batch_format: str = "pandas", batch_size: Optional[int] = 4096):
        self.fn = fn
        self.batch_format = batch_format
        self.batch_size = batch_size

    def fit(self, ds: Dataset) -> "BatchMapper":
        return self

    def transform(self, ds: Dataset) -> Dataset:
        return ds.map_batches(self.fn, batch_size=self.batch_size, batch_format=self.batch_format)

    def fit_transform(self, ds: Dataset) -> Dataset:
        return self.transform(ds)

    def transform_batch(self, batch: Any) -> Any:
        blk = _to_block(batch)
        arg = _to_pandas(blk) if self.batch_format == "pandas" else blk
        return self.fn(arg)


# ---- constructors (ray.data.from_*)
def from_pandas(dfs: Union[pd.DataFrame, List[pd.DataFrame]]) -> Dataset:
    dfs = [dfs] if isinstance(dfs, pd.DataFrame) else list(dfs)
    return Dataset([_to_block(df) for df in dfs])


def from_items(items: List[Any]) -> Dataset:
    if items and isinstance(items[0], dict):
        return from_pandas(pd.DataFrame(items))
    return from_pandas(pd.DataFrame({"item": items}))


def from_numpy(arr: Union[np.ndarray, Dict[str, np.ndarray]]) -> Dataset:
    return Dataset._from_block(_to_block(arr))


def from_huggingface(dataset: Any):
    """`datasets.Dataset` -> Dataset ; `datasets.DatasetDict` -> dict of Datasets (notebook :184)."""
    if isinstance(dataset, dict) and dataset and all(isinstance(v, (list, np.ndarray)) for v in dataset.values()):
        # plain column dict (the synthetic Alpaca rows)
        return Dataset._from_block({k: np.asarray(v, dtype=object if len(v) and isinstance(v[0], str) else None)
                                    for k, v in dataset.items()})
    if hasattr(dataset, "to_pandas"):
        return Dataset._from_block(_to_block(dataset.to_pandas()))
    if hasattr(dataset, "keys"):
        return {k: from_huggingface(dataset[k]) for k in dataset.keys()}
    raise TypeError(f"cannot build a Dataset from {type(dataset)}")
