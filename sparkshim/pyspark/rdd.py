"""Lazy partitioned collections with the RDD surface SparkTorch-style trainers
touch: ``mapPartitions``, ``mapPartitionsWithIndex``,
``repartition``, ``getNumPartitions``, ``barrier()``, ``foreach``, ``collect``
and the private ``PipelinedRDD(prev, f(idx, it), preserves, isFromBarrier)``
constructor.

A source partition is a Python list whose elements may be :class:`RowBlock`
objects (columnar chunks standing for many rows).  Functions see blocks expanded
into rows unless they carry ``_accepts_blocks = True``.
"""
from __future__ import annotations

import itertools
import random
from typing import Any, Callable, Iterable, Iterator, List, Optional


class RowBlock:
    """A columnar chunk: ``columns`` names and equally long ``arrays``.
    2-D arrays are vector columns (one DenseVector per row when expanded)."""

    def __init__(self, columns: List[str], arrays: list):
        self.columns = list(columns)
        self.arrays = list(arrays)

    def __len__(self) -> int:
        return len(self.arrays[0]) if self.arrays else 0

    def column(self, name: str):
        return self.arrays[self.columns.index(name)]

    def take(self, idx) -> "RowBlock":
        return RowBlock(self.columns, [a[idx] for a in self.arrays])

    def rows(self) -> Iterator[Any]:
        from pyspark.ml.linalg import DenseVector
        from pyspark.sql.types import Row

        cols = self.columns
        # scalar columns are converted to Python values in one vectorised call, vector columns row by row
        per_col = []
        for a in self.arrays:
            if getattr(a, "ndim", 1) >= 2:
                per_col.append((True, a))
            elif hasattr(a, "tolist"):
                per_col.append((False, a.tolist()))
            else:
                per_col.append((False, a))
        for i in range(len(self)):
            yield Row._make(cols, [DenseVector(a[i]) if vec else a[i] for vec, a in per_col])


def expand_blocks(it: Iterable) -> Iterator[Any]:
    for e in it:
        if isinstance(e, RowBlock):
            yield from e.rows()
        else:
            yield e


class RDD:
    def __init__(self, ctx, partitions: Optional[List[list]] = None, prev: "Optional[RDD]" = None,
                 func: Optional[Callable[[int, Iterator], Iterable]] = None, barrier: bool = False):
        self.ctx = ctx
        self._parts = partitions
        self._prev = prev
        self._func = func
        self._barrier = barrier or (prev is not None and prev._barrier)
        self._cached: Optional[List[list]] = None

    # ---------------------------------------------------------------- lineage
    def getNumPartitions(self) -> int:
        return len(self._parts) if self._parts is not None else self._prev.getNumPartitions()

    def _compute(self, idx: int) -> Iterator:
        if self._cached is not None:
            return iter(self._cached[idx])
        if self._parts is not None:
            return iter(self._parts[idx])
        it = self._prev._compute(idx)
        if not getattr(self._func, "_accepts_blocks", False):
            it = expand_blocks(it)
        return iter(self._func(idx, it))

    def _prune(self, idx: int) -> "RDD":
        """Copy of the lineage that only carries partition ``idx``'s source
        data – this is what travels to an executor."""
        if self._cached is not None or self._parts is not None:
            src = self._cached if self._cached is not None else self._parts
            parts: List[Any] = [None] * len(src)
            parts[idx] = src[idx]
            return RDD(None, partitions=parts)
        r = RDD(None, prev=self._prev._prune(idx), func=self._func)
        return r

    # --------------------------------------------------------- transformations
    def mapPartitionsWithIndex(self, f, preservesPartitioning: bool = False) -> "RDD":
        return PipelinedRDD(self, f, preservesPartitioning)

    def mapPartitions(self, f, preservesPartitioning: bool = False) -> "RDD":
        def func(_idx, it):
            return f(it)

        if getattr(f, "_accepts_blocks", False):
            func._accepts_blocks = True  # type: ignore[attr-defined]
        return PipelinedRDD(self, func, preservesPartitioning)

    def map(self, f) -> "RDD":
        return PipelinedRDD(self, lambda _i, it: (f(x) for x in it))

    def filter(self, f) -> "RDD":
        return PipelinedRDD(self, lambda _i, it: (x for x in it if f(x)))

    def barrier(self) -> "RDDBarrier":
        return RDDBarrier(self)

    def _materialize(self) -> List[list]:
        return [list(self._compute(i)) for i in range(self.getNumPartitions())]

    def repartition(self, numPartitions: int) -> "RDD":
        """Round-robin redistribution (a shuffle); blocks are split, not expanded."""
        parts = self._materialize()
        out: List[list] = [[] for _ in range(numPartitions)]
        start = random.randrange(numPartitions) if numPartitions > 0 else 0
        pos = start
        for p in parts:
            for e in p:
                if isinstance(e, RowBlock):
                    n = len(e)
                    for k in range(numPartitions):
                        sl = slice((k - pos) % numPartitions, n, numPartitions)
                        chunk = e.take(sl)
                        if len(chunk):
                            out[k].append(chunk)
                    pos = (pos + n) % numPartitions
                else:
                    out[pos].append(e)
                    pos = (pos + 1) % numPartitions
        return RDD(self.ctx, partitions=out)

    def coalesce(self, numPartitions: int, shuffle: bool = False) -> "RDD":
        parts = self._materialize()
        k = max(1, min(numPartitions, len(parts)))
        out: List[list] = [[] for _ in range(k)]
        for i, p in enumerate(parts):
            out[i * k // max(1, len(parts))].extend(p)
        return RDD(self.ctx, partitions=out)

    def persist(self, *_a, **_k) -> "RDD":
        if self._cached is None and self._parts is None:
            self._cached = self._materialize()
        return self

    cache = persist

    # ---------------------------------------------------------------- actions
    def _run(self, concurrent: bool) -> List[list]:
        n = self.getNumPartitions()
        cluster = self.ctx._cluster if self.ctx is not None else None
        if cluster is None or not (concurrent or self._barrier):
            return [list(expand_blocks(self._compute(i))) for i in range(n)]

        def make_task(i: int):
            pruned = self._prune(i)
            return lambda _ctx, pruned=pruned, i=i: list(expand_blocks(pruned._compute(i)))

        return cluster.run_stage(make_task, n, barrier=self._barrier)

    def collect(self) -> list:
        return list(itertools.chain.from_iterable(self._run(concurrent=False)))

    def count(self) -> int:
        return len(self.collect())

    def take(self, n: int) -> list:
        out: list = []
        if self._barrier:
            return self.collect()[:n]
        for i in range(self.getNumPartitions()):
            for x in expand_blocks(self._compute(i)):
                out.append(x)
                if len(out) >= n:
                    return out
        return out

    def first(self):
        return self.take(1)[0]

    def foreach(self, f) -> None:
        self.mapPartitions(lambda it: [f(x) for x in it] and [])._run(concurrent=True)

    def foreachPartition(self, f) -> None:
        self.mapPartitions(lambda it: f(it) or [])._run(concurrent=True)

    def glom(self) -> "RDD":
        return PipelinedRDD(self, lambda _i, it: [list(it)])

    def toLocalIterator(self):
        return iter(self.collect())


class PipelinedRDD(RDD):
    """``PipelinedRDD(prev, func(split_index, iterator), preservesPartitioning,
    isFromBarrier)`` – same positional signature as pyspark's private class,
    which the reference trainer constructs directly (distributed.py:63)."""

    def __init__(self, prev: RDD, func, preservesPartitioning: bool = False,
                 isFromBarrier: bool = False, isFromBarrierCheck: bool = False):
        super().__init__(prev.ctx, prev=prev, func=func, barrier=bool(isFromBarrier))
        self.preservesPartitioning = preservesPartitioning
        self.is_barrier = self._barrier


class RDDBarrier:
    def __init__(self, rdd: RDD):
        self.rdd = rdd

    def mapPartitions(self, f, preservesPartitioning: bool = False) -> RDD:
        def func(_idx, it):
            return f(it)

        if getattr(f, "_accepts_blocks", False):
            func._accepts_blocks = True  # type: ignore[attr-defined]
        return PipelinedRDD(self.rdd, func, preservesPartitioning, isFromBarrier=True)

    def mapPartitionsWithIndex(self, f, preservesPartitioning: bool = False) -> RDD:
        return PipelinedRDD(self.rdd, f, preservesPartitioning, isFromBarrier=True)
