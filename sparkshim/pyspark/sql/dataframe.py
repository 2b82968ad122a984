"""DataFrame over an RDD of Rows (or columnar RowBlocks)."""
from __future__ import annotations

import random
from typing import List, Sequence

from pyspark.rdd import RDD, RowBlock, expand_blocks
from pyspark.sql.functions import Column, _as_column
from pyspark.sql.types import Row


class DataFrame:
    def __init__(self, rdd: RDD, columns: Sequence[str], session):
        self._rdd = rdd
        self._columns = list(columns)
        self.sparkSession = session
        self.sql_ctx = session
        self._sc = session.sparkContext

    # ------------------------------------------------------------- metadata
    @property
    def columns(self) -> List[str]:
        return list(self._columns)

    @property
    def rdd(self) -> RDD:
        return self._rdd

    def __getitem__(self, name):
        if isinstance(name, str):
            return _as_column(name)
        raise TypeError(name)

    # -------------------------------------------------------- transformations
    def _derive(self, rdd: RDD, columns=None) -> "DataFrame":
        return DataFrame(rdd, self._columns if columns is None else columns, self.sparkSession)

    def withColumn(self, name: str, column: Column) -> "DataFrame":
        cols = self._columns if name in self._columns else self._columns + [name]
        replace = name in self._columns

        def add(_i, it, column=column, cols=cols, name=name, replace=replace):
            for r in it:
                v = column.eval(r)
                if replace:
                    d = r.asDict()
                    d[name] = v
                    yield Row._make(cols, [d[c] for c in cols])
                else:
                    yield Row._make(cols, tuple(r) + (v,))

        from pyspark.rdd import PipelinedRDD

        return self._derive(PipelinedRDD(self._rdd, add), cols)

    def select(self, *names) -> "DataFrame":
        if len(names) == 1 and isinstance(names[0], (list, tuple)):
            names = tuple(names[0])
        cs = [_as_column(n) for n in names]
        cols = [c._name for c in cs]
        srcs = [getattr(c, "_source", None) for c in cs]
        if all(s is not None and s in self._columns for s in srcs):
            # plain column references: columnar blocks stay columnar (no per-row objects for dropped columns)
            def project(_i, it, cols=cols, srcs=srcs):
                for e in it:
                    if isinstance(e, RowBlock):
                        yield RowBlock(cols, [e.column(s) for s in srcs])
                    else:
                        yield Row._make(cols, [e[s] for s in srcs])

            project._accepts_blocks = True
            from pyspark.rdd import PipelinedRDD

            return self._derive(PipelinedRDD(self._rdd, project), cols)
        return self._derive(self._rdd.map(lambda r: Row._make(cols, [c.eval(r) for c in cs])), cols)

    def drop(self, *names) -> "DataFrame":
        keep = [c for c in self._columns if c not in names]
        return self.select(*keep)

    def repartition(self, n: int, *_cols) -> "DataFrame":
        return self._derive(self._rdd.repartition(n))

    def coalesce(self, n: int) -> "DataFrame":
        return self._derive(self._rdd.coalesce(n))

    def orderBy(self, *cols, **_kw) -> "DataFrame":
        cs = [_as_column(c) for c in cols]
        parts = self._rdd._materialize()
        n_parts = max(1, len(parts))
        if cs and all(c._is_rand for c in cs):
            # shuffle; keep columnar blocks columnar
            out = []
            for p in parts:
                rows, blocks = [], []
                for e in p:
                    (blocks if isinstance(e, RowBlock) else rows).append(e)
                import numpy as np

                for b in blocks:
                    out.append(b.take(np.random.permutation(len(b))))
                random.shuffle(rows)
                out.extend(rows)
            random.shuffle(out)
            k = n_parts
            new_parts = [out[i::k] for i in range(k)]
            return self._derive(RDD(self._sc, partitions=new_parts))
        rows = [r for p in parts for r in expand_blocks(p)]
        rows.sort(key=lambda r: tuple(c.eval(r) for c in cs))
        k = n_parts
        new_parts = [rows[i * len(rows) // k:(i + 1) * len(rows) // k] for i in range(k)]
        return self._derive(RDD(self._sc, partitions=new_parts))

    sort = orderBy

    def persist(self, *_a, **_k) -> "DataFrame":
        self._rdd.persist()
        return self

    cache = persist

    def unpersist(self, *_a) -> "DataFrame":
        return self

    def limit(self, n: int) -> "DataFrame":
        rows = self.take(n)
        return self._derive(RDD(self._sc, partitions=[rows]))

    # ---------------------------------------------------------------- actions
    def collect(self) -> list:
        return self._rdd.collect()

    def take(self, n: int) -> list:
        return self._rdd.take(n)

    def head(self, n: int = 1):
        rows = self.take(n)
        return rows[0] if n == 1 and rows else rows

    def first(self):
        return self.head(1)

    def count(self) -> int:
        total = 0
        for i in range(self._rdd.getNumPartitions()):
            for e in self._rdd._compute(i):
                total += len(e) if isinstance(e, RowBlock) else 1
        return total

    def show(self, n: int = 20, *_a, **_k) -> None:
        print(self._columns)
        for r in self.take(n):
            print(tuple(r))

    def toPandas(self):
        import pandas as pd

        return pd.DataFrame([r.asDict() for r in self.collect()], columns=self._columns)
