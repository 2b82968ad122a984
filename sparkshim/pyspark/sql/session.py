"""SparkSession: builder, createDataFrame (row- and block-backed), csv reader."""
from __future__ import annotations

import os
from typing import Optional, Sequence

from pyspark.conf import SparkConf
from pyspark.context import SparkContext
from pyspark.rdd import RDD, RowBlock
from pyspark.sql.dataframe import DataFrame
from pyspark.sql.types import Row


class _Builder:
    def __init__(self):
        self._master = None
        self._name = "sparkshim"
        self._conf = {}

    def master(self, m: str) -> "_Builder":
        self._master = m
        return self

    def appName(self, n: str) -> "_Builder":
        self._name = n
        return self

    def config(self, key=None, value=None, conf=None, **_k) -> "_Builder":
        if key is not None:
            self._conf[key] = value
        return self

    def enableHiveSupport(self) -> "_Builder":
        return self

    def getOrCreate(self) -> "SparkSession":
        if SparkSession._active is not None:
            return SparkSession._active
        # under torch.distributed.run the non-zero ranks are the executors
        from pyspark._cluster import serve_as_executor_rank, under_torchrun

        if under_torchrun() and int(os.environ.get("RANK", "0")) != 0:
            serve_as_executor_rank()  # never returns
        master = self._master or os.environ.get("SPARKSHIM_MASTER", "local[*]")
        sc = SparkContext._active_spark_context or SparkContext(master, self._name, SparkConf(self._conf))
        SparkSession._active = SparkSession(sc)
        return SparkSession._active


class _Reader:
    def __init__(self, session):
        self._session = session
        self._opts = {}

    def option(self, k, v) -> "_Reader":
        self._opts[k] = v
        return self

    def options(self, **kw) -> "_Reader":
        self._opts.update(kw)
        return self

    def csv(self, path: str, header=None, inferSchema=None, **_k) -> DataFrame:
        import numpy as np

        has_header = str(header if header is not None else self._opts.get("header", "false")).lower() == "true"
        data = np.loadtxt(path, delimiter=",", skiprows=1 if has_header else 0, dtype=np.float64, ndmin=2)
        if has_header:
            with open(path) as fh:
                cols = fh.readline().strip().split(",")
        else:
            cols = ["_c%d" % i for i in range(data.shape[1])]
        # columnar: every CSV column is a 1-D array inside one block per partition
        k = self._session.sparkContext.defaultParallelism
        parts = []
        for i in range(k):
            sl = slice(i * len(data) // k, (i + 1) * len(data) // k)
            parts.append([RowBlock(cols, [data[sl, j] for j in range(data.shape[1])])])
        return DataFrame(RDD(self._session.sparkContext, partitions=parts), cols, self._session)


class SparkSession:
    _active: "Optional[SparkSession]" = None
    builder = None  # set below (fresh builder per access)

    def __init__(self, sc: SparkContext):
        self.sparkContext = sc
        self._sc = sc

    @property
    def read(self) -> _Reader:
        return _Reader(self)

    def createDataFrame(self, data, schema: Optional[Sequence[str]] = None, numPartitions: Optional[int] = None) -> DataFrame:
        """``data``: list of tuples/Rows, a pandas DataFrame, or – columnar fast
        path – a dict ``{column: ndarray}`` (2-D arrays become vector columns)."""
        k = numPartitions or self.sparkContext.defaultParallelism
        if isinstance(data, dict):
            cols = list(schema) if schema else list(data.keys())
            arrays = [data[c] for c in cols]
            n = len(arrays[0])
            k = max(1, min(k, n))
            parts = []
            for i in range(k):
                sl = slice(i * n // k, (i + 1) * n // k)
                parts.append([RowBlock(cols, [a[sl] for a in arrays])])
            return DataFrame(RDD(self.sparkContext, partitions=parts), cols, self)
        if hasattr(data, "to_dict") and hasattr(data, "columns"):  # pandas
            cols = list(data.columns)
            data = [tuple(r) for r in data.itertuples(index=False)]
            schema = schema or cols
        data = list(data)
        if schema is None:
            first = data[0]
            schema = list(first.__fields__) if hasattr(first, "__fields__") else ["_%d" % (i + 1) for i in range(len(first))]
        cols = list(schema)
        rows = [Row._make(cols, tuple(r)) for r in data]
        k = max(1, min(k, max(1, len(rows))))
        parts = [rows[i * len(rows) // k:(i + 1) * len(rows) // k] for i in range(k)]
        return DataFrame(RDD(self.sparkContext, partitions=parts), cols, self)

    def stop(self) -> None:
        self.sparkContext.stop()
        SparkSession._active = None

    @classmethod
    def getActiveSession(cls):
        return cls._active


class _BuilderDescriptor:
    def __get__(self, _obj, _cls):
        return _Builder()


SparkSession.builder = _BuilderDescriptor()
