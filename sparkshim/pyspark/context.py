"""SparkContext: owns the executor pool; ``parallelize``/``broadcast`` and the
tiny py4j look-alike that the pipeline carrier code path touches
(``sc._gateway.jvm.java.lang.String``, ``sc._gateway.new_array``)."""
from __future__ import annotations

import re
from typing import Any, List, Optional

from pyspark.conf import SparkConf


class Broadcast:
    def __init__(self, value: Any):
        self.value = value

    def unpersist(self, *_a):
        pass

    def destroy(self, *_a):
        pass


class _JavaStringClass:
    pass


class _Jvm:
    class java:  # noqa: N801
        class lang:  # noqa: N801
            String = _JavaStringClass


class _Gateway:
    jvm = _Jvm

    @staticmethod
    def new_array(_cls, n: int) -> list:
        return [None] * n


def _stub_context():
    return SparkContext._active_spark_context


class SparkContext:
    _active_spark_context: "Optional[SparkContext]" = None
    _gateway = _Gateway

    def __init__(self, master: str = "local[*]", appName: str = "sparkshim", conf: Optional[SparkConf] = None):
        import os

        from pyspark._cluster import LocalCluster

        m = re.match(r"local\[(\d+|\*)\]", master or "local[*]")
        n = (os.cpu_count() or 1) if (m is None or m.group(1) == "*") else int(m.group(1))
        if master == "local":
            n = 1
        self.master = master
        self.appName = appName
        self._conf = conf or SparkConf()
        self._conf.set("spark.driver.host", self._conf.get("spark.driver.host", "127.0.0.1"))
        self._conf.set("spark.master", master)
        self.defaultParallelism = n
        self._cluster = LocalCluster(n)
        self.defaultParallelism = self._cluster.n_slots
        SparkContext._active_spark_context = self

    # executors get a stub (never the pool) when closures capture the context
    def __reduce__(self):
        return (_stub_context, ())

    def getConf(self) -> SparkConf:
        return self._conf

    def broadcast(self, value: Any) -> Broadcast:
        return Broadcast(value)

    def parallelize(self, data, numSlices: Optional[int] = None):
        from pyspark.rdd import RDD

        data = list(data)
        k = max(1, numSlices or self.defaultParallelism)
        parts: List[list] = [data[i * len(data) // k:(i + 1) * len(data) // k] for i in range(k)]
        return RDD(self, partitions=parts)

    def stop(self) -> None:
        self._cluster.shutdown()
        if SparkContext._active_spark_context is self:
            SparkContext._active_spark_context = None

    @classmethod
    def getOrCreate(cls, conf: Optional[SparkConf] = None) -> "SparkContext":
        return cls._active_spark_context or cls(conf=conf)
