"""sparkshim.pyspark – a single-node, pure-Python stand-in for the slice of the
PySpark API that SparkTorch-style estimators use (inventory: SURVEY.md §2.9).

It is only put on ``sys.path`` when the real ``pyspark`` is not importable (this
image has no JVM).  It runs real concurrent executors with barrier-stage
semantics (one process per partition/GPU), a Param/Estimator/Pipeline layer and
Spark's on-disk ML metadata layout, so both this repo's estimator and the
unmodified reference package run on it.
"""
import functools

__version__ = "3.5.0+sparkshim"
__shim__ = True


def keyword_only(func):
    """Force keyword arguments and stash them in ``self._input_kwargs``."""

    @functools.wraps(func)
    def wrapper(self, *args, **kwargs):
        if args:
            raise TypeError("Method %s forces keyword arguments." % func.__name__)
        self._input_kwargs = kwargs
        return func(self, **kwargs)

    return wrapper


from pyspark.conf import SparkConf  # noqa: E402
from pyspark.context import SparkContext  # noqa: E402
from pyspark.rdd import RDD  # noqa: E402
from pyspark.taskcontext import BarrierTaskContext, BarrierTaskInfo, TaskContext  # noqa: E402
from pyspark.sql.types import Row  # noqa: E402

__all__ = ["SparkConf", "SparkContext", "RDD", "TaskContext", "BarrierTaskContext", "BarrierTaskInfo", "Row",
           "keyword_only"]
