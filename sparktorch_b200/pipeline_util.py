"""Persistence of Python stages inside Spark ML pipelines.

Format contract (reference sparktorch/pipeline_util.py:17-131, SURVEY.md §2.6
item 3): a Python stage is saved as a ``StopWordsRemover`` whose ``stopWords``
are ``[csv, GUID]`` with ``csv = "b0,b1,...,bn,"`` – the decimal bytes of
``zlib(dill(stage))`` – and ``GUID = 4c1740b00d3c4ff6806a1402321572cb``.

The byte<->csv codec is split out (:func:`encode_stage` / :func:`decode_stage`)
so the format is testable without a JVM, and is vectorised with numpy (the
reference builds one Python string per byte).
"""
from __future__ import annotations

import zlib
from typing import List

import dill
import numpy as np

from sparktorch_b200._spark import ensure_pyspark

ensure_pyspark()

from pyspark.context import SparkContext  # noqa: E402
from pyspark.ml import Pipeline, PipelineModel  # noqa: E402
from pyspark.ml.feature import StopWordsRemover  # noqa: E402
from pyspark.ml.util import JavaMLReader, JavaMLWriter  # noqa: E402
from pyspark.ml.wrapper import JavaParams  # noqa: E402

PY_OBJ_GUID = "4c1740b00d3c4ff6806a1402321572cb"
CARRIER_JAVA_CLASS = "org.apache.spark.ml.feature.StopWordsRemover"


class PysparkObjId(object):
    """Constants identifying carrier stages."""

    @staticmethod
    def _getPyObjId() -> str:
        return PY_OBJ_GUID

    @staticmethod
    def _getCarrierClass(javaName: bool = False):
        return CARRIER_JAVA_CLASS if javaName else StopWordsRemover


# ------------------------------------------------------------------ codec
def encode_stage(obj) -> List[str]:
    """stage → ``[csv, GUID]``."""
    raw = np.frombuffer(zlib.compress(dill.dumps(obj)), dtype=np.uint8)
    csv = ",".join(raw.astype(str).tolist()) + "," if raw.size else ""
    return [csv, PY_OBJ_GUID]


def decode_stage(stop_words) -> object:
    """``[csv]`` or ``[csv, GUID]`` → stage."""
    fields = stop_words[0].split(",")[:-1]
    raw = np.asarray(fields, dtype=np.int64).astype(np.uint8).tobytes()
    return dill.loads(zlib.decompress(raw))


def load_byte_array(stop_words):
    """Name kept from the reference (pipeline_util.py:35-47)."""
    return decode_stage(stop_words)


def is_carrier(stage) -> bool:
    if not isinstance(stage, StopWordsRemover):
        return False
    words = stage.getStopWords()
    return bool(words) and words[-1] == PY_OBJ_GUID


# ------------------------------------------------------------------ unwrap
class PysparkPipelineWrapper(object):
    """Swap carrier stages of a loaded (nested) Pipeline/PipelineModel back to
    the Python objects they carry (reference pipeline_util.py:58-78)."""

    @staticmethod
    def unwrap(pipeline):
        if not isinstance(pipeline, (Pipeline, PipelineModel)):
            raise TypeError("Cannot recognize a pipeline of type %s." % type(pipeline))
        stages = pipeline.getStages() if isinstance(pipeline, Pipeline) else pipeline.stages
        for i, stage in enumerate(stages):
            if isinstance(stage, (Pipeline, PipelineModel)):
                stages[i] = PysparkPipelineWrapper.unwrap(stage)
            elif is_carrier(stage):
                stages[i] = decode_stage(stage.getStopWords()[:-1])
        if isinstance(pipeline, Pipeline):
            pipeline.setStages(stages)
        else:
            pipeline.stages = stages
        return pipeline


# ------------------------------------------------------------------ mixin
class PysparkReaderWriter(object):
    """Mixin giving a Python Estimator/Model ``write()/read()/load()`` through
    the carrier (reference pipeline_util.py:81-131)."""

    def write(self):
        return JavaMLWriter(self)

    @classmethod
    def read(cls):
        return JavaMLReader(StopWordsRemover)

    @classmethod
    def load(cls, path):
        return cls._from_java(cls.read().load(path))

    @classmethod
    def _from_java(cls, java_obj):
        return decode_stage(list(java_obj.getStopWords())[:-1])

    def _to_java(self):
        words = encode_stage(self)
        sc = SparkContext._active_spark_context
        java_array = sc._gateway.new_array(sc._gateway.jvm.java.lang.String, len(words))
        java_array[0:2] = words[0:2]
        java_obj = JavaParams._new_java_obj(CARRIER_JAVA_CLASS, self.uid)
        java_obj.setStopWords(java_array)
        return java_obj
