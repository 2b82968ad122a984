"""ML persistence in Spark's on-disk layout.

``<path>/metadata/part-00000`` holds one JSON line ``{"class", "timestamp",
"sparkVersion", "uid", "paramMap", "defaultParamMap"}`` – the same metadata
record ``DefaultParamsWriter`` produces in Spark – so a saved carrier stage
(StopWordsRemover with the dill payload in ``stopWords``) looks on disk like
one written by a real cluster."""
from __future__ import annotations

import json
import os
import shutil
import time

from pyspark.ml.param import Identifiable  # re-export  # noqa: F401

_JAVA_CLASS_TO_PY = {}


def _register(java_name: str):
    def deco(cls):
        cls._java_class_name = java_name
        _JAVA_CLASS_TO_PY[java_name] = cls
        return cls

    return deco


def _write_metadata(path: str, java_class: str, uid: str, param_map: dict, default_map: dict, extra=None) -> None:
    os.makedirs(os.path.join(path, "metadata"), exist_ok=True)
    rec = {
        "class": java_class,
        "timestamp": int(time.time() * 1000),
        "sparkVersion": "3.5.0",
        "uid": uid,
        "paramMap": param_map,
        "defaultParamMap": default_map,
    }
    if extra:
        rec.update(extra)
    with open(os.path.join(path, "metadata", "part-00000"), "w") as fh:
        fh.write(json.dumps(rec) + "\n")
    open(os.path.join(path, "metadata", "_SUCCESS"), "w").close()


def _read_metadata(path: str) -> dict:
    with open(os.path.join(path, "metadata", "part-00000")) as fh:
        return json.loads(fh.readline())


def _jsonable(v):
    import numpy as np

    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating,)):
        return float(v)
    if isinstance(v, (list, tuple)):
        return [_jsonable(x) for x in v]
    return v


class MLWriter:
    def __init__(self):
        self.shouldOverwrite = False

    def overwrite(self):
        self.shouldOverwrite = True
        return self

    def option(self, *_a):
        return self

    def session(self, *_a):
        return self

    def _prepare(self, path: str) -> None:
        if os.path.exists(path):
            if not self.shouldOverwrite:
                raise IOError("Path %s already exists. To overwrite it, please use write.overwrite().save(path)" % path)
            shutil.rmtree(path)
        os.makedirs(path)

    def save(self, path: str) -> None:
        self._prepare(path)
        self.saveImpl(path)

    def saveImpl(self, path: str) -> None:
        raise NotImplementedError()


class JavaMLWriter(MLWriter):
    """Writer for instances that have a JVM twin: ``instance._to_java()`` gives
    the object whose params are persisted."""

    def __init__(self, instance):
        super().__init__()
        self._instance = instance

    def saveImpl(self, path: str) -> None:
        jobj = self._instance._to_java()
        jobj._save_params(path)


class MLReader:
    def load(self, path: str):
        raise NotImplementedError()


class JavaMLReader(MLReader):
    def __init__(self, clazz):
        self._clazz = clazz

    def load(self, path: str):
        meta = _read_metadata(path)
        cls = _JAVA_CLASS_TO_PY.get(meta["class"], self._clazz)
        return cls._load_params(path, meta)


class MLWritable:
    def write(self) -> MLWriter:
        raise NotImplementedError()

    def save(self, path: str) -> None:
        self.write().save(path)


class MLReadable:
    @classmethod
    def read(cls) -> MLReader:
        raise NotImplementedError()

    @classmethod
    def load(cls, path: str):
        return cls.read().load(path)


class JavaMLWritable(MLWritable):
    def write(self):
        return JavaMLWriter(self)


class JavaMLReadable(MLReadable):
    @classmethod
    def read(cls):
        return JavaMLReader(cls)


DefaultParamsWritable = JavaMLWritable
DefaultParamsReadable = JavaMLReadable
