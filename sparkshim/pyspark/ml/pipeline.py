"""Pipeline / PipelineModel with Spark's directory layout
(``metadata/`` + ``stages/<idx>_<uid>/``)."""
from __future__ import annotations

import os

from pyspark import keyword_only
from pyspark.ml.base import Estimator, Model, Transformer
from pyspark.ml.param import Param, Params
from pyspark.ml.util import (MLReadable, MLReader, MLWritable, MLWriter, _JAVA_CLASS_TO_PY, _read_metadata,
                             _write_metadata)


def _stage_dir(path: str, idx: int, n: int, uid: str) -> str:
    width = len(str(n))
    return os.path.join(path, "stages", str(idx).zfill(width) + "_" + uid)


class _PipelineWriter(MLWriter):
    def __init__(self, instance, java_class: str, stages):
        super().__init__()
        self._instance, self._java_class, self._stages = instance, java_class, stages

    def saveImpl(self, path: str) -> None:
        uids = [s.uid for s in self._stages]
        _write_metadata(path, self._java_class, self._instance.uid, {"stageUids": uids}, {})
        for i, s in enumerate(self._stages):
            w = s.write()
            w.shouldOverwrite = True
            w.save(_stage_dir(path, i, len(self._stages), s.uid))


class _PipelineReader(MLReader):
    def __init__(self, cls):
        self._cls = cls

    def load(self, path: str):
        import pyspark.ml.feature  # noqa: F401

        meta = _read_metadata(path)
        uids = meta["paramMap"]["stageUids"]
        stages = []
        for i, uid in enumerate(uids):
            d = _stage_dir(path, i, len(uids), uid)
            smeta = _read_metadata(d)
            jc = smeta["class"]
            if jc.endswith(".PipelineModel"):
                stages.append(PipelineModel.load(d))
            elif jc.endswith(".Pipeline"):
                stages.append(Pipeline.load(d))
            else:
                stages.append(_JAVA_CLASS_TO_PY[jc].load(d))
        obj = PipelineModel(stages) if self._cls is PipelineModel else Pipeline(stages=stages)
        obj.uid = meta["uid"]
        return obj


class Pipeline(Estimator, MLReadable, MLWritable):
    stages = Param(Params._dummy(), "stages", "a list of pipeline stages")

    @keyword_only
    def __init__(self, *, stages=None):
        super().__init__()
        self.setParams(**self._input_kwargs)

    @keyword_only
    def setParams(self, *, stages=None):
        return self._set(**self._input_kwargs)

    def setStages(self, value):
        return self._set(stages=value)

    def getStages(self):
        return self.getOrDefault(self.stages)

    def _fit(self, dataset):
        stages = self.getStages()
        last_est = -1
        for i, s in enumerate(stages):
            if isinstance(s, Estimator):
                last_est = i
        fitted = []
        for i, s in enumerate(stages):
            if i <= last_est:
                if isinstance(s, Transformer):
                    fitted.append(s)
                    dataset = s.transform(dataset)
                else:
                    m = s.fit(dataset)
                    fitted.append(m)
                    if i < last_est:
                        dataset = m.transform(dataset)
            else:
                fitted.append(s)
        return PipelineModel(fitted)

    def write(self):
        return _PipelineWriter(self, "org.apache.spark.ml.Pipeline", self.getStages())

    @classmethod
    def read(cls):
        return _PipelineReader(cls)


class PipelineModel(Model, MLReadable, MLWritable):
    def __init__(self, stages):
        super().__init__()
        self.stages = stages

    def _transform(self, dataset):
        for s in self.stages:
            dataset = s.transform(dataset)
        return dataset

    def write(self):
        return _PipelineWriter(self, "org.apache.spark.ml.PipelineModel", self.stages)

    @classmethod
    def read(cls):
        return _PipelineReader(cls)
