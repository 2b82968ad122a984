"""VectorAssembler and StopWordsRemover (the latter doubles as the pipeline
carrier for Python stages)."""
from __future__ import annotations

import numpy as np

from pyspark import keyword_only
from pyspark.ml.base import Transformer
from pyspark.ml.linalg import DenseVector
from pyspark.ml.param import Param, Params, TypeConverters
from pyspark.ml.param.shared import HasInputCol, HasInputCols, HasOutputCol
from pyspark.ml.util import JavaMLReadable, JavaMLWritable, _register
from pyspark.ml.wrapper import JavaParams
from pyspark.rdd import PipelinedRDD, RowBlock
from pyspark.sql.types import Row


@_register("org.apache.spark.ml.feature.VectorAssembler")
class VectorAssembler(Transformer, HasInputCols, HasOutputCol, JavaParams, JavaMLReadable, JavaMLWritable):
    handleInvalid = Param(Params._dummy(), "handleInvalid", "error|skip|keep", typeConverter=TypeConverters.toString)

    @keyword_only
    def __init__(self, *, inputCols=None, outputCol=None, handleInvalid="error"):
        super().__init__()
        self._setDefault(handleInvalid="error")
        self.setParams(**self._input_kwargs)

    @keyword_only
    def setParams(self, *, inputCols=None, outputCol=None, handleInvalid="error"):
        return self._set(**self._input_kwargs)

    def _transform(self, dataset):
        in_cols = self.getInputCols()
        out = self.getOutputCol()
        cols = dataset.columns + [out]

        def assemble(_i, it, in_cols=in_cols, cols=cols, out=out):
            for e in it:
                if isinstance(e, RowBlock):  # stay columnar
                    parts = [np.asarray(e.column(c), dtype=np.float64) for c in in_cols]
                    parts = [p[:, None] if p.ndim == 1 else p for p in parts]
                    mat = np.concatenate(parts, axis=1) if len(parts) > 1 else parts[0]
                    yield RowBlock(e.columns + [out], e.arrays + [mat])
                else:
                    vals = []
                    for c in in_cols:
                        v = e[c]
                        if isinstance(v, DenseVector):
                            vals.extend(v.toArray().tolist())
                        else:
                            vals.append(float(v))
                    yield Row._make(cols, tuple(e) + (DenseVector(vals),))

        assemble._accepts_blocks = True
        from pyspark.sql.dataframe import DataFrame

        return DataFrame(PipelinedRDD(dataset.rdd, assemble), cols, dataset.sparkSession)


@_register("org.apache.spark.ml.feature.StopWordsRemover")
class StopWordsRemover(Transformer, HasInputCol, HasOutputCol, JavaParams, JavaMLReadable, JavaMLWritable):
    stopWords = Param(Params._dummy(), "stopWords", "The words to be filtered out",
                      typeConverter=TypeConverters.toListString)
    caseSensitive = Param(Params._dummy(), "caseSensitive", "case sensitive", typeConverter=TypeConverters.toBoolean)
    locale = Param(Params._dummy(), "locale", "locale", typeConverter=TypeConverters.toString)

    @keyword_only
    def __init__(self, *, inputCol=None, outputCol=None, stopWords=None, caseSensitive=False, locale=None):
        super().__init__()
        self._setDefault(stopWords=["a", "an", "the"], caseSensitive=False, locale="en_US")
        self.setParams(**self._input_kwargs)

    @keyword_only
    def setParams(self, *, inputCol=None, outputCol=None, stopWords=None, caseSensitive=False, locale=None):
        return self._set(**self._input_kwargs)

    def setStopWords(self, value):
        return self._set(stopWords=list(value))

    def getStopWords(self):
        return self.getOrDefault(self.stopWords)

    def _transform(self, dataset):
        inp, out = self.getInputCol(), self.getOutputCol()
        sw = set(self.getStopWords())
        from pyspark.sql.functions import udf

        return dataset.withColumn(out, udf(lambda ws: [w for w in ws if w not in sw])(inp))
