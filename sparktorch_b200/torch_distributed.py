"""Spark ML integration: the ``SparkTorch`` Estimator and ``SparkTorchModel``
Transformer.

Param names, defaults and getters are the reference's (sparktorch/
torch_distributed.py:59-358; table in SURVEY.md §2.5) so saved pipelines and
user code carry over.  Both classes are published under the module path
``sparktorch.torch_distributed`` because pipeline blobs reference them by name.

New optional Params (all default to "decide for me"):

* ``precision``  – ``auto|fp32|bf16`` compute dtype on GPU (fp32 master weights);
* ``engine``     – ``auto|eager|fused``;
* ``seed``       – reproducible sampling / validation split;
* ``streaming``  – keep partitions in pinned host memory and copy one batch per
  step (for partitions larger than HBM);
* on the model: ``device`` and ``batchSize`` for batched inference.
"""
from __future__ import annotations

import time

import numpy as np
import torch  # noqa: F401

from sparktorch_b200._spark import ensure_pyspark

_SPARK_KIND = ensure_pyspark()

from pyspark import SparkContext, keyword_only  # noqa: E402
from pyspark.ml import Model  # noqa: E402
from pyspark.ml.base import Estimator  # noqa: E402
from pyspark.ml.linalg import DenseVector, VectorUDT, Vectors  # noqa: E402,F401
from pyspark.ml.param import Param, Params, TypeConverters  # noqa: E402
from pyspark.ml.param.shared import HasInputCol, HasLabelCol, HasPredictionCol  # noqa: E402
from pyspark.ml.util import Identifiable, MLReadable, MLWritable  # noqa: E402
from pyspark.sql.types import DoubleType  # noqa: E402

from sparktorch_b200.distributed import SYNC_MODES, train_distributed  # noqa: E402
from sparktorch_b200.pipeline_util import PysparkReaderWriter  # noqa: E402
from sparktorch_b200.util import DataObj, load_torch_model, torch_decoder, torch_encoder  # noqa: E402


# --------------------------------------------------------------------- ingest
def handle_data(input_col, label_col):
    """Partition iterator → ``[DataObj]`` with ONE stacked block per partition
    (columnar; the reference makes one DataObj per row, torch_distributed.py:44-56)."""

    def inner(dataset):
        xs, ys = [], []
        for item in dataset:
            cols = getattr(item, "columns", None)
            if cols is not None and hasattr(item, "arrays"):       # sparkshim RowBlock: zero-copy
                xs.append(np.asarray(item.column(input_col), dtype=np.float32).reshape(len(item), -1))
                if label_col:
                    ys.append(np.asarray(item.column(label_col), dtype=np.float32).reshape(len(item), -1))
            else:
                v = item[input_col]
                xs.append(np.asarray(v.toArray() if hasattr(v, "toArray") else v, dtype=np.float32)[None, :])
                if label_col:
                    ys.append(np.asarray(item[label_col], dtype=np.float32).reshape(1, -1))
        if not xs:
            return []
        X = xs[0] if len(xs) == 1 else np.concatenate(xs, axis=0)
        Y = (ys[0] if len(ys) == 1 else np.concatenate(ys, axis=0)) if ys else None
        return [DataObj(x_train=X, y_train=Y, x_val=None, y_val=None)]

    inner._accepts_blocks = True
    return inner


def _param_or(obj, name, default):
    """Instances unpickled from pipelines saved by older code lack newer Params."""
    try:
        return obj.getOrDefault(getattr(obj, name))
    except (KeyError, AttributeError, ValueError):
        return default


# --------------------------------------------------------------------- model
class SparkTorchModel(Model, HasInputCol, HasPredictionCol, PysparkReaderWriter, MLReadable, MLWritable, Identifiable):
    modStr = Param(Params._dummy(), "modStr", "base64(dill(nn.Module))", typeConverter=TypeConverters.toString)
    useVectorOut = Param(Params._dummy(), "useVectorOut", "emit the raw output vector", typeConverter=TypeConverters.toBoolean)
    device = Param(Params._dummy(), "device", "inference device: auto|cpu|cuda", typeConverter=TypeConverters.toString)
    batchSize = Param(Params._dummy(), "batchSize", "rows per forward pass", typeConverter=TypeConverters.toInt)

    @keyword_only
    def __init__(self, inputCol=None, predictionCol=None, modStr=None, useVectorOut=None, device=None, batchSize=None):
        super().__init__()
        self._setDefault(inputCol="encoded", predictionCol="predicted", modStr="", useVectorOut=False,
                         device="auto", batchSize=65536)
        self.setParams(**self._input_kwargs)
        self.trainingHistory = None

    @keyword_only
    def setParams(self, inputCol=None, predictionCol=None, modStr=None, useVectorOut=None, device=None, batchSize=None):
        return self._set(**self._input_kwargs)

    def getPytorchModel(self):
        return torch_decoder(self.getOrDefault(self.modStr))

    def getTrainingHistory(self):
        return getattr(self, "trainingHistory", None)

    def _transform(self, dataset):
        inp = self.getOrDefault(self.inputCol)
        out = self.getOrDefault(self.predictionCol)
        vector_out = self.getOrDefault(self.useVectorOut)
        device = _param_or(self, "device", "auto")
        batch = _param_or(self, "batchSize", 65536)
        model_bc = dataset._sc.broadcast(torch_decoder(self.getOrDefault(self.modStr)))
        if _SPARK_KIND == "sparkshim":
            return _transform_shim(dataset, inp, out, vector_out, model_bc, device, batch)
        return _transform_pyspark(dataset, inp, out, vector_out, model_bc, device, batch)


def _make_partition_scorer(inp, out, vector_out, model_bc, device, batch):
    """Returns f(partition_index, iterator) scoring a whole partition in batches."""

    def score(idx, it):
        from sparktorch_b200.runtime.predict import BatchPredictor, pick_device, to_prediction_column

        predictor = None
        pending = []

        def flush_rows(rows):
            nonlocal predictor
            if not rows:
                return
            if predictor is None:
                predictor = BatchPredictor(model_bc.value, pick_device(device, idx), batch)
            X = np.stack([np.asarray(r[inp].toArray(), dtype=np.float32) for r in rows])
            col = to_prediction_column(predictor.predict_block(X), vector_out)
            for r, v in zip(rows, col):
                yield r, (DenseVector(v) if vector_out else float(v))

        for item in it:
            if hasattr(item, "arrays") and hasattr(item, "columns"):  # RowBlock → stays columnar
                for r, v in flush_rows(pending):
                    yield _append(r, out, v)
                pending = []
                if predictor is None:
                    predictor = BatchPredictor(model_bc.value, pick_device(device, idx), batch)
                X = np.asarray(item.column(inp), dtype=np.float32).reshape(len(item), -1)
                col = to_prediction_column(predictor.predict_block(X), vector_out)
                yield type(item)(item.columns + [out], item.arrays + [col])
            else:
                pending.append(item)
                if len(pending) >= batch:
                    for r, v in flush_rows(pending):
                        yield _append(r, out, v)
                    pending = []
        for r, v in flush_rows(pending):
            yield _append(r, out, v)

    score._accepts_blocks = True
    return score


def _append(row, name, value):
    from pyspark.sql.types import Row

    fields = list(row.__fields__) + [name]
    vals = tuple(row) + (value,)
    if hasattr(Row, "_make"):
        return Row._make(fields, vals)
    return Row(**dict(zip(fields, vals)))  # real pyspark


def _transform_shim(dataset, inp, out, vector_out, model_bc, device, batch):
    from pyspark.rdd import PipelinedRDD
    from pyspark.sql.dataframe import DataFrame

    scorer = _make_partition_scorer(inp, out, vector_out, model_bc, device, batch)
    cols = [c for c in dataset.columns if c != out] + [out]
    return DataFrame(PipelinedRDD(dataset.rdd, scorer), cols, dataset.sparkSession)


def _transform_pyspark(dataset, inp, out, vector_out, model_bc, device, batch):  # pragma: no cover (needs a JVM)
    from pyspark.sql.types import StructField, StructType

    scorer = _make_partition_scorer(inp, out, vector_out, model_bc, device, batch)
    schema = StructType(list(dataset.schema.fields) + [StructField(out, VectorUDT() if vector_out else DoubleType(), True)])
    rdd = dataset.rdd.mapPartitionsWithIndex(scorer)
    return dataset.sparkSession.createDataFrame(rdd, schema)


# --------------------------------------------------------------------- estimator
class SparkTorch(Estimator, HasInputCol, HasPredictionCol, HasLabelCol, PysparkReaderWriter, MLReadable, MLWritable,
                 Identifiable):
    torchObj = Param(Params._dummy(), "torchObj", "The serialized torch object", typeConverter=TypeConverters.toString)
    mode = Param(Params._dummy(), "mode", "synchronous (alias async) | hogwild", typeConverter=TypeConverters.toString)
    device = Param(Params._dummy(), "device", "cpu | cuda", typeConverter=TypeConverters.toString)
    iters = Param(Params._dummy(), "iters", "iterations per partition", typeConverter=TypeConverters.toInt)
    partitions = Param(Params._dummy(), "partitions", "repartition to this many replicas", typeConverter=TypeConverters.toInt)
    verbose = Param(Params._dummy(), "verbose", "print per-iteration losses", typeConverter=TypeConverters.toInt)
    acquireLock = Param(Params._dummy(), "acquireLock", "hogwild: consistent reads/updates", typeConverter=TypeConverters.toBoolean)
    partitionShuffles = Param(Params._dummy(), "partitionShuffles", "hogwild: data reshuffle rounds", typeConverter=TypeConverters.toInt)
    port = Param(Params._dummy(), "port", "hogwild: parameter-server port", typeConverter=TypeConverters.toInt)
    useBarrier = Param(Params._dummy(), "useBarrier", "hogwild: run workers as a barrier stage", typeConverter=TypeConverters.toBoolean)
    useVectorOut = Param(Params._dummy(), "useVectorOut", "model emits vectors", typeConverter=TypeConverters.toBoolean)
    earlyStopPatience = Param(Params._dummy(), "earlyStopPatience", ">0 enables early stopping", typeConverter=TypeConverters.toInt)
    miniBatch = Param(Params._dummy(), "miniBatch", "rows per step per replica", typeConverter=TypeConverters.toInt)
    validationPct = Param(Params._dummy(), "validationPct", "per-partition validation split", typeConverter=TypeConverters.toFloat)
    compileMode = Param(Params._dummy(), "compileMode", "accepted, ignored", typeConverter=TypeConverters.toString)
    precision = Param(Params._dummy(), "precision", "auto|fp32|bf16", typeConverter=TypeConverters.toString)
    engine = Param(Params._dummy(), "engine", "auto|eager|fused", typeConverter=TypeConverters.toString)
    seed = Param(Params._dummy(), "seed", "sampling seed (<0: nondeterministic)", typeConverter=TypeConverters.toInt)
    streaming = Param(Params._dummy(), "streaming", "stream batches from pinned host memory", typeConverter=TypeConverters.toBoolean)
    checkpointDir = Param(Params._dummy(), "checkpointDir", "directory for mid-training checkpoints", typeConverter=TypeConverters.toString)
    checkpointEvery = Param(Params._dummy(), "checkpointEvery", "iterations between checkpoints (0: off)", typeConverter=TypeConverters.toInt)
    resume = Param(Params._dummy(), "resume", "continue from the checkpoint in checkpointDir", typeConverter=TypeConverters.toBoolean)
    sampling = Param(Params._dummy(), "sampling", "random (fresh sample per step) | sweep (walk the shuffled partition)", typeConverter=TypeConverters.toString)

    @keyword_only
    def __init__(self, inputCol=None, labelCol=None, torchObj=None, iters=None, predictionCol=None, partitions=None,
                 acquireLock=None, verbose=None, partitionShuffles=None, port=None, useBarrier=None, useVectorOut=None,
                 earlyStopPatience=None, miniBatch=None, validationPct=None, mode=None, device=None, compileMode=None,
                 precision=None, engine=None, seed=None, streaming=None, sampling=None, checkpointDir=None,
                 checkpointEvery=None, resume=None):
        super().__init__()
        self._setDefault(inputCol="features", labelCol=None, torchObj="", iters=10, predictionCol="predicted",
                         partitions=-1, acquireLock=True, verbose=0, partitionShuffles=1, port=3000, useBarrier=False,
                         useVectorOut=False, earlyStopPatience=-1, miniBatch=-1, validationPct=0.0,
                         mode="synchronous", device="cpu", compileMode=None, precision="auto", engine="auto",
                         seed=-1, streaming=False, sampling="random", checkpointDir=None, checkpointEvery=0, resume=False)
        self.setParams(**self._input_kwargs)

    @keyword_only
    def setParams(self, inputCol=None, labelCol=None, torchObj=None, iters=None, predictionCol=None, partitions=None,
                  acquireLock=None, verbose=None, partitionShuffles=None, port=None, useBarrier=None, useVectorOut=None,
                  earlyStopPatience=None, miniBatch=None, validationPct=None, mode=None, device=None, compileMode=None,
                  precision=None, engine=None, seed=None, streaming=None, sampling=None, checkpointDir=None,
                  checkpointEvery=None, resume=None):
        return self._set(**self._input_kwargs)

    # getters (names as in the reference, incl. the historical misspelling)
    def getTorchObj(self): return self.getOrDefault(self.torchObj)
    def getEarlyStopPatience(self): return self.getOrDefault(self.earlyStopPatience)
    def getIters(self): return self.getOrDefault(self.iters)
    def getPartitions(self): return self.getOrDefault(self.partitions)
    def getVerbose(self): return self.getOrDefault(self.verbose)
    def getAqcuireLock(self): return self.getOrDefault(self.acquireLock)
    getAcquireLock = getAqcuireLock
    def getPartitionShuffles(self): return self.getOrDefault(self.partitionShuffles)
    def getPort(self): return self.getOrDefault(self.port)
    def getBarrier(self): return self.getOrDefault(self.useBarrier)
    def getVectorOut(self): return self.getOrDefault(self.useVectorOut)
    def getMiniBatch(self): return self.getOrDefault(self.miniBatch)
    def getValidationPct(self): return self.getOrDefault(self.validationPct)
    def getMode(self): return self.getOrDefault(self.mode)
    def getDevice(self): return self.getOrDefault(self.device)
    def getCompileMode(self): return self.getOrDefault(self.compileMode)

    def _fit(self, dataset):
        inp_col, label = self.getInputCol(), self.getLabelCol()
        prediction, torch_obj = self.getPredictionCol(), self.getTorchObj()
        partitions, mode, device = self.getPartitions(), self.getMode(), self.getDevice()
        seed = _param_or(self, "seed", -1)
        seed = None if seed is None or seed < 0 else seed
        precision, engine = _param_or(self, "precision", "auto"), _param_or(self, "engine", "auto")
        streaming = _param_or(self, "streaming", False)

        rdd = dataset.rdd
        if partitions > 0:                      # row-level shuffle happens before stacking
            rdd = rdd.repartition(partitions)
        rdd = rdd.mapPartitions(handle_data(inp_col, label))
        partitions = partitions if partitions > 0 else rdd.getNumPartitions()

        history = None
        if mode in SYNC_MODES:
            state_dict, history = train_distributed(
                rdd=rdd, torch_obj=torch_obj, iters=self.getIters(), partition_shuffles=1,
                verbose=self.getVerbose(), mini_batch=self.getMiniBatch(), validation_pct=self.getValidationPct(),
                device=device, early_stop_patience=self.getEarlyStopPatience(), compile_mode=self.getCompileMode(),
                precision=precision, engine=engine, seed=seed, streaming=streaming, return_history=True,
                time_window=getattr(self, "_time_window", None), sampling=_param_or(self, "sampling", "random"),
                checkpoint_dir=_param_or(self, "checkpointDir", None), checkpoint_every=_param_or(self, "checkpointEvery", 0),
                resume=_param_or(self, "resume", False))
        elif mode == "hogwild" and str(device).startswith("cuda"):
            from sparktorch_b200.hogwild import train_gpu

            state_dict, history = train_gpu(
                rdd, torch_obj, self.getIters(), self.getPartitionShuffles(), verbose=self.getVerbose(),
                early_stop_patience=self.getEarlyStopPatience(), mini_batch=self.getMiniBatch(),
                validation_pct=self.getValidationPct(), device=device, seed=seed,
                acquire_lock=self.getAqcuireLock(), window_len=partitions, time_window=getattr(self, "_time_window", None),
                sampling=_param_or(self, "sampling", "random"))
        elif mode == "hogwild":
            from sparktorch_b200.hogwild import get_main, train
            from sparktorch_b200.server import Server

            sc = SparkContext._active_spark_context
            host = str(sc.getConf().get("spark.driver.host")) if sc is not None else "127.0.0.1"
            port = self.getPort()
            if self.getBarrier():
                rdd = rdd.barrier().mapPartitions(lambda it: it)
            server = Server(torch_obj=torch_obj, master_url=host + ":" + str(port), port=port,
                            acquire_lock=self.getAqcuireLock(), early_stop_patience=self.getEarlyStopPatience(),
                            window_len=partitions)
            server.start_server()
            server.master_url = host + ":" + str(server.port)
            deadline = time.time() + 10
            while True:                           # readiness probe instead of sleep(5)
                try:
                    get_main(server.master_url)
                    break
                except Exception:
                    if time.time() > deadline:
                        server.stop_server()
                        raise
                    time.sleep(0.05)
            state_dict = train(rdd, torch_obj, server, self.getIters(), self.getPartitionShuffles(),
                               verbose=self.getVerbose(), early_stop_patience=self.getEarlyStopPatience(),
                               mini_batch=self.getMiniBatch(), validation_pct=self.getValidationPct(),
                               device=device, seed=seed)
        else:
            raise RuntimeError(f"Mode: {mode} not recognized.")

        model = load_torch_model(torch_obj, from_json=True, build_optimizer=False).model
        model.load_state_dict(state_dict)
        stm = SparkTorchModel(inputCol=inp_col, predictionCol=prediction, modStr=torch_encoder(model),
                              useVectorOut=self.getVectorOut())
        stm.trainingHistory = history
        return stm


# pipeline blobs reference these classes by module path
SparkTorchModel.__module__ = "sparktorch.torch_distributed"
SparkTorch.__module__ = "sparktorch.torch_distributed"
