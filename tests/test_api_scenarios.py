"""API-level scenarios, run on a local[2] cluster with a real world_size-2 gloo
group – the reference's 12 integration scenarios (sparktorch/tests/
test_sparktorch.py:68-269) plus what it never tested: hogwild mode, pipeline
save/load/unwrap, attach_pytorch_model_to_pipeline, replica equality, loss
decrease, history."""
import numpy as np
import pytest
import torch
import torch.nn as nn

import sparktorch_b200  # noqa: F401
from pyspark.ml.linalg import Vectors
from sparktorch_b200 import (PysparkPipelineWrapper, SparkTorch, attach_pytorch_model_to_pipeline,
                             create_spark_torch_model, serialize_torch_obj, serialize_torch_obj_lazy)
from tests.nets import AutoEncoder, Classifier, Regressor, WidthRegressor


@pytest.fixture()
def data(spark):
    rng = np.random.default_rng(0)
    rows = [(1.0, Vectors.dense(rng.normal(0, 1, 10))) for _ in range(200)]
    rows += [(0.0, Vectors.dense(rng.normal(2, 1, 10))) for _ in range(200)]
    return spark.createDataFrame(rows, ["label", "features"]).repartition(2)


@pytest.fixture()
def sequential_model():
    net = nn.Sequential(nn.Linear(10, 20), nn.ReLU(), nn.Linear(20, 1))
    return serialize_torch_obj(net, nn.MSELoss(), torch.optim.Adam, lr=0.001)


@pytest.fixture()
def lazy_model():
    return serialize_torch_obj_lazy(Regressor, nn.MSELoss, torch.optim.Adam, optimizer_params={"lr": 0.001})


@pytest.fixture()
def general_model():
    return serialize_torch_obj(Regressor(), nn.MSELoss(), torch.optim.Adam, lr=0.001)


def _fit(data, obj, **kw):
    kw.setdefault("inputCol", "features")
    kw.setdefault("labelCol", "label")
    kw.setdefault("predictionCol", "predictions")
    return SparkTorch(torchObj=obj, **kw).fit(data)


def test_early_stopping(data, general_model):
    stm = _fit(data, general_model, iters=25, verbose=1, earlyStopPatience=2)
    res = stm.transform(data).take(1)
    assert "predictions" in res[0]


def test_early_stop_triggers_and_skips_update(data):
    # lr=0 → loss never improves → must stop after `patience` bad steps
    obj = serialize_torch_obj(Regressor(), nn.MSELoss(), torch.optim.SGD, lr=0.0)
    stm = _fit(data, obj, iters=50, earlyStopPatience=3)
    h = stm.getTrainingHistory()
    assert h["stopped_at"] == 3 and h["iters_run"] == 4


def test_model_parameters(data):
    obj = serialize_torch_obj_lazy(WidthRegressor, nn.MSELoss, torch.optim.Adam,
                                   optimizer_params={"lr": 0.001}, model_parameters={"param": 40})
    stm = _fit(data, obj, iters=5, verbose=1)
    net = stm.getPytorchModel()
    assert net.fc1.out_features == 40 and net.fc2 is not None


def test_inference_roundtrip(lazy_model, data):
    stm = _fit(data, lazy_model, iters=10)
    first = stm.transform(data).take(1)
    again = create_spark_torch_model(stm.getPytorchModel(), "features", "predictions").transform(data).take(1)
    assert first == again


def test_lazy(lazy_model, data):
    res = _fit(data, lazy_model, iters=5).transform(data).take(1)
    assert "predictions" in res[0] and type(res[0]["predictions"]) is float


def test_sequential(data, sequential_model):
    res = _fit(data, sequential_model, iters=5).transform(data).take(1)
    assert type(res[0]["predictions"]) is float


def test_module(data, general_model):
    res = _fit(data, general_model, iters=5).transform(data).take(1)
    assert type(res[0]["predictions"]) is float


def test_barrier_flag(data, general_model):
    res = _fit(data, general_model, iters=5, partitions=2, useBarrier=True).transform(data).take(1)
    assert "predictions" in res[0]


def test_autoencoder_vector_out(data):
    obj = serialize_torch_obj(AutoEncoder(), nn.MSELoss(), torch.optim.Adam, lr=0.001)
    stm = SparkTorch(inputCol="features", predictionCol="predictions", torchObj=obj, iters=5, partitions=2,
                     useVectorOut=True).fit(data)
    res = stm.transform(data).take(1)
    assert len(res[0]["predictions"]) == 10


def test_classification(data):
    obj = serialize_torch_obj(Classifier(), nn.CrossEntropyLoss(), torch.optim.Adam, lr=0.01)
    stm = _fit(data, obj, iters=60, partitions=2, seed=1)
    rows = stm.transform(data).collect()
    acc = np.mean([r["predictions"] == r["label"] for r in rows])
    assert acc > 0.9          # two well separated blobs
    assert set(r["predictions"] for r in rows) <= {0.0, 1.0}


def test_mini_batch(data, general_model):
    stm = _fit(data, general_model, iters=10, partitions=2, miniBatch=5, acquireLock=True)
    assert stm.getTrainingHistory()["samples_per_step"] == 10


def test_cpu_device(data, general_model):
    res = _fit(data, general_model, iters=10, partitions=2, miniBatch=5, device="cpu").transform(data).take(1)
    assert "predictions" in res[0]


def test_validation_pct(data, general_model):
    res = _fit(data, general_model, iters=10, partitions=2, validationPct=0.25, earlyStopPatience=5).transform(data).take(1)
    assert "predictions" in res[0]


def test_async_alias_and_bad_mode(data, general_model):
    _fit(data, general_model, iters=2, mode="async")
    with pytest.raises(RuntimeError):
        _fit(data, general_model, iters=2, mode="nope")


def test_loss_decreases_and_history(data):
    obj = serialize_torch_obj(Regressor(), nn.MSELoss(), torch.optim.Adam, lr=0.01)
    h = _fit(data, obj, iters=40).getTrainingHistory()
    assert h["world_size"] == 2 and len(h["train_loss"]) == 40
    assert h["train_loss"][-1] < 0.5 * h["train_loss"][0]


def test_sgd_momentum_and_other_optimizer(data):
    obj = serialize_torch_obj(Regressor(), nn.MSELoss(), torch.optim.SGD, lr=0.01, momentum=0.9, nesterov=True)
    h = _fit(data, obj, iters=20).getTrainingHistory()
    assert h["train_loss"][-1] < h["train_loss"][0]
    obj = serialize_torch_obj(Regressor(), nn.MSELoss(), torch.optim.RMSprop, lr=0.01)
    h = _fit(data, obj, iters=20).getTrainingHistory()
    assert h["train_loss"][-1] < h["train_loss"][0]


def test_hogwild(data, general_model):
    stm = _fit(data, general_model, iters=10, mode="hogwild", port=0, partitions=2)
    res = stm.transform(data).take(1)
    assert type(res[0]["predictions"]) is float


def test_hogwild_learns_with_lock_and_early_stop(data):
    obj = serialize_torch_obj(Classifier(), nn.CrossEntropyLoss(), torch.optim.Adam, lr=0.02)
    stm = _fit(data, obj, iters=60, mode="hogwild", port=0, acquireLock=True, miniBatch=64,
               earlyStopPatience=50, partitionShuffles=2)
    rows = stm.transform(data).collect()
    assert np.mean([r["predictions"] == r["label"] for r in rows]) > 0.9


def test_pipeline_save_load_unwrap(tmp_path, data, general_model):
    from pyspark.ml.feature import VectorAssembler
    from pyspark.ml.pipeline import Pipeline, PipelineModel

    est = SparkTorch(inputCol="features", labelCol="label", predictionCol="predictions", torchObj=general_model, iters=3)
    pm = Pipeline(stages=[est]).fit(data)
    path = str(tmp_path / "pipe")
    pm.write().overwrite().save(path)
    loaded = PysparkPipelineWrapper.unwrap(PipelineModel.load(path))
    assert loaded.transform(data).take(1) == pm.transform(data).take(1)
    # the estimator itself can be saved too (unfitted pipeline)
    p2 = str(tmp_path / "est")
    Pipeline(stages=[est]).write().overwrite().save(p2)
    est2 = PysparkPipelineWrapper.unwrap(Pipeline.load(p2)).getStages()[0]
    assert est2.getIters() == 3
    # nested + attach
    att = attach_pytorch_model_to_pipeline(pm.stages[0].getPytorchModel(), pm, "features", "p2")
    p3 = str(tmp_path / "nested")
    att.write().overwrite().save(p3)
    back = PysparkPipelineWrapper.unwrap(PipelineModel.load(p3))
    row = back.transform(data).take(1)[0]
    assert row["p2"] == row["predictions"]
    assert VectorAssembler is not None


def test_direct_model_save_load(tmp_path, data, general_model):
    from sparktorch_b200.torch_distributed import SparkTorchModel

    stm = _fit(data, general_model, iters=2)
    p = str(tmp_path / "m")
    stm.write().overwrite().save(p)
    back = SparkTorchModel.load(p)
    assert back.transform(data).take(1) == stm.transform(data).take(1)


def test_columnar_frame_matches_row_frame(spark, general_model):
    rng = np.random.default_rng(3)
    X = rng.normal(size=(64, 10))
    y = rng.integers(0, 2, size=64).astype(np.float64)
    block_df = spark.createDataFrame({"label": y, "features": X}, numPartitions=2)
    row_df = spark.createDataFrame([(float(a), Vectors.dense(b)) for a, b in zip(y, X)], ["label", "features"])
    stm = _fit(block_df, general_model, iters=3)
    a = [r["predictions"] for r in stm.transform(block_df).collect()]
    b = [r["predictions"] for r in stm.transform(row_df).collect()]
    np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6)
