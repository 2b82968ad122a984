"""Wire/disk format compatibility with the reference implementation
(SURVEY.md §2.6).  The reference's pure-Python modules are loaded straight from
its source tree (no pyspark needed for util/early_stopper/rw_lock)."""
import base64
import codecs
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time
import zlib

import dill
import numpy as np
import pytest
import torch
import torch.nn as nn

import sparktorch_b200  # noqa: F401
from sparktorch_b200 import util
from sparktorch_b200.early_stopper import EarlyStopping
from sparktorch_b200.pipeline_util import PY_OBJ_GUID, decode_stage, encode_stage
from sparktorch_b200.rw_lock import RWLock
from tests.nets import Regressor

REF_DIRS = ["/root/reference/sparktorch", os.path.join(os.path.dirname(__file__), "..", "baseline", "_ref", "sparktorch")]
REF = next((d for d in REF_DIRS if os.path.isfile(os.path.join(d, "util.py"))), None)
needs_ref = pytest.mark.skipif(REF is None, reason="reference sources not present")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_ref(name):
    spec = importlib.util.spec_from_file_location("_ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_base64_codec_is_mime_style():
    raw = os.urandom(1000)
    assert base64.encodebytes(raw) == codecs.encode(raw, "base64")
    s = util.torch_encoder({"a": 1})
    assert all(len(line) <= 76 for line in s.splitlines())
    assert util.torch_decoder(s) == {"a": 1}


def test_envelope_layout():
    net = nn.Sequential(nn.Linear(4, 3), nn.ReLU(), nn.Linear(3, 2))
    env = json.loads(util.serialize_torch_obj(net, nn.MSELoss(), torch.optim.SGD, lr=0.5))
    assert set(env) == {"torch_obj", "shapes"}
    assert env["shapes"] == [[3, 4], [3], [2, 3], [2]]
    obj = util.torch_decoder(env["torch_obj"])
    assert type(obj).__name__ == "TorchObj" and type(obj).__module__ == "sparktorch.util"
    assert obj.optimizer.__name__ == "SGD" and obj.optimizer_params == {"lr": 0.5} and obj.is_lazy is False
    lazy = json.loads(util.serialize_torch_obj_lazy(Regressor, nn.MSELoss, torch.optim.Adam, {"lr": 1e-3}))
    assert lazy["shapes"] == [[20, 10], [20], [1, 20], [1]]
    lobj = util.torch_decoder(lazy["torch_obj"])
    assert lobj.is_lazy and lobj.model is Regressor and lobj.criterion.__name__ == "MSELoss"


def test_load_torch_model_builds_objects():
    s = util.serialize_torch_obj_lazy(Regressor, nn.MSELoss, torch.optim.Adam, {"lr": 0.01})
    t = util.load_torch_model(s, from_json=True)
    assert isinstance(t.model, Regressor) and isinstance(t.criterion, nn.MSELoss)
    assert type(t.optimizer).__name__ == "Adam" and t.optimizer.defaults["lr"] == 0.01


@needs_ref
def test_blobs_interoperate_with_reference_util():
    ref = _load_ref("util")
    # dill pickles TorchObj by reference to sparktorch.util → both directions must resolve
    import sparktorch.util as alias  # the alias package in this repo

    assert alias.TorchObj is util.TorchObj
    net = nn.Linear(3, 2)
    mine = util.serialize_torch_obj(net, nn.MSELoss(), torch.optim.Adam, lr=0.1)
    theirs = ref.serialize_torch_obj(net, nn.MSELoss(), torch.optim.Adam, lr=0.1)
    assert json.loads(mine)["shapes"] == json.loads(theirs)["shapes"]
    a = ref.load_torch_model(mine, from_json=True)      # reference reads our blob
    b = util.load_torch_model(theirs, from_json=True)   # we read the reference's blob
    for m in (a.model, b.model):
        assert torch.equal(m.weight, net.weight)
    assert a.optimizer.defaults["lr"] == b.optimizer.defaults["lr"] == 0.1
    lz = ref.serialize_torch_obj_lazy(Regressor, nn.MSELoss, torch.optim.Adam, {"lr": 0.2}, None)
    assert isinstance(util.load_torch_model(lz, from_json=True).model, Regressor)


@needs_ref
def test_handle_features_matches_reference_without_split():
    ref = _load_ref("util")
    rng = np.random.default_rng(0)
    rows = [util.DataObj(rng.normal(size=5), float(i % 2), None, None) for i in range(9)]
    a = util.handle_features(rows)
    b = ref.handle_features([ref.DataObj(*r) for r in rows])
    assert torch.equal(a.x_train, b.x_train) and torch.equal(a.y_train, b.y_train)
    c = util.handle_features(rows, validation_pct=0.34, seed=1)
    assert c.x_val.shape[0] == 3 and c.x_train.shape[0] == 6
    assert util.handle_features([]).x_train is None


@needs_ref
def test_early_stopping_matches_reference():
    ref = _load_ref("early_stopper")
    rng = np.random.default_rng(1)
    for mode in ("min", "max"):
        for pct in (False, True):
            for patience in (0, 1, 3):
                mine = EarlyStopping(mode=mode, min_delta=0.5, patience=patience, percentage=pct)
                theirs = ref.EarlyStopping(mode=mode, min_delta=0.5, patience=patience, percentage=pct)
                seq = list(rng.normal(1.0, 0.3, size=40))
                seq[25] = float("nan")
                for v in seq:
                    assert mine.step(v) == theirs.step(v)
    with pytest.raises(ValueError):
        EarlyStopping(mode="sideways")
    es = EarlyStopping(patience=2)
    es.step(1.0), es.step(2.0)
    vec = es.state_vector()
    es2 = EarlyStopping(patience=2)
    es2.load_state_vector(vec)
    assert es2.step(3.0) is True


def test_pipeline_carrier_codec():
    payload = {"x": list(range(50)), "s": "abc"}
    words = encode_stage(payload)
    assert words[1] == PY_OBJ_GUID == "4c1740b00d3c4ff6806a1402321572cb"
    # reference construction: ''.join(str(b)+',' for b in bytearray(zlib(dill(obj))))
    raw = bytes(int(t) for t in words[0].split(",")[:-1])
    assert dill.loads(zlib.decompress(raw)) == payload
    assert words[0].endswith(",")
    assert decode_stage(words[:-1]) == payload


def test_rwlock_semantics():
    lock = RWLock()
    order = []
    lock.acquire_read()
    lock.acquire_read()          # many readers

    def writer():
        lock.acquire_write()
        order.append("w")
        lock.release()

    def late_reader():
        lock.acquire_read()
        order.append("r")
        lock.release()

    tw = threading.Thread(target=writer)
    tw.start()
    time.sleep(0.05)
    tr = threading.Thread(target=late_reader)   # must queue behind the waiting writer
    tr.start()
    time.sleep(0.05)
    assert order == []
    lock.release()
    lock.release()
    tw.join(2), tr.join(2)
    assert order == ["w", "r"]
    with lock.writing():
        pass
    with lock.reading():
        pass


_REF_SAVE = r"""
import sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(1, sys.argv[3]); sys.path.append(sys.argv[2])
import numpy as np, torch, torch.nn as nn
from pyspark.sql import SparkSession
from pyspark.ml.linalg import Vectors
from pyspark.ml.pipeline import PipelineModel
from sparktorch import SparkTorch, serialize_torch_obj
from tests.nets import Regressor
spark = SparkSession.builder.master('local[1]').getOrCreate()
rng = np.random.default_rng(0)
df = spark.createDataFrame([(1.0, Vectors.dense(rng.normal(size=10))) for _ in range(20)], ['label', 'features'])
obj = serialize_torch_obj(Regressor(), nn.MSELoss(), torch.optim.Adam, lr=0.01)
stm = SparkTorch(inputCol='features', labelCol='label', predictionCol='predictions', torchObj=obj, iters=2, verbose=0).fit(df)
PipelineModel([stm]).write().overwrite().save(sys.argv[4])
print(repr(stm.transform(df).take(1)[0]['predictions']))
"""


@pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "sparktorch")),
                    reason="reference not installed under baseline/_ref")
def test_pipeline_saved_by_reference_loads_here(tmp_path, spark):
    """A PipelineModel written by the UNMODIFIED reference package is read back
    by this package (same carrier format, same class paths)."""
    from pyspark.ml.linalg import Vectors
    from pyspark.ml.pipeline import PipelineModel
    from sparktorch_b200 import PysparkPipelineWrapper
    from sparktorch_b200.torch_distributed import SparkTorchModel

    path = str(tmp_path / "ref_pipe")
    out = subprocess.run(
        [sys.executable, "-c", _REF_SAVE, os.path.join(ROOT, "baseline", "_ref"), os.path.join(ROOT, "sparkshim"), ROOT, path],
        capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert out.returncode == 0, out.stderr[-2000:]
    expect = float(out.stdout.strip().splitlines()[-1])
    pm = PysparkPipelineWrapper.unwrap(PipelineModel.load(path))
    assert isinstance(pm.stages[0], SparkTorchModel)
    rng = np.random.default_rng(0)
    df = spark.createDataFrame([(1.0, Vectors.dense(rng.normal(size=10))) for _ in range(20)], ["label", "features"])
    got = pm.transform(df).take(1)[0]["predictions"]
    assert abs(got - expect) < 1e-6
