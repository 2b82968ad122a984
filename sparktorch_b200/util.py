"""Serialisation envelope and tensor marshalling.

Wire formats are kept bit-compatible with the reference so that blobs travel in
both directions (reference: sparktorch/util.py:31-209):

* ``torchObj``  = ``json({"torch_obj": base64(dill(TorchObj)), "shapes": [[...]]})``
* ``modStr``    = ``base64(dill(nn.Module))``

``TorchObj``/``DataObj`` are published under the module path ``sparktorch.util``
(the ``sparktorch`` alias package re-exports this module) because dill pickles
namedtuple classes by reference.

What is new here compared to the reference:

* :func:`rows_to_block` / :func:`ingest_partition` – columnar ingest. A partition
  is turned into ONE contiguous float32 matrix (zero-copy when the launcher
  hands us a numpy block) instead of a Python list of per-row arrays
  (reference: torch_distributed.py:44-56 + util.py:58-101, quirk Q16).
* label dtype is decided once from the criterion (quirk Q3) by
  :func:`criterion_wants_class_index`.
"""
from __future__ import annotations

import base64
import collections
import json
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple, Type

import dill
import numpy as np
import torch
import torch.nn as nn

TorchObj = collections.namedtuple(
    "TorchObj",
    ["model", "criterion", "optimizer", "optimizer_params", "is_lazy", "model_parameters"],
)
DataObj = collections.namedtuple("DataObj", ["x_train", "y_train", "x_val", "y_val"])
# dill pickles namedtuple classes by reference -> keep the reference's import path.
TorchObj.__module__ = "sparktorch.util"
DataObj.__module__ = "sparktorch.util"


# --------------------------------------------------------------------------- codec
def torch_encoder(obj: Any) -> str:
    """dill -> base64 text with a newline every 76 chars (MIME style, like
    ``codecs.encode(..., 'base64')`` in the reference, util.py:38-46)."""
    return base64.encodebytes(dill.dumps(obj)).decode()


def torch_decoder(payload: str) -> Any:
    """Inverse of :func:`torch_encoder` (reference util.py:49-55)."""
    return dill.loads(base64.decodebytes(payload.encode()))


def _param_shapes(module: nn.Module) -> List[List[int]]:
    return [list(p.shape) for p in module.parameters()]


def _envelope(obj: TorchObj, shapes: List[List[int]]) -> str:
    return json.dumps({"torch_obj": torch_encoder(obj), "shapes": shapes})


def serialize_torch_obj(model: nn.Module, criterion: Any, optimizer: Type, **kwargs) -> str:
    """Eager envelope: the instantiated model/criterion travel in the blob and
    ``kwargs`` become the optimizer parameters (reference util.py:183-202)."""
    obj = TorchObj(model, criterion, optimizer, kwargs, False, None)
    return _envelope(obj, _param_shapes(model))


def serialize_torch_obj_lazy(
    model: Type[nn.Module],
    criterion: Type,
    optimizer: Type,
    optimizer_params: Optional[Dict] = None,
    model_parameters: Optional[Dict] = None,
) -> str:
    """Lazy envelope: classes + ctor kwargs travel, instances are built on the
    worker (reference util.py:149-180).  A throw-away instance is built here
    only to record the parameter shapes; on the meta device so no memory or
    RNG state is touched (quirk Q13)."""
    try:
        with torch.device("meta"):
            probe = model(**model_parameters) if model_parameters else model()
    except Exception:  # ctor does something meta tensors cannot do
        probe = model(**model_parameters) if model_parameters else model()
    obj = TorchObj(model, criterion, optimizer, optimizer_params, True, model_parameters)
    return _envelope(obj, _param_shapes(probe))


def load_base_torch(torch_obj: str) -> Tuple[str, List[List[int]]]:
    """Split the JSON envelope into (blob, shapes) (reference util.py:104-111)."""
    d = json.loads(torch_obj)
    return d["torch_obj"], d["shapes"]


def load_optimizer(optimizer: Type, params: Optional[Dict], model: nn.Module):
    """Instantiate the optimizer class on the model (reference util.py:205-209)."""
    return optimizer(model.parameters(), **params) if params else optimizer(model.parameters())


def load_torch_model(torch_obj: str, from_json: bool = False, build_optimizer: bool = True) -> TorchObj:
    """Decode a blob into live objects (reference util.py:114-146).

    ``build_optimizer=False`` keeps the optimizer *class* in the tuple; the
    fused engines own the optimizer state themselves and only need the class
    and its hyper-parameters.
    """
    if from_json:
        torch_obj, _ = load_base_torch(torch_obj)
    raw = torch_decoder(torch_obj)
    if raw.is_lazy:
        model = raw.model(**raw.model_parameters) if raw.model_parameters else raw.model()
        criterion = raw.criterion()
    else:
        model, criterion = raw.model, raw.criterion
    opt = load_optimizer(raw.optimizer, raw.optimizer_params, model) if build_optimizer else raw.optimizer
    return TorchObj(model, criterion, opt, raw.optimizer_params, False, None)


# --------------------------------------------------------------------------- data
def criterion_wants_class_index(criterion: Any) -> bool:
    """True when the loss expects integer class targets of shape [N].

    The reference discovers this by letting the first ``criterion(...)`` call
    raise and retrying with ``flatten().long()`` labels – every iteration when
    mini-batching (distributed.py:152-157, quirk Q3).  We decide once."""
    return isinstance(criterion, (nn.CrossEntropyLoss, nn.NLLLoss, nn.MultiMarginLoss))


def _label_to_row(y) -> np.ndarray:
    if isinstance(y, (int, float, np.integer, np.floating)):
        return np.asarray([y], dtype=np.float32)
    return np.asarray(y, dtype=np.float32).reshape(-1)


def rows_to_block(data: Iterable) -> Tuple[Optional[np.ndarray], Optional[np.ndarray]]:
    """Collect a partition into (X[n,d] float32, Y[n,k] float32 | None).

    Accepts an iterable of :class:`DataObj` (reference element format,
    torch_distributed.py:44-56) or a single pre-stacked ``DataObj`` whose
    ``x_train`` is already a 2-D block (our columnar launcher path)."""
    if isinstance(data, DataObj):
        data = [data]
    xs: List[np.ndarray] = []
    ys: List[np.ndarray] = []
    for item in data:
        x = np.asarray(item.x_train)
        if x.ndim >= 2:  # already a block of rows
            xs.append(x.reshape(x.shape[0], -1).astype(np.float32, copy=False))
            if item.y_train is not None:
                y = np.asarray(item.y_train, dtype=np.float32)
                ys.append(y.reshape(x.shape[0], -1))
        else:
            xs.append(x.astype(np.float32, copy=False)[None, :])
            if item.y_train is not None:
                ys.append(_label_to_row(item.y_train)[None, :])
    if not xs:
        return None, None
    X = xs[0] if len(xs) == 1 else np.concatenate(xs, axis=0)
    Y = None
    if ys:
        Y = ys[0] if len(ys) == 1 else np.concatenate(ys, axis=0)
    return np.ascontiguousarray(X), (np.ascontiguousarray(Y) if Y is not None else None)


def split_validation(n: int, validation_pct: float, rng: np.random.Generator):
    """Random train/validation index split (reference util.py:80-95)."""
    n_val = int(n * validation_pct)
    perm = rng.permutation(n)
    val_idx = np.sort(perm[:n_val])
    train_idx = np.sort(perm[n_val:])
    return train_idx, val_idx


def handle_features(data: Sequence, validation_pct: float = 0.0, seed: Optional[int] = None) -> DataObj:
    """list[DataObj] -> DataObj of float32 tensors, with an optional random
    validation split (API of reference util.py:58-101)."""
    X, Y = rows_to_block(data)
    if X is None:
        return DataObj(None, None, None, None)
    xt = torch.from_numpy(X)
    yt = torch.from_numpy(Y) if Y is not None else None
    if validation_pct > 0:
        tr, va = split_validation(X.shape[0], validation_pct, np.random.default_rng(seed))
        tr_t, va_t = torch.from_numpy(tr), torch.from_numpy(va)
        return DataObj(
            xt[tr_t], yt[tr_t] if yt is not None else None, xt[va_t], yt[va_t] if yt is not None else None
        )
    return DataObj(xt, yt, None, None)
