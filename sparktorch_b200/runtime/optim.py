"""Optimizer description + reference update math on flat buffers.

The trainers never run ``torch.optim`` in the hot loop (reference:
distributed.py:198 ``optimizer.step()``, server.py:139).  The optimizer *class*
and kwargs that travel in the ``TorchObj`` envelope are parsed into an
:class:`OptSpec`; the update is applied to ONE flat fp32 parameter buffer – by
the fused all-reduce+update CUDA kernels on B200 (csrc/comm_ar.cu, csrc/mlp_step.cu) or
by :func:`apply_update_` below (CPU / oracle path, same math, used by the
kernel numerics tests).  Semantics follow ``torch.optim.{Adam,AdamW,SGD}``.
Anything else (``kind == 'other'``) falls back to stepping the user's optimizer
on views of the flat buffers.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, Optional

import torch

KIND_SGD, KIND_ADAM, KIND_ADAMW = 0, 1, 2


@dataclasses.dataclass
class OptSpec:
    kind: str = "adam"            # adam | adamw | sgd | other
    lr: float = 1e-3
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    weight_decay: float = 0.0
    momentum: float = 0.0
    dampening: float = 0.0
    nesterov: bool = False
    maximize: bool = False

    @property
    def kind_id(self) -> int:
        return {"sgd": KIND_SGD, "adam": KIND_ADAM, "adamw": KIND_ADAMW}[self.kind]

    @property
    def n_state(self) -> int:
        """Number of fp32 state buffers the size of the parameters."""
        if self.kind in ("adam", "adamw"):
            return 2
        if self.kind == "sgd":
            return 1 if self.momentum != 0 else 0
        return 0


def _torch_optim_base(opt_cls):
    """('adam'|'adamw'|'sgd'|None, overridden_step).  Classes are matched by
    module+name along the MRO, not by identity: dill re-creates torch optimizer
    classes *by value* when a TorchObj blob is decoded, so ``issubclass`` against
    the live ``torch.optim`` classes is False for them."""
    overridden = False
    for k in getattr(opt_cls, "__mro__", ()):
        mod = getattr(k, "__module__", "") or ""
        if mod.startswith("torch.optim"):
            if k.__name__ == "AdamW":
                return "adamw", overridden
            if k.__name__ == "Adam":
                return "adam", overridden
            if k.__name__ == "SGD":
                return "sgd", overridden
            return None, overridden
        if "step" in k.__dict__:
            overridden = True      # user subclass with its own update rule
    return None, overridden


def parse_optimizer(opt_cls, params: Optional[Dict]) -> OptSpec:
    p = dict(params or {})
    base, overridden = _torch_optim_base(opt_cls)
    if base is None or overridden:
        return OptSpec(kind="other")
    if p.get("foreach") is not None:
        p.pop("foreach")
    if p.get("fused") is not None:
        p.pop("fused")
    if base in ("adam", "adamw"):
        if p.get("amsgrad") or p.get("capturable") or p.get("differentiable") or p.get("decoupled_weight_decay"):
            return OptSpec(kind="other")
        if isinstance(p.get("lr", 0.0), torch.Tensor):
            return OptSpec(kind="other")
        b1, b2 = p.get("betas", (0.9, 0.999))
        return OptSpec(
            kind=base,
            lr=float(p.get("lr", 1e-3)), beta1=float(b1), beta2=float(b2), eps=float(p.get("eps", 1e-8)),
            weight_decay=float(p.get("weight_decay", 1e-2 if base == "adamw" else 0.0)),
            maximize=bool(p.get("maximize", False)),
        )
    return OptSpec(
        kind="sgd", lr=float(p.get("lr", 1e-3)), momentum=float(p.get("momentum", 0.0)),
        dampening=float(p.get("dampening", 0.0)), nesterov=bool(p.get("nesterov", False)),
        weight_decay=float(p.get("weight_decay", 0.0)), maximize=bool(p.get("maximize", False)),
    )


def apply_update_(spec: OptSpec, p: torch.Tensor, g: torch.Tensor, state: Dict[str, torch.Tensor], step: int) -> None:
    """In-place update of flat fp32 ``p`` with (already averaged) gradient ``g``.
    ``step`` is 1-based.  ``state`` holds 'm','v' (adam) or 'buf' (sgd)."""
    if spec.maximize:
        g = -g
    if spec.kind in ("adam", "adamw"):
        if spec.kind == "adamw":
            if spec.weight_decay:
                p.mul_(1.0 - spec.lr * spec.weight_decay)
        elif spec.weight_decay:
            g = g.add(p, alpha=spec.weight_decay)
        m, v = state["m"], state["v"]
        m.mul_(spec.beta1).add_(g, alpha=1.0 - spec.beta1)
        v.mul_(spec.beta2).addcmul_(g, g, value=1.0 - spec.beta2)
        bc1 = 1.0 - spec.beta1 ** step
        bc2 = 1.0 - spec.beta2 ** step
        denom = (v.sqrt() / math.sqrt(bc2)).add_(spec.eps)
        p.addcdiv_(m, denom, value=-spec.lr / bc1)
    elif spec.kind == "sgd":
        if spec.weight_decay:
            g = g.add(p, alpha=spec.weight_decay)
        if spec.momentum:
            buf = state["buf"]
            if step == 1:
                buf.copy_(g)
            else:
                buf.mul_(spec.momentum).add_(g, alpha=1.0 - spec.dampening)
            g = g.add(buf, alpha=spec.momentum) if spec.nesterov else buf
        p.add_(g, alpha=-spec.lr)
    else:  # pragma: no cover
        raise ValueError(spec.kind)


def new_state(spec: OptSpec, like: torch.Tensor) -> Dict[str, torch.Tensor]:
    if spec.kind in ("adam", "adamw"):
        return {"m": torch.zeros_like(like), "v": torch.zeros_like(like)}
    if spec.kind == "sgd" and spec.momentum:
        return {"buf": torch.zeros_like(like)}
    return {}
