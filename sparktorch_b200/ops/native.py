"""Loader for the native sm_100a extension.

The .so is built in-tree by ``sparktorch_b200.csrc.build`` (``__graft_entry__.
build()``); at import time we only *load* it – no JIT, no rebuild, so a GPU box
never spends time compiling.  ``load(required=True)`` raises if it is missing:
on a GPU box the CUDA path must be the one that runs, never a silent fallback.
"""
from __future__ import annotations

import importlib.util
import os
from typing import Optional

_mod = None
_err: Optional[str] = None


def lib_path() -> str:
    from sparktorch_b200.csrc.build import so_path

    return so_path()


def load(required: bool = False):
    global _mod, _err
    if _mod is not None:
        return _mod
    path = lib_path()
    if not os.path.exists(path):
        _err = f"{path} not built (run: python -m sparktorch_b200.csrc.build)"
    else:
        try:
            import torch  # noqa: F401  (libtorch must be loaded first)

            spec = importlib.util.spec_from_file_location("sparktorch_b200._C", path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)  # type: ignore[union-attr]
            _mod = mod
            return _mod
        except Exception as e:  # pragma: no cover
            _err = f"loading {path} failed: {e!r}"
    if required:
        raise ImportError("sparktorch_b200 native extension unavailable: " + str(_err))
    return None


def available() -> bool:
    return load() is not None


def error() -> Optional[str]:
    return _err
