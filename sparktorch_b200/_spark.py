"""Locate a Spark API: the real ``pyspark`` if importable, else the in-repo
single-node runtime (``sparkshim/pyspark``), which implements the same contract
 with one executor process per partition/GPU."""
from __future__ import annotations

import importlib.util
import os
import sys

_SHIM_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sparkshim")


def locate_pyspark() -> None:
    """Make ``import pyspark`` resolvable (real package, else the in-repo runtime) WITHOUT importing it: only sys.path is
    touched, so ``import sparktorch_b200`` stays cheap and scripts that import pyspark right after it keep working."""
    if "pyspark" in sys.modules:
        return
    if importlib.util.find_spec("pyspark") is None and _SHIM_DIR not in sys.path:
        sys.path.append(_SHIM_DIR)


def ensure_pyspark() -> str:
    """Returns 'pyspark' or 'sparkshim'."""
    if "pyspark" in sys.modules:
        return "sparkshim" if getattr(sys.modules["pyspark"], "__shim__", False) else "pyspark"
    if importlib.util.find_spec("pyspark") is None:
        if _SHIM_DIR not in sys.path:
            sys.path.append(_SHIM_DIR)
    import pyspark

    return "sparkshim" if getattr(pyspark, "__shim__", False) else "pyspark"
