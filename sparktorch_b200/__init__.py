"""sparktorch_b200 – a Blackwell-native re-implementation of SparkTorch's
capabilities (Spark ML Estimator/Transformer for distributed PyTorch training
and inference) – see DESIGN.md.  Public names match the reference package
(sparktorch/__init__.py:1-4).

The public names are resolved lazily (PEP 562): importing a sub-module such as
``sparktorch_b200.csrc.build`` or ``sparktorch_b200.ops.native`` does not pull in
the Spark layer, so the extension can be built and the kernels used on a box
without any Spark API on ``sys.path``."""
import importlib

from sparktorch_b200._spark import locate_pyspark

__version__ = "0.2.0"
locate_pyspark()          # sys.path only; nothing of Spark is imported here

_LAZY = {
    "serialize_torch_obj": "sparktorch_b200.util",
    "serialize_torch_obj_lazy": "sparktorch_b200.util",
    "SparkTorch": "sparktorch_b200.torch_distributed",
    "SparkTorchModel": "sparktorch_b200.torch_distributed",
    "PysparkPipelineWrapper": "sparktorch_b200.pipeline_util",
    "attach_pytorch_model_to_pipeline": "sparktorch_b200.inference",
    "create_spark_torch_model": "sparktorch_b200.inference",
}
__all__ = ["serialize_torch_obj", "serialize_torch_obj_lazy", "SparkTorch", "SparkTorchModel",
           "PysparkPipelineWrapper", "create_spark_torch_model", "attach_pytorch_model_to_pipeline"]


def __getattr__(name):
    target = _LAZY.get(name)
    if target is None:
        raise AttributeError(f"module 'sparktorch_b200' has no attribute {name!r}")
    from sparktorch_b200._spark import ensure_pyspark

    ensure_pyspark()
    value = getattr(importlib.import_module(target), name)
    globals()[name] = value
    return value


def __dir__():
    return sorted(list(globals()) + list(_LAZY))
