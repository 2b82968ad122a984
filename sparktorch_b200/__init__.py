"""sparktorch_b200 – a Blackwell-native re-implementation of SparkTorch's
capabilities (Spark ML Estimator/Transformer for distributed PyTorch training
and inference) – see DESIGN.md.  Public names match the reference package
(sparktorch/__init__.py:1-4)."""
__version__ = "0.1.0"

from sparktorch_b200._spark import ensure_pyspark

ensure_pyspark()

from sparktorch_b200.util import serialize_torch_obj, serialize_torch_obj_lazy  # noqa: E402
from sparktorch_b200.torch_distributed import SparkTorch, SparkTorchModel  # noqa: E402
from sparktorch_b200.pipeline_util import PysparkPipelineWrapper  # noqa: E402
from sparktorch_b200.inference import attach_pytorch_model_to_pipeline, create_spark_torch_model  # noqa: E402

__all__ = ["serialize_torch_obj", "serialize_torch_obj_lazy", "SparkTorch", "SparkTorchModel",
           "PysparkPipelineWrapper", "create_spark_torch_model", "attach_pytorch_model_to_pipeline"]
