"""Helpers for inference-only use (reference sparktorch/inference.py:10-61)."""
from __future__ import annotations

import torch.nn as nn

from sparktorch_b200._spark import ensure_pyspark

ensure_pyspark()

from pyspark.ml.pipeline import PipelineModel  # noqa: E402

from sparktorch_b200.torch_distributed import SparkTorchModel  # noqa: E402
from sparktorch_b200.util import torch_encoder  # noqa: E402


def convert_to_serialized_torch(network: nn.Module) -> str:
    """``modStr`` = base64(dill(module)) – the whole module, not a state_dict."""
    return torch_encoder(network)


def create_spark_torch_model(network: nn.Module, inputCol: str = "features", predictionCol: str = "predicted",
                             useVectorOut: bool = False) -> SparkTorchModel:
    """Wrap an already trained network as a Spark ML transformer."""
    return SparkTorchModel(inputCol=inputCol, predictionCol=predictionCol,
                           modStr=convert_to_serialized_torch(network), useVectorOut=useVectorOut)


def attach_pytorch_model_to_pipeline(network: nn.Module, pipeline_model: PipelineModel, inputCol: str = "features",
                                     predictionCol: str = "predicted", useVectorOut: bool = False) -> PipelineModel:
    """``PipelineModel([existing_pipeline_model, SparkTorchModel(network)])``."""
    return PipelineModel(stages=[pipeline_model, create_spark_torch_model(network, inputCol, predictionCol, useVectorOut)])
