from pyspark.ml.base import Estimator, Model, Transformer
from pyspark.ml.pipeline import Pipeline, PipelineModel
from pyspark.ml import evaluation, feature, linalg, param, util, wrapper

__all__ = ["Estimator", "Model", "Transformer", "Pipeline", "PipelineModel", "evaluation", "feature", "linalg",
           "param", "util", "wrapper"]
