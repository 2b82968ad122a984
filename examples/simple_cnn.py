"""MNIST CNN (conv5x5 -> conv3x3 -> maxpool -> dropout2d -> fc) through the same
pipeline flow (reference examples/simple_cnn.py:13-62).  ``--mode hogwild`` trains
against the parameter server instead of the synchronous all-reduce."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn as nn

from examples.cnn_network import Net
from sparktorch import PysparkPipelineWrapper, SparkTorch, serialize_torch_obj
from pyspark.ml.evaluation import MulticlassClassificationEvaluator
from pyspark.ml.feature import VectorAssembler
from pyspark.ml.pipeline import Pipeline, PipelineModel
from pyspark.sql import SparkSession
from pyspark.sql.functions import rand

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--partitions", type=int, default=2)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--mode", default="synchronous")
    ap.add_argument("--csv", default="examples/mnist_train.csv")
    a = ap.parse_args()
    if not os.path.exists(a.csv):
        from examples.make_synthetic_mnist import make

        make(a.csv, rows=4096)
    spark = SparkSession.builder.appName("examples").master(f"local[{a.partitions}]").getOrCreate()
    df = spark.read.option("inferSchema", "true").csv(a.csv).orderBy(rand()).repartition(a.partitions)
    torch_obj = serialize_torch_obj(model=Net(), criterion=nn.CrossEntropyLoss(), optimizer=torch.optim.Adam, lr=0.001)
    vector_assembler = VectorAssembler(inputCols=df.columns[1:785], outputCol="features")
    spark_model = SparkTorch(inputCol="features", labelCol="_c0", predictionCol="predictions", torchObj=torch_obj,
                             iters=a.iters, verbose=1, validationPct=0.2, miniBatch=128, device=a.device, mode=a.mode,
                             port=0)
    p = Pipeline(stages=[vector_assembler, spark_model]).fit(df)
    p.write().overwrite().save("simple_cnn")
    loaded_pipeline = PysparkPipelineWrapper.unwrap(PipelineModel.load("simple_cnn"))
    predictions = loaded_pipeline.transform(df).persist()
    evaluator = MulticlassClassificationEvaluator(labelCol="_c0", predictionCol="predictions", metricName="accuracy")
    print("Train accuracy = %g" % evaluator.evaluate(predictions))
