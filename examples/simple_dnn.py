"""MNIST MLP with SparkTorch on this framework: VectorAssembler -> SparkTorch inside
a Spark ML Pipeline, save/load through PysparkPipelineWrapper, accuracy evaluation.
Same flow and hyper-parameters as the reference example (examples/simple_dnn.py:12-66);
``--device cuda`` trains on the B200(s), one barrier task per GPU.

    python examples/make_synthetic_mnist.py            # once (no dataset download here)
    python examples/simple_dnn.py [--device cuda] [--partitions 2] [--iters 1000]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn as nn

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
    ap.add_argument("--iters", type=int, default=1000)
    ap.add_argument("--csv", default="examples/mnist_train.csv")
    ap.add_argument("--mode", default="synchronous")
    a = ap.parse_args()
    if not os.path.exists(a.csv):
        from examples.make_synthetic_mnist import make

        make(a.csv, rows=8192)
    spark = SparkSession.builder.appName("examples").master(f"local[{a.partitions}]").getOrCreate()
    df = spark.read.option("inferSchema", "true").csv(a.csv).orderBy(rand()).repartition(a.partitions)
    network = nn.Sequential(nn.Linear(784, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, 10))
    torch_obj = serialize_torch_obj(model=network, criterion=nn.CrossEntropyLoss(), optimizer=torch.optim.Adam, lr=0.001)
    vector_assembler = VectorAssembler(inputCols=df.columns[1:785], outputCol="features")
    spark_model = SparkTorch(inputCol="features", labelCol="_c0", predictionCol="predictions", torchObj=torch_obj,
                             iters=a.iters, verbose=0, miniBatch=256, earlyStopPatience=40, validationPct=0.2,
                             device=a.device, mode=a.mode)
    p = Pipeline(stages=[vector_assembler, spark_model]).fit(df)
    p.write().overwrite().save("simple_dnn")
    loaded_pipeline = PysparkPipelineWrapper.unwrap(PipelineModel.load("simple_dnn"))
    predictions = loaded_pipeline.transform(df).persist()
    evaluator = MulticlassClassificationEvaluator(labelCol="_c0", predictionCol="predictions", metricName="accuracy")
    print("history:", {k: v for k, v in p.stages[-1].getTrainingHistory().items() if k != "train_loss"})
    print("Train accuracy = %g" % evaluator.evaluate(predictions))
