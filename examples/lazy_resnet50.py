"""ResNet-50 through the lazy envelope (the fourth workload of BASELINE.json): only the class and its
constructor kwargs travel to the executors, every barrier task binds one GPU, the step runs as a CUDA graph of
the fused residual-network plan (runtime/fused_resnet.py) and gradients meet in the K-AR kernel.

    python examples/lazy_resnet50.py --device cuda --partitions 1 --image 64 --rows 256 --iters 20
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/lazy_resnet50.py --device cuda --partitions 8

Synthetic rows (there is no dataset in the sandbox): flat 3*H*W feature vectors, labels drawn from a fixed random
linear teacher so that the loss can actually go down.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.nn as nn

import sparktorch_b200  # noqa: F401  (puts the bundled single-node pyspark runtime on the path when Spark is absent)
from pyspark.sql import SparkSession
from sparktorch import SparkTorch, serialize_torch_obj_lazy
from sparktorch_b200.models import ResNet50

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--partitions", type=int, default=1)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--image", type=int, default=32)
    ap.add_argument("--classes", type=int, default=10)
    ap.add_argument("--rows", type=int, default=128, help="rows per partition")
    ap.add_argument("--batch", type=int, default=32)
    a = ap.parse_args()
    spark = SparkSession.builder.appName("lazy_resnet50").master(f"local[{a.partitions}]").getOrCreate()
    rng = np.random.default_rng(0)
    n, d = a.rows * a.partitions, 3 * a.image * a.image
    x = rng.standard_normal((n, d), dtype=np.float32)
    teacher = rng.standard_normal((d, a.classes), dtype=np.float32)
    y = (x @ teacher).argmax(1).astype(np.float64)
    df = spark.createDataFrame({"label": y, "features": x}, numPartitions=a.partitions)
    torch_obj = serialize_torch_obj_lazy(model=ResNet50, criterion=nn.CrossEntropyLoss, optimizer=torch.optim.SGD,
                                         optimizer_params={"lr": 0.005, "momentum": 0.9},
                                         model_parameters={"num_classes": a.classes, "image_size": a.image})
    est = SparkTorch(inputCol="features", labelCol="label", predictionCol="predictions", torchObj=torch_obj, iters=a.iters,
                     miniBatch=a.batch, device=a.device, verbose=0, seed=0)
    model = est.fit(df)
    hist = model.getTrainingHistory()
    print("engine:", hist.get("engine"), "comm:", hist.get("comm"), "launches/step:", hist.get("gpu_launches_per_step"))
    tl = hist.get("train_loss") or []
    if tl:
        print("loss: first %.4f  last %.4f" % (tl[0], tl[-1]))
    pred = model.transform(df).select("label", "predictions").collect()
    acc = float(np.mean([float(r[0]) == float(r[1]) for r in pred]))
    print("Train accuracy = %g" % acc)
    spark.stop()
