"""SparkTorchModel.transform throughput (rows/s), both arms, MNIST-DNN shaped rows.

The reference scores one ROW per Python UDF call (batch size 1, ``model.eval()`` per row, CPU;
torch_distributed.py:97-130); this package scores whole partitions in large batches on the executor's device
(fused tcgen05 forward kernels for recognised nets).  Wall clock of ``transform(df).count()``-equivalent
(all predictions materialised on the driver), synthetic rows.

    python bench/inference_bench.py --impl ours --rows 200000 --device cuda
    python bench/inference_bench.py --impl reference --rows 20000
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=100000)
    ap.add_argument("--device", default="auto")
    ap.add_argument("--partitions", type=int, default=1)
    a = ap.parse_args()
    import numpy as np
    import torch

    import sparktorch_b200  # noqa: F401  (bundled pyspark runtime)
    from sparktorch_b200.models import mnist_dnn

    if a.impl == "reference":
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        for k in [k for k in sys.modules if k == "sparktorch" or k.startswith("sparktorch.")]:
            del sys.modules[k]
        import sparktorch as st                       # the unmodified reference package
        assert "baseline/_ref" in st.__file__, st.__file__
        from sparktorch import create_spark_torch_model
    else:
        from sparktorch_b200 import create_spark_torch_model
    from pyspark.sql import SparkSession

    spark = SparkSession.builder.master(f"local[{a.partitions}]").appName("inference").getOrCreate()
    rng = np.random.default_rng(0)
    x = rng.standard_normal((a.rows, 784), dtype=np.float32)
    df = spark.createDataFrame({"features": x}, numPartitions=a.partitions)
    torch.manual_seed(0)
    net = mnist_dnn()
    model = create_spark_torch_model(net, inputCol="features", predictionCol="predictions")
    if a.impl == "ours" and a.device != "auto":
        model = model.setParams(device=a.device) if hasattr(model, "setParams") else model
    # warm-up on a small slice (module import, kernel load), then the timed pass
    small = spark.createDataFrame({"features": x[:256]}, numPartitions=1)
    model.transform(small).select("predictions").collect()
    t0 = time.time()
    out = model.transform(df).select("predictions").collect()
    dt = time.time() - t0
    assert len(out) == a.rows
    print(json.dumps({"metric": "transform rows/s (MNIST-DNN rows, wall clock incl. driver collect)", "impl": a.impl,
                      "rows": a.rows, "seconds": dt, "value": a.rows / dt, "unit": "rows/s", "partitions": a.partitions,
                      "device": a.device if a.impl == "ours" else "cpu (reference ignores devices at inference)"}), flush=True)
    spark.stop()


if __name__ == "__main__":
    main()
