"""ResNet-50 (lazy envelope, synthetic 3x224x224 rows) through SparkTorch.fit on N GPUs.

--engine fused (default): every convolution is a tcgen05 GEMM over channels-last bf16 activations, batch-norm /
pooling / loss are the kernels of csrc/norm.cu, the whole step is one CUDA graph and the gradient exchange +
optimizer is the fused K-AR kernel (runtime/fused_resnet.py).  --engine eager: torch autograd with cuDNN
convolutions under bf16 autocast + the same K-AR kernel (the library baseline on the same box).

    python bench/resnet50_bench.py --gpus 1 --steps 10 --warmup 3 [--batch 32] [--engine fused|eager]
    python -m torch.distributed.run --nproc-per-node 8 ... bench/resnet50_bench.py --gpus 8
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def reference_arm(a):
    import glob
    import tempfile

    import numpy as np
    import torch
    import torch.nn as nn

    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    sys.path.insert(0, ref_dir)
    import sparktorch_b200  # noqa: F401  (bundled pyspark runtime)
    from baseline.ref_harness import TimedSGD, install_backend_patch

    for k in [k for k in sys.modules if k == "sparktorch" or k.startswith("sparktorch.")]:
        del sys.modules[k]
    install_backend_patch()
    out_dir = tempfile.mkdtemp(prefix="refresnet_")
    from pyspark.sql import SparkSession

    spark = SparkSession.builder.master(f"local[{a.gpus}]").appName("resnet50-reference").getOrCreate()
    import sparktorch
    from sparktorch import SparkTorch, serialize_torch_obj_lazy
    from sparktorch_b200.models import ResNet50

    assert os.path.realpath(sparktorch.__file__).startswith(os.path.realpath(ref_dir)), sparktorch.__file__
    n = a.batch * a.gpus
    rng = np.random.default_rng(0)
    x = rng.standard_normal((n, 3 * a.image * a.image), dtype=np.float32)
    y = rng.integers(0, 1000, size=n).astype(np.float64)
    df = spark.createDataFrame({"label": y, "features": x.astype(np.float64)}, numPartitions=a.gpus)
    obj = serialize_torch_obj_lazy(ResNet50, nn.CrossEntropyLoss, TimedSGD,
                                   {"lr": 0.1, "momentum": 0.9, "warmup": a.warmup, "steps": a.steps, "out_dir": out_dir},
                                   {"num_classes": 1000, "image_size": a.image})
    est = SparkTorch(inputCol="features", labelCol="label", predictionCol="p", torchObj=obj, iters=a.warmup + a.steps + 1,
                     verbose=0, mode="synchronous", device=a.device, miniBatch=-1, earlyStopPatience=-1)
    est.fit(df)
    recs = [json.load(open(f)) for f in sorted(glob.glob(os.path.join(out_dir, "ref_timing_rank*.json")))]
    ms = max(r["device_ms"] for r in recs)
    print(json.dumps({"metric": "ResNet-50 samples/sec (whole job)", "impl": "reference", "value": n * a.steps / (ms / 1e3),
                      "unit": "samples/s", "n_gpus": a.gpus, "ms_per_step": ms / a.steps, "global_batch": n,
                      "dtype": "fp32 (the reference has no mixed precision)", "data": "synthetic",
                      "note": "unmodified reference loop, full batch per partition, gloo->nccl"}), flush=True)
    spark.stop()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--image", type=int, default=224)
    ap.add_argument("--engine", default="fused")
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"],
                    help="reference: the unmodified reference package (baseline/_ref) through its own SparkTorch.fit, fp32, NCCL")
    a = ap.parse_args()
    if a.impl == "reference":
        return reference_arm(a)
    import numpy as np
    import torch
    import torch.nn as nn

    import sparktorch_b200  # noqa: F401
    from pyspark.sql import SparkSession
    from sparktorch_b200 import SparkTorch, serialize_torch_obj_lazy
    from sparktorch_b200.models import ResNet50

    spark = SparkSession.builder.master(f"local[{a.gpus}]").appName("resnet50").getOrCreate()
    n = a.batch * a.gpus
    rng = np.random.default_rng(0)
    x = rng.standard_normal((n, 3 * a.image * a.image), dtype=np.float32)
    y = rng.integers(0, 1000, size=n).astype(np.float64)
    df = spark.createDataFrame({"label": y, "features": x}, numPartitions=a.gpus)
    obj = serialize_torch_obj_lazy(ResNet50, nn.CrossEntropyLoss, torch.optim.SGD, {"lr": 0.1, "momentum": 0.9},
                                   {"num_classes": 1000, "image_size": a.image})
    est = SparkTorch(inputCol="features", labelCol="label", predictionCol="p", torchObj=obj, iters=a.warmup + a.steps,
                     device="cuda", precision="bf16", engine=a.engine, seed=0)
    est._time_window = (a.warmup, a.steps)
    h = est.fit(df).getTrainingHistory()
    ms = h["window_ms"]["device_ms"]
    print(json.dumps({"metric": "ResNet-50 samples/sec (whole job)", "value": n * a.steps / (ms / 1e3), "unit": "samples/s",
                      "n_gpus": a.gpus, "ms_per_step": ms / a.steps, "global_batch": n, "engine": h["engine"], "comm": h["comm"],
                      "gpu_launches_per_step": h.get("gpu_launches_per_step"),
                      "dtype": "bf16 activations/weights, fp32 accumulate + fp32 master/K-AR", "data": "synthetic"}), flush=True)
    spark.stop()


if __name__ == "__main__":
    main()
