"""Packaging (reference: setup.py:1-17).  The CUDA extension is built in-tree by
``sparktorch_b200.csrc.build`` (sm_100a only) – see ``make build``."""
from setuptools import find_packages, setup

setup(
    name="sparktorch-b200",
    version="0.2.0",
    description="Blackwell-native distributed training/inference of PyTorch networks with the SparkTorch API",
    # sparkshim = the single-node PySpark-compatible runtime used when no real pyspark is importable
    # (sparktorch_b200/_spark.py appends <site-packages>/sparkshim to sys.path in that case)
    packages=find_packages(include=["sparktorch", "sparktorch.*", "sparktorch_b200", "sparktorch_b200.*", "sparkshim", "sparkshim.*"]),
    package_data={"sparktorch_b200": ["_C*.so", "csrc/*.cu", "csrc/*.cuh", "csrc/*.h", "csrc/*.cpp"]},
    install_requires=["torch", "dill", "numpy", "cloudpickle"],
    extras_require={"spark": ["pyspark>=3.2"]},
    python_requires=">=3.10",
    zip_safe=False,
)
