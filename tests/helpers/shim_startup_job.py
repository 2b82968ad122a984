"""One rank of a torchrun job that only brings the Spark stand-in up (rank 0 = driver, the others become executors), runs one
barrier stage and leaves.  Used by tests/test_shim_startup.py."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import sparktorch_b200  # noqa: E402,F401  (puts the pyspark stand-in on sys.path when no real pyspark is installed)

t0 = time.time()
from pyspark.sql import SparkSession  # noqa: E402

n = int(os.environ["WORLD_SIZE"])
spark = SparkSession.builder.master(f"local[{n}]").appName("startup").getOrCreate()
out = spark.sparkContext.parallelize(list(range(n * 4)), n).barrier().mapPartitions(lambda it: [sum(it)]).collect()
assert sum(out) == sum(range(n * 4)), out
print("SHIM_OK", round(time.time() - t0, 2), flush=True)
spark.stop()
