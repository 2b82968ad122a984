"""Writes a synthetic stand-in for the reference's examples/mnist_train.csv
(42 000 rows x 785 columns, label first, no header; this image has no network).
Digits are class-dependent blob patterns so that the example nets can learn."""
import sys

import numpy as np


def make(path: str = "examples/mnist_train.csv", rows: int = 42000, seed: int = 0) -> str:
    rng = np.random.default_rng(seed)
    protos = (rng.random((10, 28, 28)) < 0.18).astype(np.float32)
    labels = rng.integers(0, 10, size=rows)
    img = protos[labels] * rng.integers(128, 256, size=(rows, 28, 28)) * (rng.random((rows, 28, 28)) < 0.9)
    data = np.concatenate([labels[:, None], img.reshape(rows, -1)], axis=1).astype(np.int64)
    np.savetxt(path, data, fmt="%d", delimiter=",")
    return path


if __name__ == "__main__":
    print(make(*(sys.argv[1:2] or ["examples/mnist_train.csv"]), rows=int(sys.argv[2]) if len(sys.argv) > 2 else 42000))
