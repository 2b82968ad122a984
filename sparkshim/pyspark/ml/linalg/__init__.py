"""DenseVector / Vectors / VectorUDT."""
from __future__ import annotations

import numpy as np

from pyspark.sql.types import DataType


class DenseVector:
    __slots__ = ("array",)

    def __init__(self, ar):
        if isinstance(ar, DenseVector):
            ar = ar.array
        self.array = np.asarray(ar, dtype=np.float64).reshape(-1)

    def toArray(self) -> np.ndarray:
        return self.array

    @property
    def values(self) -> np.ndarray:
        return self.array

    @property
    def size(self) -> int:
        return len(self.array)

    def __len__(self) -> int:
        return len(self.array)

    def __getitem__(self, i):
        return self.array[i]

    def __iter__(self):
        return iter(self.array)

    def __eq__(self, other):
        if isinstance(other, DenseVector):
            return np.array_equal(self.array, other.array)
        return False

    def __ne__(self, other):
        return not self == other

    def __hash__(self):
        return hash(self.array.tobytes())

    def __reduce__(self):
        return (DenseVector, (self.array,))

    def __repr__(self):
        return "DenseVector([%s])" % ", ".join("%g" % v for v in self.array)

    def dot(self, other):
        return float(np.dot(self.array, np.asarray(other.toArray() if hasattr(other, "toArray") else other)))

    def norm(self, p):
        return float(np.linalg.norm(self.array, p))


Vector = DenseVector


class Vectors:
    @staticmethod
    def dense(*elements) -> DenseVector:
        if len(elements) == 1 and not isinstance(elements[0], (float, int)):
            elements = elements[0]
        return DenseVector(elements)


class VectorUDT(DataType):
    def convert(self, v):
        return v if isinstance(v, DenseVector) or v is None else DenseVector(v)
