"""Column expressions: ``udf``, ``col``, ``rand``, ``lit``."""
from __future__ import annotations

import random
from typing import Any, Callable


class Column:
    def __init__(self, fn: Callable[[Any], Any], name: str = "col", is_rand: bool = False, source: str = None):
        self._fn = fn          # Row -> value
        self._name = name
        self._is_rand = is_rand
        self._source = source  # set for plain column references: lets select() stay columnar

    def eval(self, row):
        return self._fn(row)

    def alias(self, name: str) -> "Column":
        return Column(self._fn, name, self._is_rand, self._source)

    def cast(self, dtype) -> "Column":
        conv = dtype.convert if hasattr(dtype, "convert") else (lambda v: v)
        return Column(lambda r: conv(self._fn(r)), self._name)


def _as_column(c) -> Column:
    if isinstance(c, Column):
        return c
    return col(c)


def col(name: str) -> Column:
    return Column(lambda r, name=name: r[name], name, source=name)


column = col


def lit(v) -> Column:
    return Column(lambda _r: v, "lit")


def rand(seed=None) -> Column:
    rng = random.Random(seed)
    return Column(lambda _r: rng.random(), "rand", is_rand=True)


class UserDefinedFunction:
    def __init__(self, f: Callable, returnType=None):
        self.func = f
        self.returnType = returnType

    def __call__(self, *cols) -> Column:
        cs = [_as_column(c) for c in cols]
        f, rt = self.func, self.returnType

        def run(row):
            out = f(*[c.eval(row) for c in cs])
            return rt.convert(out) if hasattr(rt, "convert") else out

        return Column(run, getattr(f, "__name__", "udf"))


def udf(f=None, returnType=None):
    if f is None or not callable(f):  # used as decorator with a type
        rt = f if f is not None else returnType
        return lambda fn: UserDefinedFunction(fn, rt)
    return UserDefinedFunction(f, returnType)
