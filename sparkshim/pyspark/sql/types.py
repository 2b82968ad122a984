"""Row and the handful of SQL types used as UDF return types."""
from __future__ import annotations


class Row(tuple):
    """Tuple with field names: ``row['a']``, ``row.a``, ``row[0]``, ``'a' in row``."""

    def __new__(cls, *args, **kwargs):
        if args and kwargs:
            raise ValueError("Can not use both args and kwargs to create Row")
        if kwargs:
            row = tuple.__new__(cls, list(kwargs.values()))
            row.__fields__ = list(kwargs.keys())
            return row
        return tuple.__new__(cls, args)

    @classmethod
    def _make(cls, fields, values) -> "Row":
        row = tuple.__new__(cls, values)
        row.__fields__ = list(fields)
        return row

    def asDict(self):
        return dict(zip(self.__fields__, self))

    def __contains__(self, item):
        if hasattr(self, "__fields__"):
            return item in self.__fields__
        return super().__contains__(item)

    def __getitem__(self, item):
        if isinstance(item, (int, slice)):
            return super().__getitem__(item)
        try:
            return super().__getitem__(self.__fields__.index(item))
        except (ValueError, AttributeError):
            raise ValueError(item)

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        try:
            return self[item]
        except ValueError:
            raise AttributeError(item)

    def __reduce__(self):
        if hasattr(self, "__fields__"):
            return (Row._make, (self.__fields__, tuple(self)))
        return tuple.__reduce__(self)

    def __repr__(self):
        if hasattr(self, "__fields__"):
            return "Row(%s)" % ", ".join("%s=%r" % kv for kv in zip(self.__fields__, tuple(self)))
        return "<Row(%s)>" % ", ".join(repr(f) for f in self)


class DataType:
    def __eq__(self, other):
        return type(self) is type(other)

    def __hash__(self):
        return hash(type(self).__name__)

    def __repr__(self):
        return type(self).__name__ + "()"

    def convert(self, v):
        return v


class DoubleType(DataType):
    def convert(self, v):
        return None if v is None else float(v)


class FloatType(DoubleType):
    pass


class IntegerType(DataType):
    def convert(self, v):
        return None if v is None else int(v)


class LongType(IntegerType):
    pass


class StringType(DataType):
    def convert(self, v):
        return None if v is None else str(v)


class BooleanType(DataType):
    pass


class ArrayType(DataType):
    def __init__(self, elementType=None, containsNull=True):
        self.elementType = elementType
