"""Param / Params / TypeConverters – the Spark ML parameter machinery
(``_setDefault``, ``_set``, ``getOrDefault``, ``_input_kwargs`` …)."""
from __future__ import annotations

import copy
import uuid
from typing import Any, Dict, List, Optional

import numpy as np


class Param:
    def __init__(self, parent, name: str, doc: str, typeConverter=None):
        self.parent = parent.uid if isinstance(parent, Identifiable) else parent
        self.name = str(name)
        self.doc = str(doc)
        self.typeConverter = TypeConverters.identity if typeConverter is None else typeConverter

    def _copy_new_parent(self, parent: "Params") -> "Param":
        if self.parent == "undefined":
            p = copy.copy(self)
            p.parent = parent.uid
            return p
        raise ValueError("Cannot copy from non-dummy parent %s." % self.parent)

    def __str__(self):
        return str(self.parent) + "__" + self.name

    def __repr__(self):
        return "Param(parent=%r, name=%r, doc=%r)" % (self.parent, self.name, self.doc)

    def __hash__(self):
        return hash(str(self))

    def __eq__(self, other):
        return isinstance(other, Param) and self.parent == other.parent and self.name == other.name


class TypeConverters:
    @staticmethod
    def identity(value):
        return value

    @staticmethod
    def toString(value):
        if isinstance(value, str):
            return value
        raise TypeError("Could not convert %s to string type" % type(value))

    @staticmethod
    def toInt(value):
        if isinstance(value, (bool, np.bool_)):
            raise TypeError("Could not convert %s to int" % value)
        if isinstance(value, (int, np.integer)) or (isinstance(value, float) and value.is_integer()):
            return int(value)
        raise TypeError("Could not convert %s to int" % value)

    @staticmethod
    def toFloat(value):
        if isinstance(value, (int, float, np.integer, np.floating)) and not isinstance(value, bool):
            return float(value)
        raise TypeError("Could not convert %s to float" % value)

    @staticmethod
    def toBoolean(value):
        if isinstance(value, (bool, np.bool_)):
            return bool(value)
        raise TypeError("Boolean Param requires value of type bool. Found %s." % type(value))

    @staticmethod
    def toList(value):
        if isinstance(value, (list, tuple, np.ndarray)):
            return list(value)
        raise TypeError("Could not convert %s to list" % value)

    @staticmethod
    def toListString(value):
        return [TypeConverters.toString(v) for v in TypeConverters.toList(value)]

    @staticmethod
    def toListFloat(value):
        return [TypeConverters.toFloat(v) for v in TypeConverters.toList(value)]

    @staticmethod
    def toListInt(value):
        return [TypeConverters.toInt(v) for v in TypeConverters.toList(value)]


class Identifiable:
    def __init__(self):
        self.uid = self._randomUID()

    @classmethod
    def _randomUID(cls) -> str:
        return str(cls.__name__ + "_" + uuid.uuid4().hex[-12:])

    def __repr__(self):
        return self.uid


class Params(Identifiable):
    def __init__(self):
        super().__init__()
        self._paramMap: Dict[Param, Any] = {}
        self._defaultParamMap: Dict[Param, Any] = {}
        self._params: Optional[List[Param]] = None
        self._copy_params()

    @staticmethod
    def _dummy():
        d = Params.__new__(Params)
        d.uid = "undefined"
        return d

    def _copy_params(self) -> None:
        cls = type(self)
        for name in dir(cls):
            if name.startswith("__"):
                continue
            attr = getattr(cls, name, None)
            if isinstance(attr, Param):
                setattr(self, name, attr._copy_new_parent(self))

    @property
    def params(self) -> List[Param]:
        if self._params is None:
            found = []
            for name in dir(self):
                if name in ("params",) or name.startswith("__"):
                    continue
                if isinstance(getattr(type(self), name, None), property):
                    continue
                attr = getattr(self, name, None)
                if isinstance(attr, Param):
                    found.append(attr)
            self._params = sorted(found, key=lambda p: p.name)
        return self._params

    def hasParam(self, paramName: str) -> bool:
        return isinstance(getattr(self, paramName, None), Param) if isinstance(paramName, str) else False

    def getParam(self, paramName: str) -> Param:
        p = getattr(self, paramName, None)
        if isinstance(p, Param):
            return p
        raise ValueError("Cannot find param with name %s." % paramName)

    def _resolveParam(self, param) -> Param:
        if isinstance(param, Param):
            if param.parent != self.uid:
                return self.getParam(param.name)
            return param
        return self.getParam(param)

    def isSet(self, param) -> bool:
        return self._resolveParam(param) in self._paramMap

    def hasDefault(self, param) -> bool:
        return self._resolveParam(param) in self._defaultParamMap

    def isDefined(self, param) -> bool:
        return self.isSet(param) or self.hasDefault(param)

    def getOrDefault(self, param):
        p = self._resolveParam(param)
        if p in self._paramMap:
            return self._paramMap[p]
        if p in self._defaultParamMap:
            return self._defaultParamMap[p]
        raise KeyError("Failed to find a default value for " + p.name)

    def getDefault(self, param):
        return self._defaultParamMap[self._resolveParam(param)]

    def extractParamMap(self, extra=None) -> Dict[Param, Any]:
        m = dict(self._defaultParamMap)
        m.update(self._paramMap)
        if extra:
            m.update(extra)
        return m

    def explainParams(self) -> str:
        return "\n".join("%s: %s" % (p.name, p.doc) for p in self.params)

    def set(self, param, value) -> None:
        p = self._resolveParam(param)
        self._paramMap[p] = p.typeConverter(value) if value is not None else None

    def clear(self, param) -> None:
        self._paramMap.pop(self._resolveParam(param), None)

    def _set(self, **kwargs):
        for name, value in kwargs.items():
            p = self.getParam(name)
            if value is not None:
                try:
                    value = p.typeConverter(value)
                except (ValueError, TypeError) as e:
                    raise TypeError('Invalid param value given for param "%s". %s' % (p.name, e))
            self._paramMap[p] = value
        return self

    def _setDefault(self, **kwargs):
        for name, value in kwargs.items():
            p = self.getParam(name)
            if value is not None:
                try:
                    value = p.typeConverter(value)
                except (ValueError, TypeError) as e:
                    raise TypeError('Invalid default param value given for param "%s". %s' % (p.name, e))
            self._defaultParamMap[p] = value
        return self

    def copy(self, extra=None):
        that = copy.copy(self)
        that._paramMap = dict(self._paramMap)
        that._defaultParamMap = dict(self._defaultParamMap)
        if extra:
            for p, v in extra.items():
                that._paramMap[that._resolveParam(p)] = v
        return that


__all__ = ["Param", "Params", "TypeConverters", "Identifiable"]
