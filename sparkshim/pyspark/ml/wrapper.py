"""The sliver of the py4j wrapper layer that custom-stage persistence uses:
``JavaParams._new_java_obj(java_class_name, uid)``."""
from __future__ import annotations

from pyspark.ml.param import Params
from pyspark.ml.util import _JAVA_CLASS_TO_PY, _jsonable, _write_metadata


class JavaParams(Params):
    """Python stages whose params are persisted through their (here: simulated)
    JVM twin.  ``_to_java()`` of a plain shim stage is the stage itself."""

    _java_class_name = "org.apache.spark.ml.Unknown"

    @staticmethod
    def _new_java_obj(java_class: str, *args):
        import pyspark.ml.feature  # noqa: F401  (registers classes)

        cls = _JAVA_CLASS_TO_PY[java_class]
        obj = cls()
        if args:
            obj.uid = args[0]
            obj._reparent()
        return obj

    def _reparent(self) -> None:
        pm, dm = self._paramMap, self._defaultParamMap
        self._paramMap, self._defaultParamMap = {}, {}
        self._params = None
        for p in list(pm) + list(dm):
            pass
        for name in dir(type(self)):
            from pyspark.ml.param import Param

            attr = getattr(type(self), name, None)
            if isinstance(attr, Param):
                setattr(self, name, attr._copy_new_parent(self))
        for p, v in pm.items():
            self._paramMap[self.getParam(p.name)] = v
        for p, v in dm.items():
            self._defaultParamMap[self.getParam(p.name)] = v

    def _to_java(self):
        return self

    def _save_params(self, path: str) -> None:
        _write_metadata(
            path,
            self._java_class_name,
            self.uid,
            {p.name: _jsonable(v) for p, v in self._paramMap.items()},
            {p.name: _jsonable(v) for p, v in self._defaultParamMap.items()},
        )

    @classmethod
    def _load_params(cls, path: str, meta: dict):
        obj = cls()
        obj.uid = meta["uid"]
        obj._reparent()
        for k, v in meta.get("defaultParamMap", {}).items():
            if obj.hasParam(k):
                obj._defaultParamMap[obj.getParam(k)] = v
        for k, v in meta.get("paramMap", {}).items():
            if obj.hasParam(k):
                obj._paramMap[obj.getParam(k)] = v
        return obj


class JavaTransformer(JavaParams):
    pass
