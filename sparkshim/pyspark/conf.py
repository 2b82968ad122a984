class SparkConf:
    def __init__(self, entries=None):
        self._d = dict(entries or {})

    def set(self, k, v):
        self._d[k] = v
        return self

    def get(self, k, default=None):
        return self._d.get(k, default)

    def getAll(self):
        return list(self._d.items())
