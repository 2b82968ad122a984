"""Estimator / Transformer / Model bases."""
from pyspark.ml.param import Params


class Estimator(Params):
    def _fit(self, dataset):
        raise NotImplementedError()

    def fit(self, dataset, params=None):
        if isinstance(params, (list, tuple)):
            return [self.fit(dataset, p) for p in params]
        if params:
            return self.copy(params)._fit(dataset)
        return self._fit(dataset)


class Transformer(Params):
    def _transform(self, dataset):
        raise NotImplementedError()

    def transform(self, dataset, params=None):
        if params:
            return self.copy(params)._transform(dataset)
        return self._transform(dataset)


class Model(Transformer):
    pass
