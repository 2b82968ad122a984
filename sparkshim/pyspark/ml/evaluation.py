"""MulticlassClassificationEvaluator (accuracy / f1-free subset)."""
from pyspark import keyword_only
from pyspark.ml.param import Param, Params, TypeConverters
from pyspark.ml.param.shared import HasLabelCol, HasPredictionCol


class Evaluator(Params):
    def evaluate(self, dataset, params=None):
        return self._evaluate(dataset)


class MulticlassClassificationEvaluator(Evaluator, HasLabelCol, HasPredictionCol):
    metricName = Param(Params._dummy(), "metricName", "metric name", typeConverter=TypeConverters.toString)

    @keyword_only
    def __init__(self, *, predictionCol="prediction", labelCol="label", metricName="f1"):
        super().__init__()
        self._setDefault(metricName="f1")
        self._set(**self._input_kwargs)

    def _evaluate(self, dataset):
        pred, lab = self.getPredictionCol(), self.getLabelCol()
        metric = self.getOrDefault(self.metricName)
        rows = dataset.select(pred, lab).collect()
        if not rows:
            return 0.0
        if metric != "accuracy":
            raise NotImplementedError("sparkshim evaluator supports metricName='accuracy'")
        return sum(1 for r in rows if float(r[0]) == float(r[1])) / len(rows)
