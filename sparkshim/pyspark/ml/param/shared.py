"""Shared single-param mixins."""
from pyspark.ml.param import Param, Params, TypeConverters


class HasInputCol(Params):
    inputCol = Param(Params._dummy(), "inputCol", "input column name.", typeConverter=TypeConverters.toString)

    def __init__(self):
        super().__init__()

    def getInputCol(self):
        return self.getOrDefault(self.inputCol)


class HasInputCols(Params):
    inputCols = Param(Params._dummy(), "inputCols", "input column names.", typeConverter=TypeConverters.toListString)

    def __init__(self):
        super().__init__()

    def getInputCols(self):
        return self.getOrDefault(self.inputCols)


class HasOutputCol(Params):
    outputCol = Param(Params._dummy(), "outputCol", "output column name.", typeConverter=TypeConverters.toString)

    def __init__(self):
        super().__init__()
        self._setDefault(outputCol=self.uid + "__output")

    def getOutputCol(self):
        return self.getOrDefault(self.outputCol)


class HasLabelCol(Params):
    labelCol = Param(Params._dummy(), "labelCol", "label column name.", typeConverter=TypeConverters.toString)

    def __init__(self):
        super().__init__()
        self._setDefault(labelCol="label")

    def getLabelCol(self):
        return self.getOrDefault(self.labelCol)


class HasPredictionCol(Params):
    predictionCol = Param(Params._dummy(), "predictionCol", "prediction column name.",
                          typeConverter=TypeConverters.toString)

    def __init__(self):
        super().__init__()
        self._setDefault(predictionCol="prediction")

    def getPredictionCol(self):
        return self.getOrDefault(self.predictionCol)


class HasFeaturesCol(Params):
    featuresCol = Param(Params._dummy(), "featuresCol", "features column name.", typeConverter=TypeConverters.toString)

    def __init__(self):
        super().__init__()
        self._setDefault(featuresCol="features")

    def getFeaturesCol(self):
        return self.getOrDefault(self.featuresCol)
