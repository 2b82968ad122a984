from pyspark.sql.types import Row
from pyspark.sql.dataframe import DataFrame
from pyspark.sql.session import SparkSession
from pyspark.sql import functions, types
from pyspark.sql.functions import Column

__all__ = ["Row", "DataFrame", "SparkSession", "functions", "types", "Column"]
