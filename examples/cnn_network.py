"""Importable CNN for the examples (custom modules must be importable on the
executors because dill pickles classes by reference)."""
from sparktorch_b200.models import MnistCNN as Net  # noqa: F401
