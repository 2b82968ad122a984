"""Alias package: ``import sparktorch`` gives the sparktorch_b200 implementation
under the reference's import paths (needed because dill pickles ``TorchObj`` and
``SparkTorchModel`` by reference)."""
from sparktorch_b200 import *  # noqa: F401,F403
from sparktorch_b200 import __all__, __version__  # noqa: F401
