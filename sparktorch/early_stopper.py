"""Import-path compatibility: pickles written by SparkTorch reference
`sparktorch.early_stopper`; the implementation lives in `sparktorch_b200.early_stopper`."""
import sys as _sys

import sparktorch_b200.early_stopper as _impl

_sys.modules[__name__] = _impl
