"""Import-path compatibility: pickles written by SparkTorch reference
`sparktorch.inference`; the implementation lives in `sparktorch_b200.inference`."""
import sys as _sys

import sparktorch_b200.inference as _impl

_sys.modules[__name__] = _impl
