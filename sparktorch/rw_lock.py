"""Import-path compatibility: pickles written by SparkTorch reference
`sparktorch.rw_lock`; the implementation lives in `sparktorch_b200.rw_lock`."""
import sys as _sys

import sparktorch_b200.rw_lock as _impl

_sys.modules[__name__] = _impl
