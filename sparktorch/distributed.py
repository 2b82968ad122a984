"""Import-path compatibility: pickles written by SparkTorch reference
`sparktorch.distributed`; the implementation lives in `sparktorch_b200.distributed`."""
import sys as _sys

import sparktorch_b200.distributed as _impl

_sys.modules[__name__] = _impl
