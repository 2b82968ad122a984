"""Import-path compatibility: pickles written by SparkTorch reference
`sparktorch.server`; the implementation lives in `sparktorch_b200.server`."""
import sys as _sys

import sparktorch_b200.server as _impl

_sys.modules[__name__] = _impl
