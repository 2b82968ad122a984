"""Import-path compatibility: pickles written by SparkTorch reference
`sparktorch.util`; the implementation lives in `sparktorch_b200.util`."""
import sys as _sys

import sparktorch_b200.util as _impl

_sys.modules[__name__] = _impl
