"""Import-path compatibility: pickles written by SparkTorch reference
`sparktorch.hogwild`; the implementation lives in `sparktorch_b200.hogwild`."""
import sys as _sys

import sparktorch_b200.hogwild as _impl

_sys.modules[__name__] = _impl
