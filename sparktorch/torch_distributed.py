"""Import-path compatibility: pickles written by SparkTorch reference
`sparktorch.torch_distributed`; the implementation lives in `sparktorch_b200.torch_distributed`."""
import sys as _sys

import sparktorch_b200.torch_distributed as _impl

_sys.modules[__name__] = _impl
