"""Import-path compatibility: pickles written by SparkTorch reference
`sparktorch.pipeline_util`; the implementation lives in `sparktorch_b200.pipeline_util`."""
import sys as _sys

import sparktorch_b200.pipeline_util as _impl

_sys.modules[__name__] = _impl
