"""ptxas resource report of the hot kernels: no spills and no local-memory stack frames
(an accumulator array that lands in local memory costs ~4 us per tiny GEMM)."""
import glob
import os
import re

import pytest

BUILD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sparktorch_b200", "csrc", "_build")


@pytest.mark.skipif(not glob.glob(os.path.join(BUILD, "*.ptxas.txt")), reason="extension not built here")
def test_gemm_kernels_have_no_stack_frame_or_spills():
    txt = open(os.path.join(BUILD, "gemm_sm100.cu.o.ptxas.txt")).read()
    frames = [int(x) for x in re.findall(r"(\d+) bytes stack frame", txt)]
    spills = [int(x) for x in re.findall(r"(\d+) bytes spill stores", txt)]
    assert frames and max(frames) == 0, f"stack frames: {sorted(set(frames))}"
    assert max(spills) == 0
    for other in ("comm_ar.cu.o.ptxas.txt", "loss.cu.o.ptxas.txt", "data.cu.o.ptxas.txt"):
        t = open(os.path.join(BUILD, other)).read()
        assert max(int(x) for x in re.findall(r"(\d+) bytes spill stores", t)) == 0
