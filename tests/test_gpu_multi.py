"""Launches tests/multi_gpu_check.py on all visible GPUs (needs >= 2)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_multi_gpu_peer_memory_paths():
    n = min(8, torch.cuda.device_count())
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
         "--master-port", "29611", os.path.join(ROOT, "tests", "multi_gpu_check.py")],
        capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert "MULTI_GPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_multi_gpu_step_kernel():
    """K-STEP across GPUs (tests/multi_gpu_kstep.py): loss history identical on every rank, first 40 losses against the
    kernel-per-op engine, parameters after 20 SGD steps, ragged launch split, gather + step graph, early stop."""
    n = min(8, torch.cuda.device_count())
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
         "--master-port", "29631", os.path.join(ROOT, "tests", "multi_gpu_kstep.py")],
        capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert "KSTEP_MULTI_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
