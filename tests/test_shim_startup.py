"""Start-up of the Spark stand-in under torchrun with many ranks.

Regression test for the intermittent "8-GPU hang" of round 1 / round 2 (profiles/hang_8gpu_startup.md): every executor rank reads
the driver's port from the store and connects in the same instant; multiprocessing's Listener has a listen backlog of ONE, the
kernel drops the overflowing handshakes, and a client whose final ACK was dropped waits for an authentication challenge for ever
while the driver sits in accept().  The old code hangs 5 out of 5 times with 32 ranks on this box."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _run(world: int, port: int, extra_env=None, timeout=240):
    env = dict(os.environ, **(extra_env or {}))
    env.pop("SPARKSHIM_MASTER", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "helpers", "shim_startup_job.py")]
    return subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("world,backlog", [(16, None), (8, "1")])
def test_torchrun_ranks_all_reach_the_driver(world, backlog):
    """backlog=1 forces the overflow the old code died of: the executors' bounded handshake + reconnect must still get through."""
    r = _run(world, 29900 + world + (1 if backlog else 0), {"SPARKSHIM_LISTEN_BACKLOG": backlog} if backlog else None)
    assert r.returncode == 0 and "SHIM_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
