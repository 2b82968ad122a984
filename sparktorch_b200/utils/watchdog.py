"""Stall detection for device-side training loops.

The fused engines hand the whole step loop to the GPU (CUDA-graph replays driven from C++); across GPUs the
steps synchronise through flag barriers in peer memory.  If one rank dies, or a kernel of one rank faults, its
peers would spin in that barrier forever and the host threads would sit in a stream synchronise that never
returns - the reference has the same failure mode in gloo's collectives and relies on Spark's task timeout.
Here a small host thread watches a progress counter (the iteration number the update kernel stores to the pinned
mailbox every step) and, if it has not advanced for ``timeout_s``, reports where every Python thread is and
terminates the process, so the launcher (Spark barrier stage / torchrun) fails the stage loudly instead of hanging.

``SPARKTORCH_STALL_TIMEOUT_S`` overrides the timeout (0 disables).
"""
from __future__ import annotations

import faulthandler
import os
import sys
import threading
import time
from typing import Callable, Optional


def default_timeout() -> float:
    try:
        return float(os.environ.get("SPARKTORCH_STALL_TIMEOUT_S", "300"))
    except ValueError:
        return 300.0


def _in_driver_process() -> bool:
    """Barrier task 0 of the single-node runtime runs inline in the user's driver / notebook process."""
    return os.environ.get("SPARKSHIM_EXECUTOR") != "1" and os.environ.get("SPARKTORCH_WORKER_PROCESS") != "1"


def _abort(message: str) -> None:      # pragma: no cover - kills the process
    sys.stderr.write(message + "\n")
    try:
        faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
    finally:
        sys.stderr.flush()
    if _in_driver_process():
        # never take the user's driver / notebook down: make the stuck CUDA call return with an error instead (the
        # training thread then raises, the stage fails and is retried / reported by the normal path)
        try:
            import ctypes

            ctypes.CDLL("libcudart.so").cudaDeviceReset()
        except Exception:
            pass
        return
    os._exit(86)


class ProgressWatchdog:
    """Calls ``on_stall(message)`` when ``get_progress()`` stops changing for ``timeout_s`` seconds."""

    def __init__(self, get_progress: Callable[[], int], timeout_s: Optional[float] = None, label: str = "",
                 on_stall: Callable[[str], None] = _abort, poll_s: float = 1.0, grace_first_s: float = 0.0):
        self.get_progress = get_progress
        self.timeout_s = default_timeout() if timeout_s is None else float(timeout_s)
        self.label = label
        self.on_stall = on_stall
        self.poll_s = min(poll_s, max(0.01, self.timeout_s / 4)) if self.timeout_s > 0 else poll_s
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self.fired = False
        self.grace_first_s = grace_first_s      # extra allowance until the FIRST progress (module load, first huge step)

    def _run(self) -> None:
        last, since = self.get_progress(), time.monotonic()
        moved = False
        while not self._stop.wait(self.poll_s):
            cur = self.get_progress()
            now = time.monotonic()
            if cur != last:
                last, since, moved = cur, now, True
            elif now - since >= self.timeout_s + (0.0 if moved else self.grace_first_s):
                self.fired = True
                self.on_stall(f"[sparktorch_b200] {self.label}: no progress for {now - since:.0f} s at iteration {cur} "
                              f"(a peer rank died or a kernel is stuck in a cross-GPU barrier); aborting this task")
                return

    def __enter__(self) -> "ProgressWatchdog":
        if self.timeout_s > 0:
            self._thread = threading.Thread(target=self._run, name="sparktorch-stall-watchdog", daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc) -> None:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=5)
