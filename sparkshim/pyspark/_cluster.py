"""Single-node executor pool with Spark's barrier-stage contract.

This is the piece of Spark the trainer actually depends on:
N concurrently running Python tasks, one per partition, that can discover each
other (``getTaskInfos``), ``allGather`` short strings and ``barrier()``.

Topology
--------
slot 0 is the *driver process itself* (tasks for slot 0 run inline on the thread
that called the action, so with one partition nothing is ever spawned).
slots 1..N-1 are executor processes:

* ``spawn`` mode  – started on demand with ``python -m pyspark._worker`` and
  kept alive for reuse (like Spark's python worker reuse);
* ``torchrun`` mode – when the program was launched by ``torch.distributed.run``
  the non-zero ranks *become* the executors: ``SparkSession.getOrCreate()`` on
  those ranks never returns, it serves tasks for rank 0 and exits with it.

Executor slot ``s`` binds CUDA device ``s % device_count`` before a task runs
(the GPU-resource assignment a Spark cluster manager would do).

Tasks and results travel as cloudpickle blobs over ``multiprocessing.connection``
sockets on 127.0.0.1.
"""
from __future__ import annotations

import atexit
import multiprocessing
import os
import queue
import socket
import subprocess
import sys
import threading
import time
import traceback
from multiprocessing.connection import Listener
from typing import Any, Callable, List, Optional

import cloudpickle

_AUTH = b"sparkshim-local"


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def under_torchrun() -> bool:
    """True when this process is one rank of a ``torch.distributed.run`` job."""
    return (int(os.environ.get("WORLD_SIZE", "1")) > 1 and "RANK" in os.environ
            and ("TORCHELASTIC_RUN_ID" in os.environ or "LOCAL_RANK" in os.environ))


def bind_gpu(slot: int) -> None:
    """Executor → GPU assignment (only if torch+CUDA are present)."""
    try:
        import torch

        if torch.cuda.is_available():
            torch.cuda.set_device(slot % torch.cuda.device_count())
    except Exception:  # pragma: no cover - torch missing / no driver
        pass


class StageSync:
    """Driver-side coordinator for allGather/barrier of one barrier stage."""

    def __init__(self, n: int):
        self.n = n
        self.cv = threading.Condition()
        self.round = 0
        self.msgs: List[Optional[str]] = [None] * n
        self.arrived = 0
        self.result: List[str] = []
        self.failed: Optional[str] = None

    def all_gather(self, idx: int, msg: str, timeout: float = 600.0) -> List[str]:
        with self.cv:
            my_round = self.round
            self.msgs[idx] = msg
            self.arrived += 1
            if self.arrived == self.n:
                self.result = list(self.msgs)  # type: ignore[arg-type]
                self.msgs = [None] * self.n
                self.arrived = 0
                self.round += 1
                self.cv.notify_all()
                return self.result
            deadline = time.time() + timeout
            while self.round == my_round and self.failed is None:
                left = deadline - time.time()
                if left <= 0:
                    raise RuntimeError("barrier allGather timed out")
                self.cv.wait(left)
            if self.failed is not None and self.round == my_round:
                raise RuntimeError("barrier stage failed: " + self.failed)
            return self.result

    def fail(self, why: str) -> None:
        with self.cv:
            self.failed = why
            self.cv.notify_all()


class _Slot:
    """Driver-side handle of one remote executor."""

    def __init__(self, slot: int, conn, proc=None):
        self.slot = slot
        self.conn = conn
        self.proc = proc
        self.lock = threading.Lock()


class LocalCluster:
    def __init__(self, n_slots: int):
        self.n_slots = max(1, int(n_slots))
        self.mode = "torchrun" if under_torchrun() else "spawn"
        self._slots: dict[int, _Slot] = {}
        self._listener: Optional[Listener] = None
        self._port: Optional[int] = None
        self._stage_id = 0
        self._lock = threading.Lock()
        self._closed = False
        atexit.register(self.shutdown)
        if self.mode == "torchrun":
            # the other ranks are already running and waiting for our port
            self.n_slots = int(os.environ["WORLD_SIZE"])
            self._ensure_slots(self.n_slots)

    # ------------------------------------------------------------------ pool
    def _ensure_listener(self) -> None:
        if self._listener is None:
            self._port = _free_port()
            # backlog: multiprocessing's default is 1.  Under torchrun the executors all read the port from the store and connect in
            # the same instant; with a one-entry accept queue the kernel drops the overflowing handshakes, and a client whose final
            # ACK was dropped believes it is connected and then waits for an authentication challenge that never comes, while
            # the driver waits in accept() - the intermittent 8-rank start-up hang (stack dumps: profiles/hang_8gpu_startup.md)
            backlog = int(os.environ.get("SPARKSHIM_LISTEN_BACKLOG", "0")) or max(128, 4 * self.n_slots)
            self._listener = Listener(("127.0.0.1", self._port), backlog=backlog, authkey=_AUTH)
            if self.mode == "torchrun":
                _publish_port(self._port)

    def _ensure_slots(self, upto: int) -> None:
        """Make sure executors 1..upto-1 are connected."""
        need = [s for s in range(1, upto) if s not in self._slots]
        if not need:
            return
        self._ensure_listener()
        procs = {}
        if self.mode == "spawn":
            env = dict(os.environ)
            env["PYTHONPATH"] = os.pathsep.join([p for p in sys.path if p])
            for s in need:
                env_s = dict(env, SPARKSHIM_SLOT=str(s))
                procs[s] = subprocess.Popen(
                    [sys.executable, "-m", "pyspark._worker", "127.0.0.1", str(self._port), str(s)],
                    env=env_s,
                )
        pending = set(need)
        self._listener._listener._socket.settimeout(300.0)  # type: ignore[attr-defined]
        while pending:
            try:
                conn = self._listener.accept()          # includes the authentication handshake with that client
                slot = conn.recv()
            except TimeoutError:
                raise RuntimeError(f"executors {sorted(pending)} did not connect to the driver within 300 s") from None
            except (EOFError, OSError, multiprocessing.AuthenticationError):
                continue                                # a client that gave up on a stale connection and is reconnecting
            self._slots[slot] = _Slot(slot, conn, procs.get(slot))
            pending.discard(slot)

    def shutdown(self) -> None:
        if self._closed:
            return
        self._closed = True
        for s in self._slots.values():
            try:
                s.conn.send(("exit", None))
                s.conn.close()
            except Exception:
                pass
        for s in self._slots.values():
            if s.proc is not None:
                try:
                    s.proc.wait(timeout=10)
                except Exception:
                    s.proc.kill()
        self._slots.clear()
        if self._listener is not None:
            try:
                self._listener.close()
            except Exception:
                pass
            self._listener = None

    # ------------------------------------------------------------------ jobs
    def run_stage(self, make_task: Callable[[int], Any], n_tasks: int, barrier: bool) -> List[Any]:
        """Run ``n_tasks`` tasks concurrently; ``make_task(i)`` returns a
        picklable zero-arg... (callable taking the TaskContext)``.
        Barrier stages pin task i to slot i and need n_tasks <= slots."""
        from pyspark.taskcontext import BarrierTaskContext, TaskContext, _set_context

        if barrier and n_tasks > self.n_slots:
            raise RuntimeError(
                f"[SPARK-24819] barrier stage needs {n_tasks} concurrent tasks but only "
                f"{self.n_slots} slots are available (master local[{self.n_slots}])"
            )
        with self._lock:
            self._stage_id += 1
            stage_id = self._stage_id
        width = min(n_tasks, self.n_slots)
        self._ensure_slots(width)
        sync = StageSync(n_tasks) if barrier else None
        addresses = [f"127.0.0.1:{40000 + s}" for s in range(n_tasks)]
        results: List[Any] = [None] * n_tasks
        errors: List[Optional[str]] = [None] * n_tasks
        todo: "queue.Queue[int]" = queue.Queue()
        for i in range(n_tasks):
            todo.put(i)

        def remote_loop(slot: _Slot):
            while True:
                if barrier:
                    i = slot.slot
                    if i >= n_tasks:
                        return
                else:
                    try:
                        i = todo.get_nowait()
                    except queue.Empty:
                        return
                try:
                    blob = cloudpickle.dumps(make_task(i))
                    slot.conn.send(("task", (stage_id, i, n_tasks, barrier, addresses, blob)))
                    while True:
                        kind, payload = slot.conn.recv()
                        if kind == "allgather":
                            try:
                                slot.conn.send(("gathered", sync.all_gather(i, payload)))  # type: ignore[union-attr]
                            except Exception as e:  # propagate to the task
                                slot.conn.send(("gather_failed", str(e)))
                        elif kind == "result":
                            results[i] = cloudpickle.loads(payload)
                            break
                        elif kind == "error":
                            errors[i] = payload
                            if sync is not None:
                                sync.fail(payload)
                            break
                except (EOFError, OSError) as e:
                    errors[i] = f"executor {slot.slot} died: {e!r}"
                    if sync is not None:
                        sync.fail(errors[i])
                    return
                if barrier:
                    return

        threads = [threading.Thread(target=remote_loop, args=(self._slots[s],), daemon=True) for s in range(1, width)]
        for t in threads:
            t.start()

        # slot 0 = this thread
        bind_gpu(0)
        while True:
            if barrier:
                i = 0
            else:
                try:
                    i = todo.get_nowait()
                except queue.Empty:
                    break
            ctx = (
                BarrierTaskContext(stage_id, i, n_tasks, addresses, lambda m, i=i: sync.all_gather(i, m))  # type: ignore[union-attr]
                if barrier
                else TaskContext(stage_id, i, n_tasks)
            )
            _set_context(ctx)
            try:
                results[i] = make_task(i)(ctx)
            except BaseException:
                errors[i] = traceback.format_exc()
                if sync is not None:
                    sync.fail(errors[i])
            finally:
                _set_context(None)
            if barrier:
                break
        for t in threads:
            t.join()
        bad = [e for e in errors if e]
        if bad:
            raise RuntimeError("task failed:\n" + bad[0])
        return results


# ---------------------------------------------------------------- torchrun glue
def _store():
    from torch.distributed import TCPStore
    from datetime import timedelta

    return TCPStore(
        os.environ.get("MASTER_ADDR", "127.0.0.1"),
        int(os.environ["MASTER_PORT"]),
        is_master=False,
        timeout=timedelta(seconds=600),
        wait_for_workers=False,
    )


def _detach_from_agent_store() -> None:
    """Tasks create their own process groups on their own ports; they must not
    be redirected to torchrun's agent store (which only exists on MASTER_PORT)."""
    os.environ.pop("TORCHELASTIC_USE_AGENT_STORE", None)


def _publish_port(port: int) -> None:
    _store().set("sparkshim/driver_port", str(port))
    _detach_from_agent_store()


def serve_as_executor_rank() -> None:
    """Non-zero torchrun ranks: become an executor for rank 0, then exit."""
    from pyspark import _worker

    port = int(_store().get("sparkshim/driver_port").decode())
    _detach_from_agent_store()
    slot = int(os.environ["RANK"])
    _worker.serve("127.0.0.1", port, slot)
    sys.exit(0)
