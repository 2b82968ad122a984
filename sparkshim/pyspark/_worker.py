"""Executor process: connect to the driver, run tasks until told to exit."""
from __future__ import annotations

import sys
import traceback

import cloudpickle


def _connect(host: str, port: int, authkey: bytes, slot: int, attempts: int = 6, handshake_s: float = 20.0):
    """``multiprocessing.connection.Client`` with a bounded handshake: if the driver's challenge does not arrive in time (the
    connection never reached its accept queue) the socket is dropped and the connect is retried, slightly staggered per slot."""
    import socket
    import time
    from multiprocessing.connection import Connection, answer_challenge, deliver_challenge

    last = None
    for a in range(attempts):
        time.sleep(0.01 * (slot % 16) + 0.25 * a)
        try:
            sock = socket.create_connection((host, port), timeout=handshake_s)
            sock.setblocking(True)
            c = Connection(sock.detach())
        except OSError as e:
            last = e
            continue
        try:
            if not c.poll(handshake_s):
                raise TimeoutError("no authentication challenge from the driver")
            answer_challenge(c, authkey)
            deliver_challenge(c, authkey)
            return c
        except Exception as e:      # noqa: BLE001 - any failed handshake is retried on a fresh socket
            last = e
            try:
                c.close()
            except OSError:
                pass
    raise ConnectionError(f"executor {slot}: could not reach the driver at {host}:{port}: {last!r}")


def serve(host: str, port: int, slot: int) -> None:
    import os

    os.environ["SPARKSHIM_EXECUTOR"] = "1"      # a dedicated worker process (stall watchdogs may terminate it, never the driver)
    from pyspark._cluster import _AUTH, bind_gpu
    from pyspark.taskcontext import BarrierTaskContext, TaskContext, _set_context

    conn = _connect(host, port, _AUTH, slot)
    conn.send(slot)

    def all_gather(msg: str):
        conn.send(("allgather", msg))
        kind, payload = conn.recv()
        if kind != "gathered":
            raise RuntimeError(f"barrier allGather failed: {payload}")
        return payload

    while True:
        try:
            kind, payload = conn.recv()
        except (EOFError, OSError):
            return
        if kind == "exit":
            return
        stage_id, idx, n_tasks, barrier, addresses, blob = payload
        bind_gpu(slot)
        ctx = (
            BarrierTaskContext(stage_id, idx, n_tasks, addresses, all_gather)
            if barrier
            else TaskContext(stage_id, idx, n_tasks)
        )
        _set_context(ctx)
        try:
            out = cloudpickle.loads(blob)(ctx)
            conn.send(("result", cloudpickle.dumps(out)))
        except BaseException:
            conn.send(("error", traceback.format_exc()))
        finally:
            _set_context(None)


if __name__ == "__main__":
    serve(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))
