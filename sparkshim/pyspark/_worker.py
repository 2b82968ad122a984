"""Executor process: connect to the driver, run tasks until told to exit."""
from __future__ import annotations

import sys
import traceback
from multiprocessing.connection import Client

import cloudpickle


def serve(host: str, port: int, slot: int) -> None:
    import os

    os.environ["SPARKSHIM_EXECUTOR"] = "1"      # a dedicated worker process (stall watchdogs may terminate it, never the driver)
    from pyspark._cluster import _AUTH, bind_gpu
    from pyspark.taskcontext import BarrierTaskContext, TaskContext, _set_context

    conn = Client((host, port), authkey=_AUTH)
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
