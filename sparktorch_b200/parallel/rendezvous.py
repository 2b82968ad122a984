"""Barrier-stage → torch.distributed bootstrap.

Contract taken from the launcher (real Spark or sparkshim):
``BarrierTaskContext.get()``, ``getTaskInfos()[i].address``, ``partitionId()``,
``allGather(str)`` (reference: distributed.py:35-50,95-110).  Rank 0 probes a
free TCP port and shares it through ``allGather``; every task then joins an
``env://`` process group – ``nccl`` when training on GPUs (the reference
hard-codes gloo even for CUDA tensors), ``gloo`` on CPU.
"""
from __future__ import annotations

import datetime
import socket
from typing import Tuple

import torch
import torch.distributed as dist


def find_free_port(address: str) -> int:
    s = socket.socket()
    try:
        s.bind((address, 0))
        return s.getsockname()[1]
    finally:
        s.close()


def share_port(address: str, context) -> int:
    port = ""
    if context.partitionId() == 0:
        try:
            port = str(find_free_port(address))
        except OSError:
            port = ""
    got = context.allGather(port)[0]
    if not got:
        raise RuntimeError("Failed to find free port for distributed training.")
    return int(got)


def init_from_barrier(context, device: torch.device) -> Tuple[int, int]:
    """Returns (rank, world_size); creates the default process group when
    world_size > 1 (a single task needs no group at all)."""
    addrs = [info.address.split(":")[0] for info in context.getTaskInfos()]
    world = len(addrs)
    rank = context.partitionId()
    if dist.is_initialized():  # python worker reuse (reference distributed.py:95-96)
        dist.destroy_process_group()
    if world == 1:
        return 0, 1
    port = share_port(addrs[0], context)
    kwargs = {}
    if device.type == "cuda":
        backend = "nccl"
        kwargs["device_id"] = device
    else:
        backend = "gloo"
    # tcp:// rendezvous: nothing is written to os.environ (the driver process
    # hosts task 0 inline and must not inherit a stale WORLD_SIZE/RANK)
    dist.init_process_group(backend, init_method=f"tcp://{addrs[0]}:{port}", rank=rank, world_size=world,
                            timeout=datetime.timedelta(seconds=600), **kwargs)
    return rank, world


def shutdown() -> None:
    if dist.is_initialized():
        try:
            dist.destroy_process_group()
        except Exception:
            pass
