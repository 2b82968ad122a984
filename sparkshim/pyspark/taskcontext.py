"""TaskContext / BarrierTaskContext with the surface the trainer uses
(SURVEY.md §2.9): ``get()``, ``partitionId()``, ``getTaskInfos()[i].address``,
``allGather(str)``, ``barrier()``."""
from __future__ import annotations

import threading
from typing import Callable, List, Optional

_local = threading.local()


def _set_context(ctx) -> None:
    _local.ctx = ctx


class BarrierTaskInfo:
    def __init__(self, address: str):
        self.address = address

    def __repr__(self):
        return f"BarrierTaskInfo({self.address!r})"


class TaskContext:
    def __init__(self, stage_id: int, partition_id: int, num_partitions: int):
        self._stage_id = stage_id
        self._partition_id = partition_id
        self._num_partitions = num_partitions

    @classmethod
    def get(cls) -> Optional["TaskContext"]:
        return getattr(_local, "ctx", None)

    def stageId(self) -> int:
        return self._stage_id

    def partitionId(self) -> int:
        return self._partition_id

    def attemptNumber(self) -> int:
        return 0

    def taskAttemptId(self) -> int:
        return self._stage_id * 100000 + self._partition_id


class BarrierTaskContext(TaskContext):
    def __init__(self, stage_id, partition_id, num_partitions, addresses: List[str], gather: Callable[[str], List[str]]):
        super().__init__(stage_id, partition_id, num_partitions)
        self._addresses = addresses
        self._gather = gather

    @classmethod
    def get(cls) -> "BarrierTaskContext":
        ctx = getattr(_local, "ctx", None)
        if not isinstance(ctx, BarrierTaskContext):
            raise RuntimeError("It is not in a barrier stage")
        return ctx

    def getTaskInfos(self) -> List[BarrierTaskInfo]:
        return [BarrierTaskInfo(a) for a in self._addresses]

    def allGather(self, message: str = "") -> List[str]:
        if not isinstance(message, str):
            raise TypeError("Argument `message` must be of type `str`")
        return self._gather(message)

    def barrier(self) -> None:
        self._gather("")
