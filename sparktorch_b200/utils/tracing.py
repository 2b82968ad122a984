"""Tracing / profiling hooks (absent in the reference).

* :func:`nvtx_range` – NVTX ranges around engine phases (visible in Nsight tools).
* :class:`PhaseTimer` – device-side timing of named phases with CUDA events (falls back to
  wall clock on CPU); ``SPARKTORCH_TRACE=/path/trace.json`` makes the engines dump a Chrome
  ``chrome://tracing`` file per rank.
"""
from __future__ import annotations

import contextlib
import json
import os
import time
from typing import Dict, List

import torch


@contextlib.contextmanager
def nvtx_range(name: str):
    pushed = False
    if torch.cuda.is_available():
        try:
            torch.cuda.nvtx.range_push(name)
            pushed = True
        except Exception:
            pushed = False
    try:
        yield
    finally:
        if pushed:
            torch.cuda.nvtx.range_pop()


class PhaseTimer:
    def __init__(self, device: torch.device, enabled: bool = True):
        self.device, self.enabled = device, enabled
        self.cuda = device.type == "cuda"
        self.events: List = []
        self.t0 = time.perf_counter()

    @contextlib.contextmanager
    def phase(self, name: str):
        if not self.enabled:
            yield
            return
        with nvtx_range(name):
            if self.cuda:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                w0 = time.perf_counter()
                e0.record()
                yield
                e1.record()
                self.events.append((name, w0 - self.t0, e0, e1))
            else:
                w0 = time.perf_counter()
                yield
                self.events.append((name, w0 - self.t0, time.perf_counter() - w0, None))

    def summary(self) -> Dict[str, Dict[str, float]]:
        if self.cuda:
            torch.cuda.synchronize(self.device)
        out: Dict[str, Dict[str, float]] = {}
        for name, _start, a, b in self.events:
            ms = a.elapsed_time(b) if b is not None else a * 1e3
            rec = out.setdefault(name, {"count": 0, "total_ms": 0.0})
            rec["count"] += 1
            rec["total_ms"] += ms
        return out

    def dump_chrome_trace(self, path: str, rank: int = 0) -> None:
        if self.cuda:
            torch.cuda.synchronize(self.device)
        ev = []
        for name, start, a, b in self.events:
            dur = (a.elapsed_time(b) if b is not None else a * 1e3) * 1e3
            ev.append({"name": name, "ph": "X", "ts": start * 1e6, "dur": dur, "pid": rank, "tid": 0})
        with open(path, "w") as fh:
            json.dump({"traceEvents": ev}, fh)


def trace_path(rank: int):
    base = os.environ.get("SPARKTORCH_TRACE")
    if not base:
        return None
    root, ext = os.path.splitext(base)
    return f"{root}.rank{rank}{ext or '.json'}"
