"""Mid-training checkpoint / resume (the reference has none: SURVEY.md §5 – only the
end-of-fit ``modStr``).  Every rank writes ``<dir>/rank<r>.pt`` atomically:
fp32 master weights, its optimizer-state shard, step/iteration counters, the early-stopping
record and the loss history so far.  ``resume=True`` continues from the newest complete set."""
from __future__ import annotations

import os
import tempfile
from typing import Any, Dict, Optional

import torch

FORMAT = 1


def path_for(directory: str, rank: int) -> str:
    return os.path.join(directory, f"rank{rank}.pt")


def save(directory: str, rank: int, payload: Dict[str, Any]) -> str:
    os.makedirs(directory, exist_ok=True)
    payload = dict(payload, format=FORMAT)
    fd, tmp = tempfile.mkstemp(dir=directory, prefix=f".rank{rank}.", suffix=".tmp")
    os.close(fd)
    torch.save(payload, tmp)
    final = path_for(directory, rank)
    os.replace(tmp, final)          # atomic: a crash never leaves a torn checkpoint
    return final


def load(directory: Optional[str], rank: int) -> Optional[Dict[str, Any]]:
    if not directory:
        return None
    p = path_for(directory, rank)
    if not os.path.exists(p):
        return None
    payload = torch.load(p, map_location="cpu", weights_only=False)
    if payload.get("format") != FORMAT:
        return None
    return payload
