"""Mid-training checkpoint / resume (the reference has none: SURVEY.md §5 – only the
end-of-fit ``modStr``).  Every rank writes ``<dir>/rank<r>.pt`` atomically:
fp32 master weights, its optimizer-state shard, step/iteration counters, the early-stopping
record and the loss history so far, keeping the previous generation (``rank<r>.prev.pt``).  On resume the
ranks agree (one MIN/MAX exchange, :func:`agree_on_iteration`) on the newest iteration that EVERY rank holds;
if there is none, all of them restart from iteration 0."""
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
    if os.path.exists(final):       # keep one older generation: a crash between two ranks' saves leaves a common one
        os.replace(final, prev_path_for(directory, rank))
    os.replace(tmp, final)          # atomic: a crash never leaves a torn checkpoint
    return final


def prev_path_for(directory: str, rank: int) -> str:
    return os.path.join(directory, f"rank{rank}.prev.pt")


def load_generations(directory: Optional[str], rank: int) -> Dict[int, Dict[str, Any]]:
    """{iters_done: payload} of the (up to two) generations this rank holds."""
    out: Dict[int, Dict[str, Any]] = {}
    if not directory:
        return out
    for p in (prev_path_for(directory, rank), path_for(directory, rank)):
        if os.path.exists(p):
            try:
                payload = torch.load(p, map_location="cpu", weights_only=False)
            except Exception:
                continue
            if payload.get("format") == FORMAT:
                out[int(payload["iters_done"])] = payload
    return out


def agree_on_iteration(mine, world: int, device) -> int:
    """Newest iteration present on every rank (0: none).  ``mine``: sorted iterations this rank holds (at most two)."""
    mine = [int(x) for x in mine][-2:]
    if world <= 1:
        return mine[-1] if mine else 0
    import torch.distributed as dist

    dev = device if (device is not None and getattr(device, "type", "cpu") == "cuda" and dist.get_backend() == "nccl") else torch.device("cpu")
    slots = torch.zeros(world, 2, dtype=torch.int64, device=dev)
    for j, it in enumerate(mine):
        slots[dist.get_rank(), j] = it
    dist.all_reduce(slots, op=dist.ReduceOp.SUM)
    held = [set(int(v) for v in row.tolist() if v > 0) for row in slots.cpu()]
    common = set.intersection(*held) if held else set()
    return max(common) if common else 0


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
