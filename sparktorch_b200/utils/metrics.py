"""Training metrics returned to the driver (the reference only prints; its loss history is lost,
SURVEY.md §5).  ``SparkTorchModel.getTrainingHistory()`` carries the raw dict; these helpers
render it as JSON lines or Prometheus text."""
from __future__ import annotations

import json
from typing import Any, Dict


def throughput(history: Dict[str, Any]) -> float:
    """samples/s over the training loop (whole job)."""
    secs = max(1e-9, float(history.get("loop_seconds") or 0.0))
    return float(history.get("samples_per_step", 0)) * float(history.get("iters_run", 0)) / secs


def to_jsonl(history: Dict[str, Any]) -> str:
    lines = []
    for i, loss in enumerate(history.get("train_loss", [])):
        lines.append(json.dumps({"iter": i, "loss": loss}))
    meta = {k: v for k, v in history.items() if k != "train_loss"}
    meta["samples_per_second"] = throughput(history)
    lines.append(json.dumps({"summary": meta}))
    return "\n".join(lines)


def to_prometheus(history: Dict[str, Any], job: str = "sparktorch") -> str:
    g = []
    losses = history.get("train_loss") or [float("nan")]
    g.append(f'sparktorch_last_loss{{job="{job}"}} {losses[-1]}')
    g.append(f'sparktorch_iterations_total{{job="{job}"}} {history.get("iters_run", 0)}')
    g.append(f'sparktorch_samples_per_second{{job="{job}"}} {throughput(history)}')
    g.append(f'sparktorch_world_size{{job="{job}"}} {history.get("world_size", 1)}')
    g.append(f'sparktorch_early_stopped{{job="{job}"}} {0 if history.get("stopped_at") is None else 1}')
    return "\n".join(g) + "\n"
