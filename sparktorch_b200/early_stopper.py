"""Patience-based early stopping.

Behavioural contract (reference: sparktorch/early_stopper.py:8-56):
``step(metric) -> bool``; the first call only seeds ``best``; NaN stops
immediately; ``patience == 0`` never stops; ``mode`` is ``min``/``max``;
``min_delta`` is absolute or, with ``percentage=True``, a percentage of
``best``.

The state is a plain 4-float record so the very same rule can run on the
device: :meth:`state_vector` / :meth:`load_state_vector` mirror the struct the
fused kernels update in-band (``StepState`` in csrc/kernels.h: K-AR in
csrc/comm_ar.cu, the DECIDE task of K-STEP in csrc/mlp_step.cu)."""
from __future__ import annotations

import math


class EarlyStopping:
    def __init__(self, mode: str = "min", min_delta: float = 0, patience: int = 10, percentage: bool = False):
        if mode not in ("min", "max"):
            raise ValueError("mode " + str(mode) + " is unknown!")
        self.mode = mode
        self.min_delta = min_delta
        self.patience = patience
        self.percentage = percentage
        self.best = None
        self.num_bad_epochs = 0

    # -- comparison ---------------------------------------------------------
    def is_better(self, value: float, best: float) -> bool:
        if self.patience == 0:
            return True
        margin = best * self.min_delta / 100 if self.percentage else self.min_delta
        return value < best - margin if self.mode == "min" else value > best + margin

    # -- update -------------------------------------------------------------
    def step(self, metrics: float) -> bool:
        if self.patience == 0:
            return False
        if self.best is None:
            self.best = metrics
            return False
        if math.isnan(metrics):
            return True
        if self.is_better(metrics, self.best):
            self.num_bad_epochs = 0
            self.best = metrics
        else:
            self.num_bad_epochs += 1
        return self.num_bad_epochs >= self.patience

    # -- device mirror ------------------------------------------------------
    def state_vector(self):
        """[seeded, best, num_bad, patience] as floats."""
        return [
            0.0 if self.best is None else 1.0,
            0.0 if self.best is None else float(self.best),
            float(self.num_bad_epochs),
            float(self.patience),
        ]

    def load_state_vector(self, vec) -> None:
        seeded, best, bad, _ = [float(v) for v in vec[:4]]
        self.best = best if seeded > 0 else None
        self.num_bad_epochs = int(bad)
