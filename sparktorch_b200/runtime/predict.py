"""Batched inference for ``SparkTorchModel.transform``.

The reference calls the network once per ROW from a Python UDF (batch size 1,
``model.eval()`` per row, CPU only; torch_distributed.py:106-130, quirk Q15).
Here a partition is scored in large batches on the executor's device; output
conversion rules are the reference's: ``useVectorOut`` → the flattened output
vector, else ``float(argmax)`` for multi-output nets and the raw scalar for
single-output nets.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.nn as nn


def pick_device(device: Optional[str], index: int = 0) -> torch.device:
    if device in (None, "", "auto"):
        device = "cuda" if torch.cuda.is_available() else "cpu"
    if device == "cuda":
        return torch.device("cuda", index % max(1, torch.cuda.device_count()))
    return torch.device(device)


class BatchPredictor:
    def __init__(self, model: nn.Module, device: torch.device, batch_size: int = 65536, precision: str = "auto"):
        self.device = device
        self.batch_size = int(batch_size)
        self.model = model.to(device).eval()
        self.fused = None
        if device.type == "cuda":
            try:  # recognised MLPs run on our sm_100a kernels (no autograd, bf16 tensor cores)
                from sparktorch_b200.runtime.fused_mlp import try_fused_predictor

                self.fused = try_fused_predictor(self.model, device, precision)
            except ImportError:
                self.fused = None

    @torch.no_grad()
    def predict_block(self, X: np.ndarray) -> np.ndarray:
        """float32 [n, d] → float32 [n, k] (model output flattened per row)."""
        n = X.shape[0]
        outs = []
        pin = self.device.type == "cuda"
        # bound one host->device transfer to ~256 MiB: image-sized rows (ResNet: 602 KB each) must not be staged
        # 65536 at a time
        row_bytes = max(1, int(np.prod(X.shape[1:])) * 4)
        step = max(1, min(self.batch_size, (256 << 20) // row_bytes))
        for s in range(0, n, step):
            xb = torch.from_numpy(np.ascontiguousarray(X[s:s + step], dtype=np.float32))
            if pin:
                xb = xb.pin_memory().to(self.device, non_blocking=True)
            out = self.fused(xb) if self.fused is not None else self.model(xb)
            outs.append(out.reshape(out.shape[0], -1).float().cpu().numpy())
        return np.concatenate(outs, axis=0) if len(outs) > 1 else outs[0]


def to_prediction_column(raw: np.ndarray, vector_out: bool):
    """[n, k] → list of DenseVector | 1-D float64 array."""
    if vector_out:
        return raw.astype(np.float64)
    if raw.shape[1] > 1:
        return raw.argmax(axis=1).astype(np.float64)
    return raw[:, 0].astype(np.float64)
