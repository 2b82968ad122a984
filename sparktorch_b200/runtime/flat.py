"""Flat parameter / gradient arenas.

Every trainable parameter of the user's module becomes a *view* into one
contiguous fp32 buffer, and every ``.grad`` a view into a second one of the same
layout.  This turns the reference's one-collective-per-parameter loop
(distributed.py:179-181) into a single collective over one buffer and lets the
optimizer run as one fused pass.  The gradient arena has ``TAIL`` extra fp32
slots at the end that ride in the same collective:

    tail[0] = this rank's loss for the step (summed → mean by the reduce)
    tail[1] = this rank's "I want to stop" vote
    tail[2..3] reserved
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn as nn

TAIL = 4
ALIGN = 64  # elements; every parameter starts 256 B aligned (vector/TMA friendly)


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class FlatArena:
    @staticmethod
    def measure(module: nn.Module) -> int:
        """Padded number of fp32 elements the arena of ``module`` needs."""
        return sum(_round_up(p.numel(), ALIGN) for p in module.parameters() if p.requires_grad)

    def __init__(self, module: nn.Module, device: torch.device, grad_buffer: torch.Tensor = None,
                 param_buffer: torch.Tensor = None):
        self.module = module
        self.params: List[nn.Parameter] = [p for p in module.parameters() if p.requires_grad]
        self.offsets: List[int] = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += _round_up(p.numel(), ALIGN)
        self.numel = off                      # padded parameter elements
        self.total = off + TAIL               # + piggy-back slots
        self.flat_p = param_buffer if param_buffer is not None else torch.zeros(self.numel, dtype=torch.float32, device=device)
        self.flat_g = grad_buffer if grad_buffer is not None else torch.zeros(self.total, dtype=torch.float32, device=device)
        assert self.flat_p.numel() >= self.numel and self.flat_g.numel() >= self.total
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                n = p.numel()
                self.flat_p[o:o + n].copy_(p.detach().reshape(-1).to(torch.float32))
                p.data = self.flat_p[o:o + n].view(p.shape)
                p.grad = self.flat_g[o:o + n].view(p.shape)

    @property
    def grads(self) -> torch.Tensor:
        return self.flat_g[: self.numel]

    @property
    def tail(self) -> torch.Tensor:
        return self.flat_g[self.numel: self.numel + TAIL]

    def zero_grad(self) -> None:
        self.flat_g.zero_()
        # autograd accumulates into existing .grad tensors in place, keeping the views
        for p, o in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * o:
                p.grad = self.flat_g[o:o + p.numel()].view(p.shape)

    def detach_state_dict(self) -> Dict[str, torch.Tensor]:
        """A CPU state_dict with private storage (safe to pickle)."""
        return {k: v.detach().to("cpu", copy=True) for k, v in self.module.state_dict().items()}
