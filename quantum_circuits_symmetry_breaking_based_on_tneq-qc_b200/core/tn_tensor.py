"""TNTensor: a tensor with an out-of-band scale so fp32 products of many tiny
measurement matrices neither underflow nor lose the loss contribution of the
scale.  Mirrors the reference class one to one
(reference: tneq_qc/core/tn_tensor.py:4-125): value = tensor * scale,
log_scale = ln|scale| carried as a host float.

Two methods the reference forgot (SURVEY defect D2: the gradient path calls
them, greedy_strategy.py:677-681) are provided: is_complex() and conj().
"""
from __future__ import annotations

import math
from typing import Any


class TNTensor:
    def __init__(self, tensor: Any, scale: Any = 1.0, log_scale: float = None):
        self._tensor = tensor
        self.scale = float(scale)
        if log_scale is None:
            log_scale = math.log(abs(self.scale)) if self.scale != 0 else float("-inf")
        self.log_scale = log_scale

    # -- views on the wrapped tensor ---------------------------------------
    @property
    def tensor(self):
        return self._tensor

    @property
    def ndim(self) -> int:
        return self._tensor.ndim

    @property
    def shape(self) -> tuple:
        return self._tensor.shape

    @property
    def dtype(self):
        return self._tensor.dtype

    def is_complex(self) -> bool:
        return self._tensor.is_complex()

    def conj(self) -> "TNTensor":
        return TNTensor(self._tensor.conj(), self.scale, self.log_scale)

    # -- rescaling (represented value never changes) -----------------------
    def auto_scale(self):
        """Normalise so that max|tensor| == 1 (one host sync, as in the reference)."""
        peak = self._tensor.abs().max()
        peak = peak.item() if hasattr(peak, "item") else float(peak)
        if peak == 0:
            return
        self._tensor /= peak
        self.scale *= peak
        self.log_scale += math.log(abs(peak))

    def scale_to(self, new_scale: float):
        new_scale = float(new_scale)
        if new_scale == 0:
            raise ValueError("Cannot scale to 0.")
        self._tensor = self._tensor * (self.scale / new_scale)
        self.scale = new_scale
        self.log_scale = math.log(abs(new_scale))

    def scale_with(self, factor: float):
        factor = float(factor)
        if factor == 0:
            raise ValueError("Cannot scale with factor 0.")
        self._tensor = self._tensor / factor
        self.scale *= factor
        self.log_scale += math.log(abs(factor))

    def __repr__(self):
        return f"TNTensor(shape={getattr(self._tensor, 'shape', 'unknown')}, scale={self.scale})"
