"""BPR loss of the PairSampling path, backed by ``ngacf_bpr_loss``.

The reference's module of the same name (graphattention/BPRLoss.py:4-9) evaluates the naive ``-log(sigmoid(pos - neg)).mean()`` with
three elementwise torch kernels and a reduction; here the loss and both score gradients come from one single-block kernel in the
numerically stable form ``softplus(neg - pos).mean()`` (identical wherever the naive form is finite, finite where it overflows),
summed with a fixed reduction tree so the value is run-to-run deterministic.  The autograd glue is ``propagation.BPRLossFn``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ._lib import NgacfError
from .propagation import BPRLossFn


class BPRLoss(nn.Module):
    """``loss = BPRLoss()(pos_scores, neg_scores)`` -- 0-dim float32, differentiable w.r.t. both score vectors."""

    def forward(self, pos_scores: torch.Tensor, neg_scores: torch.Tensor) -> torch.Tensor:
        if pos_scores.shape != neg_scores.shape or pos_scores.dim() != 1:
            raise ValueError("BPRLoss expects two score vectors of the same length, got %s and %s" % (tuple(pos_scores.shape), tuple(neg_scores.shape)))
        if pos_scores.numel() == 0:
            raise ValueError("BPRLoss of an empty batch is undefined (the reference's mean() returns NaN)")
        if not (pos_scores.is_cuda and neg_scores.is_cuda):
            raise NgacfError("BPRLoss runs on CUDA tensors only; there is no CPU path")
        return BPRLossFn.apply(pos_scores.float(), neg_scores.float())

    def extra_repr(self) -> str:
        return "stable softplus form, single-kernel forward+backward"
