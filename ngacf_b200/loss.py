"""Drop-in ``BPRLoss`` (graphattention/BPRLoss.py:4-9): ``-log(sigmoid(pos - neg)).mean()``."""
from __future__ import annotations

import torch.nn as nn

from .propagation import BPRLossFn


class BPRLoss(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, pos_scores, neg_scores):
        return BPRLossFn.apply(pos_scores, neg_scores)
