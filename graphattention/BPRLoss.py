"""``from graphattention.BPRLoss import BPRLoss`` keeps working (run_Gowalla.py:33)."""
from ngacf_b200.loss import BPRLoss  # noqa: F401
