"""``from graphattention.SPUIGACF import SPUIGACF`` keeps working (run_Gowalla.py:23)."""
from ngacf_b200.model import SPUIGACF, SpUIGAT, SpUIGraphAttentionLayer  # noqa: F401
