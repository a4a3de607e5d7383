"""``from graphattention.SPUIGACF import SPUIGACF, SPUIMultiGACF`` keeps working (run_Gowalla.py:23)."""
from ngacf_b200.model import SPUIGACF, SPUIMultiGACF, SPUIMultiGAT, SpUIGAT, SpUIGraphAttentionLayer  # noqa: F401
