"""``from graphattention.SPUIGACF import SPUIGACF, SPUIMultiGACF, SPUIGAGPCF`` keeps working (run_Gowalla.py:23)."""
from ngacf_b200.gp import GPLayer, SPUIGAGPCF  # noqa: F401
from ngacf_b200.model import SPUIGACF, SPUIMultiGACF, SPUIMultiGAT, SpUIGAT, SpUIGraphAttentionLayer  # noqa: F401
