"""``from graphattention.SPGA import SPGACF`` keeps working (run_Gowalla.py:20 of the reference)."""
from ngacf_b200.spgat import SPGACF, HomoGraph, SpGAT, SpGraphAttentionLayer  # noqa: F401
