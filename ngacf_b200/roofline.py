"""Algorithmic-byte models of the kernels (the `roofline.achieved` numerators of bench.py) and of the
whole training step (SURVEY.md 8d "compulsory model": every array crosses HBM once per kernel that needs
it; gathered tables count once because they stay L2-resident at the BASELINE shapes)."""
from __future__ import annotations

R = 256  # bytes of one 64-float row


def kernel_bytes(name: str, U: int, I: int, E: int, H: int, dropout: bool, B: int = 2048, n_params: int = 0) -> int:
    N = U + I
    sc = N * H * 4          # one per-head scalar array
    if name == "ngacf_transform_fwd":
        return 2 * N * R + sc + (N * 8 if dropout else 0)                       # read X, write h + s (+ mask words)
    if name == "ngacf_aggregate_fwd":
        return 3 * N * R + 2 * sc + 2 * E * 4 + ((2 * E * 4 + E) if dropout else 0)   # h own + h gathered(once) + Z; s, norm; adj (+eid, mask)
    if name == "ngacf_stage_bwd_prep":
        return 4 * N * R + 2 * sc                                               # G, Z, h -> Ghat; norm -> dN
    if name == "ngacf_stage_bwd_edges_users":
        return (N + I) * R + 2 * U * R + 3 * sc + 2 * E * 4 + E * H * 4 + (E if dropout else 0)
    if name == "ngacf_stage_bwd_edges_items":
        return (U + 2 * I) * R + 2 * sc + 2 * E * 4 + E * H * 4 + (E if dropout else 0)
    if name == "ngacf_transform_bwd":
        return 4 * N * R + sc + (N * 8 if dropout else 0)                       # dh, X, h -> dX ; dS
    if name == "ngacf_adam_step_dev":
        return 7 * n_params * 4
    if name == "ngacf_score_pairs_bwd":
        return 3 * 2 * B * R
    if name == "ngacf_score_pairs":
        return 2 * B * R
    raise KeyError(name)


def step_bytes_compulsory(U: int, I: int, E: int, S: int = 2, B: int = 2048) -> int:
    """fwd_stage = 5 N r + 12 E, bwd_stage = 9 N r + 12 E,
    step = 2 S (fwd + bwd) + 7 N r (Adam) + 6 B r (BPR gather + scatter)."""
    N = U + I
    fwd = 5 * N * R + 12 * E
    bwd = 9 * N * R + 12 * E
    return 2 * S * (fwd + bwd) + 7 * N * R + 6 * B * R


def eval_flops(n_users: int, I: int, D: int = 64) -> int:
    return 2 * n_users * I * D
