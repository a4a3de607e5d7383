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
        return 3 * N * R + sc + (N * 8 if dropout else 0)                       # dh, X -> dX ; dS (da goes through Xd^T dS: h is not read)
    if name == "ngacf_transform_bwd_dx":
        return 3 * N * R + (N * 8 if dropout else 0)                            # dh, Zprev (ELU') -> dX
    if name == "ngacf_transform_bwd_dw":
        return 2 * N * R + sc + (N * 8 if dropout else 0)                       # dh, X (recomputed input), dS
    if name == "ngacf_adam_step_dev":
        return 7 * n_params * 4
    if name == "ngacf_score_pairs_bwd":
        return 3 * 2 * B * R
    if name == "ngacf_score_pairs":
        return 2 * B * R
    raise KeyError(name)


GATHER_TABLE_L2_LIMIT = 63 * 2 ** 20     # above this the gathered table no longer stays L2-resident


def gather_regime(U: int, I: int) -> bool:
    return (U + I) * R > GATHER_TABLE_L2_LIMIT


def kernel_bytes_gather(name: str, U: int, I: int, E: int, H: int, dropout: bool) -> int:
    """Gather model (tables >> L2): every gathered 256-byte row and per-head scalar crosses HBM once per EDGE."""
    N = U + I
    sc = 4 * H
    if name == "ngacf_aggregate_fwd":                     # 2E directed edges: row + scalar + index (+eid, mask)
        return 2 * E * (R + sc + 4 + (5 if dropout else 0)) + 2 * N * R + 2 * N * sc
    if name == "ngacf_stage_bwd_edges_users":             # E edges: Ghat row + h row + s + dN + idx + eid + ds write
        return E * (2 * R + 2 * sc + 8 + sc + (1 if dropout else 0)) + 4 * U * R + 3 * U * sc
    if name == "ngacf_stage_bwd_edges_items":             # E edges: Ghat row + s + idx + eid + ds read
        return E * (R + sc + 8 + sc + (1 if dropout else 0)) + 2 * I * R + 2 * I * sc
    raise KeyError(name)


def step_bytes_compulsory(U: int, I: int, E: int, S: int = 2, B: int = 2048, propagations: int = 2, pairs_per_row: int = 2) -> int:
    """fwd_stage = 5 N r + 12 E, bwd_stage = 9 N r + 12 E,
    step = 2 S (fwd + bwd) + 7 N r (Adam) + 6 B r (BPR gather + scatter).  A NegSampling step (8f-3) has ONE propagation and
    5 (user, item) pairs per train row: propagations=1, pairs_per_row=5."""
    N = U + I
    fwd = 5 * N * R + 12 * E
    bwd = 9 * N * R + 12 * E
    return propagations * S * (fwd + bwd) + 7 * N * R + 3 * pairs_per_row * B * R


def eval_flops(n_users: int, I: int, D: int = 64) -> int:
    return 2 * n_users * I * D
