"""FusedAdam: torch.optim.Adam semantics (L2 weight decay added to the gradient; run_Gowalla.py:114)
in ONE table-driven kernel launch over all 29 parameter tensors (ngacf_adam_step).  The per-parameter
state uses torch.optim.Adam's keys (step, exp_avg, exp_avg_sq) so `optim.state_dict()` checkpoints are
interchangeable with the reference's (run_Gowalla.py:130,143)."""
from __future__ import annotations

import torch

from . import ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False, foreach=None,
                        capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False)
        super().__init__(params, defaults)
        self._tables = {}

    def _group_table(self, gi, group):
        ps = [p for p in group["params"] if p.grad is not None]
        for p in ps:
            st = self.state[p]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(), self.state[p]["exp_avg_sq"].data_ptr())
                    for p in ps)
        cached = self._tables.get(gi)
        if cached is None or cached[0] != key:
            rows = []
            for p in ps:
                if p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("FusedAdam needs contiguous float32 parameters and gradients")
                st = self.state[p]
                rows.append([p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()])
            tab = torch.tensor(rows, dtype=torch.int64).to(ps[0].device)
            cached = (key, tab, sum(p.numel() for p in ps))
            self._tables[gi] = cached
        return ps, cached[1], cached[2]

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for gi, group in enumerate(self.param_groups):
            if not any(p.grad is not None for p in group["params"]):
                continue
            ps, tab, total = self._group_table(gi, group)
            step = int(self.state[ps[0]]["step"].item()) + 1
            b1, b2 = group["betas"]
            for i in range(0, len(ps), 64):       # the kernel takes up to 64 tensors per launch
                chunk = ps[i:i + 64]
                ops.adam_step(tab[i:i + 64], len(chunk), sum(p.numel() for p in chunk), group["lr"], b1, b2, group["eps"],
                              group["weight_decay"], step)
            for p in ps:
                self.state[p]["step"] = torch.tensor(float(step))
                torch.autograd.graph.increment_version(p)   # the kernel wrote p in place: invalidate cached eval features
        return loss
