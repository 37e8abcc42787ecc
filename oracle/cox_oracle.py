"""CPU oracle of the survival loss (TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's CPU legs may import this).

Restates ``neg_partial_log_likelihood`` (src/stamp/modeling/models/cox.py:107-268; Cox partial likelihood :19-34, Efron's
ties :37-81, Breslow's :84-104, reduction "mean") as one vectorised fp64 expression over the [times x samples] risk
matrix, differentiable by autograd.  Pinned by tests/golden/cox_loss.npz = the reference function itself (loss and
gradient, oracle/make_golden_cox.py).
"""

from __future__ import annotations

import torch


def neg_partial_log_likelihood(log_hz: torch.Tensor, time: torch.Tensor, event: torch.Tensor, ties_method: str = "efron") -> torch.Tensor:
    s = log_hz.double().flatten()
    t = time.double().flatten()
    e = event.bool().flatten()
    if not bool(e.any()):
        return s.sum() * 0.0
    if ties_method == "breslow":
        u, owners = t[e], e.nonzero().flatten()
        in_h = torch.zeros(len(u), len(t), dtype=torch.bool)
        in_h[torch.arange(len(u)), owners] = True
        tie_weight = torch.zeros(len(u), 1, dtype=torch.float64)          # no correction
    elif ties_method == "efron":
        u = torch.unique(t[e])
        in_h = (t[None, :] == u[:, None]) & e[None, :]
        tie_weight = torch.ones(len(u), 1, dtype=torch.float64)
    else:
        raise ValueError(ties_method)
    at_risk = t[None, :] >= u[:, None]
    w = torch.exp(s)
    D = (at_risk * w).sum(1, keepdim=True)
    T = (in_h * w).sum(1, keepdim=True) * tie_weight
    m = in_h.sum(1, keepdim=True)
    k = torch.arange(int(m.max()), dtype=torch.float64)[None, :]
    live = k < m
    log_den = torch.where(live, torch.log(torch.where(live, D - k / m * T, torch.ones(()).double())), torch.zeros(()).double())
    terms = (in_h * s).sum(1) - log_den.sum(1)
    return -terms.mean()
