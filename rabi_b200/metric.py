"""Robustness metric of the hot path, mirroring ``rbibm.metric.robustness_metric.ReverseKLRobMetric``
(rbibm/rbibm/metric/base.py:84-170,225-346 and metric/robustness_metric.py:54-99).

For every observation row: x~ = attack.perturb(x); loss = KL(q(.|x~) || q(.|x)) with ``mc_samples`` MC samples
(``eval_mc_budget``, 256 in config/eval_rob/metric/rKL.yaml); losses below -1e4 are replaced by the median; the best
of ``attack_attemps`` is kept; the summary is mean + quantiles over the finite values.

B200-first difference in *scheduling only*: the reference walks the data in ``batch_size`` chunks to bound memory
(rbibm/utils/batched_processing.py:5-15).  Rows are independent except for the 1/batch factor in the attack
gradient, so here all rows of a shard go through ONE call with that factor passed explicitly
(``chunk_size``); results per row are identical to the chunked walk.
"""
from __future__ import annotations

from typing import Any, Callable, Optional, Tuple, Union

import torch
from torch import Tensor

from .attacks import Attack, L2PGDAttack
from .loss import ReverseKLLoss


def summary_reduce(m: Tensor, reduction: Optional[str]):
    """``Metric._reduce`` (metric/base.py:84-144)."""
    if reduction is None:
        return m
    parts = reduction.split("_")
    main_out, supp_out = parts[0], (parts[1] if len(parts) > 1 else None)
    m = m[torch.isfinite(m)]
    if main_out == "mean":
        out = m.mean()
    elif main_out == "median":
        out = m.median()
    elif main_out == "sum":
        out = m.sum()
    elif main_out == "max":
        out = m.max()
    elif main_out == "min":
        out = m.min()
    else:
        out = m
    supp = None
    if supp_out == "std":
        supp = {"std": m.std()}
    elif supp_out == "q95":
        supp = {"q02_5": m.quantile(0.025), "q97_5": m.quantile(0.975)}
    elif supp_out == "summary":
        qs = {"q00_5": 0.005, "q02_5": 0.025, "q05": 0.05, "q15": 0.15, "q20": 0.20, "q30": 0.30, "q50": 0.50,
              "q70": 0.70, "q80": 0.80, "q85": 0.85, "q95": 0.95, "q97_5": 0.975, "q99_5": 0.995}
        supp = {k: m.quantile(v) for k, v in qs.items()}
        supp.update({"min": m.min(), "max": m.max(), "std": m.std()})
    return out if supp is None else (out, supp)


class ReverseKLRobMetric:
    def __init__(self, model, attack: Attack, targeted: bool = False, target_wrapper: Callable = lambda x: x,
                 attack_attemps: int = 5, device: str = "cuda", mask: Optional[Tensor] = None,
                 reduction: Optional[str] = "mean_quantile95", batch_size: Optional[int] = None,
                 descending: bool = True, **kwargs):
        if mask is not None:
            raise NotImplementedError("masked-feature attacks are outside the maf_pyro / l2pgd / rKL path")
        self.model = model
        self.attack = attack
        self.loss_fn = ReverseKLLoss(**kwargs)  # kwargs carries mc_samples (run_rob_eval.py:178-189)
        self.targeted = targeted
        self.target_wrapper = target_wrapper
        self.attack_attemps = attack_attemps
        self.device = device
        self.reduction = reduction
        self.batch_size = batch_size
        self.descending = descending
        self._cached_x = None
        self._cached_xs_tilde = None
        self._cached_losses = None

    # robustness_metric.py:93-99
    @staticmethod
    def _ensure_loss_constraint(loss: Tensor) -> Tensor:
        mask_neg = loss < -10000
        if mask_neg.any():
            loss[mask_neg] = loss[~mask_neg].median()
        return loss

    def _eval(self, x: Tensor, target: Optional[Any] = None) -> Tensor:
        target = self.target_wrapper(target)
        losses = None
        xs_tilde = None
        # chunk_size: the reference attacks `batch_size` rows at a time; ReverseKLLoss(reduction="mean") then puts a
        # 1/batch_size factor in front of every row's gradient before advertorch's 1e-6 norm floor.
        chunk = self.batch_size if (self.batch_size is not None and self.batch_size < x.shape[0]) else None
        for i in range(self.attack_attemps):
            kw = {"chunk_size": chunk} if isinstance(self.attack, L2PGDAttack) else {}
            if target is None or not self.attack.targeted:
                x_tilde = self.attack.perturb(x, **kw)
            else:
                x_tilde = self.attack.perturb(x, target, **kw)
            with torch.no_grad():
                q_tilde = self.model(x_tilde)
                q = target if self.targeted else self.model(x)
                loss = ReverseKLLoss(mc_samples=self.loss_fn.mc_samples, reduction=None)(q_tilde, q)
                loss = self._ensure_loss_constraint(loss)
            if i == 0:
                xs_tilde, losses = x_tilde.detach(), loss.detach()
            else:
                better = (loss > losses) if (self.descending != self.targeted) else (loss < losses)
                xs_tilde[better] = x_tilde[better].detach()
                losses[better] = loss[better].detach()
        self._cached_x = x if self._cached_x is None else torch.vstack([self._cached_x, x])
        self._cached_xs_tilde = xs_tilde if self._cached_xs_tilde is None else torch.vstack([self._cached_xs_tilde, xs_tilde])
        self._cached_losses = losses if self._cached_losses is None else torch.cat([self._cached_losses, losses])
        return losses

    def eval(self, x: Tensor, other: Any = None):
        if x.ndim == 1:
            x = x.unsqueeze(0)
        x = x.to(self.device)
        if other is not None and hasattr(other, "to"):
            other = other.to(self.device)
        out = self._eval(x, other)
        return summary_reduce(out, self.reduction)

    def generate_adversarial_examples(self, x: Tensor, target: Optional[Any] = None,
                                      num_examples: int = 100) -> Tuple[Tensor, Tensor]:
        if self._cached_x is None or self._cached_x.shape != x.shape or \
                not torch.isclose(self._cached_x.to(x.device), x).all():
            self._cached_x = self._cached_xs_tilde = self._cached_losses = None
            self.eval(x, target)
        order_val = (not self.targeted) if self.descending else self.targeted
        _, idx = torch.sort(self._cached_losses.flatten(), descending=order_val)
        idx = idx[torch.isfinite(self._cached_losses.flatten()[idx])]
        return idx[:num_examples], self._cached_xs_tilde[idx][:num_examples]
