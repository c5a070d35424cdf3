"""Metrics of the hot path, mirroring ``rbibm.metric`` (hydra: ``metric_module: rabi_b200.metric``).

* ``Metric`` / ``summary_reduce``                 rbibm/rbibm/metric/base.py:27-170
* ``RobustnessMetric``                            rbibm/rbibm/metric/base.py:173-355
* ``ReverseKLRobMetric`` / ``ForwardKLRobMetric`` rbibm/rbibm/metric/robustness_metric.py:54-155
* ``NegativeLogLikelihoodMetric`` / ``ExpectedCoverageMetric``   rbibm/rbibm/metric/approximation_metric.py:205-324

Robustness metrics, for every observation row: x~ = attack.perturb(x); loss = D(q(.|x~), q(.|x)) with ``mc_samples`` MC
samples (``eval_mc_budget``, 256 in config/eval_rob/metric/rKL.yaml); losses below -1e4 are replaced by the median; the
best of ``attack_attemps`` is kept; the summary is mean + quantiles over the finite values.

B200-first difference in *scheduling only*: the reference walks the data in ``batch_size`` chunks to bound memory
(rbibm/utils/batched_processing.py:5-15).  Rows are independent except for the 1/batch factor in the attack
gradient, so here all rows of a shard go through ONE call with that factor passed explicitly
(``chunk_size``); results per row are those of the chunked walk (overflowed estimates are replaced by the median of their own
chunk, as there).  One stated deviation: a last, PARTIAL chunk is attacked with 1/chunk in front of its row gradients where the
reference has 1/len(partial chunk); the factor cancels in the normalised step and only moves advertorch's 1e-6 gradient-norm
floor, and the reference itself cannot evaluate ragged chunk splits (its ``vstack`` of per-chunk losses fails).
"""
from __future__ import annotations

from typing import Any, Callable, Optional, Tuple

import torch
from torch import Tensor

from .attacks import Attack, L2PGDAttack
from .loss import EvalLoss, ForwardKLLoss, NegativeLogLikelihoodLoss, ReverseKLLoss


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


class Metric:
    """Common surface of every metric: ``eval(x, other)`` -> reduced value (metric/base.py:27-170)."""

    def __init__(self, model, reduction: Optional[str] = None, batch_size: Optional[int] = None,
                 descending: bool = False, device: str = "cuda") -> None:
        self.model = model
        self.batch_size = batch_size
        self.descending = descending
        self.device = device
        self._reduction = reduction  # not validated at construction (metric/base.py:47), only by the setter

    @property
    def reduction(self):
        return self._reduction

    @reduction.setter
    def reduction(self, x: Optional[str]) -> None:
        self._set_reduction(x)

    def _set_reduction(self, x: Optional[str]) -> None:
        if x is None:
            self._reduction = None
            return
        if not isinstance(x, str):
            raise ValueError("Unknown reduction type")
        parts = x.split("_")
        main_ok = parts[0] in ("mean", "median", "sum", "max", "min")
        supp_ok = parts[1] in ("std", "q95", "summary") if len(parts) > 1 else True
        if not (main_ok and supp_ok):
            raise ValueError("Unknown reduction type")
        self._reduction = x

    def _reduce(self, m: Tensor):
        return summary_reduce(m, self.reduction)

    def _eval(self, x, other=None) -> Tensor:
        raise NotImplementedError

    def eval(self, x: Tensor, other: Any = None):
        if x.ndim == 1:
            x = x.unsqueeze(0)
        x = x.to(self.device)
        if other is not None and hasattr(other, "to"):
            other = other.to(self.device)
        return self._reduce(self._eval(x, other))


class RobustnessMetric(Metric):
    """max_d D(q(x + d), q(x))  (or min_d D(q(x + d), t) for a target t)."""

    def __init__(self, model, loss_fn: EvalLoss, attack: Attack, attack_attemps: int = 5, targeted: bool = False,
                 target_wrapper: Callable = lambda x: x, mask: Optional[Tensor] = None, descending: bool = True,
                 batch_size: Optional[int] = None, reduction: Optional[str] = None, device: str = "cuda"):
        if mask is not None:
            mask = torch.as_tensor(mask).bool()
            if not isinstance(attack, L2PGDAttack):
                raise NotImplementedError("masked-feature attacks are served for L2PGDAttack")
        super().__init__(model, reduction=reduction, batch_size=batch_size, descending=descending, device=device)
        self.loss_fn = loss_fn
        self.loss_fn.reduction = None
        self.attack_attemps = attack_attemps
        self.targeted = targeted
        self.target_wrapper = target_wrapper
        self.attack = attack
        self.mask = mask
        self._cached_x = None
        self._cached_xs_tilde = None
        self._cached_losses = None

    def _get_goal(self, x, target):
        raise NotImplementedError

    @staticmethod
    def _ensure_loss_constraint(loss: Tensor) -> Tensor:
        return loss

    def _eval(self, x: Tensor, target: Optional[Any] = None) -> Tensor:
        target = self.target_wrapper(target)
        losses = None
        xs_tilde = None
        # chunk_size: the reference attacks `batch_size` rows at a time; a loss with reduction="mean" then puts a
        # 1/batch_size factor in front of every row's gradient before advertorch's 1e-6 norm floor.
        chunk = self.batch_size if (self.batch_size is not None and self.batch_size < x.shape[0]) else None
        for i in range(self.attack_attemps):
            kw = {"chunk_size": chunk} if isinstance(self.attack, L2PGDAttack) else {}
            if self.mask is not None:   # metric/base.py:232-250: only the masked features are attacked
                kw["feature_mask"] = self.mask
            if target is None or not self.attack.targeted:
                x_tilde = self.attack.perturb(x, **kw)
            else:
                x_tilde = self.attack.perturb(x, target, **kw)
            with torch.no_grad():
                q_tilde = self.model(x_tilde)
                q = self._get_goal(x, target)
                loss = self.loss_fn(q_tilde, q)
                if chunk is None:
                    loss = self._ensure_loss_constraint(loss)
                else:   # the reference repairs overflowed estimates inside every ``batch_size`` chunk (median of THAT chunk)
                    loss = torch.cat([self._ensure_loss_constraint(l) for l in torch.split(loss, chunk)])
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

    def eval_sweep(self, x: Tensor, eps_values, eps_relative_to_std: bool = True) -> dict:
        """The metric at several attack radii from ONE batched attack (rows x radii are independent): what the reference
        obtains from one ``run_rob_eval`` process per eps (config/experiment/eval_l2_pyloric.yaml:13-16).  eps is relative to
        ``x.std(1).mean()`` as in run_rob_eval.py:115-121, eps_iter "auto" per value (:137-140); the attack keeps its other
        settings.  Returns {eps: reduced value}; untargeted L2PGDAttack only."""
        if self.targeted or not isinstance(self.attack, L2PGDAttack):
            raise NotImplementedError("eval_sweep batches the untargeted L2-PGD attack")
        x = x.to(self.device)
        scale = float(x.std(1).mean()) if eps_relative_to_std else 1.0
        eps_abs = [float(e) * scale for e in eps_values]
        B = x.shape[0]
        chunk = self.batch_size if (self.batch_size is not None and self.batch_size < B) else B
        best = None
        for _ in range(self.attack_attemps):
            xs_t = self.attack.perturb_sweep(x, eps_abs, chunk_size=chunk)              # [n_eps, B, dx]
            with torch.no_grad():
                loss = self.loss_fn(self.model(xs_t.reshape(-1, x.shape[1])), self._get_goal(x.repeat(len(eps_abs), 1), None))
                loss = torch.stack([self._ensure_loss_constraint(l) for l in loss.reshape(len(eps_abs), B)])
            best = loss if best is None else (torch.maximum(best, loss) if self.descending else torch.minimum(best, loss))
        return {float(e): self._reduce(best[i]) for i, e in enumerate(eps_values)}

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


def _median_for_overflow(loss: Tensor) -> Tensor:
    # robustness_metric.py:93-99 / 148-153: overflowed (very negative) MC estimates are replaced by the median
    mask_neg = loss < -10000
    if mask_neg.any():
        loss[mask_neg] = loss[~mask_neg].median()
    return loss


class ReverseKLRobMetric(RobustnessMetric):
    """D_KL(q(.|x~) || q(.|x)), samples from the perturbed posterior (config/eval_rob/metric/rKL.yaml)."""

    def __init__(self, model, attack: Attack, targeted: bool = False, target_wrapper: Callable = lambda x: x,
                 attack_attemps: int = 5, device: str = "cuda", mask: Optional[Tensor] = None,
                 reduction: Optional[str] = "mean_quantile95", batch_size: Optional[int] = None,
                 descending: bool = True, **kwargs):
        super().__init__(model, loss_fn=ReverseKLLoss(**kwargs), attack=attack, attack_attemps=attack_attemps,
                         targeted=targeted, target_wrapper=target_wrapper, device=device, mask=mask,
                         batch_size=batch_size, reduction=reduction, descending=descending)

    def _get_goal(self, x, target):
        if self.targeted:
            return target
        with torch.no_grad():
            return self.model(x)

    _ensure_loss_constraint = staticmethod(_median_for_overflow)


class ForwardKLRobMetric(RobustnessMetric):
    """D_KL(q(.|x) || q(.|x~)), samples from the clean posterior (config/eval_rob/metric/fKL.yaml)."""

    def __init__(self, model, attack: Attack, targeted: bool = False, target_wrapper: Callable = lambda x: x,
                 attack_attemps: int = 5, device: str = "cuda", mask: Optional[Tensor] = None,
                 reduction: Optional[str] = "mean_quantile95", batch_size: Optional[int] = None,
                 descending: bool = True, empirical_approx: bool = False, empirical_mc_samples=20, **kwargs):
        if empirical_approx:
            raise NotImplementedError("empirical_approx (EmpiricalDistribution targets) is outside the maf_pyro path")
        self.empirical_approx = empirical_approx
        self.empirical_mc_samples = empirical_mc_samples
        super().__init__(model, loss_fn=ForwardKLLoss(**kwargs), attack=attack, attack_attemps=attack_attemps,
                         targeted=targeted, target_wrapper=target_wrapper, device=device, mask=mask,
                         batch_size=batch_size, reduction=reduction, descending=descending)

    def _get_goal(self, x, target):
        if self.targeted or self.attack.targeted:
            return target
        with torch.no_grad():
            return self.model(x)

    _ensure_loss_constraint = staticmethod(_median_for_overflow)


# ---- approximation metrics on clean / adversarial x (rbibm_script.py:232-257) -----------------------------------------
class ApproximationMetric(Metric):
    requires_posterior = False
    requires_thetas = False
    requires_potential = False
    can_eval_single = True
    can_eval_x_xtilde = False

    def __init__(self, model, prior=None, simulator=None, ground_truth=None, potential_fn=None,
                 reduction: Optional[str] = "mean", batch_size: Optional[int] = None, descending: bool = False,
                 device: str = "cuda") -> None:
        super().__init__(model, reduction=reduction, batch_size=batch_size, descending=descending, device=device)
        self.prior = prior
        self.simulator = simulator
        self.ground_truth = ground_truth
        self.potential_fn = potential_fn


class NegativeLogLikelihoodMetric(ApproximationMetric):
    """-log q(theta | x) per row.  All rows go through one log-prob launch; the reference's ``batch_size`` walk only
    bounds memory and does not change any row's value."""
    requires_thetas = True

    def __init__(self, model, reduction: Optional[str] = "mean", batch_size: Optional[int] = None,
                 device: str = "cuda", **kwargs) -> None:
        self.loss_fn = NegativeLogLikelihoodLoss(reduction=None)
        super().__init__(model, reduction=reduction, batch_size=batch_size, device=device)

    def _eval(self, xs: Tensor, target: Optional[Tensor]) -> Tensor:
        with torch.no_grad():
            return self.loss_fn(self.model(xs), target)


class ExpectedCoverageMetric(ApproximationMetric):
    """Expected coverage of the highest-density regions: for every level alpha the fraction of rows whose true theta has
    a higher log q than the (1-alpha)-quantile of ``mc_samples`` posterior samples' log q.  Per ``batch_size`` chunk one
    [alphas] vector (the chunk mean), then the mean over chunks -- the reference's ``vstack`` + ``mean(0)``
    (approximation_metric.py:275-324, batched_processing.py:5-15).  Sampling and log q of the samples are one fused
    launch (the "free" log-prob of a fresh sample)."""
    requires_thetas = True
    can_eval_single = False

    def __init__(self, model, mc_samples: int = 100, alphas: Tensor = None, device: str = "cuda",
                 batch_size: int = 1000, reduction: Optional[str] = "absAUC", **kwargs):
        super().__init__(model=model, device=device, batch_size=batch_size, reduction=reduction)
        self.mc_samples = mc_samples
        self.alphas = (torch.linspace(0, 1, 100) if alphas is None else alphas).to(device)

    def _set_reduction(self, x):
        if x not in ("AUC", "absAUC", "AUC_full", "absAUC_full", None):
            raise ValueError("Unknown reduction type")
        self._reduction = x

    def _reduce(self, m: Tensor):
        # The reference's branch test (`== "absAUC" or "absAUC_full"`) is always true: every reduction, None included,
        # takes the |coverage - alpha| area branch (approximation_metric.py:281-291).  Reproduced.
        while m.ndim > 1:
            m = m.mean(0)
        val = torch.trapz(torch.abs(m - self.alphas), self.alphas)
        if self._reduction == "absAUC":
            return val
        return val, {"alphas": self.alphas.cpu(), "quantiles": m.cpu()}

    def eval(self, x: Tensor, other: Any = None):
        x = x.to(self.device)
        other = other.to(self.device)
        if self.batch_size is not None and self.batch_size < x.shape[0]:
            out = torch.vstack([self._eval(xb, tb) for xb, tb in
                                zip(torch.split(x, self.batch_size), torch.split(other, self.batch_size))])
        else:
            out = self._eval(x, other)
        return self._reduce(out)

    def coverage_curve(self, logq_theta: Tensor, logq_samples: Tensor) -> Tensor:
        """All levels at once: sorted sample log-densities [S, B]; the level-alpha threshold is row ``int(S*(1-alpha))``
        of the sort (== ``sorted[cut:].min(0)``); alpha = 0 / 1 are defined as 0 / 1."""
        S = logq_samples.shape[0]
        srt, _ = torch.sort(logq_samples, dim=0)
        out = []
        for a in self.alphas.cpu():
            if float(a) == 0:
                out.append(torch.zeros((), device=logq_theta.device))
            elif float(a) == 1:
                out.append(torch.ones((), device=logq_theta.device))
            else:
                cut = int(self.mc_samples * (1 - a))  # fp32 tensor arithmetic, as in the reference (0.95*40 -> 38)
                out.append((srt[min(cut, S - 1)] < logq_theta).float().mean())
        return torch.stack(out)

    def _eval(self, xs: Tensor, thetas: Tensor) -> Tensor:
        with torch.no_grad():
            q = self.model(xs)
            logq_theta = q.log_prob(thetas)
            samples = q.sample((self.mc_samples,))
            logq_samples = q.log_prob(samples)
            return self.coverage_curve(logq_theta, logq_samples)
