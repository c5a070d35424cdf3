"""L2-PGD attack on the amortized posterior, mirroring ``rbi.attacks.L2PGDAttack``
(hydra: ``attack_module: rabi_b200.attacks``, ``attack_class: L2PGDAttack``), plus the random-noise "attacks"
(rbi/rbi/attacks/noise_attacks.py:12-69) that the noise-augmentation defenses and the l2noise baseline use.

Reference behaviour reproduced (rbi/rbi/attacks/base.py:9-84, attacks/advertorch_attack.py:23-42 and advertorch's
``PGDAttack.perturb`` / ``perturb_iterative(ord=2)`` / ``rand_init_delta``):
  * untargeted: the target is ``predict(x)`` evaluated under ``no_grad``;
  * clip_min / clip_max default to -1000 / 1000 unless passed as keywords (``OverWriteDefaults``);
  * delta0 = U(clip_min, clip_max) - x, projected onto the L2 eps-ball, then box-clamped (the reference's
    ``L2PGDAttack`` resolves to advertorch's own ``perturb``; see oracle/thirdparty/advertorch_restated.py);
  * every iteration: loss = loss_fn(predict(x + delta), y) (mean over the batch, Q2), gradient normalised with the
    1e-6 floor, step, box clamp, L2 projection; returns clamp(x + delta).
With ``loss_fn = ReverseKLLoss`` and an identity(+z-score) embedding the whole loop runs on the fused kernels
(rabi_rkl_pgd_run): no autograd graph, no per-op launches.
"""
from __future__ import annotations

from typing import Any, Callable, Optional

import torch
from torch import Tensor

from . import noise
from .loss import ForwardKLLoss, ReverseKLLoss
from .models import InverseAffineAutoregressiveModel, MAFPosterior


class Attack:
    def __init__(self, predict: Callable, loss_fn: Optional[Callable] = None, targeted: bool = False,
                 clip_min: float = -1e3, clip_max: float = 1e3, *args, **kwargs) -> None:
        self.predict = predict
        self.loss_fn = loss_fn
        self.clip_min = clip_min
        self.clip_max = clip_max
        self.targeted = targeted

    def _get_predicted_label(self, x: Tensor) -> Any:
        with torch.no_grad():
            return self.predict(x)

    def _verify_and_process_inputs(self, x: Tensor, y: Any):
        if self.targeted:
            assert y is not None
        elif y is None:
            y = self._get_predicted_label(x)
        if isinstance(x, torch.Tensor):
            x = x.detach().clone()
        if isinstance(y, torch.Tensor):
            y = y.detach().clone()
        return x, y

    def perturb(self, x: Tensor, target: Optional[Any] = None, **kwargs):
        raise NotImplementedError("This attack is not implemented...")

    def __call__(self, *args, **kwargs):
        return self.perturb(*args, **kwargs)


def rand_init_delta_l2(x: Tensor, eps: float, clip_min: float, clip_max: float) -> Tensor:
    """advertorch ``rand_init_delta`` (ord=2) followed by the box clamp of ``PGDAttack.perturb``."""
    delta = torch.empty_like(x).uniform_(clip_min, clip_max) - x
    norm = delta.abs().pow(2).sum(dim=1).pow(0.5)
    delta = delta * torch.min(eps / norm, torch.ones_like(norm))[:, None]   # eps: float or a [B] tensor (eps-sweep)
    delta = torch.clamp(x + delta, clip_min, clip_max) - x
    return torch.clamp(x + delta, clip_min, clip_max) - x


def _scalar(v) -> float:
    """eps / eps_iter as the scalar argument of the C ABI (ignored by the kernels while per-row scales are set)."""
    return float(v.max()) if isinstance(v, Tensor) else float(v)


class L2PGDAttack(Attack):
    def __init__(self, predict, loss_fn=None, eps: float = 0.3, nb_iter: int = 40, eps_iter: float = 0.01,
                 rand_init: bool = True, clip_min: float = -1000.0, clip_max: float = 1000.0,
                 targeted: bool = False, **kwargs):
        super().__init__(predict, loss_fn, targeted, clip_min, clip_max)
        if loss_fn is None:
            raise ValueError("L2PGDAttack on a posterior needs a divergence loss_fn (e.g. ReverseKLLoss)")
        self.eps = eps
        self.nb_iter = nb_iter
        self.eps_iter = eps_iter
        self.rand_init = rand_init
        self.ord = 2
        self.l1_sparsity = None

    # -- dispatch ---------------------------------------------------------------------------------------
    def _fused_ok(self, y) -> bool:
        m = self.predict
        if not isinstance(m, InverseAffineAutoregressiveModel) or not isinstance(y, MAFPosterior) or y.model is not m:
            return False
        return type(self.loss_fn) in (ReverseKLLoss, ForwardKLLoss) and self.loss_fn.reduction == "mean"

    def perturb(self, x: Tensor, y: Optional[Any] = None, delta_init: Optional[Tensor] = None,
                chunk_size: Optional[int] = None, feature_mask: Optional[Tensor] = None) -> Tensor:
        """x [B, dx] -> x_adv [B, dx] (detached).

        ``delta_init`` injects delta0 (parity tests); ``chunk_size`` makes the loss mean (Q2: the 1/B factor in
        front of every row gradient) behave as if rows were attacked in independent chunks of that size, so a
        whole sharded data set can go through in one call.  ``feature_mask`` (bool [dx]): only these features are attacked
        (``RobustnessMetric(mask=...)``, rbibm/rbibm/metric/base.py:232-250 attacks the sub-vector x[..., mask] through a wrapper
        model): the perturbation, its L2 ball and the clip box live on the masked columns, the others stay exactly x.
        """
        x, y = self._verify_and_process_inputs(x, y)
        if x.dim() != 2:
            raise ValueError("perturb expects x of shape [B, dx]")
        if not self._fused_ok(y):
            raise NotImplementedError(
                "L2PGDAttack supports predict=InverseAffineAutoregressiveModel (any embedding net) with "
                f"loss_fn=ReverseKLLoss / ForwardKLLoss (reduction='mean'); got loss_fn={type(self.loss_fn).__name__}")
        model: InverseAffineAutoregressiveModel = self.predict
        eng = model.engine()
        x = x.float().contiguous()
        B, dx = x.shape
        per_row = isinstance(self.eps, Tensor) and self.eps.dim() == 1   # batched eps-sweep: one radius per row
        if per_row and self.eps.shape[0] != B:
            raise ValueError(f"per-row eps has {self.eps.shape[0]} entries for {B} rows")
        eps0 = self.eps.to(x) if per_row else float(self.eps)
        fm = None
        if feature_mask is not None:
            fm = torch.as_tensor(feature_mask, device=x.device).bool().reshape(-1)
            if fm.numel() != dx:
                raise ValueError(f"feature_mask has {fm.numel()} entries for {dx} features")
        if delta_init is not None:
            delta = delta_init.detach().to(x).clone().contiguous()
            if fm is not None and delta.shape[1] == int(fm.sum()):      # given on the sub-vector, like the reference's attack sees it
                full = torch.zeros_like(x)
                full[:, fm] = delta
                delta = full
        elif self.rand_init:
            if fm is None:
                delta = rand_init_delta_l2(x, eps0, float(self.clip_min), float(self.clip_max)).contiguous()
            else:                                                       # drawn in the subspace the attack lives in
                delta = torch.zeros_like(x)
                delta[:, fm] = rand_init_delta_l2(x[:, fm].contiguous(), eps0, float(self.clip_min), float(self.clip_max))
        else:
            delta = torch.zeros_like(x)
        if fm is not None:
            delta = (delta * fm.to(delta.dtype)).contiguous()
        S = int(self.loss_fn.mc_samples)
        D = model.output_dim
        if noise._queue is not None:
            z_all = torch.stack([noise.standard_normal((S, B, D), x.device) for _ in range(self.nb_iter)]) \
                if self.nb_iter > 0 else torch.empty((0, S, B, D), device=x.device)
        else:
            z_all = noise.standard_normal((self.nb_iter, S, B, D), x.device)
        inv_B = 1.0 / float(min(B, chunk_size) if chunk_size else B)
        if self.targeted:
            inv_B = -inv_B  # perturb_iterative(minimize=True): descend on the loss
        from .models import _split_embedding
        if per_row:
            step = self.eps_iter.to(x) if isinstance(self.eps_iter, Tensor) else torch.full_like(eps0, float(self.eps_iter))
            eng.set_row_scales(eps0, step)
        if fm is not None:
            eng.set_column_mask(fm.to(torch.float32))
        try:
            if _split_embedding(model.embedding_net)[1] is not None:
                x_adv = self._run_embedded(model, eng, x, delta, y, z_all, inv_B).detach()
                return torch.where(fm, x_adv, x) if fm is not None else x_adv
            _, zs, _ = model.condition_inputs(x)
            zs = zs if zs is not None else (None, None)
            # A_q: the target's projection (== projection of the clean x when untargeted)
            x_adv = self._run(eng, x, delta, zs, y, z_all, inv_B)
            if fm is not None:
                x_adv = torch.where(fm, x_adv, x)
            return x_adv.detach()
        finally:
            if per_row:
                eng.set_row_scales()
            if fm is not None:
                eng.set_column_mask()

    def perturb_sweep(self, x: Tensor, eps_values, eps_iter_values=None, delta_init: Optional[Tensor] = None,
                      chunk_size: Optional[int] = None) -> Tensor:
        """The attack at SEVERAL radii in one call: x [B, dx] -> x_adv [n_eps, B, dx].

        The reference sweeps eps with one process per value (config/experiment/eval_l2_pyloric.yaml:13-16, hydra multirun);
        rows x radii are independent, so they are stacked into one batch of n_eps * B rows with a per-row radius / step size
        (eps_iter "auto" per value as run_rob_eval.py:137-140 when ``eps_iter_values`` is None).  ``chunk_size`` defaults to B, so
        every row sees the 1/B loss scaling of its own single-eps run."""
        eps_values = [float(e) for e in eps_values]
        n, B = len(eps_values), x.shape[0]
        if eps_iter_values is None:
            eps_iter_values = [min(e / max(self.nb_iter, 1) * 20, e) for e in eps_values]
        eps_rows = torch.tensor(eps_values, device=x.device, dtype=torch.float32).repeat_interleave(B)
        step_rows = torch.tensor([float(e) for e in eps_iter_values], device=x.device, dtype=torch.float32).repeat_interleave(B)
        saved = (self.eps, self.eps_iter)
        self.eps, self.eps_iter = eps_rows, step_rows
        try:
            out = self.perturb(x.repeat(n, 1), delta_init=delta_init, chunk_size=chunk_size or B)
        finally:
            self.eps, self.eps_iter = saved
        return out.reshape(n, B, -1)

    def _run_embedded(self, model, eng, x, delta, y, z_all, inv_B):
        """x-space PGD through a non-identity embedding net (e.g. pyloric's CNN): every iteration the embedding runs in
        torch, the flow part (divergence and its gradient w.r.t. the embedded context h) runs on the fused kernels, the
        gradient is chained back through the embedding by autograd and the step / projection runs on the update
        kernel.  Same arithmetic as perturb_iterative(ord=2) on ``loss_fn(predict(x + delta), y)``."""
        fkl = type(self.loss_fn) is ForwardKLLoss
        for it in range(z_all.shape[0]):
            xd = (x + delta).requires_grad_()
            with torch.enable_grad():
                h = model.embedding_net(xd)
            gh, _ = eng.rkl_input_grad(h.detach().float().contiguous(), None, None, y.A, z_all[it], inv_B,
                                       want_kl=False, forward_kl=fkl)
            (gx,) = torch.autograd.grad(h, xd, gh)
            eng.pgd_update(x, delta, gx.contiguous(), _scalar(self.eps), _scalar(self.eps_iter), self.clip_min, self.clip_max)
        return torch.clamp(x + delta, self.clip_min, self.clip_max)

    def _run(self, eng, x, delta, zs, y, z_all, inv_B):
        return eng.pgd_run(x, delta, zs[0], zs[1], y.A, z_all, _scalar(self.eps), _scalar(self.eps_iter), self.clip_min, self.clip_max,
                           inv_B, forward_kl=type(self.loss_fn) is ForwardKLLoss)


# ---- random perturbations (no model evaluation; elementwise host-side torch) --------------------------------------------
def sample_lp_uniformly(N: int, d: int, p=2.0, eps=0.5, device="cuda") -> Tensor:
    """N points uniform in the d-dimensional l_p ball of radius eps (rbi/rbi/utils/distributions.py:22-43): direction
    from i.i.d. generalised-Gamma(1/p, p) magnitudes with random signs, normalised in l_p; radius eps * U^(1/d)."""
    if p in ("inf", float("inf")):
        return torch.rand(N, d, device=device) * eps * 2 - eps
    if not isinstance(p, (int, float)):
        raise ValueError("Unknown order...")
    conc = torch.full((d,), 1.0 / p, device=device)
    mag = torch.distributions.Gamma(conc, 1.0).sample((N,)) ** (1.0 / p)
    direction = mag * (torch.rand((N, d), device=device) - 0.5).sign()
    radius = torch.rand(N, 1, device=device) ** (1.0 / d)
    return eps * radius * (direction / torch.linalg.norm(direction, keepdim=True, ord=p, dim=-1))


class UniformNoiseAttack(Attack):
    """x + noise, noise uniform in the l_p eps-ball, clipped to the box."""

    def __init__(self, predict: Callable, eps: float = 0.5, ord=2.0, **kwargs) -> None:
        super().__init__(predict, **kwargs)
        self.eps = eps
        self.ord = ord

    def perturb(self, x: Tensor, *args, **kwargs) -> Tensor:
        with torch.no_grad():
            noise_ = sample_lp_uniformly(x.shape[0], x.shape[-1], p=self.ord, eps=self.eps, device=x.device)
            return torch.clip(x + noise_, min=self.clip_min, max=self.clip_max)


class L2UniformNoiseAttack(UniformNoiseAttack):
    def __init__(self, predict, eps=0.5, **kwargs) -> None:
        super().__init__(predict, eps, ord=2.0, **kwargs)


class L1UniformNoiseAttack(UniformNoiseAttack):
    def __init__(self, predict, eps=0.5, **kwargs) -> None:
        super().__init__(predict, eps, ord=1.0, **kwargs)


class LinfUniformNoiseAttack(UniformNoiseAttack):
    def __init__(self, predict, eps=0.5, **kwargs) -> None:
        super().__init__(predict, eps, ord="inf", **kwargs)


class GaussianNoiseAttack(Attack):
    def __init__(self, predict, eps=0.5, **kwargs) -> None:
        super().__init__(predict, **kwargs)
        self.eps = eps
        if self.targeted:
            raise ValueError("This attack cannot be targeted")

    def perturb(self, x, **kwargs):
        with torch.no_grad():
            return torch.clip(x + torch.randn_like(x) * self.eps, min=self.clip_min, max=self.clip_max)
