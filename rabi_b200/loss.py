"""Losses / divergences of the hot path, mirroring ``rbi.loss`` (hydra: ``loss_module: rabi_b200.loss``).

* ``reduce`` / ``EvalLoss`` / ``TrainLoss``        rbi/rbi/loss/base.py:13-175
* ``ReverseKLLoss`` / ``ForwardKLLoss``            rbi/rbi/loss/eval_loss.py:108-150
* ``NLLLoss``                                      rbi/rbi/loss/train_loss.py:16-29
* MC KL registered for ``MAFPosterior`` pairs      rbi/rbi/loss/kl_div.py:14-70  (``set_mc_budget``, global budget)

KL(p || q) between two posteriors of the same flow runs as ONE fused kernel (sampling pass under p, density pass
under q, reduction over the S samples); everything else falls back to rsample + log_prob on the same kernels.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Union

import torch
from torch import Tensor
from torch.distributions import Distribution, kl_divergence, register_kl
from torch.nn.modules.loss import _Loss

from . import noise
from .models import MAFPosterior

MC_SAMPLES = 10


def set_mc_budget(n: int):
    global MC_SAMPLES
    MC_SAMPLES = int(n)


def _generic_mc_kl(p: Distribution, q: Distribution) -> Tensor:
    batch_shape = torch.broadcast_shapes(p.batch_shape, q.batch_shape)
    p, q = p.expand(batch_shape), q.expand(batch_shape)
    samples = p.rsample((MC_SAMPLES,)) if p.has_rsample else p.sample((MC_SAMPLES,))
    logq = q.log_prob(samples)
    logp = p.log_prob(samples)
    return torch.mean(logp - logq, 0)


@register_kl(MAFPosterior, MAFPosterior)
def _kl_maf_maf(p: MAFPosterior, q: MAFPosterior) -> Tensor:
    batch_shape = torch.broadcast_shapes(p.batch_shape, q.batch_shape)
    p, q = p.expand(batch_shape), q.expand(batch_shape)
    same_flow = p.model is q.model and p.output_transform is q.output_transform
    if not same_flow or p._needs_grad() or q._needs_grad():
        return _generic_mc_kl(p, q)
    # the output transform (if any) is shared, so its log-dets cancel: KL is evaluated in unconstrained space
    D = p.event_shape[0]
    Bf = int(batch_shape.numel())
    z = noise.standard_normal((MC_SAMPLES,) + tuple(batch_shape) + (D,), p.ctx_in.device).reshape(MC_SAMPLES, Bf, D)
    if p is q:
        # same conditioned object: the reference's log p and log q both come from the transform cache and cancel
        # exactly (rbi/tests/test_loss_fn.py:105-111 relies on KL(Q, Q) == 0)
        return torch.zeros(batch_shape, device=z.device)
    kl = p.model.engine().rkl(p.A, q.A, z)
    return kl.reshape(batch_shape)


@register_kl(MAFPosterior, Distribution)
def _kl_maf_any(p, q):
    return _generic_mc_kl(p, q)


@register_kl(Distribution, MAFPosterior)
def _kl_any_maf(p, q):
    return _generic_mc_kl(p, q)


def reduce(loss: Tensor, reduction: Optional[str], weight: Optional[Union[Tensor, float]] = None) -> Tensor:
    if weight is None:
        weight = 1.0
    if reduction == "mean":
        return torch.mean(weight * loss)
    if reduction == "sum":
        return torch.sum(weight * loss)
    if reduction == "median":
        return torch.median(weight * loss)
    if reduction == "min":
        return torch.min(weight * loss)
    if reduction == "max":
        return torch.max(weight * loss)
    if reduction is None:
        return loss
    raise ValueError("We only support the reduction operations: mean, sum, median, min, max and None.")


class Loss(_Loss):
    def __init__(self, reduction: Optional[str] = "mean") -> None:
        super().__init__(None, None, reduction)  # type: ignore


class EvalLoss(Loss):
    def _loss(self, output, target) -> Tensor:
        raise NotImplementedError("Loss not implemented")

    def forward(self, output, target) -> Tensor:
        return reduce(self._loss(output, target), self.reduction)


class ReverseKLLoss(EvalLoss):
    """D_KL(output || target): samples are drawn from ``output``."""

    def __init__(self, mc_samples: int = 10, reduction: Optional[str] = "mean", **kwargs) -> None:
        super().__init__(reduction)
        self.mc_samples = mc_samples

    def _loss(self, output: Distribution, target: Distribution) -> Tensor:
        if not isinstance(target, Distribution):
            raise ValueError(r"The target must be a distribution...")
        set_mc_budget(self.mc_samples)
        kl_div = kl_divergence(output, target)
        return torch.mean(kl_div) if self.reduction == "mean" else kl_div


class ForwardKLLoss(EvalLoss):
    """D_KL(target || output): samples are drawn from ``target``."""

    def __init__(self, reduction: Optional[str] = "mean", mc_samples: int = 10, **kwargs) -> None:
        super().__init__(reduction)
        self.mc_samples = mc_samples

    def _loss(self, output: Distribution, target: Distribution) -> Tensor:
        if not isinstance(target, Distribution):
            raise ValueError(r"The target must be a distribution...")
        set_mc_budget(self.mc_samples)
        return kl_divergence(target, output)


class NegativeLogLikelihoodLoss(EvalLoss):
    def _loss(self, output: Distribution, target: Tensor) -> Tensor:
        if not isinstance(target, Tensor):
            raise ValueError("This loss function requires events y ~ p(y|x) given as tensors.")
        sample_shape = output.batch_shape + output.event_shape
        return -output.log_prob(target.float().reshape(*sample_shape))


class TrainLoss(Loss):
    """Training objective with pre-/post-loss regulariser registries (defenses hook in here)."""

    def __init__(self, model, reduction: Optional[str] = "mean") -> None:
        super().__init__(reduction)
        self.model = model
        self._pre_loss_regularizers: Dict[str, Callable] = {}
        self._post_loss_regularizers: Dict[str, Callable] = {}

    @property
    def pre_loss_regularizers(self) -> Dict:
        return self._pre_loss_regularizers

    @property
    def post_loss_regularizers(self) -> Dict:
        return self._post_loss_regularizers

    def register_post_loss_regularizer(self, name: str, regularizer: Callable) -> None:
        self._post_loss_regularizers[name] = regularizer

    def remove_post_loss_regularizer(self, name: str) -> None:
        del self._post_loss_regularizers[name]

    def register_pre_loss_regularizer(self, name: str, regularizer: Callable) -> None:
        self._pre_loss_regularizers[name] = regularizer

    def remove_pre_loss_regularizer(self, name: str) -> None:
        del self._pre_loss_regularizers[name]

    def forward(self, input: Tensor, target: Tensor) -> Tensor:
        loss = 0
        if self.training:
            if len(self.pre_loss_regularizers) == 0:
                output = self.model(input)
                loss += self._loss(output, target, input)
            else:
                for pre_regularizer in self.pre_loss_regularizers.values():
                    input, target = pre_regularizer(input, target)
                    output = self.model(input)
                    loss += self._loss(output, target, input) / len(self.pre_loss_regularizers)
            if len(self.post_loss_regularizers) == 0:
                # nothing to add: the reference's "+ (zeros(1) - zeros(1))" only broadcasts the result to at least one dimension
                out = reduce(loss, self.reduction, None)
                return out.reshape(1) if out.dim() == 0 else out
            post_reg = torch.zeros(1, device=input.device)
            for post_regularizer in self.post_loss_regularizers.values():
                post_reg += post_regularizer(input, output, target)
            # value unchanged, gradient of the regulariser added (loss/base.py:171)
            return reduce(loss, self.reduction, None) + (post_reg - post_reg.detach())
        output = self.model(input)
        loss += self._loss(output, target, input)
        return reduce(loss, self.reduction, None)

    def _loss(self, output, target, input) -> Tensor:
        raise NotImplementedError


class NLLLoss(TrainLoss):
    """-log q(theta | x) per row (rbi/rbi/loss/train_loss.py:16-29).  ``cuda_graph``: after two eager calls of a batch
    shape, forward and parameter-gradient computation are replayed as CUDA graphs (rabi_b200/graphs.py)."""

    def __init__(self, model, reduction: Optional[str] = "mean", cuda_graph: bool = True) -> None:
        super().__init__(model, reduction)
        self.cuda_graph = cuda_graph
        self._graphed = None

    def _nll(self, x: Tensor, theta: Tensor) -> Tensor:
        q = self.model(x)
        return -q.log_prob(theta.reshape(*(q.batch_shape + q.event_shape)))

    def _loss(self, output: Distribution, target: Tensor, input: Tensor) -> Tensor:
        if not isinstance(target, Tensor):
            raise ValueError("This loss function requires events y ~ p(y|x) given as tensors.")
        if (self.cuda_graph and self.training and torch.is_grad_enabled() and isinstance(output, MAFPosterior)
                and input.is_cuda and input.dim() == 2 and target.dim() == 2 and not input.requires_grad
                and not target.requires_grad and noise._queue is None):
            if self._graphed is None or self._graphed.module is not self.model:
                from .graphs import GraphedParamFunction
                self._graphed = GraphedParamFunction(self.model, self._nll)
            if self._graphed.ready((input, target)):
                return self._graphed(input, target.to(input.dtype))
        sample_shape = output.batch_shape + output.event_shape
        return -output.log_prob(target.reshape(*sample_shape))
