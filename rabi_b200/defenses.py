"""Defenses of the hot path, mirroring ``rbi.defenses`` (hydra: ``defense_module: rabi_b200.defenses``).

* ``AdditiveRegularizer``            rbi/rbi/defenses/base.py:38-128   (post-loss regulariser registry on TrainLoss)
* ``FIMTraceRegularizer``            rbi/rbi/defenses/fisher_regularization.py:15-95   (config/defense/fisher_trace.yaml)
* ``TradesAdversarialRegularizer`` / ``L2PGDrKLTrades`` / ``L2PGDTrades`` (fKL)
                                     rbi/rbi/defenses/trades_regularizers.py:26-133   (config/defense/l2pgdTrades_rKL.yaml)
* ``DataAugmentationRegularizer``    rbi/rbi/defenses/base.py:131-176   (pre-loss seam of TrainLoss)
* ``AdversarialTraining`` / ``L2PGDAdversarialTraining`` / ``L2UniformNoiseTraining``
                                     rbi/rbi/defenses/adversarial_training.py:21-135   (config/defense/l2pgdAdvTraining.yaml)
The inner L2-PGD of TRADES runs on the fused attack kernels; the regulariser values and their (first/second order)
weight gradients run through rabi_b200/autograd.py.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
from torch import Tensor
from torch.distributions import Distribution
from torch.nn import Module

from .attacks import Attack, L2PGDAttack, L2UniformNoiseAttack
from .fisher import ExponentialMovingAverageEstimator, monte_carlo_fisher
from .loss import EvalLoss, ForwardKLLoss, ReverseKLLoss, TrainLoss


class Defense:
    def __init__(self, model: Module, loss_fn: Callable) -> None:
        self.model = model
        self.loss_fn = loss_fn

    def activate(self, **kwargs):
        raise NotImplementedError("Defense not implemented")

    def deactivate(self, **kwargs):
        raise NotImplementedError("Defense not implemented")


class AdditiveRegularizer(Defense):
    """Adds a regularisation term to the loss through ``TrainLoss.register_post_loss_regularizer``."""

    def __init__(self, model: Module, loss_fn: Callable, reduce: str = "mean") -> None:
        super().__init__(model, loss_fn)
        self.reduce = reduce
        self._regularizer = None
        self._algorithm = None

    @property
    def regularizer(self) -> Callable:
        if self._regularizer is None:
            raise NotImplementedError("Your regularizer is not implemented...")
        return self._regularizer

    @property
    def algorithm(self) -> str:
        if self._algorithm is None:
            raise NotImplementedError("this regulariser has no algorithm selected: call activate(algorithm=...)")
        return self._algorithm

    @algorithm.setter
    def algorithm(self, method):
        self._set_algorithm(method)

    def _set_algorithm(self, method):
        pass

    def _reduce(self, reg: Tensor):
        if self.reduce in ("mean", "sum"):  # the reference returns the mean for "sum" as well (base.py:109-110)
            return torch.mean(reg)
        if self.reduce == "max":
            return torch.max(reg)
        if self.reduce == "min":
            return torch.min(reg)
        raise NotImplementedError()

    def activate(self, algorithm=None, **kwargs):
        if algorithm is not None:
            self.algorithm = algorithm
        self.loss_fn.register_post_loss_regularizer(self.__class__.__name__, self.regularizer)

    def deactivate(self):
        self.loss_fn.remove_post_loss_regularizer(self.__class__.__name__)


class FIMTraceRegularizer(AdditiveRegularizer):
    """beta * mean_x tr F(x), F the Fisher information of q(theta|x) w.r.t. x.  ``algorithm="ema"`` (the config's
    choice) estimates the trace with ``ema_mc_samples`` scores per row, takes its weight gradient (double backward),
    smooths it with an EMA and OVERWRITES ``p.grad`` with the bias-corrected average; the caller's ``loss.backward()``
    then accumulates the NLL gradient on top."""

    def __init__(self, model: Module, loss_fn: TrainLoss, beta: float = 1.0, algorithm: str = "jac_exact",
                 mc_samples: int = 20, ema_mc_samples: int = 1, ema_decay: float = 0.99, clamp_val=None,
                 grad_clamp_val=None, reduce: str = "mean", cuda_graph: bool = True):
        super().__init__(model, loss_fn, reduce=reduce)
        # ``cuda_graph``: capture (trace estimate + double backward + EMA update) once per input shape and replay it
        # every step -- the unfused path is ~1.7k tiny launches, i.e. launch-bound; replay runs them back to back.
        self.cuda_graph = cuda_graph
        self._graphs = {}
        self._ema_static = None
        self.beta = beta
        self.algorithm = algorithm
        self.mc_samples = mc_samples
        self.ema = ExponentialMovingAverageEstimator(ema_decay)
        self.ema_mc_samples = ema_mc_samples
        self.clamp_val = clamp_val
        self.grad_clamp_val = grad_clamp_val

    def _set_algorithm(self, method, **kwargs) -> None:
        self._algorithm = method
        if method == "jac_exact":
            self._regularizer = self._regularizer_exact
        elif method == "ema":
            self._regularizer = self._regularizer_ema

    def _trace(self, input: Tensor, output: Distribution, mc_samples: int) -> Tensor:
        from .autograd import differentiable_only
        with differentiable_only():   # the score is differentiated again: build the op-by-op graph straight away
            reg = monte_carlo_fisher(self.model, input, output, mc_samples=mc_samples, create_graph=True,
                                     retain_graph=True, typ="trace")
        if self.reduce in ("mean", "sum"):
            # mean over the finite entries without a data-dependent shape (== reg[isfinite(reg)].mean(); capturable)
            mask = torch.isfinite(reg)
            reg = torch.where(mask, reg, torch.zeros_like(reg))
            if self.clamp_val is not None:
                reg = reg.clamp(min=0, max=self.clamp_val)
            return self.beta * (reg.sum() / mask.sum())
        reg = reg[torch.isfinite(reg)]
        if self.clamp_val is not None:
            reg = reg.clamp(min=0, max=self.clamp_val)
        return self.beta * self._reduce(reg)

    def _closed_form_ok(self, input: Tensor) -> bool:
        """The graph-free closed form (rabi_b200/fisher.py::fisher_trace_weight_grad) serves two-hidden-layer conditioners with an
        identity (+ z-score) embedding; ``RABI_FIM_CLOSED_FORM=0`` keeps the op-by-op double backward."""
        import os
        from .models import _split_embedding
        if os.environ.get("RABI_FIM_CLOSED_FORM", "1") == "0" or not input.is_cuda or input.dim() != 2:
            return False
        if self.reduce not in ("mean", "sum") or _split_embedding(self.model.embedding_net)[1] is not None:
            return False
        return self.model.engine().fused_training_ok()

    def _reg_and_grads(self, input: Tensor, output, mc_samples: int, params):
        """(beta * reduce tr F, its gradient w.r.t. ``params``): closed form on the kernels where it serves, else the op-by-op
        double backward."""
        if self._closed_form_ok(input):
            from .fisher import fisher_trace_weight_grad
            # reduce "sum" is the mean as well in the reference (defenses/base.py:109-110, kept by _reduce above)
            return fisher_trace_weight_grad(self.model, input, mc_samples, beta=self.beta, reduce="mean",
                                            clamp_val=self.clamp_val)
        reg = self._trace(input, output if output is not None else self.model(input), mc_samples)
        return reg, torch.autograd.grad(reg, params, retain_graph=True)

    def _regularizer_exact(self, input: Tensor, output: Distribution, target: Tensor) -> Tensor:
        # for a flow the reference's fisher_information_matrix dispatches to the MC score estimator as well
        return self._trace(input, output, self.mc_samples)

    # ---- CUDA-graph replay of the ema step ---------------------------------------------------------------------
    def _graph_eligible(self, input: Tensor) -> bool:
        from . import noise
        return (self.cuda_graph and input.is_cuda and input.dim() == 2 and noise._queue is None
                and self.reduce in ("mean", "sum") and self.ema.t >= 2)  # two eager steps warm everything up

    def _ema_graphed(self, input: Tensor) -> Tensor:
        from . import noise
        params = list(self.model.parameters())
        key = (tuple(input.shape), input.device, tuple(id(p) for p in params))
        D = self.model.output_dim
        # ONE set of static EMA buffers (average m, step count t) shared by every captured graph: a second input shape
        # (e.g. the smaller last batch of an epoch) must continue the same average, not start its own
        shared = self._ema_static
        if shared is None or shared["key"] != (input.device, tuple(id(p) for p in params)):
            shared = self._ema_static = {"key": (input.device, tuple(id(p) for p in params)),
                                         "t": torch.zeros((), device=input.device),
                                         "m": [m.detach().clone() for m in self.ema.state]}
            self._graphs = {}
        if self.ema.state is not shared["m"]:   # an eager step ran in between: take its state over
            for m_s, m_h in zip(shared["m"], self.ema.state):
                m_s.copy_(m_h.detach())
        shared["t"].fill_(float(self.ema.t))
        st = self._graphs.get(key)
        if st is None:
            d = float(self.ema.decay)
            st = {"x": input.detach().clone(),
                  "z": torch.empty((self.ema_mc_samples, input.shape[0], D), device=input.device),
                  "t": shared["t"],
                  "m": shared["m"],
                  "grads": [torch.zeros_like(p) for p in params]}
            side = torch.cuda.Stream(device=input.device)
            side.wait_stream(torch.cuda.current_stream(input.device))
            graph = torch.cuda.CUDAGraph()
            st["z"].normal_()
            with torch.cuda.stream(side):
                with torch.cuda.graph(graph, stream=side):
                    # Differentiate w.r.t. fresh leaves that alias the live parameters: the parameters' own
                    # AccumulateGrad nodes belong to the (still alive) NLL graph on the default stream, and letting
                    # the engine touch them would join the capturing stream with the legacy stream.
                    from torch.nn.utils import stateless
                    leaves = {n: p.detach().requires_grad_() for n, p in self.model.named_parameters()}
                    self.model.invalidate()   # the graph carries its own re-pack of the weight image (see graphs.py)
                    with stateless._reparametrize_module(self.model, leaves):
                        with noise.injected([st["z"]]):
                            _, grads = self._reg_and_grads(st["x"], None, self.ema_mc_samples,
                                                           [leaves[n] for n, _ in self.model.named_parameters()])
                    st["t"].add_(1.0)
                    corr = 1.0 - torch.pow(torch.full_like(st["t"], 1.0 - d), st["t"])
                    for m, g_, out in zip(st["m"], grads, st["grads"]):
                        m.mul_(1.0 - d).add_(torch.nan_to_num(g_), alpha=d)
                        out.copy_(torch.nan_to_num(m / corr))
            torch.cuda.current_stream(input.device).wait_stream(side)
            st["graph"] = graph
            self._graphs[key] = st
        st["x"].copy_(input.detach())
        st["z"].normal_()
        st["graph"].replay()
        self.ema.t += 1
        self.ema.state = shared["m"]
        for p, g_ in zip(params, st["grads"]):
            p.grad = g_
        if self.grad_clamp_val is not None:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.grad_clamp_val, norm_type=2.0)
        return torch.zeros(1, device=input.device)

    def _regularizer_ema(self, input: Tensor, output: Distribution, target: Tensor) -> Tensor:
        if self._graph_eligible(input):
            return self._ema_graphed(input)
        _, grads = self._reg_and_grads(input, output, self.ema_mc_samples, list(self.model.parameters()))
        self.ema(grads)
        for params, grad in zip(self.model.parameters(), self.ema.value):
            params.grad = torch.nan_to_num(grad.detach())
        if self.grad_clamp_val is not None:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.grad_clamp_val, norm_type=2.0)
        return torch.zeros(1, device=input.device)


class TradesAdversarialRegularizer(AdditiveRegularizer):
    """beta * D(q(.|x) || q(.|x')) with x' = attack(x)   (clean distribution FIRST)."""

    def __init__(self, model: Module, loss_fn: TrainLoss, attack: Attack, reg_loss_fn: Callable, beta: float = 1.0,
                 reduce: str = "mean", cuda_graph: bool = True, **kwargs):
        super().__init__(model, loss_fn, reduce)
        self.attack = attack
        self.reg_loss_fn = reg_loss_fn
        self.beta = beta
        self._regularizer = self._regularizer_attack
        # ``cuda_graph``: after two eager calls of a batch shape the regulariser and its weight gradient are replayed as
        # CUDA graphs with the base noise drawn into a static buffer (rabi_b200/graphs.py)
        self.cuda_graph = cuda_graph
        self._graphed = None

    def _set_algorithm(self, method: str):
        self._regularizer = self._regularizer_attack

    def _regularizer_attack(self, input: Tensor, output: Tensor, target: Tensor):
        if self.beta == 0.0:
            return torch.zeros(1, device=input.device)
        # NOTE: the reference's inner ``loss.backward()`` calls also accumulate the attack loss' weight gradient into
        # ``p.grad`` (an unintended side effect of advertorch's loop).  The fused attack has no weight
        # gradient, so that pollution is NOT reproduced; the regulariser value and its own gradient are identical.
        x_pert = self.attack.perturb(input)
        from . import noise
        S = getattr(self.reg_loss_fn, "mc_samples", None)
        if (self.cuda_graph and S is not None and input.is_cuda and input.dim() == 2 and torch.is_grad_enabled()
                and noise._queue is None and not input.requires_grad
                and getattr(self.reg_loss_fn, "reduction", None) == "mean"):
            if self._graphed is None:
                from .graphs import GraphedParamFunction

                def reg(x, xp, z):
                    with noise.injected([z]):
                        return self.reg_loss_fn(self.model(x), self.model(xp)).reshape(1)

                self._graphed = GraphedParamFunction(self.model, reg)
            z = noise.standard_normal((int(S), input.shape[0], self.model.output_dim), input.device)
            if self._graphed.ready((input, x_pert, z)):
                return self.beta * self._graphed(input, x_pert, z)
            with noise.injected([z]):
                return self.beta * self.reg_loss_fn(self.model(input), self.model(x_pert))
        return self.beta * self.reg_loss_fn(self.model(input), self.model(x_pert))


class L2PGDrKLTrades(TradesAdversarialRegularizer):
    def __init__(self, model: Module, loss_fn: TrainLoss, attack_loss_fn: Optional[EvalLoss] = None,
                 reg_loss_fn: Optional[EvalLoss] = None, eps: float = 0.5, beta: float = 1.0, reduce: str = "mean",
                 **kwargs):
        if attack_loss_fn is None:
            attack_loss_fn = ReverseKLLoss(mc_samples=1)
        if reg_loss_fn is None:
            reg_loss_fn = _default_reg_loss(attack_loss_fn)
        attack = L2PGDAttack(model, attack_loss_fn, eps=eps, **kwargs)
        super().__init__(model, loss_fn, attack, reg_loss_fn, beta, reduce, **kwargs)


def _default_reg_loss(attack_loss_fn):
    if hasattr(attack_loss_fn, "mc_samples"):
        return attack_loss_fn.__class__(mc_samples=int(attack_loss_fn.mc_samples))
    return attack_loss_fn.__class__()


class L2PGDTrades(TradesAdversarialRegularizer):
    """TRADES with the forward KL (samples from the clean posterior) as attack and regulariser loss."""

    def __init__(self, model: Module, loss_fn: TrainLoss, attack_loss_fn: Optional[EvalLoss] = None,
                 reg_loss_fn: Optional[EvalLoss] = None, eps: float = 0.5, beta: float = 1.0, reduce: str = "mean",
                 **kwargs):
        if attack_loss_fn is None:
            attack_loss_fn = ForwardKLLoss(mc_samples=1)
        if reg_loss_fn is None:
            reg_loss_fn = _default_reg_loss(attack_loss_fn)
        attack = L2PGDAttack(model, attack_loss_fn, eps=eps, **kwargs)
        super().__init__(model, loss_fn, attack, reg_loss_fn, beta, reduce, **kwargs)


# ---- defenses that change the data BEFORE the loss ------------------------------------------------------------------------
class DataAugmentationRegularizer(Defense):
    """(x, theta) -> (x', theta') through ``TrainLoss.register_pre_loss_regularizer``."""

    @property
    def regularizer(self) -> Callable:
        return self._transform

    def _transform(self, input: Tensor, target):
        raise NotImplementedError("a data-augmentation defense must define how it transforms (input, target)")

    def activate(self):
        self.loss_fn.register_pre_loss_regularizer(self.__class__.__name__, self.regularizer)

    def deactivate(self):
        self.loss_fn.remove_pre_loss_regularizer(self.__class__.__name__)


class AdversarialTraining(DataAugmentationRegularizer):
    """The training loss is evaluated on attack(x) instead of x."""

    def __init__(self, model: Module, loss_fn: TrainLoss, attack: Attack, **kwargs):
        super().__init__(model, loss_fn, **kwargs)
        self.attack = attack

    def _transform(self, input: Tensor, target):
        if not self.attack.targeted:
            x_tilde = self.attack.perturb(input)
        else:
            x_tilde = self.attack.perturb(input, target)
        return x_tilde.detach(), target


class L2PGDAdversarialTraining(AdversarialTraining):
    def __init__(self, model: Module, loss_fn: TrainLoss, attack_loss_fn: Optional[Callable] = None,
                 eps: float = 0.5, **kwargs):
        if attack_loss_fn is None:
            attack_loss_fn = ForwardKLLoss(mc_samples=1)
        super().__init__(model, loss_fn, L2PGDAttack(model, attack_loss_fn, eps=eps, **kwargs))


class L2UniformNoiseTraining(AdversarialTraining):
    def __init__(self, model: Module, loss_fn: TrainLoss, eps: float = 0.5, **kwargs):
        super().__init__(model, loss_fn, L2UniformNoiseAttack(model, eps=eps, **kwargs))
