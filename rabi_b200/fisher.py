"""Monte-Carlo Fisher information of the amortized posterior w.r.t. the observation, mirroring
rbi/rbi/utils/fisher_info.py:185-280 and rbi/rbi/utils/streaming_estimators.py:33-74."""
from __future__ import annotations

from typing import Callable, Optional

import torch
from torch import Tensor
from torch.distributions import Distribution


def score_function(x: Tensor, parameter: Tensor, generator: Callable, create_graph: bool = True,
                   retain_graph: bool = True) -> Tensor:
    """d/dX sum log p(x | X) with X = ``parameter`` repeated along a new leading sample axis (fisher_info.py:246-280)."""
    assert len(parameter.shape) == 2, "Should be of dimension [param_batch, parameter_dim]"
    batch_size = x.shape[0]
    parameters = parameter.clone().repeat(batch_size, *(1,) * len(x.shape[1:])).requires_grad_()
    p = generator(parameters)
    y = p.log_prob(x)
    return torch.autograd.grad(y.sum(), parameters, create_graph=create_graph, retain_graph=retain_graph)[0]


def monte_carlo_fisher(generator: Callable, parameters: Tensor, output: Optional[Distribution] = None,
                       mc_samples: int = 20, create_graph: bool = True, typ: str = "full",
                       method: str = "score_based", **kwargs) -> Tensor:
    """MC estimate of the FIM from scores of reparameterised samples (NOT detached: the weight gradient of the
    estimate includes the path through the samples)."""
    if method != "score_based":
        raise NotImplementedError("only method='score_based' is on the fisher_trace path")
    if output is None:
        output = generator(parameters)
    xs = output.rsample((mc_samples,)) if output.has_rsample else output.sample((mc_samples,))
    score = score_function(xs, parameters, generator, create_graph=create_graph)
    if typ == "full":
        return torch.einsum("mbj, mbi -> bji", score, score) / mc_samples
    if typ == "diag":
        return torch.mean(score ** 2, 0)
    if typ == "trace":
        return torch.mean(score ** 2, 0).sum(-1)
    raise NotImplementedError("Unknown typ.")


class ExponentialMovingAverageEstimator:
    """``decay`` weights the NEW value; bias-corrected read-out (streaming_estimators.py:33-74)."""

    def __init__(self, decay: float = 0.95) -> None:
        self.t = 0
        self.decay = decay
        self._x_old = None

    def __call__(self, x_new):
        self.t += 1
        if isinstance(x_new, Tensor):
            x_new = torch.nan_to_num(x_new)
            self._x_old = self.decay * x_new if self._x_old is None else self.decay * x_new + (1 - self.decay) * self._x_old
        elif isinstance(x_new, (list, tuple)):
            if self._x_old is None:
                self._x_old = [self.decay * torch.nan_to_num(x_i) for x_i in x_new]
            else:
                self._x_old = [self.decay * torch.nan_to_num(x_i) + (1 - self.decay) * c_i
                               for x_i, c_i in zip(x_new, self._x_old)]
        else:
            raise ValueError("Unknown value")

    @property
    def value(self):
        if isinstance(self._x_old, Tensor):
            return self._x_old / (1 - (1 - self.decay) ** self.t)
        if isinstance(self._x_old, (list, tuple)):
            return [val / (1 - (1 - self.decay) ** self.t) for val in self._x_old]
        raise ValueError("Unknonw value")
