"""Fisher information of the amortized posterior q(theta | x) with respect to the observation x -- host side of the
Fisher-trace kernels.

What the reference computes (rbi/rbi/utils/fisher_info.py:185-280, score-based Monte-Carlo estimator): draw S
reparameterised samples theta_s ~ q(.|x), take the score  s_{s,b} = d log q(theta_s | x_b) / d x_b  and average its outer
product; FIMTraceRegularizer (rbi/rbi/defenses/fisher_regularization.py:67-95) uses the trace  mean_s ||s_{s,b}||^2.

Two execution paths:

* ``fisher_trace(model, x, mc_samples)`` -- value only, entirely on the kernels (``rabi_fisher_trace_fwd``): sampling chain,
  per-pair density passes with their reverse on the tensor cores, back-projection of every (sample, row) pair, squared norms.
  No autograd graph; used for monitoring / evaluation and by ``FIMTraceRegularizer`` when no gradient is requested.
* ``fisher_scores`` / ``monte_carlo_fisher`` with ``create_graph=True`` -- the estimate as a differentiable function of the
  weights (the regulariser's gradient needs the derivative of the score, including the path through the reparameterised
  samples); built on the op-by-op graph of rabi_b200/autograd.py.

``StreamingGradientAverage`` is the bias-corrected exponential average the regulariser keeps of its weight gradient
(rbi/rbi/utils/streaming_estimators.py:33-74: the decay weights the NEW value).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Union

import torch
from torch import Tensor
from torch.distributions import Distribution

from . import noise

_REDUCERS = {
    "full": lambda s: torch.einsum("sbi,sbj->bij", s, s) / s.shape[0],
    "diag": lambda s: s.square().mean(0),
    "trace": lambda s: s.square().mean(0).sum(-1),
}


def fisher_trace(model, x: Tensor, mc_samples: int = 20, return_scores: bool = False):
    """tr F(x_b) for every row of x [B, dx], on the kernels (no gradient).  -> trace [B] (and the scores [S, B, dx])."""
    from .models import MAFPosterior

    with torch.no_grad():
        from .models import _split_embedding
        if _split_embedding(model.embedding_net)[1] is not None:
            raise NotImplementedError("fisher_trace runs on the flow kernels: identity (+ z-score) embedding only")
        q = model(x)
        if not isinstance(q, MAFPosterior):
            raise NotImplementedError("fisher_trace needs the MAF posterior of rabi_b200.models")
        eng = model.engine()
        if not eng.fused_training_ok():
            raise NotImplementedError("fisher_trace serves the two-hidden-layer conditioners of the tensor-core path")
        Bf, D = q.ctx_in.shape[0], model.output_dim
        z = noise.standard_normal((int(mc_samples), Bf, D), q.ctx_in.device)
        rows = q.ctx_in if q.zscore is None else (q.ctx_in - q.zscore[0]) / q.zscore[1]
        trace, score, _ = eng.fisher_trace(rows, q.A, z, q.zscore[1] if q.zscore is not None else None)
    trace = trace.reshape(q.batch_shape)
    return (trace, score.reshape((int(mc_samples),) + tuple(q.batch_shape) + (-1,))) if return_scores else trace


def fisher_scores(generator: Callable, x: Tensor, samples: Tensor, create_graph: bool = True) -> Tensor:
    """Scores of ``samples`` [S, B, D] under the posteriors of the rows x [B, dx]: d log q(theta_s | x_b) / d x_b, [S, B, dx].
    The rows are replicated along the sample axis into a fresh leaf, so the result is the per-pair derivative (not summed over
    the samples of a row); with ``create_graph`` it stays differentiable w.r.t. the weights and the samples."""
    if x.dim() != 2:
        raise ValueError("x must be [batch, dx]")
    S = samples.shape[0]
    leaf = x.detach().unsqueeze(0).expand(S, *x.shape).reshape(S * x.shape[0], x.shape[1]).clone().requires_grad_()
    logq = generator(leaf).log_prob(samples.reshape(S * x.shape[0], *samples.shape[2:]))
    (g,) = torch.autograd.grad(logq.sum(), leaf, create_graph=create_graph, retain_graph=True)
    return g.reshape(S, *x.shape)


def score_function(x: Tensor, parameter: Tensor, generator: Callable, create_graph: bool = True,
                   retain_graph: bool = True) -> Tensor:
    """Reference-compatible signature (fisher_info.py:246): samples [S, B, D] first, conditioning rows second; [S, B, dx]."""
    return fisher_scores(generator, parameter, x, create_graph=create_graph)


def monte_carlo_fisher(generator: Callable, parameters: Tensor, output: Optional[Distribution] = None,
                       mc_samples: int = 20, create_graph: bool = True, typ: str = "full",
                       method: str = "score_based", **_ignored) -> Tensor:
    """Score-based MC estimate of the Fisher information of ``generator(parameters)`` w.r.t. ``parameters`` [B, dx]:
    ``typ`` "full" -> [B, dx, dx], "diag" -> [B, dx], "trace" -> [B].  The samples keep their graph (reparameterisation), so
    a weight gradient of the estimate includes the path through them."""
    if method != "score_based":
        raise NotImplementedError("only the score-based estimator is on the fisher_trace path")
    if typ not in _REDUCERS:
        raise NotImplementedError(f"unknown typ {typ!r} (full | diag | trace)")
    if not create_graph and typ == "trace" and output is None and hasattr(generator, "engine"):
        try:
            return fisher_trace(generator, parameters, mc_samples)
        except NotImplementedError:
            pass
    q = generator(parameters) if output is None else output
    samples = q.rsample((mc_samples,)) if q.has_rsample else q.sample((mc_samples,))
    return _REDUCERS[typ](fisher_scores(generator, parameters, samples, create_graph=create_graph))


class StreamingGradientAverage:
    """Exponential average of a tensor or a list of tensors, m <- decay * new + (1 - decay) * m, read out with the bias
    correction 1 / (1 - (1 - decay)^t).  Non-finite entries of a new value count as zero."""

    def __init__(self, decay: float = 0.95) -> None:
        self.decay = float(decay)
        self.t = 0
        self.state: Union[None, Tensor, List[Tensor]] = None

    def __call__(self, new: Union[Tensor, Sequence[Tensor]]) -> None:
        single = isinstance(new, Tensor)
        if not single and not isinstance(new, (list, tuple)):
            raise TypeError(f"cannot average a {type(new).__name__}")
        vals = [torch.nan_to_num(v) for v in ([new] if single else new)]
        if self.state is None:
            upd = [self.decay * v for v in vals]
        else:
            old = [self.state] if single else self.state
            upd = [torch.lerp(m, v, self.decay) for m, v in zip(old, vals)]
        self.t += 1
        self.state = upd[0] if single else upd

    @property
    def value(self):
        if self.state is None:
            raise RuntimeError("no value has been averaged yet")
        corr = 1.0 - (1.0 - self.decay) ** self.t
        return self.state / corr if isinstance(self.state, Tensor) else [m / corr for m in self.state]


# name used by the reference's regulariser code
ExponentialMovingAverageEstimator = StreamingGradientAverage
