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


def fisher_trace_weight_grad(model, x: Tensor, mc_samples: int, beta: float = 1.0, reduce: str = "mean",
                             clamp_val: Optional[float] = None, dense_products: Optional[bool] = None):
    """Value and WEIGHT GRADIENT of the regulariser  R = beta * reduce_b tr F(x_b)  (tr F estimated from ``mc_samples`` scores
    per row) without an autograd graph -- what ``FIMTraceRegularizer`` needs every training step
    (rbi/rbi/defenses/fisher_regularization.py:67-95: ``autograd.grad(reg, parameters)`` of the estimator with
    ``create_graph=True``, including the path through the reparameterised samples).

    Closed form (derivation and fp64 check against autograd's double backward: oracle/fim_closed_form_check.py).  With the
    scores g_p held constant, dR = sum_p 2 w_p g_p . dg_p = sum_p d phi_p, phi_p = xdot_p . grad_x log q(theta_p | x) = the
    derivative of log q along xdot_p = 2 w_p g_p.  So:
      1. kernels (``rabi_fisher_trace_fwd``): theta_p, the primal density passes with their reverse, the scores g_p; the per-pair
         layer activations h0, h1 and inputs y come back through ``rabi_train_emissions``;
      2. tangent (forward-mode) density pass along xdot: for ReLU conditioners the primal pass with the stored masks in place of
         the sign tests, zero biases;
      3. reverse over the (primal, tangent) program seeded on the tangent of log q: two masked reverse passes per transform
         (adjoints of the tangent and of the primal variables);
      4. weight products  ot (x) h1dot + ob (x) h1,  a1t (x) h0dot + a1b (x) h0,  a0t (x) [ctxdot, ydot] + a0b (x) [ctx, y];
      5. the path through the samples: ``rabi_maf_rsample_bwd`` seeded with d phi / d theta.
    Steps 2-5 run as ``rabi_fisher_trace_bwd``: one kernel for the tangent pass and the dual reverse (``k_fim_dual``, 64 pairs per
    CTA), one launch for the 24 stacked weight products (``k_wgrad``) and the sampling-chain reverse -- about 10 launches per
    step instead of the ~1700 of the op-by-op double backward.  ``dense_products=True`` (or ``RABI_FIM_DENSE=1``) runs steps 2-4 as
    [units x pairs] products on ``rabi_sgemm`` with elementwise glue instead: the same arithmetic written out, kept as the
    cross-check of the kernel.  -> (R (0-dim tensor), [grad per parameter, in ``model.parameters()`` order])."""
    import os
    if dense_products is None:
        dense_products = os.environ.get("RABI_FIM_DENSE", "0") == "1"
    from .autograd import _sgemm
    from .models import MAFPosterior, _split_embedding

    if _split_embedding(model.embedding_net)[1] is not None:
        raise NotImplementedError("closed-form Fisher gradient: identity (+ z-score) embedding only")
    if reduce not in ("mean", "sum"):
        raise NotImplementedError("closed-form Fisher gradient: reduce = mean | sum")
    with torch.no_grad():
        q = model(x)
        eng = model.engine()
        if not isinstance(q, MAFPosterior) or not eng.fused_training_ok():
            raise NotImplementedError("closed-form Fisher gradient serves the two-hidden-layer conditioners of the tensor-core path")
        S, D, K = int(mc_samples), model.output_dim, model.num_transforms
        B, C = q.ctx_in.shape
        P = S * B
        z = noise.standard_normal((S, B, D), q.ctx_in.device)
        std = q.zscore[1] if q.zscore is not None else None
        rows = q.ctx_in if q.zscore is None else (q.ctx_in - q.zscore[0]) / q.zscore[1]
        A = q.A
        # ---- 1. kernels: samples, primal passes, scores
        trace, score, theta = eng.fisher_trace(rows, A, z, std)
        finite = torch.isfinite(trace)
        tr = torch.where(finite, trace, torch.zeros_like(trace))
        live = finite
        if clamp_val is not None:
            live = live & (tr > 0) & (tr < clamp_val)     # clamp(min=0, max=clamp_val) passes no gradient outside
            tr = tr.clamp(min=0, max=clamp_val)
        n = finite.sum().clamp_min(1) if reduce == "mean" else torch.ones((), device=trace.device)
        reg = beta * tr.sum() / n
        w_row = torch.where(live, beta / (n * S), torch.zeros_like(trace))               # d R / d ||g_{s,b}||^2
        if not dense_products:
            mw, bo_ = [], []
            for k in range(K):
                net = getattr(model, "t" + str(k)).nn
                mw.append([(l.weight.detach() * l.mask).contiguous() for l in net.layers])
                bo_.append(net.layers[-1].bias.detach())
            grads = eng.fisher_trace_bwd(rows, A, z, std, score, w_row, mw, bo_)
            return reg, [t for gk in grads for wb in gk for t in wb]
        em = eng.train_emissions(S, B)
        g = torch.nan_to_num(score.reshape(P, C))
        xdot = 2.0 * w_row.repeat(S).unsqueeze(1) * g                                     # [P, C]
        cdot = (xdot / std if std is not None else xdot).t().contiguous()                 # [C, P] tangent of the context rows
        ctx_p = rows.t().contiguous().repeat(1, S)                                        # [C, P]
        params = model.transforms_params()
        mm = lambda a, b: _sgemm(a, b, False, False)          # [M, K] @ [K, P]
        mmt = lambda a, b: _sgemm(a, b, True, False)          # [K, M]^T @ [K, P]
        outer = lambda a, b: _sgemm(a, b, False, True)        # [M, P] @ [N, P]^T  (reduction over the pairs)
        lo, hi = model.log_scale_min_clip, model.log_scale_max_clip
        st = []
        for k in range(K):
            net = getattr(model, "t" + str(k)).nn
            (W0, b0), (W1, b1), (Wo, bo) = [(l.weight.detach() * l.mask, l.bias.detach()) for l in net.layers]
            h0, h1, y = em["h0"][k], em["h1"][k], em["y"][k]
            o = mm(Wo, h1) + bo.unsqueeze(1)
            s_ = o[D:]                                       # clamp_preserve_gradients: identity for every derivative
            es = torch.exp(s_.clamp(lo, hi))
            st.append(dict(W0c=W0[:, :C].contiguous(), W0t=W0[:, C:].contiguous(), W1=W1, Wo=Wo, h0=h0, h1=h1, y=y, es=es,
                           m0=(h0 > 0).float(), m1=(h1 > 0).float(), u=es * y + o[:D]))
        # ---- 2. tangent density pass (transform K-1 first: theta itself has no tangent)
        ydot = torch.zeros((D, P), device=g.device)
        for k in range(K - 1, -1, -1):
            d = st[k]
            a0d = mm(d["W0c"], cdot) + mm(d["W0t"], ydot)
            d["h0d"] = a0d * d["m0"]
            d["h1d"] = mm(d["W1"], d["h0d"]) * d["m1"]
            od = mm(d["Wo"], d["h1d"])
            d["sd"], d["ydot"] = od[D:], ydot
            ydot = d["es"] * (d["sd"] * d["y"] + ydot) + od[:D]
        # ---- 3. + 4. reverse over the dual program, seed 1 on the tangent of log q = sum sd - u0 . u0dot
        ut, ub = -st[0]["u"], -ydot
        grads = []
        for k in range(K):
            d = st[k]
            es, y, sd, yd = d["es"], d["y"], d["sd"], d["ydot"]
            ot = torch.cat([ut, 1.0 + ut * es * y], 0)
            ob = torch.cat([ub, ub * es * y + ut * es * (sd * y + yd)], 0)
            a1t = mmt(d["Wo"], ot) * d["m1"]
            a1b = mmt(d["Wo"], ob) * d["m1"]
            a0t = mmt(d["W1"], a1t) * d["m0"]
            a0b = mmt(d["W1"], a1b) * d["m0"]
            gWo = outer(ot, d["h1d"]) + outer(ob, d["h1"])
            gW1 = outer(a1t, d["h0d"]) + outer(a1b, d["h0"])
            gW0 = torch.cat([outer(a0t, cdot) + outer(a0b, ctx_p), outer(a0t, yd) + outer(a0b, y)], 1)
            grads.append([(gW0, a0b.sum(1)), (gW1, a1b.sum(1)), (gWo, ob.sum(1))])
            yt = ut * es + mmt(d["W0t"], a0t)
            yb = ub * es + ut * es * sd + mmt(d["W0t"], a0b)
            ut, ub = yt, yb
        # ---- 5. the path through the reparameterised samples: d phi / d theta = ub (adjoint of the last transform's input)
        theta_bar = ub.t().contiguous().reshape(S, B, D)
        _, samp = eng.rsample_bwd(rows, A, z, theta_bar, torch.zeros((S, B), device=g.device), want_weights=True)
        out = []
        for k in range(K):
            net = getattr(model, "t" + str(k)).nn
            for l, layer in enumerate(net.layers):
                gw, gb = grads[k][l]
                out.append(gw * layer.mask + samp[k][l][0])
                out.append(gb + samp[k][l][1])
    return reg, out


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
