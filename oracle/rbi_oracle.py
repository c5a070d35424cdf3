"""ORACLE (test infrastructure, not product code).

Standalone CPU restatement (pure PyTorch fp32, eager, op-for-op) of RABI's MAF hot path, on top of
the restated third-party arithmetic in ``oracle/thirdparty``.  It exists so that the parity tests and the
CPU baseline can run on the GPU box, where /root/reference is absent.  Each function cites the reference
lines it follows (paths relative to /root/reference/src).

Pinning: ``oracle/gen_golden.py`` executes the UNMODIFIED reference ``rbi`` package in the build container
(over ``oracle/thirdparty/install_stubs``) with recorded base noise and asserts that this restatement
reproduces it bit-for-bit before writing ``tests/golden/*.pt``; ``tests/test_oracle.py`` re-checks the
restatement against those fixtures.  The Pyro / advertorch layer underneath is restated from the pinned
releases and is itself "parity unpinned" (see the headers in oracle/thirdparty/).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this module.  The product (``rabi_b200``) never does.
"""
from __future__ import annotations

import contextlib
from functools import partial
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
from torch import Tensor
from torch.distributions import Distribution, Normal, TransformedDistribution

from .thirdparty import advertorch_restated as adv
from .thirdparty import pyro_restated as pr


# ----------------------------------------------------------------------------- noise injection
@contextlib.contextmanager
def injected_base_noise(noise: Sequence[Tensor]):
    """Serve ``Normal.rsample``'s standard-normal draws from ``noise`` (in call order).

    torch.distributions.Normal.rsample calls ``torch.distributions.normal._standard_normal``; the parity
    tests replace it so that oracle and CUDA kernels consume the *same* z (CPU and CUDA
    generators give different streams).  Shapes must match the requests exactly.
    """
    import torch.distributions.normal as tn

    queue = list(noise)
    orig = tn._standard_normal

    def fake(shape, dtype, device):
        z = queue.pop(0)
        assert tuple(z.shape) == tuple(shape), (z.shape, shape)
        return z.to(dtype=dtype, device=device)

    orig_sample = Normal.sample

    def fake_sample(self, sample_shape=torch.Size()):
        # ``q.sample`` (no rsample) draws with torch.normal(loc, scale); the flow's base is N(0, 1), so the draw IS z
        if len(torch.Size(sample_shape)) == 0:  # the throw-away draw of every model(x) (utils/distributions.py:457)
            return orig_sample(self, sample_shape)
        shape = self._extended_shape(sample_shape)
        z = queue.pop(0)
        assert tuple(z.shape) == tuple(shape), (z.shape, shape)
        with torch.no_grad():
            return self.loc.expand(shape) + self.scale.expand(shape) * z.to(self.loc)

    tn._standard_normal = fake
    Normal.sample = fake_sample
    try:
        yield queue
    finally:
        tn._standard_normal = orig
        Normal.sample = orig_sample


@contextlib.contextmanager
def recorded_base_noise(record: List[Tensor]):
    """Record every standard-normal draw made by ``Normal.rsample`` (used by gen_golden.py)."""
    import torch.distributions.normal as tn

    orig = tn._standard_normal

    def rec(shape, dtype, device):
        z = orig(shape, dtype, device)
        record.append(z.detach().clone())
        return z

    orig_sample = Normal.sample

    def rec_sample(self, sample_shape=torch.Size()):
        out = orig_sample(self, sample_shape)
        if len(torch.Size(sample_shape)) == 0:
            return out
        assert bool((self.loc == 0).all()) and bool((self.scale == 1).all()), "only a standard-normal base is recorded"
        record.append(out.detach().clone())
        return out

    tn._standard_normal = rec
    Normal.sample = rec_sample
    try:
        yield record
    finally:
        tn._standard_normal = orig
        Normal.sample = orig_sample


# ----------------------------------------------------------------------------- model
class ZScoreLayer(nn.Module):
    """rbi/rbi/models/module.py:109-122."""

    def __init__(self, mean, std):
        super().__init__()
        self.register_buffer("mean", mean)
        self.register_buffer("std", std)

    def forward(self, x):
        return (x - self.mean) / self.std


class _CondInvAffineAR(nn.Module):
    """rbi/rbi/models/transforms.py:24-46 -- ``condition(h)`` -> ``AffineAutoregressive(...).inv``."""

    def __init__(self, arn, **kwargs):
        super().__init__()
        self.nn = arn
        self.kwargs = kwargs

    def condition(self, context):
        cond_nn = partial(self.nn, context=context)
        cond_nn.permutation = cond_nn.func.permutation
        cond_nn.get_permutation = cond_nn.func.get_permutation
        return pr.AffineAutoregressive(cond_nn, **self.kwargs).inv


class OracleMAF(nn.Module):
    """``InverseAffineAutoregressiveModel`` restated (rbi/rbi/models/flows.py:72-113 over
    rbi/rbi/models/base.py:274-446 and rbi/rbi/utils/distributions.py:448-492).

    Same constructor arguments, same module tree (``t{i}.nn.layers.{j}.{weight,bias,mask}``,
    ``t{i}.nn.permutation``, ``base_dist_mean/scale``) and therefore the same RNG consumption at
    construction and an interchangeable ``state_dict``.
    """

    def __init__(
        self,
        input_dim: int,
        output_dim: int,
        hidden_dims: List[int] = [50, 50],
        num_transforms: int = 5,
        shuffle: bool = True,
        with_cache: bool = False,
        output_transform=None,
        embedding_net: Optional[nn.Module] = None,
        **kwargs,
    ):
        super().__init__()
        self.embedding_net = embedding_net if embedding_net is not None else nn.Identity()
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.with_cache = with_cache
        ctx_dim = self.embedding_net(torch.zeros((1, input_dim))).shape[-1]  # base.py:314
        self.context_dim = ctx_dim
        self.flow_transforms = []  # conditional transforms (and Permute layers if shuffle)
        for i in range(num_transforms):
            arn = pr.ConditionalAutoRegressiveNN(output_dim, ctx_dim, hidden_dims)  # transforms.py:92-98
            t = _CondInvAffineAR(arn, **kwargs)
            self.add_module("t" + str(i), t)
            self.flow_transforms.append(t)
            if shuffle:  # base.py:326-333
                pl = pr.permute(output_dim)
                self.register_buffer("permute" + str(i), pl.permutation)
                self.register_buffer("inv_permute" + str(i), pl.inv_permutation)
                self.flow_transforms.append(pl)
        self.register_buffer("base_dist_mean", torch.zeros(output_dim))  # base.py:345-353
        self.register_buffer("base_dist_scale", torch.ones(output_dim))
        self.output_transform = output_transform

    def forward(self, x: Tensor) -> Distribution:
        # base.py:417-446
        expand = False
        if x.shape == torch.Size([]):
            x = x.reshape(1, 1)
        if x.shape.numel() > 1 and x.shape[0] == 1:
            x = x.squeeze(0)
            expand = True
        h = self.embedding_net(x)
        base = Normal(self.base_dist_mean, self.base_dist_scale, validate_args=False)
        # utils/distributions.py:455-461
        if len(h.shape) > 1 and h.shape[0] > 1:
            base = base.expand((*h.shape[:-1], *base.sample().shape))
        ts = []
        for t in self.flow_transforms:
            ts.append(t.condition(h) if isinstance(t, _CondInvAffineAR) else t)
        if self.output_transform is not None:
            ts.append(self.output_transform)
        q = TransformedDistribution(base, ts)
        if expand:
            q = q.expand((1,) + q.batch_shape)
        # base.py:442-444 -- sets the flag of the *wrapper* (_InverseTransform) / Permute / output transform
        # only; the inner AffineAutoregressive keeps cache_size=1
        for t in q.transforms:
            if not isinstance(t, pr.IndependentTransform):
                t._cache_size = int(self.with_cache)
        return q


# ----------------------------------------------------------------------------- divergences / losses
def mc_kl(p: Distribution, q: Distribution, mc_samples: int) -> Tensor:
    """KL(p||q) ~= mean_S[log p(th_s) - log q(th_s)], th_s = p.rsample((S,)).  rbi/rbi/loss/kl_div.py:38-56."""
    batch_shape = torch.broadcast_shapes(p.batch_shape, q.batch_shape)
    p = p.expand(batch_shape)
    q = q.expand(batch_shape)
    samples = p.rsample((mc_samples,))
    logq = q.log_prob(samples)
    logp = p.log_prob(samples)
    return torch.mean(logp - logq, 0)


class ReverseKLLoss:
    """rbi/rbi/loss/eval_loss.py:129-150 + loss/base.py:13-70: KL(output || target), samples from ``output``."""

    def __init__(self, mc_samples: int = 10, reduction: Optional[str] = "mean"):
        self.mc_samples = mc_samples
        self.reduction = reduction

    def __call__(self, output, target):
        kl = mc_kl(output, target, self.mc_samples)
        return kl.mean() if self.reduction == "mean" else kl


class ForwardKLLoss:
    """rbi/rbi/loss/eval_loss.py:108-126: KL(target || output), samples from ``target``; then ``reduce``."""

    def __init__(self, mc_samples: int = 10, reduction: Optional[str] = "mean"):
        self.mc_samples = mc_samples
        self.reduction = reduction

    def __call__(self, output, target):
        kl = mc_kl(target, output, self.mc_samples)
        return kl.mean() if self.reduction == "mean" else kl


def nll(model, x, theta):
    """rbi/rbi/loss/train_loss.py:16-29 (unreduced)."""
    q = model(x)
    return -q.log_prob(theta.reshape(*(q.batch_shape + q.event_shape)))


# ----------------------------------------------------------------------------- L2-PGD attack
def l2pgd_perturb(
    model,
    x: Tensor,
    loss_fn,
    eps: float,
    nb_iter: int,
    eps_iter: float,
    clip_min: float = -1000.0,
    clip_max: float = 1000.0,
    delta0: Optional[Tensor] = None,
    target=None,
) -> Tensor:
    """rbi ``L2PGDAttack.perturb`` (untargeted).

    attacks/base.py:35-71 (y = predict(x) under no_grad), then advertorch's PGDAttack.perturb /
    perturb_iterative(ord=2) -- see oracle/thirdparty/advertorch_restated.py for why it is advertorch's
    own perturb (rand_init_delta) and not ``pgd_randint_change_perturb``.  ``delta0`` injects the initial
    perturbation (already projected/clamped) instead of drawing it.
    """
    x = x.detach().clone()
    if target is None:
        with torch.no_grad():
            target = model(x)
    delta = nn.Parameter(torch.zeros_like(x))
    if delta0 is None:
        adv.rand_init_delta(delta, x, 2, eps, clip_min, clip_max)
        delta.data = adv.clamp(x + delta.data, min=clip_min, max=clip_max) - x
    else:
        delta.data = delta0.detach().clone()
    out = adv.perturb_iterative(
        x, target, model, nb_iter=nb_iter, eps=eps, eps_iter=eps_iter, loss_fn=loss_fn,
        minimize=False, ord=2, clip_min=clip_min, clip_max=clip_max, delta_init=delta, l1_sparsity=None)
    return out.data


def rkl_rob_metric(model, x, x_tilde, mc_samples: int) -> Tensor:
    """rbibm/rbibm/metric/base.py:261-271 + robustness_metric.py:84-99: KL(q(.|x~) || q(.|x)) per row."""
    with torch.no_grad():
        q_tilde = model(x_tilde)
        q = model(x)
        loss = ReverseKLLoss(mc_samples=mc_samples, reduction=None)(q_tilde, q)
        mask_neg = loss < -10000
        if mask_neg.any():
            loss[mask_neg] = loss[~mask_neg].median()
    return loss


# ----------------------------------------------------------------------------- Fisher trace
def score_function(theta_s: Tensor, x: Tensor, model, create_graph=True) -> Tensor:
    """rbi/rbi/utils/fisher_info.py:246-280: d/dX sum log q(theta_s | X), X = x repeated over S."""
    S = theta_s.shape[0]
    X = x.clone().repeat(S, *(1,) * len(theta_s.shape[1:])).requires_grad_()
    y = model(X).log_prob(theta_s)
    return torch.autograd.grad(y.sum(), X, create_graph=create_graph, retain_graph=True)[0]


def mc_fisher_trace(model, x: Tensor, q: Distribution, mc_samples: int, create_graph=True) -> Tensor:
    """rbi/rbi/utils/fisher_info.py:185-226 (method=score_based, typ=trace); samples are NOT detached (Q4)."""
    theta_s = q.rsample((mc_samples,))
    score = score_function(theta_s, x, model, create_graph=create_graph)
    return torch.mean(score ** 2, 0).sum(-1)


class EMA:
    """rbi/rbi/utils/streaming_estimators.py:33-74 -- note ``decay`` weights the NEW value."""

    def __init__(self, decay=0.95):
        self.t = 0
        self.decay = decay
        self.m = None

    def __call__(self, xs):
        self.t += 1
        xs = [torch.nan_to_num(x) for x in xs]
        if self.m is None:
            self.m = [self.decay * x for x in xs]
        else:
            self.m = [self.decay * x + (1 - self.decay) * m for x, m in zip(xs, self.m)]

    @property
    def value(self):
        return [m / (1 - (1 - self.decay) ** self.t) for m in self.m]


def fim_trace_regularizer_ema(model, x, q, beta, ema: EMA, ema_mc_samples, clamp_val=None, grad_clamp_val=None):
    """rbi/rbi/defenses/fisher_regularization.py:67-95: writes the EMA'd grad of beta*mean tr F into p.grad."""
    reg = mc_fisher_trace(model, x, q, ema_mc_samples, create_graph=True)
    reg = reg[torch.isfinite(reg)]
    if clamp_val is not None:
        reg = reg.clamp(min=0, max=clamp_val)
    reg = beta * torch.mean(reg)
    grads = torch.autograd.grad(reg, list(model.parameters()), retain_graph=True)
    ema(grads)
    for p, g in zip(model.parameters(), ema.value):
        p.grad = torch.nan_to_num(g.detach())
    if grad_clamp_val is not None:
        torch.nn.utils.clip_grad_norm_(model.parameters(), grad_clamp_val, norm_type=2.0)
    return reg.detach()


def train_loss_with_regularizer(model, x, theta, regularizer=None):
    """rbi/rbi/loss/base.py:142-175 in training mode with one post-loss regulariser."""
    q = model(x)
    loss = -q.log_prob(theta.reshape(*(q.batch_shape + q.event_shape)))
    post = torch.zeros(1)
    if regularizer is not None:
        post = post + regularizer(x, q, theta)
    return torch.mean(loss) + (post - post.detach())


def trades_rkl_regularizer(model, x, beta, eps, eps_iter, nb_iter, mc_samples=1, delta0=None,
                           clip_min=-1000.0, clip_max=1000.0):
    """rbi/rbi/defenses/trades_regularizers.py:58-64,101-133: beta * mean_B KL(q(.|x) || q(.|x_pert))."""
    x_pert = l2pgd_perturb(model, x, ReverseKLLoss(mc_samples=mc_samples), eps, nb_iter, eps_iter,
                           clip_min, clip_max, delta0=delta0)
    return beta * ReverseKLLoss(mc_samples=mc_samples)(model(x), model(x_pert)), x_pert


# ============================================================================= SURVEY 8f "next" rows
def rob_metric_eval(model, x, loss_fn, perturb_fn, attack_attemps=5, descending=True, targeted=False):
    """rbibm/rbibm/metric/base.py:225-300 (``RobustnessMetric._eval``) for the untargeted KL metrics of
    robustness_metric.py:54-155: per attempt x~ = perturb_fn(i); loss = loss_fn(q(.|x~), q(.|x)) (reduction None);
    overflow guard; keep per row the best attempt.  ``loss_fn`` is a Reverse/ForwardKLLoss with reduction None."""
    losses = xs_tilde = None
    for i in range(attack_attemps):
        x_tilde = perturb_fn(i)
        with torch.no_grad():
            q_tilde = model(x_tilde)
            q = model(x)
            loss = loss_fn(q_tilde, q)
            mask_neg = loss < -10000
            loss[mask_neg] = loss[~mask_neg].median()
        if i == 0:
            xs_tilde, losses = x_tilde.detach(), loss.detach()
        else:
            if descending:
                better = loss > losses if not targeted else loss < losses
            else:
                better = loss < losses if not targeted else loss > losses
            xs_tilde[better] = x_tilde[better].detach()
            losses[better] = loss[better].detach()
    return losses, xs_tilde


def nll_metric(model, x, theta):
    """rbibm/rbibm/metric/approximation_metric.py:205-232: -log q(theta|x) per row under no_grad."""
    with torch.no_grad():
        return nll(model, x, theta)


def expected_coverage(model, x, theta, mc_samples, alphas):
    """rbibm/rbibm/metric/approximation_metric.py:300-324: one [alphas] vector for one batch."""
    q = model(x)
    logq_theta = q.log_prob(theta)
    samples = q.sample((mc_samples,))
    logq_samples = q.log_prob(samples)
    out = []
    for a in alphas:
        if float(a) == 0:
            out.append(torch.zeros(1).mean())
        elif float(a) == 1:
            out.append(torch.ones(1).mean())
        else:
            cut_off = int(mc_samples * (1 - a))
            srt, _ = torch.sort(logq_samples, dim=0)
            out.append((srt[cut_off:].min(0).values < logq_theta).float().mean())
    return torch.stack(out)


def coverage_abs_auc(cov, alphas):
    """approximation_metric.py:281-291 (the branch every reduction takes): area of |coverage - alpha|."""
    while cov.ndim > 1:
        cov = cov.mean(0)
    return torch.trapz(torch.abs(cov - alphas), alphas)


def sample_l2_uniformly(N, d, eps):
    """rbi/rbi/utils/distributions.py:22-40 with p = 2 (draw order: Gamma, sign uniforms, radius uniforms)."""
    c = 2.0 * torch.ones(d)  # a tensor exponent on purpose: pow(t, tensor) and pow(t, 0.5) (= sqrt) differ by an ulp
    xi = torch.distributions.Gamma(1 / 2.0 * torch.ones(d), 1.0).sample((N,)) ** (1 / c)
    signs = (torch.rand((N, d)) - 0.5).sign()
    v = xi * signs
    z = torch.rand(N, 1) ** (1 / d)
    return eps * z * (v / torch.linalg.norm(v, keepdim=True, ord=2.0, dim=-1))


def l2_uniform_noise_perturb(x, eps, clip_min=-1000.0, clip_max=1000.0, noise=None):
    """rbi/rbi/attacks/noise_attacks.py:30-35."""
    if noise is None:
        noise = sample_l2_uniformly(x.shape[0], x.shape[-1], eps)
    return torch.clip(x + noise, min=clip_min, max=clip_max)


def train_loss_with_pre_regularizer(model, x, theta, transform):
    """rbi/rbi/loss/base.py:156-163 with ONE pre-loss regulariser (defenses/adversarial_training.py:34-50): the NLL is
    evaluated at (x', theta') = transform(x, theta)."""
    x2, theta2 = transform(x, theta)
    q = model(x2)
    loss = -q.log_prob(theta2.reshape(*(q.batch_shape + q.event_shape)))
    post = torch.zeros(1)
    return torch.mean(loss) + (post - post.detach()), x2


def trades_regularizer(model, x, beta, eps, eps_iter, nb_iter, loss_cls, mc_samples=1, delta0=None,
                       clip_min=-1000.0, clip_max=1000.0):
    """rbi/rbi/defenses/trades_regularizers.py:58-99 for a KL loss class (ForwardKLLoss: ``L2PGDTrades``)."""
    x_pert = l2pgd_perturb(model, x, loss_cls(mc_samples=mc_samples), eps, nb_iter, eps_iter, clip_min, clip_max,
                           delta0=delta0)
    return beta * loss_cls(mc_samples=mc_samples)(model(x), model(x_pert)), x_pert
