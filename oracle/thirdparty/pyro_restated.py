"""ORACLE (test infrastructure, not product code).

CPU restatement of the un-vendored third-party arithmetic RABI's MAF path runs on:
``pyro_ppl==1.8.4`` (pinned at /root/reference/src/requirements.txt:6).  Pyro is
NOT installed in the build image and is not vendored under /root/reference, so the
published algorithm is restated here from the pinned release
(pyro/nn/auto_reg_nn.py, pyro/distributions/transforms/affine_autoregressive.py,
pyro/distributions/transforms/utils.py, pyro/distributions/conditional.py,
pyro/distributions/transforms/permute.py) -- "[3P-memory]" in SURVEY.md 8c.
PARITY UNPINNED for this layer: no Pyro wheel exists here to diff against; it is
anchored by (i) the reference call sites it has to satisfy
(src/rbi/rbi/models/transforms.py:8-46,92-98; src/rbi/rbi/models/base.py:11,320-336;
src/rbi/rbi/utils/distributions.py:5-8,437-483), (ii) structural known-answer tests
(strict autoregressive Jacobian in permutation order, _inverse(_call(x)) == x,
state-dict key layout ``t{i}.nn.layers.{j}.{weight,bias,mask}``, ``t{i}.nn.permutation``),
see tests/test_oracle.py.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import warnings
from abc import ABC, abstractmethod

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.distributions import constraints
from torch.distributions import Transform
from torch.distributions import TransformedDistribution  # pyro's == torch's + mixin
from torch.distributions.transforms import IndependentTransform  # noqa: F401 (re-export)


# --------------------------------------------------------------------------- utils
def clamp_preserve_gradients(x, min, max):
    """pyro/distributions/transforms/utils.py: clamp in the forward value, identity gradient."""
    return x + (x.clamp(min, max) - x).detach()


def sum_leftmost(value, dim):
    """pyro/distributions/util.py::sum_leftmost."""
    if dim < 0:
        dim += value.dim()
    if dim == 0:
        return value
    if dim >= value.dim():
        return value.sum()
    return value.reshape(-1, *value.shape[dim:]).sum(0)


# ------------------------------------------------------------------- MADE masks
def sample_mask_indices(input_dim, hidden_dim, simple=True):
    """pyro/nn/auto_reg_nn.py::sample_mask_indices (simple=True branch is the one used)."""
    indices = torch.linspace(1, input_dim, steps=hidden_dim, device="cpu").to(torch.Tensor().device)
    if simple:
        return torch.round(indices)
    ints = indices.floor()
    ints += torch.bernoulli(indices - ints)
    return ints


def create_mask(input_dim, context_dim, hidden_dims, permutation, output_dim_multiplier):
    """pyro/nn/auto_reg_nn.py::create_mask -- MADE masks for a conditional density."""
    var_index = torch.empty(permutation.shape, dtype=torch.get_default_dtype())
    var_index[permutation] = torch.arange(input_dim, dtype=torch.get_default_dtype())

    input_indices = torch.cat((torch.zeros(context_dim), 1 + var_index))

    if context_dim > 0:
        hidden_indices = [sample_mask_indices(input_dim, h) - 1 for h in hidden_dims]
    else:
        hidden_indices = [sample_mask_indices(input_dim - 1, h) for h in hidden_dims]

    output_indices = (var_index + 1).repeat(output_dim_multiplier)

    mask_skip = (output_indices.unsqueeze(-1) > input_indices.unsqueeze(0)).type_as(var_index)

    masks = [(hidden_indices[0].unsqueeze(-1) >= input_indices.unsqueeze(0)).type_as(var_index)]
    for i in range(1, len(hidden_dims)):
        masks.append(
            (hidden_indices[i].unsqueeze(-1) >= hidden_indices[i - 1].unsqueeze(0)).type_as(var_index)
        )
    masks.append((output_indices.unsqueeze(-1) > hidden_indices[-1].unsqueeze(0)).type_as(var_index))
    return masks, mask_skip


class MaskedLinear(nn.Linear):
    """pyro/nn/auto_reg_nn.py::MaskedLinear -- the mask is re-applied on every call."""

    def __init__(self, in_features, out_features, mask, bias=True):
        super().__init__(in_features, out_features, bias)
        self.register_buffer("mask", mask.data)

    def forward(self, _input):
        masked_weight = self.weight * self.mask
        return F.linear(_input, masked_weight, self.bias)


class ConditionalAutoRegressiveNN(nn.Module):
    """pyro/nn/auto_reg_nn.py::ConditionalAutoRegressiveNN (MADE with a context input)."""

    def __init__(
        self,
        input_dim,
        context_dim,
        hidden_dims,
        param_dims=[1, 1],
        permutation=None,
        skip_connections=False,
        nonlinearity=nn.ReLU(),
    ):
        super().__init__()
        if input_dim == 1:
            warnings.warn(
                "ConditionalAutoRegressiveNN input_dim = 1. "
                "Consider using an affine transformation instead."
            )
        self.input_dim = input_dim
        self.context_dim = context_dim
        self.hidden_dims = hidden_dims
        self.param_dims = param_dims
        self.count_params = len(param_dims)
        self.output_multiplier = sum(param_dims)
        self.all_ones = (torch.tensor(param_dims) == 1).all().item()

        ends = torch.cumsum(torch.tensor(param_dims), dim=0)
        starts = torch.cat((torch.zeros(1).type_as(ends), ends[:-1]))
        self.param_slices = [slice(s.item(), e.item()) for s, e in zip(starts, ends)]

        for h in hidden_dims:
            if h < input_dim:
                raise ValueError("Hidden dimension must not be less than input dimension.")

        if permutation is None:
            P = torch.randperm(input_dim, device="cpu").to(torch.Tensor().device)
        else:
            P = permutation.type(dtype=torch.int64)
        self.register_buffer("permutation", P)

        self.masks, self.mask_skip = create_mask(
            input_dim=input_dim,
            context_dim=context_dim,
            hidden_dims=hidden_dims,
            permutation=self.permutation,
            output_dim_multiplier=self.output_multiplier,
        )

        layers = [MaskedLinear(input_dim + context_dim, hidden_dims[0], self.masks[0])]
        for i in range(1, len(hidden_dims)):
            layers.append(MaskedLinear(hidden_dims[i - 1], hidden_dims[i], self.masks[i]))
        layers.append(
            MaskedLinear(hidden_dims[-1], input_dim * self.output_multiplier, self.masks[-1])
        )
        self.layers = nn.ModuleList(layers)

        if skip_connections:
            self.skip_layer = MaskedLinear(
                input_dim + context_dim,
                input_dim * self.output_multiplier,
                self.mask_skip,
                bias=False,
            )
        else:
            self.skip_layer = None

        self.f = nonlinearity

    def get_permutation(self):
        return self.permutation

    def forward(self, x, context=None):
        if context is None:
            context = self.context
        context = context.expand(x.size()[:-1] + (context.size(-1),))
        x = torch.cat([context, x], dim=-1)
        return self._forward(x)

    def _forward(self, x):
        h = x
        for layer in self.layers[:-1]:
            h = self.f(layer(h))
        h = self.layers[-1](h)

        if self.skip_layer is not None:
            h = h + self.skip_layer(x)

        if self.output_multiplier == 1:
            return h
        h = h.reshape(list(x.size()[:-1]) + [self.output_multiplier, self.input_dim])
        if self.count_params == 1:
            return h
        if self.all_ones:
            return torch.unbind(h, dim=-2)
        return tuple([h[..., s, :] for s in self.param_slices])


class DenseNN(nn.Module):  # imported by the reference (models/transforms.py:8), unused on the MAF path
    def __init__(self, *a, **k):
        raise NotImplementedError("DenseNN is not on the MAF hot path")


# ---------------------------------------------------------- transform plumbing
class TransformModule(Transform, nn.Module):
    """pyro/distributions/torch_transform.py::TransformModule."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

    def __hash__(self):
        return super(nn.Module, self).__hash__()


class ConditionalDistribution(ABC):
    @abstractmethod
    def condition(self, context):
        raise NotImplementedError


class ConditionalTransform(ABC):
    @abstractmethod
    def condition(self, context):
        raise NotImplementedError


class ConditionalTransformModule(ConditionalTransform, nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

    def __hash__(self):
        return super(nn.Module, self).__hash__()


class ConstantConditionalTransform(ConditionalTransform):
    def __init__(self, transform):
        assert isinstance(transform, Transform)
        self.transform = transform

    def condition(self, context):
        return self.transform

    def clear_cache(self):
        self.transform.clear_cache()


class AffineAutoregressive(TransformModule):
    """pyro/distributions/transforms/affine_autoregressive.py::AffineAutoregressive.

    ``_call`` (one MADE pass): y = exp(clamp(log_scale)) * x + mean.
    ``_inverse`` (D MADE passes, in ``arn.permutation`` order):
        x[idx] = (y[idx] - mean[idx]) * exp(-clamp(log_scale[idx])).
    ``cache_size=1``: ``(x, y)`` of the last call and its log-scale are cached, so
    ``log_prob`` of a sample just drawn through ``.inv`` costs no MADE pass.
    """

    domain = constraints.real_vector
    codomain = constraints.real_vector
    bijective = True
    sign = +1
    autoregressive = True

    def __init__(
        self,
        autoregressive_nn,
        log_scale_min_clip=-5.0,
        log_scale_max_clip=3.0,
        sigmoid_bias=2.0,
        stable=False,
    ):
        super().__init__(cache_size=1)
        self.arn = autoregressive_nn
        self._cached_log_scale = None
        self.log_scale_min_clip = log_scale_min_clip
        self.log_scale_max_clip = log_scale_max_clip
        self.sigmoid = nn.Sigmoid()
        self.logsigmoid = nn.LogSigmoid()
        self.sigmoid_bias = sigmoid_bias
        self.stable = stable
        if stable:
            raise NotImplementedError("stable=True is not used by maf_pyro (config/model/maf_pyro.yaml:8)")

    def _call(self, x):
        mean, log_scale = self.arn(x)
        log_scale = clamp_preserve_gradients(log_scale, self.log_scale_min_clip, self.log_scale_max_clip)
        self._cached_log_scale = log_scale
        scale = torch.exp(log_scale)
        y = scale * x + mean
        return y

    def _inverse(self, y):
        x_size = y.size()[:-1]
        perm = self.arn.permutation
        input_dim = y.size(-1)
        x = [torch.zeros(x_size, device=y.device)] * input_dim

        # NOTE: inversion costs one full MADE pass per dimension.
        for idx in perm:
            mean, log_scale = self.arn(torch.stack(x, dim=-1))
            inverse_scale = torch.exp(
                -clamp_preserve_gradients(
                    log_scale[..., idx], min=self.log_scale_min_clip, max=self.log_scale_max_clip
                )
            )
            mean = mean[..., idx]
            x[idx] = (y[..., idx] - mean) * inverse_scale

        x = torch.stack(x, dim=-1)
        log_scale = clamp_preserve_gradients(
            log_scale, min=self.log_scale_min_clip, max=self.log_scale_max_clip
        )
        self._cached_log_scale = log_scale
        return x

    def log_abs_det_jacobian(self, x, y):
        x_old, y_old = self._cached_x_y
        if x is not x_old or y is not y_old:
            # updates the cache and re-computes y and the log-scale
            self(x)

        if self._cached_log_scale is not None:
            log_scale = self._cached_log_scale
        else:
            _, log_scale = self.arn(x)
            log_scale = clamp_preserve_gradients(
                log_scale, self.log_scale_min_clip, self.log_scale_max_clip
            )
        return log_scale.sum(-1)


class Permute(Transform):
    """pyro/distributions/transforms/permute.py::Permute (only with shuffle=True; maf_pyro: false)."""

    bijective = True
    volume_preserving = True

    def __init__(self, permutation, *, dim=-1, cache_size=1):
        super().__init__(cache_size=cache_size)
        if dim >= 0:
            raise ValueError("'dim' keyword argument must be negative")
        self.permutation = permutation
        self.dim = dim

    @constraints.dependent_property(is_discrete=False)
    def domain(self):
        return constraints.independent(constraints.real, -self.dim)

    @constraints.dependent_property(is_discrete=False)
    def codomain(self):
        return constraints.independent(constraints.real, -self.dim)

    @property
    def inv_permutation(self):
        result = torch.empty_like(self.permutation, dtype=torch.long)
        result[self.permutation] = torch.arange(
            self.permutation.size(0), dtype=torch.long, device=self.permutation.device
        )
        return result

    @inv_permutation.setter
    def inv_permutation(self, value):  # the reference assigns it in PyroFlowModel.to (models/base.py:372)
        pass

    def _call(self, x):
        return x.index_select(self.dim, self.permutation)

    def _inverse(self, y):
        return y.index_select(self.dim, self.inv_permutation)

    def clear_cache(self):
        # pyro/distributions/torch_patch.py patches ``Transform.clear_cache`` onto every torch transform
        if self._cache_size == 1:
            self._cached_x_y = None, None

    def log_abs_det_jacobian(self, x, y):
        return torch.zeros(x.size()[: -(-self.dim)], dtype=x.dtype, layout=x.layout, device=x.device)

    def with_cache(self, cache_size=1):
        if self._cache_size == cache_size:
            return self
        return Permute(self.permutation, cache_size=cache_size)


def permute(input_dim, permutation=None, dim=-1):
    if permutation is None:
        permutation = torch.randperm(input_dim)
    return Permute(permutation, dim=dim)


__all__ = [
    "clamp_preserve_gradients",
    "sum_leftmost",
    "sample_mask_indices",
    "create_mask",
    "MaskedLinear",
    "ConditionalAutoRegressiveNN",
    "DenseNN",
    "TransformModule",
    "ConditionalDistribution",
    "ConditionalTransform",
    "ConditionalTransformModule",
    "ConstantConditionalTransform",
    "AffineAutoregressive",
    "Permute",
    "permute",
    "TransformedDistribution",
    "IndependentTransform",
]
