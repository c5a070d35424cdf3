"""Host-side mirror of the reference model interface for ``model=maf_pyro``.

``InverseAffineAutoregressiveModel`` here is a drop-in for ``rbi.models.InverseAffineAutoregressiveModel``
(rbi/rbi/models/flows.py:72-113 over PyroFlowModel, rbi/rbi/models/base.py:274-446): same constructor, same
attribute surface (``input_dim``, ``output_dim``, assignable ``embedding_net``, settable ``output_transform``),
same ``state_dict`` layout (``t{i}.nn.permutation``, ``t{i}.nn.layers.{j}.{weight,bias,mask}``,
``base_dist_mean/scale``) so reference checkpoints load, and ``forward(x)`` returns a
``torch.distributions.Distribution`` (``MAFPosterior``) whose ``log_prob / rsample / sample / expand`` run on the
sm_100a kernels behind include/rabi_b200.h.  Hydra usage: ``module_path: rabi_b200.models``.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn
from torch import Tensor
from torch.distributions import Distribution, constraints
from torch.distributions.transforms import Transform

from . import noise
from ._lib import RabiError
from .engine import MafEngine


# ----------------------------------------------------------------------------------------------- structure
def made_masks(D: int, C: int, hidden: List[int], permutation: Tensor) -> List[Tensor]:
    """MADE masks of one conditional autoregressive net (what Pyro 1.8.4's ``create_mask`` produces for
    ``ConditionalAutoRegressiveNN(D, C, hidden, param_dims=[1, 1])`` with C > 0), as [out, in] float tensors.

    Degrees: context inputs 0, variable d has degree 1 + (its position in ``permutation``); hidden unit i of a
    layer of width h has degree round(linspace(1, D, h))[i] - 1; outputs (mean block then log-scale block)
    have the degree of their variable.  Hidden units connect to inputs of degree <= theirs, outputs to hidden
    units of degree < theirs.
    """
    if C <= 0:
        raise ValueError("the conditional MAF needs a context dimension > 0")
    pos = torch.empty(D, dtype=torch.float32)
    pos[permutation.cpu()] = torch.arange(D, dtype=torch.float32)
    in_deg = torch.cat([torch.zeros(C), pos + 1])
    hid_deg = [torch.round(torch.linspace(1, D, steps=h)) - 1 for h in hidden]
    out_deg = (pos + 1).repeat(2)
    masks = [(hid_deg[0][:, None] >= in_deg[None, :]).float()]
    for a, b in zip(hid_deg[1:], hid_deg[:-1]):
        masks.append((a[:, None] >= b[None, :]).float())
    masks.append((out_deg[:, None] > hid_deg[-1][None, :]).float())
    return masks


class MaskedLinear(nn.Linear):
    """Parameter container with the reference's key names (weight, bias, mask).  The masked product
    ``F.linear(x, W * mask, b)`` of Pyro's MaskedLinear is executed by the CUDA kernels, never here."""

    def __init__(self, in_features: int, out_features: int, mask: Tensor):
        super().__init__(in_features, out_features)
        self.register_buffer("mask", mask)

    def forward(self, *_):
        raise RabiError("MaskedLinear holds parameters only; evaluation happens in librabi_b200 (no torch fallback)")


class _AutoRegressiveNN(nn.Module):
    def __init__(self, D: int, C: int, hidden: List[int]):
        super().__init__()
        for h in hidden:
            if h < D:
                raise ValueError("Hidden dimension must not be less than input dimension.")
        # same construction order (and RNG consumption) as Pyro: permutation first, then the layers
        self.register_buffer("permutation", torch.randperm(D, device="cpu"))
        masks = made_masks(D, C, hidden, self.permutation)
        dims = [C + D] + list(hidden) + [2 * D]
        self.layers = nn.ModuleList(MaskedLinear(dims[i], dims[i + 1], masks[i]) for i in range(len(dims) - 1))


class _ConditionalInverseAffineAutoregressive(nn.Module):
    """Holds ``nn`` so that parameters live under ``t{i}.nn.*`` (rbi/rbi/models/transforms.py:24-35)."""

    def __init__(self, D: int, C: int, hidden: List[int]):
        super().__init__()
        self.nn = _AutoRegressiveNN(D, C, hidden)


class ZScoreLayer(nn.Module):
    """(x - mean) / std; rbi/rbi/models/module.py:109-122.  Fused into the context-projection kernel when it is
    the first module of ``embedding_net`` followed by identities."""

    def __init__(self, mean: Tensor, std: Tensor):
        super().__init__()
        self.register_buffer("mean", mean)
        self.register_buffer("std", std)

    def forward(self, x):
        return (x - self.mean) / self.std


def _is_zscore(m) -> bool:
    return type(m).__name__ == "ZScoreLayer" and hasattr(m, "mean") and hasattr(m, "std")


def _split_embedding(net):
    """-> (zscore_layer | None, remaining module | None).  ``None`` remaining means identity."""
    if net is None or isinstance(net, nn.Identity):
        return None, None
    if _is_zscore(net):
        return net, None
    if isinstance(net, nn.Sequential) and len(net) > 0 and _is_zscore(net[0]):
        rest = [m for m in list(net)[1:] if not isinstance(m, nn.Identity)]
        flat = []
        for m in rest:  # run_train nests Sequential(ZScore, old_embedding); flatten identity-only tails
            z2, r2 = _split_embedding(m)
            if z2 is None and r2 is None:
                continue
            flat.append(m)
        if not flat:
            return net[0], None
        return net[0], (flat[0] if len(flat) == 1 else nn.Sequential(*flat))
    if isinstance(net, nn.Sequential) and all(isinstance(m, nn.Identity) for m in net):
        return None, None
    return None, net


# ----------------------------------------------------------------------------------------------- distribution
class MAFPosterior(Distribution):
    """q(theta | x) for a batch of observations; the object ``model(x)`` returns.

    Mirrors the surface rbi / rbibm use on the Pyro ``TransformedDistribution`` (base.py:417-446): ``batch_shape ==
    x.shape[:-1]``, ``event_shape == [D]``, ``has_rsample``, ``rsample``, ``sample``, ``log_prob`` (broadcasting
    ``[*sample, *batch, D]``), ``expand``; and it takes part in ``torch.distributions.kl_divergence`` through the
    registration in rabi_b200/loss.py (fused MC reverse-KL kernel).
    """

    arg_constraints = {}
    has_rsample = True

    def __init__(self, model: "InverseAffineAutoregressiveModel", ctx_in: Tensor, zscore, batch_shape,
                 output_transform: Optional[Transform], A: Optional[Tensor] = None, A_epoch: Optional[int] = None):
        self.model = model
        self.ctx_in = ctx_in          # [Bf, C] flattened rows: raw x when the z-score is fused, else embedded h
        self.zscore = zscore          # (mean, std) or None
        self.output_transform = output_transform
        self._A = A
        self._A_epoch = A_epoch       # weight epoch of the engine the projection was computed with
        self._last = None             # (theta object, log q) of the last rsample: the "free" log_prob (Q9)
        super().__init__(torch.Size(batch_shape), torch.Size([model.output_dim]), validate_args=False)

    # -- structure
    @property
    def support(self):
        if self.output_transform is not None:
            return self.output_transform.codomain
        return constraints.real_vector

    @property
    def A(self) -> Tensor:
        """Layer-0 context projection of all K transforms, [K, H0, Bf] (computed once per conditioned object)."""
        eng = self.model.engine()   # re-packs the weight image if a parameter changed since the last call
        if self._A is None or self._A_epoch != eng.epoch:   # a projection made with older weights is stale
            zs = self.zscore if self.zscore is not None else (None, None)
            self._A = eng.ctx_project(self.ctx_in, zs[0], zs[1])
            self._A_epoch = eng.epoch
        return self._A

    def expand(self, batch_shape, _instance=None):
        batch_shape = torch.Size(batch_shape)
        if batch_shape == self.batch_shape:
            return self
        Bf = self.ctx_in.shape[0]
        if batch_shape.numel() == Bf:  # pure reshape (e.g. [B] -> [1, B])
            return MAFPosterior(self.model, self.ctx_in, self.zscore, batch_shape, self.output_transform, self._A,
                                self._A_epoch)
        ctx = self.ctx_in.reshape(self.batch_shape + self.ctx_in.shape[-1:]).expand(batch_shape + self.ctx_in.shape[-1:])
        return MAFPosterior(self.model, ctx.reshape(-1, ctx.shape[-1]).contiguous(), self.zscore, batch_shape,
                            self.output_transform)

    # -- helpers
    def _needs_grad(self, *tensors) -> bool:
        if not torch.is_grad_enabled():
            return False
        if any(t is not None and t.requires_grad for t in tensors) or self.ctx_in.requires_grad:
            return True
        return any(p.requires_grad for p in self.model.parameters())

    def _flatten_value(self, value: Tensor):
        D = self.event_shape[0]
        nb = len(self.batch_shape)
        full = torch.broadcast_shapes(value.shape[:-1], self.batch_shape)
        sample_shape = full[: len(full) - nb]
        # explicit sizes: reshape(-1, ...) is ambiguous for empty batches
        v = value.expand(full + (D,)).reshape(int(sample_shape.numel()), int(self.batch_shape.numel()), D)
        return v, full, sample_shape

    # -- Distribution API
    def rsample(self, sample_shape=torch.Size()) -> Tensor:
        sample_shape = torch.Size(sample_shape)
        D = self.event_shape[0]
        Bf = int(self.batch_shape.numel())
        S = int(sample_shape.numel())
        z = noise.standard_normal(sample_shape + self.batch_shape + (D,), self.ctx_in.device).reshape(S, Bf, D)
        if self._needs_grad():
            from .autograd import rsample_training
            theta, logq = rsample_training(self, z)
        else:
            theta, logq = self.model.engine().rsample(self.A, z, want_logq=True)
        out = theta.reshape(sample_shape + self.batch_shape + (D,))
        if self.output_transform is not None:
            y = self.output_transform(out)
            logq = logq.reshape(sample_shape + self.batch_shape) - self.output_transform.log_abs_det_jacobian(out, y)
            out = y
        self._last = (out, logq.reshape(sample_shape + self.batch_shape))
        return out

    def sample(self, sample_shape=torch.Size()) -> Tensor:
        with torch.no_grad():
            return self.rsample(sample_shape)

    def log_prob(self, value: Tensor) -> Tensor:
        if self._last is not None and value is self._last[0]:
            # samples just drawn from this very object: no MADE pass (the reference's cache hit) -- but only
            # when the cached value carries what the caller can ask of it: a sample drawn under no_grad has a detached
            # log q, and a later differentiable log_prob(theta) must be recomputed with its graph
            if self._last[1].requires_grad or not self._needs_grad(value):
                return self._last[1]
        nb = len(self.batch_shape)
        full = torch.broadcast_shapes(value.shape[:-1], self.batch_shape)
        if nb and full[len(full) - nb:] != self.batch_shape:
            # the value's batch dims are larger than ours (e.g. q = model(x[1, dx]); q.log_prob(thetas[N, D])): Pyro's
            # ConditionalAutoRegressiveNN expands the context, so do we
            return self.expand(full[len(full) - nb:]).log_prob(value)
        ldj = None
        if self.output_transform is not None:
            u = self.output_transform.inv(value)
            ldj = self.output_transform.log_abs_det_jacobian(u, value)
            value = u
        v, full, _ = self._flatten_value(value)
        if self._needs_grad(v):
            from .autograd import log_prob_training
            lp = log_prob_training(self, v)
        else:
            lp = self.model.engine().log_prob(self.A, v)
        lp = lp.reshape(full)
        if ldj is not None:
            lp = lp - ldj
        return lp


def _map_transform_tensors(obj, fn, _seen=None):
    """Apply ``fn`` to every tensor held by a (possibly composed) torch.distributions transform, in place."""
    if obj is None:
        return
    _seen = set() if _seen is None else _seen
    if id(obj) in _seen:
        return
    _seen.add(id(obj))
    if isinstance(obj, (list, tuple)):
        for o in obj:
            _map_transform_tensors(o, fn, _seen)
        return
    if not isinstance(obj, (Transform, constraints.Constraint)):
        return
    for name, val in list(vars(obj).items()):
        if name == "_inv":
            continue
        if isinstance(val, Tensor):
            setattr(obj, name, fn(val))
        elif name == "_cached_x_y":
            obj._cached_x_y = (None, None)
        elif isinstance(val, (Transform, constraints.Constraint, list, tuple)):
            _map_transform_tensors(val, fn, _seen)


# ----------------------------------------------------------------------------------------------- model
class InverseAffineAutoregressiveModel(nn.Module):
    """Masked autoregressive flow posterior ("MAF": fast density, sequential sampling) on sm_100a kernels.

    Args mirror rbi/rbi/models/flows.py:73-84; ``**kwargs`` takes ``log_scale_min_clip`` / ``log_scale_max_clip`` /
    ``stable`` exactly as config/model/maf_pyro.yaml passes them (Pyro defaults -5 / 3 / False).
    """

    def __init__(
        self,
        input_dim: int,
        output_dim: int,
        hidden_dims: List[int] = [50, 50],
        num_transforms: int = 5,
        shuffle: bool = True,
        with_cache: bool = False,
        output_transform: Optional[Transform] = None,
        embedding_net: Optional[nn.Module] = None,
        **kwargs,
    ):
        super().__init__()
        self.embedding_net = embedding_net if embedding_net is not None else nn.Identity()
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.with_cache = with_cache
        self.hidden_dims = list(hidden_dims)
        self.num_transforms = num_transforms
        self._kwargs = dict(kwargs)
        if kwargs.get("stable", False):
            raise NotImplementedError("stable=True is not on the maf_pyro path (config/model/maf_pyro.yaml:8)")
        self.shuffle = bool(shuffle)
        self.log_scale_min_clip = float(kwargs.get("log_scale_min_clip", -5.0))
        self.log_scale_max_clip = float(kwargs.get("log_scale_max_clip", 3.0))
        # context dimension is probed like the reference does (base.py:314)
        ctx_dim = self.embedding_net(torch.zeros((1, input_dim))).shape[-1]
        self.context_dim = int(ctx_dim)
        for i in range(num_transforms):
            self.add_module("t" + str(i), _ConditionalInverseAffineAutoregressive(output_dim, self.context_dim,
                                                                                 self.hidden_dims))
            if shuffle:  # Permute layer after every transform (base.py:326-333); folded into kernel indices
                perm = torch.randperm(output_dim)
                inv = torch.empty_like(perm)
                inv[perm] = torch.arange(output_dim)
                self.register_buffer("permute" + str(i), perm)
                self.register_buffer("inv_permute" + str(i), inv)
        self.register_buffer("base_dist_mean", torch.zeros(output_dim))
        self.register_buffer("base_dist_scale", torch.ones(output_dim))
        self._output_transform = output_transform
        self._engine: Optional[MafEngine] = None

    # -- attribute surface the runners rely on (run_train.py:111,123-128,250-251)
    @property
    def output_transform(self):
        return self._output_transform

    @output_transform.setter
    def output_transform(self, transform):
        self._output_transform = transform

    def forward_embedding(self, x: Tensor) -> Tensor:
        return self.embedding_net(x)

    # -- engine management
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engine"] = None  # device handles are rebuilt lazily; keeps deepcopy / torch.save(model) working
        state.pop("_ag_struct", None)
        return state

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._engine = None
        self.__dict__.pop("_ag_struct", None)
        # the output transform is not a module: move its tensors along (the reference's ``to`` does the same through
        # ``move_all_tensor_to_device``, models/base.py:365-382)
        _map_transform_tensors(self._output_transform, fn)
        return out

    def load_state_dict(self, state_dict, *a, **k):
        out = super().load_state_dict(state_dict, *a, **k)
        self._engine = None  # permutation / masks may have changed
        return out

    def transforms_params(self):
        return [[(l.weight, l.bias) for l in getattr(self, "t" + str(i)).nn.layers] for i in range(self.num_transforms)]

    def invalidate(self):
        """Force a re-pack of the device weight image on the next call.  Needed only after in-place updates that do
        not bump the parameters' version counters (``p.data.copy_()`` / ``p.data.mul_()``, e.g. EMA-of-weights or
        weight-clipping code); optimizer steps, ``load_state_dict`` and ``.to()`` are detected automatically."""
        if self._engine is not None:
            self._engine.invalidate()

    def engine(self) -> MafEngine:
        dev = self.base_dist_mean.device
        if dev.type != "cuda":
            raise RabiError("rabi_b200 has no CPU path: move the model to a CUDA device (model.to('cuda'))")
        want = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        if self._engine is None or self._engine.device != want:
            eng = MafEngine(self.num_transforms, self.output_dim, self.context_dim, self.hidden_dims,
                            self.log_scale_min_clip, self.log_scale_max_clip, dev)
            nets = [getattr(self, "t" + str(i)).nn for i in range(self.num_transforms)]
            post = [getattr(self, "permute" + str(i)) for i in range(self.num_transforms)] if self.shuffle else None
            eng.set_structure([n.permutation for n in nets], [[l.mask for l in n.layers] for n in nets], post)
            self._engine = eng
        self._engine.set_weights(self.transforms_params())
        return self._engine

    # -- forward
    def condition_inputs(self, x: Tensor):
        """-> (ctx_in [Bf, C'], zscore (mean, std) | None, batch_shape).  The z-score layer is fused into the
        projection kernel when the rest of the embedding is the identity."""
        if x.shape == torch.Size([]):
            x = x.reshape(1, 1)
        zl, rest = _split_embedding(self.embedding_net)
        batch_shape = x.shape[:-1]
        if rest is None:
            flat = x.reshape(-1, x.shape[-1])
            zs = (zl.mean.float().contiguous(), zl.std.float().contiguous()) if zl is not None else None
            return flat, zs, batch_shape
        h = self.embedding_net(x)
        return h.reshape(-1, h.shape[-1]), None, h.shape[:-1]

    def forward(self, x: Tensor) -> Distribution:
        ctx_in, zs, batch_shape = self.condition_inputs(x)
        if ctx_in.shape[-1] != self.context_dim:
            raise ValueError(f"embedding produced {ctx_in.shape[-1]} features, the flow was built for {self.context_dim}")
        return MAFPosterior(self, ctx_in.float().contiguous(), zs, batch_shape, self._output_transform)
