"""ORACLE (test infrastructure, not product code).

Makes the UNMODIFIED reference package ``rbi`` (under /root/reference/src/rbi) importable in the
build container, where its third-party dependencies (pyro, advertorch, sbi, zuko, cvxpy) are
absent: the arithmetic the MAF path needs from them is served by the restatements in
``pyro_restated`` / ``advertorch_restated``; every other imported name resolves to an inert
placeholder class (those code paths -- sbi posteriors, zuko flows, C2ST -- are not on the hot path and
are never called).  With this in place the reference's own ``rbi`` code (model/loss/attack/defense/
fisher glue) runs here and generates the golden vectors (oracle/gen_golden.py).

This only works where /root/reference exists (the build container); nothing at run time on the GPU box
may call it.
"""
from __future__ import annotations

import importlib.machinery
import sys
import types

REFERENCE_SRC = "/root/reference/src/rbi"


class _Placeholder:
    """Inert stand-in for a third-party symbol that is imported but never used on the MAF path."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return None


class _AutoModule(types.ModuleType):
    __path__: list = []

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        cls = type(name, (_Placeholder,), {"__module__": self.__name__})
        setattr(self, name, cls)
        return cls


def _module(name, **attrs):
    m = _AutoModule(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


def install(reference_src: str = REFERENCE_SRC):
    """Register stub modules and put the reference's ``rbi`` on sys.path.  Idempotent."""
    if getattr(install, "_done", False):
        return
    from torch.distributions import constraints
    import torch.distributions.transforms as tt

    from . import advertorch_restated as adv
    from . import pyro_restated as pr

    def move_all_tensor_to_device(obj, device):  # sbi helper: no-op on CPU
        return None

    # pyro/distributions/torch_patch.py gives every torch Transform a ``clear_cache`` (PyroFlowModel.to calls it on the
    # output transform as well)
    if not hasattr(tt.Transform, "clear_cache"):
        def _transform_clear_cache(self):
            if getattr(self, "_cache_size", 0) == 1:
                self._cached_x_y = None, None
            for part in getattr(self, "parts", []) or ([self.base_transform] if hasattr(self, "base_transform") else []):
                part.clear_cache()

        tt.Transform.clear_cache = _transform_clear_cache

    # ---- pyro
    _module("pyro")
    _module(
        "pyro.nn",
        ConditionalAutoRegressiveNN=pr.ConditionalAutoRegressiveNN,
        DenseNN=pr.DenseNN,
    )
    _module(
        "pyro.distributions",
        constraints=constraints,
        TransformedDistribution=pr.TransformedDistribution,
        ConditionalDistribution=pr.ConditionalDistribution,
    )
    _module("pyro.distributions.constraints", **{k: getattr(constraints, k) for k in dir(constraints) if not k.startswith("__")})
    _module(
        "pyro.distributions.transforms",
        AffineAutoregressive=pr.AffineAutoregressive,
        Transform=tt.Transform,
        IndependentTransform=tt.IndependentTransform,
        ComposeTransform=tt.ComposeTransform,
        permute=pr.permute,
        Permute=pr.Permute,
    )
    _module(
        "pyro.distributions.conditional",
        ConditionalDistribution=pr.ConditionalDistribution,
        ConditionalTransform=pr.ConditionalTransform,
        ConditionalTransformModule=pr.ConditionalTransformModule,
        ConstantConditionalTransform=pr.ConstantConditionalTransform,
    )
    _module("pyro.distributions.util", sum_leftmost=pr.sum_leftmost)

    # ---- advertorch
    _module("advertorch")
    attacks = _module("advertorch.attacks")
    base = _module("advertorch.attacks.base", Attack=adv.Attack, LabelMixin=adv.LabelMixin)
    for name in (
        "Attack", "PGDAttack", "LinfPGDAttack", "L2PGDAttack", "L1PGDAttack", "L2BasicIterativeAttack",
        "LinfBasicIterativeAttack", "MomentumIterativeAttack", "L2MomentumIterativeAttack",
        "LinfMomentumIterativeAttack",
    ):
        setattr(attacks, name, getattr(adv, name))
    attacks.base = base
    _module(
        "advertorch.attacks.iterative_projected_gradient",
        **{k: getattr(adv, k) for k in (
            "PGDAttack", "perturb_iterative", "MomentumIterativeAttack", "L2MomentumIterativeAttack",
            "LinfMomentumIterativeAttack", "L2BasicIterativeAttack", "LinfBasicIterativeAttack")},
    )
    _module(
        "advertorch.utils",
        clamp=adv.clamp,
        normalize_by_pnorm=adv.normalize_by_pnorm,
        clamp_by_pnorm=adv.clamp_by_pnorm,
        batch_multiply=adv.batch_multiply,
        batch_clamp=adv.batch_clamp,
    )

    # ---- sbi / zuko / cvxpy: imported by the reference, unused on the MAF path
    for name in (
        "sbi", "sbi.inference", "sbi.inference.posteriors", "sbi.inference.posteriors.vi_posterior",
        "sbi.samplers", "sbi.samplers.vi", "sbi.utils", "sbi.utils.metrics",
        "sbi.utils.user_input_checks_utils", "zuko", "zuko.flows", "cvxpy",
    ):
        _module(name)
    _module("sbi.samplers.vi.vi_utils", move_all_tensor_to_device=move_all_tensor_to_device)

    if reference_src not in sys.path:
        sys.path.insert(0, reference_src)
    install._done = True


REFERENCE_RBIBM_SRC = "/root/reference/src/rbibm"


def install_rbibm(reference_src: str = REFERENCE_RBIBM_SRC):
    """Additionally make the reference's ``rbibm.metric`` modules importable (metric base classes, robustness and
    approximation metrics).  ``rbibm.tasks`` (simulators; pulls in MCMC code that imports ``turtle``/tkinter by accident)
    and ``rbibm.utils`` helpers outside batched processing are not needed by the metrics and become placeholders."""
    install()
    if getattr(install_rbibm, "_done", False):
        return
    _module("turtle")
    if reference_src not in sys.path:
        sys.path.insert(0, reference_src)
    import rbibm  # the real (empty-ish) package __init__
    _module("rbibm.tasks")
    _module("rbibm.tasks.base")
    install_rbibm._done = True
