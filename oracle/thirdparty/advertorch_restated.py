"""ORACLE (test infrastructure, not product code).

CPU restatement of the ``advertorch`` pieces RABI's L2-PGD attack runs on.  advertorch is
installed by the reference from un-pinned git master (/root/reference/src/requirements.txt:29),
is not vendored and is absent from this image, so the published algorithm is restated
(advertorch/attacks/iterative_projected_gradient.py::perturb_iterative, PGDAttack, L2PGDAttack;
advertorch/attacks/base.py::Attack, LabelMixin; advertorch/utils.py::clamp, batch_multiply,
_get_norm_batch, normalize_by_pnorm, clamp_by_pnorm, batch_clamp) -- "[3P-memory]" in SURVEY 8c.
PARITY UNPINNED for this layer; anchored on the reference call site
src/rbi/rbi/attacks/advertorch_attack.py:77-118 (keyword names, ord=2 branch, minimize flag) and
on the reference test's constraint check (src/rbi/tests/test_attacks.py:56-57).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


# ------------------------------------------------------------------ advertorch.utils
def clamp(input, min=None, max=None):
    if min is not None:
        if isinstance(min, (float, int)):
            input = torch.clamp(input, min=min)
        else:
            input = torch.max(input, min)
    if max is not None:
        if isinstance(max, (float, int)):
            input = torch.clamp(input, max=max)
        else:
            input = torch.min(input, max)
    return input


def _batch_multiply_tensor_by_vector(vector, batch_tensor):
    return (batch_tensor.transpose(0, -1) * vector).transpose(0, -1).contiguous()


def batch_multiply(float_or_vector, tensor):
    if isinstance(float_or_vector, torch.Tensor):
        assert len(float_or_vector) == len(tensor)
        tensor = _batch_multiply_tensor_by_vector(float_or_vector, tensor)
    elif isinstance(float_or_vector, float):
        tensor = tensor * float_or_vector
    else:
        raise TypeError("Value has to be float or torch.Tensor")
    return tensor


def _batch_clamp_tensor_by_vector(vector, batch_tensor):
    return torch.min(torch.max(batch_tensor.transpose(0, -1), -vector), vector).transpose(0, -1).contiguous()


def batch_clamp(float_or_vector, tensor):
    if isinstance(float_or_vector, torch.Tensor):
        assert len(float_or_vector) == len(tensor)
        return _batch_clamp_tensor_by_vector(float_or_vector, tensor)
    if isinstance(float_or_vector, float):
        return clamp(tensor, -float_or_vector, float_or_vector)
    raise TypeError("Value has to be float or torch.Tensor")


def _get_norm_batch(x, p):
    batch_size = x.size(0)
    return x.abs().pow(p).view(batch_size, -1).sum(dim=1).pow(1.0 / p)


def normalize_by_pnorm(x, p=2, small_constant=1e-6):
    assert isinstance(p, (float, int))
    norm = _get_norm_batch(x, p)
    norm = torch.max(norm, torch.ones_like(norm) * small_constant)
    return batch_multiply(1.0 / norm, x)


def clamp_by_pnorm(x, p, r):
    assert isinstance(p, (float, int))
    norm = _get_norm_batch(x, p)
    if isinstance(r, torch.Tensor):
        assert norm.size() == r.size()
    else:
        assert isinstance(r, float)
    factor = torch.min(r / norm, torch.ones_like(norm))
    return batch_multiply(factor, x)


def is_float_or_torch_tensor(x):
    return isinstance(x, (torch.Tensor, float))


# --------------------------------------------------------- advertorch.attacks.utils
def rand_init_delta(delta, x, ord, eps, clip_min, clip_max):
    """advertorch/attacks/utils.py::rand_init_delta (ord=inf and ord=2 branches)."""
    if isinstance(eps, torch.Tensor):
        assert len(eps) == len(delta)
    if ord == np.inf:
        delta.data.uniform_(-1, 1)
        delta.data = batch_multiply(eps, delta.data)
    elif ord == 2:
        delta.data.uniform_(clip_min, clip_max)
        delta.data = delta.data - x
        delta.data = clamp_by_pnorm(delta.data, ord, eps)
    else:
        raise NotImplementedError("Only ord = inf and ord = 2 are restated (L2-PGD path).")
    delta.data = clamp(x + delta.data, min=clip_min, max=clip_max) - x
    return delta.data


# ------------------------------------------------- iterative_projected_gradient
def perturb_iterative(
    xvar,
    yvar,
    predict,
    nb_iter,
    eps,
    eps_iter,
    loss_fn,
    delta_init=None,
    minimize=False,
    ord=np.inf,
    clip_min=0.0,
    clip_max=1.0,
    l1_sparsity=None,
):
    """advertorch PGD loop; only the ord=2 and ord=inf branches are restated."""
    if delta_init is not None:
        delta = delta_init
    else:
        delta = torch.zeros_like(xvar)

    delta.requires_grad_()
    for ii in range(nb_iter):
        outputs = predict(xvar + delta)
        loss = loss_fn(outputs, yvar)
        if minimize:
            loss = -loss

        loss.backward()
        if ord == np.inf:
            grad_sign = delta.grad.data.sign()
            delta.data = delta.data + batch_multiply(eps_iter, grad_sign)
            delta.data = batch_clamp(eps, delta.data)
            delta.data = clamp(xvar.data + delta.data, clip_min, clip_max) - xvar.data
        elif ord == 2:
            grad = delta.grad.data
            grad = normalize_by_pnorm(grad)
            delta.data = delta.data + batch_multiply(eps_iter, grad)
            delta.data = clamp(xvar.data + delta.data, clip_min, clip_max) - xvar.data
            if eps is not None:
                delta.data = clamp_by_pnorm(delta.data, ord, eps)
        else:
            raise NotImplementedError("Only ord = inf and ord = 2 are restated (L2-PGD path).")
        delta.grad.data.zero_()

    x_adv = clamp(xvar + delta, clip_min, clip_max)
    return x_adv


class Attack(object):
    """advertorch/attacks/base.py::Attack."""

    def __init__(self, predict, loss_fn, clip_min, clip_max):
        self.predict = predict
        self.loss_fn = loss_fn
        self.clip_min = clip_min
        self.clip_max = clip_max

    def perturb(self, x, **kwargs):
        raise NotImplementedError("Sub-classes must implement perturb.")

    def __call__(self, *args, **kwargs):
        return self.perturb(*args, **kwargs)


class LabelMixin(object):
    def _get_predicted_label(self, x):
        with torch.no_grad():
            outputs = self.predict(x)
        _, y = torch.max(outputs, dim=1)
        return y

    def _verify_and_process_inputs(self, x, y):
        if self.targeted:
            assert y is not None
        if not self.targeted:
            if y is None:
                y = self._get_predicted_label(x)
        x = x.detach().clone()
        y = y.detach().clone()
        return x, y


class PGDAttack(Attack, LabelMixin):
    def __init__(
        self,
        predict,
        loss_fn=None,
        eps=0.3,
        nb_iter=40,
        eps_iter=0.01,
        rand_init=True,
        clip_min=0.0,
        clip_max=1.0,
        ord=np.inf,
        l1_sparsity=None,
        targeted=False,
    ):
        super(PGDAttack, self).__init__(predict, loss_fn, clip_min, clip_max)
        self.eps = eps
        self.nb_iter = nb_iter
        self.eps_iter = eps_iter
        self.rand_init = rand_init
        self.ord = ord
        self.targeted = targeted
        if self.loss_fn is None:
            self.loss_fn = nn.CrossEntropyLoss(reduction="sum")
        self.l1_sparsity = l1_sparsity
        assert is_float_or_torch_tensor(self.eps_iter)
        assert is_float_or_torch_tensor(self.eps)

    def perturb(self, x, y=None):
        """advertorch's own PGDAttack.perturb.

        NOTE (verified by executing the reference over these stubs, oracle/gen_golden.py): rbi's
        ``setattr(PGDAttack, "perturb", pgd_randint_change_perturb)`` (attacks/advertorch_attack.py:183)
        lands on the *re-bound, derived* ``PGDAttack`` created by the metaclass loop (:32-42), so
        ``rbi.attacks.L2PGDAttack`` -- whose bases are the original advertorch classes -- still resolves
        ``perturb`` to THIS method: delta0 = rand_init_delta (uniform in the clip box minus x, projected
        to the L2 ball), not ``sample_lp_uniformly``.
        """
        x, y = self._verify_and_process_inputs(x, y)

        delta = torch.zeros_like(x)
        delta = nn.Parameter(delta)
        if self.rand_init:
            rand_init_delta(delta, x, self.ord, self.eps, self.clip_min, self.clip_max)
            delta.data = clamp(x + delta.data, min=self.clip_min, max=self.clip_max) - x

        rval = perturb_iterative(
            x, y, self.predict, nb_iter=self.nb_iter, eps=self.eps, eps_iter=self.eps_iter,
            loss_fn=self.loss_fn, minimize=self.targeted, ord=self.ord, clip_min=self.clip_min,
            clip_max=self.clip_max, delta_init=delta, l1_sparsity=self.l1_sparsity)
        return rval.data


class LinfPGDAttack(PGDAttack):
    def __init__(self, predict, loss_fn=None, eps=0.3, nb_iter=40, eps_iter=0.01, rand_init=True,
                 clip_min=0.0, clip_max=1.0, targeted=False):
        super(LinfPGDAttack, self).__init__(
            predict=predict, loss_fn=loss_fn, eps=eps, nb_iter=nb_iter, eps_iter=eps_iter,
            rand_init=rand_init, clip_min=clip_min, clip_max=clip_max, targeted=targeted, ord=np.inf)


class L2PGDAttack(PGDAttack):
    def __init__(self, predict, loss_fn=None, eps=0.3, nb_iter=40, eps_iter=0.01, rand_init=True,
                 clip_min=0.0, clip_max=1.0, targeted=False):
        super(L2PGDAttack, self).__init__(
            predict=predict, loss_fn=loss_fn, eps=eps, nb_iter=nb_iter, eps_iter=eps_iter,
            rand_init=rand_init, clip_min=clip_min, clip_max=clip_max, targeted=targeted, ord=2)


class L1PGDAttack(PGDAttack):
    def __init__(self, predict, loss_fn=None, eps=10.0, nb_iter=40, eps_iter=0.01, rand_init=True,
                 clip_min=0.0, clip_max=1.0, targeted=False):
        super(L1PGDAttack, self).__init__(
            predict=predict, loss_fn=loss_fn, eps=eps, nb_iter=nb_iter, eps_iter=eps_iter,
            rand_init=rand_init, clip_min=clip_min, clip_max=clip_max, targeted=targeted, ord=1,
            l1_sparsity=None)


class L2BasicIterativeAttack(PGDAttack):
    def __init__(self, predict, loss_fn=None, eps=0.1, nb_iter=10, eps_iter=0.05,
                 clip_min=0.0, clip_max=1.0, targeted=False):
        super(L2BasicIterativeAttack, self).__init__(
            predict, loss_fn, eps, nb_iter, eps_iter, False, clip_min, clip_max, 2, None, targeted)


class LinfBasicIterativeAttack(PGDAttack):
    def __init__(self, predict, loss_fn=None, eps=0.1, nb_iter=10, eps_iter=0.05,
                 clip_min=0.0, clip_max=1.0, targeted=False):
        super(LinfBasicIterativeAttack, self).__init__(
            predict, loss_fn, eps, nb_iter, eps_iter, False, clip_min, clip_max, np.inf, None, targeted)


class MomentumIterativeAttack(Attack, LabelMixin):
    def __init__(self, predict, loss_fn=None, eps=0.3, nb_iter=40, decay_factor=1.0, eps_iter=0.01,
                 clip_min=0.0, clip_max=1.0, targeted=False, ord=np.inf):
        super(MomentumIterativeAttack, self).__init__(predict, loss_fn, clip_min, clip_max)
        self.eps = eps
        self.nb_iter = nb_iter
        self.decay_factor = decay_factor
        self.eps_iter = eps_iter
        self.targeted = targeted
        self.ord = ord


class L2MomentumIterativeAttack(MomentumIterativeAttack):
    def __init__(self, predict, loss_fn=None, eps=0.3, nb_iter=40, decay_factor=1.0, eps_iter=0.01,
                 clip_min=0.0, clip_max=1.0, targeted=False):
        super(L2MomentumIterativeAttack, self).__init__(
            predict, loss_fn, eps, nb_iter, decay_factor, eps_iter, clip_min, clip_max, targeted, 2)


class LinfMomentumIterativeAttack(MomentumIterativeAttack):
    def __init__(self, predict, loss_fn=None, eps=0.3, nb_iter=40, decay_factor=1.0, eps_iter=0.01,
                 clip_min=0.0, clip_max=1.0, targeted=False):
        super(LinfMomentumIterativeAttack, self).__init__(
            predict, loss_fn, eps, nb_iter, decay_factor, eps_iter, clip_min, clip_max, targeted, np.inf)
