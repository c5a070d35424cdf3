"""Differentiable (training-side) evaluation of the MAF posterior: gradients w.r.t. weights, context and theta, to any
order.

Everything the defenses need -- the NLL gradient (rbi/rbi/loss/train_loss.py:16-29), the TRADES regulariser's gradient
through both posteriors and through the reparameterised samples (defenses/trades_regularizers.py:58-64), the Fisher
score d log q / dx and the *second-order* gradient of its squared norm (utils/fisher_info.py:185-280,
defenses/fisher_regularization.py:67-95) -- is built from ONE CUDA primitive, ``rabi_sgemm``, wrapped in an
``autograd.Function`` whose backward is expressed with the same function, so autograd can differentiate it again.
All matrix products (the FLOPs) therefore run in the hand-written kernel at every order of differentiation; torch
only supplies the elementwise glue (ReLU, exp, clamp with straight-through gradient, sums) and the graph.

Relative to the reference this path (i) multiplies the masks into the weights once per call instead of once per MADE
pass, and (ii) draws samples with ONE pass-equivalent: at autoregressive step j only the hidden units of MADE degree
j are evaluated (Pyro's ``AffineAutoregressive._inverse`` runs D full passes).  The attack / metric side does not
use this module at all -- it runs on the fused kernels (rabi_b200/attacks.py, rabi_b200/loss.py).
"""
from __future__ import annotations

import ctypes
import math
from typing import List, Tuple

import torch
from torch import Tensor

from . import _lib


def _rowmajor(t: Tensor) -> Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    if t.dim() != 2:
        raise ValueError("mm expects 2-D tensors")
    if t.stride(1) != 1 or t.stride(0) < max(1, t.shape[1]):
        t = t.contiguous()
    return t


def _sgemm(A: Tensor, B: Tensor, ta: bool, tb: bool) -> Tensor:
    A, B = _rowmajor(A), _rowmajor(B)
    M, K = (A.shape[1], A.shape[0]) if ta else (A.shape[0], A.shape[1])
    Kb, N = (B.shape[1], B.shape[0]) if tb else (B.shape[0], B.shape[1])
    if K != Kb:
        raise ValueError(f"mm: inner dimensions differ ({K} vs {Kb})")
    if not A.is_cuda:
        raise _lib.RabiError("rabi_b200 has no CPU path: tensors must live on a CUDA device")
    C = torch.empty((M, N), device=A.device, dtype=torch.float32)
    if M and N:
        if K == 0:
            return C.zero_()
        L = _lib.lib()
        idx = A.device.index if A.device.index is not None else torch.cuda.current_device()
        rc = L.rabi_sgemm(idx, int(ta), int(tb), M, N, K, ctypes.c_void_p(A.data_ptr()), A.stride(0),
                          ctypes.c_void_p(B.data_ptr()), B.stride(0), ctypes.c_void_p(C.data_ptr()), C.stride(0),
                          0.0, ctypes.c_void_p(torch.cuda.current_stream(A.device).cuda_stream))
        if rc != 0:
            raise _lib.RabiError(L.rabi_last_error(None).decode())
    return C


class _MM(torch.autograd.Function):
    """C = op(A) @ op(B) on ``rabi_sgemm``; backward calls ``mm`` again, hence differentiable to any order."""

    @staticmethod
    def forward(ctx, A, B, ta: bool, tb: bool):
        ctx.save_for_backward(A, B)
        ctx.ta, ctx.tb = ta, tb
        return _sgemm(A.detach(), B.detach(), ta, tb)

    @staticmethod
    def backward(ctx, dC):
        A, B = ctx.saved_tensors
        ta, tb = ctx.ta, ctx.tb
        dA = dB = None
        if ctx.needs_input_grad[0]:
            if not ta and not tb:
                dA = mm(dC, B, False, True)
            elif not ta and tb:
                dA = mm(dC, B, False, False)
            elif ta and not tb:
                dA = mm(B, dC, False, True)
            else:
                dA = mm(B, dC, True, True)
        if ctx.needs_input_grad[1]:
            if not ta and not tb:
                dB = mm(A, dC, True, False)
            elif not ta and tb:
                dB = mm(dC, A, True, False)
            elif ta and not tb:
                dB = mm(A, dC, False, False)
            else:
                dB = mm(dC, A, True, True)
        return dA, dB, None, None


def mm(A: Tensor, B: Tensor, ta: bool = False, tb: bool = False) -> Tensor:
    return _MM.apply(A, B, ta, tb)


def linear(x: Tensor, W: Tensor, b: Tensor = None) -> Tensor:
    """x [M, in] @ W[out, in]^T (+ b)."""
    y = mm(x, W, False, True)
    return y if b is None else y + b


def clamp_preserve_gradients(x: Tensor, lo: float, hi: float) -> Tensor:
    return x + (x.clamp(lo, hi) - x).detach()


# ----------------------------------------------------------------------------------------------- structure
class _Struct:
    """Per-transform MADE bookkeeping derived from the mask buffers (cached on the model)."""

    def __init__(self, model):
        self.K, self.D, self.C = model.num_transforms, model.output_dim, model.context_dim
        self.perm: List[List[int]] = []
        self.groups: List[List[List[int]]] = []  # [k][layer][j] -> start of degree group j (len D+1)
        for k in range(self.K):
            net = getattr(model, "t" + str(k)).nn
            perm = [int(v) for v in net.permutation.tolist()]
            self.perm.append(perm)
            gl = []
            deg = None
            for l, layer in enumerate(net.layers[:-1]):
                m = layer.mask.detach().cpu()
                if l == 0:
                    deg = m[:, self.C:].sum(1).long()          # number of visible variables == MADE degree
                else:
                    # unit degree = the degree group of the previous layer its prefix ends at
                    cnt = m.sum(1).long()
                    prev = gl[-1]
                    deg = torch.tensor([prev.index(int(c), 1) - 1 for c in cnt])
                if (deg[1:] < deg[:-1]).any():
                    raise ValueError("MADE hidden degrees are not sorted; not a Pyro create_mask structure")
                starts = [int((deg < j).sum()) for j in range(self.D + 1)]
                starts[self.D] = int(deg.numel())
                gl.append(starts)
            self.groups.append(gl)
        dev = model.base_dist_mean.device
        self.perm_prefix = [[torch.as_tensor(p[:j], dtype=torch.long, device=dev) for j in range(self.D)] for p in self.perm]
        self.out_rows = [[torch.as_tensor([p[j], self.D + p[j]], dtype=torch.long, device=dev) for j in range(self.D)]
                         for p in self.perm]
        self.device = dev
        self.post = None
        if getattr(model, "shuffle", False):
            self.post = [getattr(model, "permute" + str(k)) for k in range(self.K)]
            self.post_inv = [getattr(model, "inv_permute" + str(k)) for k in range(self.K)]


def _struct(model) -> _Struct:
    st = getattr(model, "_ag_struct", None)
    if st is None or st.device != model.base_dist_mean.device:
        st = _Struct(model)
        object.__setattr__(model, "_ag_struct", st)
    return st


def _masked(model, k):
    net = getattr(model, "t" + str(k)).nn
    return [(l.weight * l.mask, l.bias) for l in net.layers]


def _context(post) -> Tensor:
    h = post.ctx_in
    if post.zscore is not None:
        h = (h - post.zscore[0]) / post.zscore[1]
    return h


HALF_LOG_2PI = 0.5 * math.log(2 * math.pi)


# ----------------------------------------------------------------------------------------------- density pass
def _log_prob_ops(post, v: Tensor) -> Tensor:
    """v [S, Bf, D] (unconstrained space) -> log q(v | ctx) [S, Bf]; K full MADE passes."""
    model = post.model
    st = _struct(model)
    S, Bf, D = v.shape
    C = st.C
    h = _context(post)                                   # [Bf, C]
    y = v.reshape(S * Bf, D)
    logdet = 0.0
    lo, hi = model.log_scale_min_clip, model.log_scale_max_clip
    for k in range(st.K - 1, -1, -1):
        if st.post is not None:
            y = y.index_select(-1, st.post_inv[k])
        layers = _masked(model, k)
        W0, b0 = layers[0]
        A = linear(h, W0[:, :C], b0)                      # [Bf, H0], shared by the S samples of a row
        a = linear(y, W0[:, C:]) + A.unsqueeze(0).expand(S, Bf, A.shape[-1]).reshape(S * Bf, -1)
        hcur = torch.relu(a)
        for (W, b) in layers[1:-1]:
            hcur = torch.relu(linear(hcur, W, b))
        Wo, bo = layers[-1]
        out = linear(hcur, Wo, bo)
        m, s = out[:, :D], out[:, D:]
        sh = clamp_preserve_gradients(s, lo, hi)
        y = torch.exp(sh) * y + m
        logdet = logdet + sh.sum(-1)
    base = (-0.5 * y * y - HALF_LOG_2PI).sum(-1)
    return (base + logdet).reshape(S, Bf)


def log_prob_differentiable(post, v: Tensor) -> Tensor:
    """The op-by-op (any-order differentiable) density pass; kept as the public name of the slow path."""
    return _log_prob_ops(post, v)


# ----------------------------------------------------------------------------------------------- sampling pass
def rsample_differentiable(post, z: Tensor) -> Tuple[Tensor, Tensor]:
    """Public name of the op-by-op (any-order differentiable) sampling pass."""
    return _rsample_ops(post, z)


def _rsample_ops(post, z: Tensor) -> Tuple[Tensor, Tensor]:
    """z [S, Bf, D] -> (theta [S, Bf, D], log q(theta) [S, Bf]); one pass-equivalent per transform."""
    model = post.model
    st = _struct(model)
    S, Bf, D = z.shape
    C = st.C
    h = _context(post)
    M = S * Bf
    x = z.reshape(M, D)
    logq = (-0.5 * x * x - HALF_LOG_2PI).sum(-1)
    lo, hi = model.log_scale_min_clip, model.log_scale_max_clip
    for k in range(st.K):
        layers = _masked(model, k)
        nl = len(layers) - 1                                  # hidden layers
        W0, b0 = layers[0]
        A = linear(h, W0[:, :C], b0)
        A = A.unsqueeze(0).expand(S, Bf, A.shape[-1]).reshape(M, -1)
        Wo, bo = layers[-1]
        perm, grp = st.perm[k], st.groups[k]
        hs: List[List[Tensor]] = [[] for _ in range(nl)]      # per layer: activations of the degree groups so far
        ys: List[Tensor] = []                                  # y in visiting order
        ycols = [None] * D
        for j in range(D):
            d = perm[j]
            a0, a1 = grp[0][j], grp[0][j + 1]
            pre = A[:, a0:a1]
            if j > 0 and a1 > a0:
                Wt = W0[a0:a1, C:].index_select(1, st.perm_prefix[k][j])
                pre = pre + linear(torch.stack(ys, -1), Wt)
            hs[0].append(torch.relu(pre))
            for l in range(1, nl):
                W, b = layers[l]
                v0, v1 = grp[l][j], grp[l][j + 1]
                e = grp[l - 1][j + 1]
                hin = torch.cat(hs[l - 1], -1)
                hs[l].append(torch.relu(linear(hin, W[v0:v1, :e], b[v0:v1])))
            e = grp[nl - 1][j + 1]
            hin = torch.cat(hs[nl - 1], -1)
            rows = st.out_rows[k][j]
            ms = linear(hin, Wo.index_select(0, rows)[:, :e], bo.index_select(0, rows))
            sh = clamp_preserve_gradients(ms[:, 1], lo, hi)
            yj = (x[:, d] - ms[:, 0]) * torch.exp(-sh)
            ys.append(yj)
            ycols[d] = yj
            logq = logq + sh
        x = torch.stack(ycols, -1)
        if st.post is not None:
            x = x.index_select(-1, st.post[k])
    return x.reshape(S, Bf, D), logq.reshape(S, Bf)


# ----------------------------------------------------------------------------------------------- fused first-order path
_force_differentiable = 0


class differentiable_only:
    """Inside this block ``log_prob`` builds the op-by-op graph above even where the fused backward would serve: for code
    that is known to differentiate the result twice (the Fisher-trace regulariser)."""

    def __enter__(self):
        global _force_differentiable
        _force_differentiable += 1

    def __exit__(self, *exc):
        global _force_differentiable
        _force_differentiable -= 1
        return False


def _params(model) -> List[Tensor]:
    out = []
    for k in range(model.num_transforms):
        for layer in getattr(model, "t" + str(k)).nn.layers:
            out += [layer.weight, layer.bias]
    return out


class _FusedLogProb(torch.autograd.Function):
    """log q(v | ctx) whose backward is ``rabi_maf_log_prob_bwd``: one per-pair tensor-core launch (density passes + their
    reverse) and one launch for all weight-gradient products; no graph, no per-layer launches.  When the backward itself must
    be differentiable (``create_graph=True``) it hands over to the op-by-op graph of ``log_prob_differentiable``."""

    @staticmethod
    def forward(ctx, post, v, ctx_in, *params):
        ctx.post = post
        ctx.save_for_backward(v, ctx_in)
        return post.model.engine().log_prob(post.A, v)

    @staticmethod
    def backward(ctx, g):
        post = ctx.post
        v, ctx_in = ctx.saved_tensors
        need = ctx.needs_input_grad
        params = _params(post.model)
        if torch.is_grad_enabled():   # double backward requested: rebuild the differentiable graph and differentiate that
            with torch.enable_grad():
                vv = v if v.requires_grad else v.detach().requires_grad_(need[1])
                lp = _log_prob_graph(post, vv)
                wanted = [t for t, n in zip([vv, post.ctx_in] + params, need[1:]) if n]
                got = iter(torch.autograd.grad(lp, wanted, g, create_graph=True, allow_unused=True))
            return (None,) + tuple(next(got) if n else None for n in need[1:])
        eng = post.model.engine()
        A = post.A
        rows = _context(post).detach()
        want_w = any(need[3:])
        g_v, g_A, grads = eng.log_prob_bwd(rows, A, v.detach(), g.detach(), want_weights=want_w)
        g_ctx = None
        if need[2]:
            g_ctx = eng.ctx_project_bwd(g_A, post.zscore[1] if post.zscore is not None else None)
        flat = [None] * len(params)
        if want_w:
            flat = [t for gk in grads for wb in gk for t in wb]
            flat = [t if n else None for t, n in zip(flat, need[3:])]
        return (None, g_v if need[1] else None, g_ctx) + tuple(flat)


class _FusedRSample(torch.autograd.Function):
    """theta, log q(theta) = rsample(z) whose backward is ``rabi_maf_rsample_bwd`` (sampling chain + its reverse on the tensor
    cores, one launch for the weight-gradient products); a differentiable backward hands over to ``rsample_differentiable``."""

    @staticmethod
    def forward(ctx, post, z, ctx_in, *params):
        ctx.post = post
        ctx.save_for_backward(z, ctx_in)
        theta, logq = post.model.engine().rsample(post.A, z, want_logq=True)
        return theta, logq

    @staticmethod
    def backward(ctx, g_theta, g_logq):
        post = ctx.post
        z, ctx_in = ctx.saved_tensors
        need = ctx.needs_input_grad
        params = _params(post.model)
        if g_theta is None:
            g_theta = torch.zeros_like(z)
        if g_logq is None:
            g_logq = torch.zeros(z.shape[:-1], device=z.device, dtype=z.dtype)
        if torch.is_grad_enabled():
            with torch.enable_grad():
                theta, logq = _rsample_ops(post, z)
                wanted = [t for t, n in zip([z, post.ctx_in] + params, need[1:]) if n]
                got = iter(torch.autograd.grad([theta, logq], wanted, [g_theta, g_logq], create_graph=True, allow_unused=True))
            return (None,) + tuple(next(got) if n else None for n in need[1:])
        if need[1]:
            raise NotImplementedError("gradient w.r.t. the base noise is not provided by the fused sampling backward")
        eng = post.model.engine()
        want_w = any(need[3:])
        g_A, grads = eng.rsample_bwd(_context(post).detach(), post.A, z.detach(), g_theta.detach(), g_logq.detach(),
                                     want_weights=want_w)
        g_ctx = eng.ctx_project_bwd(g_A, post.zscore[1] if post.zscore is not None else None) if need[2] else None
        flat = [None] * len(params)
        if want_w:
            flat = [t for gk in grads for wb in gk for t in wb]
            flat = [t if n else None for t, n in zip(flat, need[3:])]
        return (None, None, g_ctx) + tuple(flat)


def rsample_training(post, z: Tensor) -> Tuple[Tensor, Tensor]:
    """Entry point of ``MAFPosterior.rsample`` when gradients are needed."""
    eng = post.model.engine()
    if _force_differentiable or not eng.fused_training_ok() or z.numel() == 0 or z.requires_grad:
        return _rsample_ops(post, z)
    return _FusedRSample.apply(post, z, post.ctx_in, *_params(post.model))


def _log_prob_graph(post, v: Tensor) -> Tensor:
    return _log_prob_ops(post, v)


def log_prob_training(post, v: Tensor) -> Tensor:
    """Entry point of ``MAFPosterior.log_prob`` when gradients are needed: the fused kernels where they serve (two-hidden-layer
    conditioners without Permute layers), else the op-by-op graph."""
    eng = post.model.engine()
    if _force_differentiable or not eng.fused_training_ok() or v.numel() == 0:
        return _log_prob_ops(post, v)
    return _FusedLogProb.apply(post, v, post.ctx_in, *_params(post.model))
